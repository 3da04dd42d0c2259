#!/bin/bash
# Same-box A/B of a kernel change (GPU boxes differ by a few percent, so before/after numbers from two gpurun
# calls are not comparable).  Run HERE, with the change in the working tree:
#     tools/ab_build.sh            # builds ab_old.so (HEAD) and ab_new.so (working tree)
#     gpurun --timeout 600 -- 'tools/ab_run.sh 2 40'
#     rm ab_old.so ab_new.so       # they are git-ignored, but travel with every gpurun snapshot
# Only csrc/ changes are A/B-able this way (both libraries must export the header's symbols).
set -eu
cd "$(dirname "$0")/.."
if git diff --quiet -- 3dvnet_b200/csrc; then
  echo "no uncommitted change under 3dvnet_b200/csrc: nothing to compare" >&2
  exit 1
fi
python 3dvnet_b200/build.py > /dev/null
cp 3dvnet_b200/lib3dvnet_b200.so ab_new.so
git stash -q
trap 'git stash pop -q' EXIT
python 3dvnet_b200/build.py > /dev/null
cp 3dvnet_b200/lib3dvnet_b200.so ab_old.so
trap - EXIT
git stash pop -q
python 3dvnet_b200/build.py > /dev/null
touch ab_old.so ab_new.so 3dvnet_b200/lib3dvnet_b200.so
ls -la ab_old.so ab_new.so
