#!/bin/bash
# On the GPU box (under gpurun): alternate the two libraries of tools/ab_build.sh under bench.py.
#     tools/ab_run.sh [rounds=2] [steps=40]
set -u
ROUNDS=${1:-2}
STEPS=${2:-40}
cp 3dvnet_b200/lib3dvnet_b200.so /tmp/lib_keep.so
for r in $(seq 1 "$ROUNDS"); do
  for v in old new; do
    cp ab_$v.so 3dvnet_b200/lib3dvnet_b200.so
    timeout 200 python bench.py --steps "$STEPS" --warmup 5 --streams 0 2>&1 | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('$v', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'gemm0_us', round(1e3 * d['roofline']['kernel_ms'], 1),
      {k: round(x, 3) for k, x in d['stages_ms_per_step'].items()}, 'abs_rel', d['abs_rel_vs_oracle'])"
  done
done
cp /tmp/lib_keep.so 3dvnet_b200/lib3dvnet_b200.so
