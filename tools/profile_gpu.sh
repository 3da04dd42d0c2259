#!/bin/bash
# Run on the GPU box (under gpurun): ncu --set full captures of the hot kernels of one bench step.
#   tools/profile_gpu.sh <tag>       -> gpurun_out/<tag>_gemm.ncu-rep, gpurun_out/<tag>_path_a.ncu-rep
# Steps 0..2 are warm-up (weights packed, caches built); the captured launches belong to step 3.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
GEMM_PER_STEP=${GEMM_PER_STEP:-60}
A_PER_STEP=${A_PER_STEP:-38}
ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc \
    -s $((3 * GEMM_PER_STEP)) -c $GEMM_PER_STEP -f -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_gemm.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k 'regex:planesweep_var|conv3d|prob_softargmin|points_var|sparse_interp' \
    -s $((3 * A_PER_STEP)) -c $A_PER_STEP -f -o gpurun_out/${TAG}_path_a $CMD > gpurun_out/${TAG}_path_a.log 2>&1
ls -la gpurun_out/
