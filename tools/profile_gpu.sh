#!/bin/bash
# Run on the GPU box (under gpurun): ncu --set full captures of the hot kernels of one pass.
#   tools/profile_gpu.sh <tag>  -> gpurun_out/<tag>_{mvs,scene,flow}.ncu-rep
# tools/prof_step.py brackets one region of its last pass with cudaProfilerStart/Stop.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
for region in mvs flow scene; do
  timeout 240 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/${TAG}_${region} \
      python tools/prof_step.py --region $region > gpurun_out/${TAG}_${region}.log 2>&1
  echo "$region rc=$?"
done
ls -la gpurun_out/ | head -30
