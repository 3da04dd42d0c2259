#!/bin/bash
# Run on the GPU box (under gpurun): ncu --set full captures of the hot kernels of one pass.
#   tools/profile_gpu.sh <tag>  -> gpurun_out/<tag>_{mvs,scene,flow}.ncu-rep (+ raw-page CSVs)
# tools/prof_step.py brackets one region of its last pass with cudaProfilerStart/Stop.
# A --set full record is ~2 MB per launch and gpurun_out/ is capped at 64 MiB: keep the counts small.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
cap() {  # region, kernel regex, launch-skip, launch-count
  timeout 300 ncu --set full --clock-control none --profile-from-start off -k "regex:$2" -s $3 -c $4 -f \
      -o gpurun_out/${TAG}_$1 python tools/prof_step.py --region $1 > gpurun_out/${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
}
cap mvs 'planesweep_var|s1_tiled|prob_softargmin|conv3d_direct|deconv3d' 0 6
cap flow 'points_var|sparse_interp|gather_gemm|decoder_head' 0 8
cap scene 'gather_gemm|pair_reduce|pair_fill|kernel_map' 6 8
du -sh gpurun_out; ls -la gpurun_out/ | head -30
