#!/bin/bash
# Standard A/B check on the GPU box: a subset of the GPU tests, the in-kernel phase timings and one bench line.
#   gpurun -- 'bash tools/gpu_check.sh TAG [pytest args...]'   -> gpurun_out/TAG_bench.json
TAG=${1:-check}; shift
TESTS=${@:-tests/test_gpu_gemm_tc.py tests/test_gpu_engine.py tests/test_gpu_decoder_fused.py tests/test_gpu_parity.py}
timeout 900 python -m pytest $TESTS -q -x 2>&1 | tail -4
python tools/gemm_phases.py 2>&1 | grep -A10 "K=128 N=128\|density 0.100"
python tools/decoder_phases.py 2>&1 | tail -9
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --streams 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
print({k: round(v, 3) for k, v in d["stages_ms_per_step"].items()})
print("roofline", round(d["roofline"]["frac"], 4), d["roofline"]["kernel_ms"], "warp", round(d["roofline_warp"]["frac"], 4))
PY
