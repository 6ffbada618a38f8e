#!/bin/bash
# FFN pair epilogue (gate|up GEMV writes silu(gate)*up): parity + decode-step A/B + timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llama_step.py tests/test_gpu_tp.py tests/test_gpu_reference_parity.py -q -m gpu -x 2>&1 | tail -3
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-batch --steps 128 --warmup 8 > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2l_$name.json").read().strip().splitlines()[-1])
    print("$name", "tok/s %.1f" % d["value"], "ms %.4f" % d["ms_per_step"], "gemv frac %.4f" % d["roofline"]["frac"], "avg_us %.3f" % d["roofline"]["avg_launch_us"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2l_$name.err").read()[-1500:])
PY
}
run pair0 GGML_B200_BS1_PAIR=0
run pair1 GGML_B200_BS1_PAIR=1
run pair0b GGML_B200_BS1_PAIR=0
run pair1b GGML_B200_BS1_PAIR=1
timeout 300 python tools/step_prof.py 3 2>&1 | tail -5
