#!/bin/bash
# bs1 GEMV first-fill window (GGML_B200_BS1_WINDOW, KB per SM): correctness subset, in-situ timeline, decode-step A/B
mkdir -p gpurun_out
GGML_B200_BS1_WINDOW=64 timeout 900 python -m pytest tests/test_gpu_mulmat.py tests/test_gpu_llama_step.py -q -m gpu -x 2>&1 | tail -3
for w in 0 64; do echo "== timeline window $w"; GGML_B200_BS1_WINDOW=$w timeout 300 python tools/step_prof.py 3 2>&1 | tail -5; done
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-batch --steps 128 --warmup 8 > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2j_$name.json").read().strip().splitlines()[-1])
    print("$name", "tok/s %.1f" % d["value"], "ms %.4f" % d["ms_per_step"], "gemv frac %.4f" % d["roofline"]["frac"], "avg_us %.3f" % d["roofline"]["avg_launch_us"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2j_$name.err").read()[-1500:])
PY
}
run w0 GGML_B200_BS1_WINDOW=0
run w32 GGML_B200_BS1_WINDOW=32
run w48 GGML_B200_BS1_WINDOW=48
run w64 GGML_B200_BS1_WINDOW=64
run w96 GGML_B200_BS1_WINDOW=96
run w128 GGML_B200_BS1_WINDOW=128
run w64_l2pf2 GGML_B200_BS1_WINDOW=64 GGML_B200_L2_PREFETCH=2
