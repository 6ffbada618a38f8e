#!/bin/bash
# down projection prologue: cluster size A/B with the FFN pair epilogue; gemm_tc timeline at 32 token columns
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-batch --steps 128 --warmup 8 > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2m_$name.json").read().strip().splitlines()[-1])
    print("$name", "tok/s %.1f" % d["value"], "ms %.4f" % d["ms_per_step"], "gemv frac %.4f" % d["roofline"]["frac"], "avg_us %.3f" % d["roofline"]["avg_launch_us"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2m_$name.err").read()[-1500:])
PY
}
run cl4 GGML_B200_BS1_CLUSTER=4
run cl2 GGML_B200_BS1_CLUSTER=2
run cl1 GGML_B200_BS1_CLUSTER=1
run cl4_nopair GGML_B200_BS1_CLUSTER=4 GGML_B200_BS1_PAIR=0
run cl1_nopair GGML_B200_BS1_CLUSTER=1 GGML_B200_BS1_PAIR=0
GGML_B200_BS1_CLUSTER=1 timeout 300 python tools/step_prof.py 3 2>&1 | grep "down"
GGML_B200_TC_PROF=1 GGML_B200_TC_MIN_M=0 timeout 300 python tools/bench_gemv.py --types q4_K --cols 32 --shapes 14336x4096 --iters 3 > gpurun_out/r2m_tc32.log 2>&1; head -60 gpurun_out/r2m_tc32.log | cut -c1-260
