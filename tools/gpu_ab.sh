#!/bin/bash
# Decode-step A/B on one B200 (bench.py --no-cpu --no-batch: Llama-3-8B Q4_K_M bs1, depth 512): one line per variant, then the in-situ
# GEMV timeline (tools/step_prof.py) of the LAST variant.  Usage (through gpurun):
#   bash tools/gpu_ab.sh base: win0:GGML_B200_BS1_WINDOW=0 nopair:GGML_B200_BS1_PAIR=0 cl4:GGML_B200_BS1_PAIR=0,GGML_B200_BS1_CLUSTER_F32=1
# This is how the round-2 experiments of DESIGN.md 8 were measured: L2 look-ahead (GGML_B200_L2_PREFETCH=1|2, GGML_B200_L2PF_MB), first-fill
# window (GGML_B200_BS1_WINDOW=KB), cluster size of the activation prologue (GGML_B200_BS1_CLUSTER=1|2|4), FFN pair epilogue (GGML_B200_BS1_PAIR).
mkdir -p gpurun_out
last=""
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}; last=$(echo "$envs" | tr ',' ' ')
  env $last timeout 600 python bench.py --no-cpu --no-batch --steps 128 --warmup 8 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % name).read().strip().splitlines()[-1])
    print(name, "tok/s %.1f" % d["value"], "ms %.4f" % d["ms_per_step"], "gemv frac %.4f" % d["roofline"]["frac"], "avg_us %.3f" % d["roofline"]["avg_launch_us"])
except Exception as e:
    print(name, "FAILED", e); print(open("gpurun_out/ab_%s.err" % name).read()[-1500:])
PY
done
env $last timeout 300 python tools/step_prof.py 3 2>&1 | tail -5
