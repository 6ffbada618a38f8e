#!/bin/bash
# round-2 re-entry: GPU suite + bs1 L2 look-ahead / cluster-size A/B on the decode step (bench.py --no-cpu --no-batch)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-batch --steps 128 --warmup 8 > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2h_$name.json").read().strip().splitlines()[-1])
    print("$name", "tok/s %.1f" % d["value"], "ms %.4f" % d["ms_per_step"], "gemv frac %.4f" % d["roofline"]["frac"], "avg_us %.3f" % d["roofline"]["avg_launch_us"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2h_$name.err").read()[-1500:])
PY
}
run base A=1
run l2pf2 GGML_B200_L2_PREFETCH=2
run l2pf2_48 GGML_B200_L2_PREFETCH=2 GGML_B200_L2PF_MB=48
run l2pf1 GGML_B200_L2_PREFETCH=1
run cl2 GGML_B200_BS1_CLUSTER=2
run cl2_l2pf2 GGML_B200_BS1_CLUSTER=2 GGML_B200_L2_PREFETCH=2
