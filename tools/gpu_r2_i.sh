#!/bin/bash
mkdir -p gpurun_out
GGML_B200_DSTEP=0 python bench.py --no-cpu --no-batch --steps 64 --warmup 8 > gpurun_out/r2i_bench_dstep0.json 2> gpurun_out/r2i_err0.log
GGML_B200_DSTEP=1 python bench.py --no-cpu --no-batch --steps 64 --warmup 8 > gpurun_out/r2i_bench_dstep1.json 2> gpurun_out/r2i_err1.log
python - <<'PY'
import json
for f in ("gpurun_out/r2i_bench_dstep0.json","gpurun_out/r2i_bench_dstep1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f tok/s ms/step %.3f e2e %.1f launches %s roofline frac %.3f (%.2f us/launch) whole-step frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["whole_step"]["frac"]))
    except Exception as ex:
        print(f, "failed", ex); print(open(f.replace("bench_dstep","err").replace(".json",".log")).read()[-1500:])
PY
