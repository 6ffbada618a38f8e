#!/bin/bash
# fattn_tc.cu: parity, pp512 and depth-4096 prefill, one ncu --set full capture of the kernel (summary: profiles/r2_fattn_tc.md)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fattn.py tests/test_gpu_llama_step.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python tools/prefill_prof.py 512 2 prefill 2>&1 | tail -1
timeout 600 python tools/bench_configs.py llama3-8b:q4_k_m --depth 4096 --steps 16 > gpurun_out/fa_depth4096.jsonl 2> gpurun_out/fa_depth4096.err; cut -c1-700 gpurun_out/fa_depth4096.jsonl
timeout 600 python bench.py --no-cpu --steps 128 --warmup 8 > gpurun_out/fa_bench.json 2> gpurun_out/fa_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/fa_bench.json").read().strip().splitlines()[-1])
print("bs1 tok/s %.1f frac %.4f" % (d["value"], d["roofline"]["frac"]), "bs32", d["batched"]["bs32_decode"]["value"], "pp512", d["batched"]["prefill_pp512"]["value"], d["batched"]["prefill_pp512"]["ms_per_ubatch"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b200_fattn_tc -s 4 -c 1 -f -o gpurun_out/fa_fattn_tc python tools/prefill_prof.py 512 2 prefill > gpurun_out/fa_ncu.log 2>&1
ncu -i gpurun_out/fa_fattn_tc.ncu-rep --page raw --csv > gpurun_out/fa_fattn_tc_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/fa_fattn_tc_raw.csv'))); d={h:rows[2][i] for i,h in enumerate(rows[0])}
for k in ['gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','launch__registers_per_thread','launch__shared_mem_per_block_dynamic']: print(k, d.get(k))
PY
