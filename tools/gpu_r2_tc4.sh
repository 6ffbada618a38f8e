#!/bin/bash
# evidence for gemm_tc: one --set full capture of the kernel, and the launch list of a 2-layer pp512 ubatch
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:b200_gemm_tc_kernel -c 1 -o gpurun_out/r3e_tc python tools/bench_gemv.py --types q4_K --cols 512 --shapes 4096x4096 --iters 2 > gpurun_out/r3e_ncu.log 2>&1; tail -2 gpurun_out/r3e_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r3e_pp512.csv python tools/prefill_prof.py 512 2 prefill > gpurun_out/r3e_pp512.log 2>&1; tail -1 gpurun_out/r3e_pp512.log
python tools/summarize_launches.py gpurun_out/r3e_pp512.csv > gpurun_out/r3e_pp512_summary.txt 2>&1; head -16 gpurun_out/r3e_pp512_summary.txt
