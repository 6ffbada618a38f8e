#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:b200_fattn_kernel -s 2 -c 1 -o gpurun_out/r2s_fa -f python tools/batched_prof.py bs32 2 1 > gpurun_out/r2s_ncu.log 2>&1
tail -n 2 gpurun_out/r2s_ncu.log
