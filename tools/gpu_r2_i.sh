#!/bin/bash
# in-situ bs1 timeline + one ncu --set full capture of the tcgen05 flash-attention kernel
mkdir -p gpurun_out
timeout 300 python tools/step_prof.py 3 2>&1 | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b200_fattn_tc -s 4 -c 1 -f -o gpurun_out/r2i_fattn_tc python tools/prefill_prof.py 512 2 prefill > gpurun_out/r2i_ncu.log 2>&1
tail -3 gpurun_out/r2i_ncu.log
ncu -i gpurun_out/r2i_fattn_tc.ncu-rep --page raw --csv > gpurun_out/r2i_fattn_tc_raw.csv 2>/dev/null
ls -la gpurun_out/
