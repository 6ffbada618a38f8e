#!/bin/bash
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows8 -s 2 -c 1 -o gpurun_out/r2g_rows8 -f python tools/batched_prof.py bs32 2 1 > /dev/null 2>&1
ls -la gpurun_out/r2g_rows8.ncu-rep
