#!/bin/bash
mkdir -p gpurun_out
./tools/micro/imma_bench.bin 2>&1 | tee gpurun_out/r2r_imma.txt
timeout 900 python -m pytest tests/test_gpu_fattn.py -x -q -m gpu 2>&1 | tail -5
python tools/batched_prof.py bs32 32 8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2r_bs32.csv python tools/batched_prof.py bs32 4 1 > gpurun_out/r2r_bs32.log 2>&1
python tools/summarize_launches.py gpurun_out/r2r_bs32.csv 2>&1 | grep -v "at::" | head -12
