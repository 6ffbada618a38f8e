#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== pytest gpu"; timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-300
echo "== ncu full gemv"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:b200_gemv -s 40 -c 6 -f -o $O/prof_gemv_r1 python bench.py --steps 1 --warmup 3 --graphs 0 --no-cpu --pdl 0 > $O/ncu_full.log 2>&1; echo "rc=$?"; ls -la $O/*.ncu-rep
