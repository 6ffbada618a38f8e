#!/bin/bash
# GEMV diagnosis: in-kernel timeline stamps, micro-bench GB/s, one ncu --set full capture with source per shape.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp "== prof stamps"
for s in "4096 4096" "28672 4096" "4096 14336"; do echo "-- N K = $s"; timeout 120 python tools/gemv_prof.py $s 0 2>&1 | tail -9; done
stamp "== gemv micro"; timeout 300 python tools/bench_gemv.py --types q4_K,q6_K --cols 1 --shapes 4096x4096,28672x4096,4096x14336,128256x4096 2>&1 | tail -10
stamp "== ncu full"; 
timeout 300 ncu --set full --clock-control none --import-source on -k regex:b200_gemv -s 3 -c 1 -f -o $O/prof_gemv_big python tools/gemv_prof.py 28672 4096 0 > $O/ncu_big.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:b200_gemv -s 3 -c 1 -f -o $O/prof_gemv_small python tools/gemv_prof.py 4096 4096 0 > $O/ncu_small.log 2>&1; echo "rc=$?"
ls -la $O/*.ncu-rep
stamp done
