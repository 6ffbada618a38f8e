#!/bin/bash
# First-contact GPU run: smoke, parity tests, GEMV microbench, bench line, reference op harness, e2e parity, ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -3 $O/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_gpu.log
echo "== bench_gemv"; timeout 300 python tools/bench_gemv.py --types q4_K,q6_K --cols 1 --shapes 4096x4096,1024x4096,14336x4096,4096x14336,128256x4096 > $O/bench_gemv.log 2>&1; echo "rc=$?"; cat $O/bench_gemv.log
echo "== bench_gemv other"; timeout 300 python tools/bench_gemv.py --types q5_K,q4_0,q8_0 --cols 1,4 --shapes 4096x4096,14336x4096 > $O/bench_gemv2.log 2>&1; echo "rc=$?"; cat $O/bench_gemv2.log
echo "== bench.py"; timeout 900 python bench.py --steps 32 --warmup 4 > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
echo "== bench.py no graphs"; timeout 300 python bench.py --steps 16 --warmup 3 --graphs 0 --no-cpu > $O/bench_nograph.json 2>> $O/bench.err; echo "rc=$?"; cat $O/bench_nograph.json
echo "== backend ops"; timeout 900 bash tools/run_backend_ops.sh > $O/backend_ops.log 2>&1; echo "rc=$?"; cat $O/backend_ops.log | tail -40
echo "== e2e parity"; for cfg in "tiny-d128 q4_k_m" "tiny-d128 q4_0" "tiny-d64 q4_0"; do timeout 300 bash tools/e2e_parity.sh $cfg 2>&1 | tail -4; done > $O/e2e_parity.log 2>&1; cat $O/e2e_parity.log
timeout 300 bash tools/e2e_parity.sh tiny-d128 q4_k_m -ctk q8_0 -ctv q8_0 2>&1 | tail -3 >> $O/e2e_parity.log
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_r1a.csv python bench.py --steps 2 --warmup 3 --graphs 0 --no-cpu > $O/ncu_bench.log 2>&1; echo "rc=$?"; wc -l $O/launches_r1a.csv
