#!/bin/bash
# Re-entry evidence run: smoke, GPU parity suite, bench line, ncu launch list + one full capture of the GEMV kernel.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1; nproc > $O/nproc.txt
stamp "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -3 $O/smoke.log | cut -c1-300
stamp "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -16 $O/pytest_gpu.log | cut -c1-300
stamp "== bench.py"; timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
stamp "== bench.py --impl reference"; timeout 400 python bench.py --impl reference --steps 8 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err; echo "rc=$?"; cat $O/bench_ref.json
stamp "== gemv micro"; timeout 200 python tools/bench_gemv.py --types q4_K,q6_K --cols 1 --shapes 4096x4096,14336x4096,4096x14336,128256x4096 2>&1 | tail -10
stamp "== ncu launches"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:b200_ -c 600 --csv --log-file $O/launches_r1i.csv python bench.py --steps 1 --warmup 3 --graphs 0 --no-cpu > $O/ncu_bench.log 2>&1; echo "rc=$?"; python tools/summarize_launches.py $O/launches_r1i.csv | tee $O/launches_r1i_summary.txt
stamp "== ncu full gemv"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:b200_gemv -s 40 -c 4 -f -o $O/prof_gemv_r1i python bench.py --steps 1 --warmup 3 --graphs 0 --no-cpu > $O/ncu_full.log 2>&1; echo "rc=$?"; ls -la $O/*.ncu-rep
stamp "done"
