#!/bin/bash
# Round evidence: smoke, GPU parity suite, bench line (both arms), ncu launch list + one full capture of the bs1 GEMV kernel.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1; nproc > $O/nproc.txt
stamp "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -3 $O/smoke.log | cut -c1-300
stamp "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -10 $O/pytest_gpu.log | cut -c1-300
stamp "== bench.py"; timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
stamp "== bench.py --impl reference"; timeout 400 python bench.py --impl reference --steps 8 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err; echo "rc=$?"; cat $O/bench_ref.json
stamp "== ncu launches"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:b200_ -c 700 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --graphs 0 --no-cpu > $O/ncu_bench.log 2>&1; echo "rc=$?"; python tools/summarize_launches.py $O/launches.csv | tee $O/launches_summary.txt
stamp "== ncu full bs1 gemv"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:b200_gemv_bs1 -s 40 -c 4 -f -o $O/prof_bs1 python bench.py --steps 1 --warmup 3 --graphs 0 --no-cpu > $O/ncu_full.log 2>&1; echo "rc=$?"; ls -la $O/*.ncu-rep
stamp "done"
