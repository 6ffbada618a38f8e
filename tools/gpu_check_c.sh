#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== repro prefill (plain)"; timeout 300 python tools/repro_prefill.py 512 2>&1 | tail -3
echo "== repro prefill 16 (sanitizer)"; timeout 600 compute-sanitizer --tool memcheck python tools/repro_prefill.py 16 > $O/sanitizer_prefill.log 2>&1; grep -E "Invalid|Error|at 0x|by thread|Address|=========     in|FAILED|ok " $O/sanitizer_prefill.log | head -30
echo "== q4_0 gemv (sanitizer)"; timeout 600 compute-sanitizer --tool memcheck python tools/bench_gemv.py --types q4_0 --cols 4 --shapes 14336x4096 --iters 1 > $O/sanitizer_q40.log 2>&1; grep -E "Invalid|Error|at 0x|by thread|Address|=========     in" $O/sanitizer_q40.log | head -30
echo "== nodes q4_k_m q8_0 (fusion off)"; GGML_B200_FUSION=0 timeout 300 bash tools/_exp.sh q4_k_m q8_0 90 2>&1 | cut -c1-200
