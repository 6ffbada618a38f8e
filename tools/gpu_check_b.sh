#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_gpu.log
echo "== pdl gemv"; timeout 300 python tools/bench_gemv.py --types q4_K --cols 1 --pdl 1 --shapes 4096x4096,14336x4096,4096x14336 2>&1 | tail -4
echo "== bench pdl=1"; GGML_B200_PDL=1 timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-400
echo "== logits parity"
for cfg in "tiny-d128 q4_0 f16" "tiny-d64 q4_0 f16" "tiny-d128 q4_k_m q8_0" "tiny-d128 q4_k_m q4_0" "tiny-d128 q5_k_m f16" "tiny-d128 q8_0 f16"; do echo "-- $cfg"; timeout 300 bash tools/logits_parity.sh $cfg 64 32 1 2>&1 | tail -3 | cut -c1-700; done
echo "-- multi-slot"; timeout 300 bash tools/logits_parity.sh tiny-d128 q4_k_m q8_0 32 16 4 2>&1 | tail -3 | cut -c1-700
echo "== harness 8B"; python tools/make_gguf.py --model llama3-8b --ftype q4_k_m --out /tmp/l3.gguf 2>&1 | tail -1
export LD_LIBRARY_PATH=$PWD/cortex.llamacpp_b200:$PWD/oracle/_ref
GGML_BACKEND_PATH=$PWD/cortex.llamacpp_b200/libggml-b200.so GGML_B200_GRAPHS=1 LOGITS_DUMP_WARMUP=4 timeout 600 oracle/_ref/logits_dump /tmp/l3.gguf - 99 512 32 f16 1 4 > $O/harness_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/harness_gpu.log | cut -c1-300
echo "== ncu launches (own kernels only)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:b200_ -c 1500 --csv --log-file $O/launches_r1b.csv python bench.py --steps 1 --warmup 3 --graphs 0 --no-cpu > $O/ncu_bench.log 2>&1; echo "rc=$?"; wc -l $O/launches_r1b.csv
