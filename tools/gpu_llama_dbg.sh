#!/bin/bash
cd "$(dirname "$0")/.."
HERE=$(pwd)
export LD_LIBRARY_PATH="$HERE/cortex.llamacpp_b200:$HERE/oracle/_ref:$LD_LIBRARY_PATH"
G=/tmp/e2e_mid-d128_q4_k_m.gguf
[ -f "$G" ] || python tools/make_gguf.py --model mid-d128 --ftype q4_k_m --out $G 2>&1 | tail -1
D=oracle/_ref/logits_dump
echo "== graph debug (mid model, 16 prompt + 12 gen)"
GGML_B200_GRAPH_DEBUG=1 GGML_BACKEND_PATH=$HERE/cortex.llamacpp_b200/libggml-b200.so $D $G /tmp/x.bin 99 16 12 f16 1 4 2>&1 | grep "b200 graph" | head -12
$D $G /tmp/lg_cpu.bin 0 64 48 f16 1 16 > /dev/null 2>&1
for v in "A=1" "GGML_B200_GRAPHS=0" "GGML_B200_BS1_OFF=1 GGML_B200_GRAPHS=0" "GGML_B200_BS1_OFF=1 GGML_B200_GRAPHS=0 GGML_B200_FUSION=0"; do
  echo "== $v"; env $v GGML_BACKEND_PATH=$HERE/cortex.llamacpp_b200/libggml-b200.so $D $G /tmp/lg_gpu.bin 99 64 48 f16 1 4 > /dev/null 2>&1; python tools/compare_logits.py /tmp/lg_cpu.bin /tmp/lg_gpu.bin | cut -c1-260
done
