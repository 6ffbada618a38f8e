#!/bin/bash
# usage: tools/logits_parity.sh MODEL FTYPE KV N_PROMPT N_GEN [N_PARALLEL]
HERE="$(cd "$(dirname "$0")/.." && pwd)"
MODEL=$1; FTYPE=$2; KV=${3:-f16}; NP=${4:-64}; NG=${5:-32}; PAR=${6:-1}
G=/tmp/e2e_${MODEL}_${FTYPE}${LAYERS:+_L$LAYERS}.gguf
[ -f "$G" ] || python "$HERE/tools/make_gguf.py" --model "$MODEL" --ftype "$FTYPE" --out "$G" ${LAYERS:+--layers $LAYERS} 2>&1 | tail -1
export LD_LIBRARY_PATH="$HERE/cortex.llamacpp_b200:$HERE/oracle/_ref:$LD_LIBRARY_PATH"
D="$HERE/oracle/_ref/logits_dump"
"$D" "$G" /tmp/lg_cpu.bin 0 $NP $NG $KV $PAR ${THREADS:-$(nproc)} 2>/tmp/lg_cpu.err | tail -1
GGML_BACKEND_PATH="$HERE/cortex.llamacpp_b200/libggml-b200.so" "$D" "$G" /tmp/lg_gpu.bin 99 $NP $NG $KV $PAR 4 2>/tmp/lg_gpu.err | tail -1
python "$HERE/tools/compare_logits.py" /tmp/lg_cpu.bin /tmp/lg_gpu.bin
