#!/bin/bash
# Greedy-decode parity: the same synthetic GGUF through the reference CPU backend (-ngl 0) and through the B200 backend
# (-ngl 99), identical prompt, --temp 0.  Prints both token streams and whether they are identical.
# usage: tools/e2e_parity.sh MODEL FTYPE [extra llama-cli args...]   e.g. tiny-d128 q4_k_m -ctk q8_0 -ctv q8_0
HERE="$(cd "$(dirname "$0")/.." && pwd)"
MODEL=$1; FTYPE=$2; shift 2
NTOK=${NTOK:-64}
G=/tmp/e2e_${MODEL}_${FTYPE}.gguf
[ -f "$G" ] || python "$HERE/tools/make_gguf.py" --model "$MODEL" --ftype "$FTYPE" --out "$G" ${LAYERS:+--layers $LAYERS} 2>&1 | tail -1
export LD_LIBRARY_PATH="$HERE/cortex.llamacpp_b200:$HERE/oracle/_ref:$LD_LIBRARY_PATH"
CLI="$HERE/oracle/_ref/llama-cli"
PROMPT="${PROMPT:-the quick brown fox jumps over the lazy dog and keeps running through the forest}"
COMMON=(-m "$G" -p "$PROMPT" -n "$NTOK" --temp 0 --top-k 1 -no-cnv -fa -s 1 --no-warmup "$@")
"$CLI" "${COMMON[@]}" -ngl 0 -t "${THREADS:-8}" > /tmp/e2e_cpu.txt 2> /tmp/e2e_cpu.err
GGML_BACKEND_PATH="$HERE/cortex.llamacpp_b200/libggml-b200.so" "$CLI" "${COMMON[@]}" -ngl 99 -t 4 > /tmp/e2e_gpu.txt 2> /tmp/e2e_gpu.err
echo "--- cpu:"; cat /tmp/e2e_cpu.txt | head -c 600; echo
echo "--- b200:"; cat /tmp/e2e_gpu.txt | head -c 600; echo
grep -E "offloaded|B200|eval time" /tmp/e2e_gpu.err | head -8
grep -E "eval time" /tmp/e2e_cpu.err | head -3
if cmp -s /tmp/e2e_cpu.txt /tmp/e2e_gpu.txt; then echo "PARITY: IDENTICAL ($MODEL $FTYPE $*)"; else echo "PARITY: DIFFERENT ($MODEL $FTYPE $*)"; exit 1; fi
