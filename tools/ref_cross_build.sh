#!/bin/bash
# How reproducible is the REFERENCE against itself?  Builds a second copy of the unmodified reference CPU backend with a
# different SIMD level (SSE4.2 instead of AVX2+FMA: other vec_dot branches, other float summation order, no FMA) next to
# oracle/_ref and compares the logits of the two builds on the parity fixtures (same GGUF, same tokens, same host).
# Needs /root/reference (run in the build container).  Result: profiles/r2_reference_cross_build.md
HERE="$(cd "$(dirname "$0")/.." && pwd)"
OUT=${OUT:-/tmp/ref_sse42}
make -C "$HERE/oracle" -j"$(nproc)" ref OUT="$OUT" OBJ="$OUT/obj" ARCHFL="-msse4.2" > "$OUT.build.log" 2>&1 || { tail "$OUT.build.log"; exit 1; }
run() { # model ftype layers kv seed
  G=/tmp/xb_$1_$2_L$3.gguf
  [ -f "$G" ] || python "$HERE/tools/make_gguf.py" --model "$1" --ftype "$2" ${3:+--layers $3} --out "$G" 2>/dev/null
  LD_LIBRARY_PATH="$HERE/oracle/_ref" LOGITS_DUMP_SEED=$5 "$HERE/oracle/_ref/logits_dump" "$G" /tmp/xb_a.bin 0 32 128 "$4" 1 "$(nproc)" > /dev/null 2>&1
  LD_LIBRARY_PATH="$OUT" LOGITS_DUMP_SEED=$5 "$OUT/logits_dump" "$G" /tmp/xb_b.bin 0 32 128 "$4" 1 "$(nproc)" > /dev/null 2>&1
  echo "$1 $2 layers=${3:-all} kv=$4: $(python "$HERE/tools/compare_logits.py" /tmp/xb_a.bin /tmp/xb_b.bin)"
}
run tinyllama q4_0 "" f16 9
run llama3-8b q4_k_m 2 q8_0 10
