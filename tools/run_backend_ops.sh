#!/bin/bash
# Runs the reference's own op-parity harness (llama.cpp tests/test-backend-ops.cpp, built unmodified into oracle/_ref)
# against the B200 backend loaded through GGML_BACKEND_PATH.  Prints per-op pass/fail/unsupported counts.
# usage: tools/run_backend_ops.sh [OP ...]
HERE="$(cd "$(dirname "$0")/.." && pwd)"
export GGML_BACKEND_PATH="$HERE/cortex.llamacpp_b200/libggml-b200.so"
export LD_LIBRARY_PATH="$HERE/cortex.llamacpp_b200:$HERE/oracle/_ref:$LD_LIBRARY_PATH"
OPS=("$@")
[ ${#OPS[@]} -eq 0 ] && OPS=(MUL_MAT MUL_MAT_ID FLASH_ATTN_EXT RMS_NORM ROPE CPY CONT DUP ADD MUL DIV SILU SOFT_MAX GET_ROWS ARGSORT SUM_ROWS SCALE)
rc=0
for op in "${OPS[@]}"; do
  out=$("$HERE/oracle/_ref/test-backend-ops" test -b B2000 -o "$op" 2>&1)
  ok=$(echo "$out" | grep -c "OK$\|\[1;32mOK")
  fail=$(echo "$out" | grep -c "FAIL")
  ns=$(echo "$out" | grep -c "not supported")
  printf "%-16s ok=%-5s fail=%-4s not_supported=%-5s\n" "$op" "$ok" "$fail" "$ns"
  if [ "$fail" != "0" ]; then rc=1; echo "$out" | grep "FAIL" | head -${FAILS_SHOWN:-6}; fi
  if [ -n "$SHOW_UNSUPPORTED" ]; then echo "$out" | grep "not supported" | head -${SHOW_UNSUPPORTED}; fi
done
exit $rc
