#!/bin/bash
# the unmodified llama.cpp runtime (oracle/_ref) with libggml-b200.so as its backend: logits + greedy parity vs the reference CPU
# backend (graphs + KV-table indirection on), then the decode throughput of the 8B bench model through llama_decode
cd "$(dirname "$0")/.."
echo "== logits parity tiny q4_k_m q8_0 KV (64 prompt + 48 generated, teacher forcing)"; bash tools/logits_parity.sh tiny-d128 q4_k_m q8_0 64 48 2>&1 | tail -3
echo "== logits parity mid q4_k_m f16 KV"; bash tools/logits_parity.sh mid-d128 q4_k_m f16 64 48 2>&1 | tail -3
echo "== greedy parity (llama-cli --temp 0)"; NTOK=48 bash tools/e2e_parity.sh mid-d128 q4_k_m 2>&1 | tail -4
echo "== 8B through llama_decode"
python - <<'PY'
import bench, json
g = bench.ensure_gguf()
for graphs in ("1", "0"):
    import os
    os.environ["GGML_B200_GRAPHS"] = graphs
    r = bench.run_harness(99, 512, 64, 8, 4, g)
    print("GGML_B200_GRAPHS=%s decode %.1f tok/s prefill %.1f tok/s" % (graphs, r["decode_tok_s"], r["prefill_tok_s"]))
PY
