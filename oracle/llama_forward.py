"""TEST INFRASTRUCTURE ONLY -- CPU forward of one llama ubatch composed from the C oracle (oracle.c).

Follows the node order of llm_build_llama (llama.cpp/src/llama-model.cpp:4093-4230) with the arithmetic of the reference
CPU backend (ggml-cpu.c: rms_norm, mul_mat -> vec_dot over q8_0/q8_K activations, rope, KV-store quantisation,
flash_attn_ext, silu*mul), operating on host copies of a ``LlamaGraph``'s weights and caches.  Used only by tests/,
``__graft_entry__.smoke()`` and bench.py's cpu_baseline leg as the checker / timed CPU baseline.
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import reflib as R  # noqa: E402


def _host(buf):
    return buf.cpu().numpy()


class HostModel:
    """host copy of a LlamaGraph (weights as raw GGUF bytes, caches as bytes)"""

    def __init__(self, g):
        self.g = g
        self._bytes = {buf.data_ptr(): buf for buf in g.keep}

    def w(self, tensor):
        """bytes (uint8) of a weight tensor described by a b200 Tensor"""
        buf = self._bytes[tensor.data]
        raw = _host(buf.view(self.g.torch.uint8) if buf.dtype != self.g.torch.uint8 else buf)
        t, K, N = tensor.type, tensor.ne[0], tensor.ne[1]
        return t, raw[:N * R.row_size(t, K)], K, N


def forward(g, emb, pos, mask, kv_head, n_kv, layers=None, caches=None):
    """emb f32 [T, E], pos i32 [T], mask f32 [Tp, n_kv] (0/-inf) -> (logits f32 [T, V], caches)

    caches: list of (k_bytes, v_bytes) uint8 arrays [n_ctx, row_size(Hkv*D)] per layer; taken from the device when None.
    The KV rows of this ubatch are written at kv_head before attention, as the graph does.
    """
    hm = HostModel(g)
    T = emb.shape[0]
    E, H, Hkv, D, FF = g.E, g.H, g.Hkv, g.D, g.FF
    kvt = g.kv_type
    rs_row = R.row_size(kvt, Hkv * D)
    rs_head = R.row_size(kvt, D)
    mask16 = mask.astype(np.float16)
    x = emb.astype(np.float32)
    nl = len(g.layers) if layers is None else layers
    if caches is None:
        caches = [(_host(lw["k_cache"]).reshape(g.n_ctx, rs_row).copy(), _host(lw["v_cache"]).reshape(g.n_ctx, rs_row).copy())
                  for lw in g.layers[:nl]]

    def mm(tensor, act):
        t, raw, K, N = hm.w(tensor)
        return R.orc_mul_mat(t, raw, act, N, K)

    def normw(tensor):
        buf = hm._bytes[tensor.data]
        return _host(buf)

    def kv_store(rows):                     # rows f32 [T, Hkv*D] -> bytes [T, rs_row]
        if kvt == R.F16:
            return rows.astype(np.float16).view(np.uint8).reshape(T, -1)
        return R.orc_quantize_act(kvt, rows.reshape(T * Hkv, D)).reshape(T, rs_row)

    for il in range(nl):
        lw = g.layers[il]
        xn = R.orc_rms_norm(x, 1e-5) * normw(lw["attn_norm"])
        q, k, v = mm(lw["wq"], xn), mm(lw["wk"], xn), mm(lw["wv"], xn)
        qr = R.orc_rope(q.reshape(T, H, D), pos, D, 0, g.rope_base, n_ctx_orig=8192)
        kr = R.orc_rope(k.reshape(T, Hkv, D), pos, D, 0, g.rope_base, n_ctx_orig=8192)
        kc, vc = caches[il]
        kc[kv_head:kv_head + T] = kv_store(kr.reshape(T, Hkv * D))
        vc[kv_head:kv_head + T] = kv_store(v)
        # cache rows are [cell][head][D]; the oracle wants [Hkv, n_kv, row_size(D)]
        kb = np.ascontiguousarray(kc[:n_kv].reshape(n_kv, Hkv, rs_head).transpose(1, 0, 2))
        vb = np.ascontiguousarray(vc[:n_kv].reshape(n_kv, Hkv, rs_head).transpose(1, 0, 2))
        att = R.orc_flash_attn(np.ascontiguousarray(qr.transpose(1, 0, 2)), kb, vb, mask16, D, n_kv, Hkv, kvt, kvt, 1.0 / math.sqrt(D))
        x = mm(lw["wo"], att.reshape(T, H * D)) + x
        xn = R.orc_rms_norm(x, 1e-5) * normw(lw["ffn_norm"])
        up, gate = mm(lw["up"], xn), mm(lw["gate"], xn)
        x = mm(lw["down"], R.orc_silu_mul(gate, up)) + x
    xn = R.orc_rms_norm(x, 1e-5) * normw(g.output_norm)
    logits = mm(g.output, xn)
    return logits, caches
