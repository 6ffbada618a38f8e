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


ROUTER_MARGINS = []          # filled by moe_ffn: tests use it to tell routing near-ties (where any two builds may pick different experts)


def moe_ffn(g, hm, lw, xn):
    """llm_graph_context::build_moe_ffn (llama-graph.cpp:834-952) with the CPU backend's arithmetic: F32 router matmul, soft_max over the
    experts, top-k by a descending argsort, weights = probs[selected] / their sum, experts through mul_mat_id, f32 weighted sum"""
    T = xn.shape[0]
    nE, nu = g.n_expert, g.n_used

    def raw3(tensor):
        buf = hm._bytes[tensor.data]
        t, K, N = tensor.type, tensor.ne[0], tensor.ne[1]
        return t, _host(buf)[:nE * N * R.row_size(t, K)], K, N

    t_, raw, K, N = hm.w(lw["gate_inp"])
    logits = R.orc_mul_mat(t_, raw, xn, N, K)                                    # [T, nE]
    probs = R.orc_soft_max(logits, None, 1.0)
    order = np.argsort(-probs, axis=1, kind="stable")
    sel = order[:, :nu].astype(np.int32)                                          # ggml argsort desc; probabilities do not tie
    srt = np.take_along_axis(probs, order, axis=1)
    ROUTER_MARGINS.append(srt[:, nu - 1] - srt[:, nu])                            # per token: gap between the last used and the first unused expert
    w = np.take_along_axis(probs, sel, axis=1)
    wsum = w.astype(np.float64).sum(axis=1).astype(np.float32)                    # ggml_vec_sum_f32: ggml_float accumulator
    wn = (w / wsum[:, None]).astype(np.float32)
    t_, raw, K, N = raw3(lw["up_exps"])
    up = R.orc_mul_mat_id(t_, raw, xn[:, None, :], sel, N, K, nE)                 # [T, nu, FF]
    t_, raw, K, N = raw3(lw["gate_exps"])
    gate = R.orc_mul_mat_id(t_, raw, xn[:, None, :], sel, N, K, nE)
    par = R.orc_silu_mul(gate, up)
    t_, raw, K, N = raw3(lw["down_exps"])
    ex = R.orc_mul_mat_id(t_, raw, par, sel, N, K, nE)                            # [T, nu, E]
    ex = (ex * wn[:, :, None]).astype(np.float32)
    out = ex[:, 0]
    for i in range(1, nu):
        out = (out + ex[:, i]).astype(np.float32)
    return out


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
        if getattr(g, "n_expert", 0):
            x = moe_ffn(g, hm, lw, xn) + x
            continue
        up, gate = mm(lw["up"], xn), mm(lw["gate"], xn)
        x = mm(lw["down"], R.orc_silu_mul(gate, up)) + x
    xn = R.orc_rms_norm(x, 1e-5) * normw(g.output_norm)
    logits = mm(g.output, xn)
    return logits, caches
