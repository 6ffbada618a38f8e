"""ctypes access to (a) our C oracle (oracle/liboracle.so) and (b) the reference's own CPU code
(oracle/_ref/libggml-{base,cpu}.so, built from /root/reference/llama.cpp by oracle/Makefile).

TEST INFRASTRUCTURE ONLY -- imported by tests/, golden generation and bench.py's cpu_baseline.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

# ggml type ids (ggml/include/ggml.h:350-390)
F32, F16, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K, Q8_K, I32 = 0, 1, 2, 8, 12, 13, 14, 15, 26
TYPE_NAMES = {F32: "f32", F16: "f16", Q4_0: "q4_0", Q8_0: "q8_0", Q4_K: "q4_K", Q5_K: "q5_K", Q6_K: "q6_K", Q8_K: "q8_K"}
BLOCK = {F32: (1, 4), F16: (1, 2), Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176),
         Q6_K: (256, 210), Q8_K: (256, 292)}
QUANT_TYPES = [Q4_0, Q8_0, Q4_K, Q5_K, Q6_K]


def row_size(t, k):
    be, bb = BLOCK[t]
    assert k % be == 0
    return k // be * bb


def act_type(t):
    """vec_dot_type of the CPU backend (ggml-cpu.c:266-341)"""
    return Q8_0 if t in (Q4_0, Q8_0) else Q8_K


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- our oracle
_orc = None


def oracle():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        L = C.CDLL(ORACLE_SO)
        L.orc_vec_dot.restype = C.c_float
        L.orc_f16_to_f32.restype = C.c_float
        L.orc_f16_to_f32.argtypes = [C.c_uint16]
        L.orc_f32_to_f16.restype = C.c_uint16
        L.orc_f32_to_f16.argtypes = [C.c_float]
        _orc = L
    return _orc


def orc_quantize_act(t_act, x):
    """x f32 [rows, K] -> uint8 [rows, row_size]"""
    x = np.ascontiguousarray(x, np.float32)
    rows, K = x.shape
    out = np.zeros((rows, row_size(t_act, K)), np.uint8)
    fn = {Q8_0: oracle().orc_quantize_row_q8_0, Q8_K: oracle().orc_quantize_row_q8_K, Q4_0: oracle().orc_quantize_row_q4_0}[t_act]
    for r in range(rows):
        fn(_ptr(x[r]), _ptr(out[r]), C.c_int64(K))
    return out


def orc_dequantize(t, w, K):
    w = np.ascontiguousarray(w, np.uint8).reshape(-1, row_size(t, K))
    out = np.zeros((w.shape[0], K), np.float32)
    for r in range(w.shape[0]):
        oracle().orc_dequantize_row(t, _ptr(w[r]), _ptr(out[r]), C.c_int64(K))
    return out


def orc_block_sums(t, wrow, arow, K):
    nb = K // BLOCK[t][0]
    P = np.zeros(nb, np.int32)
    M = np.zeros(nb, np.int32)
    oracle().orc_block_sums(t, _ptr(np.ascontiguousarray(wrow)), _ptr(np.ascontiguousarray(arow)), C.c_int64(K), _ptr(P), _ptr(M))
    return P, M


def orc_mul_mat(t, W, x, N, K):
    """W uint8 [N*row_size] (or f32/f16 raw bytes), x f32 [M, K] -> f32 [M, N]"""
    x = np.ascontiguousarray(x, np.float32)
    M = x.shape[0]
    W = np.ascontiguousarray(W)
    dst = np.zeros((M, N), np.float32)
    oracle().orc_mul_mat(t, _ptr(W), _ptr(x), _ptr(dst), C.c_int64(N), C.c_int64(K), C.c_int64(M))
    return dst


def orc_mul_mat_id(t, As, b, ids, N, K, n_expert):
    """As bytes of [n_expert, N, K]; b f32 [n_tok, b_ne1, K]; ids i32 [n_tok, n_used] -> [n_tok, n_used, N]"""
    b = np.ascontiguousarray(b, np.float32)
    ids = np.ascontiguousarray(ids, np.int32)
    n_tok, n_used = ids.shape
    dst = np.zeros((n_tok, n_used, N), np.float32)
    oracle().orc_mul_mat_id(t, _ptr(np.ascontiguousarray(As)), _ptr(b), _ptr(ids), _ptr(dst), C.c_int64(N), C.c_int64(K),
                            C.c_int64(n_expert), C.c_int64(n_used), C.c_int64(n_tok), C.c_int64(b.shape[1]))
    return dst


def orc_rms_norm(x, eps):
    x = np.ascontiguousarray(x, np.float32)
    y = np.zeros_like(x)
    oracle().orc_rms_norm(_ptr(x), _ptr(y), C.c_int64(x.shape[-1]), C.c_int64(x.size // x.shape[-1]), C.c_float(eps))
    return y


def orc_rope(x, pos, n_dims, mode, freq_base, freq_scale=1.0, ext_factor=0.0, attn_factor=1.0, beta_fast=32.0,
             beta_slow=1.0, n_ctx_orig=4096, freq_factors=None):
    """x f32 [n_tok, n_head, ne0]"""
    x = np.ascontiguousarray(x, np.float32)
    pos = np.ascontiguousarray(pos, np.int32)
    n_tok, n_head, ne0 = x.shape
    y = np.zeros_like(x)
    ff = None if freq_factors is None else np.ascontiguousarray(freq_factors, np.float32)
    oracle().orc_rope(_ptr(x), _ptr(y), _ptr(pos), _ptr(ff) if ff is not None else None, C.c_int64(ne0), C.c_int64(n_head),
                      C.c_int64(n_tok), C.c_int(n_dims), C.c_int(mode), C.c_int(n_ctx_orig), C.c_float(freq_base),
                      C.c_float(freq_scale), C.c_float(ext_factor), C.c_float(attn_factor), C.c_float(beta_fast),
                      C.c_float(beta_slow))
    return y


def orc_soft_max(x, mask_f16, scale):
    x = np.ascontiguousarray(x, np.float32)
    y = np.zeros_like(x)
    m = None if mask_f16 is None else np.ascontiguousarray(mask_f16, np.float16)
    oracle().orc_soft_max(_ptr(x), _ptr(m) if m is not None else None, _ptr(y), C.c_int64(x.shape[-1]),
                          C.c_int64(x.size // x.shape[-1]), C.c_float(scale))
    return y


def orc_silu_mul(g, u):
    g = np.ascontiguousarray(g, np.float32)
    u = np.ascontiguousarray(u, np.float32)
    y = np.zeros_like(g)
    oracle().orc_silu_mul(_ptr(g), _ptr(u), _ptr(y), C.c_int64(g.size))
    return y


def orc_flash_attn(q, k_bytes, v_bytes, mask_f16, D, n_kv, Hkv, type_k, type_v, scale, softcap=0.0):
    """q f32 [H, n_q, D]; k_bytes/v_bytes uint8 [Hkv, n_kv, row_size]; mask f16 [n_q_pad, n_kv] -> f32 [n_q, H, D]"""
    q = np.ascontiguousarray(q, np.float32)
    H, n_q, _ = q.shape
    k_bytes = np.ascontiguousarray(k_bytes, np.uint8)
    v_bytes = np.ascontiguousarray(v_bytes, np.uint8)
    krs, vrs = row_size(type_k, D), row_size(type_v, D)
    dst = np.zeros((n_q, H, D), np.float32)
    m = None if mask_f16 is None else np.ascontiguousarray(mask_f16, np.float16)
    oracle().orc_flash_attn_ext(_ptr(q), _ptr(k_bytes), _ptr(v_bytes), _ptr(m) if m is not None else None, _ptr(dst),
                                C.c_int64(D), C.c_int64(n_q), C.c_int64(H), C.c_int64(n_kv), C.c_int64(Hkv),
                                C.c_int(type_k), C.c_int(type_v), C.c_size_t(krs), C.c_size_t(krs * n_kv), C.c_size_t(vrs),
                                C.c_size_t(vrs * n_kv), C.c_size_t(2 * n_kv), C.c_float(scale), C.c_float(softcap))
    return dst


# ----------------------------------------------------------------------------- the reference itself
_ref = None


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libggml-cpu.so"))


class _Ref:
    def __init__(self):
        self.base = C.CDLL(os.path.join(REF_DIR, "libggml-base.so"), mode=C.RTLD_GLOBAL)
        self.cpu = C.CDLL(os.path.join(REF_DIR, "libggml-cpu.so"), mode=C.RTLD_GLOBAL)
        self.base.ggml_quantize_chunk.restype = C.c_size_t
        self.base.ggml_quantize_chunk.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
        self.cpu.ggml_cpu_init()

    def quantize_weights(self, t, w):
        """f32 [rows, K] -> reference-quantised bytes [rows*row_size] via ggml_quantize_chunk"""
        w = np.ascontiguousarray(w, np.float32)
        rows, K = w.shape
        out = np.zeros(rows * row_size(t, K), np.uint8)
        n = self.base.ggml_quantize_chunk(t, _ptr(w), _ptr(out), 0, rows, K, None)
        assert n == out.size
        return out

    def quantize_act(self, t_act, x):
        """the CPU backend's from_float for the vec_dot_type (SIMD build as compiled)"""
        x = np.ascontiguousarray(x, np.float32)
        rows, K = x.shape
        out = np.zeros((rows, row_size(t_act, K)), np.uint8)
        fn = {Q8_0: self.cpu.quantize_row_q8_0, Q8_K: self.cpu.quantize_row_q8_K, Q4_0: self.base.quantize_row_q4_0_ref}[t_act]
        for r in range(rows):
            fn(_ptr(x[r]), _ptr(out[r]), C.c_int64(K))
        return out

    def dequantize(self, t, w, K):
        w = np.ascontiguousarray(w, np.uint8).reshape(-1, row_size(t, K))
        out = np.zeros((w.shape[0], K), np.float32)
        fn = getattr(self.base, "dequantize_row_" + TYPE_NAMES[t])
        for r in range(w.shape[0]):
            fn(_ptr(w[r]), _ptr(out[r]), C.c_int64(K))
        return out

    def vec_dot(self, t, wrow, arow, K):
        fn = getattr(self.cpu, "ggml_vec_dot_%s_%s" % (TYPE_NAMES[t], TYPE_NAMES[act_type(t)]))
        s = C.c_float(0)
        fn(C.c_int(K), C.byref(s), C.c_size_t(0), _ptr(np.ascontiguousarray(wrow)), C.c_size_t(0),
           _ptr(np.ascontiguousarray(arow)), C.c_size_t(0), C.c_int(1))
        return s.value

    def mul_mat(self, t, W, x, N, K):
        """reference result of MUL_MAT as ggml_compute_forward_mul_mat computes it: quantise each src1 row
        with the CPU from_float, one vec_dot per (row, col)."""
        x = np.ascontiguousarray(x, np.float32)
        act = self.quantize_act(act_type(t), x)
        W = np.ascontiguousarray(W, np.uint8).reshape(N, row_size(t, K))
        out = np.zeros((x.shape[0], N), np.float32)
        for m in range(x.shape[0]):
            for n in range(N):
                out[m, n] = self.vec_dot(t, W[n], act[m], K)
        return out


def ref():
    global _ref
    if _ref is None:
        _ref = _Ref()
    return _ref


# ----------------------------------------------------------------------------- one-op ggml graphs on the reference CPU backend
class _InitParams(C.Structure):
    _fields_ = [("mem_size", C.c_size_t), ("mem_buffer", C.c_void_p), ("no_alloc", C.c_bool)]


class RefGraph:
    """Builds and runs a small ggml graph on the reference's CPU backend through its public C API (ggml.h / ggml-cpu.h),
    the way tests/test-backend-ops.cpp does.  Used by tests/golden/make_golden_ops.py (fixtures) and, when oracle/_ref is
    present, directly by tests/test_oracle_pin.py."""

    def __init__(self, mem_mb=256):
        r = ref()
        self.b, self.c = r.base, r.cpu
        b = self.b
        vp, i64, f32, i32 = C.c_void_p, C.c_int64, C.c_float, C.c_int
        b.ggml_init.restype = vp
        b.ggml_init.argtypes = [_InitParams]
        b.ggml_free.argtypes = [vp]
        b.ggml_new_tensor_4d.restype = vp
        b.ggml_new_tensor_4d.argtypes = [vp, i32, i64, i64, i64, i64]
        b.ggml_get_data.restype = vp
        b.ggml_get_data.argtypes = [vp]
        b.ggml_nbytes.restype = C.c_size_t
        b.ggml_nbytes.argtypes = [vp]
        for name, args in {
            "ggml_rms_norm": [vp, vp, f32], "ggml_mul": [vp, vp, vp], "ggml_add": [vp, vp, vp], "ggml_silu": [vp, vp],
            "ggml_rope_ext": [vp, vp, vp, vp, i32, i32, i32, f32, f32, f32, f32, f32, f32],
            "ggml_soft_max_ext": [vp, vp, vp, f32, f32],
            "ggml_flash_attn_ext": [vp, vp, vp, vp, vp, f32, f32, f32],
            "ggml_cpy": [vp, vp, vp], "ggml_cont": [vp, vp], "ggml_permute": [vp, vp, i32, i32, i32, i32],
            "ggml_get_rows": [vp, vp, vp], "ggml_mul_mat": [vp, vp, vp], "ggml_new_graph": [vp],
            "ggml_cast": [vp, vp, i32], "ggml_rope_ext_inplace": [vp, vp, vp, vp, i32, i32, i32, f32, f32, f32, f32, f32, f32],
            "ggml_argmax": [vp, vp],
        }.items():
            fn = getattr(b, name)
            fn.restype = vp
            fn.argtypes = args
        b.ggml_flash_attn_ext_set_prec.argtypes = [vp, i32]
        b.ggml_build_forward_expand.argtypes = [vp, vp]
        self.c.ggml_graph_compute_with_ctx.argtypes = [vp, vp, i32]
        self.c.ggml_graph_compute_with_ctx.restype = i32
        self.ctx = b.ggml_init(_InitParams(mem_mb << 20, None, False))
        assert self.ctx

    def close(self):
        self.b.ggml_free(self.ctx)
        self.ctx = None

    def tensor(self, t, ne, data=None):
        ne = list(ne) + [1] * (4 - len(ne))
        x = self.b.ggml_new_tensor_4d(self.ctx, t, *ne)
        if data is not None:
            data = np.ascontiguousarray(data)
            n = self.b.ggml_nbytes(x)
            assert data.nbytes == n, (data.nbytes, n)
            C.memmove(self.b.ggml_get_data(x), _ptr(data), n)
        return x

    def run(self, out, dtype, shape, threads=4):
        g = self.b.ggml_new_graph(self.ctx)
        self.b.ggml_build_forward_expand(g, out)
        assert self.c.ggml_graph_compute_with_ctx(self.ctx, g, threads) == 0
        n = self.b.ggml_nbytes(out)
        buf = np.zeros(n, np.uint8)
        C.memmove(_ptr(buf), self.b.ggml_get_data(out), n)
        return buf.view(dtype).reshape(shape).copy()


def ref_rms_norm(x, eps):
    g = RefGraph()
    rows, n = x.shape
    y = g.run(g.b.ggml_rms_norm(g.ctx, g.tensor(F32, [n, rows], x.astype(np.float32)), eps), np.float32, (rows, n))
    g.close()
    return y


def ref_rope(x, pos, n_dims, mode, freq_base, freq_scale=1.0, ext_factor=0.0, attn_factor=1.0, beta_fast=32.0, beta_slow=1.0,
             n_ctx_orig=4096, freq_factors=None):
    """x f32 [n_tok, n_head, ne0]"""
    g = RefGraph()
    n_tok, n_head, ne0 = x.shape
    a = g.tensor(F32, [ne0, n_head, n_tok], x.astype(np.float32))
    p = g.tensor(I32, [n_tok], np.asarray(pos, np.int32))
    ff = g.tensor(F32, [ne0 // 2], np.asarray(freq_factors, np.float32)) if freq_factors is not None else None
    out = g.b.ggml_rope_ext(g.ctx, a, p, ff, n_dims, mode, n_ctx_orig, freq_base, freq_scale, ext_factor, attn_factor, beta_fast, beta_slow)
    y = g.run(out, np.float32, (n_tok, n_head, ne0))
    g.close()
    return y


def ref_soft_max(x, mask_f16, scale):
    g = RefGraph()
    rows, n = x.shape
    a = g.tensor(F32, [n, rows], x.astype(np.float32))
    m = g.tensor(F16, [n, mask_f16.shape[0]], mask_f16.astype(np.float16)) if mask_f16 is not None else None
    y = g.run(g.b.ggml_soft_max_ext(g.ctx, a, m, scale, 0.0), np.float32, (rows, n))
    g.close()
    return y


def ref_silu_mul(gate, up):
    g = RefGraph()
    n = gate.size
    a = g.tensor(F32, [n], gate.astype(np.float32))
    b = g.tensor(F32, [n], up.astype(np.float32))
    y = g.run(g.b.ggml_mul(g.ctx, g.b.ggml_silu(g.ctx, a), b), np.float32, gate.shape)
    g.close()
    return y


def ref_flash_attn(q, k_bytes, v_bytes, mask_f16, D, n_kv, Hkv, type_k, type_v, scale, softcap=0.0):
    """same contract as orc_flash_attn: q f32 [H, n_q, D]; k/v uint8 [Hkv, n_kv, row_size]; mask f16 [n_q_pad, n_kv] -> [n_q, H, D]"""
    g = RefGraph()
    H, n_q, _ = q.shape
    qt = g.tensor(F32, [D, n_q, H], q.astype(np.float32))
    kt = g.tensor(type_k, [D, n_kv, Hkv], k_bytes)
    vt = g.tensor(type_v, [D, n_kv, Hkv], v_bytes)
    mt = g.tensor(F16, [n_kv, mask_f16.shape[0]], mask_f16.astype(np.float16)) if mask_f16 is not None else None
    out = g.b.ggml_flash_attn_ext(g.ctx, qt, kt, vt, mt, scale, 0.0, softcap)
    g.b.ggml_flash_attn_ext_set_prec(out, 1)         # GGML_PREC_F32 (ggml.h:395-398), as llama.cpp sets it (llama-graph.cpp:1197)
    y = g.run(out, np.float32, (n_q, H, D))
    g.close()
    return y


def ref_argmax_row(x):
    """ggml_vec_argmax_f32 (ggml-cpu.c:2393-2401) restated: running maximum, index updated whenever the element equals it"""
    x = np.asarray(x, np.float32)
    m = np.maximum.accumulate(x)
    hit = np.nonzero(m == x)[0]
    return int(hit[-1]) if hit.size else 0


def ref_k_shift(kbytes, ktype, D, Hkv, n_cells, shift, n_rot, mode, freq_base, freq_scale=1.0, n_ctx_orig=8192):
    """llama_context::build_rope_shift (llama-context.cpp:464-514) run on the reference CPU backend: the cached K of one layer
    [D, Hkv, n_cells] re-rotated by the per-cell `shift`; quantised K goes through ggml_cast -> f32, rope in place, ggml_cpy back"""
    g = RefGraph()
    cur = g.tensor(ktype, [D, Hkv, n_cells], np.asarray(kbytes, np.uint8))
    sh = g.tensor(I32, [n_cells], np.asarray(shift, np.int32))
    if ktype in (Q8_0, Q4_0):
        tmp = g.b.ggml_cast(g.ctx, cur, F32)
        tmp = g.b.ggml_rope_ext_inplace(g.ctx, tmp, sh, None, n_rot, mode, n_ctx_orig, freq_base, freq_scale, 0.0, 1.0, 32.0, 1.0)
        out = g.b.ggml_cpy(g.ctx, tmp, cur)
    else:
        out = g.b.ggml_rope_ext_inplace(g.ctx, cur, sh, None, n_rot, mode, n_ctx_orig, freq_base, freq_scale, 0.0, 1.0, 32.0, 1.0)
    gr = g.b.ggml_new_graph(g.ctx)
    g.b.ggml_build_forward_expand(gr, out)
    assert g.c.ggml_graph_compute_with_ctx(g.ctx, gr, 4) == 0
    n = g.b.ggml_nbytes(cur)
    buf = np.zeros(n, np.uint8)
    C.memmove(_ptr(buf), g.b.ggml_get_data(cur), n)
    g.close()
    return buf


def ref_mul_mat_f32(w, x):
    """w f32 [N, K], x f32 [M, K] -> [M, N] through ggml_mul_mat on the reference CPU backend"""
    g = RefGraph()
    N, K = w.shape
    M = x.shape[0]
    out = g.b.ggml_mul_mat(g.ctx, g.tensor(F32, [K, N], w.astype(np.float32)), g.tensor(F32, [K, M], x.astype(np.float32)))
    y = g.run(out, np.float32, (M, N), threads=1)
    g.close()
    return y
