"""GPU parity for the glue ops between the matmuls (rms_norm, rope, KV-store cpy, add/mul, swiglu, get_rows, soft_max,
argsort, sum_rows, scale) against the oracle.  Integer/byte results (quantised KV store, argsort) are bit-exact."""
import numpy as np
import pytest

import reflib as R
from util import dev_bytes, to_dev

pytestmark = pytest.mark.gpu


def run(b200, ctx, op_id, out_shape, out_type, srcs, params=None, out_nbytes=None):
    """srcs: list of (np array or None, type, ne[, nb]).  Returns raw bytes of dst."""
    devs, tens = [], []
    for s in srcs:
        if s is None:
            tens.append(None)
            continue
        arr, t, ne = s[0], s[1], s[2]
        d = to_dev(arr.view(np.uint8).reshape(-1))
        devs.append(d)
        tens.append(b200.tensor(d.data_ptr(), t, ne, s[3] if len(s) > 3 else None))
    be, bb = b200.BLOCK[out_type]
    n = int(np.prod(out_shape))
    nbytes = out_nbytes or n // be * bb
    out = dev_bytes(nbytes, 0xEE)
    op = b200.make_op(op_id, b200.tensor(out.data_ptr(), out_type, list(out_shape)), tens, params)
    assert b200.supports(op), "op %d refused" % op_id
    ctx.compute_op(op)
    ctx.sync()
    return out.cpu().numpy()


def test_rms_norm_and_fused_mul(b200, ctx):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((5, 4096)) * 3).astype(np.float32)
    w = (1 + 0.02 * rng.standard_normal(4096)).astype(np.float32)
    got = run(b200, ctx, b200.OP_RMS_NORM, [4096, 5], b200.F32, [(x, b200.F32, [4096, 5])], [1e-5]).view(np.float32).reshape(5, 4096)
    want = R.orc_rms_norm(x, 1e-5)
    assert np.array_equal(got, want) or np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    got2 = run(b200, ctx, b200.OP_RMS_NORM_MUL, [4096, 5], b200.F32, [(x, b200.F32, [4096, 5]), (w, b200.F32, [4096])], [1e-5]).view(np.float32).reshape(5, 4096)
    assert np.abs(got2 - want * w).max() <= 1e-6 * np.abs(want).max()


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("ff", [False, True])
def test_rope(b200, ctx, mode, ff):
    rng = np.random.default_rng(2)
    n_tok, n_head, D = 7, 8, 128
    x = rng.standard_normal((n_tok, n_head, D)).astype(np.float32)
    pos = np.array([0, 1, 5, 100, 1000, 4095, 8191], np.int32)
    freq = rng.uniform(1.0, 8.0, D // 2).astype(np.float32) if ff else None
    params = [0, D, mode, 0, 8192, 500000.0, 1.0, 0.0, 1.0, 32.0, 1.0]
    srcs = [(x, b200.F32, [D, n_head, n_tok]), (pos, b200.I32, [n_tok]), (freq, b200.F32, [D // 2]) if ff else None]
    got = run(b200, ctx, b200.OP_ROPE, [D, n_head, n_tok], b200.F32, srcs, params).view(np.float32).reshape(n_tok, n_head, D)
    want = R.orc_rope(x, pos, D, mode, 500000.0, n_ctx_orig=8192, freq_factors=freq)
    # CUDA sinf/cosf vs glibc: <= 2 ulp on the trig values
    assert np.abs(got - want).max() <= 2e-6 * np.abs(x).max()


@pytest.mark.parametrize("mode", [0, 2])
def test_rope_cpu_exact_mode_bit_identical(b200, ctx, mode):
    """option cpu_exact: sin/cos through the restated glibc sinf/cosf (common.cuh) -> the oracle's rope (same libm, itself
    bit-identical to the reference golden) bit for bit, including positions that take glibc's large-argument reduction"""
    rng = np.random.default_rng(12)
    n_tok, n_head, D = 9, 4, 128
    x = rng.standard_normal((n_tok, n_head, D)).astype(np.float32)
    pos = np.array([0, 1, 5, 100, 119, 121, 4095, 8191, 131071], np.int32)
    ctx.set_option("cpu_exact", 1)
    try:
        for base, ff in ((10000.0, None), (500000.0, rng.uniform(1.0, 8.0, D // 2).astype(np.float32))):
            params = [0, D, mode, 0, 8192, base, 1.0, 0.0, 1.0, 32.0, 1.0]
            srcs = [(x, b200.F32, [D, n_head, n_tok]), (pos, b200.I32, [n_tok]), (ff, b200.F32, [D // 2]) if ff is not None else None]
            got = run(b200, ctx, b200.OP_ROPE, [D, n_head, n_tok], b200.F32, srcs, params).view(np.float32).reshape(n_tok, n_head, D)
            want = R.orc_rope(x, pos, D, mode, base, n_ctx_orig=8192, freq_factors=ff)
            assert np.array_equal(got, want), (base, np.abs(got - want).max())
    finally:
        ctx.set_option("cpu_exact", 0)


def test_rope_yarn(b200, ctx):
    rng = np.random.default_rng(3)
    n_tok, n_head, D = 4, 4, 64
    x = rng.standard_normal((n_tok, n_head, D)).astype(np.float32)
    pos = np.array([3, 77, 2048, 9000], np.int32)
    params = [0, D, 0, 0, 4096, 10000.0, 0.25, 1.0, 1.1, 32.0, 1.0]
    got = run(b200, ctx, b200.OP_ROPE, [D, n_head, n_tok], b200.F32, [(x, b200.F32, [D, n_head, n_tok]), (pos, b200.I32, [n_tok]), None], params)
    want = R.orc_rope(x, pos, D, 0, 10000.0, freq_scale=0.25, ext_factor=1.0, attn_factor=1.1, n_ctx_orig=4096)
    assert np.abs(got.view(np.float32).reshape(n_tok, n_head, D) - want).max() <= 4e-6 * np.abs(x).max()


@pytest.mark.parametrize("dt", [R.Q8_0, R.Q4_0, R.F16])
def test_kv_store_cpy_bit_exact(b200, ctx, dt):
    """f32 K/V rows -> cache type, exactly the bytes the CPU backend would write (KV-store CPY)"""
    rng = np.random.default_rng(4)
    n_tok, Hkv, D = 5, 8, 128
    x = (rng.standard_normal((n_tok, Hkv, D)) * 2).astype(np.float32)
    x[1, 2, :32] = 0
    total = n_tok * Hkv * D
    got = run(b200, ctx, b200.OP_CPY, [total], dt, [(x, b200.F32, [D, Hkv, n_tok])])
    if dt == R.F16:
        want = x.astype(np.float16).view(np.uint8).reshape(-1)
    else:
        want = R.orc_quantize_act(dt, x.reshape(-1, D)).reshape(-1)
    assert np.array_equal(got, want)


def test_cpy_strided_and_dequant(b200, ctx):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((6, 4, 64)).astype(np.float32)          # ne = [64, 4, 6]
    # permuted source view (swap dims 1,2) -> contiguous
    nb = [4, 64 * 4 * 4, 64 * 4, 64 * 4 * 6]
    got = run(b200, ctx, b200.OP_CONT, [64, 6, 4], b200.F32, [(x, b200.F32, [64, 6, 4], nb)]).view(np.float32).reshape(4, 6, 64)
    assert np.array_equal(got, x.transpose(1, 0, 2))
    q = R.orc_quantize_act(R.Q8_0, x.reshape(-1, 64))
    got = run(b200, ctx, b200.OP_CPY, [64, 24], b200.F32, [(q, b200.Q8_0, [64, 24])]).view(np.float32).reshape(24, 64)
    assert np.array_equal(got, R.orc_dequantize(R.Q8_0, q.reshape(-1), 64))


def test_add_mul_div_broadcast_and_swiglu(b200, ctx):
    rng = np.random.default_rng(6)
    a = rng.standard_normal((3, 5, 256)).astype(np.float32)
    b = rng.standard_normal((256,)).astype(np.float32)
    c = rng.standard_normal((3, 1, 256)).astype(np.float32) + 3
    for op, fn in [(b200.OP_ADD, np.add), (b200.OP_MUL, np.multiply), (b200.OP_SUB, np.subtract)]:
        got = run(b200, ctx, op, [256, 5, 3], b200.F32, [(a, b200.F32, [256, 5, 3]), (b, b200.F32, [256])]).view(np.float32).reshape(a.shape)
        assert np.array_equal(got, fn(a, b))
    got = run(b200, ctx, b200.OP_DIV, [256, 5, 3], b200.F32, [(a, b200.F32, [256, 5, 3]), (c, b200.F32, [256, 1, 3])]).view(np.float32).reshape(a.shape)
    assert np.array_equal(got, a / c)
    g, u = rng.standard_normal((2, 14336)).astype(np.float32), rng.standard_normal((2, 14336)).astype(np.float32)
    got = run(b200, ctx, b200.OP_SWIGLU_FUSED, [14336, 2], b200.F32, [(g, b200.F32, [14336, 2]), (u, b200.F32, [14336, 2])]).view(np.float32).reshape(2, -1)
    g[0, :16] = [-200, -130, -100, -90, -88, -20, 0, 1e-8, 20, 88, 90, 100, 130, 200, -0.0, 7.5]
    got = run(b200, ctx, b200.OP_SWIGLU_FUSED, [14336, 2], b200.F32, [(g, b200.F32, [14336, 2]), (u, b200.F32, [14336, 2])]).view(np.float32).reshape(2, -1)
    want = R.orc_silu_mul(g, u)
    # bit-exact: kernels and oracle both restate the reference's polynomial ggml_v_expf (pinned to the reference golden)
    assert np.array_equal(got, want), np.abs(got - want).max()
    got = run(b200, ctx, b200.OP_SCALE, [256, 5, 3], b200.F32, [(a, b200.F32, [256, 5, 3])], [0.125]).view(np.float32).reshape(a.shape)
    assert np.array_equal(got, a * np.float32(0.125))


def test_get_rows_f32_and_quantised(b200, ctx):
    rng = np.random.default_rng(7)
    src = rng.standard_normal((50, 512)).astype(np.float32)
    idx = np.array([3, 49, 0, 3, 17], np.int32)
    got = run(b200, ctx, b200.OP_GET_ROWS, [512, 5], b200.F32, [(src, b200.F32, [512, 50]), (idx, b200.I32, [5])]).view(np.float32).reshape(5, 512)
    assert np.array_equal(got, src[idx])
    from util import rand_quant_rows
    for t in (R.Q4_K, R.Q6_K, R.Q8_0, R.Q4_0, R.Q5_K):
        W = rand_quant_rows(t, 50, 512, rng)
        got = run(b200, ctx, b200.OP_GET_ROWS, [512, 5], b200.F32, [(W, t, [512, 50]), (idx, b200.I32, [5])]).view(np.float32).reshape(5, 512)
        want = R.orc_dequantize(t, W, 512)[idx]
        assert np.array_equal(got, want), R.TYPE_NAMES[t]


def test_soft_max_argsort_sum_rows(b200, ctx):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((6, 8)).astype(np.float32)          # MoE router: 8 experts
    got = run(b200, ctx, b200.OP_SOFT_MAX, [8, 6], b200.F32, [(x, b200.F32, [8, 6]), None], [1.0, 0.0]).view(np.float32).reshape(6, 8)
    want = R.orc_soft_max(x, None, 1.0)
    assert np.abs(got - want).max() <= 2e-7
    got = run(b200, ctx, b200.OP_ARGSORT, [8, 6], b200.I32, [(x, b200.F32, [8, 6])], [1]).view(np.int32).reshape(6, 8)
    assert np.array_equal(got, np.argsort(-x, axis=1, kind="stable"))
    got = run(b200, ctx, b200.OP_SUM_ROWS, [1, 6], b200.F32, [(x, b200.F32, [8, 6])]).view(np.float32).reshape(6)
    assert np.abs(got - x.astype(np.float64).sum(1)).max() <= 1e-6
    big = rng.standard_normal((3, 1000)).astype(np.float32)
    m = np.where(rng.random((3, 1000)) < 0.3, -np.inf, 0).astype(np.float16)
    got = run(b200, ctx, b200.OP_SOFT_MAX, [1000, 3], b200.F32, [(big, b200.F32, [1000, 3]), (m, b200.F16, [1000, 3])], [0.5, 0.0]).view(np.float32).reshape(3, 1000)
    want = R.orc_soft_max(big, m, 0.5)
    assert np.abs(got - want).max() <= 1e-6


def test_argmax_last_maximum_like_the_cpu(b200, ctx):
    """greedy sampling on the device: ggml_vec_argmax_f32 keeps the LAST index that holds the maximum (ggml-cpu.c:2393-2401)"""
    rng = np.random.default_rng(12)
    x = rng.standard_normal((5, 128256)).astype(np.float32)          # full Llama-3 vocabulary rows
    x[1, 7] = x[1, 100000] = 9.0                                     # tie: the later index wins
    x[2, :] = -np.inf                                                # every element equals the running maximum: last index
    x[3, 128255] = 11.0
    x[4, 0] = 12.0
    got = run(b200, ctx, b200.OP_ARGMAX, [5], b200.I32, [(x, b200.F32, [128256, 5])]).view(np.int32).reshape(5)
    want = np.array([R.ref_argmax_row(r) for r in x], np.int32)
    assert np.array_equal(got, want), (got, want)
    assert got[1] == 100000 and got[2] == 128255 and got[3] == 128255 and got[4] == 0
