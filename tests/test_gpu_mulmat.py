"""GPU parity tests (through the C ABI): activation quantisers and the decode GEMV against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import reflib as R
from util import dev_bytes, rand_quant_rows, to_dev

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath_golden.npz"))


def gpu_quantize(b200, ctx, ta, x):
    rows, K = x.shape
    xd = to_dev(x)
    out = dev_bytes(rows * R.row_size(ta, K), 0)
    b200.check(b200.lib().b200_quantize_act(ctx.h, ta, xd.data_ptr(), out.data_ptr(), K, rows), "quantize_act")
    ctx.sync()
    return out.cpu().numpy().reshape(rows, -1)


@pytest.mark.parametrize("ta", [R.Q8_0, R.Q8_K])
def test_act_quantiser_bit_exact_vs_golden(b200, ctx, ta):
    got = gpu_quantize(b200, ctx, ta, G["act_x"])
    want = G["act_" + R.TYPE_NAMES[ta]]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ta", [R.Q8_0, R.Q8_K])
def test_act_quantiser_bit_exact_vs_oracle_large(b200, ctx, ta):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((7, 14336)) * rng.uniform(0.01, 30, (7, 1))).astype(np.float32)
    x[3, 512:768] = 0
    assert np.array_equal(gpu_quantize(b200, ctx, ta, x), R.orc_quantize_act(ta, x))


def gpu_mul_mat(b200, ctx, t, W, x, N, K, flags=0):
    M = x.shape[0]
    pad = b200.lib().b200_alloc_size(t, (C.c_int64 * 4)(K, N, 1, 1), W.size)
    Wd = dev_bytes(pad, 0)
    Wd[:W.size] = to_dev(W)
    xd = to_dev(x)
    out = dev_bytes(M * N * 4, 0xFF)
    import torch
    torch.cuda.synchronize()
    op = b200.make_op(b200.OP_MUL_MAT, b200.tensor(out.data_ptr(), b200.F32, [N, M]),
                      [b200.tensor(Wd.data_ptr(), t, [K, N], flags=flags), b200.tensor(xd.data_ptr(), b200.F32, [K, M])])
    assert b200.supports(op)
    ctx.compute_op(op)
    ctx.sync()
    return out.cpu().numpy().view(np.float32).reshape(M, N)


@pytest.mark.parametrize("K", [256, 1280])
@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_gemv_vs_golden_reference_output(b200, ctx, t, K):
    name = "%s_%d" % (R.TYPE_NAMES[t], K)
    got = gpu_mul_mat(b200, ctx, t, G["mm_W_" + name], G["mm_x_%d" % K], 12, K)
    want = G["mm_out_" + name]
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


SHAPES = [(16, 256), (17, 512), (333, 2048), (1024, 4096), (600, 5632), (4096, 4096), (300, 14336), (64, 28672)]


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_block_sums_bit_exact(b200, ctx, t, N, K):
    """per-block integer dot products of the CUDA pipeline == the oracle's, bit for bit"""
    rng = np.random.default_rng(N * 31 + K + t)
    N = min(N, 96)       # the oracle is scalar C; keep it to seconds
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((1, K)).astype(np.float32)
    nb = K // R.BLOCK[t][0]
    Wd = dev_bytes(W.size + 256, 0)
    Wd[:W.size] = to_dev(W)
    xd = to_dev(x)
    P = dev_bytes(N * nb * 4)
    M = dev_bytes(N * nb * 4)
    b200.check(b200.lib().b200_block_sums(ctx.h, t, Wd.data_ptr(), xd.data_ptr(), N, K, P.data_ptr(), M.data_ptr()), "block_sums")
    ctx.sync()
    Pg = P.cpu().numpy().view(np.int32).reshape(N, nb)
    Mg = M.cpu().numpy().view(np.int32).reshape(N, nb)
    act = R.orc_quantize_act(R.act_type(t), x)[0]
    rb = R.row_size(t, K)
    for n in range(N):
        p, m = R.orc_block_sums(t, W.reshape(N, rb)[n], act, K)
        assert np.array_equal(p, Pg[n]), "P differs row %d" % n
        assert np.array_equal(m, Mg[n]), "M differs row %d" % n


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_gemv_vs_oracle(b200, ctx, t, N, K, M):
    if M > 1 and (N, K) not in [(17, 512), (1024, 4096), (300, 14336)]:
        pytest.skip("multi-column cases on a subset of shapes")
    rng = np.random.default_rng(N + K + t + M)
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((M, K)).astype(np.float32)
    got = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    rows = np.unique(np.concatenate([np.arange(min(N, 24)), rng.integers(0, N, 24), [N - 1]]))
    rb = R.row_size(t, K)
    Wsub = W.reshape(N, rb)[rows].reshape(-1)
    want = R.orc_mul_mat(t, Wsub, x, len(rows), K)
    scale = max(np.abs(want).max(), 1e-6)
    assert np.abs(got[:, rows] - want).max() <= 3e-6 * scale
    assert np.isfinite(got).all()


def test_gemv_deterministic(b200, ctx):
    rng = np.random.default_rng(9)
    N, K, t = 4096, 14336, R.Q4_K
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((1, K)).astype(np.float32)
    a = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    b = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("wt", ["f32", "f16"])
def test_float_weight_matmul(b200, ctx, wt):
    rng = np.random.default_rng(2)
    N, K, M = 8, 4096, 3
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
    x = rng.standard_normal((M, K)).astype(np.float32)
    if wt == "f16":
        wb, t = w.astype(np.float16), R.F16
    else:
        wb, t = w, R.F32
    Wd, xd = to_dev(wb.view(np.uint8).reshape(-1)), to_dev(x)
    out = dev_bytes(M * N * 4)
    op = b200.make_op(b200.OP_MUL_MAT, b200.tensor(out.data_ptr(), b200.F32, [N, M]),
                      [b200.tensor(Wd.data_ptr(), t, [K, N]), b200.tensor(xd.data_ptr(), b200.F32, [K, M])])
    ctx.compute_op(op)
    ctx.sync()
    got = out.cpu().numpy().view(np.float32).reshape(M, N)
    want = R.orc_mul_mat(t, wb.view(np.uint8).reshape(-1), x, N, K)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_batched_columns_prequantised_path(b200, ctx, t):
    """M > 8 goes through the stand-alone quantiser + the pre-quantised activation path"""
    rng = np.random.default_rng(77 + t)
    N, K, M = 96, 2048, 19
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((M, K)).astype(np.float32)
    got = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    want = R.orc_mul_mat(t, W, x, N, K)
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


@pytest.mark.parametrize("M", [9, 16, 17, 32])
@pytest.mark.parametrize("N,K", [(16, 256), (48, 512), (1024, 4096), (4096, 4096), (304, 14336), (64, 28672), (160, 2304)])
@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_small_batch_mma_kernel_vs_oracle(b200, ctx, t, N, K, M):
    """continuous-batching decode (9..32 token columns): gemv_mma.cu (mma.sync s8 over the streamed GGUF rows, split-K ranges of
    2048 with a fixed-order reduction) against the oracle; every row of the first / last tile and a random sample"""
    if M not in (17, 32) and (N, K) not in [(48, 512), (1024, 4096)]:
        pytest.skip("column-count sweep on a subset of shapes")
    rng = np.random.default_rng(N + K + t + M)
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((M, K)).astype(np.float32)
    l0 = ctx.launches()
    got = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    assert ctx.launches() - l0 == (3 if K > 2048 else 2), "expected quantise + gemv_mma (+ split-K reduce)"
    rows = np.unique(np.concatenate([np.arange(min(N, 16)), rng.integers(0, N, 16), np.arange(N - 16, N)]))
    rb = R.row_size(t, K)
    Wsub = W.reshape(N, rb)[rows].reshape(-1)
    want = R.orc_mul_mat(t, Wsub, x, len(rows), K)
    scale = max(np.abs(want).max(), 1e-6)
    assert np.abs(got[:, rows] - want).max() <= 3e-6 * scale
    assert np.isfinite(got).all()
    again = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    assert np.array_equal(got, again), "split-K reduction must be deterministic"


@pytest.mark.parametrize("t", [R.Q4_0, R.Q8_0])
def test_q4_0_q8_0_prefill_in_column_chunks_of_32(b200, ctx, t):
    """the legacy block formats have no tcgen05 path: a prompt batch goes through gemv_mma in chunks of 32 columns"""
    rng = np.random.default_rng(5 + t)
    N, K, M = 128, 2048, 77
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((M, K)).astype(np.float32)
    got = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    want = R.orc_mul_mat(t, W, x, N, K)
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


@pytest.mark.parametrize("t", [R.Q4_K, R.Q6_K, R.Q8_0])
@pytest.mark.parametrize("b_ne1", [1, 2])
def test_mul_mat_id_device_side_routing(b200, ctx, t, b_ne1):
    """MoE: experts chosen on the device from `ids` (a strided view of the argsort output, as build_moe_ffn makes it)"""
    rng = np.random.default_rng(11 + t + b_ne1)
    n_expert, n_used, n_tok, N, K = 8, 2, 3, 160, 1024
    As = rand_quant_rows(t, n_expert * N, K, rng)
    b = rng.standard_normal((n_tok, b_ne1, K)).astype(np.float32)
    sel = np.stack([rng.permutation(n_expert) for _ in range(n_tok)]).astype(np.int32)      # [n_tok, n_expert] "argsort" rows
    ids = np.ascontiguousarray(sel[:, :n_used])
    rb = R.row_size(t, K)
    Ad = dev_bytes(As.size + 256, 0)
    Ad[:As.size] = to_dev(As)
    bd, seld = to_dev(b), to_dev(sel)
    out = dev_bytes(n_tok * n_used * N * 4, 0xFF)
    op = b200.make_op(b200.OP_MUL_MAT_ID, b200.tensor(out.data_ptr(), b200.F32, [N, n_used, n_tok]),
                      [b200.tensor(Ad.data_ptr(), t, [K, N, n_expert], flags=1), b200.tensor(bd.data_ptr(), b200.F32, [K, b_ne1, n_tok]),
                       b200.tensor(seld.data_ptr(), b200.I32, [n_used, n_tok], [4, n_expert * 4, n_expert * 4 * n_tok, n_expert * 4 * n_tok])])
    assert b200.supports(op)
    ctx.compute_op(op)
    ctx.sync()
    got = out.cpu().numpy().view(np.float32).reshape(n_tok, n_used, N)
    want = R.orc_mul_mat_id(t, As, b, ids, N, K, n_expert)
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


@pytest.mark.parametrize("t", [R.Q4_K, R.Q6_K, R.Q5_K, R.Q4_0, R.Q8_0])
@pytest.mark.parametrize("b_ne1,n_tok,K", [(1, 40, 1024), (2, 40, 1024), (1, 97, 4352), (2, 5, 2304)])
def test_mul_mat_id_grouped_by_expert_on_device(b200, ctx, t, b_ne1, n_tok, K):
    """MoE batches: the (token, slot) pairs are counting-sorted by expert on the device and every expert's matrix is streamed once per
    32 pairs by the small-batch mma kernel (no host copy of `ids`); unbalanced routing, an idle expert, split-K ranges"""
    rng = np.random.default_rng(21 + t + b_ne1 + n_tok)
    n_expert, n_used, N = 8, 2, 160
    As = rand_quant_rows(t, n_expert * N, K, rng)
    b = rng.standard_normal((n_tok, b_ne1, K)).astype(np.float32)
    w = np.array([8, 4, 2, 1, 1, 0.5, 0.5, 0.0])                 # expert 7 is never chosen, expert 0 gets more than 32 pairs
    sel = np.stack([np.concatenate([c := rng.choice(n_expert, n_used, replace=False, p=w / w.sum()), np.setdiff1d(np.arange(n_expert), c)])
                    for _ in range(n_tok)]).astype(np.int32)
    ids = np.ascontiguousarray(sel[:, :n_used])
    Ad = dev_bytes(As.size + 256, 0)
    Ad[:As.size] = to_dev(As)
    bd, seld = to_dev(b), to_dev(sel)
    out = dev_bytes(n_tok * n_used * N * 4, 0xFF)
    op = b200.make_op(b200.OP_MUL_MAT_ID, b200.tensor(out.data_ptr(), b200.F32, [N, n_used, n_tok]),
                      [b200.tensor(Ad.data_ptr(), t, [K, N, n_expert], flags=1), b200.tensor(bd.data_ptr(), b200.F32, [K, b_ne1, n_tok]),
                       b200.tensor(seld.data_ptr(), b200.I32, [n_used, n_tok], [4, n_expert * 4, n_expert * 4 * n_tok, n_expert * 4 * n_tok])])
    assert b200.supports(op)
    l0 = ctx.launches()
    ctx.compute_op(op)
    ctx.sync()
    assert ctx.launches() - l0 <= 4, "quantise + group + grouped mma (+ split-K reduce), independent of the number of tokens"
    got = out.cpu().numpy().view(np.float32).reshape(n_tok, n_used, N)
    want = R.orc_mul_mat_id(t, As, b, ids, N, K, n_expert)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


@pytest.mark.parametrize("t", [R.Q4_K, R.Q6_K, R.Q5_K, R.Q4_0, R.Q8_0])
@pytest.mark.parametrize("b_ne1,n_tok,K", [(1, 200, 1024), (2, 160, 2304)])
def test_mul_mat_id_prompt_batch_grouped_gemm(b200, ctx, t, b_ne1, n_tok, K):
    """MoE prompt batches (>= 32 pairs per expert on average): the pairs of an expert form token tiles of 128 for the tile GEMMs in grouped
    mode (K-quants: gemm_tc.cu on tcgen05, Q4_0 / Q8_0: gemm_mma.cu); unbalanced routing -> one expert with several tiles, experts with a
    ragged or no tile"""
    rng = np.random.default_rng(31 + t + b_ne1 + n_tok)
    n_expert, n_used, N = 8, 2, 256
    As = rand_quant_rows(t, n_expert * N, K, rng)
    b = rng.standard_normal((n_tok, b_ne1, K)).astype(np.float32)
    w = np.array([10, 4, 2, 1, 1, 0.5, 0.5, 0.0])
    sel = np.stack([np.concatenate([c := rng.choice(n_expert, n_used, replace=False, p=w / w.sum()), np.setdiff1d(np.arange(n_expert), c)])
                    for _ in range(n_tok)]).astype(np.int32)
    ids = np.ascontiguousarray(sel[:, :n_used])
    Ad = dev_bytes(As.size + 256, 0)
    Ad[:As.size] = to_dev(As)
    bd, seld = to_dev(b), to_dev(sel)
    out = dev_bytes(n_tok * n_used * N * 4, 0xFF)
    op = b200.make_op(b200.OP_MUL_MAT_ID, b200.tensor(out.data_ptr(), b200.F32, [N, n_used, n_tok]),
                      [b200.tensor(Ad.data_ptr(), t, [K, N, n_expert], flags=1), b200.tensor(bd.data_ptr(), b200.F32, [K, b_ne1, n_tok]),
                       b200.tensor(seld.data_ptr(), b200.I32, [n_used, n_tok], [4, n_expert * 4, n_expert * 4 * n_tok, n_expert * 4 * n_tok])])
    assert b200.supports(op)
    l0 = ctx.launches()
    ctx.compute_op(op)
    ctx.sync()
    assert ctx.launches() - l0 == 3, "group + pack + grouped GEMM"
    got = out.cpu().numpy().view(np.float32).reshape(n_tok, n_used, N)
    want = R.orc_mul_mat_id(t, As, b, ids, N, K, n_expert)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


# ----------------------------------------------------------------------------- cpu-exact mode (exact.cu)
@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_cpu_exact_mode_bit_identical_to_reference_golden(b200, ctx, t):
    """option cpu_exact: float sums in the SIMD order of the reference build -> the committed reference outputs, bit for bit"""
    ctx.set_option("cpu_exact", 1)
    try:
        for K in (256, 1280):
            name = "%s_%d" % (R.TYPE_NAMES[t], K)
            got = gpu_mul_mat(b200, ctx, t, G["mm_W_" + name], G["mm_x_%d" % K], 12, K)
            assert np.array_equal(got, G["mm_out_" + name]), (name, np.abs(got - G["mm_out_" + name]).max())
    finally:
        ctx.set_option("cpu_exact", 0)


@pytest.mark.parametrize("t", R.QUANT_TYPES)
@pytest.mark.parametrize("N,K,M", [(64, 4096, 1), (24, 14336, 3), (40, 2048, 33)])
def test_cpu_exact_mode_bit_identical_to_oracle(b200, ctx, t, N, K, M):
    """real decode / prefill shapes against the oracle's vec_dot (itself bit-identical to the reference, tests/test_oracle_pin.py)"""
    rng = np.random.default_rng(N + K + M + t)
    W = rand_quant_rows(t, N, K, rng)
    x = (rng.standard_normal((M, K)) * rng.uniform(0.2, 5, (M, 1))).astype(np.float32)
    want = R.orc_mul_mat(t, W, x, N, K)
    fast = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    ctx.set_option("cpu_exact", 1)
    try:
        got = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    finally:
        ctx.set_option("cpu_exact", 0)
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert np.abs(fast - want).max() <= 3e-6 * np.abs(want).max()          # the fast path: same integers, its own float order
