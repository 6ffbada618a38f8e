"""GPU parity of the tcgen05 int8 prefill GEMM (M > 8 tokens, K-quant weights) against the oracle's mul_mat
(ggml_compute_forward_mul_mat with q8_K activations).  The integer stage is the CPU's exactly, so the results agree to
f32 summation order: 3e-6 relative."""
import numpy as np
import pytest

import reflib as R
from util import dev_bytes, rand_quant_rows, to_dev

pytestmark = pytest.mark.gpu


def gpu_mul_mat(b200, ctx, t, W, x, N, K):
    M = x.shape[0]
    Wd = dev_bytes(W.size + 256, 0)
    Wd[:W.size] = to_dev(W)
    xd = to_dev(x)
    out = dev_bytes(M * N * 4, 0xFF)
    op = b200.make_op(b200.OP_MUL_MAT, b200.tensor(out.data_ptr(), b200.F32, [N, M]),
                      [b200.tensor(Wd.data_ptr(), t, [K, N], flags=1), b200.tensor(xd.data_ptr(), b200.F32, [K, M])])
    assert b200.supports(op)
    ctx.compute_op(op)
    ctx.sync()
    return out.cpu().numpy().view(np.float32).reshape(M, N)


@pytest.mark.parametrize("N,K,M", [(128, 256, 16), (128, 512, 128), (256, 1024, 130), (128, 4096, 9), (384, 2048, 300), (1024, 4096, 512),
                                   (256, 5632, 32), (1024, 14336, 32)])      # the last two: split-K with uneven / 8 slices, a quarter-full token tile
@pytest.mark.parametrize("t", [R.Q4_K, R.Q5_K, R.Q6_K])
def test_gemm_i8_vs_oracle(b200, ctx, t, N, K, M):
    rng = np.random.default_rng(N + K + M + t)
    W = rand_quant_rows(t, N, K, rng)
    x = (rng.standard_normal((M, K)) * rng.uniform(0.1, 4.0, (M, 1))).astype(np.float32)
    got = gpu_mul_mat(b200, ctx, t, W, x, N, K)
    assert np.isfinite(got).all()
    rows = np.unique(np.concatenate([np.arange(min(N, 8)), rng.integers(0, N, 8), [N - 1, N - 128, 127 if N > 127 else 0]]))
    cols = np.unique(np.concatenate([np.arange(min(M, 4)), rng.integers(0, M, 6), [M - 1]]))
    rb = R.row_size(t, K)
    want = R.orc_mul_mat(t, W.reshape(N, rb)[rows].reshape(-1), x[cols], len(rows), K)
    sub = got[np.ix_(cols, rows)]
    scale = max(np.abs(want).max(), 1e-6)
    assert np.abs(sub - want).max() <= 3e-6 * scale, np.abs(sub - want).max() / scale
