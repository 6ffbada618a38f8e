"""CPU tests: the C oracle (oracle/oracle.c) against (a) golden vectors produced by the reference itself and
(b) the reference libraries directly when oracle/_ref is present.  Bit-exact for every integer/byte result."""
import os

import numpy as np
import pytest

import reflib as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath_golden.npz"))


@pytest.mark.parametrize("ta", [R.Q8_0, R.Q8_K, R.Q4_0])
def test_act_quantisers_match_golden(ta):
    got = R.orc_quantize_act(ta, G["act_x"])
    want = G["act_" + R.TYPE_NAMES[ta]]
    assert got.shape == want.shape
    assert np.array_equal(got, want), "quantiser %s differs from the reference bytes" % R.TYPE_NAMES[ta]


@pytest.mark.parametrize("K", [256, 1280])
@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_dequant_and_matmul_match_golden(t, K):
    name = "%s_%d" % (R.TYPE_NAMES[t], K)
    W, x = G["mm_W_" + name], G["mm_x_%d" % K]
    assert np.array_equal(R.orc_dequantize(t, W, K), G["mm_deq_" + name])
    out = R.orc_mul_mat(t, W, x, 12, K)
    want = G["mm_out_" + name]
    # only the f32 summation order differs from the reference's AVX2 build
    assert np.abs(out - want).max() <= 2e-6 * np.abs(want).max()


def test_fp16_roundtrip_exhaustive():
    o = R.oracle()
    h = np.arange(65536, dtype=np.uint16)
    f = h.view(np.float16).astype(np.float32)
    for v in list(range(0, 65536, 97)) + [0, 1, 0x3ff, 0x400, 0x7bff, 0x7c00, 0x8000, 0xfbff]:
        got = o.orc_f16_to_f32(v)
        if np.isnan(f[v]):
            assert np.isnan(got)
        else:
            assert got == f[v]
            assert o.orc_f32_to_f16(float(f[v])) == v or f[v] == 0
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-8, 5, 2000), [65504.0, 65519.9, 65520.0, 1e-8, 6e-8]]).astype(np.float32)
    want = xs.astype(np.float16).view(np.uint16)
    got = np.array([o.orc_f32_to_f16(float(v)) for v in xs], np.uint16)
    assert np.array_equal(got, want)


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_block_sums_consistent_with_reference_vec_dot(t):
    """exact integer sums recombined in float must reproduce the reference vec_dot to float-order noise"""
    rng = np.random.default_rng(11 + t)
    r = R.ref()
    K, N = 2048, 4
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
    x = rng.standard_normal((1, K)).astype(np.float32)
    W = r.quantize_weights(t, w).reshape(N, -1)
    ta = R.act_type(t)
    a_ref = r.quantize_act(ta, x)
    a_orc = R.orc_quantize_act(ta, x)
    assert np.array_equal(a_ref, a_orc)
    for n in range(N):
        want = r.vec_dot(t, W[n], a_ref[0], K)
        got = R.oracle().orc_vec_dot(t, W[n].ctypes.data, a_orc[0].ctypes.data, K)
        assert abs(got - want) <= 2e-6 * max(1.0, abs(want))


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
def test_reference_own_quantize_tests_pass():
    """the reference's test-quantize-fns (dot error <= 0.02 etc.) passes on the oracle build"""
    import subprocess
    exe = os.path.join(R.REF_DIR, "test-quantize-fns")
    if not os.path.exists(exe):
        pytest.skip("tool not built")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:]
