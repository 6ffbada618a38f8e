"""CPU tests: the GEMV per-lane item decoders (cortex.llamacpp_b200/csrc/gemv_items.cuh), compiled for the host, must
reproduce the oracle's exact integer block sums for every format, alignment phase and ragged K."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import reflib as R
from util import rand_quant_rows, scratch_from_blocks

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host_emul", "libgemv_host.so")


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "host_emul", "gemv_host.cpp")
    hdr = os.path.join(R.ROOT, "cortex.llamacpp_b200", "csrc", "gemv_items.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, src])
    return C.CDLL(SO)


@pytest.mark.parametrize("K", [256, 512, 2048, 4096, 5632])
@pytest.mark.parametrize("t", R.QUANT_TYPES)
def test_item_decoders_integer_exact(emul, t, K):
    rng = np.random.default_rng(K + t)
    N = 3
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((1, K)).astype(np.float32)
    ta = R.act_type(t)
    act = R.orc_quantize_act(ta, x)[0]
    sc = scratch_from_blocks(ta, act, K)
    rb = R.row_size(t, K)
    nb = K // R.BLOCK[t][0]
    Wp = np.concatenate([W, np.zeros(256, np.uint8)])
    want = R.orc_mul_mat(t, W, x, N, K)[0]
    for phase in (0, 2):
        dst = np.zeros(N, np.float32)
        P = np.zeros((N, nb), np.int32)
        M = np.zeros((N, nb), np.int32)
        rc = emul.emul_gemv(t, Wp.ctypes.data_as(C.c_void_p), C.c_size_t(rb), N, K, sc.ctypes.data_as(C.c_void_p),
                            dst.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p), M.ctypes.data_as(C.c_void_p), phase)
        assert rc == 0
        for n in range(N):
            p, m = R.orc_block_sums(t, W.reshape(N, rb)[n], act, K)
            assert np.array_equal(p, P[n]) and np.array_equal(m, M[n])
        assert np.abs(dst - want).max() <= 3e-6 * np.abs(want).max()


@pytest.mark.parametrize("t", [R.Q4_0, R.Q8_0])
@pytest.mark.parametrize("K", [32, 96, 160, 288])
def test_ragged_k_legacy_quants(emul, t, K):
    """K not a multiple of the 128-element item: the tail item must only use its valid blocks"""
    rng = np.random.default_rng(K * 7 + t)
    N = 2
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((1, K)).astype(np.float32)
    act = R.orc_quantize_act(R.Q8_0, x)[0]
    sc = scratch_from_blocks(R.Q8_0, act, K)
    rb = R.row_size(t, K)
    nb = K // 32
    Wp = np.concatenate([W, np.full(512, 0xAB, np.uint8)])
    dst = np.zeros(N, np.float32)
    P = np.zeros((N, nb + 4), np.int32)
    M = np.zeros((N, nb + 4), np.int32)
    emul.emul_gemv(t, Wp.ctypes.data_as(C.c_void_p), C.c_size_t(rb), N, K, sc.ctypes.data_as(C.c_void_p),
                   dst.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p), M.ctypes.data_as(C.c_void_p), 0)
    want = R.orc_mul_mat(t, W, x, N, K)[0]
    assert np.abs(dst - want).max() <= 3e-6 * max(np.abs(want).max(), 1e-6)


BS1_SO = os.path.join(HERE, "host_emul", "libbs1_host.so")


@pytest.fixture(scope="module")
def emul_bs1():
    src = os.path.join(HERE, "host_emul", "bs1_host.cpp")
    hdrs = [os.path.join(R.ROOT, "cortex.llamacpp_b200", "csrc", h) for h in ("gemv_bs1_items.cuh", "gemv_items.cuh")]
    if not os.path.exists(BS1_SO) or os.path.getmtime(BS1_SO) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", BS1_SO, src])
    return C.CDLL(BS1_SO)


@pytest.mark.parametrize("K", [256, 2048, 4096, 14336])
@pytest.mark.parametrize("t", [R.Q4_K, R.Q5_K, R.Q6_K])
def test_bs1_block_decoders_integer_exact(emul_bs1, t, K):
    """the headline kernel's decoders (gemv_bs1_items.cuh) on the CPU: per-block integers P (and M) identical to the oracle's
    ggml_vec_dot_q*_K_q8_K integer stage, for both 2-byte alignment phases of a Q6_K row; float result within summation order"""
    rng = np.random.default_rng(K * 3 + t)
    N = 3
    W = rand_quant_rows(t, N, K, rng)
    x = (rng.standard_normal((1, K)) * 3.0).astype(np.float32)
    act = R.orc_quantize_act(R.Q8_K, x)[0]                      # block_q8_K: float d; int8 qs[256]; int16 bsums[16]  (292 B)
    blocks = np.frombuffer(act.tobytes(), dtype=np.uint8).reshape(K // 256, 292)
    d = np.ascontiguousarray(blocks[:, 0:4]).view(np.float32).reshape(-1).copy()
    q = np.ascontiguousarray(blocks[:, 4:260]).view(np.int8).reshape(-1).copy()
    bs = np.ascontiguousarray(blocks[:, 260:292]).view(np.int16).reshape(-1).copy()
    rb = R.row_size(t, K)
    nb = K // 256
    Wp = np.concatenate([W, np.zeros(256, np.uint8)])
    want = R.orc_mul_mat(t, W, x, N, K)[0]
    for phase in ((0, 2) if t == R.Q6_K else (0,)):
        dst = np.zeros(N, np.float32)
        P = np.zeros((N, nb), np.int32)
        M = np.zeros((N, nb), np.int32)
        rc = emul_bs1.emul_bs1(t, Wp.ctypes.data_as(C.c_void_p), C.c_size_t(rb), N, K, q.ctypes.data_as(C.c_void_p),
                               d.ctypes.data_as(C.c_void_p), bs.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p),
                               P.ctypes.data_as(C.c_void_p), M.ctypes.data_as(C.c_void_p), phase)
        assert rc == 0
        for n in range(N):
            p, m = R.orc_block_sums(t, W.reshape(N, rb)[n], act, K)
            assert np.array_equal(p, P[n]), (t, K, phase, n)
            if t != R.Q6_K:
                assert np.array_equal(m, M[n]), (t, K, phase, n)
        assert np.abs(dst - want).max() <= 3e-6 * np.abs(want).max()
