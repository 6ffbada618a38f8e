"""Generates tests/golden/*.npz by running THE REFERENCE ITSELF (oracle/_ref/libggml-{base,cpu}.so, built from
/root/reference/llama.cpp by oracle/Makefile) on seeded inputs.  The reference ships no golden vectors for the hot
path, so these fixtures are what pins the oracle and the CUDA kernels when the reference libraries are
not around.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reflib as R  # noqa: E402


def main():
    r = R.ref()
    rng = np.random.default_rng(20260117)
    out = {}
    # ---- activation quantisers: edge cases included (zero block, +/- tie on max, tiny and huge values)
    K = 1024
    x = rng.standard_normal((6, K)).astype(np.float32)
    x[1, 256:512] = 0.0
    x[2, 0] = 3.5; x[2, 1] = -3.5; x[2, 2:32] *= 0.1            # tie on |max| inside a block: first wins
    x[3] *= 1e-6
    x[4] *= 1e4
    x[5, 32:64] = 0.0
    out["act_x"] = x
    out["act_q8_0"] = r.quantize_act(R.Q8_0, x)
    out["act_q8_K"] = r.quantize_act(R.Q8_K, x)
    out["act_q4_0"] = r.quantize_act(R.Q4_0, x)
    # ---- weights through the reference quantiser + reference matmul (from_float + vec_dot per element)
    N = 12
    for K in (256, 1280):
        w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
        xs = rng.standard_normal((3, K)).astype(np.float32)
        out["mm_x_%d" % K] = xs
        for t in R.QUANT_TYPES:
            W = r.quantize_weights(t, w)
            name = "%s_%d" % (R.TYPE_NAMES[t], K)
            out["mm_W_" + name] = W
            out["mm_deq_" + name] = r.dequantize(t, W, K)
            out["mm_out_" + name] = r.mul_mat(t, W, xs, N, K)
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "hotpath_golden.npz"), {k: v.shape for k, v in out.items() if k.startswith("act")})


if __name__ == "__main__":
    main()
