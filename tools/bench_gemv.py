"""Micro-benchmark of the decode GEMV through the C ABI: GB/s of algorithmic bytes, CUDA events on the library's stream.
Usage: python tools/bench_gemv.py [--types q4_K,q6_K] [--cols 1]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package  # noqa: E402
from util import dev_bytes, rand_quant_rows, to_dev  # noqa: E402
import reflib as R  # noqa: E402

NAMES = {"q4_0": R.Q4_0, "q8_0": R.Q8_0, "q4_K": R.Q4_K, "q5_K": R.Q5_K, "q6_K": R.Q6_K}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--types", default="q4_K,q6_K,q5_K,q4_0,q8_0")
    ap.add_argument("--cols", default="1")
    ap.add_argument("--shapes", default="4096x4096,1024x4096,14336x4096,4096x14336,128256x4096,28672x8192,8192x28672")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--pdl", type=int, default=0)
    a = ap.parse_args()
    import torch
    b200 = load_package()
    ctx = b200.Context(0)
    ctx.set_option("pdl", a.pdl)
    L = b200.lib()
    rng = np.random.default_rng(0)
    for shp in a.shapes.split(","):
        N, K = [int(v) for v in shp.split("x")]
        for tn in a.types.split(","):
            t = NAMES[tn]
            rb = R.row_size(t, K)
            # a few distinct random rows tiled: content does not matter for speed, allocation does
            base = rand_quant_rows(t, 64, K, rng)
            # rotate over enough distinct copies of the weights that consecutive launches never hit L2 (126 MB)
            ncopy = max(2, int((300 << 20) // (N * rb)) + 1)
            tile = to_dev(base).repeat((N + 63) // 64)[:N * rb]
            Wds = []
            for _ in range(ncopy):
                Wd = dev_bytes(N * rb + 256, 0)
                Wd[:N * rb] = tile
                Wds.append(Wd)
            for M in [int(v) for v in a.cols.split(",")]:
                xd = to_dev(rng.standard_normal((M, K)).astype(np.float32))
                out = dev_bytes(M * N * 4)
                ops = [b200.make_op(b200.OP_MUL_MAT, b200.tensor(out.data_ptr(), b200.F32, [N, M]),
                                    [b200.tensor(Wd.data_ptr(), t, [K, N], flags=1), b200.tensor(xd.data_ptr(), b200.F32, [K, M])]) for Wd in Wds]
                for i in range(3):
                    ctx.compute_op(ops[i % ncopy])
                ctx.sync()
                e0, e1 = L.b200_event_create(0), L.b200_event_create(0)
                ts = []
                for i in range(a.iters):
                    L.b200_event_record(ctx.h, e0)
                    ctx.compute_op(ops[i % ncopy])
                    L.b200_event_record(ctx.h, e1)
                    L.b200_event_synchronize(e1)
                    ts.append(L.b200_event_elapsed_ms(e0, e1))
                # back-to-back inside ONE captured CUDA graph (what a decode step looks like): total time / count
                ctx.set_option("cuda_graphs", 1)
                ctx.set_option("fusion", 0)
                lst = [ops[i % ncopy] for i in range(max(a.iters, 8))]
                for _ in range(3):
                    ctx.compute(lst)
                ctx.sync()
                L.b200_event_record(ctx.h, e0)
                ctx.compute(lst)
                L.b200_event_record(ctx.h, e1)
                L.b200_event_synchronize(e1)
                b2b = L.b200_event_elapsed_ms(e0, e1) / len(lst)
                ctx.set_option("cuda_graphs", 0)
                ms = float(np.median(ts))
                bytes_ = N * rb + M * K * 4 + M * N * 4
                print("%-5s N=%6d K=%6d M=%d  %8.2f us  %7.1f GB/s | back-to-back %8.2f us %7.1f GB/s" % (tn, N, K, M, ms * 1e3, bytes_ / ms / 1e6, b2b * 1e3, bytes_ / b2b / 1e6), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
