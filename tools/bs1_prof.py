"""Timeline of consecutive batch-1 GEMV launches (gemv_bs1.cu): per-CTA %globaltimer stamps laid on one clock.
Usage: python tools/bs1_prof.py N K [pdl] [graphs] [type]   -- launches 6 matmuls back to back, prints launches 2..4"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package
from util import dev_bytes, rand_quant_rows, to_dev
import reflib as R
import torch

b200 = load_package(); ctx = b200.Context(0); L = b200.lib()
N, K = int(sys.argv[1]), int(sys.argv[2])
pdl = int(sys.argv[3]) if len(sys.argv) > 3 else 1
graphs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
t = {"q4_K": R.Q4_K, "q6_K": R.Q6_K}[sys.argv[5] if len(sys.argv) > 5 else "q4_K"]
ctx.set_option("pdl", pdl); ctx.set_option("fusion", 0)
rng = np.random.default_rng(0)
rb = R.row_size(t, K)
tile = to_dev(rand_quant_rows(t, 64, K, rng)).repeat((N + 63) // 64)[:N * rb]
Wds = []
for _ in range(max(8, (300 << 20) // (N * rb) + 1)):
    Wd = dev_bytes(N * rb + 256, 0); Wd[:N * rb] = tile; Wds.append(Wd)
xd = to_dev(rng.standard_normal((1, K)).astype(np.float32)); outs = [dev_bytes(N * 4) for _ in range(2)]
# a dependent chain like a decode step: every matmul reads x (constant) and writes its own output
ops = [b200.make_op(b200.OP_MUL_MAT, b200.tensor(outs[i % 2].data_ptr(), b200.F32, [N, 1]),
                    [b200.tensor(W.data_ptr(), t, [K, N], flags=1), b200.tensor(xd.data_ptr(), b200.F32, [K, 1])]) for i, W in enumerate(Wds)]
NL = 6
prof = torch.zeros(8 * 296 * 32, dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
ctx.set_option("cuda_graphs", graphs)
lst = ops[:NL]
for _ in range(3): ctx.compute(lst)
ctx.sync()
L.b200_debug_set_prof(ctx.h, prof.data_ptr())
ctx.set_option("cuda_graphs", 0)          # stamps need the prof pointer in the params: run eagerly (PDL still applies on the stream)
ctx.compute(lst); ctx.sync()
p = prof.cpu().numpy().reshape(8, 296, 32).astype(np.int64)
names = {0: "cta start", 1: "producer ready", 2: "first copies issued", 3: "all copies issued", 4: "consumers past pdl_wait", 5: "prologue done",
         6: "first stage landed", 7: "first chunk done"}
t0 = None
for li in range(1, 5):
    q = p[li]
    if t0 is None: t0 = q[:, 0][q[:, 0] > 0].min()
    print("-- launch %d" % li)
    for i in range(8):
        v = q[:, i][q[:, i] > 0] - t0
        if len(v): print("  %-24s min %7d  median %7d  max %7d ns  (n=%d)" % (names[i], v.min(), np.median(v), v.max(), len(v)))
    for i, nm in ((24, "x landed (warp 0)"), (25, "quantised (warp 0)")):
        v = q[:, i][q[:, i] > 0] - t0
        if len(v): print("  %-24s min %7d  median %7d  max %7d ns  (n=%d)" % (nm, v.min(), np.median(v), v.max(), len(v)))
    w = q[:, 8:23]; v = w[w > 0] - t0
    if len(v): print("  %-24s min %7d  median %7d  max %7d ns" % ("consumer warps done", v.min(), np.median(v), v.max()))
