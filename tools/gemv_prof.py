"""Where does a GEMV launch spend its time?  Per-CTA %globaltimer stamps from inside the kernel (debug hook)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package
from util import dev_bytes, rand_quant_rows, to_dev
import reflib as R
import torch

b200 = load_package(); ctx = b200.Context(0); L = b200.lib()
if len(sys.argv) > 3: ctx.set_option("pdl", int(sys.argv[3]))
N, K = int(sys.argv[1]), int(sys.argv[2]); t = R.Q4_K
rng = np.random.default_rng(0)
rb = R.row_size(t, K)
tile = to_dev(rand_quant_rows(t, 64, K, rng)).repeat((N + 63) // 64)[:N * rb]
Wds = []
for _ in range(max(2, (300 << 20) // (N * rb) + 1)):
    Wd = dev_bytes(N * rb + 256, 0); Wd[:N * rb] = tile; Wds.append(Wd)
xd = to_dev(rng.standard_normal((1, K)).astype(np.float32)); out = dev_bytes(N * 4)
ops = [b200.make_op(b200.OP_MUL_MAT, b200.tensor(out.data_ptr(), b200.F32, [N, 1]),
                    [b200.tensor(W.data_ptr(), t, [K, N], flags=1), b200.tensor(xd.data_ptr(), b200.F32, [K, 1])]) for W in Wds]
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
for i in range(4): ctx.compute_op(ops[i % len(ops)])
ctx.sync()
L.b200_debug_set_prof(ctx.h, prof.data_ptr())
ctx.compute_op(ops[4 % len(ops)]); ctx.sync()
p = prof.cpu().numpy().reshape(148, 16).astype(np.int64)
t0 = p[:, 0][p[:, 0] > 0].min()
names = ["cta start", "after init+sync", "first copy issued", "all copies issued", "prologue done", "after prologue bar", "first stage landed", "first chunk done", "last chunk done"]
for i, n in enumerate(names):
    v = p[:, i][p[:, i] > 0] - t0
    if len(v) == 0:
        continue
    print("%-22s min %6d  median %6d  max %6d ns  (n=%d)" % (n, v.min(), np.median(v), v.max(), len(v)))
