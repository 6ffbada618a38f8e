"""In-situ timeline of the fused GEMV launches of a real decode step (eager, PDL on): per-CTA %globaltimer stamps of the
last 8 bs1 launches.  Usage: python tools/step_prof.py [layers]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_llama_graph, load_package
import torch
b200 = load_package(); lg = load_llama_graph(); L = b200.lib(); ctx = b200.Context(0)
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = lg.LlamaGraph(b200, model="llama3-8b", ftype="q4_k_m", kv="f16", n_ctx=1024, layers=layers, max_tokens=1)
g.fill_cache(512)
rng = np.random.default_rng(0)
emb, pos, mask = g.set_inputs_host(1, 512, 768, rng)
g.inp_embd[:g.E] = torch.from_numpy(emb.reshape(-1)).cuda(); g.pos[:1] = torch.from_numpy(pos).cuda(); g.mask_f32[:mask.size] = torch.from_numpy(mask.reshape(-1)).cuda()
ops = g.build(1, 512, 768)
ctx.set_option("pdl", 1); ctx.set_option("fusion", 2); ctx.set_option("cuda_graphs", 0)
for _ in range(3): ctx.compute(ops)
ctx.sync()
prof = torch.zeros(8 * 296 * 32, dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
L.b200_debug_set_prof(ctx.h, prof.data_ptr())
n_gemv = 4 * layers + 1
ctx.compute(ops); ctx.sync()
p = prof.cpu().numpy().reshape(8, 296, 32).astype(np.int64)
names = {0: "cta start", 1: "producer ready", 2: "first copies issued", 3: "all copies issued", 4: "past pdl_wait", 5: "prologue done",
         6: "first stage landed", 7: "first chunk done", 24: "x landed (w0)", 25: "quantised (w0)", 2: "first copies issued", 26: "x in regs (w0)",
         27: "sumsq done (w0)", 28: "norm scale (w0)"}
kinds = ["qkv(norm)", "wo(+res)", "gate|up(norm)", "down(swiglu+res)"]
order = sorted(range(8), key=lambda s: p[s][:, 0][p[s][:, 0] > 0].min() if (p[s][:, 0] > 0).any() else 1 << 62)
first_idx = n_gemv - 8
prev_end = None
for j, sidx in enumerate(order):
    q = p[sidx]
    if not (q[:, 0] > 0).any(): continue
    li = first_idx + j
    kind = "output(norm)" if li == n_gemv - 1 else kinds[li % 4]
    t0 = q[:, 0][q[:, 0] > 0].min()
    w = q[:, 8:8 + 16]; end = w[w > 0].max()
    line = "launch %2d %-18s gap-from-prev-end %6s | " % (li, kind, "-" if prev_end is None else str(t0 - prev_end))
    for i in (2, 4, 26, 27, 28, 24, 25, 5, 6, 7, 3):
        v = q[:, i][q[:, i] > 0] - t0
        if len(v): line += "%s %d | " % (names[i], int(np.median(v)))
    line += "last warp done %d" % (end - t0)
    print(line)
    prev_end = end
