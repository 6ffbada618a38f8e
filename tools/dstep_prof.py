"""Per-phase timeline of the persistent decode-step kernel (dstep.cu) on the bench workload: %globaltimer stamps of every CTA at
phase start / prologue done / work done (warp 0) / barrier passed.  usage: python tools/dstep_prof.py [layers] [skip_attn]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_llama_graph, load_package
b200 = load_package(); lg = load_llama_graph(); L = b200.lib(); ctx = b200.Context(0)
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = lg.LlamaGraph(b200, model="llama3-8b", ftype="q4_k_m", kv="f16", n_ctx=1024, layers=layers, max_tokens=1)
g.fill_cache(512)
rng = np.random.default_rng(0)
emb, pos, mask = g.set_inputs_host(1, 512, 768, rng)
g.inp_embd[:g.E] = torch.from_numpy(emb.reshape(-1)).cuda(); g.pos[:1] = torch.from_numpy(pos).cuda(); g.mask_f32[:mask.size] = torch.from_numpy(mask.reshape(-1)).cuda()
torch.cuda.synchronize()
ops = g.build(1, 512, 768)
ctx.set_option("cuda_graphs", 0); ctx.set_option("fusion", 2); ctx.set_option("debug_skip", 3 if skip else 0)
for _ in range(3): ctx.compute(ops)
ctx.sync()
nph = (6 if not skip else 4) * layers + 1
prof = torch.zeros(148 * nph * 8, dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
L.b200_debug_set_prof(ctx.h, prof.data_ptr())
ctx.compute(ops); ctx.sync()
L.b200_debug_set_prof(ctx.h, None)
p = prof.cpu().numpy().reshape(148, nph, 8).astype(np.int64)
t0 = p[:, 0, 0].min()
names = ["qkv", "attn", "combine", "wo", "gate|up", "down"] if not skip else ["qkv", "wo", "gate|up", "down"]
print("phase: start(min..max) | prologue(median) | work(median, max) | barrier wait(median) | phase total   [ns]")
for ph in range(min(nph, 2 * len(names) + 1)):
    st, pr, wk, br = p[:, ph, 0] - t0, p[:, ph, 1] - p[:, ph, 0], p[:, ph, 2] - p[:, ph, 0], p[:, ph, 3] - p[:, ph, 2]
    nm = names[ph % len(names)] if ph < nph - 1 else "output"
    print("%2d %-8s start %7d..%7d | prologue %6d | work %6d max %6d | barrier %6d max %6d | total %6d" % (
        ph, nm, st.min(), st.max(), int(np.median(pr[pr > 0])) if (pr > 0).any() else 0, int(np.median(wk)), wk.max(), int(np.median(br)), br.max(), (p[:, ph, 3].max() - p[:, ph, 0].min())))
for ph in range(nph):
    if ph % len(names) == 1 and not skip and ph < 8:
        a = p[:, ph, :]
        ok = a[:, 4] > 0
        f = lambda i, j: int(np.median(a[ok, i] - a[ok, j]))
        print("   attn phase %d: inputs %d | rope+store %d | walk %d | bar %d | merge+publish %d  [ns, median over %d unit CTAs]" % (ph, f(4, 0), f(5, 4), f(6, 5), f(7, 6), f(2, 7), ok.sum()))
print("kernel total %.1f us for %d phases" % ((p[:, -1, 3].max() - t0) / 1e3, nph))
