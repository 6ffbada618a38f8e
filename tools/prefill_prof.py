"""One prefill ubatch (or a bs32 decode ubatch) of a few layers, for ncu launch lists.  Usage: prefill_prof.py [T] [layers] [mode: prefill|slots]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_llama_graph, load_package
import torch
b200 = load_package(); lg = load_llama_graph(); L = b200.lib(); ctx = b200.Context(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "prefill"
depth = 512
n_ctx = ((T * depth + T + 255) // 256 + 1) * 256 if mode == "slots" else 1024
g = lg.LlamaGraph(b200, model="llama3-8b", ftype="q4_k_m", kv="q8_0", n_ctx=n_ctx, layers=layers, max_tokens=T)
rng = np.random.default_rng(0)
if mode == "slots":
    g.fill_cache(T * depth)
    n_kv = (T * depth + T + 255) // 256 * 256
    emb, pos, mask, kv_head = g.set_inputs_slots_host(T, depth, n_kv, rng)
else:
    n_kv, kv_head = (T + 255) // 256 * 256, 0
    emb, pos, mask = g.set_inputs_host(T, 0, n_kv, rng)
g.inp_embd[:T * g.E] = torch.from_numpy(emb.reshape(-1)).cuda(); g.pos[:T] = torch.from_numpy(pos).cuda(); g.mask_f32[:mask.size] = torch.from_numpy(mask.reshape(-1)).cuda()
ops = g.build(T, kv_head, n_kv)
ctx.set_option("cuda_graphs", 0)
for _ in range(3): ctx.compute(ops)
ctx.sync()
e0, e1 = L.b200_event_create(0), L.b200_event_create(0)
L.b200_event_record(ctx.h, e0); ctx.compute(ops); L.b200_event_record(ctx.h, e1); L.b200_event_synchronize(e1)
print("T=%d layers=%d mode=%s: %.3f ms" % (T, layers, mode, L.b200_event_elapsed_ms(e0, e1)))
