"""One eager bs32 continuous-batching decode step (or a 512-token prefill ubatch) for an ncu launch list.
Usage: python tools/batched_prof.py bs32|pp512 [layers] [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_llama_graph, load_package
import torch
b200 = load_package(); lg = load_llama_graph(); L = b200.lib(); ctx = b200.Context(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "bs32"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
slots, depth, PP = 32, 512, 512
n_ctx = ((slots * depth + slots + 255) // 256 + 1) * 256
g = lg.LlamaGraph(b200, model="llama3-8b", ftype="q4_k_m", kv="q8_0", n_ctx=n_ctx, layers=layers, max_tokens=PP)
g.fill_cache(slots * depth)
rng = np.random.default_rng(1)
def upload(emb, pos, mask, T):
    g.inp_embd[:T * g.E] = torch.from_numpy(emb.reshape(-1)).cuda(); g.pos[:T] = torch.from_numpy(pos).cuda()
    g.mask_f32[:mask.size] = torch.from_numpy(mask.reshape(-1)).cuda(); torch.cuda.synchronize()
if mode == "bs32":
    n_kv = (slots * depth + slots + 255) // 256 * 256
    emb, pos, mask, kv_head = g.set_inputs_slots_host(slots, depth, n_kv, rng)
    upload(emb, pos, mask, slots)
    ops = g.build(slots, kv_head, n_kv)
else:
    emb, pos, mask = g.set_inputs_host(PP, 0, PP, rng)
    upload(emb, pos, mask, PP)
    ops = g.build(PP, 0, PP, n_outputs=1)
ctx.set_option("pdl", 1); ctx.set_option("fusion", 2); ctx.set_option("cuda_graphs", 0)
e0, e1 = L.b200_event_create(0), L.b200_event_create(0)
ctx.compute(ops); ctx.sync()
L.b200_event_record(ctx.h, e0)
for _ in range(reps): ctx.compute(ops)
L.b200_event_record(ctx.h, e1); L.b200_event_synchronize(e1)
print("%s layers=%d: %.3f ms per step (eager)" % (mode, layers, L.b200_event_elapsed_ms(e0, e1) / reps))
