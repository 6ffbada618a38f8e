"""BASELINE.json configs #1, #2 and #4 through the C ABI (random-init GGUF blocks of the named shapes): bs1 decode tok/s at depth 512 with
the HBM roofline fraction, and a 512-token prefill ubatch.  Extra numbers beside bench.py's headline (config #3 shape); one JSON line per config.
Usage: python tools/bench_configs.py [tinyllama:q4_0 llama2-7b:q5_k_m mixtral:q4_k_m] [--depth 512] [--steps 32]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_llama_graph, load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["tinyllama:q4_0", "llama2-7b:q5_k_m", "mixtral:q4_k_m"])
    ap.add_argument("--depth", type=int, default=512)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--pp", type=int, default=512)
    a = ap.parse_args()
    import torch
    b200 = load_package(); lg = load_llama_graph(); L = b200.lib()
    peak = 6561.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for cfg in a.configs:
        model, ftype = cfg.split(":")
        ctx = b200.Context(0)
        ctx.set_option("cuda_graphs", 1); ctx.set_option("pdl", 1); ctx.set_option("fusion", 2)
        n_ctx = (a.depth + a.pp + 255) // 256 * 256 + 256
        g = lg.LlamaGraph(b200, model=model, ftype=ftype, kv="f16", n_ctx=n_ctx, max_tokens=a.pp)
        g.fill_cache(a.depth)
        rng = np.random.default_rng(0)
        e0, e1 = L.b200_event_create(0), L.b200_event_create(0)

        def upload(emb, pos, mask, T):
            g.inp_embd[:T * g.E] = torch.from_numpy(emb.reshape(-1)).cuda(); g.pos[:T] = torch.from_numpy(pos).cuda()
            g.mask_f32[:mask.size] = torch.from_numpy(mask.reshape(-1)).cuda(); torch.cuda.synchronize()

        def timed(ops, reps, warm):
            arr = (b200.Op * len(ops))(*ops)
            n0 = ctx.launches()
            b200.check(L.b200_graph_compute(ctx.h, arr, len(ops)), "step"); ctx.sync()
            launches = ctx.launches() - n0
            for _ in range(warm):
                b200.check(L.b200_graph_compute(ctx.h, arr, len(ops)), "step")
            ctx.sync()
            L.b200_event_record(ctx.h, e0)
            for _ in range(reps):
                b200.check(L.b200_graph_compute(ctx.h, arr, len(ops)), "step")
            L.b200_event_record(ctx.h, e1); L.b200_event_synchronize(e1)
            return L.b200_event_elapsed_ms(e0, e1) / reps, int(launches)
        n_kv = (a.depth + 1 + 255) // 256 * 256
        emb, pos, mask = g.set_inputs_host(1, a.depth, n_kv, rng)
        upload(emb, pos, mask, 1)
        ms, launches = timed(g.build(1, a.depth, n_kv), a.steps, 4)
        sb = g.step_bytes(1, n_kv)
        if g.n_expert:          # a token touches n_used of the n_expert expert matrices
            sb["weights"] -= g.expert_bytes * (1.0 - g.n_used / g.n_expert)
            sb["total"] = sb["weights"] + sb["kv_read"] + sb["kv_write"] + sb["act"]
        out = {"config": "%s %s bs1 decode, f16 KV, depth %d" % (model, ftype, a.depth), "decode_tok_s": 1e3 / ms, "ms_per_step": ms, "gpu_launches_per_step": launches,
               "hbm_bytes_per_step": sb["total"], "achieved_gbs": sb["total"] / ms / 1e6, "hbm_frac": sb["total"] / ms / 1e6 / peak,
               "finite": bool(torch.isfinite(g.logits[:g.V]).all().item())}
        n_kv2 = (a.depth + a.pp + 255) // 256 * 256
        emb, pos, mask = g.set_inputs_host(a.pp, a.depth, n_kv2, rng)
        upload(emb, pos, mask, a.pp)
        ms, launches = timed(g.build(a.pp, a.depth, n_kv2, n_outputs=1), 3, 1)
        out.update({"prefill_tok_s": a.pp / (ms / 1e3), "prefill_ms_per_ubatch": ms, "prefill_launches": launches, "pp": a.pp})
        print(json.dumps(out), flush=True)
        g.keep.clear(); g.layers.clear(); del g
        ctx.close()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
