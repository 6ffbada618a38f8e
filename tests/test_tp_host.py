"""CPU tests of the tensor-parallel host logic: the Megatron shard plan, the GGUF-block shard extraction, and --
with a world_size-2 gloo group -- that summing the per-rank partial matmuls of K-split shards reproduces the full matmul of
the CPU oracle, and that row shards concatenate back to the full result.  No GPU, no compute through the product library."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lg():
    from __graft_entry__ import load_llama_graph
    return load_llama_graph()


def test_tp_plan_north_star_models():
    G = lg()
    from make_gguf import MODELS
    for model, worlds in (("llama3-70b", (2, 4, 8)), ("llama3-8b", (2, 4, 8)), ("mixtral", (2, 4, 8))):
        if model not in MODELS:
            continue
        L, E, H, Hkv, D, FF = MODELS[model][:6]
        for w in worlds:
            Hl, Hkvl, FFl = G.tp_plan((H, Hkv, D, FF), w)
            assert Hl * w == H and Hkvl * w == Hkv and FFl * w == FF and FFl % 256 == 0 and (Hl * D) % 256 == 0
    # Llama-2-7B: FF = 11008 = 43 * 256 cannot be cut evenly at super-block boundaries
    L, E, H, Hkv, D, FF = MODELS["llama2-7b"][:6]
    with pytest.raises(ValueError):
        G.tp_plan((H, Hkv, D, FF), 2)
    with pytest.raises(ValueError):
        G.tp_plan((4, 2, 128, 1024), 4)          # more ranks than KV heads


@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K", "Q8_0"])
def test_shards_partition_the_tensor(tname):
    import reflib as R
    from util import rand_quant_rows
    G = lg()
    t = getattr(R, tname)
    N, K, world = 8, 1024, 4
    W = rand_quant_rows(t, N, K, np.random.default_rng(0))
    rb = R.row_size(t, K)
    rows = [G.shard_rows(W, t, K, N, r, world) for r in range(world)]
    assert np.array_equal(np.concatenate(rows), W[:N * rb])
    ks = [G.shard_k(W, t, K, N, r, world).reshape(N, -1) for r in range(world)]
    assert np.array_equal(np.concatenate(ks, axis=1).reshape(-1), W[:N * rb])


def _tp_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reflib as R
    from util import rand_quant_rows
    G = lg()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    import torch
    rng = np.random.default_rng(3)                      # same stream on every rank: the FULL tensors are identical
    N, K, FF = 64, 1024, 512
    ok = True
    for t in (R.Q4_K, R.Q6_K):
        Wdown = rand_quant_rows(t, N, K, rng)            # row-parallel (K split)
        Wup = rand_quant_rows(t, FF, K, rng)             # column-parallel (row split)
        x = rng.standard_normal((1, K)).astype(np.float32)
        full = R.orc_mul_mat(t, Wdown, x, N, K)
        kl = K // world
        part = R.orc_mul_mat(t, G.shard_k(Wdown, t, K, N, rank, world), x[:, rank * kl:(rank + 1) * kl], N, kl)
        tot = torch.from_numpy(part.copy())
        dist.all_reduce(tot)                             # the exchange B200_OP_ALLREDUCE performs on the GPU
        ok &= bool(np.abs(tot.numpy() - full).max() <= 1e-5 * np.abs(full).max())
        fullu = R.orc_mul_mat(t, Wup, x, FF, K)
        mine = torch.from_numpy(R.orc_mul_mat(t, G.shard_rows(Wup, t, K, FF, rank, world), x, FF // world, K).copy())
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        ok &= bool(np.array_equal(np.concatenate([p.numpy() for p in parts], axis=1), fullu))     # row shards are bit-exact
    # the blob exchange Context.comm_init relies on
    blobs = [None] * world
    dist.all_gather_object(blobs, bytes([rank]) * 64)
    ok &= blobs == [bytes([r]) * 64 for r in range(world)]
    dist.destroy_process_group()
    q.put((rank, ok))


def test_tp_partial_sums_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_tp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(60)
    assert res == [(0, True), (1, True)]
