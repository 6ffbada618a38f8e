"""GPU parity of row-split tensor parallelism (needs >= 2 GPUs, skipped otherwise): TP=2 logits of a decode
sequence vs the TP=1 model drawn from the same seed on rank 0, bit-identical logits across ranks, one-shot peer-memory
all-reduce vs NCCL.  The reference pins nothing for multi-GPU ("parity unpinned"); TP=1 is itself pinned against the CPU
oracle in test_gpu_llama_step.py.  Tolerance as there: 1e-2 of the largest logit on the 2048-wide model, identical greedy tokens."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, oneshot, graphs, q):
    try:
        os.environ["GGML_B200_ONESHOT"] = str(oneshot)
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        from __graft_entry__ import load_llama_graph, load_package
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        b200 = load_package()
        lg = load_llama_graph()
        ctx = b200.Context(rank)

        def exchange(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out
        ctx.comm_init(rank, world, exchange)
        ctx.set_option("cuda_graphs", graphs)
        ctx.set_option("pdl", 1)
        model = "mid-d128"
        g = lg.LlamaGraph(b200, model=model, ftype="q4_k_m", kv="q8_0", n_ctx=256, max_tokens=2, device=rank, tp_rank=rank, tp_world=world)
        ref = lg.LlamaGraph(b200, model=model, ftype="q4_k_m", kv="q8_0", n_ctx=256, max_tokens=2, device=rank) if rank == 0 else None
        ctx1 = b200.Context(rank) if rank == 0 else None
        rng = np.random.default_rng(11)
        worst, same_tok, same_ranks = 0.0, True, True
        n_kv = 256
        schedule = [(2, 0), (1, 2), (1, 3), (1, 3), (1, 3), (2, 4)]         # repeated op lists exercise CUDA-graph capture + replay
        for T, kv_head in schedule:
            emb, pos, mask = g.set_inputs_host(T, kv_head, n_kv, rng)
            for m in (g, ref):
                if m is None:
                    continue
                m.inp_embd[:T * m.E] = torch.from_numpy(emb.reshape(-1)).cuda(rank)
                m.pos[:T] = torch.from_numpy(pos).cuda(rank)
                m.mask_f32[:mask.size] = torch.from_numpy(mask.reshape(-1)).cuda(rank)
            torch.cuda.synchronize()
            ctx.compute(g.build(T, kv_head, n_kv))
            ctx.sync()
            got = g.logits[:T * g.V].cpu()
            both = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(both, got)
            same_ranks &= all(torch.equal(both[0], b) for b in both)
            if rank == 0:
                ctx1.compute(ref.build(T, kv_head, n_kv))
                ctx1.sync()
                want = ref.logits[:T * ref.V].cpu().numpy().reshape(T, ref.V)
                gotn = got.numpy().reshape(T, g.V)
                worst = max(worst, float(np.abs(gotn - want).max() / np.abs(want).max()))
                same_tok &= bool((gotn.argmax(1) == want.argmax(1)).all())
        n_launch = ctx.launches()
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
        q.put((rank, "ok", worst, same_tok, same_ranks, n_launch))
    except Exception as ex:      # surface the failure in the parent instead of a hang
        import traceback
        q.put((rank, "error: %s\n%s" % (ex, traceback.format_exc()), 0.0, False, False, 0))


@pytest.mark.parametrize("oneshot,graphs", [(1, 0), (0, 0), (1, 1)])
def test_tp2_matches_tp1(oneshot, graphs):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29600 + (os.getpid() + 7 * oneshot + 3 * graphs) % 2000
    ps = [mpc.Process(target=_worker, args=(r, 2, port, oneshot, graphs, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(60)
    for r in res:
        assert r[1] == "ok", r[1]
        assert r[4], "logits differ between ranks (the all-reduce must be bit-identical on every rank)"
    assert res[0][2] <= 1e-2, "TP=2 vs TP=1 logits: rel %.3g" % res[0][2]
    assert res[0][3], "greedy tokens differ between TP=2 and TP=1"


def _lost_peer_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        from __graft_entry__ import load_package
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        b200 = load_package()
        ctx = b200.Context(rank)

        def exchange(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out
        ctx.comm_init(rank, world, exchange)
        x = torch.ones(4096, dtype=torch.float32, device="cuda")
        y = torch.zeros_like(x)
        op = b200.make_op(b200.OP_ALLREDUCE, b200.tensor(y.data_ptr(), b200.F32, [4096]), [b200.tensor(x.data_ptr(), b200.F32, [4096])])
        ctx.compute_op(op); ctx.sync()                      # a healthy exchange first
        ok_first = bool((y == world).all().item())
        dist.barrier()
        status = "ok"
        if rank == 0:                                       # rank 1 never joins the second all-reduce
            ctx.compute_op(op)
            try:
                ctx.sync()
                status = "no error raised"
            except b200.B200Error as ex:
                status = "failed as expected: %s" % ex
        dist.barrier()
        q.put((rank, status, ok_first))
        dist.destroy_process_group()
    except Exception as ex:
        import traceback
        q.put((rank, "error: %s\n%s" % (ex, traceback.format_exc()), False))


def test_lost_peer_is_reported_not_summed():
    """an all-reduce whose peer never arrives gives up after its bounded spin and the next b200_synchronize returns B200_ERR_FAILED
    (it used to return success with garbage sums)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29600 + (os.getpid() + 911) % 2000
    ps = [mpc.Process(target=_lost_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(60)
    assert res[0][2] and res[1][2], res
    assert res[0][1].startswith("failed as expected") and "timed out" in res[0][1], res[0][1]
    assert res[1][1] == "ok", res[1][1]


@pytest.mark.parametrize("M", [1, 3, 17, 64])
def test_row_split_mul_mat_c_abi(M):
    """B200_TENSOR_FLAG_SPLIT (the op the plugin's split buffer type emits): a weight cut by rows over two devices; the matmul runs on
    both GPUs from ONE host thread, activations and results stay in the main device's memory (NVLink peer access).  The result must be
    bit-identical to the single-GPU matmul: every dst element is produced by the same kernel."""
    import ctypes as C
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from __graft_entry__ import load_package
    import reflib as R
    from util import rand_quant_rows
    b200 = load_package()
    rng = np.random.default_rng(70 + M)
    N, K, t = 1024 + 256, 4096, R.Q4_K
    rb = R.row_size(t, K)
    W = rand_quant_rows(t, N, K, rng)
    x = rng.standard_normal((M, K)).astype(np.float32)
    ctx = b200.Context(0)

    class Split(C.Structure):
        _fields_ = [("n_dev", C.c_int32), ("device", C.c_int32 * 16), ("shard", C.c_void_p * 16), ("row_low", C.c_int64 * 17)]
    cut = 768
    Wt = torch.from_numpy(W)
    pad = torch.zeros(512, dtype=torch.uint8)
    s0 = torch.cat([Wt[:cut * rb], pad]).cuda(0)
    s1 = torch.cat([Wt[cut * rb:], pad]).cuda(1)
    full = torch.cat([Wt, pad]).cuda(0)
    sp = Split()
    sp.n_dev = 2
    sp.device[0], sp.device[1] = 0, 1
    sp.shard[0], sp.shard[1] = s0.data_ptr(), s1.data_ptr()
    sp.row_low[0], sp.row_low[1], sp.row_low[2] = 0, cut, N
    xd = torch.from_numpy(x).cuda(0)
    out_split = torch.zeros(M * N, dtype=torch.float32, device="cuda:0")
    out_one = torch.zeros(M * N, dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    wsplit = b200.tensor(C.addressof(sp), t, [K, N], flags=b200.TENSOR_FLAG_WEIGHT | 2)
    op_s = b200.make_op(b200.OP_MUL_MAT, b200.tensor(out_split.data_ptr(), b200.F32, [N, M]), [wsplit, b200.tensor(xd.data_ptr(), b200.F32, [K, M])])
    op_1 = b200.make_op(b200.OP_MUL_MAT, b200.tensor(out_one.data_ptr(), b200.F32, [N, M]),
                        [b200.tensor(full.data_ptr(), t, [K, N], flags=b200.TENSOR_FLAG_WEIGHT), b200.tensor(xd.data_ptr(), b200.F32, [K, M])])
    assert b200.supports(op_s)
    for _ in range(2):                       # second pass: peer context already exists
        ctx.compute_op(op_s); ctx.sync()
    ctx.compute_op(op_1); ctx.sync()
    a, b = out_split.cpu().numpy(), out_one.cpu().numpy()
    assert np.isfinite(a).all()
    assert np.array_equal(a, b), float(np.abs(a - b).max())
    want = R.orc_mul_mat(t, W[:16 * rb], x, 16, K)
    assert np.abs(a.reshape(M, N)[:, :16] - want).max() <= 3e-6 * np.abs(want).max()
    ctx.close()
