"""Shared helpers for tests: synthetic GGUF-layout blocks, layout conversion, torch<->device glue."""
import numpy as np

import reflib as R


def scratch_from_blocks(ta, blocks, K):
    """reference activation blocks of ONE row (block_q8_0 / block_q8_K bytes) -> the library's split scratch layout"""
    q8k = ta == R.Q8_K
    nd = K // (256 if q8k else 32)
    ns = K // (16 if q8k else 32)
    off_d = (K + 15) & ~15
    off_s = off_d + ((nd * 4 + 15) & ~15)
    col = off_s + ((ns * 2 + 15) & ~15)
    out = np.zeros(col, np.uint8)
    if q8k:
        b = blocks.reshape(-1, 292)
        out[:K] = b[:, 4:260].reshape(-1)
        out[off_d:off_d + nd * 4] = b[:, 0:4].reshape(-1)
        out[off_s:off_s + ns * 2] = b[:, 260:292].reshape(-1)
    else:
        b = blocks.reshape(-1, 34)
        out[:K] = b[:, 2:34].reshape(-1)
        d = b[:, 0:2].copy().view(np.float16).astype(np.float32).reshape(-1)
        out[off_d:off_d + nd * 4] = d.view(np.uint8)
        s = b[:, 2:34].view(np.int8).astype(np.int32).sum(1).astype(np.int16)
        out[off_s:off_s + ns * 2] = s.view(np.uint8)
    return out


def rand_quant_rows(t, N, K, rng, scale=0.02):
    """Random but VALID quantised rows in GGUF block layout (no reference needed): random quants and
    6/8-bit sub-scales, fp16 super-scales drawn so dequantised weights are O(scale)."""
    be, bb = R.BLOCK[t]
    nb = N * (K // be)
    blk = rng.integers(0, 256, size=(nb, bb), dtype=np.uint8)

    def f16(vals):
        return np.asarray(vals, np.float16).view(np.uint8).reshape(nb, 2)

    if t == R.Q4_0:
        blk[:, 0:2] = f16(rng.uniform(0.5, 1.5, nb) * scale / 4 * rng.choice([-1, 1], nb))
    elif t == R.Q8_0:
        blk[:, 0:2] = f16(rng.uniform(0.5, 1.5, nb) * scale / 64 * rng.choice([-1, 1], nb))
    elif t in (R.Q4_K, R.Q5_K):
        qmax = 15 if t == R.Q4_K else 31
        blk[:, 0:2] = f16(rng.uniform(0.5, 1.5, nb) * scale / (32 * qmax))
        blk[:, 2:4] = f16(rng.uniform(0.5, 1.5, nb) * scale / 64)
    elif t == R.Q6_K:
        blk[:, 208:210] = f16(rng.uniform(0.5, 1.5, nb) * scale / (64 * 32) * rng.choice([-1, 1], nb))
    return blk.reshape(-1)


def to_dev(arr):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    torch.cuda.synchronize()
    return t


def dev_bytes(n, fill=None):
    import torch
    t = torch.empty(int(n), dtype=torch.uint8, device="cuda")
    if fill is not None:
        t.fill_(fill)
    torch.cuda.synchronize()
    return t


def nmse(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    return float(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30))
