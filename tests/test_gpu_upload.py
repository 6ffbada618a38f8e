"""GGUF -> device load path: pipelined upload from pageable host memory and the sliced variants a tensor-parallel
rank uses to take its shard of a row-split (rows) or K-split (bytes of every row) weight straight out of the mmap'd file."""
import ctypes as C

import numpy as np
import pytest

from util import dev_bytes

pytestmark = pytest.mark.gpu


def test_upload_whole_and_slices_roundtrip(b200, ctx):
    rng = np.random.default_rng(60)
    rows, rb = 4096 + 37, 2304 + 144                      # a Q4_K matrix with K = 4352: 17 blocks of 144 bytes per row
    host = rng.integers(0, 256, rows * rb, dtype=np.uint8)        # pageable memory, like the mmap'd file
    L = b200.lib()
    # whole tensor (crosses several 16 MiB staging chunks? 10 MB here; the 70 MB case below does)
    d = dev_bytes(host.size)
    b200.check(L.b200_upload(ctx.h, d.data_ptr(), host.ctypes.data, host.size), "upload")
    assert np.array_equal(d.cpu().numpy(), host)
    # row shard: rank 1 of 4 of a row-split matrix
    r0, nr = rows // 4, rows // 4
    d = dev_bytes(nr * rb)
    b200.check(L.b200_upload_slice(ctx.h, d.data_ptr(), host.ctypes.data, rb, r0, nr, 0, rb), "upload rows")
    assert np.array_equal(d.cpu().numpy(), host.reshape(rows, rb)[r0:r0 + nr].reshape(-1))
    # K shard: blocks 5..11 of every row (wo / down split along K)
    c0, cb = 5 * 144, 7 * 144
    d = dev_bytes(rows * cb)
    b200.check(L.b200_upload_slice(ctx.h, d.data_ptr(), host.ctypes.data, rb, 0, rows, c0, cb), "upload cols")
    assert np.array_equal(d.cpu().numpy(), host.reshape(rows, rb)[:, c0:c0 + cb].reshape(-1))


def test_upload_large_multichunk(b200, ctx):
    rng = np.random.default_rng(61)
    host = rng.integers(0, 256, (70 << 20) + 12345, dtype=np.uint8)
    d = dev_bytes(host.size)
    b200.check(b200.lib().b200_upload(ctx.h, d.data_ptr(), host.ctypes.data, host.size), "upload")
    assert np.array_equal(d.cpu().numpy(), host)
