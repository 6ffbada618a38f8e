"""cortex.llamacpp_b200 -- Python host-side binding of the B200 compute library.

The product is native: ``libggml_b200_kernels.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/ggml_b200.h``) and ``libggml-b200.so`` (the ggml backend llama.cpp loads).  This module only
binds the C ABI through ctypes so that tests and ``bench.py`` can drive the same entry points the ggml
backend calls; torch is used for device memory and nothing else.  There is NO CPU fallback: if the
library or a GPU is missing, calls raise.

Directory name contains a dot, so import it with ``load_package()`` from ``__graft_entry__`` /
``tests/conftest.py`` (registers the module as ``cortex_llamacpp_b200``).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
KERNELS_SO = os.path.join(HERE, "libggml_b200_kernels.so")
BACKEND_SO = os.path.join(HERE, "libggml-b200.so")

# ggml type ids (include/ggml_b200.h)
F32, F16, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K, Q8_K, I32, BF16 = 0, 1, 2, 8, 12, 13, 14, 15, 26, 30
BLOCK = {F32: (1, 4), F16: (1, 2), BF16: (1, 2), I32: (1, 4), Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144),
         Q5_K: (256, 176), Q6_K: (256, 210), Q8_K: (256, 292)}

(OP_NONE, OP_MUL_MAT, OP_MUL_MAT_ID, OP_FLASH_ATTN_EXT, OP_RMS_NORM, OP_ROPE, OP_CPY, OP_CONT, OP_ADD, OP_SUB, OP_MUL,
 OP_DIV, OP_SILU, OP_GELU, OP_RELU, OP_TANH, OP_SIGMOID, OP_GET_ROWS, OP_SOFT_MAX, OP_ARGSORT, OP_SUM_ROWS, OP_SCALE,
 OP_SWIGLU_FUSED, OP_RMS_NORM_MUL, OP_ALLREDUCE, OP_ARGMAX, OP_COUNT) = range(27)

TENSOR_FLAG_WEIGHT = 1


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("type", C.c_int32), ("flags", C.c_uint32), ("ne", C.c_int64 * 4), ("nb", C.c_uint64 * 4)]


class Op(C.Structure):
    _fields_ = [("op", C.c_int32), ("n_src", C.c_int32), ("params", C.c_int32 * 16), ("dst", Tensor), ("src", Tensor * 4)]


def row_size(t, k):
    be, bb = BLOCK[t]
    assert k % be == 0
    return k // be * bb


_lib = None


class B200Error(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(KERNELS_SO):
        raise B200Error("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % KERNELS_SO)
    L = C.CDLL(KERNELS_SO)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    sig = {
        "b200_abi_version": (C.c_int, []),
        "b200_device_count": (C.c_int, []),
        "b200_device_info": (C.c_int, [C.c_int, C.c_char_p, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "b200_last_error": (C.c_char_p, []),
        "b200_ctx_create": (vp, [C.c_int]),
        "b200_ctx_destroy": (None, [vp]),
        "b200_ctx_device": (C.c_int, [vp]),
        "b200_ctx_stream": (vp, [vp]),
        "b200_synchronize": (C.c_int, [vp]),
        "b200_malloc": (vp, [C.c_int, sz]),
        "b200_free": (None, [C.c_int, vp]),
        "b200_host_malloc": (vp, [sz]),
        "b200_host_free": (None, [vp]),
        "b200_memset": (C.c_int, [C.c_int, vp, C.c_int, sz]),
        "b200_memcpy_h2d": (C.c_int, [C.c_int, vp, vp, sz]),
        "b200_memcpy_d2h": (C.c_int, [C.c_int, vp, vp, sz]),
        "b200_memcpy_d2d": (C.c_int, [C.c_int, vp, C.c_int, vp, sz]),
        "b200_memcpy_h2d_async": (C.c_int, [vp, vp, vp, sz]),
        "b200_memcpy_d2h_async": (C.c_int, [vp, vp, vp, sz]),
        "b200_memcpy_d2d_async": (C.c_int, [vp, vp, C.c_int, vp, sz]),
        "b200_alloc_size": (sz, [i32, C.POINTER(i64), sz]),
        "b200_event_create": (vp, [C.c_int]),
        "b200_event_destroy": (None, [vp]),
        "b200_event_record": (C.c_int, [vp, vp]),
        "b200_event_wait": (C.c_int, [vp, vp]),
        "b200_event_synchronize": (C.c_int, [vp]),
        "b200_event_elapsed_ms": (C.c_float, [vp, vp]),
        "b200_supports_op": (C.c_int, [C.c_int, C.POINTER(Op)]),
        "b200_upload": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "b200_upload_slice": (C.c_int, [vp, vp, vp, C.c_size_t, i64, i64, C.c_size_t, C.c_size_t]),
        "b200_graph_compute": (C.c_int, [vp, C.POINTER(Op), C.c_int]),
        "b200_op_compute": (C.c_int, [vp, C.POINTER(Op)]),
        "b200_kernel_launches": (i64, [vp]),
        "b200_set_option": (C.c_int, [vp, C.c_char_p, C.c_int]),
        "b200_quantize_act": (C.c_int, [vp, i32, vp, vp, i64, i64]),
        "b200_block_sums": (C.c_int, [vp, i32, vp, vp, i64, i64, vp, vp]),
        "b200_debug_set_prof": (C.c_int, [vp, vp]),
        "b200_comm_unique_id": (C.c_int, [vp]),
        "b200_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "b200_comm_peer_handle": (C.c_int, [vp, vp]),
        "b200_comm_peer_attach": (C.c_int, [vp, vp]),
        "b200_comm_rank": (C.c_int, [vp]),
        "b200_comm_world": (C.c_int, [vp]),
        "b200_comm_destroy": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here == the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled by tests from include/ggml_b200.h


def check(rc, what=""):
    if rc != 0:
        raise B200Error("%s failed (%d): %s" % (what, rc, lib().b200_last_error().decode()))


class Context:
    """One device + stream (b200_ctx)."""

    def __init__(self, device=0):
        L = lib()
        if L.b200_device_count() <= device:
            raise B200Error("no sm_100 CUDA device %d (device_count=%d): the CUDA path is mandatory" % (device, L.b200_device_count()))
        self.h = L.b200_ctx_create(device)
        if not self.h:
            raise B200Error("b200_ctx_create: " + L.b200_last_error().decode())
        self.device = device

    def close(self):
        if self.h:
            lib().b200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return lib().b200_ctx_stream(self.h)

    def sync(self):
        check(lib().b200_synchronize(self.h), "synchronize")

    def launches(self):
        return lib().b200_kernel_launches(self.h)

    def set_option(self, key, value):
        check(lib().b200_set_option(self.h, key.encode(), int(value)), "set_option")

    def compute(self, ops):
        arr = (Op * len(ops))(*ops)
        check(lib().b200_graph_compute(self.h, arr, len(ops)), "graph_compute")

    def compute_op(self, op):
        check(lib().b200_op_compute(self.h, C.byref(op)), "op_compute")

    def comm_init(self, rank, world, exchange):
        """Join a tensor-parallel group of `world` processes (one GPU each).  `exchange(blob, src)` must return, on every
        rank, the list of every rank's `blob` (bytes) -- e.g. torch.distributed.all_gather_object; it is the only thing
        the library needs from the host's process launcher (b200_comm_* in include/ggml_b200.h)."""
        L = lib()
        uid = C.create_string_buffer(128)
        if rank == 0:
            check(L.b200_comm_unique_id(uid), "comm_unique_id")
        uid_all = exchange(bytes(uid.raw))
        uid0 = C.create_string_buffer(uid_all[0], 128)
        check(L.b200_comm_init(self.h, uid0, rank, world), "comm_init")
        if world > 1:
            hb = C.create_string_buffer(64)
            check(L.b200_comm_peer_handle(self.h, hb), "comm_peer_handle")
            hs = exchange(bytes(hb.raw))
            allh = C.create_string_buffer(b"".join(hs), 64 * world)
            check(L.b200_comm_peer_attach(self.h, allh), "comm_peer_attach")


def tensor(data_ptr, type_, ne, nb=None, flags=0):
    """Describe a device tensor ggml-style: ne elements, nb byte strides (contiguous when omitted)."""
    ne = list(ne) + [1] * (4 - len(ne))
    t = Tensor()
    t.data = data_ptr
    t.type = type_
    t.flags = flags
    if nb is None:
        be, bb = BLOCK[type_]
        nb = [bb, ne[0] // be * bb]
        nb.append(nb[1] * ne[1])
        nb.append(nb[2] * ne[2])
    nb = list(nb) + [0] * (4 - len(nb))
    for i in range(4):
        t.ne[i] = ne[i]
        t.nb[i] = nb[i]
    return t


def make_op(op_id, dst, srcs, params=None):
    o = Op()
    C.memset(C.byref(o), 0, C.sizeof(o))
    o.op = op_id
    o.n_src = len(srcs)
    o.dst = dst
    for i, s in enumerate(srcs):
        if s is not None:
            o.src[i] = s
    if params:
        for i, p in enumerate(params):
            if isinstance(p, float):
                o.params[i] = C.c_int32.from_buffer_copy(C.c_float(p)).value
            else:
                o.params[i] = int(p)
    return o


def supports(op, device=0):
    return bool(lib().b200_supports_op(device, C.byref(op)))
