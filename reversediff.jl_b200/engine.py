"""ctypes binding of ``librdb200.so`` (declared in ``include/rdb200.h``).

This is the Python equivalent of the ``ccall`` layer in ``julia/ReverseDiffB200.jl``.  The library
is built in-tree by ``__graft_entry__.build()`` (``nvcc -gencode arch=compute_100a,code=sm_100a``).
There is no fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librdb200.so")

RDB_F64, RDB_F32 = 0, 1


class EngineError(RuntimeError):
    pass


class _Options(C.Structure):
    _fields_ = [("fuse_elementwise", C.c_int32), ("seed_block", C.c_int32), ("use_graph", C.c_int32),
                ("profile", C.c_int32), ("gemm_f64_mode", C.c_int32), ("ozaki_slices", C.c_int32), ("reserved", C.c_int32 * 2)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the B200 engine has no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "rdb_version": (C.c_char_p, []),
        "rdb_last_error": (C.c_char_p, []),
        "rdb_ctx_create": (vp, [i32]),
        "rdb_ctx_destroy": (None, [vp]),
        "rdb_ctx_set_stream": (i32, [vp, vp]),
        "rdb_ctx_synchronize": (i32, [vp]),
        "rdb_ctx_set_gemm_f64_mode": (i32, [vp, i32, i32]),
        "rdb_ctx_gemm_guard": (i32, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(f64)]),
        "rdb_tape_set_async": (i32, [vp, i32]),
        "rdb_tape_guard_stats": (i32, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(f64)]),
        "rdb_comm_unique_id": (i32, [vp]),
        "rdb_comm_create": (vp, [vp, i32, i32, vp]),
        "rdb_comm_destroy": (None, [vp]),
        "rdb_gather_result": (i32, [vp, i32, i32, vp, C.POINTER(i64), i64, vp, i32]),
        "rdb_place_block": (i32, [i32, i32, vp, i64, i64, i64, i64, vp]),
        "rdb_tape_load": (vp, [vp, vp, C.c_size_t, C.POINTER(_Options)]),
        "rdb_tape_destroy": (None, [vp]),
        "rdb_tape_set_input": (i32, [vp, i32, vp, i32]),
        "rdb_forward": (i32, [vp]),
        "rdb_reverse_scalar_seed": (i32, [vp]),
        "rdb_gradient": (i32, [vp, C.POINTER(vp), i32, vp]),
        "rdb_gradient_async": (i32, [vp, C.POINTER(vp), i32, vp]),
        "rdb_tape_wait": (i32, [vp]),
        "rdb_jacobian": (i32, [vp, i32, i64, i64, vp, i64, i32, vp]),
        "rdb_hessian": (i32, [vp, i64, i64, vp, i64, i32, vp, vp]),
        "rdb_get_value": (i32, [vp, i32, vp]),
        "rdb_get_deriv": (i32, [vp, i32, vp]),
        "rdb_buffer_size": (i64, [vp, i32]),
        "rdb_tape_stats": (i32, [vp, C.c_char_p, C.c_size_t]),
        "rdb_tape_launch_count": (i64, [vp]),
        "rdb_tape_slice_pairs": (i64, [vp]),
        "rdb_gemm": (i32, [vp, i32, i32, i32, i64, i64, i64, f64, vp, i64, vp, i64, f64, vp, i64]),
        "rdb_add_sum": (i32, [vp, i32, i64, vp, vp, f64, vp, vp]),
        "rdb_plan_col_panels": (i32, [i64, i64, C.POINTER(i64)]),
        "rdb_schedule_bundles": (i64, [i64, C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError if the library lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "rdb_version", "rdb_last_error", "rdb_ctx_create", "rdb_ctx_destroy", "rdb_ctx_set_stream",
    "rdb_ctx_synchronize", "rdb_ctx_set_gemm_f64_mode", "rdb_ctx_gemm_guard", "rdb_tape_guard_stats", "rdb_tape_set_async", "rdb_comm_unique_id",
    "rdb_comm_create", "rdb_comm_destroy", "rdb_gather_result", "rdb_place_block", "rdb_tape_load", "rdb_tape_destroy", "rdb_tape_set_input", "rdb_forward",
    "rdb_reverse_scalar_seed", "rdb_gradient", "rdb_gradient_async", "rdb_tape_wait", "rdb_jacobian", "rdb_hessian", "rdb_get_value", "rdb_get_deriv",
    "rdb_buffer_size", "rdb_tape_stats", "rdb_tape_launch_count", "rdb_tape_slice_pairs", "rdb_gemm", "rdb_add_sum", "rdb_plan_col_panels", "rdb_schedule_bundles",
]


def _check(rc):
    if rc != 0:
        msg = lib().rdb_last_error()
        raise EngineError(f"librdb200 error {rc}: {msg.decode() if msg else '?'}")


def _is_device(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def device_array(x, device="cuda"):
    """numpy array -> CUDA tensor of the SAME shape whose memory is column-major (what the engine reads and writes; the
    reference's arrays are column-major, src/tracked.jl:77).  A C-contiguous 2-D torch tensor is the TRANSPOSE of that."""
    import torch
    a = np.asarray(x)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).to(device)
    return t.permute(*reversed(range(t.ndim))) if t.ndim >= 2 else t


def _colmajor_strides(shape):
    st, acc = [], 1
    for d in shape:
        st.append(acc)
        acc *= int(d)
    return tuple(st)


class Context:
    def __init__(self, device_id: int = 0):
        self.device_id = device_id
        self.ptr = lib().rdb_ctx_create(device_id)
        if not self.ptr:
            msg = lib().rdb_last_error()
            raise EngineError(f"rdb_ctx_create({device_id}) failed: {msg.decode() if msg else '?'}")

    def set_stream(self, cuda_stream_ptr: int):
        _check(lib().rdb_ctx_set_stream(self.ptr, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        _check(lib().rdb_ctx_synchronize(self.ptr))

    def set_gemm_f64_mode(self, mode: int, slices: int = 0):
        _check(lib().rdb_ctx_set_gemm_f64_mode(self.ptr, mode, slices))

    def gemm_guard(self):
        """(launches checked, launches flagged, worst bound/|product| ratio) of the int8-slice GEMM since the last call."""
        a, b, w = C.c_int64(0), C.c_int64(0), C.c_double(0.0)
        _check(lib().rdb_ctx_gemm_guard(self.ptr, C.byref(a), C.byref(b), C.byref(w)))
        return int(a.value), int(b.value), float(w.value)

    def gemm(self, dtype, ta, tb, m, n, k, alpha, A, lda, Bp, ldb, beta, Cp, ldc):
        _check(lib().rdb_gemm(self.ptr, dtype, int(ta), int(tb), m, n, k, alpha, A, lda, Bp, ldb, beta, Cp, ldc))

    def add_sum(self, dtype, n, a, b, sign, out, sum_out):
        _check(lib().rdb_add_sum(self.ptr, dtype, n, a, b, sign, out, sum_out))

    def close(self):
        if self.ptr:
            lib().rdb_ctx_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """``rdb_comm``: the NCCL communicator behind ``rdb_gather_result`` (one process per GPU).  ``id_bytes`` comes from
    ``Comm.unique_id()`` on rank 0 and reaches the other ranks through the host (``api.make_comm`` uses torch.distributed)."""
    ID_BYTES = 128

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(Comm.ID_BYTES)
        _check(lib().rdb_comm_unique_id(buf))
        return buf.raw

    def __init__(self, ctx: Context, world: int, rank: int, id_bytes: bytes):
        self.ctx, self.world, self.rank = ctx, int(world), int(rank)
        buf = C.create_string_buffer(bytes(id_bytes), Comm.ID_BYTES)
        self.ptr = lib().rdb_comm_create(ctx.ptr, self.world, self.rank, buf)
        if not self.ptr:
            msg = lib().rdb_last_error()
            raise EngineError(f"rdb_comm_create failed: {msg.decode() if msg else '?'}")

    def gather(self, np_dtype, kind, block, bounds, other_dim, result, root=-1):
        """``rdb_gather_result``: ``block``/``result`` are device tensors (flat, column-major); ``bounds`` the partition."""
        b = (C.c_int64 * len(bounds))(*[int(x) for x in bounds])
        dt = RDB_F64 if np.dtype(np_dtype) == np.float64 else RDB_F32
        bp = C.c_void_p(block.data_ptr()) if block is not None and block.numel() else None
        rp = C.c_void_p(result.data_ptr()) if result is not None else None
        _check(lib().rdb_gather_result(self.ptr, dt, int(kind), bp, b, int(other_dim), rp, int(root)))

    def close(self):
        if getattr(self, "ptr", None):
            lib().rdb_comm_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def place_block(kind, block, begin, end, result):
    """``rdb_place_block`` on host arrays: ``result`` is the full column-major matrix (F-contiguous numpy array)."""
    if not (isinstance(result, np.ndarray) and result.flags.f_contiguous and result.ndim == 2):
        raise TypeError("result must be a 2-D F-contiguous numpy array")
    blk = np.asfortranarray(block, dtype=result.dtype)
    m, n = result.shape
    want = (end - begin, n) if kind == 0 else (m, end - begin)
    if blk.shape != want:
        raise ValueError(f"block of shape {blk.shape}, expected {want}")
    dt = RDB_F64 if result.dtype == np.float64 else RDB_F32
    _check(lib().rdb_place_block(dt, int(kind), blk.ctypes.data_as(C.c_void_p), int(begin), int(end), m, n, result.ctypes.data_as(C.c_void_p)))
    return result


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        _default_ctx = Context(dev)
    return _default_ctx


class EngineTape:
    def __init__(self, ctx: Context, program: bytes, fuse_elementwise=1, seed_block=0, use_graph=1, profile=0,
                 gemm_f64_mode=0, ozaki_slices=0):
        self.ctx = ctx
        opts = _Options(int(fuse_elementwise), int(seed_block), int(use_graph), int(profile), int(gemm_f64_mode), int(ozaki_slices))
        buf = (C.c_char * len(program)).from_buffer_copy(program)
        self.ptr = lib().rdb_tape_load(ctx.ptr, buf, len(program), C.byref(opts))
        if not self.ptr:
            msg = lib().rdb_last_error()
            raise EngineError(f"rdb_tape_load failed: {msg.decode() if msg else '?'}")
        hdr = np.frombuffer(program, "<u4", 3, 8)
        self.np_dtype = np.dtype(np.float64 if int(hdr[1]) == RDB_F64 else np.float32)
        self._keep = []

    # ------------------------------------------------------------------------------
    def _host_in(self, x, shape):
        a = np.asarray(x, dtype=self.np_dtype)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"input of shape {a.shape} does not match the recorded shape {tuple(shape)}")
        return np.asfortranarray(a)

    def _check_device(self, x, shape, what):
        """A device tensor crosses the C ABI as a bare pointer: everything the engine assumes about it is checked here.
        Accepted: a 1-D contiguous tensor holding the column-major linearisation, or an N-D tensor of exactly the recorded
        shape with column-major strides (``device_array`` makes one; ``t.T`` of a C-contiguous transposed tensor)."""
        shape = tuple(int(d) for d in shape)
        n = int(np.prod(shape)) if shape else 1
        if not x.is_floating_point() or x.element_size() != self.np_dtype.itemsize:
            raise TypeError(f"{what}: device tensor of dtype {x.dtype} on a {self.np_dtype} tape")
        if x.device.index is not None and x.device.index != self.ctx.device_id:
            raise ValueError(f"{what}: tensor lives on cuda:{x.device.index}, the tape's context on cuda:{self.ctx.device_id}")
        if x.numel() != n:
            raise ValueError(f"{what}: {x.numel()} elements, expected {n} (shape {shape})")
        if x.ndim <= 1:
            if x.ndim == 1 and n > 1 and x.stride(0) != 1:
                raise ValueError(f"{what}: 1-D device tensors must be contiguous")
            return
        want = _colmajor_strides(shape)
        ok = tuple(x.shape) == shape and all(d == 1 or s == w for d, s, w in zip(shape, x.stride(), want))
        if not ok:
            raise ValueError(f"{what}: N-D device tensors must have the recorded shape {shape} and COLUMN-MAJOR strides {want} "
                             f"(got shape {tuple(x.shape)}, strides {tuple(x.stride())}); a C-contiguous torch matrix is the "
                             "transpose of what the engine would read - pass engine.device_array(a), m.T of the transposed "
                             "matrix, or a flat 1-D tensor holding the column-major data")

    def set_input(self, slot, x, shape):
        if _is_device(x):
            self._check_device(x, shape, f"input {slot}")
            _check(lib().rdb_tape_set_input(self.ptr, slot, C.c_void_p(x.data_ptr()), 1))
        else:
            a = self._host_in(x, shape)
            _check(lib().rdb_tape_set_input(self.ptr, slot, a.ctypes.data_as(C.c_void_p), 0))

    def forward(self):
        _check(lib().rdb_forward(self.ptr))

    def reverse_scalar_seed(self):
        _check(lib().rdb_reverse_scalar_seed(self.ptr))

    def _out_array(self, arr, shape, ld=None):
        """Return (pointer, is_device, staging) for a result array.  ``ld`` > rows: ``arr`` is a window into a larger column-major
        matrix (a rank's row block of the full Jacobian): a flat device tensor that reaches at least to the block's last element."""
        if arr is None:
            return None, 0, None
        if _is_device(arr):
            if ld is not None and len(shape) == 2 and ld != shape[0]:
                need = ld * (shape[1] - 1) + shape[0]
                if arr.ndim != 1 or (arr.numel() > 1 and arr.stride(0) != 1) or arr.numel() < need:
                    raise ValueError(f"result: a strided block (ld = {ld}) needs a flat contiguous device tensor of >= {need} elements")
                self._check_device(arr[:1], (1,), "result")
                return arr.data_ptr(), 1, None
            self._check_device(arr, shape, "result")
            return arr.data_ptr(), 1, None
        if not isinstance(arr, np.ndarray) or arr.dtype != self.np_dtype:
            raise TypeError(f"result arrays must be numpy arrays of dtype {self.np_dtype}")
        if arr.size != int(np.prod(shape)):
            raise ValueError("result array has the wrong size")
        if arr.flags.f_contiguous:
            return arr.ctypes.data, 0, None
        st = np.empty(shape, self.np_dtype, order="F")
        return st.ctypes.data, 0, st

    def gradient(self, results, shapes, want_value=True):
        n = len(results)
        ptrs = (C.c_void_p * n)()
        staged = []
        dev = None
        for i, (r, s) in enumerate(zip(results, shapes)):
            p, d, st = self._out_array(r, s)
            if p is not None:
                dev = d if dev is None else dev
                if d != dev:
                    raise ValueError("results must be all host or all device arrays")
            ptrs[i] = p
            staged.append((r, st))
        val = np.zeros(1, self.np_dtype)
        _check(lib().rdb_gradient(self.ptr, ptrs, dev or 0, val.ctypes.data_as(C.c_void_p) if want_value else None))
        for r, st in staged:
            if st is not None:
                r[...] = st.reshape(r.shape, order="F")
        return val[0] if want_value else None

    def gradient_async(self, results, shapes, want_value=True):
        """``rdb_gradient_async``: enqueue only.  Host results must be Fortran-contiguous arrays of the tape's dtype (page-locked
        ones keep the copies off the caller's thread); ``wait()`` completes the call and returns the primal."""
        n = len(results)
        ptrs = (C.c_void_p * n)()
        dev = None
        for i, (r, s) in enumerate(zip(results, shapes)):
            p, d, st = self._out_array(r, s)
            if st is not None:
                raise ValueError("asynchronous calls write straight into the result arrays: pass Fortran-contiguous arrays of the tape's dtype")
            if p is not None:
                dev = d if dev is None else dev
                if d != dev:
                    raise ValueError("results must be all host or all device arrays")
            ptrs[i] = p
        val = None
        if want_value:
            try:                                          # a page-locked scalar, or the tiny D2H copy blocks the caller until the forward sweep is done
                import torch
                self._pending_val_t = torch.zeros(1, dtype=torch.float64 if self.np_dtype == np.float64 else torch.float32).pin_memory()
                val = self._pending_val_t.numpy()
            except Exception:
                val = np.zeros(1, self.np_dtype)
        self._pending = (ptrs, results, val)              # keep everything the device writes to alive until wait()
        _check(lib().rdb_gradient_async(self.ptr, ptrs, dev or 0, val.ctypes.data_as(C.c_void_p) if want_value else None))

    def wait(self):
        """``rdb_tape_wait``: results of the pending asynchronous call are valid after this returns; returns its primal (or None)."""
        _check(lib().rdb_tape_wait(self.ptr))
        pend, self._pending = getattr(self, "_pending", None), None
        if pend is None or pend[2] is None:
            return None
        return pend[2][0].copy()

    def jacobian(self, slot, row_begin, row_end, result, shape, want_value=False, out_shape=None, ld=None):
        p, d, st = self._out_array(result, shape, ld)
        ld = (row_end - row_begin) if ld is None else ld
        val = None
        vp = None
        if want_value and not d:
            val = np.zeros(out_shape, self.np_dtype, order="F")
            vp = val.ctypes.data_as(C.c_void_p)
        _check(lib().rdb_jacobian(self.ptr, slot, row_begin, row_end, C.c_void_p(p), ld, d, vp))
        if st is not None:
            result[...] = st.reshape(result.shape, order="F")
        return val

    def hessian(self, col_begin, col_end, result, shape, want_grad=False, ld=None, grad_out=None):
        p, d, st = self._out_array(result, shape)
        n = shape[0]
        ld = n if ld is None else ld
        g = val = None
        gp = vp = None
        if want_grad:
            val = np.zeros(1, self.np_dtype)
            vp = val.ctypes.data_as(C.c_void_p)
            if _is_device(grad_out):                      # the gradient goes straight into the caller's device tensor
                self._check_device(grad_out, (n,), "gradient result")
                g, gp = grad_out, C.c_void_p(grad_out.data_ptr())
            else:
                g = np.zeros(n, self.np_dtype)
                gp = g.ctypes.data_as(C.c_void_p)
        _check(lib().rdb_hessian(self.ptr, col_begin, col_end, C.c_void_p(p), ld, d, gp, vp))
        if st is not None:
            result[...] = st.reshape(result.shape, order="F")
        return g, (val[0] if val is not None else None)

    # debug / parity accessors ---------------------------------------------------------
    def buffer_size(self, b):
        return int(lib().rdb_buffer_size(self.ptr, b))

    def get_value(self, b, shape=None):
        n = self.buffer_size(b)
        out = np.zeros(n, self.np_dtype)
        _check(lib().rdb_get_value(self.ptr, b, out.ctypes.data_as(C.c_void_p)))
        return out if shape is None else out.reshape(shape, order="F")

    def get_deriv(self, b, shape=None):
        n = self.buffer_size(b)
        out = np.zeros(n, self.np_dtype)
        _check(lib().rdb_get_deriv(self.ptr, b, out.ctypes.data_as(C.c_void_p)))
        return out if shape is None else out.reshape(shape, order="F")

    def stats(self) -> str:
        buf = C.create_string_buffer(1 << 20)
        _check(lib().rdb_tape_stats(self.ptr, buf, len(buf)))
        return buf.value.decode()

    def set_async(self, on=True):
        """``rdb_tape_set_async``: device-result ``hessian!`` calls return without synchronising (``ctx.synchronize()`` waits)."""
        _check(lib().rdb_tape_set_async(self.ptr, 1 if on else 0))

    def guard_stats(self):
        """dict(checked, flagged, fallbacks, worst_ratio): the int8-slice GEMM's accuracy certificate on this tape since load."""
        a, b, f, w = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_double(0.0)
        _check(lib().rdb_tape_guard_stats(self.ptr, C.byref(a), C.byref(b), C.byref(f), C.byref(w)))
        return {"checked": int(a.value), "flagged": int(b.value), "fallbacks": int(f.value), "worst_ratio": float(w.value),
                "slice_pairs": int(lib().rdb_tape_slice_pairs(self.ptr))}

    def launch_count(self) -> int:
        return int(lib().rdb_tape_launch_count(self.ptr))

    def close(self):
        if getattr(self, "ptr", None):
            lib().rdb_tape_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
