"""Host-side harness mirror of the reference's GPU Operation interface, over the axb C ABI.

`DeviceArray` plays the role of a Device::GPU axiom::Tensor (device storage + shape + BYTE strides + byte
offset, include/axiom/tensor.hpp:40-58); the module-level functions mirror Operation::execute_unary /
execute_binary / execute_reduction / execute_matmul / execute_where (include/axiom/operations.hpp:210-279):
same argument meaning, results freshly allocated on the device, errors raised as exceptions.

This is TEST/BENCH plumbing: every function is a direct call into libaxb.so. No computation happens here
and there is no CPU fallback — without the library (or without a B200) these calls raise.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .capi import AxbTensor, check, lib

_NP2AXB = {np.dtype(np.bool_): capi.BOOL, np.dtype(np.int8): capi.I8, np.dtype(np.int16): capi.I16,
           np.dtype(np.int32): capi.I32, np.dtype(np.int64): capi.I64, np.dtype(np.uint8): capi.U8,
           np.dtype(np.uint16): capi.U16, np.dtype(np.uint32): capi.U32, np.dtype(np.uint64): capi.U64,
           np.dtype(np.float16): capi.F16, np.dtype(np.float32): capi.F32, np.dtype(np.float64): capi.F64}
_AXB2NP = {v: k for k, v in _NP2AXB.items()}
_AXB2NP[capi.BF16] = np.dtype(np.uint16)  # bf16 travels as raw bit patterns


class Storage:
    """Owning device allocation from the stream-ordered caching allocator (== CudaStorage)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        check(lib().axb_alloc(self.nbytes, None, ctypes.byref(p)))
        self.ptr = p.value or 0

    def __del__(self):
        try:
            if self.ptr:
                lib().axb_free(ctypes.c_void_p(self.ptr), None)
                self.ptr = 0
        except Exception:
            pass


class DeviceArray:
    def __init__(self, storage: Storage, dtype: int, shape: Sequence[int], strides: Sequence[int], offset: int = 0):
        self.storage = storage
        self.dtype = int(dtype)
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self.offset = int(offset)  # byte offset of element [0,...,0]

    # ---- construction -------------------------------------------------------------------------------
    @staticmethod
    def empty(shape: Sequence[int], dtype: int) -> "DeviceArray":
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        st = Storage(n * capi.DTYPE_SIZE[dtype])
        strides, acc = [], capi.DTYPE_SIZE[dtype]
        for e in reversed(shape):
            strides.append(acc)
            acc *= max(int(e), 1)
        return DeviceArray(st, dtype, shape, list(reversed(strides)))

    @staticmethod
    def from_numpy(a: np.ndarray, bf16: bool = False) -> "DeviceArray":
        """Uploads the buffer that backs `a` and returns a view with a's strides/offset preserved
        (transposes, slices, broadcast (stride 0) and flipped (negative stride) views stay views)."""
        base = a
        while isinstance(base.base, np.ndarray):
            base = base.base
        if not base.flags.c_contiguous:
            base = np.ascontiguousarray(a)
            a = base
        lo = base.__array_interface__["data"][0]
        off = a.__array_interface__["data"][0] - lo
        st = Storage(base.nbytes)
        if base.nbytes:
            check(lib().axb_memcpy_h2d(ctypes.c_void_p(st.ptr), ctypes.c_void_p(lo), base.nbytes, None))
            check(lib().axb_stream_sync(None))  # host buffer may be freed by the caller
        dt = capi.BF16 if bf16 else _NP2AXB[a.dtype]
        if bf16:
            assert a.dtype == np.uint16
        return DeviceArray(st, dt, a.shape, a.strides, off)

    # ---- views (zero copy, like Tensor::transpose / slice / expand) -----------------------------------
    def view(self, shape, strides, offset=None) -> "DeviceArray":
        return DeviceArray(self.storage, self.dtype, shape, strides, self.offset if offset is None else offset)

    def T(self) -> "DeviceArray":
        return self.view(self.shape[::-1], self.strides[::-1])

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if len(self.shape) else 1

    @property
    def itemsize(self) -> int:
        return capi.DTYPE_SIZE[self.dtype]

    def desc(self) -> AxbTensor:
        return capi.make_tensor(self.storage.ptr, self.dtype, self.shape, self.strides, self.offset)

    def is_contiguous(self) -> bool:
        acc = self.itemsize
        for e, s in zip(reversed(self.shape), reversed(self.strides)):
            if e != 1 and s != acc:
                return False
            acc *= e
        return True

    # ---- readback -----------------------------------------------------------------------------------
    def contiguous(self) -> "DeviceArray":
        if self.is_contiguous():
            return self
        out = DeviceArray.empty(self.shape, self.dtype)
        check(lib().axb_gather_strided(ctypes.byref(self.desc()), ctypes.byref(out.desc()), None))
        return out

    def numpy(self) -> np.ndarray:
        c = self.contiguous()
        out = np.empty(self.shape, dtype=_AXB2NP[self.dtype])
        check(lib().axb_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(c.storage.ptr + c.offset),
                                   out.nbytes, None, 1))
        return out


def _bshape(*shapes: Sequence[int]) -> Tuple[int, ...]:
    """NumPy broadcast through the library's own rule (ops::compute_broadcast_info): errors surface as AxbError."""
    cur = tuple(int(e) for e in shapes[0])
    for s in shapes[1:]:
        a = (ctypes.c_int64 * capi.AXB_MAX_DIMS)(*cur)
        b = (ctypes.c_int64 * capi.AXB_MAX_DIMS)(*[int(e) for e in s])
        out = (ctypes.c_int64 * capi.AXB_MAX_DIMS)()
        n = ctypes.c_int32()
        check(lib().axb_broadcast_shape(a, len(cur), b, len(s), out, ctypes.byref(n)))
        cur = tuple(out[i] for i in range(n.value))
    return cur


def synchronize() -> None:
    check(lib().axb_stream_sync(None))


def launch_count() -> int:
    return int(lib().axb_launch_count())


# ---- family 1 -------------------------------------------------------------------------------------
def elementwise(prog: List[capi.AxbInstr], inputs: Sequence[DeviceArray], out_shape: Optional[Sequence[int]] = None,
                out: Optional[DeviceArray] = None) -> DeviceArray:
    n = len(prog)
    P = (capi.AxbInstr * n)(*prog)
    ins = (AxbTensor * max(len(inputs), 1))(*[x.desc() for x in inputs])
    if out is None:
        shape = tuple(out_shape) if out_shape is not None else _bshape(*[x.shape for x in inputs])
        out = DeviceArray.empty(shape, prog[-1].dtype)
    check(lib().axb_elementwise(P, n, ins, len(inputs), ctypes.byref(out.desc()), None))
    return out


def unary(op: str, x: DeviceArray, param: float = 0.0) -> DeviceArray:
    o = capi.OP[op]
    out = DeviceArray.empty(x.shape, lib().axb_unary_result_dtype(o, x.dtype))
    check(lib().axb_unary(o, ctypes.byref(x.desc()), ctypes.byref(out.desc()), param, None))
    return out


def binary(op: str, a: DeviceArray, b: DeviceArray) -> DeviceArray:
    o = capi.OP[op]
    dt = lib().axb_binary_result_dtype(o, a.dtype, b.dtype)
    out = DeviceArray.empty(_bshape(a.shape, b.shape), dt)
    check(lib().axb_binary(o, ctypes.byref(a.desc()), ctypes.byref(b.desc()), ctypes.byref(out.desc()), None))
    return out


def where(cond: DeviceArray, a: DeviceArray, b: DeviceArray) -> DeviceArray:
    dt = lib().axb_promote_types(a.dtype, b.dtype)
    out = DeviceArray.empty(_bshape(cond.shape, a.shape, b.shape), dt)
    check(lib().axb_where(ctypes.byref(cond.desc()), ctypes.byref(a.desc()), ctypes.byref(b.desc()),
                          ctypes.byref(out.desc()), None))
    return out


def masked_fill(x: DeviceArray, mask: DeviceArray, value: float) -> DeviceArray:
    out = DeviceArray.empty(_bshape(x.shape, mask.shape), x.dtype)
    check(lib().axb_masked_fill(ctypes.byref(x.desc()), ctypes.byref(mask.desc()), value, ctypes.byref(out.desc()),
                                None))
    return out


def cast(x: DeviceArray, dtype: int) -> DeviceArray:
    out = DeviceArray.empty(x.shape, dtype)
    check(lib().axb_cast(ctypes.byref(x.desc()), ctypes.byref(out.desc()), None))
    return out


def fill(shape: Sequence[int], dtype: int, value: float) -> DeviceArray:
    out = DeviceArray.empty(shape, dtype)
    check(lib().axb_fill(ctypes.byref(out.desc()), value, None))
    return out


def randn(shape: Sequence[int], seed: int, mean: float = 0.0, std: float = 1.0) -> DeviceArray:
    out = DeviceArray.empty(shape, capi.F32)
    check(lib().axb_randn(ctypes.byref(out.desc()), seed, mean, std, None))
    return out


# ---- family 2 -------------------------------------------------------------------------------------
def _reduced_shape(shape, axes, keep_dims):
    nd = len(shape)
    if not axes:
        return tuple([1] * nd) if keep_dims else (1,)
    out = []
    for i, e in enumerate(shape):
        if i in axes:
            if keep_dims:
                out.append(1)
        else:
            out.append(e)
    return tuple(out) if out else (1,)  # a full reduction is {1}, not rank-0 (graph_registry.cpp:101-111)


def reduce(op: str, x: DeviceArray, axes: Sequence[int] = (), keep_dims: bool = False) -> DeviceArray:
    o = capi.OP[op]
    axes = list(axes)
    out = DeviceArray.empty(_reduced_shape(x.shape, axes, keep_dims), lib().axb_reduce_result_dtype(o, x.dtype))
    arr = (ctypes.c_int32 * max(len(axes), 1))(*axes)
    check(lib().axb_reduce(o, ctypes.byref(x.desc()), arr, len(axes), int(keep_dims), ctypes.byref(out.desc()), None))
    return out


def argreduce(op: str, x: DeviceArray, axis: int, keep_dims: bool = False) -> DeviceArray:
    o = capi.OP[op]
    if axis < 0:  # flattened (cpu_operations.cpp:2240-2247): rank-0 result
        shape = tuple([1] * len(x.shape)) if keep_dims else ()
    else:
        shape = tuple(1 if i == axis else e for i, e in enumerate(x.shape) if keep_dims or i != axis)
    out = DeviceArray.empty(shape, capi.I64)
    check(lib().axb_argreduce(o, ctypes.byref(x.desc()), axis, int(keep_dims), ctypes.byref(out.desc()), None))
    return out


def softmax(x: DeviceArray, axis: int = -1, log: bool = False) -> DeviceArray:
    out = DeviceArray.empty(x.shape, x.dtype)
    check(lib().axb_softmax(ctypes.byref(x.desc()), axis, int(log), ctypes.byref(out.desc()), None))
    return out


def layer_norm(x: DeviceArray, w: Optional[DeviceArray], b: Optional[DeviceArray], axis: int = -1,
               eps: float = 1e-5, out_dtype: Optional[int] = None) -> DeviceArray:
    out = DeviceArray.empty(x.shape, x.dtype if out_dtype is None else out_dtype)
    check(lib().axb_layer_norm(ctypes.byref(x.desc()), ctypes.byref(w.desc()) if w else None,
                               ctypes.byref(b.desc()) if b else None, axis, eps, ctypes.byref(out.desc()), None))
    return out


def rms_norm(x: DeviceArray, w: Optional[DeviceArray], axis: int = -1, eps: float = 1e-5,
             out_dtype: Optional[int] = None) -> DeviceArray:
    out = DeviceArray.empty(x.shape, x.dtype if out_dtype is None else out_dtype)
    check(lib().axb_rms_norm(ctypes.byref(x.desc()), ctypes.byref(w.desc()) if w else None, axis, eps,
                             ctypes.byref(out.desc()), None))
    return out


# ---- family 3 -------------------------------------------------------------------------------------
def matmul(a: DeviceArray, b: DeviceArray, transpose_a: bool = False, transpose_b: bool = False,
           bias: Optional[DeviceArray] = None, activation: Optional[str] = None, const_ops: int = 0) -> DeviceArray:
    shape = (ctypes.c_int64 * capi.AXB_MAX_DIMS)()
    nd = ctypes.c_int32()
    ad, bd = a.desc(), b.desc()
    check(lib().axb_matmul_shape(ctypes.byref(ad), ctypes.byref(bd), int(transpose_a), int(transpose_b), shape,
                                 ctypes.byref(nd)))
    out = DeviceArray.empty([shape[i] for i in range(nd.value)], lib().axb_promote_types(a.dtype, b.dtype))
    epi = None
    if bias is not None or activation is not None or const_ops:
        e = capi.AxbMatmulEpilogue()
        e.flags = const_ops
        bdesc = bias.desc() if bias is not None else None
        e.bias = ctypes.pointer(bdesc) if bdesc is not None else None
        e.activation = capi.OP[activation] if activation else 0
        epi = ctypes.byref(e)
    check(lib().axb_matmul(ctypes.byref(ad), ctypes.byref(bd), int(transpose_a), int(transpose_b), epi,
                           ctypes.byref(out.desc()), None))
    return out


def gather(x: DeviceArray, dim: int, idx: DeviceArray) -> DeviceArray:
    out = DeviceArray.empty(idx.shape, x.dtype)
    check(lib().axb_gather(ctypes.byref(x.desc()), dim, ctypes.byref(idx.desc()), ctypes.byref(out.desc()), None))
    return out


def index_select(x: DeviceArray, dim: int, idx: DeviceArray) -> DeviceArray:
    shape = list(x.shape)
    shape[dim] = idx.size
    out = DeviceArray.empty(shape, x.dtype)
    check(lib().axb_index_select(ctypes.byref(x.desc()), dim, ctypes.byref(idx.desc()), ctypes.byref(out.desc()),
                                 None))
    return out


# ---- spatial operators next to the path (csrc/spatial.cu; ops::conv2d / conv_transpose1d,2d / interpolate) ------------
def _i32s(v, n):
    v = [int(v)] * n if isinstance(v, int) else [int(t) for t in v]
    return (ctypes.c_int32 * n)(*v)


def conv2d(x: DeviceArray, w: DeviceArray, bias: Optional[DeviceArray] = None, stride=1, padding=0, dilation=1,
           groups: int = 1) -> DeviceArray:
    st, pd, dl = _i32s(stride, 2), _i32s(padding, 2), _i32s(dilation, 2)
    n, _, h, wd = x.shape
    cout, _, kh, kw = w.shape
    oh = (h + 2 * pd[0] - dl[0] * (kh - 1) - 1) // st[0] + 1
    ow = (wd + 2 * pd[1] - dl[1] * (kw - 1) - 1) // st[1] + 1
    out = DeviceArray.empty((n, cout, oh, ow), x.dtype)
    check(lib().axb_conv2d(ctypes.byref(x.desc()), ctypes.byref(w.desc()), ctypes.byref(bias.desc()) if bias else None,
                           st, pd, dl, groups, ctypes.byref(out.desc()), None))
    return out


def conv(x: DeviceArray, w: DeviceArray, bias: Optional[DeviceArray] = None, stride=1, padding=0, dilation=1, groups: int = 1,
         activation: Optional[str] = None) -> DeviceArray:
    """axb_conv: 1-D / 2-D forward convolution; the tensor-core implicit GEMM when the problem qualifies (matmul_last_path() == 4)."""
    nsp = len(x.shape) - 2
    st, pd, dl = _i32s(stride, nsp), _i32s(padding, nsp), _i32s(dilation, nsp)
    sp = [(x.shape[2 + i] + 2 * pd[i] - dl[i] * (w.shape[2 + i] - 1) - 1) // st[i] + 1 for i in range(nsp)]
    out = DeviceArray.empty((x.shape[0], w.shape[0], *sp), x.dtype)
    check(lib().axb_conv(ctypes.byref(x.desc()), ctypes.byref(w.desc()), ctypes.byref(bias.desc()) if bias else None, nsp, st, pd, dl,
                         groups, capi.OP[activation] if activation else 0, ctypes.byref(out.desc()), None))
    return out


def conv_transpose(x: DeviceArray, w: DeviceArray, bias: Optional[DeviceArray] = None, stride=1, padding=0,
                   output_padding=0, dilation=1, groups: int = 1) -> DeviceArray:
    nsp = len(x.shape) - 2
    st, pd, op, dl = _i32s(stride, nsp), _i32s(padding, nsp), _i32s(output_padding, nsp), _i32s(dilation, nsp)
    sp = [(x.shape[2 + i] - 1) * st[i] - 2 * pd[i] + dl[i] * (w.shape[2 + i] - 1) + op[i] + 1 for i in range(nsp)]
    out = DeviceArray.empty((x.shape[0], w.shape[1] * groups, *sp), x.dtype)
    check(lib().axb_conv_transpose(ctypes.byref(x.desc()), ctypes.byref(w.desc()),
                                   ctypes.byref(bias.desc()) if bias else None, nsp, st, pd, dl, groups,
                                   ctypes.byref(out.desc()), None))
    return out


def interpolate(x: DeviceArray, size: Sequence[int], mode: str = "nearest", align_corners: bool = False) -> DeviceArray:
    out = DeviceArray.empty((x.shape[0], x.shape[1], *size), x.dtype)
    m = {"nearest": 0, "linear": 1, "bilinear": 1, "bicubic": 2}[mode]
    check(lib().axb_interpolate(ctypes.byref(x.desc()), m, int(align_corners), ctypes.byref(out.desc()), None))
    return out


def pool(x: DeviceArray, mode: str, kernel, stride=None, padding=0, count_include_pad: bool = True) -> DeviceArray:
    """ops::max_pool{1,2,3}d / avg_pool{1,2,3}d on (N, C, spatial...) (axb_pool). stride None = kernel (operations.cpp:2369)."""
    nsp = len(x.shape) - 2
    k = _i32s(kernel, nsp)
    st = _i32s(kernel if stride is None else stride, nsp)
    pd = _i32s(padding, nsp)
    sp = [(x.shape[2 + i] + 2 * pd[i] - k[i]) // st[i] + 1 for i in range(nsp)]
    out = DeviceArray.empty((x.shape[0], x.shape[1], *sp), x.dtype)
    check(lib().axb_pool(ctypes.byref(x.desc()), 1 if mode == "avg" else 0, nsp, k, st, pd, int(count_include_pad),
                         ctypes.byref(out.desc()), None))
    return out


def sort(x: DeviceArray, axis: int = -1, descending: bool = False, k: Optional[int] = None, want_values: bool = True,
         want_indices: bool = True):
    """ops::sort / argsort / topk (axb_sort): (values, indices) with extent k (default: all) along `axis`."""
    ax = axis + len(x.shape) if axis < 0 else axis
    oshape = list(x.shape)
    if k is not None:
        oshape[ax] = int(k)
    v = DeviceArray.empty(oshape, x.dtype) if want_values else None
    i = DeviceArray.empty(oshape, capi.I64) if want_indices else None
    check(lib().axb_sort(ctypes.byref(x.desc()), axis, int(descending), ctypes.byref(v.desc()) if v else None,
                         ctypes.byref(i.desc()) if i else None, None))
    return v, i


class _AdoptedStorage(Storage):
    def __init__(self, ptr: int, nbytes: int):  # a block that came out of axb_alloc inside the library
        self.nbytes, self.ptr = int(nbytes), int(ptr)


def masked_select(x: DeviceArray, mask: DeviceArray) -> DeviceArray:
    """ops::masked_select (axb_masked_select): 1-D result of the input's dtype."""
    p, n = ctypes.c_void_p(), ctypes.c_int64()
    check(lib().axb_masked_select(ctypes.byref(x.desc()), ctypes.byref(mask.desc()), ctypes.byref(p), ctypes.byref(n), None))
    if n.value == 0:
        return DeviceArray.empty((0,), x.dtype)
    return DeviceArray(_AdoptedStorage(p.value, n.value * x.itemsize), x.dtype, (n.value,), (x.itemsize,))


def elementwise_reduce(prog: List[capi.AxbInstr], inputs: Sequence[DeviceArray], op: str) -> DeviceArray:
    """FusedReductionStep: the chain's value under a FULL reduction (axb_elementwise_reduce); result shape {1}."""
    P = (capi.AxbInstr * len(prog))(*prog)
    I = (capi.AxbTensor * len(inputs))(*[t.desc() for t in inputs])
    out = DeviceArray.empty((1,), prog[-1].dtype)
    check(lib().axb_elementwise_reduce(P, len(prog), I, len(inputs), capi.OP[op], ctypes.byref(out.desc()), None))
    return out


def elementwise_reduce_axis(prog: List[capi.AxbInstr], inputs: Sequence[DeviceArray], op: str, axis: int,
                            keep_dims: bool = False) -> DeviceArray:
    """The chain's value reduced over one axis (axb_elementwise_reduce_axis): one kernel for fp32 chains over the last axis."""
    P = (capi.AxbInstr * len(prog))(*prog)
    I = (capi.AxbTensor * len(inputs))(*[t.desc() for t in inputs])
    shape = list(_bshape(*[x.shape for x in inputs]))
    ax = axis + len(shape) if axis < 0 else axis
    oshape = shape[:ax] + ([1] if keep_dims else []) + shape[ax + 1:]
    odt = lib().axb_reduce_result_dtype(capi.OP[op], prog[-1].dtype)
    out = DeviceArray.empty(tuple(oshape), odt)
    check(lib().axb_elementwise_reduce_axis(P, len(prog), I, len(inputs), capi.OP[op], axis, int(keep_dims),
                                            ctypes.byref(out.desc()), None))
    return out


def elementwise_softmax(prog: List[capi.AxbInstr], inputs: Sequence[DeviceArray], log: bool = False,
                        epilogue: Optional[str] = None) -> DeviceArray:
    """softmax / log_softmax over the last axis of the chain's value [+ row sum / mean / max] (axb_elementwise_softmax)."""
    P = (capi.AxbInstr * len(prog))(*prog)
    I = (capi.AxbTensor * len(inputs))(*[t.desc() for t in inputs])
    shape = list(_bshape(*[x.shape for x in inputs]))
    out = DeviceArray.empty(tuple(shape[:-1]) if epilogue else tuple(shape), prog[-1].dtype)
    check(lib().axb_elementwise_softmax(P, len(prog), I, len(inputs), int(log), capi.OP[epilogue] if epilogue else 0,
                                        ctypes.byref(out.desc()), None))
    return out


def elementwise_norm(prog: List[capi.AxbInstr], inputs: Sequence[DeviceArray], kind: str, weight: Optional[DeviceArray] = None,
                     bias: Optional[DeviceArray] = None, eps: float = 1e-5) -> DeviceArray:
    """layer_norm / rms_norm over the last axis of the chain's value (axb_elementwise_norm)."""
    P = (capi.AxbInstr * len(prog))(*prog)
    I = (capi.AxbTensor * len(inputs))(*[t.desc() for t in inputs])
    out = DeviceArray.empty(tuple(_bshape(*[x.shape for x in inputs])), prog[-1].dtype)
    check(lib().axb_elementwise_norm(P, len(prog), I, len(inputs), capi.OP[kind], ctypes.byref(weight.desc()) if weight else None,
                                     ctypes.byref(bias.desc()) if bias else None, eps, ctypes.byref(out.desc()), None))
    return out
