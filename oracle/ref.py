"""oracle/ref.py — ctypes driver for oracle/_ref/libaxiom_ref.so (the UNMODIFIED reference CPU backend).

TEST INFRASTRUCTURE ONLY. Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs. Never imported by axiom_b200/.

The .so is built by oracle/Makefile from the reference's own sources (compiled in place from
/root/reference); it travels to the GPU box as a prebuilt artefact. `available()` tells callers whether
it is there; tests that need it skip when it is not (the restated oracle in oracle/axiom_oracle.py is
the always-present checker).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libaxiom_ref.so")

# axiom::DType numeric values (include/axiom/dtype.hpp:170-192)
DT = dict(bool=0, int8=1, int16=2, int32=3, int64=4, uint8=5, uint16=6, uint32=7, uint64=8,
          float16=9, bfloat16=10, float32=11, float64=12)
_NP_OF = {0: np.bool_, 1: np.int8, 2: np.int16, 3: np.int32, 4: np.int64, 5: np.uint8, 6: np.uint16,
          7: np.uint32, 8: np.uint64, 9: np.float16, 10: np.uint16, 11: np.float32, 12: np.float64}

# axiom::ops::OpType numeric values (include/axiom/operations.hpp:55-208)
_OPNAMES = ["Add", "Subtract", "Multiply", "Divide", "Power", "Modulo", "Equal", "NotEqual", "Less", "LessEqual",
            "Greater", "GreaterEqual", "LogicalAnd", "LogicalOr", "LogicalXor", "LogicalNot", "BitwiseAnd",
            "BitwiseOr", "BitwiseXor", "LeftShift", "RightShift", "Maximum", "Minimum", "Atan2", "Hypot", "Negate",
            "Abs", "Sqrt", "Exp", "Log", "Sin", "Cos", "Tan", "Erf", "Sign", "Floor", "Ceil", "Trunc", "Round",
            "Reciprocal", "Square", "Cbrt", "IsNaN", "IsInf", "IsFinite", "Conj", "Real", "Imag", "ReLU",
            "LeakyReLU", "SiLU", "Sigmoid", "Tanh", "GELU", "Softmax", "LogSoftmax", "Sum", "Mean", "Max", "Min",
            "ArgMax", "ArgMin", "Any", "All", "Prod", "MatMul", "BatchMatMul", "Where", "MaskedFill",
            "MaskedSelect", "Gather", "Scatter", "IndexSelect", "Take", "TakeAlongAxis", "LayerNorm", "RMSNorm",
            "Dropout", "Cast"]
OP = {n: i for i, n in enumerate(_OPNAMES)}

_lib: Optional[ctypes.CDLL] = None


def available() -> bool:
    return os.path.exists(_SO)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"reference oracle not built: {_SO} (run `make -C oracle ref` where /root/reference exists)")
        L = ctypes.CDLL(_SO)
        vp, i32, i64, u64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float
        pi64, pi32 = ctypes.POINTER(i64), ctypes.POINTER(i32)
        sig = {
            "axref_last_error": (ctypes.c_char_p, []),
            "axref_num_threads": (i32, []),
            "axref_set_num_threads": (None, [i32]),
            "axref_init": (None, []),
            "axref_from_buffer": (vp, [vp, u64, i32, pi64, pi64, i64, i32]),
            "axref_free": (None, [vp]),
            "axref_info": (i32, [vp, pi32, pi64, pi32, pi32]),
            "axref_sync": (i32, [vp]),
            "axref_read": (i64, [vp, vp, u64]),
            "axref_unary": (vp, [i32, vp, f32]),
            "axref_binary": (vp, [i32, vp, vp]),
            "axref_where": (vp, [vp, vp, vp]),
            "axref_where_scalar": (vp, [vp, vp, f32]),
            "axref_masked_fill": (vp, [vp, vp, f32]),
            "axref_binary_scalar_f32": (vp, [i32, vp, f32]),
            "axref_binary_scalar_i32": (vp, [i32, vp, i32]),
            "axref_reduce": (vp, [i32, vp, pi32, i32, i32]),
            "axref_argreduce": (vp, [i32, vp, i32, i32]),
            "axref_softmax": (vp, [vp, i32, i32]),
            "axref_layer_norm": (vp, [vp, vp, vp, i32, f32]),
            "axref_rms_norm": (vp, [vp, vp, i32, f32]),
            "axref_matmul": (vp, [vp, vp, i32, i32]),
            "axref_astype": (vp, [vp, i32]),
            "axref_contiguous": (vp, [vp]),
            "axref_transpose": (vp, [vp]),
            "axref_permute": (vp, [vp, pi32, i32]),
            "axref_reshape": (vp, [vp, pi64, i32]),
            "axref_readme_chain": (vp, [vp]),
            "axref_config2_chain": (vp, [vp, vp, vp]),
            "axref_gather": (vp, [vp, i32, vp]),
            "axref_index_select": (vp, [vp, i32, vp]),
            "axref_conv1d": (vp, [vp, vp, vp, i32, i32, i32, i32]),
            "axref_sdpa": (vp, [vp, vp, vp, vp, f32, i32]),
            "axref_pool": (vp, [vp, i32, i32, pi32, pi32, pi32, i32]),
            "axref_sort": (vp, [vp, i32, i32, i32, i32]),
            "axref_masked_select": (vp, [vp, vp]),
            "axref_time_sync": (ctypes.c_double, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        L.axref_init()
        _lib = L
    return _lib


def _np_dtype_code(a: np.ndarray, bf16: bool) -> int:
    if bf16:
        assert a.dtype == np.uint16
        return DT["bfloat16"]
    return DT[{np.dtype(np.bool_): "bool"}.get(a.dtype, a.dtype.name)]


class RefTensor:
    """Owning handle to an axiom::Tensor living inside the reference library."""

    def __init__(self, handle: int):
        if not handle:
            raise RuntimeError("reference error: " + lib().axref_last_error().decode())
        self.h = ctypes.c_void_p(handle)

    def __del__(self):
        try:
            if self.h:
                lib().axref_free(self.h)
        except Exception:
            pass

    # -- construction -------------------------------------------------------------------------
    @staticmethod
    def from_numpy(a: np.ndarray, bf16: bool = False) -> "RefTensor":
        """Wraps `a` with ITS strides/offset (views, broadcast and negative strides are preserved).

        bf16=True: `a` is a uint16 array holding bfloat16 bit patterns.
        """
        base = a
        while isinstance(base.base, np.ndarray):
            base = base.base
        base = np.ascontiguousarray(base) if not base.flags.c_contiguous and base is a else base
        if base is not a and not base.flags.c_contiguous:
            # unusual base layout: fall back to exporting the raw byte window
            base = np.ascontiguousarray(a)
            a = base
        lo = base.__array_interface__["data"][0]
        off = a.__array_interface__["data"][0] - lo
        # Reference convention (tensor.cpp:279-291): offset_ addresses the LOWEST byte of the view and
        # data() adds sum((shape-1)*|stride|) over negative-stride dims to reach element [0,...,0].
        off -= sum((n - 1) * (-s) for n, s in zip(a.shape, a.strides) if s < 0 and n > 1)
        nbytes = base.nbytes
        shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
        strides = (ctypes.c_int64 * max(a.ndim, 1))(*a.strides)
        h = lib().axref_from_buffer(ctypes.c_void_p(lo), nbytes, a.ndim, shape, strides, off,
                                    _np_dtype_code(a, bf16))
        return RefTensor(h)

    # -- inspection ---------------------------------------------------------------------------
    def info(self):
        nd, dt, lz = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        shape = (ctypes.c_int64 * 16)()
        if lib().axref_info(self.h, ctypes.byref(nd), shape, ctypes.byref(dt), ctypes.byref(lz)) != 0:
            raise RuntimeError(lib().axref_last_error().decode())
        return tuple(shape[i] for i in range(nd.value)), dt.value, bool(lz.value)

    @property
    def shape(self):
        return self.info()[0]

    @property
    def dtype_code(self):
        return self.info()[1]

    def sync(self):
        if lib().axref_sync(self.h) != 0:
            raise RuntimeError(lib().axref_last_error().decode())
        return self

    def numpy(self) -> np.ndarray:
        """Materialise + copy out. bfloat16 comes back as uint16 bit patterns."""
        shape, dt, _ = self.info()
        out = np.empty(shape, dtype=_NP_OF[dt])
        n = lib().axref_read(self.h, out.ctypes.data_as(ctypes.c_void_p), out.nbytes)
        if n < 0:
            raise RuntimeError(lib().axref_last_error().decode())
        return out

    # -- ops (each forwards to the reference's public function) -------------------------------
    def unary(self, op: str, fparam: float = 0.0):
        return RefTensor(lib().axref_unary(OP[op], self.h, fparam))

    def binary(self, op: str, other: "RefTensor"):
        return RefTensor(lib().axref_binary(OP[op], self.h, other.h))

    def binary_scalar(self, op: str, s):
        if isinstance(s, float):
            return RefTensor(lib().axref_binary_scalar_f32(OP[op], self.h, s))
        return RefTensor(lib().axref_binary_scalar_i32(OP[op], self.h, int(s)))

    def where(self, cond: "RefTensor", other):
        if isinstance(other, RefTensor):
            return RefTensor(lib().axref_where(cond.h, self.h, other.h))
        return RefTensor(lib().axref_where_scalar(self.h, cond.h, float(other)))

    def masked_fill(self, mask: "RefTensor", value: float):
        return RefTensor(lib().axref_masked_fill(self.h, mask.h, value))

    def reduce(self, op: str, axes: Sequence[int] = (), keep_dims: bool = False):
        arr = (ctypes.c_int * max(len(axes), 1))(*axes)
        return RefTensor(lib().axref_reduce(OP[op], self.h, arr, len(axes), int(keep_dims)))

    def argreduce(self, op: str, axis: int, keep_dims: bool = False):
        return RefTensor(lib().axref_argreduce(OP[op], self.h, axis, int(keep_dims)))

    def softmax(self, axis: int = -1, log: bool = False):
        return RefTensor(lib().axref_softmax(self.h, axis, int(log)))

    def layer_norm(self, w: "RefTensor", b: "RefTensor", axis: int = -1, eps: float = 1e-5):
        return RefTensor(lib().axref_layer_norm(self.h, w.h, b.h, axis, eps))

    def rms_norm(self, w: "RefTensor", axis: int = -1, eps: float = 1e-5):
        return RefTensor(lib().axref_rms_norm(self.h, w.h, axis, eps))

    def matmul(self, other: "RefTensor", ta: bool = False, tb: bool = False):
        return RefTensor(lib().axref_matmul(self.h, other.h, int(ta), int(tb)))

    def astype(self, dtype: str):
        return RefTensor(lib().axref_astype(self.h, DT[dtype]))

    def contiguous(self):
        return RefTensor(lib().axref_contiguous(self.h))

    def T(self):
        return RefTensor(lib().axref_transpose(self.h))

    def permute(self, axes: Sequence[int]):
        arr = (ctypes.c_int * len(axes))(*axes)
        return RefTensor(lib().axref_permute(self.h, arr, len(axes)))

    def reshape(self, shape: Sequence[int]):
        arr = (ctypes.c_int64 * max(len(shape), 1))(*shape)
        return RefTensor(lib().axref_reshape(self.h, arr, len(shape)))

    def gather(self, dim: int, idx: "RefTensor"):
        return RefTensor(lib().axref_gather(self.h, dim, idx.h))

    def index_select(self, dim: int, idx: "RefTensor"):
        return RefTensor(lib().axref_index_select(self.h, dim, idx.h))


def readme_chain(x: RefTensor) -> RefTensor:
    return RefTensor(lib().axref_readme_chain(x.h))


def config2_chain(a: RefTensor, b: RefTensor, c: RefTensor) -> RefTensor:
    return RefTensor(lib().axref_config2_chain(a.h, b.h, c.h))


def conv1d(x: RefTensor, w: RefTensor, b: Optional[RefTensor], stride=1, padding=0, dilation=1, groups=1):
    return RefTensor(lib().axref_conv1d(x.h, w.h, b.h if b else None, stride, padding, dilation, groups))


def sdpa(q: RefTensor, k: RefTensor, v: RefTensor, mask: Optional[RefTensor] = None, scale: float = -1.0,
         is_causal: bool = False):
    return RefTensor(lib().axref_sdpa(q.h, k.h, v.h, mask.h if mask else None, scale, int(is_causal)))


def pool(x: RefTensor, mode: str, kernel, stride, padding, count_include_pad: bool = True) -> RefTensor:
    n = len(kernel)
    arr = lambda v: (ctypes.c_int * n)(*[int(t) for t in v])  # noqa: E731
    return RefTensor(lib().axref_pool(x.h, 1 if mode == "avg" else 0, n, arr(kernel), arr(stride), arr(padding),
                                      int(count_include_pad)))


def sort(x: RefTensor, axis: int = -1, descending: bool = False) -> RefTensor:
    return RefTensor(lib().axref_sort(x.h, 0, axis, int(descending), 0))


def argsort(x: RefTensor, axis: int = -1, descending: bool = False) -> RefTensor:
    return RefTensor(lib().axref_sort(x.h, 1, axis, int(descending), 0))


def topk(x: RefTensor, k: int, axis: int = -1, largest: bool = True):
    return (RefTensor(lib().axref_sort(x.h, 2, axis, int(largest), k)), RefTensor(lib().axref_sort(x.h, 3, axis, int(largest), k)))


def masked_select(x: RefTensor, mask: RefTensor) -> RefTensor:
    return RefTensor(lib().axref_masked_select(x.h, mask.h))


def num_threads() -> int:
    return lib().axref_num_threads()


def set_num_threads(n: int) -> None:
    lib().axref_set_num_threads(n)
