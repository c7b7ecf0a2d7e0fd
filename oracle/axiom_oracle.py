"""oracle/axiom_oracle.py — CPU restatement (numpy + oracle/axiom_oracle.c) of the reference algorithms on
the hot path. TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg; never by axiom_b200/.

Pinned against the reference itself (oracle/_ref/libaxiom_ref.so, built from /root/reference by
oracle/Makefile) in tests/test_oracle_pin.py, and against the reference's own known-answer tests
transcribed in tests/golden/.  Every function cites the reference lines it restates.

Arrays are numpy arrays; bfloat16 is carried as uint16 bit patterns and tagged with dtype code BF16.
dtype / op codes are axiom::DType / axiom::ops::OpType values (include/axiom/dtype.hpp:170-192,
include/axiom/operations.hpp:55-208).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F16, BF16, F32, F64 = range(13)
NP_OF = {BOOL: np.bool_, I8: np.int8, I16: np.int16, I32: np.int32, I64: np.int64, U8: np.uint8, U16: np.uint16,
         U32: np.uint32, U64: np.uint64, F16: np.float16, BF16: np.uint16, F32: np.float32, F64: np.float64}
CODE_OF = {np.dtype(v): k for k, v in NP_OF.items() if k != BOOL and k != BF16}
CODE_OF[np.dtype(np.bool_)] = BOOL
CODE_OF[np.dtype(np.uint16)] = U16

OP = dict(add=0, sub=1, mul=2, div=3, pow=4, mod=5, eq=6, ne=7, lt=8, le=9, gt=10, ge=11, logical_and=12,
          logical_or=13, logical_xor=14, logical_not=15, bit_and=16, bit_or=17, bit_xor=18, shl=19, shr=20,
          maximum=21, minimum=22, atan2=23, hypot=24, neg=25, abs=26, sqrt=27, exp=28, log=29, sin=30, cos=31,
          tan=32, erf=33, sign=34, floor=35, ceil=36, trunc=37, round=38, reciprocal=39, square=40, cbrt=41,
          isnan=42, isinf=43, isfinite=44, relu=48, leaky_relu=49, silu=50, sigmoid=51, tanh=52, gelu=53,
          sum=56, mean=57, max=58, min=59, argmax=60, argmin=61, any=62, all=63, prod=64)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "axiom_oracle.c")
_clib: Optional[ctypes.CDLL] = None


def build_c_oracle(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", _SRC, "-o", _SO, "-lm"])
    return _SO


def clib() -> ctypes.CDLL:
    global _clib
    if _clib is None:
        L = ctypes.CDLL(build_c_oracle())
        fp = ctypes.POINTER(ctypes.c_float)
        L.axo_unary_f32.argtypes = [ctypes.c_int, fp, fp, ctypes.c_int64, ctypes.c_int]
        L.axo_binary_f32.argtypes = [ctypes.c_int, fp, fp, fp, ctypes.c_int64, ctypes.c_int]
        L.axo_softmax_rows_f32.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        _clib = L
    return _clib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# ---- bfloat16 helpers (include/axiom/bfloat16.hpp:16-27,41-46) ---------------------------------------
def bf16_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = x.view(np.uint32).astype(np.uint64)
    r = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16).astype(np.uint16)
    return np.where(np.isnan(x), np.uint16(0x7FC0), r)


def to_f32(a: np.ndarray, dt: int) -> np.ndarray:
    if dt == BF16:
        return bf16_to_f32(a)
    return a.astype(np.float32)


# ---- type rules -------------------------------------------------------------------------------------
_RANK = {BOOL: 0, I8: 1, U8: 2, I16: 3, U16: 4, I32: 5, U32: 6, I64: 7, U64: 8, F16: 9, BF16: 10, F32: 11, F64: 12}
_SIZE = {BOOL: 1, I8: 1, I16: 2, I32: 4, I64: 8, U8: 1, U16: 2, U32: 4, U64: 8, F16: 2, BF16: 2, F32: 4, F64: 8}


def promote_types(a: int, b: int) -> int:
    """src/tensor/operations.cpp:346-397"""
    if a == b:
        return a
    t1, t2 = (a, b) if _RANK[a] > _RANK[b] else (b, a)
    if t1 in (F16, BF16, F32, F64):
        return t1
    s1, s2 = t1 in (I8, I16, I32, I64), t2 in (I8, I16, I32, I64)
    z1, z2 = _SIZE[t1], _SIZE[t2]
    if s1 == s2:
        return t1
    if z1 > z2:
        return t1
    return {1: I16, 2: I32, 4: I64}.get(max(z1, z2), F64)


def cast(a: np.ndarray, src: int, dst: int) -> np.ndarray:
    """Tensor::astype (src/tensor/tensor.cpp:1656-1701, type_conversion.cpp): static_cast semantics;
    float->bf16 RNE with NaN->0x7FC0; anything->Bool is x != 0."""
    if src == dst:
        return a
    if dst == BOOL:
        return (to_f32(a, src) != 0) if src == BF16 else (a != 0)
    if src == BF16:
        a = bf16_to_f32(a)
    if dst == BF16:
        return f32_to_bf16(a.astype(np.float32))
    if np.issubdtype(a.dtype, np.floating) and np.issubdtype(NP_OF[dst], np.integer):
        a = np.trunc(a)
    with np.errstate(all="ignore"):
        return a.astype(NP_OF[dst])


# ---- elementwise ------------------------------------------------------------------------------------
_CMP = {"eq": np.equal, "ne": np.not_equal, "lt": np.less, "le": np.less_equal, "gt": np.greater,
        "ge": np.greater_equal}


def _c_unary(op: str, x: np.ndarray, lanes: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    rc = clib().axo_unary_f32(OP[op], _fp(x), _fp(out), x.size, lanes)
    assert rc == 0, op
    return out


def unary(op: str, x: np.ndarray, dt: int, lanes: int = 0, param: float = 0.01) -> Tuple[np.ndarray, int]:
    """CPUUnaryOperation (cpu_operations.cpp:824-1273). `lanes` = SIMD width of the reference host for the
    vector-body/scalar-tail split of transcendental fp32 ops (0 = polynomial everywhere)."""
    if op == "logical_not":
        return ~(cast(x, dt, BOOL)), BOOL
    if op in ("isnan", "isinf", "isfinite"):
        if dt in (F32, F64, F16):
            return getattr(np, op)(x), BOOL
        if dt == BF16:  # not is_floating_point_v, not float16_t -> constant (cpu_operations.hpp:806-845)
            return np.full(x.shape, op == "isfinite"), BOOL
        return np.full(x.shape, op == "isfinite"), BOOL
    if dt == F32:
        x = x.astype(np.float32)
        with np.errstate(all="ignore"):
            if op in ("exp", "log", "sin", "cos", "tan", "erf", "cbrt", "silu", "sigmoid", "tanh", "gelu"):
                return _c_unary(op, x, lanes).reshape(x.shape), F32
            r = {"neg": lambda: -x, "abs": lambda: np.abs(x), "sqrt": lambda: np.sqrt(x),
                 "sign": lambda: np.sign(x).astype(np.float32), "floor": lambda: np.floor(x),
                 "ceil": lambda: np.ceil(x), "trunc": lambda: np.trunc(x), "round": lambda: np.rint(x),
                 "reciprocal": lambda: (np.float32(1) / x), "square": lambda: x * x,
                 "relu": lambda: np.where(x > 0, x, np.float32(0)),
                 "leaky_relu": lambda: np.where(x > 0, x, np.float32(param) * x)}[op]()
        return r.astype(np.float32), F32
    if dt in (BF16, F16):
        # scalar functors on half types: evaluate in double (cpu_operations.hpp:383-700), round once to float,
        # then to the half type
        xf = to_f32(x, dt)
        xd = xf.astype(np.float64)
        with np.errstate(all="ignore"):
            if op in ("neg", "abs", "relu", "square", "sign", "reciprocal", "leaky_relu", "floor", "ceil", "trunc",
                      "round"):
                if dt == BF16 and op in ("floor", "ceil", "trunc", "round"):
                    rf = xf  # identity for bfloat16 (cpu_operations.hpp:728-770)
                elif op == "leaky_relu":
                    slope = to_f32(cast(np.array([param], np.float32), F32, dt), dt)[0]
                    rf = np.where(xf > 0, xf, slope * xf)
                elif op == "round":
                    rf = np.where(xf >= 0, np.floor(xf + np.float32(0.5)), np.ceil(xf - np.float32(0.5)))  # std::round
                else:
                    rf = {"neg": lambda: -xf, "abs": lambda: np.abs(xf), "relu": lambda: np.where(xf > 0, xf, 0),
                          "square": lambda: xf * xf, "sign": lambda: np.sign(xf),
                          "reciprocal": lambda: np.float32(1) / xf, "floor": lambda: np.floor(xf),
                          "ceil": lambda: np.ceil(xf), "trunc": lambda: np.trunc(xf)}[op]()
                return cast(np.asarray(rf, np.float32), F32, dt), dt
            import scipy.special as sp
            e = np.exp(-np.abs(xd))
            rd = {"sqrt": lambda: np.sqrt(xd), "exp": lambda: np.exp(xd), "log": lambda: np.log(xd),
                  "sin": lambda: np.sin(xd), "cos": lambda: np.cos(xd), "tan": lambda: np.tan(xd),
                  "erf": lambda: sp.erf(xd), "cbrt": lambda: np.cbrt(xd), "tanh": lambda: np.tanh(xd),
                  "gelu": lambda: 0.5 * xd * (1.0 + sp.erf(xd / 1.4142135623730951)),
                  "sigmoid": lambda: np.where(xd >= 0, 1.0 / (1.0 + e), e / (1.0 + e)),
                  "silu": lambda: np.where(xd >= 0, xd / (1.0 + e), (xd * e) / (1.0 + e))}[op]()
        return cast(rd.astype(np.float32), F32, dt), dt
    if dt == F64:
        import scipy.special as sp
        with np.errstate(all="ignore"):
            r = {"neg": lambda: -x, "abs": lambda: np.abs(x), "sqrt": lambda: np.sqrt(x), "exp": lambda: np.exp(x),
                 "log": lambda: np.log(x), "sin": lambda: np.sin(x), "cos": lambda: np.cos(x),
                 "tan": lambda: np.tan(x), "erf": lambda: sp.erf(x), "tanh": lambda: np.tanh(x),
                 "sign": lambda: np.sign(x), "floor": lambda: np.floor(x), "ceil": lambda: np.ceil(x),
                 "trunc": lambda: np.trunc(x), "round": lambda: np.rint(x), "reciprocal": lambda: 1.0 / x,
                 "square": lambda: x * x, "relu": lambda: np.where(x > 0, x, 0.0)}[op]()
        return r.astype(np.float64), F64
    # integers / bool
    if dt == BOOL:
        return {"neg": lambda: ~x}.get(op, lambda: x)(), BOOL
    unsigned = dt in (U8, U16, U32, U64)
    with np.errstate(all="ignore"):
        r = {"neg": lambda: (0 - x).astype(x.dtype), "abs": lambda: x if unsigned else np.abs(x),
             "sign": lambda: (x > 0).astype(x.dtype) if unsigned else np.sign(x).astype(x.dtype),
             "square": lambda: (x * x).astype(x.dtype), "relu": lambda: np.where(x > 0, x, 0).astype(x.dtype),
             "floor": lambda: x, "ceil": lambda: x, "trunc": lambda: x, "round": lambda: x}[op]()
    return r, dt


def binary(op: str, a: np.ndarray, adt: int, b: np.ndarray, bdt: int, lanes: int = 0,
           simd_path: bool = True) -> Tuple[np.ndarray, int]:
    """CPUBinaryOperation::execute_binary (cpu_operations.cpp:28-88, 109-583). simd_path=False forces the
    scalar-functor algorithms the reference uses when an operand is broadcast or non-contiguous."""
    if op in _CMP or op in ("logical_and", "logical_or", "logical_xor"):
        # both operands are converted to Float32 first (cpu_operations.cpp:115-126,286-296)
        af, bf = to_f32(a, adt) if adt == BF16 else a.astype(np.float32), \
            to_f32(b, bdt) if bdt == BF16 else b.astype(np.float32)
        if op in _CMP:
            return _CMP[op](af, bf), BOOL
        ab, bb = af != 0, bf != 0
        return {"logical_and": ab & bb, "logical_or": ab | bb, "logical_xor": ab ^ bb}[op], BOOL
    rdt = promote_types(adt, bdt)
    if rdt == BOOL:  # arithmetic on two Bool tensors runs in float and is cast back (T = bool path)
        r, _ = binary(op, a.astype(np.float32), F32, b.astype(np.float32), F32, lanes)
        return r != 0, BOOL
    x, y = cast(a, adt, rdt), cast(b, bdt, rdt)
    if rdt in (BF16, F16):
        xf, yf = to_f32(x, rdt), to_f32(y, rdt)
        with np.errstate(all="ignore"):
            if op in ("pow", "hypot", "atan2"):
                rd = {"pow": np.power, "hypot": np.hypot, "atan2": np.arctan2}[op](xf.astype(np.float64),
                                                                                     yf.astype(np.float64))
                rf = rd.astype(np.float32)
            else:
                rf, _ = binary(op, xf, F32, yf, F32)
        return cast(np.asarray(rf, np.float32), F32, rdt), rdt
    with np.errstate(all="ignore"):
        if rdt in (F32, F64):
            same_shape = (np.shape(x) == np.shape(y))
            if op in ("pow", "hypot") and rdt == F32 and same_shape and simd_path:
                # Highway kernels only on the same-shape contiguous path (cpu_operations.cpp:229-262);
                # broadcast / strided operands use the scalar libm functors (:516-583)
                xb, yb = np.ascontiguousarray(x), np.ascontiguousarray(y)
                out = np.empty_like(xb)
                clib().axo_binary_f32(OP[op], _fp(xb), _fp(yb), _fp(out), out.size, lanes)
                return out, rdt
            f = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "pow": np.power,
                 "mod": np.fmod, "maximum": lambda p, q: np.where(p < q, q, p),
                 "minimum": lambda p, q: np.where(q < p, q, p), "atan2": np.arctan2, "hypot": np.hypot}[op]
            return np.asarray(f(x, y)).astype(NP_OF[rdt]), rdt
        # integers
        if op == "div":
            q = np.where(y == 0, 0, np.trunc(x.astype(np.float64) / np.where(y == 0, 1, y).astype(np.float64)))
            return q.astype(NP_OF[rdt]), rdt
        if op == "mod":
            yy = np.where(y == 0, 1, y)
            r = np.fmod(x, yy)
            return np.where(y == 0, 0, r).astype(NP_OF[rdt]), rdt
        f = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "maximum": np.maximum, "minimum": np.minimum,
             "bit_and": np.bitwise_and, "bit_or": np.bitwise_or, "bit_xor": np.bitwise_xor,
             "shl": np.left_shift, "shr": np.right_shift}[op]
        return np.asarray(f(x, y)).astype(NP_OF[rdt]), rdt


def where(cond: np.ndarray, cdt: int, a: np.ndarray, adt: int, b: np.ndarray, bdt: int) -> Tuple[np.ndarray, int]:
    """CPUWhereOperation (cpu_operations.cpp:2408-2589): any non-zero BYTE of cond is true (:2503-2513)."""
    rdt = promote_types(adt, bdt)
    if cdt == BOOL:
        c = cond.astype(bool)
    else:
        raw = np.ascontiguousarray(cond)
        c = raw.view(np.uint8).reshape(raw.shape + (raw.dtype.itemsize,)).any(axis=-1)
    return np.where(c, cast(a, adt, rdt), cast(b, bdt, rdt)), rdt


def masked_fill(x: np.ndarray, xdt: int, mask: np.ndarray, value: float) -> Tuple[np.ndarray, int]:
    """cpu_operations.cpp:2596-2700; the value goes through a Float32 {1} tensor (operations.cpp:1585-1592)."""
    v = cast(np.array([value], np.float32), F32, xdt)
    return np.where(mask.astype(bool), v, x), xdt


# ---- reductions -------------------------------------------------------------------------------------
def _tree_lanes(acc: np.ndarray, f) -> np.ndarray:
    """hn::ReduceSum/Max/Min over one vector, in Highway's shuffle order: ReduceAcrossBlocks
    (x86_512-inl.h:7641-7646: v+=SwapAdjacentBlocks, v+=ReverseBlocks; AVX2 generic_ops-inl.h:1041-1045) then
    ReduceWithinBlocks on a 128-bit block (generic_ops-inl.h:1050-1071)."""
    L = acc.shape[-1]
    lpb = max(1, min(L, 16 // acc.dtype.itemsize))  # lanes per 128-bit block
    nb = max(1, L // lpb)
    b = acc.reshape(acc.shape[:-1] + (nb, lpb))
    if nb == 4:
        y = f(f(b[..., 0, :], b[..., 1, :]), f(b[..., 3, :], b[..., 2, :]))
    elif nb == 2:
        y = f(b[..., 0, :], b[..., 1, :])
    else:
        y = b[..., 0, :]
        for k in range(1, nb):
            y = f(y, b[..., k, :])
    if lpb == 4:
        return f(f(y[..., 0], y[..., 3]), f(y[..., 1], y[..., 2]))
    if lpb == 2:
        return f(y[..., 0], y[..., 1])
    r = y[..., 0]
    for k in range(1, lpb):
        r = f(r, y[..., k])
    return r


def _reduce_rows(x2: np.ndarray, op: str, lanes: int) -> np.ndarray:
    """Per-row SIMD reduce (hwy_kernels-inl.h:678-767): `lanes` lane-strided serial partials, horizontal
    tree, then the scalar tail. x2: (rows, n) in the accumulation dtype."""
    rows, n = x2.shape
    f = {"sum": np.add, "prod": np.multiply, "max": np.maximum, "min": np.minimum}[op]
    ident = {"sum": 0, "prod": 1, "max": np.finfo(x2.dtype).min if x2.dtype.kind == "f" else np.iinfo(x2.dtype).min,
             "min": np.finfo(x2.dtype).max if x2.dtype.kind == "f" else np.iinfo(x2.dtype).max}[op]
    L = max(int(lanes), 1)
    nb = (n // L) * L
    acc = np.full((rows, L), ident, dtype=x2.dtype)
    with np.errstate(all="ignore"):
        for i in range(0, nb, L):
            acc = f(acc, x2[:, i:i + L])
        if op == "prod":  # horizontal product is a serial loop over lanes (hwy_kernels-inl.h:752-758)
            r = np.full((rows,), 1, dtype=x2.dtype)
            for j in range(L):
                r = r * acc[:, j]
        else:
            r = _tree_lanes(acc, f)
        for i in range(nb, n):
            r = f(r, x2[:, i])
    return r


def reduce_result_dtype(op: str, dt: int) -> int:
    """What the caller of ops::sum/mean/... observes (infer_reduction_dtype graph_registry.cpp:265-291 +
    executor promotion graph_executor.cpp:396-403): Mean of half -> Float32 (value already rounded to half)."""
    if op == "mean" and dt in (F16, BF16):
        return F32
    return dt


def reduce(op: str, x: np.ndarray, dt: int, axes: Sequence[int] = (), keep_dims: bool = False,
           lanes: int = 16) -> Tuple[np.ndarray, int]:
    """CPUReductionOperation::execute_reduction_typed (cpu_operations.cpp:1419-1712)."""
    nd = x.ndim
    axes = sorted(set(int(a) for a in axes)) if len(axes) else list(range(nd))
    half = dt in (F16, BF16)
    xa = to_f32(x, dt) if half else x  # fp32 accumulation for half types (:1422-1424)
    adt = F32 if half else dt
    base = "sum" if op == "mean" else op
    n_red = int(np.prod([x.shape[a] for a in axes])) if axes else 1
    full = len(axes) == nd
    simd_ok = adt in (F32, F64, I32, I64) and base in ("sum", "max", "min", "prod")
    if op in ("any", "all"):
        xb = xa != 0
        r = xb.any(axis=tuple(axes), keepdims=keep_dims) if op == "any" else xb.all(axis=tuple(axes),
                                                                                      keepdims=keep_dims)
        if full and not keep_dims:
            r = np.asarray(r).reshape(1)
        return cast(np.asarray(r), BOOL, dt), dt
    with np.errstate(all="ignore"):
        if full and simd_ok:
            r = _reduce_rows(np.ascontiguousarray(xa).reshape(1, -1), base, lanes)
            r = r.reshape([1] * nd if keep_dims else [1])
        elif len(axes) == 1 and axes[0] == nd - 1 and nd >= 2 and simd_ok:
            r = _reduce_rows(np.ascontiguousarray(xa).reshape(-1, x.shape[-1]), base, lanes)
            r = r.reshape(x.shape[:-1] + ((1,) if keep_dims else ()))
        else:
            # first-axis SIMD accumulate and the recursive scalar fallback both add elements of each output
            # in increasing index order (cpu_operations.cpp:1620-1711)
            keep = [a for a in range(nd) if a not in axes]
            xt = np.transpose(xa, keep + axes).reshape([x.shape[a] for a in keep] + [n_red])
            f = {"sum": np.add, "prod": np.multiply, "max": lambda p, q: np.where(p < q, q, p),
                 "min": lambda p, q: np.where(q < p, q, p)}[base]
            if axes == [0] and nd >= 2 and adt in (F32, F64):
                r = xt[..., 0].copy()
                start = 1
            else:
                ident = {"sum": 0, "prod": 1,
                         "max": np.finfo(xt.dtype).min if xt.dtype.kind == "f" else np.iinfo(xt.dtype).min if xt.dtype.kind in "iu" else False,
                         "min": np.finfo(xt.dtype).max if xt.dtype.kind == "f" else np.iinfo(xt.dtype).max if xt.dtype.kind in "iu" else True}[base]
                r = np.full(xt.shape[:-1], ident, dtype=xt.dtype)
                start = 0
            for j in range(start, n_red):
                r = f(r, xt[..., j]).astype(xt.dtype)
            if keep_dims:
                shp = [1 if a in axes else x.shape[a] for a in range(nd)]
                r = r.reshape(shp)
            elif r.ndim == 0:
                r = r.reshape(1)
        if op == "mean":
            r = (r / np.asarray(n_red, dtype=r.dtype)).astype(r.dtype) if r.dtype.kind == "f" else \
                np.trunc(r / n_red).astype(r.dtype)
    r = np.asarray(r)
    if half:
        rb = cast(r.astype(np.float32), F32, dt)  # rounded once to the half type (:1613-1614)
        if op == "mean":
            return to_f32(rb, dt), F32
        return rb, dt
    return r.astype(NP_OF[adt]), adt


def argreduce(op: str, x: np.ndarray, dt: int, axis: int, keep_dims: bool = False) -> np.ndarray:
    """CPUArgMax/ArgMin (cpu_operations.cpp:2157-2350): strict compare from lowest()/max() => first extreme wins,
    NaN never wins, all-NaN (or all == lowest) -> 0; axis < 0 => flattened, rank-0 result."""
    xf = to_f32(x, dt) if dt == BF16 else x
    if axis < 0:
        xf = xf.reshape(-1)
        axis = 0
        flat = True
    else:
        flat = False
    xm = np.moveaxis(xf, axis, -1)
    if xm.dtype.kind == "f":
        fill = -np.inf if op == "argmax" else np.inf
        clean = np.where(np.isnan(xm), fill, xm)
        lowest = np.finfo(xm.dtype).min if op == "argmax" else np.finfo(xm.dtype).max
    else:
        clean = xm
        lowest = None
    idx = np.argmax(clean, axis=-1) if op == "argmax" else np.argmin(clean, axis=-1)
    if lowest is not None:
        best = np.take_along_axis(clean, idx[..., None], axis=-1)[..., 0]
        # values that never beat the initial best (-inf, or NaN everywhere) leave best_idx at 0
        never = (best <= lowest) if op == "argmax" else (best >= lowest)
        idx = np.where(never, 0, idx)
    idx = idx.astype(np.int64)
    if keep_dims:
        idx = np.expand_dims(idx, axis if not flat else 0)
        if flat:
            idx = idx.reshape([1] * x.ndim)
    return idx


def softmax(x: np.ndarray, dt: int, axis: int = -1, log: bool = False) -> np.ndarray:
    """CPUSoftmaxOperation (cpu_operations.cpp:3067-3241): per slice max, libm exp, SERIAL fp32 sum, divide;
    half dtypes round-trip through Float32."""
    xf = to_f32(x, dt) if dt in (BF16, F16) else x
    if dt == F64:
        m = xf.max(axis=axis, keepdims=True)
        e = np.exp(xf - m)
        s = e.sum(axis=axis, keepdims=True)
        return (xf - m - np.log(s)) if log else e / s
    xm = np.ascontiguousarray(np.moveaxis(xf.astype(np.float32), axis, -1))
    rows = xm.reshape(-1, xm.shape[-1])
    out = np.empty_like(rows)
    clib().axo_softmax_rows_f32(_fp(rows), _fp(out), rows.shape[0], rows.shape[1], int(log))
    r = np.moveaxis(out.reshape(xm.shape), -1, axis)
    return cast(np.ascontiguousarray(r), F32, dt) if dt in (BF16, F16) else np.ascontiguousarray(r)


def layer_norm(x: np.ndarray, dt: int, w: np.ndarray, wdt: int, b: np.ndarray, bdt: int, axis: int = -1,
               eps: float = 1e-5, lanes: int = 16) -> Tuple[np.ndarray, int]:
    """ops::layer_norm on CPU = 9 composed ops (src/tensor/operations.cpp:1803-1817): mean, sub, mul, mean,
    +eps, sqrt, div, *w, +b — each op rounded to its dtype. Half input => Float32 from the first mean on."""
    ax = axis % x.ndim
    mean, mdt = reduce("mean", x, dt, [ax], True, lanes)
    xc, cdt = binary("sub", x, dt, mean, mdt)
    sq, sdt = binary("mul", xc, cdt, xc, cdt)
    var, vdt = reduce("mean", sq, sdt, [ax], True, lanes)
    ve, edt = binary("add", var, vdt, np.array([eps], np.float32), F32)
    sd, ddt = unary("sqrt", ve, edt)
    nrm, ndt = binary("div", xc, cdt, sd, ddt)
    sc, scdt = binary("mul", nrm, ndt, w, wdt)
    return binary("add", sc, scdt, b, bdt)


def rms_norm(x: np.ndarray, dt: int, w: np.ndarray, wdt: int, axis: int = -1, eps: float = 1e-5,
             lanes: int = 16) -> Tuple[np.ndarray, int]:
    """ops::rms_norm CPU composition (src/tensor/operations.cpp:1841-1857)."""
    ax = axis % x.ndim
    sq, sdt = binary("mul", x, dt, x, dt)
    ms, mdt = reduce("mean", sq, sdt, [ax], True, lanes)
    me, edt = binary("add", ms, mdt, np.array([eps], np.float32), F32)
    rms, rdt = unary("sqrt", me, edt)
    nrm, ndt = binary("div", x, dt, rms, rdt)
    return binary("mul", nrm, ndt, w, wdt)


def matmul(a: np.ndarray, adt: int, b: np.ndarray, bdt: int, ta: bool = False, tb: bool = False) -> Tuple[np.ndarray, int]:
    """CPUMatMulOperation (cpu_operations.cpp:1886-2150). fp32/fp64 go to sgemm/dgemm (any summation order:
    compared with tolerance); the governing spec for bf16 is 'vs an fp32 reference' (north_star), so bf16
    inputs are multiplied in fp32 (fp64 accumulate) and rounded once."""
    rdt = promote_types(adt, bdt)
    x = to_f32(a, adt) if adt in (BF16, F16) else a
    y = to_f32(b, bdt) if bdt in (BF16, F16) else b
    if ta:
        x = np.swapaxes(x, -1, -2) if x.ndim >= 2 else x
    if tb:
        y = np.swapaxes(y, -1, -2) if y.ndim >= 2 else y
    r = np.matmul(x.astype(np.float64), y.astype(np.float64))
    if rdt in (BF16, F16):
        return cast(np.asarray(r, np.float32), F32, rdt), rdt
    return np.asarray(r).astype(NP_OF[rdt]), rdt


# ---- spatial operators (small-case restatements in plain loops; fp32) ---------------------------------------------------
def conv2d_f32(x, w, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1):
    """ops::conv2d CPU loop (src/tensor/operations.cpp:3263-3318): fp64 accumulator over (ic, fh, fw), + bias, one rounding."""
    n, cin, h, wd = x.shape
    cout, cg, kh, kw = w.shape
    oh = (h + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    ow = (wd + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    out = np.zeros((n, cout, oh, ow), np.float32)
    cog = cout // groups
    xd, wdd = x.astype(np.float64), w.astype(np.float64)
    for b in range(n):
        for oc in range(cout):
            g = oc // cog
            for i in range(oh):
                for j in range(ow):
                    acc = 0.0
                    for ic in range(cg):
                        for fh in range(kh):
                            ih = i * stride[0] - padding[0] + fh * dilation[0]
                            if ih < 0 or ih >= h:
                                continue
                            for fw in range(kw):
                                iw = j * stride[1] - padding[1] + fw * dilation[1]
                                if 0 <= iw < wd:
                                    acc += xd[b, g * cg + ic, ih, iw] * wdd[oc, ic, fh, fw]
                    if bias is not None:
                        acc += float(bias[oc])
                    out[b, oc, i, j] = np.float32(acc)
    return out


def conv_transpose2d_f32(x, w, bias=None, stride=(1, 1), padding=(0, 0), output_padding=(0, 0), dilation=(1, 1), groups=1):
    """ops::conv_transpose2d CPU loop (src/tensor/operations.cpp:3590-3690; 1-d: :3425-3486 with H = 1): scatter of
    out += float(double(in) * double(w)) in fp32, order (ic, ih, iw, oc, fh, fw); bias added last in fp32."""
    n, cin, h, wd = x.shape
    _, cog, kh, kw = w.shape
    cout = cog * groups
    cig = cin // groups
    oh = (h - 1) * stride[0] - 2 * padding[0] + dilation[0] * (kh - 1) + output_padding[0] + 1
    ow = (wd - 1) * stride[1] - 2 * padding[1] + dilation[1] * (kw - 1) + output_padding[1] + 1
    out = np.zeros((n, cout, oh, ow), np.float32)
    for b in range(n):
        for g in range(groups):
            for ic in range(g * cig, (g + 1) * cig):
                for ih in range(h):
                    for iw in range(wd):
                        v = float(x[b, ic, ih, iw])
                        for oc in range(cog):
                            for fh in range(kh):
                                o_h = ih * stride[0] - padding[0] + fh * dilation[0]
                                if o_h < 0 or o_h >= oh:
                                    continue
                                for fw in range(kw):
                                    o_w = iw * stride[1] - padding[1] + fw * dilation[1]
                                    if 0 <= o_w < ow:
                                        out[b, g * cog + oc, o_h, o_w] += np.float32(v * float(w[ic, oc, fh, fw]))
    if bias is not None:
        out += bias.astype(np.float32).reshape(1, cout, 1, 1)
    return out


def interpolate2d_f32(x, size, mode="nearest", align_corners=False):
    """ops::interpolate CPU loop (src/tensor/operations.cpp:3826-3955): coordinates and weights in fp32, every product
    and sum rounded on its own (no FMA contraction in the reference's x86-64 build)."""
    f = np.float32
    n, c, h, w = x.shape
    oh, ow = size
    out = np.zeros((n, c, oh, ow), np.float32)

    def src(o, i, t):
        if align_corners and t > 1:
            return f(f(o) * f(i - 1)) / f(t - 1)
        return f(f(f(f(o) + f(0.5)) * f(i)) / f(t)) - f(0.5)

    def cubic(v):
        v = f(abs(v))
        if v <= 1:
            return f(f(f(f(f(f(1.25) * v) * v) * v) - f(f(f(2.25) * v) * v)) + f(1))
        if v < 2:
            return f(f(f(f(f(f(-0.75) * v) * v) * v) - f(f(f(-3.75) * v) * v)) + f(f(-6) * v)) - f(-3)
        return f(0)

    for i in range(oh):
        for j in range(ow):
            if mode == "nearest":
                sh = min(int(f(f(i) * f(h)) / f(oh)), h - 1)
                sw = min(int(f(f(j) * f(w)) / f(ow)), w - 1)
                out[:, :, i, j] = x[:, :, sh, sw]
                continue
            fh, fw = src(i, h, oh), src(j, w, ow)
            if mode == "bilinear":
                fh, fw = max(f(0), min(fh, f(h - 1))), max(f(0), min(fw, f(w - 1)))
                h0, w0 = int(fh), int(fw)
                h1, w1 = min(h0 + 1, h - 1), min(w0 + 1, w - 1)
                wh, ww = f(fh - f(h0)), f(fw - f(w0))
                ah, aw = f(f(1) - wh), f(f(1) - ww)
                val = f(f(ah * aw) * x[:, :, h0, w0]) + f(f(ah * ww) * x[:, :, h0, w1])
                val = f(val) + f(f(wh * aw) * x[:, :, h1, w0])
                out[:, :, i, j] = f(val) + f(f(wh * ww) * x[:, :, h1, w1])
            else:
                hf, wf = int(np.floor(fh)), int(np.floor(fw))
                val = np.zeros((n, c), np.float32)
                for dh in range(-1, 3):
                    sh = max(0, min(hf + dh, h - 1))
                    a = cubic(f(fh - f(hf + dh)))
                    for dw in range(-1, 3):
                        sw = max(0, min(wf + dw, w - 1))
                        bq = cubic(f(fw - f(wf + dw)))
                        val = (val + (f(a * bq) * x[:, :, sh, sw]).astype(np.float32)).astype(np.float32)
                out[:, :, i, j] = val
    return out


# ---- pooling / sort / topk / masked_select ---------------------------------------------------------
def _from_f64(v: np.ndarray, dt: int) -> np.ndarray:
    """static_cast<T>(double) of the reference's element types: half types go through float (bfloat16.hpp:29,
    float16.hpp:16)."""
    if dt == F64:
        return v.astype(np.float64)
    f = v.astype(np.float32)
    if dt == F32:
        return f
    if dt == BF16:
        return f32_to_bf16(f)
    return f.astype(np.float16)


def pool(x: np.ndarray, dt: int, mode: str, kernel: Sequence[int], stride: Optional[Sequence[int]] = None,
         padding: Optional[Sequence[int]] = None, count_include_pad: bool = True) -> np.ndarray:
    """ops::max_pool{1,2,3}d / avg_pool{1,2,3}d on (N, C, spatial...) (src/tensor/operations.cpp:2367-2880): the window is
    walked kd, kh, kw ascending; max starts at -inf in double and takes `val > max` (NaN never wins); avg sums in double,
    counts in-bounds taps (+ padded ones with count_include_pad) and divides once; one cast to the element type."""
    nsp = x.ndim - 2
    k = list(kernel)
    s = list(stride) if stride else list(k)
    p = list(padding) if padding else [0] * nsp
    xf = to_f32(x, dt).astype(np.float64) if dt != F64 else x.astype(np.float64)
    k3, s3, p3 = [1] * (3 - nsp) + k, [1] * (3 - nsp) + s, [0] * (3 - nsp) + p
    x5 = xf.reshape(x.shape[:2] + (1,) * (3 - nsp) + x.shape[2:])
    insz = x5.shape[2:]
    osz = [(insz[d] + 2 * p3[d] - k3[d]) // s3[d] + 1 for d in range(3)]  # pool_output_size, :2360-2363
    acc = np.full(x5.shape[:2] + tuple(osz), -np.inf if mode == "max" else 0.0, np.float64)
    cnt = np.zeros(tuple(osz), np.int64)
    grids = [np.arange(osz[d]) * s3[d] - p3[d] for d in range(3)]
    for kd in range(k3[0]):
        for kh in range(k3[1]):
            for kw in range(k3[2]):
                pos = [grids[0] + kd, grids[1] + kh, grids[2] + kw]
                ok = [(pos[d] >= 0) & (pos[d] < insz[d]) for d in range(3)]
                valid = ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]
                cl = [np.clip(pos[d], 0, insz[d] - 1) for d in range(3)]
                v = x5[:, :, cl[0][:, None, None], cl[1][None, :, None], cl[2][None, None, :]]
                if mode == "max":
                    with np.errstate(invalid="ignore"):
                        acc = np.where(valid[None, None] & (v > acc), v, acc)
                else:
                    acc = np.where(valid[None, None], acc + v, acc)
                    cnt += valid.astype(np.int64) if not count_include_pad else 1
    if mode != "max":
        with np.errstate(invalid="ignore", divide="ignore"):
            acc = np.where(cnt[None, None] > 0, acc / np.maximum(cnt, 1)[None, None], 0.0)
    out = _from_f64(acc, dt)
    return out.reshape(x.shape[:2] + tuple(osz[3 - nsp:]))


def sort_order(x: np.ndarray, axis: int = -1, descending: bool = False) -> np.ndarray:
    """Positions that sort `x` along `axis` (ops::argsort, src/tensor/operations.cpp:4088-4127, std::sort by `<` or `>` on
    float32 / float64 / int32 / int64 keys; every other dtype is converted to float32 first, :4218-4234). std::sort leaves
    the order of EQUAL keys unspecified; the restatement (and the CUDA kernel) fix it as ascending position — compare
    indices only where keys are distinct, values everywhere."""
    if descending:
        # stable ascending order of the reversed slice, mapped back: largest first, ties by ascending position
        n = x.shape[axis]
        r = np.argsort(np.flip(x, axis), axis=axis, kind="stable")
        return np.flip(n - 1 - r, axis)
    return np.argsort(x, axis=axis, kind="stable")


def sort(x: np.ndarray, axis: int = -1, descending: bool = False, k: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(values, indices) of ops::sort / argsort / topk (:4053-4257); k limits the result to the first k along `axis`
    (std::partial_sort + the first k copied out, :4148-4193 — they come out sorted whether or not `sorted` is set)."""
    key = x if x.dtype in (np.float32, np.float64, np.int32, np.int64) else x.astype(np.float32)
    order = sort_order(key, axis, descending)
    if k is not None:
        order = np.take(order, np.arange(k), axis=axis)
    return np.take_along_axis(key, order, axis=axis), order.astype(np.int64)


def masked_select(x: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """ops::masked_select (src/backends/cpu/cpu_operations.cpp:2719-2776): operands broadcast together, an element is
    selected when the FIRST BYTE of its mask element is non-zero (:2738-2741 reads a uint8 whatever the mask dtype),
    row-major order, 1-D result of the input's dtype."""
    shape = np.broadcast_shapes(x.shape, mask.shape)
    xb, mb = np.broadcast_to(x, shape), np.broadcast_to(mask, shape)
    first = np.ascontiguousarray(mb).view(np.uint8).reshape(shape + (mask.dtype.itemsize,))[..., 0] != 0
    return np.ascontiguousarray(xb)[first]


def conv2d_f64(x, w, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1)):
    """The same sum as conv2d_f32 (groups == 1), accumulated in float64 tap by tap with whole-array operations, for problems
    too large for the scalar loops; returned in float64 together with sum |x| |w| per output (the scale of the rounding
    error of any fp32 evaluation order). Pinned against conv2d_f32 on a small case in tests/test_spatial_gpu.py."""
    n, cin, h, wd = x.shape
    cout, cg, kh, kw = w.shape
    oh = (h + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    ow = (wd + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    xp = np.zeros((n, cin, h + 2 * padding[0], wd + 2 * padding[1]), np.float64)
    xp[:, :, padding[0]:padding[0] + h, padding[1]:padding[1] + wd] = x
    wd64 = w.astype(np.float64)
    out = np.zeros((n, cout, oh, ow), np.float64)
    mag = np.zeros((n, cout, oh, ow), np.float64)
    for fh in range(kh):
        for fw in range(kw):
            h0, w0 = fh * dilation[0], fw * dilation[1]
            win = xp[:, :, h0:h0 + (oh - 1) * stride[0] + 1:stride[0], w0:w0 + (ow - 1) * stride[1] + 1:stride[1]]
            out += np.einsum("oc,nchw->nohw", wd64[:, :, fh, fw], win, optimize=True)
            mag += np.einsum("oc,nchw->nohw", np.abs(wd64[:, :, fh, fw]), np.abs(win), optimize=True)
    if bias is not None:
        out += np.asarray(bias, np.float64).reshape(1, -1, 1, 1)
        mag += np.abs(np.asarray(bias, np.float64)).reshape(1, -1, 1, 1)
    return out, mag
