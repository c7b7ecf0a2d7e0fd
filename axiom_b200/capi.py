"""ctypes binding of include/axb.h — the only way Python (tests, bench) reaches the CUDA backend.

The product is the C-ABI library `axiom_b200/libaxb.so`; this module is harness plumbing. It fails loudly
when the library is missing: there is NO CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaxb.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "axb.h")

AXB_MAX_DIMS = 8

# status codes
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_OOM, ERR_SHAPE = range(6)

# dtypes == axiom::DType
BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F16, BF16, F32, F64, C64, C128 = range(15)
DTYPE_SIZE = {BOOL: 1, I8: 1, I16: 2, I32: 4, I64: 8, U8: 1, U16: 2, U32: 4, U64: 8, F16: 2, BF16: 2, F32: 4,
              F64: 8, C64: 8, C128: 16}

# ops == axiom::ops::OpType
OP: Dict[str, int] = dict(
    add=0, sub=1, mul=2, div=3, pow=4, mod=5, eq=6, ne=7, lt=8, le=9, gt=10, ge=11,
    logical_and=12, logical_or=13, logical_xor=14, logical_not=15,
    bit_and=16, bit_or=17, bit_xor=18, shl=19, shr=20, maximum=21, minimum=22, atan2=23, hypot=24,
    neg=25, abs=26, sqrt=27, exp=28, log=29, sin=30, cos=31, tan=32, erf=33, sign=34, floor=35, ceil=36,
    trunc=37, round=38, reciprocal=39, square=40, cbrt=41, isnan=42, isinf=43, isfinite=44,
    relu=48, leaky_relu=49, silu=50, sigmoid=51, tanh=52, gelu=53, softmax=54, log_softmax=55,
    sum=56, mean=57, max=58, min=59, argmax=60, argmin=61, any=62, all=63, prod=64,
    matmul=65, batch_matmul=66, where=67, masked_fill=68, gather=70, index_select=72,
    layer_norm=75, rms_norm=76, cast=78)

SRC_NONE = -(2 ** 31)
SRC_IMM = -(2 ** 31) + 1


def SRC_VALUE(k: int) -> int:
    return -1 - k


class AxbTensor(ctypes.Structure):
    _fields_ = [("base", ctypes.c_void_p), ("byte_offset", ctypes.c_int64), ("ndim", ctypes.c_int32),
                ("dtype", ctypes.c_int32), ("shape", ctypes.c_int64 * AXB_MAX_DIMS),
                ("strides", ctypes.c_int64 * AXB_MAX_DIMS)]


class AxbInstr(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int32), ("dtype", ctypes.c_int32), ("src", ctypes.c_int32 * 3),
                ("imm", ctypes.c_double)]


class AxbAllocStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("bytes_in_use", "bytes_cached", "peak_bytes_in_use", "n_cuda_malloc", "n_cache_hits", "n_allocs",
                 "n_splits", "n_merges", "n_cross_stream")]


class AxbTransferStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("h2d_bytes", "h2d_calls", "d2h_bytes", "d2h_calls", "d2d_bytes", "d2d_calls")]


class AxbPlanCacheStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("entries", "hits", "misses", "evictions")]


class AxbHalfLutStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("tables", "builds", "hits", "launches")]


class AxbJitStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("mem_hits", "disk_hits", "compiles")]


class AxbMatmulEpilogue(ctypes.Structure):
    _fields_ = [("bias", ctypes.POINTER(AxbTensor)), ("activation", ctypes.c_int32), ("flags", ctypes.c_int32)]


class AxbConstCacheStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("entries", "bytes", "hits", "misses", "invalidations")]


MM_CONST_A, MM_CONST_B = 1, 2


class AxbError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"axb status {status}: {msg}")
        self.status = status


def declared_symbols() -> List[str]:
    """Every function name include/axb.h declares (used by the CPU-side export test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(axb_[a-z0-9_]+)\s*\(", text)))


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C axiom_b200/csrc). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32, f64 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64,
                                   ctypes.c_float, ctypes.c_double)
    sz = ctypes.c_size_t
    T = ctypes.POINTER(AxbTensor)
    I = ctypes.POINTER(AxbInstr)
    pi32, pi64, pvp = ctypes.POINTER(i32), ctypes.POINTER(i64), ctypes.POINTER(vp)
    sig = {
        "axb_abi_version": (i32, []),
        "axb_last_error": (ctypes.c_char_p, []),
        "axb_device_count": (i32, [pi32]),
        "axb_set_device": (i32, [i32]),
        "axb_get_device": (i32, [pi32]),
        "axb_device_name": (i32, [i32, ctypes.c_char_p, sz]),
        "axb_device_sm_count": (i32, [i32, pi32]),
        "axb_default_stream": (i32, [pvp]),
        "axb_stream_sync": (i32, [vp]),
        "axb_launch_count": (u64, []),
        "axb_event_create": (i32, [pvp]),
        "axb_event_destroy": (i32, [vp]),
        "axb_event_record": (i32, [vp, vp]),
        "axb_event_sync": (i32, [vp]),
        "axb_event_elapsed_ms": (i32, [vp, vp, ctypes.POINTER(f32)]),
        "axb_alloc": (i32, [sz, vp, pvp]),
        "axb_free": (i32, [vp, vp]),
        "axb_empty_cache": (i32, []),
        "axb_get_alloc_stats": (i32, [ctypes.POINTER(AxbAllocStats)]),
        "axb_get_transfer_stats": (i32, [ctypes.POINTER(AxbTransferStats)]),
        "axb_pool_reserve": (i32, [sz]),
        "axb_host_alloc": (i32, [sz, pvp]),
        "axb_host_free": (i32, [vp]),
        "axb_host_alloc_numa_node": (i32, []),
        "axb_memcpy_h2d": (i32, [vp, vp, sz, vp]),
        "axb_memcpy_d2h": (i32, [vp, vp, sz, vp, i32]),
        "axb_memcpy_d2d": (i32, [vp, vp, sz, vp]),
        "axb_memset": (i32, [vp, i32, sz, vp]),
        "axb_elementwise": (i32, [I, i32, T, i32, T, vp]),
        "axb_elementwise_precompile": (i32, [I, i32, T, i32, T]),
        "axb_get_jit_stats": (i32, [ctypes.POINTER(AxbJitStats)]),
        "axb_get_half_lut_stats": (i32, [ctypes.POINTER(AxbHalfLutStats)]),
        "axb_get_plan_cache_stats": (i32, [ctypes.POINTER(AxbPlanCacheStats)]),
        "axb_plan_cache_clear": (i32, []),
        "axb_promote_types": (i32, [i32, i32]),
        "axb_binary_result_dtype": (i32, [i32, i32, i32]),
        "axb_unary_result_dtype": (i32, [i32, i32]),
        "axb_dtype_size": (sz, [i32]),
        "axb_broadcast_shape": (i32, [pi64, i32, pi64, i32, pi64, pi32]),
        "axb_unary": (i32, [i32, T, T, f64, vp]),
        "axb_binary": (i32, [i32, T, T, T, vp]),
        "axb_where": (i32, [T, T, T, T, vp]),
        "axb_masked_fill": (i32, [T, T, f64, T, vp]),
        "axb_cast": (i32, [T, T, vp]),
        "axb_gather_strided": (i32, [T, T, vp]),
        "axb_fill": (i32, [T, f64, vp]),
        "axb_randn": (i32, [T, u64, f32, f32, vp]),
        "axb_reduce_result_dtype": (i32, [i32, i32]),
        "axb_reduce": (i32, [i32, T, pi32, i32, i32, T, vp]),
        "axb_argreduce": (i32, [i32, T, i32, i32, T, vp]),
        "axb_softmax": (i32, [T, i32, i32, T, vp]),
        "axb_layer_norm": (i32, [T, T, T, i32, f32, T, vp]),
        "axb_rms_norm": (i32, [T, T, i32, f32, T, vp]),
        "axb_rownorm_last_path": (i32, []),
        "axb_elementwise_reduce": (i32, [I, i32, T, i32, i32, T, vp]),
        "axb_elementwise_softmax": (i32, [I, i32, T, i32, i32, i32, T, vp]),
        "axb_elementwise_norm": (i32, [I, i32, T, i32, i32, T, T, f32, T, vp]),
        "axb_elementwise_reduce_precompile": (i32, [I, i32, T, i32, T, i32, i32]),
        "axb_elementwise_reduce_axis": (i32, [I, i32, T, i32, i32, i32, i32, T, vp]),
        "axb_matmul": (i32, [T, T, i32, i32, ctypes.POINTER(AxbMatmulEpilogue), T, vp]),
        "axb_matmul_shape": (i32, [T, T, i32, i32, pi64, pi32]),
        "axb_matmul_last_path": (i32, []),
        "axb_invalidate_cached": (i32, [vp, sz]),
        "axb_get_const_cache_stats": (i32, [ctypes.POINTER(AxbConstCacheStats)]),
        "axb_sdpa": (i32, [T, T, T, T, f32, i32, T, vp]),
        "axb_sdpa_last_path": (i32, []),
        "axb_gather": (i32, [T, i32, T, T, vp]),
        "axb_index_select": (i32, [T, i32, T, T, vp]),
        "axb_conv2d": (i32, [T, T, T, pi32, pi32, pi32, i32, T, vp]),
        "axb_conv": (i32, [T, T, T, i32, pi32, pi32, pi32, i32, i32, T, vp]),
        "axb_conv_transpose": (i32, [T, T, T, i32, pi32, pi32, pi32, i32, T, vp]),
        "axb_interpolate": (i32, [T, i32, i32, T, vp]),
        "axb_upload_file": (i32, [ctypes.c_char_p, u64, sz, vp, vp]),
        "axb_pool": (i32, [T, i32, i32, pi32, pi32, pi32, i32, T, vp]),
        "axb_sort": (i32, [T, i32, i32, T, T, vp]),
        "axb_masked_select": (i32, [T, T, pvp, pi64, vp]),
        "axb_peer_mailbox_create": (i32, [pvp, ctypes.c_char_p]),
        "axb_peer_mailbox_open": (i32, [ctypes.c_char_p, pvp]),
        "axb_peer_mailbox_close": (i32, [vp]),
        "axb_peer_mailbox_destroy": (i32, [vp]),
        "axb_peer_group_create": (i32, [i32, i32, pvp, pvp]),
        "axb_peer_group_destroy": (i32, [vp]),
        "axb_peer_group_status": (i32, [vp]),
        "axb_reduce_sharded": (i32, [i32, T, i64, vp, T, vp]),
        "axb_elementwise_host": (i32, [I, i32, T, i32, T]),
        "axb_reduce_host": (i32, [i32, T, pi32, i32, i32, T]),
        "axb_matmul_host": (i32, [T, T, i32, i32, T]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    L._axb_missing = missing  # exported-symbol test reports these
    _lib = L
    return L


def check(status: int) -> None:
    if status != OK:
        raise AxbError(status, lib().axb_last_error().decode(errors="replace"))


def make_tensor(base: int, dtype: int, shape: Sequence[int], strides_bytes: Optional[Sequence[int]] = None,
                byte_offset: int = 0) -> AxbTensor:
    t = AxbTensor()
    t.base = base
    t.byte_offset = byte_offset
    t.ndim = len(shape)
    t.dtype = dtype
    if len(shape) > AXB_MAX_DIMS:
        raise ValueError("too many dims")
    if strides_bytes is None:
        st, acc = [], DTYPE_SIZE[dtype]
        for e in reversed(shape):
            st.append(acc)
            acc *= max(int(e), 1)
        strides_bytes = list(reversed(st))
    for i, (e, s) in enumerate(zip(shape, strides_bytes)):
        t.shape[i] = int(e)
        t.strides[i] = int(s)
    return t


def make_instr(op: int, dtype: int, src0: int = SRC_NONE, src1: int = SRC_NONE, src2: int = SRC_NONE,
               imm: float = 0.0) -> AxbInstr:
    ins = AxbInstr()
    ins.op = op
    ins.dtype = dtype
    ins.src[0], ins.src[1], ins.src[2] = src0, src1, src2
    ins.imm = imm
    return ins
