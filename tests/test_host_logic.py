"""CPU-side checks of the host logic behind the C ABI (no GPU, no compute): kernel specialisation (program lowering +
NVRTC -> in-tree cubin cache), matmul shape rules (infer_matmul_shape, src/graph/graph_registry.cpp:147-220) and
argument validation."""
import ctypes
import os

import numpy as np
import pytest

from axiom_b200 import capi

BASE = 0x7F0000000000  # any 16-byte aligned non-null address: descriptors only


def _t(dtype, shape, strides=None):
    return capi.make_tensor(BASE, dtype, shape, strides)


def _shape(a, b, ta=False, tb=False):
    L = capi.lib()
    out = (ctypes.c_int64 * 8)()
    n = ctypes.c_int32()
    st = L.axb_matmul_shape(ctypes.byref(a), ctypes.byref(b), int(ta), int(tb), out, ctypes.byref(n))
    return st, [out[i] for i in range(n.value)]


def test_matmul_shape_rules():
    F = capi.F32
    assert _shape(_t(F, (2, 3)), _t(F, (3, 4))) == (capi.OK, [2, 4])
    assert _shape(_t(F, (4,)), _t(F, (4,))) == (capi.OK, [])                    # 1D x 1D -> rank-0
    assert _shape(_t(F, (3,)), _t(F, (3, 4))) == (capi.OK, [4])
    assert _shape(_t(F, (2, 3)), _t(F, (3,))) == (capi.OK, [2])
    assert _shape(_t(F, (2, 3, 4)), _t(F, (2, 4, 5))) == (capi.OK, [2, 3, 5])
    assert _shape(_t(F, (2, 1, 3, 4)), _t(F, (4, 5))) == (capi.OK, [2, 1, 3, 5])  # broadcast batch
    assert _shape(_t(F, (3, 1, 7, 4)), _t(F, (1, 5, 4, 2))) == (capi.OK, [3, 5, 7, 2])
    assert _shape(_t(F, (3, 2)), _t(F, (3, 4)), ta=True) == (capi.OK, [2, 4])
    assert _shape(_t(F, (2, 3)), _t(F, (4, 3)), tb=True) == (capi.OK, [2, 4])
    st, _ = _shape(_t(F, (2, 3)), _t(F, (4, 5)))
    assert st == capi.ERR_SHAPE and b"inner dimensions" in capi.lib().axb_last_error()
    st, _ = _shape(_t(F, (2, 2, 3)), _t(F, (3, 3, 4)))
    assert st == capi.ERR_SHAPE


def _chain():
    F, B, V = capi.F32, capi.BOOL, capi.SRC_VALUE
    prog = [capi.make_instr(capi.OP["relu"], F, 0), capi.make_instr(capi.OP["add"], F, V(0), 1),
            capi.make_instr(capi.OP["mul"], F, V(1), 2), capi.make_instr(capi.OP["gt"], B, 3, 4),
            capi.make_instr(capi.OP["where"], F, V(3), V(2), 5)]
    return (capi.AxbInstr * len(prog))(*prog), len(prog)


def test_specialised_kernel_compiles_without_a_gpu_and_is_cached(tmp_path, monkeypatch):
    """Signature = ops + dtypes + access modes; shapes and immediates must NOT create new kernels."""
    L = capi.lib()
    F = capi.F32
    P, n = _chain()

    def precompile(rows, cols):
        ins = [_t(F, (rows, 1)), _t(F, (1,)), _t(F, (1, cols)), _t(F, (rows, cols)), _t(capi.I32, (1,)), _t(F, (1,))]
        I = (capi.AxbTensor * len(ins))(*ins)
        out = _t(F, (rows, cols))
        return L.axb_elementwise_precompile(P, n, I, len(ins), ctypes.byref(out))

    s0 = capi.AxbJitStats()
    L.axb_get_jit_stats(ctypes.byref(s0))
    assert precompile(4096, 16384) == capi.OK, L.axb_last_error()
    assert precompile(64, 256) == capi.OK       # same signature, different shape
    s1 = capi.AxbJitStats()
    L.axb_get_jit_stats(ctypes.byref(s1))
    assert (s1.compiles - s0.compiles) + (s1.disk_hits - s0.disk_hits) == 2
    assert s1.compiles - s0.compiles <= 1, "a shape change must not trigger a second compile"
    cache = os.path.join(os.path.dirname(capi.LIB_PATH), "kernel_cache")
    assert any(f.endswith(".axbk") for f in os.listdir(cache))


def test_program_validation_errors():
    L = capi.lib()
    F = capi.F32
    bad = (capi.AxbInstr * 1)(capi.make_instr(capi.OP["add"], F, 0, 5))  # references input 5 of 2
    ins = (capi.AxbTensor * 2)(_t(F, (4,)), _t(F, (4,)))
    out = _t(F, (4,))
    assert L.axb_elementwise_precompile(bad, 1, ins, 2, ctypes.byref(out)) == capi.ERR_INVALID
    cmp_not_bool = (capi.AxbInstr * 1)(capi.make_instr(capi.OP["gt"], F, 0, 1))
    assert L.axb_elementwise_precompile(cmp_not_bool, 1, ins, 2, ctypes.byref(out)) == capi.ERR_INVALID
    bitand_float = (capi.AxbInstr * 1)(capi.make_instr(capi.OP["bit_and"], F, 0, 1))
    assert L.axb_elementwise_precompile(bitand_float, 1, ins, 2, ctypes.byref(out)) == capi.ERR_UNSUPPORTED
    ok = (capi.AxbInstr * 1)(capi.make_instr(capi.OP["add"], F, 0, 1))
    wrong_shape = _t(F, (5,))
    assert L.axb_elementwise_precompile(ok, 1, ins, 2, ctypes.byref(wrong_shape)) == capi.ERR_SHAPE
    cplx = (capi.AxbTensor * 2)(_t(capi.C64, (4,)), _t(capi.C64, (4,)))
    assert L.axb_elementwise_precompile(ok, 1, cplx, 2, ctypes.byref(_t(capi.C64, (4,)))) == capi.ERR_UNSUPPORTED


def test_reduce_result_dtype_rules():
    L = capi.lib()
    assert L.axb_reduce_result_dtype(capi.OP["mean"], capi.BF16) == capi.F32   # graph_executor.cpp:396-403
    assert L.axb_reduce_result_dtype(capi.OP["sum"], capi.BF16) == capi.BF16
    assert L.axb_reduce_result_dtype(capi.OP["argmax"], capi.F32) == capi.I64
    assert L.axb_reduce_result_dtype(capi.OP["max"], capi.I32) == capi.I32
