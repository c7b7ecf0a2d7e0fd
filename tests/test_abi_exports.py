"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/axb.h declares; the pure host functions (type rules) agree with the reference's rules."""
import ctypes
import itertools

import numpy as np
import pytest

from axiom_b200 import capi
from oracle import axiom_oracle as ao


def test_library_loads_and_exports_every_declared_symbol():
    L = capi.lib()
    declared = capi.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in include/axb.h but not exported by libaxb.so: {missing}"
    assert L.axb_abi_version() == 2


def test_ctypes_signatures_cover_header():
    L = capi.lib()
    assert L._axb_missing == []


def test_promote_types_matches_reference_rule():
    # src/tensor/operations.cpp:346-397 restated in the oracle and pinned against _ref in test_oracle_pin
    L = capi.lib()
    for a, b in itertools.product(range(13), repeat=2):
        assert L.axb_promote_types(a, b) == ao.promote_types(a, b), (a, b)


def test_result_dtype_rules():
    L = capi.lib()
    for op in ("eq", "ne", "lt", "le", "gt", "ge", "logical_and", "logical_or", "logical_xor"):
        assert L.axb_binary_result_dtype(capi.OP[op], capi.F32, capi.I32) == capi.BOOL
    assert L.axb_binary_result_dtype(capi.OP["add"], capi.I32, capi.F32) == capi.F32
    assert L.axb_binary_result_dtype(capi.OP["add"], capi.U8, capi.I8) == capi.I16
    for op in ("logical_not", "isnan", "isinf", "isfinite"):
        assert L.axb_unary_result_dtype(capi.OP[op], capi.F32) == capi.BOOL
    assert L.axb_unary_result_dtype(capi.OP["exp"], capi.BF16) == capi.BF16


def test_broadcast_shape_host_function():
    L = capi.lib()
    a = (ctypes.c_int64 * 2)(4096, 1)
    b = (ctypes.c_int64 * 2)(1, 16384)
    out = (ctypes.c_int64 * 8)()
    n = ctypes.c_int32()
    assert L.axb_broadcast_shape(a, 2, b, 2, out, ctypes.byref(n)) == capi.OK
    assert [out[i] for i in range(n.value)] == [4096, 16384]
    bad = (ctypes.c_int64 * 2)(3, 16384)
    assert L.axb_broadcast_shape(a, 2, bad, 2, out, ctypes.byref(n)) == capi.ERR_SHAPE
    assert b"broadcast" in L.axb_last_error()


def test_dtype_sizes():
    L = capi.lib()
    for dt, sz in capi.DTYPE_SIZE.items():
        assert L.axb_dtype_size(dt) == sz


def test_no_cpu_fallback_without_device():
    """Without a CUDA device a compute call must fail loudly (status != OK), never compute on the host."""
    L = capi.lib()
    n = ctypes.c_int32(0)
    L.axb_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present; this check is for the CPU-only container")
    host = np.zeros(16, np.float32)
    t = capi.make_tensor(host.ctypes.data, capi.F32, host.shape)
    st = L.axb_unary(capi.OP["relu"], ctypes.byref(t), ctypes.byref(t), 0.0, None)
    assert st != capi.OK
    assert L.axb_last_error()
