"""Host <-> device transfers of the storage seam (csrc/runtime.cpp): pageable buffers of 16 MiB and more go through a pool of
pinned staging buffers filled / drained by several host threads (Tensor::gpu() / .cpu() of large tensors,
src/tensor/tensor.cpp:1593-1650); file ranges go pread() -> pinned -> device (axb_upload_file, safetensors loading)."""
import ctypes
import os
import tempfile
import time

import numpy as np
import pytest

from axiom_b200 import capi, device as dev


@pytest.mark.gpu
@pytest.mark.parametrize("nbytes", [16 << 20, (16 << 20) + 4, (100 << 20) + 12345 * 4, 8 << 20])
def test_pageable_round_trip_is_exact(gpu, nbytes):
    rng = np.random.default_rng(nbytes & 0xFFFF)
    x = rng.integers(0, 2 ** 31 - 1, nbytes // 4, dtype=np.int32)
    d = dev.DeviceArray.from_numpy(x)       # pageable H2D (staged above 16 MiB)
    back = d.numpy()                        # synchronous D2H into a pageable array
    assert np.array_equal(back, x)
    # a kernel in between: the staged copies are ordered on the backend stream like the plain ones
    y = dev.binary("add", d, dev.DeviceArray.from_numpy(np.array([1], np.int32))).numpy()
    assert np.array_equal(y, x + 1)


@pytest.mark.gpu
def test_staged_copies_are_counted_and_faster_than_the_driver_path(gpu):
    L = capi.lib()
    n = 256 << 20
    x = np.ones(n // 4, np.float32)
    st0, st1 = capi.AxbTransferStats(), capi.AxbTransferStats()
    capi.check(L.axb_get_transfer_stats(ctypes.byref(st0)))
    d = dev.DeviceArray.from_numpy(x)
    capi.check(L.axb_get_transfer_stats(ctypes.byref(st1)))
    assert st1.h2d_bytes - st0.h2d_bytes == n
    t0 = time.perf_counter()
    for _ in range(3):
        d = dev.DeviceArray.from_numpy(x)
    up = 3 * n / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter()
    for _ in range(3):
        d.numpy()
    down = 3 * n / (time.perf_counter() - t0) / 1e9
    print(f"pageable H2D {up:.1f} GB/s, D2H {down:.1f} GB/s (includes numpy allocation / page faults on the way down)")
    assert up > 8.0  # the single-threaded driver path measures ~10-12 GB/s on this host; several lanes must not be slower


@pytest.mark.gpu
def test_upload_file_ranges(gpu):
    L = capi.lib()
    rng = np.random.default_rng(3)
    blob = rng.integers(0, 255, (40 << 20) + 777, dtype=np.uint8)
    with tempfile.NamedTemporaryFile(delete=False) as f:
        f.write(blob.tobytes())
        path = f.name
    try:
        for off, n in ((0, blob.size), (12345, 33 << 20), (blob.size - 1000, 1000), (5, 0)):
            d = dev.DeviceArray.empty((max(n, 1),), capi.U8)
            capi.check(L.axb_upload_file(path.encode(), off, n, ctypes.c_void_p(d.storage.ptr), None))
            if n:
                assert np.array_equal(d.numpy()[:n], blob[off:off + n])
        with pytest.raises(capi.AxbError):  # reading past the end of the file is an error, not a short upload
            d = dev.DeviceArray.empty((2000,), capi.U8)
            capi.check(L.axb_upload_file(path.encode(), blob.size - 1000, 2000, ctypes.c_void_p(d.storage.ptr), None))
        with pytest.raises(capi.AxbError):
            capi.check(L.axb_upload_file(b"/nonexistent/file", 0, 16, ctypes.c_void_p(d.storage.ptr), None))
    finally:
        os.unlink(path)
