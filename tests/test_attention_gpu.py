"""axb_sdpa (csrc/attention.cu) against an fp64 restatement of the reference's attention
(src/backends/cpu/flash_attention_cpu.cpp:73-131: S = Q K^T * scale, causal masks kv > q, explicit mask non-zero = masked
out, softmax over kv, O = P V) and — where the unmodified reference library is present — against the reference itself.
Mirrors tests/test_flash_attention.cpp (shapes, causal, masks) and adds the whisper shape {B, 6, 1500, 64}."""
import ctypes

import numpy as np
import pytest

from axiom_b200 import capi, device as dev
from oracle import ref
from tests.helpers import rand_f32

pytestmark = pytest.mark.gpu


def _sdpa(q, k, v, mask=None, scale=-1.0, causal=False):
    L = capi.lib()
    L.axb_sdpa.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int,
                           ctypes.c_void_p, ctypes.c_void_p]
    L.axb_sdpa.restype = ctypes.c_int
    dq, dk, dv = (dev.DeviceArray.from_numpy(t) for t in (q, k, v))
    out = dev.DeviceArray.empty(q.shape[:3] + (v.shape[3],), capi.F32)
    dm = dev.DeviceArray.from_numpy(mask) if mask is not None else None
    qd, kd, vd, od = dq.desc(), dk.desc(), dv.desc(), out.desc()
    md = dm.desc() if dm is not None else None
    capi.check(L.axb_sdpa(ctypes.byref(qd), ctypes.byref(kd), ctypes.byref(vd), ctypes.byref(md) if md is not None else None,
                          scale, int(causal), ctypes.byref(od), None))
    return out.numpy()


def _truth(q, k, v, mask=None, scale=-1.0, causal=False):
    q, k, v = (t.astype(np.float64) for t in (q, k, v))
    if scale <= 0:
        scale = 1.0 / np.sqrt(q.shape[-1])
    s = np.einsum("bhqd,bhkd->bhqk", q, k) * scale
    lq, lk = s.shape[-2:]
    if causal:
        s = np.where(np.arange(lk)[None, :] > np.arange(lq)[:, None], -np.inf, s)
    if mask is not None:
        s = np.where(np.broadcast_to(mask != 0, s.shape), -np.inf, s)
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(axis=-1, keepdims=True)
    return np.einsum("bhqk,bhkd->bhqd", p, v)


@pytest.mark.parametrize("shape", [(1, 1, 4, 8, 4), (2, 3, 17, 16, 33), (1, 2, 128, 64, 128), (2, 6, 300, 64, 300),
                                   (1, 4, 65, 32, 2100), (1, 1, 40, 64, 9000)])
@pytest.mark.parametrize("causal", [False, True])
def test_sdpa_matches_fp64_attention(gpu, shape, causal):
    b, h, lq, d, lk = shape
    rng = np.random.default_rng(5 + lq)
    q, k, v = rand_f32(rng, (b, h, lq, d)), rand_f32(rng, (b, h, lk, d)), rand_f32(rng, (b, h, lk, d))
    got = _sdpa(q, k, v, causal=causal)
    exp = _truth(q, k, v, causal=causal)
    assert got.shape == exp.shape
    # fp32 scores of magnitude ~|q||k| feed an exponential: the reference's own CPU/GPU tolerance for attention is 1e-4
    # (tests/test_flash_attention.cpp); measured error here is ~1e-5 absolute on O(1) outputs
    assert np.allclose(got, exp, rtol=1e-4, atol=2e-5), float(np.max(np.abs(got - exp)))


def test_sdpa_masks_scale_and_strided_operands(gpu):
    rng = np.random.default_rng(11)
    b, h, lq, lk, d = 2, 4, 96, 160, 64
    # (B, L, H, D) storage viewed as (B, H, L, D): what nn::MultiHeadAttention hands over (src/nn/attention.cpp:42-52)
    q = rand_f32(rng, (b, lq, h, d)).transpose(0, 2, 1, 3)
    k = rand_f32(rng, (b, lk, h, d)).transpose(0, 2, 1, 3)
    v = rand_f32(rng, (b, lk, h, d)).transpose(0, 2, 1, 3)
    m2 = rng.random((lq, lk)) < 0.3
    m2[:, 0] = False                                  # keep one visible key per row
    m4 = rng.random((b, 1, lq, lk)) < 0.5
    m4[..., 1] = False
    for mask in (m2, m4):
        got = _sdpa(q, k, v, mask=mask.astype(np.bool_), scale=0.2)
        assert np.allclose(got, _truth(q, k, v, mask=mask, scale=0.2), rtol=1e-4, atol=2e-5)
    if ref.available() and hasattr(ref.RefTensor, "sdpa"):
        r = ref.RefTensor.from_numpy(np.ascontiguousarray(q)).sdpa(ref.RefTensor.from_numpy(np.ascontiguousarray(k)),
                                                                   ref.RefTensor.from_numpy(np.ascontiguousarray(v))).numpy()
        assert np.allclose(_sdpa(q, k, v), r, rtol=1e-4, atol=1e-5)


def test_sdpa_whisper_shape_presplit_path(gpu):
    """{2, 6, 1500, 64}: the P V product takes the pre-split tensor-core path; the same call with the path disabled
    (probabilities written as fp32) must agree to fp32 rounding, and both with the fp64 result."""
    import os
    rng = np.random.default_rng(13)
    q, k, v = (rand_f32(rng, (2, 6, 1500, 64)) for _ in range(3))
    n0 = dev.launch_count()
    got = _sdpa(q, k, v)
    launches = dev.launch_count() - n0
    assert launches <= 6, launches          # 2 splits + QK^T, softmax, 1 split (V) + PV
    exp = _truth(q, k, v)
    assert np.allclose(got, exp, rtol=1e-4, atol=2e-5), float(np.max(np.abs(got - exp)))
    os.environ["AXB_SDPA_NO_PRESPLIT"] = "1"
    try:
        plain = _sdpa(q, k, v)
    finally:
        del os.environ["AXB_SDPA_NO_PRESPLIT"]
    assert np.allclose(got, plain, rtol=1e-5, atol=2e-6)
