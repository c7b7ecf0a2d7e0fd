"""axb_sdpa (csrc/flash_attention.cu: one tcgen05 kernel for head_dim <= 64; csrc/attention.cu: wider heads) against an fp64 restatement of the reference's attention
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


def _sdpa(q, k, v, mask=None, scale=-1.0, causal=False, half=None, expect_path=None):
    """half: None (fp32 arrays), "bf16" or "f16" (q, k, v are uint16 bit patterns / float16 arrays)."""
    L = capi.lib()
    L.axb_sdpa.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int,
                           ctypes.c_void_p, ctypes.c_void_p]
    L.axb_sdpa.restype = ctypes.c_int
    dq, dk, dv = (dev.DeviceArray.from_numpy(t, bf16=(half == "bf16")) for t in (q, k, v))
    out = dev.DeviceArray.empty(q.shape[:3] + (v.shape[3],), {None: capi.F32, "bf16": capi.BF16, "f16": capi.F16}[half])
    dm = dev.DeviceArray.from_numpy(mask) if mask is not None else None
    qd, kd, vd, od = dq.desc(), dk.desc(), dv.desc(), out.desc()
    md = dm.desc() if dm is not None else None
    capi.check(L.axb_sdpa(ctypes.byref(qd), ctypes.byref(kd), ctypes.byref(vd), ctypes.byref(md) if md is not None else None,
                          scale, int(causal), ctypes.byref(od), None))
    if expect_path is not None:
        assert L.axb_sdpa_last_path() == expect_path
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
    got = _sdpa(q, k, v, causal=causal, expect_path=1)  # head_dim <= 64: the single flash kernel
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


def test_sdpa_whisper_shape_one_kernel(gpu):
    """{2, 6, 1500, 64}: operand prologue + ONE flash kernel; the three-kernel path (score matrix materialised, the round-1
    implementation) must agree to fp32 rounding, and both with the fp64 result."""
    import os
    rng = np.random.default_rng(13)
    q, k, v = (rand_f32(rng, (2, 6, 1500, 64)) for _ in range(3))
    n0 = dev.launch_count()
    got = _sdpa(q, k, v, expect_path=1)
    launches = dev.launch_count() - n0
    assert launches == 2, launches          # prologue (split of Q, K, V in one launch) + the attention kernel
    exp = _truth(q, k, v)
    assert np.allclose(got, exp, rtol=1e-4, atol=2e-5), float(np.max(np.abs(got - exp)))
    os.environ["AXB_SDPA_NO_FLASH"] = "1"
    try:
        plain = _sdpa(q, k, v, expect_path=0)
    finally:
        del os.environ["AXB_SDPA_NO_FLASH"]
    assert np.allclose(got, plain, rtol=1e-4, atol=2e-5), float(np.max(np.abs(got - plain)))


def test_sdpa_wide_heads_keep_the_three_kernel_path(gpu):
    rng = np.random.default_rng(17)
    q, k, v = (rand_f32(rng, (1, 2, 200, 128), 1.0) for _ in range(3))
    got = _sdpa(q, k, v, expect_path=0)
    assert np.allclose(got, _truth(q, k, v), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("half", ["bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 3, 130, 64, 257), (1, 2, 31, 32, 70), (1, 6, 1500, 64, 1500)])
def test_sdpa_half_types_native(gpu, half, shape):
    """bf16 / fp16 operands run the same kernel with one piece per operand (P rounded to the half type before P V, fp32
    accumulation, one rounding of O): <= 1e-2 of the fp64 attention of the same half inputs (the reference's own half-type
    tolerance, tests/test_flash_attention.cpp:248)."""
    from oracle import axiom_oracle as ao
    b, h, lq, d, lk = shape
    rng = np.random.default_rng(23 + lq)
    f = [rand_f32(rng, s, 1.0) for s in ((b, h, lq, d), (b, h, lk, d), (b, h, lk, d))]
    if half == "bf16":
        bits = [ao.f32_to_bf16(t) for t in f]
        vals = [ao.bf16_to_f32(t) for t in bits]
        got = ao.bf16_to_f32(_sdpa(*bits, causal=True, half="bf16", expect_path=1))
    else:
        bits = [t.astype(np.float16) for t in f]
        vals = [t.astype(np.float32) for t in bits]
        got = _sdpa(*bits, causal=True, half="f16", expect_path=1).astype(np.float32)
    exp = _truth(*vals, causal=True)
    assert np.allclose(got, exp, rtol=1e-2, atol=1e-2), float(np.max(np.abs(got - exp)))


def test_sdpa_fully_masked_rows_are_nan_and_late_maxima_rescale(gpu):
    """(1) a query row whose keys are all masked is NaN in the reference (exp(-inf - -inf), flash_attention_cpu.cpp:118-126);
    (2) a row whose maximum grows by far more than 2^8 in a LATE key block exercises the lazy accumulator rescale."""
    rng = np.random.default_rng(29)
    b, h, lq, lk, d = 1, 2, 140, 330, 64
    q, k, v = rand_f32(rng, (b, h, lq, d), 0.5), rand_f32(rng, (b, h, lk, d), 0.5), rand_f32(rng, (b, h, lk, d))
    k[:, :, 300, :] = q[:, :, 7:8, :].mean(axis=2) * 40.0   # a huge score for row 7 in the fifth key block
    k[:, :, 200, :] *= 25.0
    mask = np.zeros((lq, lk), np.bool_)
    mask[3, :] = True                                        # row 3: everything masked
    mask[9, :64] = True                                      # row 9: the whole first key block masked
    got = _sdpa(q, k, v, mask=mask, expect_path=1)
    exp = _truth(q, k, v, mask=mask)
    assert np.isnan(got[:, :, 3, :]).all()
    keep = np.ones(lq, bool)
    keep[3] = False
    # row 9 starts with a fully masked block: NaN in the reference's recurrence as well (alpha = exp(-inf - -inf))
    keep[9] = False
    assert np.isnan(got[:, :, 9, :]).all()
    assert np.allclose(got[:, :, keep, :], exp[:, :, keep, :], rtol=1e-4, atol=2e-5), float(np.nanmax(np.abs(got[:, :, keep] - exp[:, :, keep])))
