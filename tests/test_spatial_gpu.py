"""conv2d / conv_transpose / interpolate through the C ABI (csrc/spatial.cu) against the loop restatements in
oracle/axiom_oracle.py (src/tensor/operations.cpp:3141-3975). fp32 is bit-exact: the kernels keep the CPU loops'
accumulation type and term order. The reference's own known-answer tests pin the restatements (CPU part); the 230-case
CPU-vs-GPU sweep through the reference API is tests/test_axiom_cpp.py::test_spatial_ops_match_cpu_backend."""
import numpy as np
import pytest

from axiom_b200 import capi, device as dev
from oracle import axiom_oracle as ao
from tests.helpers import as_bf16


# ---- CPU: the restatements against the reference's known answers --------------------------------------------------------
def test_oracle_conv_transpose_known_answers():
    # tests/test_conv_transpose.cpp: Conv1dBasic — input [1,2,3], weight [1,1], stride 1 -> [1,3,5,3]
    x = np.array([1, 2, 3], np.float32).reshape(1, 1, 1, 3)
    w = np.array([1, 1], np.float32).reshape(1, 1, 1, 2)
    assert ao.conv_transpose2d_f32(x, w).reshape(-1).tolist() == [1, 3, 5, 3]
    # stride 2 upsamples: (3-1)*2 + 2 = 6 outputs, each input lands on two positions
    assert ao.conv_transpose2d_f32(x, w, stride=(1, 2)).reshape(-1).tolist() == [1, 1, 2, 2, 3, 3]


def test_oracle_interpolate_known_answers():
    # tests/test_interpolate.cpp: nearest 2x repeats every pixel; bilinear align_corners keeps the corners
    x = np.arange(4, dtype=np.float32).reshape(1, 1, 2, 2)
    up = ao.interpolate2d_f32(x, (4, 4), "nearest")
    assert up[0, 0].tolist() == [[0, 0, 1, 1], [0, 0, 1, 1], [2, 2, 3, 3], [2, 2, 3, 3]]
    bl = ao.interpolate2d_f32(x, (3, 3), "bilinear", True)
    assert bl[0, 0, 0, 0] == 0 and bl[0, 0, 2, 2] == 3 and bl[0, 0, 1, 1] == 1.5
    same = ao.interpolate2d_f32(x, (2, 2), "bicubic", True)
    assert np.array_equal(same, x)


def test_oracle_conv2d_known_answers():
    # tests/test_conv2d.cpp: a 3x3 box filter of ones over ones counts the taps; padding 1 keeps the size
    x = np.ones((1, 1, 4, 4), np.float32)
    w = np.ones((1, 1, 3, 3), np.float32)
    assert ao.conv2d_f32(x, w)[0, 0].tolist() == [[9, 9], [9, 9]]
    p = ao.conv2d_f32(x, w, padding=(1, 1))[0, 0]
    assert p.shape == (4, 4) and p[0, 0] == 4 and p[0, 1] == 6 and p[1, 1] == 9


# ---- GPU parity ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1),
                                 dict(stride=(2, 1), padding=(1, 2), dilation=(1, 2), groups=2),
                                 dict(stride=(1, 3), padding=(2, 0), dilation=(2, 1), groups=4)])
def test_conv2d_bit_exact(gpu, cfg):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 4, 9, 10)).astype(np.float32)
    w = (rng.standard_normal((8, 4 // cfg["groups"], 3, 2)) * 0.3).astype(np.float32)
    b = rng.standard_normal(8).astype(np.float32)
    got = dev.conv2d(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b), **cfg).numpy()
    assert np.array_equal(got, ao.conv2d_f32(x, w, b, **cfg))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(stride=(1, 1), padding=(0, 0), output_padding=(0, 0), dilation=(1, 1), groups=1),
                                 dict(stride=(2, 3), padding=(1, 0), output_padding=(1, 2), dilation=(1, 1), groups=2),
                                 dict(stride=(2, 2), padding=(1, 1), output_padding=(0, 1), dilation=(2, 1), groups=1)])
def test_conv_transpose2d_bit_exact(gpu, cfg):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, 4, 5, 6)).astype(np.float32)
    w = (rng.standard_normal((4, 3, 3, 2)) * 0.5).astype(np.float32)
    b = rng.standard_normal(3 * cfg["groups"]).astype(np.float32)
    got = dev.conv_transpose(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b), **cfg).numpy()
    assert np.array_equal(got, ao.conv_transpose2d_f32(x, w, b, **cfg))


@pytest.mark.gpu
def test_conv_transpose1d_matches_the_2d_restatement(gpu):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 4, 11)).astype(np.float32)
    w = rng.standard_normal((4, 2, 5)).astype(np.float32)
    got = dev.conv_transpose(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), None, stride=3, padding=2,
                             output_padding=1, dilation=2, groups=2).numpy()
    exp = ao.conv_transpose2d_f32(x[:, :, None, :], w[:, :, None, :], None, (1, 3), (0, 2), (0, 1), (1, 2), 2)[:, :, 0, :]
    assert np.array_equal(got, exp)


def _conv_transpose_f64(x, w, bias, stride, padding, output_padding, dilation):
    """fp64 truth of ops::conv_transpose2d (operations.cpp:3590-3690), one einsum per filter tap; also returns the same sum over
    absolute values (the scale a rounding-error bound is stated against)."""
    n, cin, h, wd = x.shape
    _, cout, kh, kw = w.shape
    oh = (h - 1) * stride[0] - 2 * padding[0] + dilation[0] * (kh - 1) + output_padding[0] + 1
    ow = (wd - 1) * stride[1] - 2 * padding[1] + dilation[1] * (kw - 1) + output_padding[1] + 1
    ph, pw = max(padding[0], 0) + dilation[0] * kh, max(padding[1], 0) + dilation[1] * kw  # margin: scatter first, crop after
    full = np.zeros((2, n, cout, oh + 2 * ph + stride[0], ow + 2 * pw + stride[1]))
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    for fh in range(kh):
        for fw in range(kw):
            t = np.einsum("nchw,co->nohw", x64, w64[:, :, fh, fw])
            ta = np.einsum("nchw,co->nohw", np.abs(x64), np.abs(w64[:, :, fh, fw]))
            r0, c0 = ph - padding[0] + fh * dilation[0], pw - padding[1] + fw * dilation[1]
            full[0][:, :, r0:r0 + (h - 1) * stride[0] + 1:stride[0], c0:c0 + (wd - 1) * stride[1] + 1:stride[1]] += t
            full[1][:, :, r0:r0 + (h - 1) * stride[0] + 1:stride[0], c0:c0 + (wd - 1) * stride[1] + 1:stride[1]] += ta
    out = full[:, :, :, ph:ph + oh, pw:pw + ow]
    if bias is not None:
        out[0] += bias.astype(np.float64)[None, :, None, None]
        out[1] += np.abs(bias.astype(np.float64))[None, :, None, None]
    return out[0], out[1]


def test_conv_transpose_f64_helper_matches_the_loop_restatement():
    rng = np.random.default_rng(40)
    x = rng.standard_normal((2, 3, 4, 5)).astype(np.float32)
    w = rng.standard_normal((3, 2, 3, 2)).astype(np.float32)
    b = rng.standard_normal(2).astype(np.float32)
    for cfg in (dict(stride=(2, 3), padding=(1, 0), output_padding=(1, 2), dilation=(1, 1)),
                dict(stride=(2, 2), padding=(1, 1), output_padding=(0, 1), dilation=(2, 1))):
        t, _ = _conv_transpose_f64(x, w, b, **cfg)
        assert np.allclose(t, ao.conv_transpose2d_f32(x, w, b, groups=1, **cfg), rtol=1e-5, atol=1e-5)


def _conv_transpose_by_phases(x, w, stride, padding, output_padding, dilation):
    """The decomposition spatial.cu conv_transpose_tc hands to the tensor cores, restated with numpy: for every output phase
    (oh % s_h, ow % s_w) the taps with (r + p - k * d) % s == 0 form an arithmetic progression k0 + i * step, tap i reads input row
    oh' + q0 + i * dil with q0 = (r + p - k0 * d) / s and dil = -(step * d / s); rows outside the input count as zero. Returns the
    fp64 result, or None when some phase has no tap (the case the library leaves to the gather kernel)."""
    n, cin, h, wd = x.shape
    _, cout, kh, kw = w.shape
    oh = (h - 1) * stride[0] - 2 * padding[0] + dilation[0] * (kh - 1) + output_padding[0] + 1
    ow = (wd - 1) * stride[1] - 2 * padding[1] + dilation[1] * (kw - 1) + output_padding[1] + 1

    def plan(r, s, p, d, k):
        ks = [kk for kk in range(k) if (r + p - kk * d) % s == 0]
        if not ks:
            return None
        step = ks[1] - ks[0] if len(ks) > 1 else 1
        assert all(b - a == step for a, b in zip(ks, ks[1:])), "taps of a phase are equally spaced"
        q0 = (r + p - ks[0] * d) // s
        dil = (r + p - ks[1] * d) // s - q0 if len(ks) > 1 else 0
        assert len(ks) == 1 or dil * s == -step * d
        return ks[0], step, len(ks), q0, dil

    out = np.zeros((n, cout, oh, ow))
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    for rh in range(stride[0]):
        for rw in range(stride[1]):
            ph, pw = plan(rh, stride[0], padding[0], dilation[0], kh), plan(rw, stride[1], padding[1], dilation[1], kw)
            if ph is None or pw is None:
                return None
            ohr, owr = len(range(rh, oh, stride[0])), len(range(rw, ow, stride[1]))
            acc = np.zeros((n, cout, ohr, owr))
            for i in range(ph[2]):
                for j in range(pw[2]):
                    tap = w64[:, :, ph[0] + i * ph[1], pw[0] + j * pw[1]]           # (C_in, C_out): the strided weight view
                    r0, c0 = ph[3] + i * ph[4], pw[3] + j * pw[4]                     # input row / column of output (0, 0)
                    rows, cols = np.arange(ohr) + r0, np.arange(owr) + c0
                    rv, cv = (rows >= 0) & (rows < h), (cols >= 0) & (cols < wd)      # out-of-range = TMA zero fill
                    patch = np.zeros((n, cin, ohr, owr))
                    patch[:, :, np.ix_(rv, cv)[0], np.ix_(rv, cv)[1]] = x64[:, :, rows[rv]][:, :, :, cols[cv]]
                    acc += np.einsum("nchw,co->nohw", patch, tap)
            out[:, :, rh::stride[0], rw::stride[1]] = acc
    return out


@pytest.mark.parametrize("cfg", [dict(stride=(2, 2), padding=(1, 1), output_padding=(0, 0), dilation=(1, 1)),
                                 dict(stride=(2, 3), padding=(1, 0), output_padding=(1, 2), dilation=(1, 1)),
                                 dict(stride=(3, 2), padding=(2, 0), output_padding=(2, 1), dilation=(2, 1)),
                                 dict(stride=(1, 4), padding=(0, 2), output_padding=(0, 3), dilation=(1, 3)),
                                 dict(stride=(2, 2), padding=(0, 3), output_padding=(1, 0), dilation=(2, 2))])
def test_conv_transpose_phase_decomposition_is_the_transposed_convolution(cfg):
    """CPU pin of the algorithm behind the tensor-core path: phase by phase it is the scatter loop of ops::conv_transpose2d
    (operations.cpp:3590-3690), including dilation against stride, output_padding and negative effective padding."""
    rng = np.random.default_rng(60)
    x = rng.standard_normal((2, 3, 5, 7)).astype(np.float32)
    w = rng.standard_normal((3, 4, 5, 4)).astype(np.float32)
    got = _conv_transpose_by_phases(x, w, **cfg)
    if cfg["dilation"] == (2, 2):  # dilation 2 under stride 2: odd phases have no tap at all
        assert got is None
        return
    truth, _ = _conv_transpose_f64(x, w, None, **cfg)
    assert got is not None and np.allclose(got, truth, rtol=1e-12, atol=1e-12)
    assert np.allclose(truth, ao.conv_transpose2d_f32(x, w, None, groups=1, **cfg), rtol=1e-5, atol=1e-5)


CONVT_TC = [  # (x shape, w shape, kwargs)
    ((8, 128, 16, 16), (128, 64, 4, 4), dict(stride=(2, 2), padding=(1, 1), output_padding=(0, 0), dilation=(1, 1))),   # 16-wide phases: four rows per chunk
    ((2, 64, 32, 32), (64, 64, 4, 4), dict(stride=(2, 2), padding=(1, 1), output_padding=(0, 0), dilation=(1, 1))),   # 2x upsampling
    ((2, 96, 30, 33), (96, 72, 3, 3), dict(stride=(2, 2), padding=(1, 1), output_padding=(1, 1), dilation=(1, 1))),   # 1 / 2 / 2 / 4 taps per phase
    ((2, 40, 25, 40), (40, 96, 5, 5), dict(stride=(3, 2), padding=(2, 0), output_padding=(2, 1), dilation=(2, 1))),   # dilation 2 against stride 3
    ((3, 64, 20, 64), (64, 64, 3, 5), dict(stride=(1, 1), padding=(1, 2), output_padding=(0, 0), dilation=(1, 1))),   # stride 1: one phase, vector stores
    ((3, 64, 24, 40), (64, 64, 4, 3), dict(stride=(2, 1), padding=(1, 1), output_padding=(0, 0), dilation=(1, 1))),   # strided rows, dense columns
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CONVT_TC)))
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_conv_transpose_on_the_tensor_cores(gpu, case, dt):
    """Transposed convolutions large enough for the tensor cores run as one implicit-GEMM convolution per output phase
    (spatial.cu conv_transpose_tc): path asserted, result against the fp64 truth — fp32 to 2e-6 of sum |x||w| (the exact
    three-piece split), bf16 to half a bf16 ulp of the result plus the fp32 accumulation bound."""
    xs, ws, kw = CONVT_TC[case]
    rng = np.random.default_rng(41 + case)
    x = rng.standard_normal(xs).astype(np.float32)
    w = (rng.standard_normal(ws) * 0.2).astype(np.float32)
    b = rng.standard_normal(ws[1]).astype(np.float32)
    if dt == "bf16":
        xb, wb, bb = as_bf16(x), as_bf16(w), as_bf16(b)
        x, w, b = ao.bf16_to_f32(xb), ao.bf16_to_f32(wb), ao.bf16_to_f32(bb)
        got = ao.bf16_to_f32(dev.conv_transpose(dev.DeviceArray.from_numpy(xb, bf16=True), dev.DeviceArray.from_numpy(wb, bf16=True),
                                                dev.DeviceArray.from_numpy(bb, bf16=True), **kw).numpy())
    else:
        got = dev.conv_transpose(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b), **kw).numpy()
    assert capi.lib().axb_matmul_last_path() == 4, "expected the implicit-GEMM path"
    truth, scale = _conv_transpose_f64(x, w, b, **kw)
    assert got.shape == truth.shape
    err = np.abs(got.astype(np.float64) - truth)
    if dt == "f32":
        assert np.all(err <= 2e-6 * scale + 1e-7), float((err / (scale + 1e-30)).max())
    else:  # conv result rounded to bf16, + bias rounded to bf16 (two roundings, the reference's op order)
        # half a bf16 ulp is up to 2^-8 |v|: once on the convolution (|conv| <= |truth| + |b|), once on the sum
        assert np.all(err <= 2.0 ** -8 * (2.0 * np.abs(truth) + np.abs(b).max()) + 1e-5 * scale), float(err.max())


@pytest.mark.gpu
def test_conv_transpose1d_on_the_tensor_cores(gpu):
    rng = np.random.default_rng(49)
    x = rng.standard_normal((4, 128, 500)).astype(np.float32)
    w = (rng.standard_normal((128, 96, 8)) * 0.1).astype(np.float32)
    got = dev.conv_transpose(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), None, stride=4, padding=2).numpy()
    assert capi.lib().axb_matmul_last_path() == 4
    truth, scale = _conv_transpose_f64(x[:, :, None, :], w[:, :, None, :], None, (1, 4), (0, 2), (0, 0), (1, 1))
    assert np.all(np.abs(got - truth[:, :, 0, :]) <= 2e-6 * scale[:, :, 0, :] + 1e-7)


@pytest.mark.gpu
def test_grouped_conv_transpose_on_the_tensor_cores(gpu):
    """groups = 2: each group is the phase-wise tensor-core path on channel-slice views (weight rows g * C_in/groups ..)."""
    rng = np.random.default_rng(51)
    groups, cin_g, cout_g = 2, 48, 40
    x = rng.standard_normal((6, cin_g * groups, 24, 28)).astype(np.float32)
    w = (rng.standard_normal((cin_g * groups, cout_g, 4, 4)) * 0.1).astype(np.float32)
    b = rng.standard_normal(cout_g * groups).astype(np.float32)
    kw = dict(stride=(2, 2), padding=(1, 1), output_padding=(0, 0), dilation=(1, 1))
    got = dev.conv_transpose(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b), groups=groups, **kw).numpy()
    assert capi.lib().axb_matmul_last_path() == 4
    for g in range(groups):
        truth, scale = _conv_transpose_f64(x[:, g * cin_g:(g + 1) * cin_g], w[g * cin_g:(g + 1) * cin_g], b[g * cout_g:(g + 1) * cout_g], **kw)
        err = np.abs(got[:, g * cout_g:(g + 1) * cout_g].astype(np.float64) - truth)
        assert np.all(err <= 2e-6 * scale + 1e-7), (g, float((err / (scale + 1e-30)).max()))


@pytest.mark.gpu
def test_conv_transpose_kernel_smaller_than_stride_stays_on_the_gather_kernel(gpu):
    """A 2 x 2 filter under stride 3 leaves output phases without any tap (they hold the bias only): not a convolution the
    implicit GEMM can express, the gather kernel runs (and stays bit-identical to the loop restatement on a sample)."""
    rng = np.random.default_rng(50)
    x = rng.standard_normal((2, 64, 40, 40)).astype(np.float32)
    w = (rng.standard_normal((64, 64, 2, 2)) * 0.2).astype(np.float32)
    b = rng.standard_normal(64).astype(np.float32)
    got = dev.conv_transpose(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b), stride=3).numpy()
    assert capi.lib().axb_matmul_last_path() == 0
    truth, scale = _conv_transpose_f64(x, w, b, (3, 3), (0, 0), (0, 0), (1, 1))
    assert np.allclose(got, truth, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nearest", "bilinear", "bicubic"])
@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("size", [(8, 8), (3, 11), (7, 5)])
def test_interpolate_bit_exact(gpu, mode, align, size):
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2, 3, 5, 7)).astype(np.float32)
    got = dev.interpolate(dev.DeviceArray.from_numpy(x), size, mode, align).numpy()
    assert np.array_equal(got, ao.interpolate2d_f32(x, size, mode, align)), (mode, align, size)


@pytest.mark.gpu
def test_spatial_errors_are_reported(gpu):
    x = dev.DeviceArray.from_numpy(np.zeros((1, 3, 4, 4), np.float32))
    w = dev.DeviceArray.from_numpy(np.zeros((2, 2, 3, 3), np.float32))  # C_in / groups mismatch
    with pytest.raises(Exception):
        dev.conv2d(x, w)
    with pytest.raises(Exception):
        dev.interpolate(dev.DeviceArray.from_numpy(np.zeros((4, 4), np.float32).reshape(1, 1, 1, 4, 4)), (2, 2, 2))


# ---- convolution as an implicit GEMM on the tensor cores (gemm_tcgen05.cu conv_tc) ----------------------------------------
def test_oracle_fast_conv_matches_the_loops():
    rng = np.random.default_rng(41)
    x = rng.standard_normal((2, 3, 7, 9)).astype(np.float32)
    w = rng.standard_normal((4, 3, 3, 2)).astype(np.float32)
    b = rng.standard_normal(4).astype(np.float32)
    for cfg in (dict(stride=(1, 1), padding=(0, 0), dilation=(1, 1)), dict(stride=(2, 1), padding=(1, 2), dilation=(1, 2))):
        fast, mag = ao.conv2d_f64(x, w, b, **cfg)
        assert np.array_equal(fast.astype(np.float32), ao.conv2d_f32(x, w, b, **cfg)) and np.all(mag >= np.abs(fast) - 1e-12)


def _check_conv_tc(x, w, b, dt, cfg, activation=None, expect_path=4):
    from axiom_b200 import capi
    xs, ws = (x, w) if x.ndim == 4 else (x[:, :, None, :], w[:, :, None, :])
    c2 = {k: ((1, v) if k != "padding" else (0, v)) if x.ndim == 3 else v for k, v in cfg.items()}
    if dt == "f32":
        dx, dw, db = dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b) if b is not None else None
        xe, we, be = xs, ws, b
        tol = 4e-6  # the three-piece bf16 split is as accurate as an fp32 FMA chain: error <= a few 2^-24 of sum |x||w|
    else:
        cv = ao.f32_to_bf16 if dt == "bf16" else (lambda a: a.astype(np.float16))
        back = ao.bf16_to_f32 if dt == "bf16" else (lambda a: a.astype(np.float32))
        xh, wh, bh = cv(x), cv(w), cv(b) if b is not None else None
        dx, dw = dev.DeviceArray.from_numpy(xh, bf16=dt == "bf16"), dev.DeviceArray.from_numpy(wh, bf16=dt == "bf16")
        db = dev.DeviceArray.from_numpy(bh, bf16=dt == "bf16") if b is not None else None
        xe, we = (back(xh), back(wh)) if x.ndim == 4 else (back(xh)[:, :, None, :], back(wh)[:, :, None, :])
        be = back(bh) if b is not None else None
        tol = 1e-2 if dt == "bf16" else 2e-3  # north_star: fp32 accumulate, one rounding, <= 1e-2 relative
    got = dev.conv(dx, dw, db, activation=activation, **cfg)
    assert capi.lib().axb_matmul_last_path() == expect_path, "unexpected kernel path"
    got = got.numpy()
    gotf = got if dt == "f32" else (ao.bf16_to_f32(got) if dt == "bf16" else got.astype(np.float32))
    exp, mag = ao.conv2d_f64(xe, we, be, **c2)
    if x.ndim == 3:
        exp, mag = exp[:, :, 0, :], mag[:, :, 0, :]
    if activation == "relu":
        exp = np.maximum(exp, 0)
    elif activation == "gelu":
        from scipy.special import erf
        exp = 0.5 * exp * (1 + erf(exp / np.sqrt(2)))
    assert gotf.shape == exp.shape
    err = np.abs(gotf.astype(np.float64) - exp)
    bound = tol * mag + (1e-6 if activation == "gelu" else 0)
    assert np.all(err <= bound), (float((err / np.maximum(mag, 1e-30)).max()), dt, cfg)


CONV_TC_2D = [((2, 64, 56, 56), (128, 64, 3, 3), dict(stride=(1, 1), padding=(1, 1), dilation=(1, 1))),
              ((2, 64, 57, 61), (96, 64, 3, 3), dict(stride=(2, 2), padding=(1, 1), dilation=(1, 1))),
              ((1, 48, 40, 70), (130, 48, 5, 3), dict(stride=(1, 2), padding=(2, 0), dilation=(1, 1))),
              ((2, 80, 33, 64), (64, 80, 3, 3), dict(stride=(1, 1), padding=(2, 2), dilation=(2, 2))),
              ((3, 200, 28, 40), (256, 200, 1, 1), dict(stride=(1, 1), padding=(0, 0), dilation=(1, 1))),
              ((1, 32, 64, 64), (64, 32, 7, 7), dict(stride=(2, 2), padding=(3, 3), dilation=(1, 1))),
              ((2, 64, 60, 96), (64, 64, 3, 3), dict(stride=(3, 2), padding=(0, 1), dilation=(2, 1))),
              ((1, 64, 21, 200), (72, 64, 3, 3), dict(stride=(1, 1), padding=(1, 1), dilation=(1, 1))),   # wider than a tile: one row per tile
              ((8, 64, 3, 64), (64, 64, 3, 3), dict(stride=(1, 1), padding=(1, 1), dilation=(1, 1))),     # OH = 3: two-row groups, ragged
              ((4, 128, 14, 14), (256, 128, 3, 3), dict(stride=(1, 1), padding=(1, 1), dilation=(1, 1))), # 14 x 14 feature map: four rows per chunk
              ((5, 64, 29, 16), (96, 64, 3, 3), dict(stride=(1, 1), padding=(1, 1), dilation=(1, 1))),   # OW = 16 (vector stores), OH = 29 ragged groups
              ((6, 64, 26, 28), (64, 64, 5, 3), dict(stride=(1, 1), padding=(2, 1), dilation=(1, 1))),   # OW = 28: two rows per chunk
              ((3, 96, 40, 23), (160, 96, 3, 3), dict(stride=(2, 2), padding=(1, 1), dilation=(1, 1)))]  # strided, OW = 12 (odd row pitch)


@pytest.mark.gpu
@pytest.mark.parametrize("xs,ws,cfg", CONV_TC_2D)
@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
def test_conv2d_implicit_gemm(gpu, xs, ws, cfg, dt):
    rng = np.random.default_rng(43)
    x = rng.standard_normal(xs).astype(np.float32)
    w = (rng.standard_normal(ws) / np.sqrt(ws[1] * ws[2] * ws[3])).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    _check_conv_tc(x, w, b, dt, cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("groups,odd", [(2, False), (4, True)])
def test_grouped_conv2d_runs_one_implicit_gemm_per_group(gpu, dt, groups, odd):
    """groups > 1 with enough channels per group: every group is the tensor-core convolution of a channel slice (views of
    the same tensors, result written into its channel slice; an odd OH * OW leaves the slice unaligned for the 128-bit stores).
    Truth: the fp64 convolution of each slice. + bias + relu in the epilogue."""
    rng = np.random.default_rng(47)
    cin_g, cout_g = 32, 48
    hw = (27, 33) if odd else (28, 32)
    x = rng.standard_normal((6, cin_g * groups, *hw)).astype(np.float32)
    w = (rng.standard_normal((cout_g * groups, cin_g, 3, 3)) / np.sqrt(cin_g * 9)).astype(np.float32)
    b = rng.standard_normal(cout_g * groups).astype(np.float32)
    cfg = dict(stride=(1, 1), padding=(1, 1), dilation=(1, 1))
    if dt == "bf16":
        xb, wb, bb = as_bf16(x), as_bf16(w), as_bf16(b)
        x, w, b = ao.bf16_to_f32(xb), ao.bf16_to_f32(wb), ao.bf16_to_f32(bb)
        got = ao.bf16_to_f32(dev.conv(dev.DeviceArray.from_numpy(xb, bf16=True), dev.DeviceArray.from_numpy(wb, bf16=True),
                                      dev.DeviceArray.from_numpy(bb, bf16=True), groups=groups, activation="relu", **cfg).numpy())
    else:
        got = dev.conv(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b), groups=groups,
                       activation="relu", **cfg).numpy()
    assert capi.lib().axb_matmul_last_path() == 4
    for g in range(groups):
        exp, mag = ao.conv2d_f64(x[:, g * cin_g:(g + 1) * cin_g], w[g * cout_g:(g + 1) * cout_g], b[g * cout_g:(g + 1) * cout_g], **cfg)
        err = np.abs(got[:, g * cout_g:(g + 1) * cout_g].astype(np.float64) - np.maximum(exp, 0))
        assert np.all(err <= (4e-6 if dt == "f32" else 1e-2) * mag), (g, float((err / mag).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_conv1d_implicit_gemm_whisper_shapes(gpu, dt):
    """The two convolutions of the whisper encoder (examples/whisper_poc.cpp:57-60): 80 -> 384, k 3, pad 1 and 384 -> 384,
    k 3, stride 2, pad 1, with the bias and a fused GELU in the GEMM epilogue."""
    rng = np.random.default_rng(44)
    x = rng.standard_normal((2, 80, 3000)).astype(np.float32)
    w = (rng.standard_normal((384, 80, 3)) / 15).astype(np.float32)
    b = rng.standard_normal(384).astype(np.float32)
    _check_conv_tc(x, w, b, dt, dict(stride=1, padding=1, dilation=1))
    _check_conv_tc(x, w, b, dt, dict(stride=1, padding=1, dilation=1), activation="gelu")
    x2 = rng.standard_normal((2, 384, 3000)).astype(np.float32)
    w2 = (rng.standard_normal((384, 384, 3)) / 34).astype(np.float32)
    _check_conv_tc(x2, w2, b, dt, dict(stride=2, padding=1, dilation=1), activation="gelu")
    _check_conv_tc(x2, w2, None, dt, dict(stride=2, padding=0, dilation=3))


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["bf16", "f16"])
def test_conv_direct_nchw_path_equals_the_prepared_path(gpu, dt, monkeypatch):
    """Half-precision stride-1 convolutions with one-column filters (1x1, KHx1) read the NCHW input in place (MN-major TMA
    boxes); AXB_CONV_NO_DIRECT forces the channels-last preparation pass. Both are the same sum in the same order: the results must be IDENTICAL bit for bit."""
    import subprocess, sys, os, json
    code = r"""
import sys, json, numpy as np
sys.path.insert(0, %r)
from axiom_b200 import capi, device as dev
from oracle import axiom_oracle as ao
capi.check(capi.lib().axb_set_device(0))
rng = np.random.default_rng(51)
outs = []
for xs, ws, cfg in (((4, 256, 56, 56), (128, 256, 1, 1), dict(stride=(1, 1), padding=(0, 0), dilation=(1, 1))),
                    ((1, 72, 24, 200), (96, 72, 3, 1), dict(stride=(1, 1), padding=(2, 0), dilation=(2, 1))),
                    ((3, 200, 28, 40), (256, 200, 1, 1), dict(stride=(1, 1), padding=(0, 0), dilation=(1, 1)))):
    x = rng.standard_normal(xs).astype(np.float32); w = (rng.standard_normal(ws) / 20).astype(np.float32)
    cv = ao.f32_to_bf16 if %r == "bf16" else (lambda a: a.astype(np.float16))
    dx = dev.DeviceArray.from_numpy(cv(x), bf16=%r == "bf16"); dw = dev.DeviceArray.from_numpy(cv(w), bf16=%r == "bf16")
    o = dev.conv(dx, dw, None, **cfg); assert capi.lib().axb_matmul_last_path() == 4
    outs.append(o.numpy().view(np.uint16).tobytes().hex()[:0] + str(hash(o.numpy().view(np.uint16).tobytes())))
print(json.dumps(outs))
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), dt, dt, dt)
    env = dict(os.environ, PYTHONHASHSEED="0")
    a = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    b = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(env, AXB_CONV_NO_DIRECT="1"), timeout=600)
    assert a.returncode == 0 and b.returncode == 0, a.stderr[-800:] + b.stderr[-800:]
    assert json.loads(a.stdout.strip().splitlines()[-1]) == json.loads(b.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_conv_implicit_gemm_views_and_small_fallback(gpu):
    from axiom_b200 import capi
    rng = np.random.default_rng(45)
    # input as a strided view (channels-last storage seen as NCHW) and a sliced weight: the preparation pass reads strides
    xl = rng.standard_normal((2, 40, 50, 64)).astype(np.float32)          # N H W C
    x = xl.transpose(0, 3, 1, 2)                                          # N C H W view
    wbig = (rng.standard_normal((96, 64, 3, 5)) / 30).astype(np.float32)
    w = wbig[:, :, :, ::2]                                                # (96, 64, 3, 3) view
    got = dev.conv(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(w), None, stride=(1, 1), padding=(1, 1), dilation=(1, 1))
    assert capi.lib().axb_matmul_last_path() == 4
    exp, mag = ao.conv2d_f64(np.ascontiguousarray(x), np.ascontiguousarray(w), None, (1, 1), (1, 1), (1, 1))
    assert np.all(np.abs(got.numpy() - exp) <= 4e-6 * mag)
    # a small problem stays on the gather kernel and is bit-identical to the CPU loops
    xs = rng.standard_normal((1, 3, 9, 9)).astype(np.float32)
    wsm = rng.standard_normal((4, 3, 3, 3)).astype(np.float32)
    got = dev.conv(dev.DeviceArray.from_numpy(xs), dev.DeviceArray.from_numpy(wsm), None, stride=(1, 1), padding=(1, 1), dilation=(1, 1))
    assert capi.lib().axb_matmul_last_path() == 0
    assert np.array_equal(got.numpy(), ao.conv2d_f32(xs, wsm, None, (1, 1), (1, 1), (1, 1)))
    # off the tensor-core path the activation is a second, in-place launch: conv (bit-exact) followed by the exact relu
    got = dev.conv(dev.DeviceArray.from_numpy(xs), dev.DeviceArray.from_numpy(wsm), None, stride=(1, 1), padding=(1, 1), dilation=(1, 1),
                   activation="relu")
    assert np.array_equal(got.numpy(), np.maximum(ao.conv2d_f32(xs, wsm, None, (1, 1), (1, 1), (1, 1)), 0))
