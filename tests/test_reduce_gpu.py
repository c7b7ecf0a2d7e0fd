"""GPU parity for kernel family 2 (axis reductions, argmax/argmin, softmax, layer/RMS norm) through the C ABI.

Mirrors the reference's tests/test_tensor_reductions.cpp (known answers on arange(6).reshape({2,3}) :14-136, the
non-contiguous multi-axis GPU case :138-165), tests/test_cpu_gpu_parity.cpp reductions/argmax/softmax (:288-370,
:406-420), tests/test_tensor_argmax_argmin.cpp (:35-148) and tests/test_tensor_normalization.cpp (:9-130), extended
with what the reference never runs on a GPU: wide rows, odd sizes, split (multi-CTA) reductions, strided / offset /
permuted inputs, bf16/fp16/int dtypes, ties and NaNs for argmax, and the BASELINE config-3 shape {8192,32768}.

Bars (north_star): argmax/argmin bit-exact; fp32 reductions <= 1e-5 relative — measured, as SURVEY.md §7 explains,
against sum(|x|) (a tree and the reference's 16-lane serial order both round at 1e-7 * sum|x|, and randn rows cancel);
softmax / norms <= 1e-5 relative+absolute vs the oracle; max/min exact.
"""
import numpy as np
import pytest

from axiom_b200 import capi, device as dev
from oracle import axiom_oracle as ao
from oracle import ref
from tests.helpers import as_bf16, bits_equal, rand_f32

pytestmark = pytest.mark.gpu

REF_RED = dict(sum="Sum", mean="Mean", max="Max", min="Min", prod="Prod", any="Any", all="All")


def _sum_close(got, exp, x, axes, rtol=1e-5):
    """|got - exp| <= rtol * sum(|x|) over the reduced axes (+ tiny absolute slack)."""
    mag = np.abs(x.astype(np.float64)).sum(axis=tuple(axes) if axes else None)
    mag = np.asarray(mag).reshape(np.asarray(exp).shape)
    err = np.abs(np.asarray(got, np.float64) - np.asarray(exp, np.float64))
    assert np.all(err <= rtol * mag + 1e-30), f"max err {err.max()} vs bound {(rtol * mag).min()}"


# ---- known answers (tests/test_tensor_reductions.cpp:14-63) ----------------------------------------------------
def test_known_answers_arange6(gpu):
    x = np.arange(6, dtype=np.float32).reshape(2, 3)
    d = dev.DeviceArray.from_numpy(x)
    assert dev.reduce("sum", d).numpy().tolist() == [15.0]
    assert dev.reduce("sum", d, [0]).numpy().tolist() == [3.0, 5.0, 7.0]
    assert dev.reduce("sum", d, [1]).numpy().tolist() == [3.0, 12.0]
    assert dev.reduce("mean", d).numpy().tolist() == [2.5]
    assert dev.reduce("mean", d, [1]).numpy().tolist() == [1.0, 4.0]
    assert dev.reduce("max", d).numpy().tolist() == [5.0]
    assert dev.reduce("min", d, [0]).numpy().tolist() == [0.0, 1.0, 2.0]
    assert dev.reduce("sum", d, [1], keep_dims=True).numpy().tolist() == [[3.0], [12.0]]
    assert dev.reduce("sum", d, [], keep_dims=True).numpy().tolist() == [[15.0]]


def test_noncontiguous_sum_gpu(gpu):
    """tests/test_tensor_reductions.cpp:138-165 — sliced (non-zero offset) + permuted view, axes {1,2}."""
    base = np.arange(2 * 5 * 4 * 3, dtype=np.float32).reshape(2, 5, 4, 3)
    v = base[:, 1:4].transpose(0, 2, 1, 3)  # offset + permute
    got = dev.reduce("sum", dev.DeviceArray.from_numpy(v), [1, 2]).numpy()
    assert np.array_equal(got, v.sum(axis=(1, 2)))


SHAPES_AXES = [((4, 5), [0]), ((4, 5), [1]), ((4, 5), []), ((4, 5, 3), [1]), ((4, 5, 3), [0, 2]), ((2, 4, 5), [2]),
               ((1,), []), ((31,), [0]), ((33, 7), [0]), ((1023,), []), ((257, 129), [1]), ((257, 129), [0]),
               ((7, 1, 33), [0, 1]), ((3, 4000), [1]), ((5, 70001), [1]), ((2, 3, 5, 7, 4), [1, 3]), ((300000,), []),
               ((64, 96, 32), [1])]


@pytest.mark.parametrize("shape,axes", SHAPES_AXES)
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min"])
def test_reduce_f32_vs_oracle_and_reference(gpu, op, shape, axes):
    rng = np.random.default_rng(abs(hash((op, shape, tuple(axes)))) % 2**31)
    x = rand_f32(rng, shape)
    for keep in (False, True):
        got = dev.reduce(op, dev.DeviceArray.from_numpy(x), axes, keep).numpy()
        exp, dt = ao.reduce(op, x, ao.F32, axes, keep)
        assert got.shape == np.asarray(exp).shape
        if op in ("max", "min"):
            assert bits_equal(got, np.asarray(exp, np.float32))
        else:
            xs = x / (np.prod([shape[a] for a in (axes or range(len(shape)))]) if op == "mean" else 1)
            _sum_close(got, exp, xs, axes or list(range(len(shape))))
    if ref.available() and all(a >= 0 for a in axes):
        r = ref.RefTensor.from_numpy(x).reduce(REF_RED[op], axes, False).numpy()
        got = dev.reduce(op, dev.DeviceArray.from_numpy(x), axes, False).numpy()
        if op in ("max", "min"):
            assert bits_equal(got, r)
        else:
            xs = x / (np.prod([shape[a] for a in (axes or range(len(shape)))]) if op == "mean" else 1)
            _sum_close(got, r, xs, axes or list(range(len(shape))))


@pytest.mark.parametrize("dtname", ["bf16", "f16", "f64", "i32", "i64", "i16", "u8"])
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min", "prod"])
def test_reduce_other_dtypes(gpu, op, dtname):
    rng = np.random.default_rng(5)
    shape, axes = (6, 130), [1]
    if dtname in ("bf16", "f16", "f64"):
        xf = rand_f32(rng, shape, 1.0)
        if op == "prod":
            xf = (1.0 + 0.01 * xf).astype(np.float32)
        if dtname == "bf16":
            x, dt, bf = as_bf16(xf), ao.BF16, True
        elif dtname == "f16":
            x, dt, bf = xf.astype(np.float16), ao.F16, False
        else:
            x, dt, bf = xf.astype(np.float64), ao.F64, False
    else:
        npdt = dict(i32=np.int32, i64=np.int64, i16=np.int16, u8=np.uint8)[dtname]
        lo, hi = (0, 4) if op == "prod" else (0, 50) if dtname == "u8" else (-50, 50)
        x = rng.integers(lo, hi, size=shape).astype(npdt)
        if op == "prod":
            x = np.where(rng.random(shape) < 0.9, 1, x).astype(npdt)
        dt, bf = dict(i32=ao.I32, i64=ao.I64, i16=ao.I16, u8=ao.U8)[dtname], False
    got = dev.reduce(op, dev.DeviceArray.from_numpy(x, bf16=bf), axes).numpy()
    exp, edt = ao.reduce(op, x, dt, axes)
    exp = np.asarray(exp)
    if dtname in ("bf16", "f16"):
        g = ao.to_f32(got, edt) if edt in (ao.BF16, ao.F16) else got
        e = ao.to_f32(exp, edt) if edt in (ao.BF16, ao.F16) else exp
        assert np.allclose(g, e, rtol=2 ** -7, atol=1e-2), (op, dtname)  # one half-ulp of the rounded result
    elif dtname == "f64":
        assert np.allclose(got, exp, rtol=1e-12, atol=1e-12)
    else:
        assert np.array_equal(got, exp.astype(got.dtype)), (op, dtname)


def test_any_all(gpu):
    x = np.zeros((5, 300), np.float32)
    x[1, 7] = 2.0
    x[3, :] = 1.0
    d = dev.DeviceArray.from_numpy(x)
    assert dev.reduce("any", d, [1]).numpy().astype(bool).tolist() == [False, True, False, True, False]
    assert dev.reduce("all", d, [1]).numpy().astype(bool).tolist() == [False, False, False, True, False]
    assert bool(dev.reduce("any", d).numpy()[0]) and not bool(dev.reduce("all", d).numpy()[0])


# ---- argmax / argmin: bit-exact, first index on ties, NaN never wins ---------------------------
def test_argmax_quirks_known_rows(gpu):
    rows = np.array([[1, 5, 5], [np.nan, 2, 2], [np.nan, np.nan, np.nan], [-np.inf, -np.inf, -np.inf]], np.float32)
    got = dev.argreduce("argmax", dev.DeviceArray.from_numpy(rows), 1).numpy()
    assert got.tolist() == [1, 1, 0, 0]
    got = dev.argreduce("argmin", dev.DeviceArray.from_numpy(rows), 1).numpy()
    assert got.tolist() == [0, 1, 0, 0]
    assert got.dtype == np.int64


@pytest.mark.parametrize("shape,axis", [((4, 5), 1), ((4, 5), 0), ((2, 4, 5), 1), ((6,), 0), ((33, 1025), 1),
                                        ((1025, 33), 0), ((3, 70001), 1), ((200000,), -1), ((7, 9), -1)])
@pytest.mark.parametrize("op", ["argmax", "argmin"])
def test_argreduce_bit_exact(gpu, op, shape, axis):
    rng = np.random.default_rng(3)
    x = np.round(rand_f32(rng, shape) * 4) / 4  # plenty of exact ties
    x.flat[::97] = np.nan
    got = dev.argreduce(op, dev.DeviceArray.from_numpy(x), axis).numpy()
    exp = ao.argreduce(op, x, ao.F32, axis)
    assert got.shape == exp.shape and np.array_equal(got, exp), (op, shape, axis)
    if ref.available() and axis >= 0:
        r = ref.RefTensor.from_numpy(x).argreduce("ArgMax" if op == "argmax" else "ArgMin", axis).numpy()
        assert np.array_equal(got, r)
    for dt, conv in ((ao.I32, lambda a: np.nan_to_num(a * 4).astype(np.int32)), (ao.BF16, None)):
        if conv is None:
            xb = as_bf16(np.nan_to_num(x))
            g = dev.argreduce(op, dev.DeviceArray.from_numpy(xb, bf16=True), axis).numpy()
            e = ao.argreduce(op, xb, ao.BF16, axis)
        else:
            xi = conv(x)
            g = dev.argreduce(op, dev.DeviceArray.from_numpy(xi), axis).numpy()
            e = ao.argreduce(op, xi, dt, axis)
        assert np.array_equal(g, e), (op, shape, axis, dt)


def test_argmax_strided_input(gpu):
    rng = np.random.default_rng(9)
    base = rand_f32(rng, (40, 60))
    v = base[3:, ::2].T  # offset, strided, transposed
    got = dev.argreduce("argmax", dev.DeviceArray.from_numpy(v), 0).numpy()
    assert np.array_equal(got, np.argmax(v, axis=0))


# ---- softmax / log_softmax ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,axis", [((4, 5), -1), ((4, 5), 0), ((2, 4, 5), 1), ((3, 257), -1), ((9, 1024), -1),
                                        ((5, 1500), -1), ((3, 8192), -1), ((2, 8200), -1), ((2, 32768), -1),
                                        ((2, 40000), -1), ((1, 7), -1)])
@pytest.mark.parametrize("log", [False, True])
def test_softmax_f32(gpu, shape, axis, log):
    rng = np.random.default_rng(21)
    x = rand_f32(rng, shape, 3.0)
    got = dev.softmax(dev.DeviceArray.from_numpy(x), axis, log).numpy()
    # (1) against the exact (fp64) softmax: <= 1e-5 relative — the north_star bound, measured against the truth
    x64 = x.astype(np.float64)
    z = x64 - x64.max(axis=axis, keepdims=True)
    lse = np.log(np.exp(z).sum(axis=axis, keepdims=True))
    truth = (z - lse) if log else np.exp(z - lse)
    assert np.allclose(got, truth, rtol=1e-5, atol=1e-6 if not log else 1e-5), (shape, axis, log)
    if not log:
        assert np.allclose(got.sum(axis=axis), 1.0, atol=1e-4)
    # (2) against the oracle / the reference itself. Their denominator is a SERIAL fp32 sum (cpu_operations.cpp:3132-3158)
    # whose own rounding error grows like sqrt(n)*2^-24 (2.5e-5 at n = 32768), so beyond n ~ 4096 the comparison uses the
    # reference's own CPU/GPU parity tolerance for softmax (tests/test_cpu_gpu_parity.cpp: 1e-4).
    tol = 1e-5 if x.shape[axis] <= 4096 else 1e-4
    exp = ao.softmax(x, ao.F32, axis, log)
    assert np.allclose(got, exp, rtol=tol, atol=1e-6 if not log else tol), (shape, axis, log)
    if ref.available():
        r = ref.RefTensor.from_numpy(x).softmax(axis, log).numpy()
        assert np.allclose(got, r, rtol=tol, atol=1e-6 if not log else tol)


def test_softmax_half_and_strided(gpu):
    rng = np.random.default_rng(22)
    x = rand_f32(rng, (6, 300), 2.0)
    xb = as_bf16(x)
    got = dev.softmax(dev.DeviceArray.from_numpy(xb, bf16=True), -1).numpy()
    exp = ao.softmax(xb, ao.BF16, -1)
    assert np.max(np.abs(ao.bf16_to_f32(got) - ao.bf16_to_f32(exp))) <= 2 ** -8 * np.max(ao.bf16_to_f32(exp))
    xh = x.astype(np.float16)
    got = dev.softmax(dev.DeviceArray.from_numpy(xh), -1).numpy()
    exp = ao.softmax(xh, ao.F16, -1)
    assert np.allclose(got.astype(np.float32), np.asarray(exp).astype(np.float32), rtol=2e-3, atol=1e-6)
    v = rand_f32(rng, (64, 48))[5:, ::3].T  # strided view, softmax along a non-contiguous axis
    got = dev.softmax(dev.DeviceArray.from_numpy(v), 0).numpy()
    assert np.allclose(got, ao.softmax(np.ascontiguousarray(v), ao.F32, 0), rtol=1e-5, atol=1e-6)


# ---- layer_norm / rms_norm (tests/test_tensor_normalization.cpp:9-130) --------------------------------------------
@pytest.mark.parametrize("shape", [(2, 3, 4), (4, 384), (3, 1000), (2, 1500, 384), (2, 8192), (2, 32768), (1, 33000)])
def test_layer_and_rms_norm_f32(gpu, shape):
    rng = np.random.default_rng(31)
    x = rand_f32(rng, shape, 2.0) + np.float32(0.5)
    n = shape[-1]
    w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
    b = (0.1 * rng.standard_normal(n)).astype(np.float32)
    dx, dw, db = (dev.DeviceArray.from_numpy(t) for t in (x, w, b))
    got = dev.layer_norm(dx, dw, db, -1, 1e-5).numpy()
    exp, _ = ao.layer_norm(x, ao.F32, w, ao.F32, b, ao.F32, -1, 1e-5)
    assert np.allclose(got, exp, rtol=1e-5, atol=2e-5), shape
    got_r = dev.rms_norm(dx, dw, -1, 1e-5).numpy()
    exp_r, _ = ao.rms_norm(x, ao.F32, w, ao.F32, -1, 1e-5)
    assert np.allclose(got_r, exp_r, rtol=1e-5, atol=2e-5), shape
    if ref.available() and np.prod(shape) <= 2_000_000:
        rx, rw, rb = (ref.RefTensor.from_numpy(t) for t in (x, w, b))
        assert np.allclose(got, rx.layer_norm(rw, rb, -1, 1e-5).numpy(), rtol=1e-5, atol=2e-5)
        assert np.allclose(got_r, rx.rms_norm(rw, -1, 1e-5).numpy(), rtol=1e-5, atol=2e-5)


def test_layer_norm_known_answer_and_no_affine(gpu):
    x = np.array([[1, 2, 3, 4]], np.float32)
    got = dev.layer_norm(dev.DeviceArray.from_numpy(x), None, None, -1, 0.0).numpy()
    e = (x - 2.5) / np.sqrt(1.25)
    assert np.allclose(got, e, rtol=1e-6)


def test_layer_norm_bf16_matches_cpu_float32_result(gpu):
    """the CPU path returns Float32 for bf16 input (mean rounded to bf16 first); the GPU lazy node
    declares the input dtype. Both output dtypes are offered; the Float32 one is compared with the oracle."""
    rng = np.random.default_rng(33)
    x = as_bf16(rand_f32(rng, (5, 384), 2.0))
    w = (1 + 0.1 * rng.standard_normal(384)).astype(np.float32)
    b = (0.1 * rng.standard_normal(384)).astype(np.float32)
    dx, dw, db = dev.DeviceArray.from_numpy(x, bf16=True), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b)
    got32 = dev.layer_norm(dx, dw, db, -1, 1e-5, out_dtype=capi.F32).numpy()
    exp, edt = ao.layer_norm(x, ao.BF16, w, ao.F32, b, ao.F32, -1, 1e-5)
    assert edt == ao.F32 and np.allclose(got32, exp, rtol=1e-5, atol=2e-5)
    got16 = dev.layer_norm(dx, dw, db, -1, 1e-5).numpy()
    assert np.array_equal(got16, ao.f32_to_bf16(got32))


# ---- BASELINE config 3: {8192, 32768} rows, fp32 and bf16 ---------------------------------------------------------
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_config3_full_size(gpu, dt):
    rows, n = 8192, 32768
    x = dev.randn((rows, n), 7)
    if dt == "bf16":
        x = dev.cast(x, capi.BF16)
    # a 64-row sample goes to the host oracle; whole-tensor properties cover the rest
    sample = dev.DeviceArray(x.storage, x.dtype, (64, n), x.strides, x.offset + 4000 * x.strides[0])
    hs = sample.numpy()
    odt = ao.BF16 if dt == "bf16" else ao.F32
    hf = ao.to_f32(hs, odt) if dt == "bf16" else hs

    s = dev.reduce("sum", x, [1]).numpy()
    es, _ = ao.reduce("sum", hs, odt, [1])
    if dt == "bf16":
        assert np.max(np.abs(ao.bf16_to_f32(s[4000:4064]) - ao.bf16_to_f32(np.asarray(es)))) <= 2.0  # 1 bf16 ulp at ~256
    else:
        _sum_close(s[4000:4064], es, hf, [1])
    m = dev.reduce("max", x, [1]).numpy()
    em, _ = ao.reduce("max", hs, odt, [1])
    assert bits_equal(m[4000:4064], np.asarray(em).astype(m.dtype))
    am = dev.argreduce("argmax", x, 1).numpy()
    assert np.array_equal(am[4000:4064], ao.argreduce("argmax", hs, odt, 1))
    # property over ALL rows: the element at argmax equals the row max
    idx = dev.DeviceArray.from_numpy(am.reshape(rows, 1))
    picked = dev.gather(x, 1, idx).numpy().reshape(rows)
    assert bits_equal(picked, m)

    sm = dev.softmax(x, -1)
    sms = dev.DeviceArray(sm.storage, sm.dtype, (64, n), sm.strides, sm.offset + 4000 * sm.strides[0]).numpy()
    esm = ao.softmax(hs, odt, -1)
    if dt == "bf16":
        assert np.max(np.abs(ao.bf16_to_f32(sms) - ao.bf16_to_f32(esm))) <= 2 ** -8 * ao.bf16_to_f32(esm).max()
    else:
        assert np.allclose(sms, esm, rtol=1e-4, atol=1e-9)  # serial fp32 denominator in the reference, see test_softmax_f32
        h64 = hf.astype(np.float64)
        z = h64 - h64.max(axis=1, keepdims=True)
        assert np.allclose(sms, np.exp(z) / np.exp(z).sum(axis=1, keepdims=True), rtol=1e-5, atol=1e-12)
        rs = dev.reduce("sum", sm, [1]).numpy()  # every softmax row sums to 1
        assert np.allclose(rs, 1.0, atol=1e-4)

    w = dev.randn((n,), 8, 1.0, 0.1)
    b = dev.randn((n,), 9, 0.0, 0.1)
    ln = dev.layer_norm(x, w, b, -1, 1e-5, out_dtype=capi.F32)
    lns = dev.DeviceArray(ln.storage, ln.dtype, (64, n), ln.strides, ln.offset + 4000 * ln.strides[0]).numpy()
    eln, _ = ao.layer_norm(hs, odt, w.numpy(), ao.F32, b.numpy(), ao.F32, -1, 1e-5)
    assert np.allclose(lns, eln, rtol=1e-5, atol=3e-5)


# ---- wide-row (TMA-staged) and cluster (SM-pair, DSMEM) kernels: shapes that select them, edge cases ----------------
def _f64_softmax(x, log=False):
    x64 = x.astype(np.float64)
    z = x64 - x64.max(axis=-1, keepdims=True)
    lse = np.log(np.exp(z).sum(axis=-1, keepdims=True))
    return (z - lse) if log else np.exp(z - lse)


@pytest.mark.parametrize("shape", [(300, 4096), (37, 5000), (150, 8192), (149, 12288), (3, 32768), (200, 16388), (9, 32764)])
@pytest.mark.parametrize("log", [False, True])
def test_softmax_wide_rows_f32(gpu, shape, log):
    """Rows of 16 KiB..128 KiB take rownorm_wide.cuh (every J template, exact and masked lane maps, 1..4 stages, more
    and fewer rows than SMs, ragged vector counts)."""
    rng = np.random.default_rng(41)
    x = rand_f32(rng, shape, 3.0)
    got = dev.softmax(dev.DeviceArray.from_numpy(x), -1, log).numpy()
    truth = _f64_softmax(x, log)
    assert np.allclose(got, truth, rtol=1e-5, atol=1e-12 if not log else 1e-5), (shape, log)
    if not log:
        assert np.allclose(got.astype(np.float64).sum(axis=-1), 1.0, atol=1e-5)


def test_softmax_wide_rows_masked_inf_nan(gpu):
    """-inf entries (masked attention scores) give exact zeros; a NaN anywhere or an all -inf row gives an all-NaN row
    as in the reference loop (x - max = NaN); other rows are unaffected. bf16 NaNs carry the reference encoding 0x7FC0."""
    rng = np.random.default_rng(42)
    x = rand_f32(rng, (160, 8192), 2.0)
    x[:, 4000:] = -np.inf           # masked tail
    x[7, 13] = np.nan
    x[9, :] = -np.inf
    got = dev.softmax(dev.DeviceArray.from_numpy(x), -1).numpy()
    ok = np.ones(160, bool)
    ok[[7, 9]] = False
    assert np.all(got[ok][:, 4000:] == 0.0) and np.all(np.isnan(got[~ok]))
    assert np.allclose(got[ok], _f64_softmax(x[ok]), rtol=1e-5, atol=1e-12)
    xb = as_bf16(x)
    gb = dev.softmax(dev.DeviceArray.from_numpy(xb, bf16=True), -1).numpy()
    assert np.all(gb[~ok] == 0x7FC0), "NaN rows must use the reference's bf16 NaN encoding"
    assert np.all(gb[ok][:, 4000:] == 0)
    eb = ao.softmax(xb[ok], ao.BF16, -1)
    assert np.max(np.abs(ao.bf16_to_f32(gb[ok]) - ao.bf16_to_f32(eb))) <= 2 ** -8 * ao.bf16_to_f32(eb).max()


def test_log_softmax_wide_rows_masked_inf_nan(gpu):
    """log_softmax on the TMA-staged kernel: masked (-inf) entries stay -inf, a NaN anywhere / an all -inf row / a +inf entry
    make the whole row NaN (x - max = NaN in the reference loop, cpu_operations.cpp:3132-3158), other rows are unaffected.
    bf16 rows go through the hardware converter; rows that produced NaNs are stored again with the reference's 0x7FC0."""
    rng = np.random.default_rng(44)
    for n in (8192, 32768):
        x = rand_f32(rng, (150, n), 2.0)
        x[:, n // 2:] = -np.inf
        x[7, 13] = np.nan
        x[9, :] = -np.inf
        x[11, 5] = np.inf
        ok = np.ones(150, bool)
        ok[[7, 9, 11]] = False
        m = x[ok].astype(np.float64).max(axis=-1, keepdims=True)
        with np.errstate(divide="ignore"):
            truth = (x[ok] - m) - np.log(np.exp(x[ok].astype(np.float64) - m).sum(axis=-1, keepdims=True))
        got = dev.softmax(dev.DeviceArray.from_numpy(x), -1, log=True).numpy()
        assert np.all(np.isnan(got[~ok])) and np.all(got[ok][:, n // 2:] == -np.inf)
        assert np.allclose(got[ok][:, :n // 2], truth[:, :n // 2], rtol=1e-5, atol=1e-5)
        xb = as_bf16(x)
        gb = dev.softmax(dev.DeviceArray.from_numpy(xb, bf16=True), -1, log=True).numpy()
        assert np.all(gb[~ok] == 0x7FC0), "NaN rows must use the reference's bf16 NaN encoding"
        assert np.all(gb[ok][:, n // 2:] == 0xFF80)  # -inf
        eb = ao.bf16_to_f32(ao.softmax(xb[ok], ao.BF16, -1, log=True))[:, :n // 2]
        gf = ao.bf16_to_f32(gb[ok])[:, :n // 2]
        xf = ao.bf16_to_f32(xb[ok]).astype(np.float64)
        mb = xf.max(axis=-1, keepdims=True)
        with np.errstate(divide="ignore"):
            tb = ((xf - mb) - np.log(np.exp(xf - mb).sum(axis=-1, keepdims=True)))[:, :n // 2]
        assert np.all(np.abs(gf - tb) <= 2.0 ** -7 * np.abs(tb) + 1e-6)
        assert np.all(np.abs(gf - eb) <= 2.0 ** -7 * np.abs(eb) + 1e-6)


def test_softmax_wide_rows_two_outer_dims_and_row_pitch(gpu):
    """128 KiB fp32 rows (the four-piece fetch) under a {3, 50, 32768} tensor and under a view whose rows sit 16 elements
    apart from dense (outer strides walked with the div / mod path, every piece copy offset by the pitch)."""
    rng = np.random.default_rng(45)
    x = rand_f32(rng, (3, 50, 32768), 3.0)
    got = dev.softmax(dev.DeviceArray.from_numpy(x), -1).numpy()
    assert np.allclose(got, _f64_softmax(x), rtol=1e-5, atol=1e-12)
    wide = rand_f32(rng, (60, 32768 + 16), 3.0)
    view = dev.DeviceArray.from_numpy(wide[:, :32768])  # a view of the uploaded {60, 32784} buffer
    assert not view.is_contiguous()
    got = dev.softmax(view, -1).numpy()
    assert np.allclose(got, _f64_softmax(wide[:, :32768]), rtol=1e-5, atol=1e-12)
    assert capi.lib().axb_rownorm_last_path() == 1


@pytest.mark.parametrize("dt", ["bf16", "f16"])
@pytest.mark.parametrize("shape", [(200, 8192), (33, 32768), (150, 20000)])
def test_softmax_wide_rows_half(gpu, dt, shape):
    rng = np.random.default_rng(43)
    x = rand_f32(rng, shape, 2.0)
    if dt == "bf16":
        xh = as_bf16(x)
        got = ao.bf16_to_f32(dev.softmax(dev.DeviceArray.from_numpy(xh, bf16=True), -1).numpy())
        truth = _f64_softmax(ao.bf16_to_f32(xh))
        ulp = 2.0 ** -8
    else:
        xh = x.astype(np.float16)
        got = dev.softmax(dev.DeviceArray.from_numpy(xh), -1).numpy().astype(np.float32)
        truth = _f64_softmax(xh.astype(np.float32))
        ulp = 2.0 ** -11
    assert np.all(np.abs(got - truth) <= ulp * truth + 1e-7), float(np.max(np.abs(got - truth) / (truth + 1e-30)))


def _ln_truth(x, w, b, eps, rms=False):
    x64 = x.astype(np.float64)
    if rms:
        y = x64 / np.sqrt((x64 * x64).mean(axis=-1, keepdims=True) + eps)
        return y * w if w is not None else y
    xc = x64 - x64.mean(axis=-1, keepdims=True)
    y = xc / np.sqrt((xc * xc).mean(axis=-1, keepdims=True) + eps)
    if w is not None:
        y = y * w
    if b is not None:
        y = y + b
    return y


@pytest.mark.parametrize("shape", [(600, 8192), (1200, 16384), (593, 8320), (700, 32768), (1201, 12288)])
@pytest.mark.parametrize("affine", ["wb", "w", "none"])
def test_layer_and_rms_norm_wide_and_cluster_f32(gpu, shape, affine):
    """>= 4 x SM-count rows of >= 8192 elements with weight/bias select the SM-pair cluster kernel (parameters resident in
    shared memory, DSMEM exchange of the row statistics); without parameters the single-CTA wide kernel runs. Odd row
    counts exercise the ragged last pass, 8320 / 12288 the masked lane maps."""
    rng = np.random.default_rng(51)
    x = rand_f32(rng, shape, 2.0) + np.float32(0.75)
    n = shape[-1]
    w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32) if affine != "none" else None
    b = (0.1 * rng.standard_normal(n)).astype(np.float32) if affine == "wb" else None
    dx = dev.DeviceArray.from_numpy(x)
    dw = dev.DeviceArray.from_numpy(w) if w is not None else None
    db = dev.DeviceArray.from_numpy(b) if b is not None else None
    got = dev.layer_norm(dx, dw, db, -1, 1e-5).numpy()
    assert np.allclose(got, _ln_truth(x, w, b, 1e-5), rtol=1e-5, atol=2e-5), (shape, affine)
    if w is not None:
        got_r = dev.rms_norm(dx, dw, -1, 1e-5).numpy()
        assert np.allclose(got_r, _ln_truth(x, w, None, 1e-5, rms=True), rtol=1e-5, atol=2e-5), (shape, affine)
    # a 32-row sample against the oracle (the reference's own op sequence)
    sl = slice(100, 132)
    e, _ = ao.layer_norm(x[sl], ao.F32, w if w is not None else np.ones(n, np.float32), ao.F32,
                         b if b is not None else np.zeros(n, np.float32), ao.F32, -1, 1e-5)
    assert np.allclose(got[sl], e, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("op", ["layer_norm", "rms_norm", "softmax", "log_softmax"])
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_tma_staged_row_kernels_are_deterministic_on_l2_resident_rows(gpu, op, dt):
    """Regression for a read-before-refill race in the TMA-staged row kernels: a ring slot was re-armed once the LDS of the
    slot had been ISSUED, and a refill served from L2 (the input of a test fits the 126 MB L2) could overwrite the slot before
    the reads happened — one 256-column piece of ~10 % of the rows of layer_norm {700, 32768} came out with the NEXT row's
    values, in about half of the runs. Every repetition has to return the same bytes, and they have to be the right ones."""
    rng = np.random.default_rng(61)
    rows, n = 700, 32768
    x = rand_f32(rng, (rows, n), 2.0) + np.float32(0.75)
    w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
    b = (0.1 * rng.standard_normal(n)).astype(np.float32)
    if dt == "bf16":
        dx = dev.DeviceArray.from_numpy(as_bf16(x), bf16=True)
        x = ao.bf16_to_f32(as_bf16(x))
    else:
        dx = dev.DeviceArray.from_numpy(x)
    dw, db = dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b)
    call = {"layer_norm": lambda: dev.layer_norm(dx, dw, db, -1, 1e-5), "rms_norm": lambda: dev.rms_norm(dx, dw, -1, 1e-5),
            "softmax": lambda: dev.softmax(dx, -1), "log_softmax": lambda: dev.softmax(dx, -1, log=True)}[op]
    first = call().numpy()
    for _ in range(8):
        again = call().numpy()
        assert again.tobytes() == first.tobytes(), (op, dt, "result changes from call to call")
    got = ao.bf16_to_f32(first) if first.dtype == np.uint16 else first.astype(np.float32)
    x64 = x.astype(np.float64)
    if op in ("softmax", "log_softmax"):
        e = np.exp(x64 - x64.max(axis=-1, keepdims=True))
        truth = e / e.sum(axis=-1, keepdims=True)
        truth = np.log(truth) if op == "log_softmax" else truth
    else:
        truth = _ln_truth(x, w, b if op == "layer_norm" else None, 1e-5, rms=(op == "rms_norm"))
    tol = 1e-5 if first.dtype != np.uint16 else 2.0 ** -7
    assert np.allclose(got, truth, rtol=tol, atol=2e-5 if first.dtype != np.uint16 else 2.0 ** -7 * np.abs(truth).max()), (op, dt)


@pytest.mark.parametrize("wdt", ["f32", "bf16"])
@pytest.mark.parametrize("shape", [(1200, 16384), (601, 32768), (800, 8192)])
def test_layer_norm_cluster_bf16_rows(gpu, shape, wdt):
    """bf16 rows, Float32 or bf16 parameters, bf16 and Float32 outputs (two rows per pass in the cluster kernel)."""
    rng = np.random.default_rng(52)
    n = shape[-1]
    xb = as_bf16(rand_f32(rng, shape, 2.0) + np.float32(0.5))
    w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
    b = (0.1 * rng.standard_normal(n)).astype(np.float32)
    if wdt == "bf16":
        wq, bq = as_bf16(w), as_bf16(b)
        dw, db = dev.DeviceArray.from_numpy(wq, bf16=True), dev.DeviceArray.from_numpy(bq, bf16=True)
        w, b = ao.bf16_to_f32(wq), ao.bf16_to_f32(bq)
    else:
        dw, db = dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b)
    dx = dev.DeviceArray.from_numpy(xb, bf16=True)
    got32 = dev.layer_norm(dx, dw, db, -1, 1e-5, out_dtype=capi.F32).numpy()
    # mean is rounded to bf16 by the CPU path: compare against the oracle on a sample, and against
    # the exact result with a tolerance that covers that rounding (|mean| ~ 0.5 -> 2^-9 absolute on xc / sd ~ 2)
    sl = slice(50, 82)
    e, _ = ao.layer_norm(xb[sl], ao.BF16, w, ao.F32, b, ao.F32, -1, 1e-5)
    assert np.allclose(got32[sl], e, rtol=1e-5, atol=3e-5)
    assert np.allclose(got32, _ln_truth(ao.bf16_to_f32(xb), w, b, 1e-5), rtol=1e-2, atol=5e-3)
    got16 = dev.layer_norm(dx, dw, db, -1, 1e-5).numpy()
    assert np.array_equal(got16, ao.f32_to_bf16(got32)), "bf16 output must be the rounded Float32 output"


def test_layer_norm_cluster_nan_rows_and_strided_rows(gpu):
    """A NaN / inf input row gives an all-NaN output row with the reference's bf16 NaN encoding (hardware conversion is
    bypassed for those rows); rows that are slices of a wider tensor (outer stride != row length) take the same path."""
    rng = np.random.default_rng(53)
    n = 8192
    base = rand_f32(rng, (640, n + 64), 1.5)
    base[5, 10] = np.nan
    base[6, 11] = np.inf
    w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
    b = (0.1 * rng.standard_normal(n)).astype(np.float32)
    dbase = dev.DeviceArray.from_numpy(base)
    view = dev.DeviceArray(dbase.storage, dbase.dtype, (640, n), dbase.strides, dbase.offset)  # row stride (n + 64) * 4
    got = dev.layer_norm(view, dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b), -1, 1e-5).numpy()
    x = base[:, :n]
    ok = np.ones(640, bool)
    ok[[5, 6]] = False
    assert np.all(np.isnan(got[~ok]))
    assert np.allclose(got[ok], _ln_truth(x[ok], w, b, 1e-5), rtol=1e-5, atol=2e-5)
    xb = as_bf16(np.ascontiguousarray(x))
    gb = dev.layer_norm(dev.DeviceArray.from_numpy(xb, bf16=True), dev.DeviceArray.from_numpy(w), dev.DeviceArray.from_numpy(b),
                        -1, 1e-5).numpy()
    assert np.all(gb[~ok] == 0x7FC0)


@pytest.mark.parametrize("op", ["argmax", "argmin"])
def test_argreduce_long_rows_with_heavy_ties(gpu, op):
    """Rows drawn from 9 distinct values: thousands of exact ties per row, so any lapse in 'first occurrence wins' shows.
    {64, 300001}: the split (multi-CTA) path; {4096, 4096} / {2100, 33000}: one CTA / one warp per row."""
    rng = np.random.default_rng(61)
    for shape in ((64, 300001), (4096, 4096), (2100, 33000), (5, 1 << 20)):
        x = rng.integers(-4, 5, size=shape).astype(np.float32)
        x[0, :] = 3.0
        x[1, ::2] = np.nan
        got = dev.argreduce(op, dev.DeviceArray.from_numpy(x), 1).numpy()
        exp = (np.nanargmax if op == "argmax" else np.nanargmin)(x, axis=1)
        assert np.array_equal(got, exp), (op, shape)
        xb = as_bf16(np.nan_to_num(x))
        gb = dev.argreduce(op, dev.DeviceArray.from_numpy(xb, bf16=True), 1).numpy()
        eb = (np.argmax if op == "argmax" else np.argmin)(np.nan_to_num(x), axis=1)
        assert np.array_equal(gb, eb), (op, shape, "bf16")
        xi = np.nan_to_num(x).astype(np.int32)
        assert np.array_equal(dev.argreduce(op, dev.DeviceArray.from_numpy(xi), 1).numpy(), eb), (op, shape, "i32")


# ---- sharded full reduction: peer-mailbox exchange fused into the finish kernel (axb_reduce_sharded) ---------------------
_RANK_STREAMS = []


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min"])
def test_sharded_full_reduction_peer_mailboxes_single_gpu(gpu, op, world):
    """The protocol of axb_reduce_sharded on ONE GPU: `world` ranks = `world` mailboxes and streams of this process. Every
    rank's finish kernel stores its partial into all mailboxes and polls its own; all ranks must report the SAME value,
    equal to the rank-ordered combine of the per-shard results, over repeated calls (bank alternation, counter tags)."""
    import ctypes
    from axiom_b200 import shard
    rt = ctypes.CDLL("libcudart.so.12")
    while len(_RANK_STREAMS) < world:  # created once per process: few streams = no two of them on one hardware queue
        s = ctypes.c_void_p()
        assert rt.cudaStreamCreateWithFlags(ctypes.byref(s), 1) == 0
        _RANK_STREAMS.append(s)
    streams = _RANK_STREAMS[:world]
    boxes = [shard.PeerGroup.create_mailbox()[0] for _ in range(world)]
    groups = [shard.PeerGroup(r, world, boxes) for r in range(world)]
    rng = np.random.default_rng(5)
    sizes = [(257, 1031), (64, 4096), (3, 70001), (1, 17)][:world]
    # every device allocation of the test happens BEFORE the allocator warm-up below (a later small allocation on the
    # backend stream would take one of the blocks parked for the rank streams)
    n_calls = 5
    shards_all = [[rng.standard_normal(sz).astype(np.float32) + np.float32(r) for r, sz in enumerate(sizes)] for _ in range(n_calls)]
    xs_all = [[dev.DeviceArray.from_numpy(x) for x in shards] for shards in shards_all]
    outs_all = [[dev.DeviceArray.empty((1,), capi.F32) for _ in range(world)] for _ in range(n_calls)]
    parts_all = [[float(dev.reduce({"mean": "sum"}.get(op, op), x, [0, 1]).numpy()[0]) for x in xs] for xs in xs_all]
    dev.synchronize()
    for call in range(n_calls):
        shards, xs, outs = shards_all[call], xs_all[call], outs_all[call]
        total = sum(x.size for x in shards)
        if call == 0:
            # One process plays all ranks here, so nothing may make one rank's stream wait for another rank's polling
            # kernel. The caching allocator would: a workspace block freed on stream A (after A's finish kernel) and re-used
            # on stream B orders B after that kernel, and a class miss is a device-synchronising cudaMalloc. Give every
            # stream its own cached blocks of the small classes first (allocate on all streams, then free) and load every
            # kernel once (lazy module loading synchronises too). Separate processes have none of these couplings.
            L = capi.lib()
            held = []
            for s in streams:
                for size in range(512, 8192 + 1, 512):
                    ptr = ctypes.c_void_p()
                    capi.check(L.axb_alloc(size, s, ctypes.byref(ptr)))
                    held.append((ptr, s))
            for ptr, s in held:
                capi.check(L.axb_free(ptr, s))
            ax = (ctypes.c_int32 * 2)(0, 1)
            for r in range(world):
                capi.check(L.axb_reduce(capi.OP[op], ctypes.byref(xs[r].desc()), ax, 2, 0, ctypes.byref(outs[r].desc()), streams[r]))
            for s in streams:
                assert rt.cudaStreamSynchronize(s) == 0
        import time
        for attempt in range(3):
            dev.synchronize()
            t0 = time.perf_counter()
            for r in range(world):  # every "rank" enqueues on its own stream; the finish kernels meet on the device
                groups[r].reduce(op, xs[r], total, outs[r], streams[r])
            for s in streams:
                assert rt.cudaStreamSynchronize(s) == 0
            elapsed = time.perf_counter() - t0
            got = [float(o.numpy()[0]) for o in outs]
            if len(set(got)) == 1 or elapsed < 2.5:
                break
            # > 2.5 s: a kernel gave up polling (~3 s) because this PROCESS serialised two rank streams (see above) — a
            # harness coupling, not the protocol; the call is repeated. A disagreement without a timeout fails below.
        assert len(set(got)) == 1, (op, call, got, elapsed)  # bit-identical on every rank
        parts = parts_all[call]
        if op in ("sum", "mean"):
            acc = np.float32(parts[0])
            for v in parts[1:]:
                acc = np.float32(acc + np.float32(v))
            exp = float(acc / np.float32(total)) if op == "mean" else float(acc)
            assert abs(got[0] - exp) <= 1e-6 * max(1.0, abs(exp)), (op, call, got[0], exp)
            ref = float(np.concatenate([x.ravel() for x in shards]).astype(np.float64).sum()) / (total if op == "mean" else 1)
            assert abs(got[0] - ref) <= 1e-5 * sum(float(np.abs(x).sum()) for x in shards) / (total if op == "mean" else 1)
        else:
            assert got[0] == (max(parts) if op == "max" else min(parts))
    for g in groups:
        g._own = None
        g.close()
    L = capi.lib()
    for b in boxes:
        capi.check(L.axb_peer_mailbox_destroy(ctypes.c_void_p(b)))


# ---- chain + full reduction in ONE pass (FusedReductionStep; EW_RED in ew_prelude.cuh) -----------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min"])
@pytest.mark.parametrize("shape", [(257, 1031), (1, 5), (64, 4096), (3, 70001)])
def test_fused_chain_full_reduction_one_pass(gpu, op, shape):
    """((a + b) * c).relu() under a full reduction: the chain kernel folds its values into one partial per CTA (the chain's
    value is never written), the reduction family combines the partials. a is a {rows, 1} column (broadcast), sizes are
    ragged (vector tails). max / min are exact against the op-by-op chain; sum / mean meet the 1e-5 reduction contract."""
    rng = np.random.default_rng(9)
    a = rng.standard_normal((shape[0], 1)).astype(np.float32)
    b = rng.standard_normal(shape).astype(np.float32)
    c = rng.standard_normal(shape).astype(np.float32)
    F, V = capi.F32, capi.SRC_VALUE
    prog = [capi.make_instr(capi.OP["add"], F, 0, 1), capi.make_instr(capi.OP["mul"], F, V(0), 2),
            capi.make_instr(capi.OP["relu"], F, V(1))]
    ins = [dev.DeviceArray.from_numpy(t) for t in (a, b, c)]
    n0 = dev.launch_count()
    got = float(dev.elementwise_reduce(prog, ins, op).numpy()[0])
    assert dev.launch_count() - n0 <= 3  # chain+partials, (split) reduce, finish — no materialising launch
    chain = np.maximum((a + b) * c, np.float32(0)).astype(np.float32)  # IEEE add / mul / relu: exact op by op
    if op == "max":
        assert got == float(chain.max())
    elif op == "min":
        assert got == float(chain.min())
    else:
        ref = float(chain.astype(np.float64).sum())
        tol = 1e-5 * float(np.abs(chain).astype(np.float64).sum()) + 1e-30
        if op == "mean":
            ref, tol = ref / chain.size, tol / chain.size
        assert abs(got - ref) <= tol, (op, shape, got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min"])
@pytest.mark.parametrize("shape", [(257, 1031), (3, 70001)])
def test_fused_chain_full_reduction_bf16(gpu, op, shape):
    """bf16 chain (a + b) * c under a full reduction, one pass: the chain values are the reference's (every op rounded to
    bf16), the fold is fp32, the result is rounded once to bf16 (Mean: rounded to bf16, returned as Float32)."""
    import ctypes
    rng = np.random.default_rng(11)
    a, b, c = (as_bf16(rand_f32(rng, shape, 1.0)) for _ in range(3))
    BF, V = capi.BF16, capi.SRC_VALUE
    prog = [capi.make_instr(capi.OP["add"], BF, 0, 1), capi.make_instr(capi.OP["mul"], BF, V(0), 2)]
    ins = [dev.DeviceArray.from_numpy(t, bf16=True) for t in (a, b, c)]
    P = (capi.AxbInstr * len(prog))(*prog)
    I = (capi.AxbTensor * 3)(*[t.desc() for t in ins])
    out_dt = capi.lib().axb_reduce_result_dtype(capi.OP[op], BF)
    out = dev.DeviceArray.empty((1,), out_dt)
    n0 = dev.launch_count()
    capi.check(capi.lib().axb_elementwise_reduce(P, len(prog), I, 3, capi.OP[op], ctypes.byref(out.desc()), None))
    assert dev.launch_count() - n0 <= 5
    v, _ = ao.binary("add", a, ao.BF16, b, ao.BF16)
    v, _ = ao.binary("mul", v, ao.BF16, c, ao.BF16)
    chain = ao.bf16_to_f32(v).astype(np.float64)
    got = out.numpy()
    gotf = float(ao.bf16_to_f32(got)[0]) if out_dt == capi.BF16 else float(got[0])
    if op in ("max", "min"):
        assert gotf == float(chain.max() if op == "max" else chain.min())
    else:
        ref = chain.sum() / (chain.size if op == "mean" else 1)
        # fp32 fold (1e-5 of the absolute sum) + one rounding to bf16 (2^-8 relative)
        tol = 1e-5 * np.abs(chain).sum() / (chain.size if op == "mean" else 1) + 2.0 ** -8 * abs(ref) + 1e-30
        assert abs(gotf - ref) <= tol, (op, shape, gotf, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min"])
@pytest.mark.parametrize("shape", [(37, 5000), (1000, 300), (4, 33, 1025), (5, 2), (3, 1, 70001)])
def test_fused_chain_last_axis_reduction_one_kernel(gpu, op, shape):
    """((a + b) * c).relu() reduced over its LAST axis (axb_elementwise_reduce_axis): ONE launch — a CTA (rows > 1024 elements) or
    a warp per output row folds the chain's values as it computes them; a is a broadcast column, b a broadcast row. max / min are
    exact against the op-by-op chain, sum / mean meet the 1e-5 reduction contract of the plain reductions."""
    rng = np.random.default_rng(19)
    a = rng.standard_normal(shape[:-1] + (1,)).astype(np.float32)
    b = rng.standard_normal((shape[-1],)).astype(np.float32)
    c = rng.standard_normal(shape).astype(np.float32)
    F, V = capi.F32, capi.SRC_VALUE
    prog = [capi.make_instr(capi.OP["add"], F, 0, 1), capi.make_instr(capi.OP["mul"], F, V(0), 2),
            capi.make_instr(capi.OP["relu"], F, V(1))]
    ins = [dev.DeviceArray.from_numpy(t) for t in (a, b, c)]
    for keep in (False, True):
        n0 = dev.launch_count()
        got = dev.elementwise_reduce_axis(prog, ins, op, -1, keep).numpy()
        assert dev.launch_count() - n0 == 1, "chain + last-axis reduction must be one kernel"
        chain = np.maximum((a + b) * c, np.float32(0)).astype(np.float32)
        assert got.shape == (chain.max(axis=-1, keepdims=keep)).shape
        if op in ("max", "min"):
            assert np.array_equal(got, chain.max(axis=-1, keepdims=keep) if op == "max" else chain.min(axis=-1, keepdims=keep))
        else:
            ref = chain.astype(np.float64).sum(axis=-1, keepdims=keep)
            tol = 1e-5 * np.abs(chain).astype(np.float64).sum(axis=-1, keepdims=keep) + 1e-30
            if op == "mean":
                ref, tol = ref / shape[-1], tol / shape[-1] + 2.0 ** -23 * np.abs(ref / shape[-1])
            assert np.all(np.abs(got - ref) <= tol), (op, shape)


@pytest.mark.gpu
def test_chain_axis_reduction_unfused_cases_agree(gpu):
    """Everything outside the one-kernel case (another axis, a bf16 chain, prod) runs chain + reduction as two launches and gives
    exactly what the separate calls give."""
    rng = np.random.default_rng(20)
    x = rng.standard_normal((33, 70, 129)).astype(np.float32)
    y = rng.standard_normal((70, 1)).astype(np.float32)
    F, V = capi.F32, capi.SRC_VALUE
    prog = [capi.make_instr(capi.OP["mul"], F, 0, 1), capi.make_instr(capi.OP["abs"], F, V(0))]
    dx, dy = dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(y)
    chain = dev.elementwise(prog, [dx, dy])
    for op, axis in (("sum", 1), ("max", 0), ("prod", 2), ("mean", -2)):
        got = dev.elementwise_reduce_axis(prog, [dx, dy], op, axis).numpy()
        assert np.array_equal(got, dev.reduce(op, chain, [axis % 3]).numpy()), (op, axis)
    xb, yb = as_bf16(x), as_bf16(np.broadcast_to(y, (70, 129)).copy())
    BF = capi.BF16
    progb = [capi.make_instr(capi.OP["mul"], BF, 0, 1)]
    dxb, dyb = dev.DeviceArray.from_numpy(xb, bf16=True), dev.DeviceArray.from_numpy(yb, bf16=True)
    got = dev.elementwise_reduce_axis(progb, [dxb, dyb], "sum", -1).numpy()
    assert np.array_equal(got, dev.reduce("sum", dev.elementwise(progb, [dxb, dyb]), [2]).numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("log", [False, True])
@pytest.mark.parametrize("shape", [(37, 5000), (1000, 300), (4, 33, 1025), (5, 2), (64, 8192), (3, 1, 257)])
def test_fused_chain_softmax_one_kernel(gpu, shape, log):
    """softmax / log_softmax over the last axis of (a * 0.5 + b) * c — a broadcast column, a broadcast row, a full tensor — as ONE
    kernel (axb_elementwise_softmax): the chain's values stay in registers between max, sum and normalisation. Same 1e-5
    contract as the stand-alone softmax kernels, checked against the fp64 value of the exact op-by-op chain."""
    rng = np.random.default_rng(23)
    a = rng.standard_normal(shape[:-1] + (1,)).astype(np.float32)
    b = rng.standard_normal((shape[-1],)).astype(np.float32)
    c = (rng.standard_normal(shape) * 2).astype(np.float32)
    F, V = capi.F32, capi.SRC_VALUE
    prog = [capi.make_instr(capi.OP["mul"], F, 0, capi.SRC_IMM, imm=0.5), capi.make_instr(capi.OP["add"], F, V(0), 1),
            capi.make_instr(capi.OP["mul"], F, V(1), 2)]
    ins = [dev.DeviceArray.from_numpy(t) for t in (a, b, c)]
    n0 = dev.launch_count()
    got = dev.elementwise_softmax(prog, ins, log=log).numpy()
    assert dev.launch_count() - n0 == 1, "chain + softmax must be one kernel"
    chain = (((a * np.float32(0.5)).astype(np.float32) + b).astype(np.float32) * c).astype(np.float32).astype(np.float64)
    z = chain - chain.max(axis=-1, keepdims=True)
    ref = z - np.log(np.exp(z).sum(axis=-1, keepdims=True)) if log else np.exp(z) / np.exp(z).sum(axis=-1, keepdims=True)
    assert got.shape == ref.shape
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6 if log else 1e-8), (shape, log, float(np.abs(got - ref).max()))
    if not log:  # every row is a probability vector
        assert np.allclose(got.astype(np.float64).sum(axis=-1), 1.0, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("epi", ["sum", "mean", "max"])
def test_fused_chain_softmax_row_epilogue_and_fallbacks(gpu, epi):
    rng = np.random.default_rng(24)
    x = (rng.standard_normal((300, 1024)) * 3).astype(np.float32)
    F = capi.F32
    ident = [capi.make_instr(capi.OP["cast"], F, 0)]
    dx = dev.DeviceArray.from_numpy(x)
    n0 = dev.launch_count()
    got = dev.elementwise_softmax(ident, [dx], epilogue=epi).numpy()
    assert dev.launch_count() - n0 == 1
    z = x.astype(np.float64) - x.astype(np.float64).max(axis=1, keepdims=True)
    p = np.exp(z) / np.exp(z).sum(axis=1, keepdims=True)
    ref = {"sum": p.sum(axis=1), "mean": p.mean(axis=1), "max": p.max(axis=1)}[epi]
    assert got.shape == (300,) and np.allclose(got, ref, rtol=1e-5, atol=1e-8)
    # rows longer than 8192 elements: chain, softmax and reduction run back to back, same answer
    xl = (rng.standard_normal((7, 10000))).astype(np.float32)
    n0 = dev.launch_count()
    gl = dev.elementwise_softmax(ident, [dev.DeviceArray.from_numpy(xl)], epilogue=epi).numpy()
    assert dev.launch_count() - n0 >= 2
    zl = xl.astype(np.float64) - xl.astype(np.float64).max(axis=1, keepdims=True)
    pl = np.exp(zl) / np.exp(zl).sum(axis=1, keepdims=True)
    assert np.allclose(gl, {"sum": pl.sum(axis=1), "mean": pl.mean(axis=1), "max": pl.max(axis=1)}[epi], rtol=1e-4, atol=1e-8)
    # a NaN in a row poisons that row only; a -inf entry is an exact zero
    xn = x[:4].copy()
    xn[1, 5] = np.nan
    xn[2, 7] = -np.inf
    gn = dev.elementwise_softmax(ident, [dev.DeviceArray.from_numpy(xn)]).numpy()
    assert np.all(np.isnan(gn[1])) and not np.any(np.isnan(gn[[0, 2, 3]])) and gn[2, 7] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["layer_norm", "rms_norm"])
@pytest.mark.parametrize("shape", [(37, 5000), (1000, 300), (4, 33, 1025), (5, 2), (64, 8192)])
def test_fused_chain_norm_one_kernel(gpu, shape, kind):
    """layer_norm / rms_norm over the last axis of x + y (the post-LN pattern layer_norm(x + sublayer(x))) as ONE kernel
    (axb_elementwise_norm), with and without the affine rows; 1e-5 against the fp64 value of the reference's op sequence
    (operations.cpp:1803-1817, :1841-1857) applied to the exact fp32 sum."""
    rng = np.random.default_rng(29)
    x = (rng.standard_normal(shape) * 2 + 0.3).astype(np.float32)
    y = rng.standard_normal(shape[-1:]).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(shape[-1])).astype(np.float32)
    b = (0.1 * rng.standard_normal(shape[-1])).astype(np.float32)
    F = capi.F32
    prog = [capi.make_instr(capi.OP["add"], F, 0, 1)]
    ins = [dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(y)]
    v = (x + y).astype(np.float32).astype(np.float64)
    for affine in (True, False):
        dw = dev.DeviceArray.from_numpy(w) if affine else None
        db = dev.DeviceArray.from_numpy(b) if affine and kind == "layer_norm" else None
        n0 = dev.launch_count()
        got = dev.elementwise_norm(prog, ins, kind, dw, db, 1e-5).numpy()
        assert dev.launch_count() - n0 == 1, "chain + norm must be one kernel"
        if kind == "layer_norm":
            c = v - v.mean(axis=-1, keepdims=True)
            ref = c / np.sqrt((c * c).mean(axis=-1, keepdims=True) + 1e-5)
        else:
            ref = v / np.sqrt((v * v).mean(axis=-1, keepdims=True) + 1e-5)
        if affine:
            ref = ref * w
            if kind == "layer_norm":
                ref = ref + b
        assert np.allclose(got, ref, rtol=1e-5, atol=2e-5), (shape, kind, affine, float(np.abs(got - ref).max()))
    # rows longer than 8192 elements: chain + stand-alone kernel, same contract
    xl = rng.standard_normal((3, 9000)).astype(np.float32)
    yl = rng.standard_normal((9000,)).astype(np.float32)
    n0 = dev.launch_count()
    gl = dev.elementwise_norm(prog, [dev.DeviceArray.from_numpy(xl), dev.DeviceArray.from_numpy(yl)], kind, None, None, 1e-5).numpy()
    assert dev.launch_count() - n0 >= 2
    vl = (xl + yl).astype(np.float64)
    cl = vl - vl.mean(axis=-1, keepdims=True) if kind == "layer_norm" else vl
    assert np.allclose(gl, cl / np.sqrt((cl * cl).mean(axis=-1, keepdims=True) + 1e-5), rtol=1e-5, atol=2e-5)


@pytest.mark.gpu
def test_fused_row_kernels_at_baseline_size_agree_with_the_unfused_path(gpu):
    """BASELINE config-3 shape {8192, 32768} fp32 (and {8192, 8192} for the register-resident softmax / norm): the fused
    single-kernel paths against chain + stand-alone kernel on the SAME device data — max exact, sums / softmax / layer_norm
    inside the 1e-5 contracts — plus size-independent properties (softmax rows sum to 1, normalised rows have zero mean)."""
    F, V = capi.F32, capi.SRC_VALUE
    a, b = dev.randn((8192, 32768), 71), dev.randn((1, 32768), 72)
    prog = [capi.make_instr(capi.OP["mul"], F, 0, 1), capi.make_instr(capi.OP["abs"], F, V(0))]
    chain = dev.elementwise(prog, [a, b])
    n0 = dev.launch_count()
    fmax = dev.elementwise_reduce_axis(prog, [a, b], "max", 1)
    fsum = dev.elementwise_reduce_axis(prog, [a, b], "sum", 1)
    assert dev.launch_count() - n0 == 2
    assert np.array_equal(fmax.numpy(), dev.reduce("max", chain, [1]).numpy())
    s_ref = dev.reduce("sum", chain, [1]).numpy().astype(np.float64)
    assert np.all(np.abs(fsum.numpy() - s_ref) <= 2e-5 * s_ref)  # |values| summed: sum == sum of magnitudes
    del chain, a, b
    x, y = dev.randn((8192, 8192), 73), dev.randn((8192, 1), 74)
    prog2 = [capi.make_instr(capi.OP["add"], F, 0, 1)]
    v = dev.elementwise(prog2, [x, y])
    n0 = dev.launch_count()
    sm = dev.elementwise_softmax(prog2, [x, y]).numpy()
    ln = dev.elementwise_norm(prog2, [x, y], "layer_norm", None, None, 1e-5).numpy()
    assert dev.launch_count() - n0 == 2
    assert np.allclose(sm, dev.softmax(v, -1).numpy(), rtol=2e-5, atol=1e-9)
    assert np.allclose(sm.astype(np.float64).sum(axis=1), 1.0, atol=1e-5)
    assert np.allclose(ln, dev.layer_norm(v, None, None, -1, 1e-5).numpy(), rtol=1e-5, atol=2e-5)
    assert np.all(np.abs(ln.astype(np.float64).mean(axis=1)) < 1e-5)
