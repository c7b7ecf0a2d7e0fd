"""Pooling, sort / argsort / topk and masked_select through the C ABI (csrc/pool_sort.cu) against the restatements in
oracle/axiom_oracle.py (src/tensor/operations.cpp:2367-2880, :4053-4257, src/backends/cpu/cpu_operations.cpp:2719-2776).

CPU part: the restatements are pinned to the reference's known answers (tests/test_pooling.cpp, tests/test_sort_topk.cpp)
and to the live reference library (oracle/_ref) on random inputs. GPU part: pooling is bit-exact for every float dtype (the
kernel keeps the CPU loops' fp64 accumulation and window order); sorted values are exact everywhere and positions exact
where keys are distinct (the reference's std::sort leaves ties unspecified; the kernel and the oracle order them by
position); masked_select is exact. The same ops through the reference's Tensor API: tests/test_axiom_cpp.py."""
import numpy as np
import pytest

from axiom_b200 import capi, device as dev
from oracle import axiom_oracle as ao
from oracle import ref


# ---- CPU: restatements against the reference's known answers ------------------------------------------------------------
def test_oracle_pool_known_answers():
    x = np.arange(1, 17, dtype=np.float32).reshape(1, 1, 4, 4)
    # tests/test_pooling.cpp AvgPool2D: 2x2 windows of 1..16 -> 3.5, 5.5, 11.5, 13.5
    assert ao.pool(x, ao.F32, "avg", (2, 2), (2, 2), (0, 0)).reshape(-1).tolist() == [3.5, 5.5, 11.5, 13.5]
    # MaxPool2D: the same windows keep their bottom-right element
    assert ao.pool(x, ao.F32, "max", (2, 2), (2, 2), (0, 0)).reshape(-1).tolist() == [6, 8, 14, 16]
    # MaxPool1D on [1,3,2,4,5,1] with k=2, s=2 -> [3,4,5]
    x1 = np.array([1, 3, 2, 4, 5, 1], np.float32).reshape(1, 1, 6)
    assert ao.pool(x1, ao.F32, "max", (2,), (2,), (0,)).reshape(-1).tolist() == [3, 4, 5]
    # padding: count_include_pad decides the divisor of the border windows
    inc = ao.pool(np.ones((1, 1, 4), np.float32), ao.F32, "avg", (3,), (1,), (1,), True).reshape(-1)
    exc = ao.pool(np.ones((1, 1, 4), np.float32), ao.F32, "avg", (3,), (1,), (1,), False).reshape(-1)
    assert inc.tolist() == [np.float32(2 / 3), 1, 1, np.float32(2 / 3)] and exc.tolist() == [1, 1, 1, 1]


def test_oracle_sort_known_answers():
    # tests/test_sort_topk.cpp: Sort1D, SortDescending, Sort2DAxis0/1, Argsort1D, ArgsortDescending, TopK1D, TopKSmallest
    v, _ = ao.sort(np.array([3, 1, 4, 1, 5, 9], np.float32))
    assert v.tolist() == [1, 1, 3, 4, 5, 9]
    assert ao.sort(np.array([3, 1, 4], np.float32), -1, True)[0].tolist() == [4, 3, 1]
    m = np.array([[3, 1], [2, 4]], np.float32)
    assert ao.sort(m, 0)[0].tolist() == [[2, 1], [3, 4]] and ao.sort(m, 1)[0].tolist() == [[1, 3], [2, 4]]
    assert ao.sort(np.array([3, 1, 4], np.float32))[1].tolist() == [1, 0, 2]
    assert ao.sort(np.array([3, 1, 4], np.float32), -1, True)[1].tolist() == [2, 0, 1]
    tv, ti = ao.sort(np.array([1, 5, 3, 9, 2], np.float32), -1, True, k=3)
    assert tv.tolist() == [9, 5, 3] and ti.tolist() == [3, 1, 2]
    assert ao.sort(np.array([1, 5, 3, 9, 2], np.float32), -1, False, k=2)[0].tolist() == [1, 2]


def test_oracle_masked_select_known_answers():
    x = np.arange(6, dtype=np.float32).reshape(2, 3)
    assert ao.masked_select(x, x > 2).tolist() == [3, 4, 5]
    assert ao.masked_select(x, np.array([True, False, True])).tolist() == [0, 2, 3, 5]  # mask broadcast over rows
    # the reference reads the FIRST BYTE of a mask element: int32 256 is "false", 257 is "true"
    assert ao.masked_select(np.array([10, 20, 30], np.float32), np.array([256, 257, 1], np.int32)).tolist() == [20, 30]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built on this machine")
def test_oracle_vs_live_reference():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3, 9, 11)).astype(np.float32)
    rx = ref.RefTensor.from_numpy(x)
    for mode in ("max", "avg"):
        for k, s, p, inc in (((2, 2), (2, 2), (0, 0), True), ((3, 2), (1, 2), (1, 1), False), ((3, 3), (2, 1), (1, 0), True)):
            assert np.array_equal(ao.pool(x, ao.F32, mode, k, s, p, inc), ref.pool(rx, mode, k, s, p, inc).numpy()), (mode, k, s, p)
    x3 = rng.standard_normal((1, 2, 5, 6, 7)).astype(np.float32)
    for mode in ("max", "avg"):
        got = ref.pool(ref.RefTensor.from_numpy(x3), mode, (2, 3, 2), (1, 2, 2), (1, 1, 0), True).numpy()
        assert np.array_equal(ao.pool(x3, ao.F32, mode, (2, 3, 2), (1, 2, 2), (1, 1, 0), True), got), mode
    xb = ao.f32_to_bf16(x)
    got = ref.pool(ref.RefTensor.from_numpy(xb, bf16=True), "avg", (3, 3), (2, 2), (1, 1), True).numpy()
    assert np.array_equal(ao.pool(xb, ao.BF16, "avg", (3, 3), (2, 2), (1, 1), True), got)
    # sort family: distinct keys (a permutation), so positions are determined
    y = rng.permutation(5 * 37 * 3).astype(np.float32).reshape(5, 37, 3)
    ry = ref.RefTensor.from_numpy(y)
    for axis in (0, 1, 2, -1):
        for desc in (False, True):
            v, i = ao.sort(y, axis, desc)
            assert np.array_equal(v, ref.sort(ry, axis, desc).numpy()), (axis, desc)
            assert np.array_equal(i, ref.argsort(ry, axis, desc).numpy()), (axis, desc)
    tv, ti = ref.topk(ry, 4, 1, True)
    ev, ei = ao.sort(y, 1, True, k=4)
    assert np.array_equal(ev, tv.numpy()) and np.array_equal(ei, ti.numpy())
    yi = rng.permutation(400).astype(np.int64).reshape(20, 20) - 200
    v, i = ao.sort(yi, 0, False)
    assert np.array_equal(v, ref.sort(ref.RefTensor.from_numpy(yi), 0, False).numpy())
    assert np.array_equal(i, ref.argsort(ref.RefTensor.from_numpy(yi), 0, False).numpy())
    # masked_select with a broadcast mask and a non-bool mask
    m = rng.standard_normal((9, 11)) > 0.3
    assert np.array_equal(ao.masked_select(x, m), ref.masked_select(rx, ref.RefTensor.from_numpy(m)).numpy())
    mi = rng.integers(0, 3, (2, 3, 9, 11)).astype(np.int32) * 128  # 0, 128, 256: the last one has a zero first byte
    assert np.array_equal(ao.masked_select(x, mi), ref.masked_select(rx, ref.RefTensor.from_numpy(mi)).numpy())


# ---- GPU parity ----------------------------------------------------------------------------------------------------------
POOL_CASES = [((2, 3, 40), (3,), (2,), (1,)), ((2, 3, 40), (4,), None, (0,)), ((2, 2, 13, 17), (3, 2), (2, 1), (1, 1)),
              ((1, 4, 16, 16), (2, 2), None, (0, 0)), ((1, 2, 6, 7, 9), (2, 3, 2), (1, 2, 2), (1, 1, 0)),
              ((3, 1, 8, 8), (8, 8), (8, 8), (0, 0))]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,k,s,p", POOL_CASES)
@pytest.mark.parametrize("mode", ["max", "avg"])
@pytest.mark.parametrize("dt", ["f32", "f64", "bf16", "f16"])
def test_pool_bit_exact(gpu, shape, k, s, p, mode, dt):
    rng = np.random.default_rng(11)
    x32 = (rng.standard_normal(shape) * 3).astype(np.float32)
    code = {"f32": ao.F32, "f64": ao.F64, "bf16": ao.BF16, "f16": ao.F16}[dt]
    x = {"f32": x32, "f64": x32.astype(np.float64) * 1.000001, "bf16": ao.f32_to_bf16(x32), "f16": x32.astype(np.float16)}[dt]
    for inc in (True, False):
        got = dev.pool(dev.DeviceArray.from_numpy(x, bf16=dt == "bf16"), mode, k, s, p, inc).numpy()
        exp = ao.pool(x, code, mode, k, s, p, inc)
        assert got.shape == exp.shape and got.tobytes() == exp.tobytes(), (mode, dt, inc)


@pytest.mark.gpu
def test_pool_strided_input_and_nan_rule(gpu):
    rng = np.random.default_rng(12)
    big = rng.standard_normal((2, 3, 10, 24)).astype(np.float32)
    v = big[:, :, ::-1, ::2]  # flipped rows, every second column: the kernel reads the view in place
    got = dev.pool(dev.DeviceArray.from_numpy(v), "avg", (3, 3), (2, 2), (1, 1), False).numpy()
    assert got.tobytes() == ao.pool(np.ascontiguousarray(v), ao.F32, "avg", (3, 3), (2, 2), (1, 1), False).tobytes()
    x = rng.standard_normal((1, 1, 8)).astype(np.float32)
    x[0, 0, 2] = np.nan  # `val > max` never lets a NaN in; an all-NaN window stays at -inf
    x[0, 0, 6:8] = np.nan
    got = dev.pool(dev.DeviceArray.from_numpy(x), "max", (2,), (2,), (0,)).numpy()
    exp = ao.pool(x, ao.F32, "max", (2,), (2,), (0,))
    assert got.tobytes() == exp.tobytes() and exp[0, 0, 3] == -np.inf and exp[0, 0, 1] == x[0, 0, 3]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,axis", [((6,), 0), ((37,), -1), ((5, 300, 3), 1), ((4, 7, 2049), 2), ((3000, 5), 0),
                                        ((2, 70000), 1), ((3, 4097, 2), 1)])
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.int64])
def test_sort_argsort_exact(gpu, shape, axis, dt):
    rng = np.random.default_rng(21)
    n = int(np.prod(shape))
    x = (rng.permutation(n).astype(np.float64) - n // 2).reshape(shape)  # distinct keys
    x = (x * (0.37 if dt in (np.float32, np.float64) else 1)).astype(dt)
    d = dev.DeviceArray.from_numpy(x)
    for desc in (False, True):
        v, i = dev.sort(d, axis, desc)
        ev, ei = ao.sort(x, axis, desc)
        assert np.array_equal(v.numpy(), ev), (shape, axis, dt, desc)
        assert np.array_equal(i.numpy(), ei), (shape, axis, dt, desc)


@pytest.mark.gpu
def test_sort_with_ties_special_values_and_topk(gpu):
    rng = np.random.default_rng(22)
    x = rng.integers(-5, 5, (7, 1000)).astype(np.float32)  # heavy ties
    x[0, :5] = [np.inf, -np.inf, 0.0, -0.0, 3.5]
    d = dev.DeviceArray.from_numpy(x)
    for desc in (False, True):
        v, i = dev.sort(d, 1, desc)
        v, i = v.numpy(), i.numpy()
        assert np.array_equal(v, np.sort(x, axis=1)[:, ::-1] if desc else np.sort(x, axis=1))  # values: unique answer
        assert np.array_equal(np.take_along_axis(x, i, 1), v)  # positions name the right elements
        assert np.array_equal(np.sort(i, axis=1), np.broadcast_to(np.arange(1000), i.shape))  # ... and are a permutation
        assert np.array_equal(i, ao.sort(x, 1, desc)[1])  # ties (-0 == +0 included) by ascending position, like the restatement
    # topk = first k of the order (largest: descending), along a middle axis, values only / indices only
    y = rng.permutation(4 * 5000 * 3).astype(np.float32).reshape(4, 5000, 3)
    dy = dev.DeviceArray.from_numpy(y)
    for largest in (True, False):
        v, i = dev.sort(dy, 1, largest, k=17)
        ev, ei = ao.sort(y, 1, largest, k=17)
        assert np.array_equal(v.numpy(), ev) and np.array_equal(i.numpy(), ei)
    v, i = dev.sort(dy, 1, True, k=1, want_indices=False)
    assert i is None and np.array_equal(v.numpy(), y.max(axis=1, keepdims=True))
    v, i = dev.sort(dy, 1, False, k=2, want_values=False)
    assert v is None and np.array_equal(i.numpy(), np.argsort(y, axis=1)[:, :2])
    with pytest.raises(capi.AxbError):
        dev.sort(dev.DeviceArray.from_numpy(np.zeros((3, 4), np.float16)), 1)  # the caller converts to float32 first
    with pytest.raises(capi.AxbError):
        dev.sort(dy, 1, True, k=5001)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int8, np.uint16, np.int64, np.bool_])
def test_masked_select_exact(gpu, dt):
    rng = np.random.default_rng(31)
    x = (rng.standard_normal((5, 33, 129)) * 50).astype(dt)
    m = rng.standard_normal((5, 33, 129)) > 0.2
    got = dev.masked_select(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(m)).numpy()
    assert np.array_equal(got, ao.masked_select(x, m))
    mb = rng.standard_normal((33, 1)) > 0  # broadcast mask; transposed (strided) input
    xt = np.ascontiguousarray(x.transpose(0, 2, 1)).transpose(0, 2, 1)
    got = dev.masked_select(dev.DeviceArray.from_numpy(xt), dev.DeviceArray.from_numpy(mb)).numpy()
    assert np.array_equal(got, ao.masked_select(x, mb))


@pytest.mark.gpu
def test_masked_select_edges(gpu):
    x = np.arange(10, dtype=np.float32)
    none = dev.masked_select(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(np.zeros(10, np.bool_)))
    assert none.shape == (0,)
    allm = dev.masked_select(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(np.ones(10, np.bool_))).numpy()
    assert np.array_equal(allm, x)
    # first byte of the mask element decides (cpu_operations.cpp:2738-2741)
    mi = np.array([256, 257, 1, 0, 512, 3, 65536, 255, 1 << 24, 2], np.int32)
    got = dev.masked_select(dev.DeviceArray.from_numpy(x), dev.DeviceArray.from_numpy(mi)).numpy()
    assert np.array_equal(got, ao.masked_select(x, mi)) and got.tolist() == [1, 2, 5, 7, 9]
    # large enough for several scan chunks, with the count checked against numpy
    rng = np.random.default_rng(32)
    big = rng.standard_normal(3_000_001).astype(np.float32)
    got = dev.masked_select(dev.DeviceArray.from_numpy(big), dev.DeviceArray.from_numpy(big > 1.0)).numpy()
    assert np.array_equal(got, big[big > 1.0])
