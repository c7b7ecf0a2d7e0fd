"""conv2d / conv_transpose / interpolate through the C ABI (csrc/spatial.cu) against the loop restatements in
oracle/axiom_oracle.py (src/tensor/operations.cpp:3141-3975). fp32 is bit-exact: the kernels keep the CPU loops'
accumulation type and term order. The reference's own known-answer tests pin the restatements (CPU part); the 230-case
CPU-vs-GPU sweep through the reference API is tests/test_axiom_cpp.py::test_spatial_ops_match_cpu_backend."""
import numpy as np
import pytest

from axiom_b200 import device as dev
from oracle import axiom_oracle as ao


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
