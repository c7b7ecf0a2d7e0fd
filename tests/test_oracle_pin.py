"""Pins the CPU oracle (oracle/axiom_oracle.py + its C restatement) BEFORE anything trusts it:
  (1) the reference's own known-answer tests for this path (tests/test_tensor_reductions.cpp:14-63,
      tests/test_tensor_matmul.cpp:4-60, tests/test_tensor_argmax_argmin.cpp, SURVEY App. A.8 measured rows);
  (2) golden vectors produced by the UNMODIFIED reference library (tests/golden/make_golden.py, run where
      /root/reference exists) — checked on every machine, no reference needed at test time;
  (3) when oracle/_ref is present, fresh random inputs through the live reference library.
No GPU, no compute call into libaxb."""
import os

import numpy as np
import pytest

from oracle import axiom_oracle as ao
from oracle import ref
from tests.helpers import ulp_diff

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))

U = dict(Negate="neg", Abs="abs", Sqrt="sqrt", Exp="exp", Log="log", Sin="sin", Cos="cos", Tan="tan", Tanh="tanh",
         Erf="erf", GELU="gelu", Sigmoid="sigmoid", SiLU="silu", ReLU="relu", Square="square", Reciprocal="reciprocal",
         Sign="sign", Floor="floor", Ceil="ceil", Trunc="trunc", Round="round", Cbrt="cbrt")
B = dict(Add="add", Subtract="sub", Multiply="mul", Divide="div", Maximum="maximum", Minimum="minimum", Equal="eq",
         NotEqual="ne", Less="lt", LessEqual="le", Greater="gt", GreaterEqual="ge", LogicalAnd="logical_and",
         LogicalOr="logical_or", LogicalXor="logical_xor", Power="pow", Hypot="hypot")


# ---- (1) the reference's own known answers ------------------------------------------------------------------------
def test_known_answers_from_reference_tests():
    x = np.arange(6, dtype=np.float32).reshape(2, 3)
    assert ao.reduce("sum", x, ao.F32)[0].tolist() == [15.0]
    assert ao.reduce("sum", x, ao.F32, [0])[0].tolist() == [3.0, 5.0, 7.0]
    assert ao.reduce("sum", x, ao.F32, [1])[0].tolist() == [3.0, 12.0]
    assert ao.reduce("mean", x, ao.F32)[0].tolist() == [2.5]
    assert ao.reduce("max", x, ao.F32)[0].tolist() == [5.0]
    assert ao.reduce("min", x, ao.F32, [0])[0].tolist() == [0.0, 1.0, 2.0]
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.arange(12, dtype=np.float32).reshape(3, 4)
    assert ao.matmul(a, ao.F32, b, ao.F32)[0].tolist() == [[20, 23, 26, 29], [56, 68, 80, 92]]
    rows = np.array([[1, 5, 5], [np.nan, 2, 2], [np.nan, np.nan, np.nan]], np.float32)
    assert ao.argreduce("argmax", rows, ao.F32, 1).tolist() == [1, 1, 0]  # SURVEY App. A.8 [measured]
    ln, _ = ao.layer_norm(np.array([[1, 2, 3, 4]], np.float32), ao.F32, np.ones(4, np.float32), ao.F32,
                          np.zeros(4, np.float32), ao.F32, -1, 0.0)
    assert np.allclose(ln, (np.array([1, 2, 3, 4]) - 2.5) / np.sqrt(1.25), rtol=1e-6)


# ---- (2) golden vectors from the unmodified reference ----------------------------------------------------------------
@pytest.mark.parametrize("op", sorted(U))
def test_unary_matches_golden(op):
    src = G["unary_xpos"] if op in ("Sqrt", "Log", "Reciprocal") else G["unary_x"]
    got, _ = ao.unary(U[op], src, ao.F32, lanes=16)
    d = int(ulp_diff(got, G["unary_" + op]).max())
    assert d <= (2 if op == "Cbrt" else 0), f"{op}: {d} ulp from the reference"


@pytest.mark.parametrize("op", sorted(B))
def test_binary_matches_golden(op):
    a, b = G["bin_a"], G["bin_b"]
    if op == "Power":
        a = np.abs(a) + np.float32(0.1)
    got, _ = ao.binary(B[op], a, ao.F32, b, ao.F32, lanes=16)
    exp = G["binary_" + op]
    if exp.dtype == np.bool_ or op not in ("Power", "Hypot"):
        assert np.array_equal(np.asarray(got).astype(exp.dtype), exp, equal_nan=True), op
    else:
        assert int(ulp_diff(np.asarray(got, np.float32), exp).max()) <= 2, op


def test_broadcast_and_config2_chain_match_golden():
    got, _ = ao.binary("mul", G["bc_a"], ao.F32, G["bc_b"], ao.F32)
    assert np.array_equal(got, G["bc_Multiply"])
    a, b, c = G["c2_a"], G["c2_b"], G["c2_c"]
    v, _ = ao.unary("relu", a, ao.F32)
    v, _ = ao.binary("add", v, ao.F32, np.array([1.0], np.float32), ao.F32)
    v, _ = ao.binary("mul", v, ao.F32, b, ao.F32)
    m, _ = ao.binary("gt", c, ao.F32, np.array([0], np.int32), ao.I32)
    o, _ = ao.where(m, ao.BOOL, v, ao.F32, np.zeros(1, np.float32), ao.F32)
    assert np.array_equal(o, G["c2_out"])


@pytest.mark.parametrize("op", ["Sum", "Mean", "Max", "Min"])
def test_reductions_match_golden_bit_exact(op):
    x = G["red_x"]
    for tag, axes in (("ax1", [1]), ("ax0", [0]), ("full", [])):
        got, _ = ao.reduce(op.lower(), x, ao.F32, axes, lanes=16)
        exp = G[f"red_{op}_{tag}"]
        # the oracle restates the reference's summation order (16 lane-strided partials, Highway's shuffle tree, scalar
        # tail): bit-exact on the AVX-512 host that produced the vectors
        assert np.array_equal(np.asarray(got).reshape(exp.shape), exp), (op, tag)


def test_argreduce_softmax_norms_matmul_match_golden():
    t = G["arg_x"]
    assert np.array_equal(ao.argreduce("argmax", t, ao.F32, 1), G["arg_max_ax1"])
    assert np.array_equal(ao.argreduce("argmin", t, ao.F32, 1), G["arg_min_ax1"])
    assert np.array_equal(ao.argreduce("argmax", t, ao.F32, 0), G["arg_max_ax0"])
    x = G["red_x"]
    assert np.array_equal(ao.softmax(x, ao.F32, -1), G["softmax_ax1"])
    assert np.allclose(ao.softmax(x, ao.F32, -1, True), G["logsoftmax_ax1"], rtol=0, atol=2e-6)
    assert np.array_equal(ao.softmax(x, ao.F32, 0), G["softmax_ax0"])
    ln, _ = ao.layer_norm(x, ao.F32, G["ln_w"], ao.F32, G["ln_b"], ao.F32, -1, 1e-5)
    assert np.allclose(ln, G["layer_norm"], rtol=1e-6, atol=1e-6)
    rn, _ = ao.rms_norm(x, ao.F32, G["ln_w"], ao.F32, -1, 1e-5)
    assert np.allclose(rn, G["rms_norm"], rtol=1e-6, atol=1e-6)
    mm, _ = ao.matmul(G["mm_a"], ao.F32, G["mm_b"], ao.F32)
    assert np.allclose(mm, G["mm_out"], rtol=1e-5, atol=1e-5)  # sgemm summation order is the BLAS's own
    xr = G["readme_x"]
    s, _ = ao.matmul(xr, ao.F32, xr.T, ao.F32)
    r, _ = ao.reduce("sum", ao.softmax(np.ascontiguousarray(s), ao.F32, -1), ao.F32, [1])
    assert np.allclose(r, G["readme_out"], atol=1e-5)


def test_bf16_reduction_dtype_rules_match_golden():
    xb = G["bf16_x"]
    s, sdt = ao.reduce("sum", xb, ao.BF16, [1])
    assert sdt == ao.BF16 and np.array_equal(s, G["bf16_sum_ax1"])
    m, mdt = ao.reduce("mean", xb, ao.BF16, [1])
    assert mdt == ao.F32 and np.array_equal(m, G["bf16_mean_ax1"])  # Float32 holding a bf16-rounded value


# ---- (3) live reference, fresh inputs -------------------------------------------------------------------------------
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built on this machine")
def test_oracle_vs_live_reference_random():
    rng = np.random.default_rng(99)
    x = (rng.standard_normal((8, 304)) * 2).astype(np.float32)
    rx = ref.RefTensor.from_numpy(x)
    for op in ("Sum", "Max", "Min", "Mean"):
        got, _ = ao.reduce(op.lower(), x, ao.F32, [1], lanes=16)
        assert np.allclose(got, rx.reduce(op, [1]).numpy(), rtol=1e-6, atol=1e-6), op
    assert np.allclose(ao.softmax(x, ao.F32, -1), rx.softmax(-1).numpy(), rtol=1e-6, atol=1e-9)
    y = (rng.standard_normal((8, 304))).astype(np.float32)
    for op in ("Add", "Multiply", "Greater", "Maximum"):
        got, _ = ao.binary(B[op], x, ao.F32, y, ao.F32)
        assert np.array_equal(np.asarray(got), rx.binary(op, ref.RefTensor.from_numpy(y)).numpy()), op
    assert np.array_equal(ao.argreduce("argmax", x, ao.F32, 1), rx.argreduce("ArgMax", 1).numpy())
