"""Generates tests/golden/reference_vectors.npz by running the UNMODIFIED reference CPU backend (oracle/_ref, built from
/root/reference by oracle/Makefile) on small seeded inputs. Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The vectors travel with the repo; tests/test_oracle_pin.py checks the restated oracle (oracle/axiom_oracle.py) against
them on every machine, including the GPU box where the reference sources do not exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

UNARY = ["Negate", "Abs", "Sqrt", "Exp", "Log", "Sin", "Cos", "Tan", "Tanh", "Erf", "GELU", "Sigmoid", "SiLU", "ReLU",
         "Square", "Reciprocal", "Sign", "Floor", "Ceil", "Trunc", "Round", "Cbrt"]
BINARY = ["Add", "Subtract", "Multiply", "Divide", "Maximum", "Minimum", "Equal", "NotEqual", "Less", "LessEqual",
          "Greater", "GreaterEqual", "LogicalAnd", "LogicalOr", "LogicalXor", "Power", "Hypot"]


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    x = (rng.standard_normal(256) * 3).astype(np.float32)
    xp = np.abs(x) + np.float32(0.01)
    out["unary_x"], out["unary_xpos"] = x, xp
    for op in UNARY:
        src = xp if op in ("Sqrt", "Log", "Reciprocal") else x
        out["unary_" + op] = ref.RefTensor.from_numpy(src).unary(op).numpy()
    a = (rng.standard_normal((16, 32)) * 2).astype(np.float32)
    b = (rng.standard_normal((16, 32)) * 2).astype(np.float32)
    b.flat[::5] = a.flat[::5]
    a.flat[::7] = 0
    out["bin_a"], out["bin_b"] = a, b
    for op in BINARY:
        aa, bb = (np.abs(a) + np.float32(0.1), b) if op == "Power" else (a, b)
        out["binary_" + op] = ref.RefTensor.from_numpy(aa).binary(op, ref.RefTensor.from_numpy(bb)).numpy()
    # broadcast pair {100,1} x {1,100} (tests/test_fusion.cpp WithBroadcast)
    c, d = (rng.standard_normal((100, 1))).astype(np.float32), (rng.standard_normal((1, 100))).astype(np.float32)
    out["bc_a"], out["bc_b"] = c, d
    out["bc_Multiply"] = ref.RefTensor.from_numpy(c).binary("Multiply", ref.RefTensor.from_numpy(d)).numpy()
    # config-2 chain at reduced size
    ca, cb, cc = (rng.standard_normal(s).astype(np.float32) for s in ((64, 1), (1, 96), (64, 96)))
    out["c2_a"], out["c2_b"], out["c2_c"] = ca, cb, cc
    out["c2_out"] = ref.config2_chain(*(ref.RefTensor.from_numpy(t) for t in (ca, cb, cc))).numpy()
    # reductions
    r = (rng.standard_normal((12, 200)) * 2).astype(np.float32)
    out["red_x"] = r
    for op in ("Sum", "Mean", "Max", "Min"):
        out[f"red_{op}_ax1"] = ref.RefTensor.from_numpy(r).reduce(op, [1]).numpy()
        out[f"red_{op}_ax0"] = ref.RefTensor.from_numpy(r).reduce(op, [0]).numpy()
        out[f"red_{op}_full"] = ref.RefTensor.from_numpy(r).reduce(op, []).numpy()
    t = np.round(r * 2) / 2
    t[3, :] = np.nan
    t[4, 5] = np.nan
    out["arg_x"] = t
    out["arg_max_ax1"] = ref.RefTensor.from_numpy(t).argreduce("ArgMax", 1).numpy()
    out["arg_min_ax1"] = ref.RefTensor.from_numpy(t).argreduce("ArgMin", 1).numpy()
    out["arg_max_ax0"] = ref.RefTensor.from_numpy(t).argreduce("ArgMax", 0).numpy()
    out["softmax_ax1"] = ref.RefTensor.from_numpy(r).softmax(-1).numpy()
    out["logsoftmax_ax1"] = ref.RefTensor.from_numpy(r).softmax(-1, True).numpy()
    out["softmax_ax0"] = ref.RefTensor.from_numpy(r).softmax(0).numpy()
    w = (1 + 0.1 * rng.standard_normal(200)).astype(np.float32)
    bb = (0.1 * rng.standard_normal(200)).astype(np.float32)
    out["ln_w"], out["ln_b"] = w, bb
    rr, rw, rb = (ref.RefTensor.from_numpy(v) for v in (r, w, bb))
    out["layer_norm"] = rr.layer_norm(rw, rb, -1, 1e-5).numpy()
    out["rms_norm"] = rr.rms_norm(rw, -1, 1e-5).numpy()
    # matmul
    ma, mb = (rng.standard_normal((24, 40))).astype(np.float32), (rng.standard_normal((40, 18))).astype(np.float32)
    out["mm_a"], out["mm_b"] = ma, mb
    out["mm_out"] = ref.RefTensor.from_numpy(ma).matmul(ref.RefTensor.from_numpy(mb)).numpy()
    out["readme_x"] = (rng.standard_normal((64, 64))).astype(np.float32)
    out["readme_out"] = ref.readme_chain(ref.RefTensor.from_numpy(out["readme_x"])).numpy()
    # bf16 rows: sum / mean dtype rules
    xb = (np.ascontiguousarray(r[:4]).view(np.uint32) >> 16).astype(np.uint16)
    out["bf16_x"] = xb
    out["bf16_sum_ax1"] = ref.RefTensor.from_numpy(xb, bf16=True).reduce("Sum", [1]).numpy()
    out["bf16_mean_ax1"] = ref.RefTensor.from_numpy(xb, bf16=True).reduce("Mean", [1]).numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
