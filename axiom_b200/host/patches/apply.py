#!/usr/bin/env python3
"""Build-time patch series for the reference TUs the CUDA seam has to change (SURVEY.md §7 "Build mechanics": the
reference tree is read-only, so patched COPIES are generated into the git-ignored build directory; no reference source
is stored in this repo — only the substitutions below).

usage: apply.py <reference file> <output file>

Every rule names the reference lines it rewrites and must match EXACTLY the stated number of times, otherwise the build
fails loudly (a changed upstream file must be re-examined, not silently mis-patched)."""
import re
import sys

RULES = {
    # Tensor::where(cond, scalar) materialises a FULL-SIZE constant (`Tensor::full(shape_, v, device())`: host fill of
    # numel elements + H2D upload — 256 MiB per call for BASELINE config 2) before the backend sees anything
    # (src/tensor/tensor.cpp:1340-1355; Tensor::full is a host fill + transfer, include/axiom/tensor.hpp:636-647).
    # ops::where broadcasts its three operands (src/tensor/operations.cpp:1547-1570), so a {1} constant gives the same
    # result shape, dtype and values; on Device::GPU it is a 4-byte upload and the select reads it with stride 0.
    # Precedent for a device-side constant in the reference: gpu_fill_constant (src/backends/metal/metal_operations.hpp:57-58).
    "src/tensor/tensor.cpp": [
        (r"Tensor::full\(shape_, other, device\(\)\)", "Tensor::full({1}, other, device())", 1),
        (r"Tensor::full\(shape_, static_cast<float>\(other\), device\(\)\)",
         "Tensor::full({1}, static_cast<float>(other), device())", 2),
    ],
}


def main() -> int:
    src, dst = sys.argv[1], sys.argv[2]
    key = next((k for k in RULES if src.endswith(k)), None)
    if key is None:
        print(f"apply.py: no patch rules for {src}", file=sys.stderr)
        return 2
    text = open(src).read()
    for pattern, repl, count in RULES[key]:
        text, n = re.subn(pattern, repl, text)
        if n != count:
            print(f"apply.py: {key}: pattern {pattern!r} matched {n} times, expected {count} — upstream changed, "
                  "re-examine the patch", file=sys.stderr)
            return 1
    with open(dst, "w") as f:
        f.write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
