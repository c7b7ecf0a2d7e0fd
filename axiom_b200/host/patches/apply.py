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
        # Tensor::to(Device) hands the WHOLE storage of a contiguous view to Storage::copy_from (src/tensor/tensor.cpp:1625-
        # 1636): a contiguous view that starts inside its storage (`x.slice({Slice(1, 2)})`, a batch slice) arrives on the
        # other device with the bytes of offset 0. On Apple silicon the unified-memory branch above it (:1600-1622) keeps the
        # offset, so the reference never takes this branch with a GPU; a discrete GPU always does. The view is compacted on
        # its own device first (host memcpy / device gather), then copied as before.
        (r"\n    const Tensor &src = is_contiguous\(\) \? \*this : ascontiguousarray\(\);\n",
         "\n    Tensor compact_;\n"
         "    if (is_contiguous() && offset_ != 0 && size() > 0 && memory_order_ == MemoryOrder::RowMajor) {\n"
         "        if (device() == Device::CPU) {\n"
         "            compact_ = Tensor(shape_, dtype_, Device::CPU, memory_order_);\n"
         "            std::memcpy(compact_.data(), data(), nbytes());\n"
         "        } else {\n"
         "            compact_ = backends::cuda::gpu_compact(*this);\n"
         "        }\n"
         "    }\n"
         "    const Tensor &src = compact_.storage_ ? compact_ : (is_contiguous() ? *this : ascontiguousarray());\n", 1),
        # Tensor::zeros / ones on Device::GPU fill a host tensor and upload it (src/tensor/tensor.cpp:1785-1810). The reference
        # already has the device-side fill hook for its own GPU backend (gpu_fill_constant, metal_operations.hpp:57-58): use it.
        (r"    auto tensor = Tensor\(shape, dtype, Device::CPU, order\);\n    std::memset\(tensor\.data\(\), 0, tensor\.nbytes\(\)\);\n",
         "    if (device == Device::GPU && order == MemoryOrder::RowMajor && !is_complex_dtype(dtype))\n"
         "        return backends::metal::gpu_fill_constant(shape, dtype, 0.0);\n"
         "    auto tensor = Tensor(shape, dtype, Device::CPU, order);\n    std::memset(tensor.data(), 0, tensor.nbytes());\n", 1),
        (r"    auto tensor = Tensor\(shape, dtype, Device::CPU, order\);\n    dispatch\(dtype, \[&\]<typename DT>\(DT\) \{\n        tensor\.fill<typename DT::value_type>\(DT::one\(\)\);\n",
         "    if (device == Device::GPU && order == MemoryOrder::RowMajor && !is_complex_dtype(dtype))\n"
         "        return backends::metal::gpu_fill_constant(shape, dtype, 1.0);\n"
         "    auto tensor = Tensor(shape, dtype, Device::CPU, order);\n    dispatch(dtype, [&]<typename DT>(DT) {\n        tensor.fill<typename DT::value_type>(DT::one());\n", 1),
        # Tensor::eye on Device::GPU returns all zeros upstream (the diagonal is only written `if (device == Device::CPU)`,
        # :1827-1839; Apple's unified memory hides nothing here — it is simply missing): build on the host, then transfer.
        (r"Tensor Tensor::eye\(size_t n, DType dtype, Device device, MemoryOrder order\) \{\n    auto tensor = zeros\(\{n, n\}, dtype, device, order\);\n",
         "Tensor Tensor::eye(size_t n, DType dtype, Device device, MemoryOrder order) {\n"
         "    if (device == Device::GPU)\n        return eye(n, dtype, Device::CPU, order).to(Device::GPU, order);\n"
         "    auto tensor = zeros({n, n}, dtype, device, order);\n", 1),
        (r"\nnamespace axiom \{\n\ntemplate <typename T> std::string vec_to_string",
         "\nnamespace axiom { namespace backends { namespace cuda { Tensor gpu_compact(const Tensor &); }}}\n"
         "\nnamespace axiom {\n\ntemplate <typename T> std::string vec_to_string", 1),
    ],
    # Pooling, sort, argsort and topk have NO device branch in the reference: a GPU tensor is copied to the host, processed by
    # scalar loops and (pooling only) uploaded again (src/tensor/operations.cpp:2367-2880, :4199-4257). The rules below add
    # the branch the other spatial operators already have (cf. gpu_conv_transpose1d at :3395, gpu_interpolate at :3742):
    # on Device::GPU the call goes to the CUDA backend (host/cuda_operations.cpp gpu_pool / gpu_sort / gpu_topk) AFTER the
    # reference's own argument checks and shape computation, so errors and result shapes are the reference's.
    "src/tensor/operations.cpp": [
        # declarations, placed in front of the first namespace block
        (r"\nnamespace axiom \{\nnamespace ops \{\n",
         "\nnamespace axiom { namespace backends { namespace cuda {\n"
         "Tensor gpu_pool(const Tensor &, bool, const std::vector<int> &, const std::vector<int> &, const std::vector<int> &, bool, const Shape &);\n"
         "Tensor gpu_sort(const Tensor &, int, bool, bool);\n"
         "std::pair<Tensor, Tensor> gpu_topk(const Tensor &, int, int, bool);\n"
         "}}} // namespace axiom::backends::cuda\n"
         "\nnamespace axiom {\nnamespace ops {\n", 1),
    ],
    # io::safetensors::load / load_tensor read every tensor into a pageable host tensor and then call Tensor::gpu()
    # (src/io/io_safetensors.cpp:123-160, :175-214). On Device::GPU the patched build hands the parsed metadata to the CUDA
    # backend, which reads the file straight into pinned staging buffers and overlaps their DMA with the next read
    # (host/cuda_storage.cpp load_file_tensors -> axb_upload_file). Header parsing, dtype names, error types stay the reference's.
    "src/io/io_safetensors.cpp": [
        (r"\nnamespace axiom \{\nnamespace io \{\nnamespace safetensors \{\n",
         "\nnamespace axiom { namespace backends { namespace cuda {\n"
         "struct FileTensor { std::string name; DType dtype; Shape shape; size_t file_offset; size_t bytes; };\n"
         "std::map<std::string, Tensor> load_file_tensors(const std::string &, const std::vector<FileTensor> &);\n"
         "}}} // namespace axiom::backends::cuda\n"
         "\nnamespace axiom {\nnamespace io {\nnamespace safetensors {\n", 1),
    ],
}

# Function-scoped insertions: (file, regex matching the START of a function definition, regex of the anchor line inside it,
# text inserted in front of the anchor). The anchor is the first match after the function start.
_POOL_ANCHOR = r"    Tensor input_cpu = input\.device\(\) == Device::CPU \? input : input\.cpu\(\);\n"
_SORT_ANCHOR = r"    auto src = input\.is_contiguous\(\) \? input : input\.ascontiguousarray\(\);\n"


def _pool_call(is_avg, k, s, p, inc):
    return ("    if (input.device() == Device::GPU)\n"
            f"        return backends::cuda::gpu_pool(input, {is_avg}, {k}, {s}, {p}, {inc}, out_shape);\n")


INSERTIONS = {
    "src/io/io_safetensors.cpp": [
        (r"\nstd::map<std::string, Tensor> load\(const std::string &filename, Device device\) \{\n",
         r"    std::ifstream file\(filename, std::ios::binary\);\n",
         "    if (device == Device::GPU) {\n"
         "        std::vector<backends::cuda::FileTensor> items;\n"
         "        for (const auto &meta : tensor_metas)\n"
         "            items.push_back({meta.name, meta.dtype, meta.shape, data_offset + meta.data_start, meta.data_end - meta.data_start});\n"
         "        return backends::cuda::load_file_tensors(filename, items);\n"
         "    }\n"),
        (r"\nTensor load_tensor\(const std::string &filename, const std::string &tensor_name,\n",
         r"    size_t byte_count = found->data_end - found->data_start;\n",
         "    if (device == Device::GPU)\n"
         "        return backends::cuda::load_file_tensors(filename, {{found->name, found->dtype, found->shape, data_offset + found->data_start,\n"
         "                                                             found->data_end - found->data_start}}).begin()->second;\n"),
    ],
    "src/tensor/operations.cpp": [
        (r"\nTensor max_pool1d\(", _POOL_ANCHOR, _pool_call("false", "{kernel_size}", "{stride}", "{padding}", "false")),
        (r"\nTensor avg_pool1d\(", _POOL_ANCHOR, _pool_call("true", "{kernel_size}", "{stride}", "{padding}", "count_include_pad")),
        (r"\nTensor max_pool2d\(", _POOL_ANCHOR, _pool_call("false", "kernel_size", "actual_stride", "actual_padding", "false")),
        (r"\nTensor avg_pool2d\(", _POOL_ANCHOR, _pool_call("true", "kernel_size", "actual_stride", "actual_padding", "count_include_pad")),
        (r"\nTensor max_pool3d\(", _POOL_ANCHOR, _pool_call("false", "kernel_size", "actual_stride", "actual_padding", "false")),
        (r"\nTensor avg_pool3d\(", _POOL_ANCHOR, _pool_call("true", "kernel_size", "actual_stride", "actual_padding", "count_include_pad")),
        (r"\nTensor sort\(const Tensor &input, int axis, bool descending\) \{\n", _SORT_ANCHOR,
         "    if (input.device() == Device::GPU)\n        return backends::cuda::gpu_sort(input, axis, descending, false);\n"),
        (r"\nTensor argsort\(const Tensor &input, int axis, bool descending\) \{\n", _SORT_ANCHOR,
         "    if (input.device() == Device::GPU)\n        return backends::cuda::gpu_sort(input, axis, descending, true);\n"),
        (r"\nstd::pair<Tensor, Tensor> topk\(const Tensor &input, int k, int axis,\n", _SORT_ANCHOR,
         "    if (input.device() == Device::GPU)\n        return backends::cuda::gpu_topk(input, k, axis, largest);\n"),
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
        text, n = re.subn(pattern, lambda m, r=repl: r, text)
        if n != count:
            print(f"apply.py: {key}: pattern {pattern!r} matched {n} times, expected {count} — upstream changed, "
                  "re-examine the patch", file=sys.stderr)
            return 1
    for start, anchor, insertion in INSERTIONS.get(key, []):
        heads = list(re.finditer(start, text))
        if len(heads) != 1:
            print(f"apply.py: {key}: function start {start!r} matched {len(heads)} times, expected 1", file=sys.stderr)
            return 1
        m = re.compile(anchor).search(text, heads[0].end())
        nxt = re.compile(r"\n}\n").search(text, heads[0].end())  # end of that function (first closing brace in column 0)
        if m is None or nxt is None or m.start() > nxt.start():
            print(f"apply.py: {key}: anchor {anchor!r} not found inside the function starting at {start!r}", file=sys.stderr)
            return 1
        text = text[:m.start()] + insertion + text[m.start():]
    with open(dst, "w") as f:
        f.write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
