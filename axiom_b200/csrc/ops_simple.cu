// ops_simple.cu — type rules and the single-op entry points (Operation::execute_unary/binary/where/cast/
// masked_fill equivalents) expressed as 1-3 instruction programs of the fused elementwise family, plus the
// device fill and counter-based normal RNG used to build synthetic inputs on device.
#include <algorithm>
#include <cstring>

#include "axb_internal.h"

using namespace axb;

namespace {

// promote_types — restated from src/tensor/operations.cpp:346-397
int rank_of(int dt) {
    switch (dt) {
    case AXB_BOOL: return 0;
    case AXB_I8: return 1;
    case AXB_U8: return 2;
    case AXB_I16: return 3;
    case AXB_U16: return 4;
    case AXB_I32: return 5;
    case AXB_U32: return 6;
    case AXB_I64: return 7;
    case AXB_U64: return 8;
    case AXB_F16: return 9;
    case AXB_BF16: return 10;
    case AXB_F32: return 11;
    case AXB_F64: return 12;
    case AXB_C64: return 13;
    case AXB_C128: return 14;
    default: return -1;
    }
}

__global__ void randn_kernel(float *out, int64_t n, uint64_t seed, float mean, float std_) {
    // counter-based: each thread derives two normals from a splitmix64 hash of (seed, pair index)
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    for (; i < n; i += stride) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i / 2 + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        uint32_t a = (uint32_t)z, b = (uint32_t)(z >> 32);
        float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777217.0f); // (0,1)
        float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);          // [0,1)
        float r = sqrtf(-2.0f * logf(u1));
        float s, c;
        sincospif(2.0f * u2, &s, &c);
        out[i] = mean + std_ * r * c;
        if (i + 1 < n)
            out[i + 1] = mean + std_ * r * s;
    }
}

} // namespace

extern "C" {

int32_t axb_promote_types(int32_t a, int32_t b) {
    if (a == b)
        return a;
    int ra = rank_of(a), rb = rank_of(b);
    if (ra < 0 || rb < 0)
        return -1;
    int t1 = ra > rb ? a : b, t2 = ra > rb ? b : a;
    if (is_complex_dtype(t1) || is_float_dtype(t1))
        return t1;
    bool s1 = is_signed_int_dtype(t1), s2 = is_signed_int_dtype(t2);
    size_t z1 = dtype_size(t1), z2 = dtype_size(t2);
    if (s1 == s2)
        return t1;
    if (s1 && z1 > z2)
        return t1;
    if (!s1 && z1 > z2)
        return t1;
    switch (std::max(z1, z2)) {
    case 1: return AXB_I16;
    case 2: return AXB_I32;
    case 4: return AXB_I64;
    default: return AXB_F64;
    }
}

// cpu_operations.cpp:36-49 — comparisons/logical -> Bool, everything else -> promote
int32_t axb_binary_result_dtype(int32_t op, int32_t a, int32_t b) {
    switch (op) {
    case AXB_OP_EQ: case AXB_OP_NE: case AXB_OP_LT: case AXB_OP_LE: case AXB_OP_GT: case AXB_OP_GE:
    case AXB_OP_LOGICAL_AND: case AXB_OP_LOGICAL_OR: case AXB_OP_LOGICAL_XOR:
        return AXB_BOOL;
    default: return axb_promote_types(a, b);
    }
}
// cpu_operations.cpp:836-853
int32_t axb_unary_result_dtype(int32_t op, int32_t a) {
    switch (op) {
    case AXB_OP_LOGICAL_NOT: case AXB_OP_ISNAN: case AXB_OP_ISINF: case AXB_OP_ISFINITE: return AXB_BOOL;
    default: return a;
    }
}

axb_status axb_broadcast_shape(const int64_t *a, int na, const int64_t *b, int nb, int64_t *out, int *nout) {
    AXB_REQUIRE(out && nout && na >= 0 && nb >= 0 && na <= AXB_MAX_DIMS && nb <= AXB_MAX_DIMS, AXB_ERR_INVALID,
                "axb_broadcast_shape: bad arguments");
    int n = std::max(na, nb);
    for (int i = 0; i < n; ++i) {
        int ia = i - (n - na), ib = i - (n - nb);
        int64_t ea = ia >= 0 ? a[ia] : 1, eb = ib >= 0 ? b[ib] : 1;
        if (ea == eb || eb == 1)
            out[i] = ea;
        else if (ea == 1)
            out[i] = eb;
        else
            return fail(AXB_ERR_SHAPE, "shapes are not broadcastable: dim %d is %lld vs %lld", i, (long long)ea,
                        (long long)eb);
    }
    *nout = n;
    return AXB_OK;
}

axb_status axb_unary(int32_t op, const axb_tensor *in, const axb_tensor *out, double param, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_unary(in)"));
    AXB_TRY(validate_tensor(out, "axb_unary(out)"));
    axb_instr pr{};
    pr.op = op;
    pr.dtype = axb_unary_result_dtype(op, in->dtype);
    pr.src[0] = 0;
    pr.src[1] = pr.src[2] = AXB_SRC_NONE;
    pr.imm = param;
    AXB_REQUIRE(pr.dtype == out->dtype, AXB_ERR_INVALID, "axb_unary: out dtype %d, expected %d", out->dtype, pr.dtype);
    return axb_elementwise(&pr, 1, in, 1, out, stream);
}

axb_status axb_binary(int32_t op, const axb_tensor *a, const axb_tensor *b, const axb_tensor *out,
                      axb_stream_t stream) {
    AXB_TRY(validate_tensor(a, "axb_binary(a)"));
    AXB_TRY(validate_tensor(b, "axb_binary(b)"));
    AXB_TRY(validate_tensor(out, "axb_binary(out)"));
    axb_instr pr{};
    pr.op = op;
    pr.dtype = axb_binary_result_dtype(op, a->dtype, b->dtype);
    AXB_REQUIRE(pr.dtype >= 0, AXB_ERR_UNSUPPORTED, "axb_binary: cannot promote dtypes %d and %d", a->dtype, b->dtype);
    pr.src[0] = 0;
    pr.src[1] = 1;
    pr.src[2] = AXB_SRC_NONE;
    AXB_REQUIRE(pr.dtype == out->dtype, AXB_ERR_INVALID, "axb_binary: out dtype %d, expected %d", out->dtype, pr.dtype);
    axb_tensor ins[2] = {*a, *b};
    return axb_elementwise(&pr, 1, ins, 2, out, stream);
}

axb_status axb_where(const axb_tensor *cond, const axb_tensor *a, const axb_tensor *b, const axb_tensor *out,
                     axb_stream_t stream) {
    AXB_TRY(validate_tensor(cond, "axb_where(cond)"));
    AXB_TRY(validate_tensor(a, "axb_where(a)"));
    AXB_TRY(validate_tensor(b, "axb_where(b)"));
    AXB_TRY(validate_tensor(out, "axb_where(out)"));
    axb_instr pr{};
    pr.op = AXB_OP_WHERE;
    pr.dtype = axb_promote_types(a->dtype, b->dtype); // cpu_operations.cpp:2577-2589
    pr.src[0] = 0;
    pr.src[1] = 1;
    pr.src[2] = 2;
    AXB_REQUIRE(pr.dtype == out->dtype, AXB_ERR_INVALID, "axb_where: out dtype %d, expected %d", out->dtype, pr.dtype);
    axb_tensor ins[3] = {*cond, *a, *b};
    return axb_elementwise(&pr, 1, ins, 3, out, stream);
}

// result = mask ? (T)value : input   (cpu_operations.cpp:2596-2700; value converted to the input dtype)
axb_status axb_masked_fill(const axb_tensor *in, const axb_tensor *mask, double value, const axb_tensor *out,
                           axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_masked_fill(in)"));
    AXB_TRY(validate_tensor(mask, "axb_masked_fill(mask)"));
    AXB_TRY(validate_tensor(out, "axb_masked_fill(out)"));
    axb_instr pr{};
    pr.op = AXB_OP_WHERE;
    pr.dtype = in->dtype;
    pr.src[0] = 1;           // mask
    pr.src[1] = AXB_SRC_IMM; // fill value
    pr.src[2] = 0;           // input
    pr.imm = (double)(float)value; // the reference builds a Float32 {1} tensor first (operations.cpp:1585-1592)
    AXB_REQUIRE(out->dtype == in->dtype, AXB_ERR_INVALID, "axb_masked_fill: out dtype must equal input dtype");
    axb_tensor ins[2] = {*in, *mask};
    return axb_elementwise(&pr, 1, ins, 2, out, stream);
}

axb_status axb_cast(const axb_tensor *in, const axb_tensor *out, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_cast(in)"));
    AXB_TRY(validate_tensor(out, "axb_cast(out)"));
    axb_instr pr{};
    pr.op = AXB_OP_CAST;
    pr.dtype = out->dtype;
    pr.src[0] = 0;
    pr.src[1] = pr.src[2] = AXB_SRC_NONE;
    return axb_elementwise(&pr, 1, in, 1, out, stream);
}

axb_status axb_gather_strided(const axb_tensor *in, const axb_tensor *out, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_gather_strided(in)"));
    AXB_TRY(validate_tensor(out, "axb_gather_strided(out)"));
    AXB_REQUIRE(in->dtype == out->dtype, AXB_ERR_INVALID, "axb_gather_strided: dtype mismatch");
    if (is_complex_dtype(in->dtype)) {
        // complex is moved as raw 8/16-byte words: view c64 as u64, c128 as 2 x u64
        axb_tensor i2 = *in, o2 = *out;
        i2.dtype = o2.dtype = AXB_U64;
        if (in->dtype == AXB_C128) {
            AXB_REQUIRE(in->ndim < AXB_MAX_DIMS, AXB_ERR_UNSUPPORTED, "axb_gather_strided: complex128 rank too high");
            i2.shape[i2.ndim] = o2.shape[o2.ndim] = 2;
            i2.strides[i2.ndim] = o2.strides[o2.ndim] = 8;
            i2.ndim++;
            o2.ndim++;
        }
        return axb_cast(&i2, &o2, stream);
    }
    return axb_cast(in, out, stream);
}

axb_status axb_fill(const axb_tensor *out, double value, axb_stream_t stream) {
    AXB_TRY(validate_tensor(out, "axb_fill(out)"));
    axb_instr pr{};
    pr.op = AXB_OP_CAST;
    pr.dtype = out->dtype;
    pr.src[0] = AXB_SRC_IMM;
    pr.src[1] = pr.src[2] = AXB_SRC_NONE;
    pr.imm = value;
    return axb_elementwise(&pr, 1, nullptr, 0, out, stream);
}

axb_status axb_randn(const axb_tensor *out, uint64_t seed, float mean, float std_, axb_stream_t stream) {
    AXB_TRY(validate_tensor(out, "axb_randn(out)"));
    AXB_REQUIRE(out->dtype == AXB_F32 && is_contiguous(*out), AXB_ERR_UNSUPPORTED,
                "axb_randn: contiguous Float32 output required (cast afterwards)");
    int64_t n = numel(*out);
    if (n == 0)
        return AXB_OK;
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    int64_t blocks = std::min<int64_t>((n / 2 + 255) / 256 + 1, (int64_t)ctx->sm_count * 16);
    randn_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<float *>(tensor_ptr(*out)), n, seed, mean, std_);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

} // extern "C"
