// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" window onto the UNMODIFIED reference (Frikallo/axiom, CPU backend) so that the
// Python test-suite / bench baseline can drive the reference's own public API through ctypes.
// It is compiled together with the reference's sources (compiled in place from /root/reference by
// oracle/Makefile) into oracle/_ref/libaxiom_ref.so. Nothing in axiom_b200/ links against it.
//
// Every entry point maps 1:1 onto a reference public function (cited per function); no arithmetic
// lives here. Handles are heap-allocated axiom::Tensor objects.
#include <axiom/axiom.hpp>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace axiom;

namespace {
thread_local std::string g_err;
struct Handle {
    Tensor t;
};
inline Handle *H(void *p) { return static_cast<Handle *>(p); }
template <typename F> void *guard(F &&f) {
    try {
        return new Handle{f()};
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    } catch (...) {
        g_err = "unknown exception";
        return nullptr;
    }
}
} // namespace

extern "C" {

int axref_version(void) { return 3; }
const char *axref_last_error(void) { return g_err.c_str(); }

int axref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void axref_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
    (void)n;
}

// include/axiom/graph/graph_registry.hpp: EagerModeScope / set_eager_mode
void axref_init(void) { ops::OperationRegistry::initialize_builtin_operations(); }

// Tensor view ctor: src/tensor/tensor.cpp:120-155 (storage, shape, byte strides, dtype, byte offset)
void *axref_from_buffer(const void *data, uint64_t storage_bytes, int ndim, const int64_t *shape,
                        const int64_t *strides_bytes, int64_t offset_bytes, int dtype) {
    return guard([&] {
        std::shared_ptr<Storage> st(make_storage(storage_bytes, Device::CPU).release());
        if (storage_bytes)
            std::memcpy(st->data(), data, storage_bytes);
        Shape sh(shape, shape + ndim);
        Strides sd(strides_bytes, strides_bytes + ndim);
        return Tensor(st, sh, sd, static_cast<DType>(dtype), static_cast<size_t>(offset_bytes));
    });
}

void axref_free(void *h) { delete H(h); }

int axref_info(void *h, int *ndim, int64_t *shape, int *dtype, int *is_lazy) {
    try {
        const Tensor &t = H(h)->t;
        *ndim = static_cast<int>(t.ndim());
        for (size_t i = 0; i < t.ndim(); ++i)
            shape[i] = static_cast<int64_t>(t.shape()[i]);
        *dtype = static_cast<int>(t.dtype());
        *is_lazy = t.is_lazy() ? 1 : 0;
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// Forces materialisation (tensor.cpp:255-273) — the reference's own benchmarks force with .data().
int axref_sync(void *h) {
    try {
        H(h)->t.sync();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// Materialise, make C-contiguous (tensor.cpp:503-594) and copy the bytes out.
int64_t axref_read(void *h, void *out, uint64_t capacity) {
    try {
        Tensor c = H(h)->t.ascontiguousarray();
        uint64_t n = c.nbytes();
        if (n > capacity) {
            g_err = "axref_read: buffer too small";
            return -1;
        }
        if (n)
            std::memcpy(out, c.data(), n);
        return static_cast<int64_t>(n);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// Unary ops by OpType value (include/axiom/operations.hpp:55-208) -> ops::<fn> (operations.hpp:343-383)
void *axref_unary(int op, void *a, float fparam) {
    return guard([&]() -> Tensor {
        const Tensor &x = H(a)->t;
        switch (static_cast<ops::OpType>(op)) {
        case ops::OpType::Negate: return ops::negate(x);
        case ops::OpType::Abs: return ops::abs(x);
        case ops::OpType::Sqrt: return ops::sqrt(x);
        case ops::OpType::Exp: return ops::exp(x);
        case ops::OpType::Log: return ops::log(x);
        case ops::OpType::Sin: return ops::sin(x);
        case ops::OpType::Cos: return ops::cos(x);
        case ops::OpType::Tan: return ops::tan(x);
        case ops::OpType::Erf: return ops::erf(x);
        case ops::OpType::Sign: return ops::sign(x);
        case ops::OpType::Floor: return ops::floor(x);
        case ops::OpType::Ceil: return ops::ceil(x);
        case ops::OpType::Trunc: return ops::trunc(x);
        case ops::OpType::Round: return ops::round(x);
        case ops::OpType::Reciprocal: return ops::reciprocal(x);
        case ops::OpType::Square: return ops::square(x);
        case ops::OpType::Cbrt: return ops::cbrt(x);
        case ops::OpType::IsNaN: return ops::isnan(x);
        case ops::OpType::IsInf: return ops::isinf(x);
        case ops::OpType::IsFinite: return ops::isfinite(x);
        case ops::OpType::LogicalNot: return ops::logical_not(x);
        case ops::OpType::ReLU: return ops::relu(x);
        case ops::OpType::LeakyReLU: return ops::leaky_relu(x, fparam);
        case ops::OpType::SiLU: return ops::silu(x);
        case ops::OpType::Sigmoid: return ops::sigmoid(x);
        case ops::OpType::Tanh: return ops::tanh(x);
        case ops::OpType::GELU: return ops::gelu(x);
        default: throw ValueError("axref_unary: unsupported op");
        }
    });
}

void *axref_binary(int op, void *a, void *b) {
    return guard([&]() -> Tensor {
        const Tensor &x = H(a)->t;
        const Tensor &y = H(b)->t;
        switch (static_cast<ops::OpType>(op)) {
        case ops::OpType::Add: return ops::add(x, y);
        case ops::OpType::Subtract: return ops::subtract(x, y);
        case ops::OpType::Multiply: return ops::multiply(x, y);
        case ops::OpType::Divide: return ops::divide(x, y);
        case ops::OpType::Power: return ops::power(x, y);
        case ops::OpType::Modulo: return ops::modulo(x, y);
        case ops::OpType::Equal: return ops::equal(x, y);
        case ops::OpType::NotEqual: return ops::not_equal(x, y);
        case ops::OpType::Less: return ops::less(x, y);
        case ops::OpType::LessEqual: return ops::less_equal(x, y);
        case ops::OpType::Greater: return ops::greater(x, y);
        case ops::OpType::GreaterEqual: return ops::greater_equal(x, y);
        case ops::OpType::LogicalAnd: return ops::logical_and(x, y);
        case ops::OpType::LogicalOr: return ops::logical_or(x, y);
        case ops::OpType::LogicalXor: return ops::logical_xor(x, y);
        case ops::OpType::BitwiseAnd: return ops::bitwise_and(x, y);
        case ops::OpType::BitwiseOr: return ops::bitwise_or(x, y);
        case ops::OpType::BitwiseXor: return ops::bitwise_xor(x, y);
        case ops::OpType::LeftShift: return ops::left_shift(x, y);
        case ops::OpType::RightShift: return ops::right_shift(x, y);
        case ops::OpType::Maximum: return ops::maximum(x, y);
        case ops::OpType::Minimum: return ops::minimum(x, y);
        case ops::OpType::Atan2: return ops::atan2(x, y);
        case ops::OpType::Hypot: return ops::hypot(x, y);
        default: throw ValueError("axref_binary: unsupported op");
        }
    });
}

// ops::where operations.cpp:1548-1570
void *axref_where(void *c, void *a, void *b) {
    return guard([&] { return ops::where(H(c)->t, H(a)->t, H(b)->t); });
}
// Tensor::where(cond, scalar) tensor.cpp:1336-1355 — the exact call config 2 makes
void *axref_where_scalar(void *self, void *cond, float other) {
    return guard([&] { return H(self)->t.where(H(cond)->t, other); });
}
void *axref_masked_fill(void *x, void *mask, float value) {
    return guard([&] { return ops::masked_fill(H(x)->t, H(mask)->t, value); });
}
// scalar forms of the operators (tensor_operators.hpp:104-107,195-210): Tensor op float / Tensor op int
void *axref_binary_scalar_f32(int op, void *a, float s) {
    return guard([&]() -> Tensor {
        const Tensor &x = H(a)->t;
        switch (static_cast<ops::OpType>(op)) {
        case ops::OpType::Add: return x + s;
        case ops::OpType::Subtract: return x - s;
        case ops::OpType::Multiply: return x * s;
        case ops::OpType::Divide: return x / s;
        case ops::OpType::Greater: return x > s;
        case ops::OpType::Less: return x < s;
        case ops::OpType::GreaterEqual: return x >= s;
        case ops::OpType::LessEqual: return x <= s;
        default: throw ValueError("axref_binary_scalar_f32: unsupported op");
        }
    });
}
void *axref_binary_scalar_i32(int op, void *a, int s) {
    return guard([&]() -> Tensor {
        const Tensor &x = H(a)->t;
        switch (static_cast<ops::OpType>(op)) {
        case ops::OpType::Add: return x + s;
        case ops::OpType::Subtract: return x - s;
        case ops::OpType::Multiply: return x * s;
        case ops::OpType::Greater: return x > s;
        case ops::OpType::Less: return x < s;
        default: throw ValueError("axref_binary_scalar_i32: unsupported op");
        }
    });
}

// Reductions: ops::sum/mean/max/min/prod/any/all operations.hpp:386-412
void *axref_reduce(int op, void *a, const int *axes, int n_axes, int keep_dims) {
    return guard([&]() -> Tensor {
        const Tensor &x = H(a)->t;
        std::vector<int> ax(axes, axes + n_axes);
        bool kd = keep_dims != 0;
        switch (static_cast<ops::OpType>(op)) {
        case ops::OpType::Sum: return ops::sum(x, ax, kd);
        case ops::OpType::Mean: return ops::mean(x, ax, kd);
        case ops::OpType::Max: return ops::max(x, ax, kd);
        case ops::OpType::Min: return ops::min(x, ax, kd);
        case ops::OpType::Prod: return ops::prod(x, ax, kd);
        case ops::OpType::Any: return ops::any(x, ax, kd);
        case ops::OpType::All: return ops::all(x, ax, kd);
        default: throw ValueError("axref_reduce: unsupported op");
        }
    });
}
void *axref_argreduce(int op, void *a, int axis, int keep_dims) {
    return guard([&]() -> Tensor {
        if (static_cast<ops::OpType>(op) == ops::OpType::ArgMax)
            return ops::argmax(H(a)->t, axis, keep_dims != 0);
        return ops::argmin(H(a)->t, axis, keep_dims != 0);
    });
}
void *axref_softmax(void *a, int axis, int is_log) {
    return guard([&] { return is_log ? ops::log_softmax(H(a)->t, axis) : ops::softmax(H(a)->t, axis); });
}
// ops::layer_norm / rms_norm operations.cpp:1775-1858
void *axref_layer_norm(void *x, void *w, void *b, int axis, float eps) {
    return guard([&] { return ops::layer_norm(H(x)->t, H(w)->t, H(b)->t, axis, eps); });
}
void *axref_rms_norm(void *x, void *w, int axis, float eps) {
    return guard([&] { return ops::rms_norm(H(x)->t, H(w)->t, axis, eps); });
}
// ops::matmul operations.cpp:1542
void *axref_matmul(void *a, void *b, int ta, int tb) {
    return guard([&] { return ops::matmul(H(a)->t, H(b)->t, ta != 0, tb != 0); });
}
// Tensor::astype tensor.cpp:1656-1701 ; ascontiguousarray :503-594 ; transpose :699-756 ; reshape :623-676
void *axref_astype(void *a, int dtype) {
    return guard([&] { return H(a)->t.astype(static_cast<DType>(dtype)); });
}
void *axref_contiguous(void *a) {
    return guard([&] { return H(a)->t.ascontiguousarray(); });
}
void *axref_transpose(void *a) {
    return guard([&] { return H(a)->t.T(); });
}
void *axref_permute(void *a, const int *axes, int n) {
    return guard([&] { return H(a)->t.transpose(std::vector<int>(axes, axes + n)); });
}
void *axref_reshape(void *a, const int64_t *shape, int n) {
    return guard([&] { return H(a)->t.reshape(Shape(shape, shape + n)); });
}
// README chain (README.md:18-22): x.matmul(x.T()).softmax(-1).sum({1})
void *axref_readme_chain(void *x) {
    return guard([&] {
        const Tensor &t = H(x)->t;
        return t.matmul(t.T()).softmax(-1).sum({1});
    });
}
// Config 2 chain exactly as a user writes it: ((a.relu()+1.0f)*b).where(c>0, 0)
void *axref_config2_chain(void *a, void *b, void *c) {
    return guard([&] {
        const Tensor &A = H(a)->t;
        const Tensor &B = H(b)->t;
        const Tensor &C = H(c)->t;
        return ((A.relu() + 1.0f) * B).where(C > 0, 0);
    });
}
// gather / index_select (operations.hpp:437-451)
void *axref_gather(void *x, int dim, void *idx) {
    return guard([&] { return ops::gather(H(x)->t, dim, H(idx)->t); });
}
void *axref_index_select(void *x, int dim, void *idx) {
    return guard([&] { return ops::index_select(H(x)->t, dim, H(idx)->t); });
}
// conv1d / sdpa for the config-5 "next" rows (operations.cpp:2968-3135, :3979-4044)
void *axref_conv1d(void *x, void *w, void *b, int stride, int padding, int dilation, int groups) {
    return guard([&] {
        Tensor bias = b ? H(b)->t : Tensor();
        return ops::conv1d(H(x)->t, H(w)->t, bias, stride, padding, dilation, groups);
    });
}
void *axref_sdpa(void *q, void *k, void *v, void *mask, float scale, int is_causal) {
    return guard([&] {
        Tensor m = mask ? H(mask)->t : Tensor();
        return ops::scaled_dot_product_attention(H(q)->t, H(k)->t, H(v)->t, m, scale, is_causal != 0);
    });
}

// pooling / sort / topk / masked_select (operations.cpp:2367-2880, :4199-4257, :1631-1656)
void *axref_pool(void *x, int is_avg, int nsp, const int *kernel, const int *stride, const int *padding, int include_pad) {
    return guard([&] {
        const Tensor &X = H(x)->t;
        std::vector<int> k(kernel, kernel + nsp), s(stride, stride + nsp), p(padding, padding + nsp);
        if (nsp == 1)
            return is_avg ? ops::avg_pool1d(X, k[0], s[0], p[0], include_pad != 0) : ops::max_pool1d(X, k[0], s[0], p[0]);
        if (nsp == 2)
            return is_avg ? ops::avg_pool2d(X, k, s, p, include_pad != 0) : ops::max_pool2d(X, k, s, p);
        return is_avg ? ops::avg_pool3d(X, k, s, p, include_pad != 0) : ops::max_pool3d(X, k, s, p);
    });
}
// what: 0 sort values, 1 argsort, 2 topk values, 3 topk indices (descending doubles as `largest`)
void *axref_sort(void *x, int what, int axis, int descending, int k) {
    return guard([&] {
        const Tensor &X = H(x)->t;
        if (what == 0)
            return ops::sort(X, axis, descending != 0);
        if (what == 1)
            return ops::argsort(X, axis, descending != 0);
        auto r = ops::topk(X, k, axis, descending != 0, true);
        return what == 2 ? r.first : r.second;
    });
}
void *axref_masked_select(void *x, void *mask) {
    return guard([&] { return ops::masked_select(H(x)->t, H(mask)->t); });
}

// Wall-clock helper for the CPU baseline: runs `iters` x (build lazy op + force) for a small menu of
// workloads and returns mean seconds (methodology = benchmarks/compare/elementwise_axiom.cpp:32-61).
double axref_time_sync(void *h) {
    auto t0 = std::chrono::steady_clock::now();
    H(h)->t.sync();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
