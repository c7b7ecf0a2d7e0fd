// spatial_parity — conv2d, conv_transpose1d/2d and interpolate on Device::GPU against the SAME call on Device::CPU of the
// unmodified reference library (the reference is its own oracle). Sweeps strides, paddings,
// dilations, output paddings, groups, bias, unbatched inputs, strided (transposed-view) inputs, the three interpolate
// modes with and without align_corners, and the float dtypes. fp32 must be BIT-IDENTICAL (the kernels keep the CPU loops'
// accumulation type and term order); half types are compared through fp32 with one output ulp of slack.
// Prints one JSON line {"cases": n, "failed": m, "bit_exact_f32": k, "failures": [...]}.
#include <axiom/axiom.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace axiom;

struct Tally {
    int cases = 0, failed = 0, exact_f32 = 0, f32_cases = 0;
    std::vector<std::string> failures;
};

static void compare(Tally &t, const std::string &name, const Tensor &cpu, const Tensor &gpu) {
    ++t.cases;
    bool ok = gpu.device() == Device::GPU && cpu.shape() == gpu.shape() && cpu.dtype() == gpu.dtype();
    if (ok) {
        Tensor g = gpu.cpu(), c = cpu.is_contiguous() ? cpu : cpu.ascontiguousarray();
        if (cpu.dtype() == DType::Float32 || cpu.dtype() == DType::Float64) {
            const bool same = std::memcmp(c.data(), g.data(), c.nbytes()) == 0;
            if (cpu.dtype() == DType::Float32) {
                ++t.f32_cases;
                t.exact_f32 += same ? 1 : 0;
            }
            ok = same || c.allclose(g, 1e-6, 1e-6);
            if (cpu.dtype() == DType::Float32 && !same)
                ok = false; // the contract for fp32 is bit-exact
        } else {
            ok = c.astype(DType::Float32).allclose(g.astype(DType::Float32), 1.0 / 128, 1e-3);
        }
    }
    if (!ok) {
        ++t.failed;
        if (t.failures.size() < 12)
            t.failures.push_back(name);
    }
}

int main() {
    ops::OperationRegistry::initialize_builtin_operations();
    manual_seed(77);
    Tally t;
    const DType dts[] = {DType::Float32, DType::Float64, DType::BFloat16, DType::Float16};
    const char *dtn[] = {"f32", "f64", "bf16", "f16"};

    // ---- conv2d ----------------------------------------------------------------------------------------------------
    struct C2 { size_t n, cin, h, w, cout, kh, kw; int sh, sw, ph, pw, dh, dw, groups; bool bias, batch; };
    const C2 c2[] = {{2, 3, 9, 11, 4, 3, 3, 1, 1, 0, 0, 1, 1, 1, true, true},   {1, 4, 12, 10, 6, 3, 2, 2, 1, 1, 2, 1, 1, 2, true, true},
                     {2, 2, 14, 14, 2, 3, 3, 1, 2, 2, 1, 2, 2, 1, false, true}, {1, 6, 8, 8, 6, 1, 1, 1, 1, 0, 0, 1, 1, 6, true, true},
                     {1, 3, 7, 9, 5, 5, 3, 1, 1, 2, 1, 1, 1, 1, true, false},   {3, 8, 16, 16, 16, 3, 3, 2, 2, 1, 1, 1, 1, 4, false, true}};
    for (size_t di = 0; di < 4; ++di)
        for (const C2 &c : c2) {
            Tensor x = Tensor::randn(c.batch ? Shape{c.n, c.cin, c.h, c.w} : Shape{c.cin, c.h, c.w}).astype(dts[di]);
            Tensor w = (Tensor::randn({c.cout, c.cin / (size_t)c.groups, c.kh, c.kw}) * 0.3f).astype(dts[di]);
            Tensor b = c.bias ? Tensor::randn({c.cout}).astype(dts[di]) : Tensor();
            Tensor rc = ops::conv2d(x, w, b, {c.sh, c.sw}, {c.ph, c.pw}, {c.dh, c.dw}, c.groups);
            Tensor rg = ops::conv2d(x.gpu(), w.gpu(), c.bias ? b.gpu() : Tensor(), {c.sh, c.sw}, {c.ph, c.pw}, {c.dh, c.dw}, c.groups);
            rg.sync();
            compare(t, std::string("conv2d_") + dtn[di] + "_" + std::to_string(&c - c2), rc, rg);
        }
    { // a strided input view (W and H swapped) and a strided weight view
        Tensor x = Tensor::randn({2, 3, 10, 12}), w = Tensor::randn({4, 3, 3, 3});
        Tensor xv = x.transpose({0, 1, 3, 2}), wv = w.transpose({0, 1, 3, 2});
        Tensor rc = ops::conv2d(xv, wv, Tensor(), {1, 1}, {1, 1});
        Tensor rg = ops::conv2d(x.gpu().transpose({0, 1, 3, 2}), w.gpu().transpose({0, 1, 3, 2}), Tensor(), {1, 1}, {1, 1});
        rg.sync();
        compare(t, "conv2d_f32_views", rc, rg);
    }

    // ---- conv_transpose1d ------------------------------------------------------------------------------------------------
    struct T1 { size_t n, cin, l, coutg, k; int s, p, op, d, groups; bool bias, batch; };
    const T1 t1[] = {{1, 1, 3, 1, 2, 1, 0, 0, 1, 1, false, true}, {2, 4, 9, 3, 3, 2, 1, 1, 1, 1, true, true}, {2, 4, 7, 2, 4, 3, 2, 2, 2, 2, true, true},
                     {1, 6, 5, 1, 3, 1, 0, 0, 1, 6, false, true}, {1, 3, 8, 2, 5, 2, 0, 1, 1, 1, true, false}};
    for (size_t di = 0; di < 4; ++di)
        for (const T1 &c : t1) {
            Tensor x = Tensor::randn(c.batch ? Shape{c.n, c.cin, c.l} : Shape{c.cin, c.l}).astype(dts[di]);
            Tensor w = (Tensor::randn({c.cin, c.coutg, c.k}) * 0.5f).astype(dts[di]);
            Tensor b = c.bias ? Tensor::randn({c.coutg * (size_t)c.groups}).astype(dts[di]) : Tensor();
            Tensor rc = ops::conv_transpose1d(x, w, b, c.s, c.p, c.op, c.d, c.groups);
            Tensor rg = ops::conv_transpose1d(x.gpu(), w.gpu(), c.bias ? b.gpu() : Tensor(), c.s, c.p, c.op, c.d, c.groups);
            compare(t, std::string("conv_transpose1d_") + dtn[di] + "_" + std::to_string(&c - t1), rc, rg);
        }

    // ---- conv_transpose2d ------------------------------------------------------------------------------------------------
    struct T2 { size_t n, cin, h, w, coutg, kh, kw; int sh, sw, ph, pw, oph, opw, dh, dw, groups; bool bias; };
    const T2 t2[] = {{1, 2, 4, 4, 3, 3, 3, 1, 1, 0, 0, 0, 0, 1, 1, 1, false}, {2, 4, 5, 6, 2, 3, 2, 2, 3, 1, 0, 1, 2, 1, 1, 2, true},
                     {1, 3, 6, 5, 4, 2, 4, 2, 2, 1, 1, 0, 1, 2, 1, 1, true},  {2, 6, 3, 3, 1, 3, 3, 1, 1, 1, 1, 0, 0, 1, 1, 6, false}};
    for (size_t di = 0; di < 4; ++di)
        for (const T2 &c : t2) {
            Tensor x = Tensor::randn({c.n, c.cin, c.h, c.w}).astype(dts[di]);
            Tensor w = (Tensor::randn({c.cin, c.coutg, c.kh, c.kw}) * 0.5f).astype(dts[di]);
            Tensor b = c.bias ? Tensor::randn({c.coutg * (size_t)c.groups}).astype(dts[di]) : Tensor();
            Tensor rc = ops::conv_transpose2d(x, w, b, {c.sh, c.sw}, {c.ph, c.pw}, {c.oph, c.opw}, {c.dh, c.dw}, c.groups);
            Tensor rg = ops::conv_transpose2d(x.gpu(), w.gpu(), c.bias ? b.gpu() : Tensor(), {c.sh, c.sw}, {c.ph, c.pw}, {c.oph, c.opw},
                                              {c.dh, c.dw}, c.groups);
            compare(t, std::string("conv_transpose2d_") + dtn[di] + "_" + std::to_string(&c - t2), rc, rg);
        }

    // ---- interpolate -----------------------------------------------------------------------------------------------------
    const ops::InterpolateMode modes[] = {ops::InterpolateMode::Nearest, ops::InterpolateMode::Bilinear, ops::InterpolateMode::Bicubic};
    const char *mn[] = {"nearest", "bilinear", "bicubic"};
    struct I2 { size_t n, c, h, w, oh, ow; };
    const I2 i2[] = {{1, 3, 4, 4, 8, 8}, {2, 2, 5, 7, 9, 4}, {1, 1, 6, 6, 3, 3}, {2, 3, 8, 5, 8, 16}, {1, 2, 1, 5, 4, 10}};
    for (size_t di = 0; di < 4; ++di)
        for (int m = 0; m < 3; ++m)
            for (int ac = 0; ac < 2; ++ac) {
                for (const I2 &c : i2) {
                    Tensor x = Tensor::randn({c.n, c.c, c.h, c.w}).astype(dts[di]);
                    Tensor rc = ops::interpolate(x, {c.oh, c.ow}, {}, modes[m], ac != 0);
                    Tensor rg = ops::interpolate(x.gpu(), {c.oh, c.ow}, {}, modes[m], ac != 0);
                    compare(t, std::string("interp2d_") + dtn[di] + "_" + mn[m] + (ac ? "_ac_" : "_") + std::to_string(&c - i2), rc, rg);
                }
                if (m < 2) { // 1-D: nearest / linear
                    Tensor x = Tensor::randn({2, 3, 7}).astype(dts[di]);
                    for (size_t ol : {size_t(14), size_t(5), size_t(7)}) {
                        Tensor rc = ops::interpolate(x, {ol}, {}, modes[m], ac != 0);
                        Tensor rg = ops::interpolate(x.gpu(), {ol}, {}, modes[m], ac != 0);
                        compare(t, std::string("interp1d_") + dtn[di] + "_" + mn[m] + (ac ? "_ac_" : "_") + std::to_string(ol), rc, rg);
                    }
                }
            }
    { // scale_factor form
        Tensor x = Tensor::randn({1, 2, 6, 6});
        Tensor rc = ops::interpolate(x, {}, {1.5f, 2.0f}, ops::InterpolateMode::Bilinear);
        Tensor rg = ops::interpolate(x.gpu(), {}, {1.5f, 2.0f}, ops::InterpolateMode::Bilinear);
        compare(t, "interp2d_f32_scale_factor", rc, rg);
    }

    std::string fl;
    for (auto &f : t.failures)
        fl += (fl.empty() ? "\"" : ", \"") + f + "\"";
    std::printf("{\"cases\": %d, \"failed\": %d, \"f32_cases\": %d, \"bit_exact_f32\": %d, \"failures\": [%s]}\n", t.cases, t.failed, t.f32_cases,
                t.exact_f32, fl.c_str());
    return t.failed ? 1 : 0;
}
