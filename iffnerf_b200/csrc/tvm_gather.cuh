// tvm_gather.cuh — plane/line gathers of the VM decomposition, shared by the forward and backward
// march kernels (and compiled for the host by tests/hostcheck to validate layout + tap math on CPU).
//
// compute_densityfeature / compute_appfeature (models/tensoRF.py:216-256): plane k is sampled
// bilinearly at (p[m0], p[m1]) and line k linearly at p[v] (F.grid_sample, align_corners=True, zero
// padding; the line is a width-1 image sampled at x=0, i.e. pure linear interpolation along H).
// Factors are channel-last, so `j` selects one float4 (4 channels) of a texel with C4 float4s.
#pragma once
#include "tvm_math.cuh"

#if defined(__CUDA_ARCH__)
#define TVM_LDG4(p) __ldg(p)
#else
#define TVM_LDG4(p) (*(p))
#endif
#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v = {x, y, z, w}; return v; }
#endif

TVM_HD float4 f4_fma(float s, float4 a, float4 acc) {
    acc.x = fmaf(s, a.x, acc.x); acc.y = fmaf(s, a.y, acc.y); acc.z = fmaf(s, a.z, acc.z); acc.w = fmaf(s, a.w, acc.w);
    return acc;
}
TVM_HD float4 f4_scale(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }

struct PlaneTaps {
    int t00, t01, t10, t11;      // texel indices (row-major y*W+x)
    float w00, w01, w10, w11;    // bilinear weights (ATen: nw, ne, sw, se)
    int l0, l1;                  // line taps
    float lw0, lw1;
};

TVM_HD PlaneTaps make_taps(const tvm_field_desc& f, const float n[3], int k) {
    const int W = f.grid[TVM_M0(k)], H = f.grid[TVM_M1(k)], L = f.grid[TVM_V(k)];
    const TvmTap tx = tvm_axis_tap(n[TVM_M0(k)], W);
    const TvmTap ty = tvm_axis_tap(n[TVM_M1(k)], H);
    const TvmTap tl = tvm_axis_tap(n[TVM_V(k)], L);
    PlaneTaps p;
    p.t00 = ty.i0 * W + tx.i0; p.t01 = ty.i0 * W + tx.i1;
    p.t10 = ty.i1 * W + tx.i0; p.t11 = ty.i1 * W + tx.i1;
    p.w00 = tx.w0 * ty.w0; p.w01 = tx.w1 * ty.w0; p.w10 = tx.w0 * ty.w1; p.w11 = tx.w1 * ty.w1;
    p.l0 = tl.i0; p.l1 = tl.i1; p.lw0 = tl.w0; p.lw1 = tl.w1;
    return p;
}

// (plane (x) line) for one float4 channel slice j of a texel with C4 float4s
TVM_HD float4 vm_product(const float4* __restrict__ P, const float4* __restrict__ Ln,
                                             const PlaneTaps& t, int C4, int j) {
    const float4 a = TVM_LDG4(P + t.t00 * C4 + j);
    const float4 b = TVM_LDG4(P + t.t01 * C4 + j);
    const float4 c = TVM_LDG4(P + t.t10 * C4 + j);
    const float4 d = TVM_LDG4(P + t.t11 * C4 + j);
    const float4 l0 = TVM_LDG4(Ln + t.l0 * C4 + j);
    const float4 l1 = TVM_LDG4(Ln + t.l1 * C4 + j);
    float4 pl = f4_scale(t.w00, a);
    pl = f4_fma(t.w01, b, pl); pl = f4_fma(t.w10, c, pl); pl = f4_fma(t.w11, d, pl);
    float4 ln = f4_scale(t.lw0, l0);
    ln = f4_fma(t.lw1, l1, ln);
    return make_float4(pl.x * ln.x, pl.y * ln.y, pl.z * ln.z, pl.w * ln.w);
}

// this lane's share of sigma_feature = sum_k sum_c plane_k[c] * line_k[c]   (tensoRF.py:227-233)
TVM_HD float density_partial(const tvm_field_desc& f, const float n[3], int sub) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = f.n_sigma[k] >> 2;
        if (sub < C4) {
            const PlaneTaps t = make_taps(f, n, k);
            const float4 v = vm_product(reinterpret_cast<const float4*>(f.factors + f.dplane_off[k]),
                                        reinterpret_cast<const float4*>(f.factors + f.dline_off[k]), t, C4, sub);
            tot += (v.x + v.y) + (v.z + v.w);
        }
    }
    return tot;
}

// A[k][g] += w * (app_plane_k (x) app_line_k)[channels of this lane]   (tensoRF.py:237-254, weighted by :888)
template <int G>
TVM_HD void app_accumulate(const tvm_field_desc& f, const float n[3], float w, int sub,
                                               float4 (&A)[3][G]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = f.n_app[k] >> 2;
        const PlaneTaps t = make_taps(f, n, k);
        const float4* P = reinterpret_cast<const float4*>(f.factors + f.aplane_off[k]);
        const float4* Ln = reinterpret_cast<const float4*>(f.factors + f.aline_off[k]);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int j = sub + 4 * g;
            if (j < C4) A[k][g] = f4_fma(w, vm_product(P, Ln, t, C4, j), A[k][g]);
        }
    }
}

