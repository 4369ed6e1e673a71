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


// ------------------------------------------------------------------------------------------------
// backward: gradient scatter into the packed factor-gradient buffer (same layout as `factors`) and,
// for pose refinement, the gradient w.r.t. the normalised sample coordinate.
// Autograd equivalent in the reference: grid_sampler_2d_backward for the 12 grid_sample calls of
// compute_densityfeature / compute_appfeature (driven by train.py:338, inerf/estimate_pose_inerf.py:178).
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define TVM_RED4(p, v) atomicAdd((p), (v))          // red.global.add.v4.f32 (sm_90+)
#else
static inline void tvm_host_add4(float4* p, float4 v) { p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w; }
#define TVM_RED4(p, v) tvm_host_add4((p), (v))
#endif

TVM_HD float f4_dot(float4 a, float4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
TVM_HD float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

struct PlaneTapsG {
    PlaneTaps t;
    TvmTap tx, ty, tl;
};
TVM_HD PlaneTapsG make_taps_g(const tvm_field_desc& f, const float n[3], int k) {
    const int W = f.grid[TVM_M0(k)], H = f.grid[TVM_M1(k)], L = f.grid[TVM_V(k)];
    PlaneTapsG g;
    g.tx = tvm_axis_tap(n[TVM_M0(k)], W);
    g.ty = tvm_axis_tap(n[TVM_M1(k)], H);
    g.tl = tvm_axis_tap(n[TVM_V(k)], L);
    PlaneTaps& p = g.t;
    p.t00 = g.ty.i0 * W + g.tx.i0; p.t01 = g.ty.i0 * W + g.tx.i1;
    p.t10 = g.ty.i1 * W + g.tx.i0; p.t11 = g.ty.i1 * W + g.tx.i1;
    p.w00 = g.tx.w0 * g.ty.w0; p.w01 = g.tx.w1 * g.ty.w0; p.w10 = g.tx.w0 * g.ty.w1; p.w11 = g.tx.w1 * g.ty.w1;
    p.l0 = g.tl.i0; p.l1 = g.tl.i1; p.lw0 = g.tl.w0; p.lw1 = g.tl.w1;
    return g;
}

// One float4 channel slice j of plane/line pair k: given the upstream gradient `up` on (plane (x) line)[channels],
// scatter into gP/gL and accumulate d/dn.  Returns (plane (x) line) for this slice.
template <bool SCATTER, bool POSE>
TVM_HD float4 vm_slice_bwd(const float4* __restrict__ P, const float4* __restrict__ Ln, float4* __restrict__ gP,
                           float4* __restrict__ gL, const PlaneTapsG& g, int C4, int j, float4 up, int k,
                           float dn[3]) {
    const PlaneTaps& t = g.t;
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
    if (SCATTER) {
        const float4 up_ln = f4_mul(up, ln);        // d/d(plane value)
        const float4 up_pl = f4_mul(up, pl);        // d/d(line value)
        if (t.w00 != 0.f) TVM_RED4(gP + t.t00 * C4 + j, f4_scale(t.w00, up_ln));
        if (t.w01 != 0.f) TVM_RED4(gP + t.t01 * C4 + j, f4_scale(t.w01, up_ln));
        if (t.w10 != 0.f) TVM_RED4(gP + t.t10 * C4 + j, f4_scale(t.w10, up_ln));
        if (t.w11 != 0.f) TVM_RED4(gP + t.t11 * C4 + j, f4_scale(t.w11, up_ln));
        if (t.lw0 != 0.f) TVM_RED4(gL + t.l0 * C4 + j, f4_scale(t.lw0, up_pl));
        if (t.lw1 != 0.f) TVM_RED4(gL + t.l1 * C4 + j, f4_scale(t.lw1, up_pl));
    }
    if (POSE) {
        // d plane / d ix = wy0*(m1x*b - m0x*a) + wy1*(m1x*d - m0x*c); zero-padded taps contribute 0
        float4 dx = f4_scale(g.ty.w0, make_float4(g.tx.m1 * b.x - g.tx.m0 * a.x, g.tx.m1 * b.y - g.tx.m0 * a.y,
                                                  g.tx.m1 * b.z - g.tx.m0 * a.z, g.tx.m1 * b.w - g.tx.m0 * a.w));
        dx = f4_fma(g.ty.w1, make_float4(g.tx.m1 * d.x - g.tx.m0 * c.x, g.tx.m1 * d.y - g.tx.m0 * c.y,
                                         g.tx.m1 * d.z - g.tx.m0 * c.z, g.tx.m1 * d.w - g.tx.m0 * c.w), dx);
        float4 dy = f4_scale(g.tx.w0, make_float4(g.ty.m1 * c.x - g.ty.m0 * a.x, g.ty.m1 * c.y - g.ty.m0 * a.y,
                                                  g.ty.m1 * c.z - g.ty.m0 * a.z, g.ty.m1 * c.w - g.ty.m0 * a.w));
        dy = f4_fma(g.tx.w1, make_float4(g.ty.m1 * d.x - g.ty.m0 * b.x, g.ty.m1 * d.y - g.ty.m0 * b.y,
                                         g.ty.m1 * d.z - g.ty.m0 * b.z, g.ty.m1 * d.w - g.ty.m0 * b.w), dy);
        const float4 dl = make_float4(g.tl.m1 * l1.x - g.tl.m0 * l0.x, g.tl.m1 * l1.y - g.tl.m0 * l0.y,
                                      g.tl.m1 * l1.z - g.tl.m0 * l0.z, g.tl.m1 * l1.w - g.tl.m0 * l0.w);
        dn[TVM_M0(k)] += g.tx.scale * f4_dot(f4_mul(up, ln), dx);
        dn[TVM_M1(k)] += g.ty.scale * f4_dot(f4_mul(up, ln), dy);
        dn[TVM_V(k)] += g.tl.scale * f4_dot(f4_mul(up, pl), dl);
    }
    return f4_mul(pl, ln);
}

// density: upstream dfeat (scalar, same for every channel).  Returns this lane's share of sigma_feature.
template <bool SCATTER, bool POSE>
TVM_HD float density_bwd(const tvm_field_desc& f, const float n[3], float dfeat, int sub, float* gbuf, float dn[3]) {
    float tot = 0.f;
    const float4 up = make_float4(dfeat, dfeat, dfeat, dfeat);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = f.n_sigma[k] >> 2;
        if (sub < C4) {
            const PlaneTapsG g = make_taps_g(f, n, k);
            const float4 v = vm_slice_bwd<SCATTER, POSE>(
                reinterpret_cast<const float4*>(f.factors + f.dplane_off[k]),
                reinterpret_cast<const float4*>(f.factors + f.dline_off[k]),
                SCATTER ? reinterpret_cast<float4*>(gbuf + f.dplane_off[k]) : nullptr,
                SCATTER ? reinterpret_cast<float4*>(gbuf + f.dline_off[k]) : nullptr, g, C4, sub, up, k, dn);
            tot += (v.x + v.y) + (v.z + v.w);
        }
    }
    return tot;
}

// appearance: upstream on (plane (x) line)[c] is w * gF[c].  Returns this lane's share of gF . phi.
template <int G, bool SCATTER, bool POSE>
TVM_HD float app_bwd(const tvm_field_desc& f, const float n[3], float w, int sub, const float4 (&gF)[3][G],
                     float* gbuf, float dn[3]) {
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = f.n_app[k] >> 2;
        const PlaneTapsG g = make_taps_g(f, n, k);
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            const int j = sub + 4 * gi;
            if (j < C4) {
                const float4 phi = vm_slice_bwd<SCATTER, POSE>(
                    reinterpret_cast<const float4*>(f.factors + f.aplane_off[k]),
                    reinterpret_cast<const float4*>(f.factors + f.aline_off[k]),
                    SCATTER ? reinterpret_cast<float4*>(gbuf + f.aplane_off[k]) : nullptr,
                    SCATTER ? reinterpret_cast<float4*>(gbuf + f.aline_off[k]) : nullptr, g, C4, j,
                    f4_scale(w, gF[k][gi]), k, dn);
                dot += f4_dot(gF[k][gi], phi);
            }
        }
    }
    return dot;
}
