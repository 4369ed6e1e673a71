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
TVM_HD float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
TVM_HD float f4_dot(float4 a, float4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
TVM_HD float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

// Axis tap for a coordinate that lies inside the field's box.  Every marched sample does (it passed the aabb
// test, so n = (p-lo)*inv-1 is in [-1, 1+1ulp] and idx in [0, size-1(+eps)]); the base index is clamped to size-2
// so both taps are always in range and no zero-padding masks are needed (at idx == size-1 this yields exactly
// the zero-padded value; beyond it the difference is O(ulp)).  Same unnormalise formula as ATen
// (GridSampler.h:31).  `scale` = d(idx)/d(coord).
struct AxisTap {
    int i0;
    float w0, w1;
};
TVM_HD float tvm_unnormalize(float coord, int size) { return ((coord + 1.0f) * 0.5f) * (float)(size - 1); }
TVM_HD AxisTap tvm_axis_tap_idx(float idx, int size) {
    AxisTap t;
    t.i0 = min(max((int)idx, 0), size - 2);
    t.w1 = idx - (float)t.i0;
    t.w0 = 1.0f - t.w1;
    return t;
}
TVM_HD AxisTap tvm_axis_tap_inbox(float coord, int size) { return tvm_axis_tap_idx(tvm_unnormalize(coord, size), size); }
struct SampleTaps {
    AxisTap a[3];        // x, y, z axis of the field grid
};
TVM_HD SampleTaps make_sample_taps(const tvm_field_desc& f, const float n[3]) {
    SampleTaps s;
#pragma unroll
    for (int c = 0; c < 3; ++c) s.a[c] = tvm_axis_tap_inbox(n[c], f.grid[c]);
    return s;
}
// same from the fractional texel indices idx[c] = tvm_unnormalize(n[c], grid[c]) (computed once by the lane that owns
// the sample and handed to the quad through shared memory)
TVM_HD SampleTaps make_sample_taps_idx(const tvm_field_desc& f, const float idx[3]) {
    SampleTaps s;
#pragma unroll
    for (int c = 0; c < 3; ++c) s.a[c] = tvm_axis_tap_idx(idx[c], f.grid[c]);
    return s;
}

// float4 offsets of the 12 factor sections inside the packed buffer (32-bit: one base pointer + one IMAD.WIDE.U32 per
// address instead of a 64-bit section pointer per plane / line).  Built on the host by the launchers.
struct TvmSections {
    unsigned dP[3], dL[3], aP[3], aL[3];
    unsigned dRow[3], aRow[3];          // plane row pitch in float4s (tvm_plane_pitch * channels / 4)
};
TVM_HD TvmSections tvm_sections(const tvm_field_desc& f) {
    TvmSections s;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s.dP[k] = (unsigned)(f.dplane_off[k] >> 2); s.dL[k] = (unsigned)(f.dline_off[k] >> 2);
        s.aP[k] = (unsigned)(f.aplane_off[k] >> 2); s.aL[k] = (unsigned)(f.aline_off[k] >> 2);
        s.dRow[k] = (unsigned)tvm_plane_pitch(f.grid[TVM_M0(k)]) * (unsigned)(f.n_sigma[k] >> 2);
        s.aRow[k] = (unsigned)tvm_plane_pitch(f.grid[TVM_M0(k)]) * (unsigned)(f.n_app[k] >> 2);
    }
    return s;
}
// keeps a loop-invariant per-lane value in a register (ptxas otherwise re-derives it from %tid inside the loop)
#if defined(__CUDA_ARCH__)
#define TVM_KEEP_REG(v) asm volatile("" : "+r"(v))
#else
#define TVM_KEEP_REG(v) ((void)0)
#endif

// texel offsets (in float4 units, before adding the channel slice j) and weights of plane/line pair k.
// Offsets are UNSIGNED 32-bit float4 indices so an address is one IMAD.WIDE.U32 off the section pointer
// (signed ints cost a sign-extension + LEA pair per load in SASS).
struct PlaneTaps {
    unsigned pbase, prow, lbase;  // plane: (y0*pitch+x0)*C4, pitch*C4 ; line: l0*C4
    float w00, w01, w10, w11;     // bilinear weights (ATen: nw, ne, sw, se)
    float lw0, lw1;
};
TVM_HD PlaneTaps make_taps(const tvm_field_desc& f, const SampleTaps& s, int k, int C4) {
    const AxisTap& tx = s.a[TVM_M0(k)];
    const AxisTap& ty = s.a[TVM_M1(k)];
    const AxisTap& tl = s.a[TVM_V(k)];
    const unsigned W = (unsigned)tvm_plane_pitch(f.grid[TVM_M0(k)]);     // row pitch in texels
    PlaneTaps p;
    p.pbase = ((unsigned)ty.i0 * W + (unsigned)tx.i0) * (unsigned)C4;
    p.prow = W * (unsigned)C4;
    p.lbase = (unsigned)tl.i0 * (unsigned)C4;
    p.w00 = tx.w0 * ty.w0; p.w01 = tx.w1 * ty.w0; p.w10 = tx.w0 * ty.w1; p.w11 = tx.w1 * ty.w1;
    p.lw0 = tl.w0; p.lw1 = tl.w1;
    return p;
}

// bilinear plane value and linear line value for one float4 channel slice of a texel with C4 float4s.
// r0 / r1 point at this lane's slice of the (x0,y0) and (x0,y0+1) texels, lb at the l0 line texel; `o` is a
// further slice offset (4*g float4s) that stays an immediate in the unrolled callers.
TVM_HD float4 vm_plane(const float4* __restrict__ r0, const float4* __restrict__ r1, const PlaneTaps& t, int C4, int o) {
    const float4 a = TVM_LDG4(r0 + o);
    const float4 b = TVM_LDG4(r0 + o + C4);
    const float4 c = TVM_LDG4(r1 + o);
    const float4 d = TVM_LDG4(r1 + o + C4);
    float4 pl = f4_scale(t.w00, a);
    pl = f4_fma(t.w01, b, pl); pl = f4_fma(t.w10, c, pl); pl = f4_fma(t.w11, d, pl);
    return pl;
}
TVM_HD float4 vm_line(const float4* __restrict__ lb, float lw0, float lw1, int C4, int o) {
    const float4 l0 = TVM_LDG4(lb + o);
    const float4 l1 = TVM_LDG4(lb + o + C4);
    return f4_fma(lw1, l1, f4_scale(lw0, l0));
}
// (plane (x) line) for slice j
TVM_HD float4 vm_product(const float4* __restrict__ P, const float4* __restrict__ Ln, const PlaneTaps& t, int C4,
                         int j) {
    const float4* r0 = P + (t.pbase + (unsigned)j);
    const float4* r1 = P + (t.pbase + t.prow + (unsigned)j);
    const float4* lb = Ln + (t.lbase + (unsigned)j);
    return f4_mul(vm_plane(r0, r1, t, C4, 0), vm_line(lb, t.lw0, t.lw1, C4, 0));
}

// this lane's share of sigma_feature = sum_k sum_c plane_k[c] * line_k[c]   (tensoRF.py:227-233)
// CS4 / CA4 template arguments: float4s per texel when all three planes have the same channel count (the
// lego/truck configs: 16 -> 4, 48 -> 12), which turns the corner / channel-slice offsets into immediates;
// 0 = read the per-plane counts from the descriptor.
template <int CS4 = 0>
TVM_HD float density_partial_taps(const tvm_field_desc& f, const TvmSections& sec, const SampleTaps& st, int sub) {
    float tot = 0.f;
    const float4* F4 = reinterpret_cast<const float4*>(f.factors);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = CS4 > 0 ? CS4 : (f.n_sigma[k] >> 2);
        if (sub < C4) {
            const PlaneTaps t = make_taps(f, st, k, C4);
            const unsigned po = sec.dP[k] + (unsigned)sub, lo = sec.dL[k] + (unsigned)sub;
            const float4* r0 = F4 + (t.pbase + po);
            const float4* r1 = F4 + (t.pbase + t.prow + po);
            const float4* lb = F4 + (t.lbase + lo);
            const float4 v = f4_mul(vm_plane(r0, r1, t, C4, 0), vm_line(lb, t.lw0, t.lw1, C4, 0));
            const float s = (v.x + v.y) + (v.z + v.w);
            tot = (k == 0) ? s : tot + s;
        }
    }
    return tot;
}
template <int CS4 = 0>
TVM_HD float density_partial(const tvm_field_desc& f, const float n[3], int sub) {
    return density_partial_taps<CS4>(f, tvm_sections(f), make_sample_taps(f, n), sub);
}

// A[k][g] += w * (app_plane_k (x) app_line_k)[channels of this lane]   (tensoRF.py:237-254, weighted by :888).
// The sample weight is folded into the two line-tap weights, so a slice costs plane(16) + line(8) + 4 FFMA.
template <int G, int CA4 = 0>
TVM_HD void app_accumulate_taps(const tvm_field_desc& f, const TvmSections& sec, const SampleTaps& st, float w, int sub,
                                float4 (&A)[3][G]) {
    const float4* F4 = reinterpret_cast<const float4*>(f.factors);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = CA4 > 0 ? CA4 : (f.n_app[k] >> 2);
        const PlaneTaps t = make_taps(f, st, k, C4);
        const float wl0 = w * t.lw0, wl1 = w * t.lw1;
        const unsigned po = sec.aP[k] + (unsigned)sub, lo = sec.aL[k] + (unsigned)sub;
        const float4* r0 = F4 + (t.pbase + po);
        const float4* r1 = F4 + (t.pbase + t.prow + po);
        const float4* lb = F4 + (t.lbase + lo);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int j = sub + 4 * g;
            if (j < C4) {
                const float4 pl = vm_plane(r0, r1, t, C4, 4 * g);
                const float4 ln = vm_line(lb, wl0, wl1, C4, 4 * g);
                A[k][g].x = fmaf(pl.x, ln.x, A[k][g].x); A[k][g].y = fmaf(pl.y, ln.y, A[k][g].y);
                A[k][g].z = fmaf(pl.z, ln.z, A[k][g].z); A[k][g].w = fmaf(pl.w, ln.w, A[k][g].w);
            }
        }
    }
}
template <int G, int CA4 = 0>
TVM_HD void app_accumulate(const tvm_field_desc& f, const float n[3], float w, int sub, float4 (&A)[3][G]) {
    app_accumulate_taps<G, CA4>(f, tvm_sections(f), make_sample_taps(f, n), w, sub, A);
}

// ------------------------------------------------------------------------------------------------
// Appearance accumulation over a RUN of consecutive samples with a register texel cache (app_gather_kernel, the
// second stage of the split forward march).
//
// Consecutive samples of a ray are half a voxel apart (step_ratio 0.5): on the bench image a sample needs 1.1 new plane
// texels (of 4) and 0.28 new line taps (of 2) when the texels of its predecessor are still at hand.  A quad walks a
// contiguous run of the ray's appearance samples one plane at a time.  Lane `sub` owns the float4 channel slices
// sub + 4g and keeps, for the current plane, the 4 corner texels and the 2 line taps of the current cell in registers,
// filed by the PARITY of their index (even/odd column x even/odd row; even/odd tap) and tagged with their float4
// offset: a step to the neighbouring column, row or tap in either direction re-fetches exactly the texels that
// changed, with no register moves, and a sample that stays in the cell fetches nothing.  The loads are predicated per
// quad; the L1 data stage serves 8 lanes per wavefront, so a wavefront is saved when both quads of a lane octet skip.
//
// sw[t] = (frac_x, frac_y, frac_z, weight), si[t] = i0_x | i0_y << 10 | i0_z << 20 per compacted sample
// (tvm_axis_tap_idx's clamped base index and idx - i0; grids up to TVM_PACKED_GRID_MAX per axis); the run is [begin, end).
// ------------------------------------------------------------------------------------------------
constexpr int TVM_PACKED_GRID_MAX = 1024;
TVM_HD int tvm_unpack_i0(unsigned v, int a) { return (int)((v >> (10 * a)) & 1023u); }
TVM_HD float tvm_pick(const float4& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
TVM_HD void tvm_slot_from_idx(const tvm_field_desc& f, const float idx[3], float w, float4& sw, unsigned& si) {
    const AxisTap tx = tvm_axis_tap_idx(idx[0], f.grid[0]);
    const AxisTap ty = tvm_axis_tap_idx(idx[1], f.grid[1]);
    const AxisTap tz = tvm_axis_tap_idx(idx[2], f.grid[2]);
    sw = make_float4(tx.w1, ty.w1, tz.w1, w);
    si = (unsigned)tx.i0 | ((unsigned)ty.i0 << 10) | ((unsigned)tz.i0 << 20);
}

// plane/line pair k over the run [begin, end): A[g] += sum_t w_t * (plane (x) line)[channels sub + 4g]
// COUNT: also count the 16-byte fetches this lane issues (measurement builds: bench.py's fetched-bytes roofline)
template <int G, int CA4 = 0, bool COUNT = false>
TVM_HD void app_run_plane(const tvm_field_desc& f, const TvmSections& sec, int k, const float4* __restrict__ sw,
                          const unsigned* __restrict__ si, int begin, int end, int sub, float4 (&A)[G],
                          unsigned* n_fetch = nullptr) {
    const float4* F4 = reinterpret_cast<const float4*>(f.factors);
    const int C4 = CA4 > 0 ? CA4 : (f.n_app[k] >> 2);
    const unsigned prow = sec.aRow[k];
    unsigned po = sec.aP[k] + (unsigned)sub, lo = sec.aL[k] + (unsigned)sub;         // this lane's slice
    TVM_KEEP_REG(po);
    TVM_KEEP_REG(lo);
    // cache tags: float4 offset each register set was loaded from (texels [x parity][y parity], taps [parity])
    unsigned tEE = ~0u, tOE = ~0u, tEO = ~0u, tOO = ~0u, tLE = ~0u, tLO = ~0u;
    float4 TEE[G], TOE[G], TEO[G], TOO[G], LE[G], LO[G];
#pragma unroll
    for (int g = 0; g < G; ++g) TEE[g] = TOE[g] = TEO[g] = TOO[g] = LE[g] = LO[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int t = begin; t < end; ++t) {
        const float4 w4 = sw[t];
        const unsigned i3 = si[t];
        const int x0 = tvm_unpack_i0(i3, TVM_M0(k)), y0 = tvm_unpack_i0(i3, TVM_M1(k)), l0 = tvm_unpack_i0(i3, TVM_V(k));
        const float fx = tvm_pick(w4, TVM_M0(k)), fy = tvm_pick(w4, TVM_M1(k)), fl = tvm_pick(w4, TVM_V(k));
        // the even and the odd member of {i0, i0+1} per axis and their interpolation weights
        const unsigned xE = (unsigned)(x0 + 1) & ~1u, xO = (unsigned)x0 | 1u;
        const unsigned yE = (unsigned)(y0 + 1) & ~1u, yO = (unsigned)y0 | 1u;
        const unsigned lE = (unsigned)(l0 + 1) & ~1u, lO = (unsigned)l0 | 1u;
        const bool xev = (x0 & 1) == 0, yev = (y0 & 1) == 0, lev = (l0 & 1) == 0;
        const float wxE = xev ? 1.0f - fx : fx, wxO = xev ? fx : 1.0f - fx;
        const float wyE = yev ? 1.0f - fy : fy, wyO = yev ? fy : 1.0f - fy;
        const float wlE = lev ? 1.0f - fl : fl, wlO = lev ? fl : 1.0f - fl;
        const unsigned rE = yE * prow + po, rO = yO * prow + po;
        const unsigned oEE = rE + xE * (unsigned)C4, oOE = rE + xO * (unsigned)C4;
        const unsigned oEO = rO + xE * (unsigned)C4, oOO = rO + xO * (unsigned)C4;
        const unsigned oLE = lE * (unsigned)C4 + lo, oLO = lO * (unsigned)C4 + lo;
        if (oEE != tEE) {
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (sub + 4 * g < C4) { TEE[g] = TVM_LDG4(F4 + oEE + 4 * g); if (COUNT) ++*n_fetch; }
            tEE = oEE;
        }
        if (oOE != tOE) {
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (sub + 4 * g < C4) { TOE[g] = TVM_LDG4(F4 + oOE + 4 * g); if (COUNT) ++*n_fetch; }
            tOE = oOE;
        }
        if (oEO != tEO) {
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (sub + 4 * g < C4) { TEO[g] = TVM_LDG4(F4 + oEO + 4 * g); if (COUNT) ++*n_fetch; }
            tEO = oEO;
        }
        if (oOO != tOO) {
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (sub + 4 * g < C4) { TOO[g] = TVM_LDG4(F4 + oOO + 4 * g); if (COUNT) ++*n_fetch; }
            tOO = oOO;
        }
        if (oLE != tLE) {
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (sub + 4 * g < C4) { LE[g] = TVM_LDG4(F4 + oLE + 4 * g); if (COUNT) ++*n_fetch; }
            tLE = oLE;
        }
        if (oLO != tLO) {
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (sub + 4 * g < C4) { LO[g] = TVM_LDG4(F4 + oLO + 4 * g); if (COUNT) ++*n_fetch; }
            tLO = oLO;
        }
        const float wEE = wxE * wyE, wOE = wxO * wyE, wEO = wxE * wyO, wOO = wxO * wyO;
        const float bE = w4.w * wlE, bO = w4.w * wlO;
#pragma unroll
        for (int g = 0; g < G; ++g)
            if (sub + 4 * g < C4) {
                float4 pl = f4_scale(wEE, TEE[g]);
                pl = f4_fma(wOE, TOE[g], pl); pl = f4_fma(wEO, TEO[g], pl); pl = f4_fma(wOO, TOO[g], pl);
                const float4 ln = f4_fma(bO, LO[g], f4_scale(bE, LE[g]));
                A[g].x = fmaf(pl.x, ln.x, A[g].x); A[g].y = fmaf(pl.y, ln.y, A[g].y);
                A[g].z = fmaf(pl.z, ln.z, A[g].z); A[g].w = fmaf(pl.w, ln.w, A[g].w);
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

// One float4 channel slice j of plane/line pair k: given the upstream gradient `up` on (plane (x) line)[channels],
// scatter into gP/gL and accumulate d/dn.  Returns (plane (x) line) for this slice.
template <bool SCATTER, bool POSE>
TVM_HD float4 vm_slice_bwd(const tvm_field_desc& f, const float4* __restrict__ F4, float4* __restrict__ G4,
                           unsigned poff, unsigned loff, const SampleTaps& s, const PlaneTaps& t,
                           int C4, int j, float4 up, int k, float dn[3]) {
    // one base pointer for the factors and one for the gradient buffer (same packed layout): every address is a
    // 32-bit float4 index off it, like the forward
    const unsigned ip = t.pbase + poff + (unsigned)j, ip1 = ip + t.prow, il = t.lbase + loff + (unsigned)j;
    const float4* pb = F4 + ip;
    const float4* pb1 = F4 + ip1;
    const float4* lb = F4 + il;
    const float4 a = TVM_LDG4(pb);
    const float4 b = TVM_LDG4(pb + C4);
    const float4 c = TVM_LDG4(pb1);
    const float4 d = TVM_LDG4(pb1 + C4);
    const float4 l0 = TVM_LDG4(lb);
    const float4 l1 = TVM_LDG4(lb + C4);
    float4 pl = f4_scale(t.w00, a);
    pl = f4_fma(t.w01, b, pl); pl = f4_fma(t.w10, c, pl); pl = f4_fma(t.w11, d, pl);
    float4 ln = f4_scale(t.lw0, l0);
    ln = f4_fma(t.lw1, l1, ln);
    const float4 up_ln = f4_mul(up, ln);        // d/d(plane value)
    const float4 up_pl = f4_mul(up, pl);        // d/d(line value)
    if (SCATTER) {
        float4* gpb = G4 + ip;
        float4* gpb1 = G4 + ip1;
        float4* glb = G4 + il;
        TVM_RED4(gpb, f4_scale(t.w00, up_ln));
        TVM_RED4(gpb + C4, f4_scale(t.w01, up_ln));
        TVM_RED4(gpb1, f4_scale(t.w10, up_ln));
        TVM_RED4(gpb1 + C4, f4_scale(t.w11, up_ln));
        TVM_RED4(glb, f4_scale(t.lw0, up_pl));
        TVM_RED4(glb + C4, f4_scale(t.lw1, up_pl));
    }
    if (POSE) {
        const AxisTap& tx = s.a[TVM_M0(k)];
        const AxisTap& ty = s.a[TVM_M1(k)];
        // d plane/d ix = wy0*(b-a) + wy1*(d-c);  d plane/d iy = wx0*(c-a) + wx1*(d-b);  d line/d il = l1-l0
        const float4 dx = f4_fma(ty.w1, f4_sub(d, c), f4_scale(ty.w0, f4_sub(b, a)));
        const float4 dy = f4_fma(tx.w1, f4_sub(d, b), f4_scale(tx.w0, f4_sub(c, a)));
        const float4 dl = f4_sub(l1, l0);
        dn[TVM_M0(k)] += 0.5f * (float)(f.grid[TVM_M0(k)] - 1) * f4_dot(up_ln, dx);
        dn[TVM_M1(k)] += 0.5f * (float)(f.grid[TVM_M1(k)] - 1) * f4_dot(up_ln, dy);
        dn[TVM_V(k)] += 0.5f * (float)(f.grid[TVM_V(k)] - 1) * f4_dot(up_pl, dl);
    }
    return f4_mul(pl, ln);
}

// ------------------------------------------------------------------------------------------------
// Backward scatter over a RUN of consecutive samples with register aggregation (march_bwd_kernel, training).
//
// The gradient of a sample lands on the 4 corner texels of its plane cell and the 2 taps of its line cell.  Consecutive
// samples of a ray mostly stay in the same cell, so a quad that walks a contiguous run of the block's samples keeps
// one pending gradient per (x parity, y parity) corner and per tap parity in registers, tagged with the texel offset,
// and issues the red.global.add.v4 only when the tag changes (or the run ends): equal addresses of neighbouring samples
// are added in registers instead of in the L2 reduction units.  CellTaps is the per-(sample, plane) record the owning
// lane prepares once in shared memory.
// ------------------------------------------------------------------------------------------------
struct alignas(16) CellTaps {
    unsigned oEE, oOE, oEO, oOO;        // float4 offsets (slice 0) of the corner texels: [x parity][y parity]
    float wEE, wOE, wEO, wOO;           // their bilinear weights
    unsigned oLE, oLO;                  // float4 offsets (slice 0) of the even / odd line tap
    float wLE, wLO;
};
TVM_HD CellTaps make_cell_taps(const SampleTaps& s, int k, unsigned prow, unsigned C4, unsigned pbase, unsigned lbase) {
    const AxisTap& tx = s.a[TVM_M0(k)];
    const AxisTap& ty = s.a[TVM_M1(k)];
    const AxisTap& tl = s.a[TVM_V(k)];
    const unsigned xE = (unsigned)(tx.i0 + 1) & ~1u, xO = (unsigned)tx.i0 | 1u;
    const unsigned yE = (unsigned)(ty.i0 + 1) & ~1u, yO = (unsigned)ty.i0 | 1u;
    const unsigned lE = (unsigned)(tl.i0 + 1) & ~1u, lO = (unsigned)tl.i0 | 1u;
    const bool xev = (tx.i0 & 1) == 0, yev = (ty.i0 & 1) == 0, lev = (tl.i0 & 1) == 0;
    const float wxE = xev ? tx.w0 : tx.w1, wxO = xev ? tx.w1 : tx.w0;
    const float wyE = yev ? ty.w0 : ty.w1, wyO = yev ? ty.w1 : ty.w0;
    CellTaps c;
    const unsigned rE = pbase + yE * prow, rO = pbase + yO * prow;
    c.oEE = rE + xE * C4; c.oOE = rE + xO * C4; c.oEO = rO + xE * C4; c.oOO = rO + xO * C4;
    c.wEE = wxE * wyE; c.wOE = wxO * wyE; c.wEO = wxE * wyO; c.wOO = wxO * wyO;
    c.oLE = lbase + lE * C4; c.oLO = lbase + lO * C4;
    c.wLE = lev ? tl.w0 : tl.w1; c.wLO = lev ? tl.w1 : tl.w0;
    return c;
}
TVM_HD float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
// pending gradient of one register slot: flush when the destination changes
TVM_HD void tvm_pend(float4* __restrict__ G4, unsigned& tag, float4& acc, unsigned o, float4 v) {
    if (o != tag) {
        if (tag != ~0u) TVM_RED4(G4 + tag, acc);
        acc = v;
        tag = o;
    } else {
        acc = f4_add(acc, v);
    }
}

// One float4 channel slice j of plane/line pair k over the run [begin, end) of the block's compacted samples:
// up[t] = wts[t] * g4 is the upstream gradient on (plane (x) line)[slice]; returns nothing, adds this lane's share of
// g4 . phi_t into part[t * 4 + sub] (appearance: c_i of the alpha gradient; ignored for density).
template <bool WANT_DOT>
TVM_HD void vm_run_bwd(const float4* __restrict__ F4, float4* __restrict__ G4, const CellTaps* __restrict__ cells, int k,
                       const float* __restrict__ wts, int begin, int end, unsigned j, float4 g4, float* __restrict__ part,
                       int sub) {
    unsigned tEE = ~0u, tOE = ~0u, tEO = ~0u, tOO = ~0u, tLE = ~0u, tLO = ~0u;
    float4 aEE = make_float4(0.f, 0.f, 0.f, 0.f), aOE = aEE, aEO = aEE, aOO = aEE, aLE = aEE, aLO = aEE;
#pragma unroll 1
    for (int t = begin; t < end; ++t) {
        const CellTaps c = cells[t * 3 + k];
        const float w = wts[t];
        const unsigned oEE = c.oEE + j, oOE = c.oOE + j, oEO = c.oEO + j, oOO = c.oOO + j, oLE = c.oLE + j, oLO = c.oLO + j;
        const float4 vEE = TVM_LDG4(F4 + oEE), vOE = TVM_LDG4(F4 + oOE), vEO = TVM_LDG4(F4 + oEO), vOO = TVM_LDG4(F4 + oOO);
        const float4 vLE = TVM_LDG4(F4 + oLE), vLO = TVM_LDG4(F4 + oLO);
        float4 pl = f4_scale(c.wEE, vEE);
        pl = f4_fma(c.wOE, vOE, pl); pl = f4_fma(c.wEO, vEO, pl); pl = f4_fma(c.wOO, vOO, pl);
        const float4 ln = f4_fma(c.wLO, vLO, f4_scale(c.wLE, vLE));
        if (WANT_DOT) part[t * 4 + sub] += f4_dot(g4, f4_mul(pl, ln));
        const float4 up = f4_scale(w, g4);
        const float4 up_ln = f4_mul(up, ln), up_pl = f4_mul(up, pl);
        tvm_pend(G4, tEE, aEE, oEE, f4_scale(c.wEE, up_ln));
        tvm_pend(G4, tOE, aOE, oOE, f4_scale(c.wOE, up_ln));
        tvm_pend(G4, tEO, aEO, oEO, f4_scale(c.wEO, up_ln));
        tvm_pend(G4, tOO, aOO, oOO, f4_scale(c.wOO, up_ln));
        tvm_pend(G4, tLE, aLE, oLE, f4_scale(c.wLE, up_pl));
        tvm_pend(G4, tLO, aLO, oLO, f4_scale(c.wLO, up_pl));
    }
    if (tEE != ~0u) TVM_RED4(G4 + tEE, aEE);
    if (tOE != ~0u) TVM_RED4(G4 + tOE, aOE);
    if (tEO != ~0u) TVM_RED4(G4 + tEO, aEO);
    if (tOO != ~0u) TVM_RED4(G4 + tOO, aOO);
    if (tLE != ~0u) TVM_RED4(G4 + tLE, aLE);
    if (tLO != ~0u) TVM_RED4(G4 + tLO, aLO);
}

// density: upstream dfeat (scalar, same for every channel).  Returns this lane's share of sigma_feature.
template <bool SCATTER, bool POSE, int CS4 = 0>
TVM_HD float density_bwd(const tvm_field_desc& f, const float n[3], float dfeat, int sub, float* gbuf, float dn[3]) {
    float tot = 0.f;
    const float4 up = make_float4(dfeat, dfeat, dfeat, dfeat);
    const SampleTaps st = make_sample_taps(f, n);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = CS4 > 0 ? CS4 : (f.n_sigma[k] >> 2);
        if (sub < C4) {
            const PlaneTaps t = make_taps(f, st, k, C4);
            const float4 v = vm_slice_bwd<SCATTER, POSE>(
                f, reinterpret_cast<const float4*>(f.factors), reinterpret_cast<float4*>(gbuf),
                (unsigned)(f.dplane_off[k] >> 2), (unsigned)(f.dline_off[k] >> 2), st, t, C4, sub, up, k, dn);
            tot += (v.x + v.y) + (v.z + v.w);
        }
    }
    return tot;
}

// appearance, compact-code form: the upstream slices are re-read from the ray's d_ray_feat row (L1-resident, one
// broadcast wavefront per slice) instead of living in 9 float4 registers, and the channel-group loop is not unrolled —
// a third of the instructions of app_bwd for the same arithmetic.  gRow = d_ray_feat + r * ta (16-byte aligned).
template <int G, bool SCATTER, bool POSE, int CA4 = 0>
TVM_HD float app_bwd_rolled(const tvm_field_desc& f, const float n[3], float w, int sub, const float* __restrict__ gRow,
                            const int (&app_off)[3], float* gbuf, float dn[3]) {
    float dot = 0.f;
    const SampleTaps st = make_sample_taps(f, n);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = CA4 > 0 ? CA4 : (f.n_app[k] >> 2);
        const PlaneTaps t = make_taps(f, st, k, C4);
        const float4* gk = reinterpret_cast<const float4*>(gRow + app_off[k]);
#pragma unroll 1
        for (int j = sub; j < C4; j += 4) {
            const float4 g4 = TVM_LDG4(gk + j);
            const float4 phi = vm_slice_bwd<SCATTER, POSE>(
                f, reinterpret_cast<const float4*>(f.factors), reinterpret_cast<float4*>(gbuf),
                (unsigned)(f.aplane_off[k] >> 2), (unsigned)(f.aline_off[k] >> 2), st, t, C4, j, f4_scale(w, g4), k, dn);
            dot += f4_dot(g4, phi);
        }
    }
    return dot;
}

// appearance: upstream on (plane (x) line)[c] is w * gF[c].  Returns this lane's share of gF . phi.
template <int G, bool SCATTER, bool POSE, int CA4 = 0>
TVM_HD float app_bwd(const tvm_field_desc& f, const float n[3], float w, int sub, const float4 (&gF)[3][G],
                     float* gbuf, float dn[3]) {
    float dot = 0.f;
    const SampleTaps st = make_sample_taps(f, n);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = CA4 > 0 ? CA4 : (f.n_app[k] >> 2);
        const PlaneTaps t = make_taps(f, st, k, C4);
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            const int j = sub + 4 * gi;
            if (j < C4) {
                const float4 phi = vm_slice_bwd<SCATTER, POSE>(
                    f, reinterpret_cast<const float4*>(f.factors), reinterpret_cast<float4*>(gbuf),
                    (unsigned)(f.aplane_off[k] >> 2), (unsigned)(f.aline_off[k] >> 2), st, t, C4, j,
                    f4_scale(w, gF[k][gi]), k, dn);
                dot += f4_dot(gF[k][gi], phi);
            }
        }
    }
    return dot;
}
