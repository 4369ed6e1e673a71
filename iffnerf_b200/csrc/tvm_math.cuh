// tvm_math.cuh — per-sample math shared by every kernel of the TensoRF-VM render path.
//
// Everything here is __host__ __device__ so tests/hostcheck can compile the *same* source
// with g++ (-ffp-contract=off) and compare the bit-exact parts against the oracle without a GPU.
// The bit-exact contract: sample positions and the occupancy test
// use the reference's un-fused fp32 op order, so every op that feeds `ray_valid` goes through
// the explicitly rounded rn_* wrappers (no FMA contraction).
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/tvm_b200.h"

#if defined(__CUDACC__)
#define TVM_HD __host__ __device__ __forceinline__
#else
#define TVM_HD inline
#include <algorithm>
using std::max;
using std::min;
#endif

#if defined(__CUDA_ARCH__)
TVM_HD float rn_add(float a, float b) { return __fadd_rn(a, b); }
TVM_HD float rn_sub(float a, float b) { return __fsub_rn(a, b); }
TVM_HD float rn_mul(float a, float b) { return __fmul_rn(a, b); }
TVM_HD float rn_div(float a, float b) { return __fdiv_rn(a, b); }
#else
// host build: compiled with -ffp-contract=off, volatile blocks re-association across calls
TVM_HD float rn_add(float a, float b) { volatile float r = a + b; return r; }
TVM_HD float rn_sub(float a, float b) { volatile float r = a - b; return r; }
TVM_HD float rn_mul(float a, float b) { volatile float r = a * b; return r; }
TVM_HD float rn_div(float a, float b) { volatile float r = a / b; return r; }
#endif

// Row pitch (in texels) of a packed plane whose rows hold W texels: W rounded up to odd.  Texels are 64 B (density)
// or 192 B (appearance), so an odd pitch puts vertically adjacent texels into different halves of the 128-B L1 line
// span (bank groups); with an even pitch (300) the two quads of an 8-lane wavefront group that fetch samples one row
// apart collide on the same banks and the request takes an extra wavefront (measured 4.68 instead of 4 per LDG.128).
TVM_HD int tvm_plane_pitch(int W) { return W | 1; }

// matMode / vecMode of the VM decomposition (models/tensorBase.py:311-312)
#define TVM_M0(k) ((k) == 2 ? 1 : 0)
#define TVM_M1(k) ((k) == 0 ? 1 : 2)
#define TVM_V(k)  (2 - (k))

struct TvmRay {
    float o[3];
    float d[3];
    float t0;      // clamped slab-entry distance (sample 0)
    float jit;     // per-ray jitter (0 in eval)
    int i_off;     // sample-index offset: 0 for sample_ray, -(S/2) for sample_point_color
};

// t_min of sample_ray (models/tensorBase.py:499-502):
//   vec = d==0 ? 1e-6 : d ; t = clamp(max_c(min((hi-o)/vec,(lo-o)/vec)), near, far)
TVM_HD float tvm_ray_entry(const tvm_field_desc& f, const float o[3], const float d[3]) {
    float t = -INFINITY;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = (d[c] == 0.0f) ? 1e-6f : d[c];
        float ra = rn_div(rn_sub(f.aabb[3 + c], o[c]), v);
        float rb = rn_div(rn_sub(f.aabb[c], o[c]), v);
        float m = fminf(ra, rb);
        t = fmaxf(t, m);
    }
    t = fmaxf(t, f.near_t);   // clamp(min=near, max=far): max first, then min (ATen clamp order)
    t = fminf(t, f.far_t);
    return t;
}

// Sampler set-up.  sample_ray (models/tensorBase.py:494-536): t0 = clamped slab entry, indices 0..S-1 (+ jitter).
// sample_point_color (:623-638, TVM_F_POINT_SAMPLES): S samples centred on the origin, z_i = stepSize*(i - S/2),
// i.e. t0 = 0 (x + 0 is exact) and an index offset of -(S/2); no near/far clamp, no jitter.
TVM_HD void tvm_init_ray(const tvm_field_desc& f, TvmRay& r, float jitter, int S, bool point_samples) {
    r.t0 = point_samples ? 0.0f : tvm_ray_entry(f, r.o, r.d);
    r.jit = point_samples ? 0.0f : jitter;
    r.i_off = point_samples ? -(S / 2) : 0;
}

// z_i = t0 + stepSize * (float(i) + jitter)      (models/tensorBase.py:504-529)
TVM_HD float tvm_sample_z(const tvm_field_desc& f, const TvmRay& r, int i) {
    float rng = rn_add((float)(i + r.i_off), r.jit);
    return rn_add(r.t0, rn_mul(f.step_size, rng));
}

// xyz = o + d*z (separate roundings, :531); returns the in-aabb test (:532-536)
TVM_HD bool tvm_sample_point(const tvm_field_desc& f, const TvmRay& r, float z, float p[3]) {
    bool inside = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        p[c] = rn_add(r.o[c], rn_mul(r.d[c], z));
        inside = inside && !(f.aabb[c] > p[c]) && !(p[c] > f.aabb[3 + c]);
    }
    return inside;
}

// alphaMask.sample_alpha(xyz) > 0  (models/tensorBase.py:66-83, :832-834).
// ATen trilinear, align_corners=True, zero padding: value>0  <=>  some in-bounds corner with
// strictly positive weight product is occupied.  `cells[z][y][x]` bit (dz*4+dy*2+dx)
// = volume[z+dz][y+dy][x+dx] > 0 (0 when out of bounds), so the common case is ONE byte load.
TVM_HD bool tvm_occupancy_keep(const tvm_field_desc& f, const float p[3]) {
    float idx[3];
    int i0[3];
    bool frac[3];
    bool interior = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float n = rn_sub(rn_mul(rn_sub(p[c], f.occ_lo[c]), f.occ_inv[c]), 1.0f);        // :83
        idx[c] = rn_mul(rn_mul(rn_add(n, 1.0f), 0.5f), (float)(f.occ_dims[c] - 1));     // GridSampler.h:31
        float fl = floorf(idx[c]);
        frac[c] = idx[c] != fl;                   // weight of the "+1" corner = idx - floor > 0
        // clamp before the int conversion so wild coordinates cannot overflow
        fl = fminf(fmaxf(fl, -2.0f), (float)f.occ_dims[c] + 1.0f);
        i0[c] = (int)fl;
        interior = interior && i0[c] >= 0 && i0[c] < f.occ_dims[c];
    }
    const int Dx = f.occ_dims[0], Dy = f.occ_dims[1];
    if (interior) {
        unsigned code = f.occ_cells[((size_t)i0[2] * Dy + i0[1]) * Dx + i0[0]];
        unsigned allow = (frac[0] ? 0xFFu : 0x55u) & (frac[1] ? 0xFFu : 0x33u) & (frac[2] ? 0xFFu : 0x0Fu);
        return (code & allow) != 0u;
    }
    // rare: base corner outside the lattice (mask aabb smaller than the field aabb) — test corners one by one
    bool keep = false;
    for (int b = 0; b < 8; ++b) {
        int dx = b & 1, dy = (b >> 1) & 1, dz = b >> 2;
        if ((dx && !frac[0]) || (dy && !frac[1]) || (dz && !frac[2])) continue;
        int x = i0[0] + dx, y = i0[1] + dy, z = i0[2] + dz;
        if (x < 0 || y < 0 || z < 0 || x >= Dx || y >= Dy || z >= f.occ_dims[2]) continue;
        keep = keep || (f.occ_cells[((size_t)z * Dy + y) * Dx + x] & 1u);
    }
    return keep;
}

// Conservative block test used to skip empty space without touching bit-exactness: can ANY of the samples
// i in [i_first, i_last] be valid?  z_i is monotone in i and every op of o + d*z and of the occupancy
// unnormalise is monotone in its input, so per axis all sample coordinates / cell indices of the block lie
// between those of its two end samples.  The block is discarded only if that box misses the aabb, or if every
// 16^3 super-cell it touches has no occupied corner (occ_coarse, built by tvm_pack_occupancy).
TVM_HD bool tvm_block_may_be_valid(const tvm_field_desc& f, const TvmRay& r, int i_first, int i_last) {
    float pa[3], pb[3];
    tvm_sample_point(f, r, tvm_sample_z(f, r, i_first), pa);
    tvm_sample_point(f, r, tvm_sample_z(f, r, i_last), pb);
    int c0[3], c1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float lo = fminf(pa[c], pb[c]), hi = fmaxf(pa[c], pb[c]);
        if (hi < f.aabb[c] || lo > f.aabb[3 + c]) return false;            // whole block outside the box
        if (f.occ_cells != nullptr) {
            const float na = rn_sub(rn_mul(rn_sub(lo, f.occ_lo[c]), f.occ_inv[c]), 1.0f);
            const float nb = rn_sub(rn_mul(rn_sub(hi, f.occ_lo[c]), f.occ_inv[c]), 1.0f);
            const float ia = rn_mul(rn_mul(rn_add(na, 1.0f), 0.5f), (float)(f.occ_dims[c] - 1));
            const float ib = rn_mul(rn_mul(rn_add(nb, 1.0f), 0.5f), (float)(f.occ_dims[c] - 1));
            const float dmax = (float)(f.occ_dims[c] - 1);
            // base cells outside [0, D-1] only ever look at the boundary cells (slow path of tvm_occupancy_keep)
            c0[c] = (int)fminf(fmaxf(floorf(fminf(ia, ib)), 0.0f), dmax) >> 4;
            c1[c] = (int)fminf(fmaxf(floorf(fmaxf(ia, ib)), 0.0f), dmax) >> 4;
        }
    }
    if (f.occ_cells == nullptr || f.occ_coarse == nullptr) return true;
    const int cx = f.occ_cdims[0], cy = f.occ_cdims[1];
    for (int z = c0[2]; z <= c1[2]; ++z)
        for (int y = c0[1]; y <= c1[1]; ++y)
            for (int x = c0[0]; x <= c1[0]; ++x)
                if (f.occ_coarse[((size_t)z * cy + y) * cx + x]) return true;
    return false;
}

// normalize_coord (models/tensorBase.py:397): (xyz - aabb0) * invaabbSize - 1
TVM_HD void tvm_normalize(const tvm_field_desc& f, const float p[3], float n[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) n[c] = rn_sub(rn_mul(rn_sub(p[c], f.aabb[c]), f.inv_aabb[c]), 1.0f);
}

// One axis of an align_corners=True, zero-padded (bi)linear tap: index pair + weights.
// Out-of-range taps get weight 0 and a clamped (safe) index.  m0/m1 are the in-bounds masks and
// `scale` = d(idx)/d(coord) = (size-1)/2, both only used by the backward (coordinate gradients).
struct TvmTap {
    int i0, i1;
    float w0, w1;
    float m0, m1;
    float scale;
};
TVM_HD TvmTap tvm_axis_tap(float coord, int size) {
    float idx = ((coord + 1.0f) * 0.5f) * (float)(size - 1);
    float fl = floorf(idx);
    float fr = idx - fl;
    fl = fminf(fmaxf(fl, -2.0f), (float)size + 1.0f);
    int a = (int)fl, b = a + 1;
    TvmTap t;
    t.m0 = (a >= 0 && a < size) ? 1.0f : 0.0f;
    t.m1 = (b >= 0 && b < size) ? 1.0f : 0.0f;
    t.w0 = t.m0 * (1.0f - fr);
    t.w1 = t.m1 * fr;
    t.i0 = min(max(a, 0), size - 1);
    t.i1 = min(max(b, 0), size - 1);
    t.scale = 0.5f * (float)(size - 1);
    return t;
}

// feature2density (models/tensorBase.py:750-754); F.softplus beta=1 threshold=20
TVM_HD float tvm_density(const tvm_field_desc& f, float feat) {
    if (f.act == 1) return fmaxf(feat, 0.0f);
    float x = feat + f.density_shift;
    return x > 20.0f ? x : log1pf(expf(x));
}
// d sigma / d feat
TVM_HD float tvm_density_grad(const tvm_field_desc& f, float feat) {
    if (f.act == 1) return feat > 0.0f ? 1.0f : 0.0f;
    float x = feat + f.density_shift;
    return x > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-x));
}
