// march_bwd.cu — backward of the fused march kernel.
//
// Replaces the autograd backward the reference runs through grid_sampler_2d_backward, cumprod, softplus,
// index_put ... for TensorBase.forward (driven by train.py:338 and inerf/estimate_pose_inerf.py:178).
// Inputs are the gradients of the march-stage outputs: d(ray_feat) [n][sum n_app] (from the shade backward),
// d(acc) [n] and d(alpha) [n][S] (train.py:328 puts a loss on alpha).  Outputs: the packed factor-gradient
// buffer (same channel-last layout as `factors`) and, for pose refinement, d(rays) [n][6].
//
// Each warp re-marches its ray exactly like the forward (same sample positions, masks and compositing,
// no early termination) — nothing per-sample is stored by the forward.  With c_i = d_acc + [app_i] gF.phi_i:
//     dL/dalpha_i = g_alpha_i + T_i c_i - (sum_{k>i} w_k c_k) / (1 - alpha_i + 1e-10)
// and the suffix sum is Total - prefix, where Total = d_acc*acc + gF.ray_feat comes from the forward's
// workspace, so one front-to-back pass suffices.  Factor gradients are scattered with 16-byte vector
// reductions (red.global.add.v4.f32), one per (corner, float4 channel slice).
#include "tvm_common.cuh"
#include "tvm_gather.cuh"
#include "tvm_warp.cuh"

namespace {

#ifndef TVM_BWD_MIN_BLOCKS
#define TVM_BWD_MIN_BLOCKS 3
#endif
#ifndef TVM_BWD_ROLL_APP
#define TVM_BWD_ROLL_APP 1        // compact appearance pass (app_bwd_rolled); 0: fully unrolled app_bwd with gF in registers
#endif
#ifndef TVM_BWD_MIN_BLOCKS_POSE
#define TVM_BWD_MIN_BLOCKS_POSE 5   // pose-only instantiation (no scatter): 96 registers, 20 warps/SM (6 spills)
#endif
#ifndef TVM_BWD_WARPS
#define TVM_BWD_WARPS 4
#endif
constexpr int BWD_WARPS = TVM_BWD_WARPS;
constexpr int BWD_RAYS_PER_CTA = 16;
constexpr unsigned FULL = 0xffffffffu;

struct BwdArgs {
    tvm_field_desc f;
    const float* rays;
    long long n_rays;
    int ray_stride;
    int S;
    const float* jitter;
    unsigned flags;
    const float* d_ray_feat;   // [n][ta] or NULL
    const float* d_acc;        // [n] or NULL
    const float* d_alpha;      // [n][S] or NULL
    float* g_factors;          // packed layout or NULL
    float* g_rays;             // [n][6] or NULL
    const float* ray_feat;     // forward workspace
    const float* acc;
    int ta;
    int app_off[3];
    int rays_per_cta;
    TvmSections sec;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    return v;
}

template <int G, bool SCATTER, bool POSE, int CS4, int CA4>
__global__ void __launch_bounds__(BWD_WARPS * 32, SCATTER ? TVM_BWD_MIN_BLOCKS : TVM_BWD_MIN_BLOCKS_POSE)
march_bwd_kernel(const __grid_constant__ BwdArgs a) {
    __shared__ int s_next;
    __shared__ float4 s_slot[BWD_WARPS][32];
    __shared__ float s_ret[BWD_WARPS][32];
    __shared__ float s_z[BWD_WARPS][32];
    __shared__ float4 s_dn[BWD_WARPS][32];      // pose-only mode: d(sigma_feature)/d(normalised coords) per sample
    // training instantiation (scatter, no pose), TVM_F_BWD_RUNS: run-aggregated scatter (tvm_gather.cuh::vm_run_bwd) —
    // per-(sample, plane) tap records, the upstream scalar of each compacted sample, the per-lane shares of gF . phi.
    // Measured (65 536 rays): L2 reduction traffic -60 % (lts 56 % -> 21 %) but +71 % instructions; the kernel is
    // latency-bound at 12 warps/SM, not L2-bound, so it is slower (5.18 vs 3.74 ms) and stays opt-in.
    constexpr bool RUNS = SCATTER && !POSE;
    __shared__ CellTaps s_cell[RUNS ? BWD_WARPS : 1][RUNS ? 32 * 3 : 1];
    __shared__ float s_wt[RUNS ? BWD_WARPS : 1][32];
    __shared__ __align__(16) float s_part[RUNS ? BWD_WARPS : 1][RUNS ? 128 : 4];
    const bool runs = RUNS && (a.flags & TVM_F_BWD_RUNS);
    const float4* F4 = reinterpret_cast<const float4*>(a.f.factors);
    float4* G4 = reinterpret_cast<float4*>(a.g_factors);
    const tvm_field_desc& f = a.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 3, quad = lane >> 2;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (threadIdx.x == 0) s_next = BWD_WARPS;
    __syncthreads();
    const long long base = (long long)blockIdx.x * a.rays_per_cta;
    const int S = a.S;
    int local = warp;

    while (local < a.rays_per_cta) {
        const long long r = base + local;
        if (r >= a.n_rays) break;
        TvmRay ray;
        {
            const float* rp = a.rays + r * a.ray_stride;
#pragma unroll
            for (int c = 0; c < 3; ++c) { ray.o[c] = __ldg(rp + c); ray.d[c] = __ldg(rp + 3 + c); }
        }
        tvm_init_ray(f, ray, a.jitter ? __ldg(a.jitter + r) : 0.f, a.S, (a.flags & TVM_F_POINT_SAMPLES) != 0);

        // upstream gradients of this ray, distributed like the forward's accumulator
        float4 gF[3][G];
        float total = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int j = sub + 4 * g;
                gF[k][g] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.d_ray_feat && j < (CA4 > 0 ? CA4 : (f.n_app[k] >> 2))) {
                    gF[k][g] = __ldg(reinterpret_cast<const float4*>(a.d_ray_feat + r * a.ta + a.app_off[k]) + j);
                    const float4 Fv = __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * a.ta + a.app_off[k]) + j);
                    total += f4_dot(gF[k][g], Fv);
                }
            }
        total = quad_sum(total);                              // every quad holds the same slices
        const float g_acc = a.d_acc ? __ldg(a.d_acc + r) : 0.f;
        total = fmaf(g_acc, __ldg(a.acc + r), total);         // Total = sum_k w_k c_k

        float T = 1.f, run = 0.f;
        float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};   // sum dL/dp and sum z*dL/dp (sub==0 lanes)
        const TvmBlockMask bm = tvm_block_prepass(f, ray, S, lane);

        // TVM_F_EARLY_TERM (only valid without a loss on alpha, and with a forward that used it too): samples behind
        // T < eps carry weights <= eps, their share of every gradient is below eps relative — stop like the forward
        const bool early = (a.flags & TVM_F_EARLY_TERM) && !a.d_alpha;
        for (int i0 = 0; i0 < S; i0 += 32) {
            if (early && T < f.early_term_eps) break;
            if (!bm.test(i0 >> 5)) continue;
            const int i = i0 + lane;
            const bool in_range = i < S;
            const float z = tvm_sample_z(f, ray, i);
            float p[3];
            const bool inside = tvm_sample_point(f, ray, z, p) && in_range;
            bool keep = inside;
            if (f.occ_cells != nullptr && inside) keep = tvm_occupancy_keep(f, p);
            const unsigned vmask = __ballot_sync(FULL, keep);
            if (vmask) {
                float n[3];
                tvm_normalize(f, p, n);
                const float zn = tvm_sample_z(f, ray, i + 1);
                const float dist = (i < S - 1) ? rn_sub(zn, z) : 0.f;
                const float delta = rn_mul(dist, f.distance_scale);
                const int nv = __popc(vmask), rank = __popc(vmask & lt_mask);
                // ---- recompute sigma_feature (quads)
                if (keep) s_slot[warp][rank] = make_float4(n[0], n[1], n[2], 0.f);
                __syncwarp();
                for (int g = 0; g * 8 < nv; ++g) {
                    const int ci = g * 8 + quad;
                    float part = 0.f;
                    float dq[3] = {0.f, 0.f, 0.f};
                    if (ci < nv) {
                        const float4 s = s_slot[warp][ci];
                        const float q[3] = {s.x, s.y, s.z};
                        // pose-only mode (frozen factors): the coordinate derivative of the feature does not depend on
                        // the upstream scalar, so it is taken here, on the texels this pass loads anyway, and the
                        // second density pass below disappears
                        if (POSE && !SCATTER) part = density_bwd<false, true, CS4>(f, q, 1.0f, sub, nullptr, dq);
                        else part = density_partial<CS4>(f, q, sub);
                    }
                    part = quad_sum(part);
                    if (POSE && !SCATTER) {
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) dq[cc] = quad_sum(dq[cc]);
                        if (sub == 0 && ci < nv) s_dn[warp][ci] = make_float4(dq[0], dq[1], dq[2], 0.f);
                    }
                    if (sub == 0 && ci < nv) s_ret[warp][ci] = part;
                }
                __syncwarp();
                const float feat = keep ? s_ret[warp][rank] : 0.f;
                const float sigma = keep ? tvm_density(f, feat) : 0.f;
                const float alpha = 1.f - expf(-sigma * delta);
                float incl = 1.f - alpha + 1e-10f;
                const float one_m = incl;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl *= v;
                }
                float excl = __shfl_up_sync(FULL, incl, 1);
                if (lane == 0) excl = 1.f;
                const float Ti = T * excl;
                const float w = alpha * Ti;
                T *= __shfl_sync(FULL, incl, 31);
                // ---- appearance: c_i = g_acc + gF . phi_i, scatter w_i * gF into the app factors
                float c = g_acc;
                const bool app = keep && (w > f.weight_thres);
                const unsigned amask = __ballot_sync(FULL, app);
                if (RUNS && runs && amask && a.d_ray_feat) {
                    // quad q walks the contiguous run q of the block's appearance samples, one (plane, slice) at a time,
                    // with the pending corner / tap gradients in registers
                    const int na = __popc(amask), ranka = __popc(amask & lt_mask);
                    __syncwarp();
                    if (app) {
                        const SampleTaps st = make_sample_taps(f, n);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const unsigned C4 = CA4 > 0 ? CA4 : (unsigned)(f.n_app[k] >> 2);
                            s_cell[warp][ranka * 3 + k] = make_cell_taps(st, k, a.sec.aRow[k], C4, a.sec.aP[k], a.sec.aL[k]);
                        }
                        s_wt[warp][ranka] = w;
                        *reinterpret_cast<float4*>(&s_part[warp][ranka * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    __syncwarp();
                    const int R = (na + 7) >> 3, b = quad * R, e = min(b + R, na);
                    const float* gRow = a.d_ray_feat + r * a.ta;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int C4 = CA4 > 0 ? CA4 : (f.n_app[k] >> 2);
                        const float4* gk = reinterpret_cast<const float4*>(gRow + a.app_off[k]);
#pragma unroll 1
                        for (int j = sub; j < C4; j += 4)
                            vm_run_bwd<true>(F4, G4, s_cell[warp], k, s_wt[warp], b, e, (unsigned)j, TVM_LDG4(gk + j),
                                             s_part[warp], sub);
                    }
                    __syncwarp();
                    if (app) {
                        const float4 pp = *reinterpret_cast<const float4*>(&s_part[warp][ranka * 4]);
                        c += (pp.x + pp.y) + (pp.z + pp.w);
                    }
                } else if (amask && a.d_ray_feat) {
                    const int na = __popc(amask), ranka = __popc(amask & lt_mask);
                    __syncwarp();
                    if (app) { s_slot[warp][ranka] = make_float4(n[0], n[1], n[2], w); s_z[warp][ranka] = z; }
                    __syncwarp();
                    for (int g = 0; g * 8 < na; ++g) {
                        const int ci = g * 8 + quad;
                        float dot = 0.f;
                        float dn[3] = {0.f, 0.f, 0.f};
                        if (ci < na) {
                            const float4 s = s_slot[warp][ci];
                            const float q[3] = {s.x, s.y, s.z};
#if TVM_BWD_ROLL_APP
                            dot = app_bwd_rolled<G, SCATTER, POSE, CA4>(f, q, s.w, sub, a.d_ray_feat + r * a.ta, a.app_off,
                                                                        a.g_factors, dn);
#else
                            dot = app_bwd<G, SCATTER, POSE, CA4>(f, q, s.w, sub, gF, a.g_factors, dn);
#endif
                        }
                        dot = quad_sum(dot);
                        if (POSE) {
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) dn[cc] = quad_sum(dn[cc]);
                            if (sub == 0 && ci < na) {
                                const float zz = s_z[warp][ci];
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) {
                                    const float dp = dn[cc] * f.inv_aabb[cc];
                                    go[cc] += dp;
                                    gd[cc] = fmaf(dp, zz, gd[cc]);
                                }
                            }
                        }
                        if (sub == 0 && ci < na) s_ret[warp][ci] = dot;
                    }
                    __syncwarp();
                    if (app) c += s_ret[warp][ranka];
                }
                // ---- suffix sums via Total - prefix, then dL/dalpha -> dL/dsigma -> dL/dfeat
                const float wc = keep ? w * c : 0.f;
                float pre = wc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v = __shfl_up_sync(FULL, pre, o);
                    if (lane >= o) pre += v;
                }
                const float suffix = total - (run + pre);
                run += __shfl_sync(FULL, pre, 31);
                float dalpha = fmaf(Ti, c, -suffix / one_m);
                if (a.d_alpha && in_range) dalpha += __ldg(a.d_alpha + r * S + i);
                const float dsigma = dalpha * delta * (1.f - alpha);
                const float dfeat = keep ? dsigma * tvm_density_grad(f, feat) : 0.f;
                // ---- density: pose-only mode finishes in the owning lane (dL/dn = dfeat * d feat/dn from the recompute pass)
                if (POSE && !SCATTER) {
                    if (keep) {
                        const float4 dnv = s_dn[warp][rank];
                        const float dnc[3] = {dnv.x, dnv.y, dnv.z};
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            const float dp = dfeat * dnc[cc] * f.inv_aabb[cc];
                            go[cc] += dp;
                            gd[cc] = fmaf(dp, z, gd[cc]);
                        }
                    }
                }
                // ---- density scatter (quads again)
                const bool live = !(POSE && !SCATTER) && keep && dfeat != 0.f;
                const unsigned dmask = __ballot_sync(FULL, live);
                if (RUNS && runs && dmask) {
                    const int nd = __popc(dmask), rankd = __popc(dmask & lt_mask);
                    __syncwarp();
                    if (live) {
                        const SampleTaps st = make_sample_taps(f, n);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const unsigned C4 = CS4 > 0 ? CS4 : (unsigned)(f.n_sigma[k] >> 2);
                            s_cell[warp][rankd * 3 + k] = make_cell_taps(st, k, a.sec.dRow[k], C4, a.sec.dP[k], a.sec.dL[k]);
                        }
                        s_wt[warp][rankd] = dfeat;
                    }
                    __syncwarp();
                    const int R = (nd + 7) >> 3, b = quad * R, e = min(b + R, nd);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int C4 = CS4 > 0 ? CS4 : (f.n_sigma[k] >> 2);
                        if (sub < C4)
                            vm_run_bwd<false>(F4, G4, s_cell[warp], k, s_wt[warp], b, e, (unsigned)sub,
                                              make_float4(1.f, 1.f, 1.f, 1.f), nullptr, sub);
                    }
                    __syncwarp();
                } else if (dmask) {
                    const int nd = __popc(dmask), rankd = __popc(dmask & lt_mask);
                    __syncwarp();
                    if (live) { s_slot[warp][rankd] = make_float4(n[0], n[1], n[2], dfeat); s_z[warp][rankd] = z; }
                    __syncwarp();
                    for (int g = 0; g * 8 < nd; ++g) {
                        const int ci = g * 8 + quad;
                        float dn[3] = {0.f, 0.f, 0.f};
                        if (ci < nd) {
                            const float4 s = s_slot[warp][ci];
                            const float q[3] = {s.x, s.y, s.z};
                            density_bwd<SCATTER, POSE, CS4>(f, q, s.w, sub, a.g_factors, dn);
                        }
                        if (POSE) {
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) dn[cc] = quad_sum(dn[cc]);
                            if (sub == 0 && ci < nd) {
                                const float zz = s_z[warp][ci];
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) {
                                    const float dp = dn[cc] * f.inv_aabb[cc];
                                    go[cc] += dp;
                                    gd[cc] = fmaf(dp, zz, gd[cc]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }

        if (POSE && a.g_rays) {
            // xyz_i = o + d*z_i, z_i = t0 + const  =>  dL/do = sum dp, dL/dd = sum z*dp, dL/dt0 = d . sum dp;
            // t0 = clamp(max_c min((hi-o)/v, (lo-o)/v)) routes its gradient to the selected slab (autograd of
            // minimum/amax/clamp, models/tensorBase.py:499-502)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) { go[cc] = warp_sum(go[cc]); gd[cc] = warp_sum(gd[cc]); }
            if (lane == 0) {
                const float gt0 = go[0] * ray.d[0] + go[1] * ray.d[1] + go[2] * ray.d[2];
                float best = -INFINITY, bv = 1.f;
                int bc = 0;
                bool bzero = false;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const bool zero = ray.d[cc] == 0.0f;
                    const float v = zero ? 1e-6f : ray.d[cc];
                    const float ra = rn_div(rn_sub(f.aabb[3 + cc], ray.o[cc]), v);
                    const float rb = rn_div(rn_sub(f.aabb[cc], ray.o[cc]), v);
                    const float m = fminf(ra, rb);
                    if (m > best) { best = m; bc = cc; bv = v; bzero = zero; }
                }
                if (!(a.flags & TVM_F_POINT_SAMPLES) && best >= f.near_t && best <= f.far_t) {   // clamp passes the gradient only inside [near, far]
                    const float inv = 1.f / bv;
                    go[bc] -= gt0 * inv;
                    if (!bzero) gd[bc] -= gt0 * best * inv;
                }
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) { a.g_rays[r * 6 + cc] = go[cc]; a.g_rays[r * 6 + 3 + cc] = gd[cc]; }
            }
        }
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(&s_next, 1);
        local = __shfl_sync(FULL, nxt, 0);
    }
}

template <int G, int CS4, int CA4>
int dispatch(const BwdArgs& a, cudaStream_t st) {
    const long long ctas = (a.n_rays + a.rays_per_cta - 1) / a.rays_per_cta;
    const bool scatter = a.g_factors != nullptr, pose = a.g_rays != nullptr;
    if (scatter && pose) { tvm_count_launch(); march_bwd_kernel<G, true, true, CS4, CA4><<<(unsigned)ctas, BWD_WARPS * 32, 0, st>>>(a); }
    else if (scatter) { tvm_count_launch(); march_bwd_kernel<G, true, false, CS4, CA4><<<(unsigned)ctas, BWD_WARPS * 32, 0, st>>>(a); }
    else if (pose) { tvm_count_launch(); march_bwd_kernel<G, false, true, CS4, CA4><<<(unsigned)ctas, BWD_WARPS * 32, 0, st>>>(a); }
    TVM_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int tvm_march_bwd(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                             int n_samples, const float* jitter, uint32_t flags, const float* d_ray_feat,
                             const float* d_acc, const float* d_alpha, float* g_factors, float* g_rays,
                             const void* ws, size_t ws_bytes, void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (ray_stride < 6 || n_samples <= 0 || n_samples > 32 * TVM_MAX_BLOCKS || n_rays < 0) return TVM_E_SHAPE;
    if (n_rays == 0 || (!g_factors && !g_rays)) return 0;
    if (!rays || !ws || !desc->factors) return TVM_E_NULL;
    const TvmWorkspace w = tvm_ws_layout(desc, n_rays);
    if (ws_bytes < w.total) return TVM_E_WORKSPACE;
    BwdArgs a{};
    a.f = *desc;
    a.rays = rays; a.n_rays = n_rays; a.ray_stride = ray_stride; a.S = n_samples; a.jitter = jitter;
    a.flags = flags;
    a.d_ray_feat = d_ray_feat; a.d_acc = d_acc; a.d_alpha = d_alpha;
    a.g_factors = g_factors; a.g_rays = g_rays;
    a.ray_feat = (const float*)((const char*)ws + w.ray_feat);
    a.acc = (const float*)((const char*)ws + w.acc);
    a.ta = tvm_total_app(desc);
    a.app_off[0] = 0; a.app_off[1] = desc->n_app[0]; a.app_off[2] = desc->n_app[0] + desc->n_app[1];
    a.sec = tvm_sections(*desc);
    {
        long long rpc = n_rays / (TVM_SM_COUNT * 8);
        a.rays_per_cta = (int)(rpc < BWD_WARPS ? BWD_WARPS : (rpc > BWD_RAYS_PER_CTA ? BWD_RAYS_PER_CTA : rpc));
#ifdef TVM_BWD_RPC_FIXED
        a.rays_per_cta = TVM_BWD_RPC_FIXED;
#endif
    }
    int gmax = 0;
    for (int k = 0; k < 3; ++k) gmax = max(gmax, (desc->n_app[k] + 15) / 16);
    cudaStream_t st = (cudaStream_t)stream;
    bool lego = true;
    for (int k = 0; k < 3; ++k) lego = lego && desc->n_sigma[k] == 16 && desc->n_app[k] == 48;
    if (lego) return dispatch<3, 4, 12>(a, st);
    if (gmax <= 1) return dispatch<1, 0, 0>(a, st);
    if (gmax == 2) return dispatch<2, 0, 0>(a, st);
    return dispatch<3, 0, 0>(a, st);
}
