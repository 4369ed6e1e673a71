// shade_bwd.cu — backward of the per-ray shading stage (fp32).
//
// Replaces the autograd backward the reference runs for the once-per-ray tail of TensorBase.forward
// (models/tensorBase.py:886-904: basis_mat, MLPRender_Fea :185-195 with positional_encoding :14-20, background
// blend + clamp), driven by train.py:338 and inerf/estimate_pose_inerf.py:178.
//   in : d(rgb_map) [n][3]
//   out: d(ray_feat) [n][sum n_app] and d(acc) [n] (consumed by tvm_march_bwd), d(viewdirs) [n][3] (pose mode),
//        parameter gradients of basis_mat ([app_dim][sum n_app], torch layout) and of the MLP (packed layout of
//        tvm_pack_mlp, unpacked by tvm_unpack_mlp_grads).
// One CTA re-runs the forward for a tile of 64 rays keeping every activation in shared memory (X, h1, h2, F:
// ~200 KB), then walks the chain backwards; each weight gradient is a [out x 64] . [64 x in] product formed in
// registers and flushed with 16-byte vector reductions. Batches that would leave SMs idle with 64-ray tiles
// (n <= 148 * 32 rays: the 4096-ray train step, the 1024-ray iNeRF step) run the 32-ray instantiation instead.
#include "tvm_common.cuh"

namespace {

constexpr int SB_THREADS = 256;
#ifndef TVM_SHADE_BWD_SMALL_BLOCKS
#define TVM_SHADE_BWD_SMALL_BLOCKS 1      // resident CTAs/SM the 32-ray instantiation is compiled for
#endif
#ifndef TVM_SHADE_BWD_SMALL_ALWAYS
#define TVM_SHADE_BWD_SMALL_ALWAYS 0      // 1: 32-ray tiles at every batch size
#endif
constexpr int FC = TVM_FEATURE_C;
constexpr unsigned FULL = 0xffffffffu;
constexpr int BS = 36;          // row stride of the transposed basis in smem: 16-B aligned rows, float4 reads by
                                // consecutive rows land on distinct bank groups (36 mod 32 = 4)

struct ShadeBwdArgs {
    const float* rays;
    long long n_rays;
    int ray_stride;
    const float* bg;
    const float* d_rgb;        // [n][3]
    const float* d_acc_in;     // [n] or NULL: upstream gradient of the acc_map output
    float* d_ray_feat;         // [n][ta]
    float* d_acc;              // [n]
    float* d_view;             // [n][3] or NULL
    float* g_basis;            // [app_dim][ta] or NULL
    float* g_mlp;              // packed layout or NULL
    const float* ray_feat;
    const float* acc;
    const int* app_count;
    const float* basis;
    const float* mlp;          // packed weights
    TvmMlpLayout m;
    int ta, app_dim, fea_pe, view_pe;
};

// acc[R][4] += A[R rows][K] (smem, row stride lda) * W[K][ldw] (global), columns 4*tx .. 4*tx+3.
// The weight rows of GEMM_DEPTH k-steps (4 rows each) are kept in flight in registers: with 8 warps per CTA a k-step
// that waited for its own loads paid the full L2 latency (~0.4 us) 166 times per tile.
#ifndef TVM_SHADE_BWD_DEPTH
#define TVM_SHADE_BWD_DEPTH 2
#endif
constexpr int GEMM_DEPTH = TVM_SHADE_BWD_DEPTH;
template <int R>
__device__ __forceinline__ void rows_gemm(float (&acc)[R][4], const float* __restrict__ sA, int lda, int row0, int K,
                                          const float* __restrict__ W, int ldw, int col0) {
    float4 w[GEMM_DEPTH][4];
    const int n = K >> 2;
    const float* wp = W + col0;
#pragma unroll
    for (int d = 0; d < GEMM_DEPTH; ++d)
        if (d < n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) w[d][e] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(d * 4 + e) * ldw));
        }
    for (int s0 = 0; s0 < n; s0 += GEMM_DEPTH) {
#pragma unroll
        for (int d = 0; d < GEMM_DEPTH; ++d) {
            const int s = s0 + d;
            if (s < n) {
                const float4 w0 = w[d][0], w1 = w[d][1], w2 = w[d][2], w3 = w[d][3];
                if (s + GEMM_DEPTH < n) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[d][e] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)((s + GEMM_DEPTH) * 4 + e) * ldw));
                }
                const int k = s * 4;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 x = *reinterpret_cast<const float4*>(sA + (row0 + r) * lda + k);
                    acc[r][0] = fmaf(x.x, w0.x, fmaf(x.y, w1.x, fmaf(x.z, w2.x, fmaf(x.w, w3.x, acc[r][0]))));
                    acc[r][1] = fmaf(x.x, w0.y, fmaf(x.y, w1.y, fmaf(x.z, w2.y, fmaf(x.w, w3.y, acc[r][1]))));
                    acc[r][2] = fmaf(x.x, w0.z, fmaf(x.y, w1.z, fmaf(x.z, w2.z, fmaf(x.w, w3.z, acc[r][2]))));
                    acc[r][3] = fmaf(x.x, w0.w, fmaf(x.y, w1.w, fmaf(x.z, w2.w, fmaf(x.w, w3.w, acc[r][3]))));
                }
            }
        }
    }
}

// G[v0+j][u0+i] += sum_ray U[ray][u0+i] * V[ray][v0+j]  (8x8 register block; G is global, row stride ldg, i contiguous)
template <int RAYS>
__device__ __forceinline__ void outer8x8_flush(const float* __restrict__ sU, int ldu, int u0, const float* __restrict__ sV,
                                               int ldv, int v0, int vmax, float* __restrict__ G, int ldg) {
    float p[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) p[j][i] = 0.f;
    for (int ray = 0; ray < RAYS; ++ray) {
        const float4 ua = *reinterpret_cast<const float4*>(sU + ray * ldu + u0);
        const float4 ub = *reinterpret_cast<const float4*>(sU + ray * ldu + u0 + 4);
        const float4 va = *reinterpret_cast<const float4*>(sV + ray * ldv + v0);
        const float4 vb = *reinterpret_cast<const float4*>(sV + ray * ldv + v0 + 4);
        const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
        const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) p[j][i] = fmaf(u[i], v[j], p[j][i]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (v0 + j >= vmax) break;
        float4* g = reinterpret_cast<float4*>(G + (size_t)(v0 + j) * ldg + u0);
        atomicAdd(g, make_float4(p[j][0], p[j][1], p[j][2], p[j][3]));
        atomicAdd(g + 1, make_float4(p[j][4], p[j][5], p[j][6], p[j][7]));
    }
}

template <int RAYS>      // rays per CTA: 64, or 32 for batches that would not fill the SMs
__global__ void __launch_bounds__(SB_THREADS, RAYS == 32 ? TVM_SHADE_BWD_SMALL_BLOCKS : 1) shade_bwd_kernel(const __grid_constant__ ShadeBwdArgs a) {
    constexpr int SB_RAYS = RAYS, R = RAYS / 8;   // R rays per warp in the GEMM phases
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ta = a.ta, k1 = a.m.k1, fs = ta + 4;
    float* sB = smem;                          // basis^T padded [ta][BS] (columns >= app_dim zero)
    float* sF = sB + ta * BS;                  // ray_feat tile [64][ta+4]
    float* sX = sF + SB_RAYS * fs;             // MLP input [64][k1]
    float* sH1 = sX + SB_RAYS * k1;            // [64][FC]
    float* sH2 = sH1 + SB_RAYS * FC;           // [64][FC]
    float* sG = sH2 + SB_RAYS * FC;            // gradient scratch [64][max(k1,FC)]  (g_h2 / g_h1: stride FC)
    float* sS = sG + SB_RAYS * (k1 > FC ? k1 : FC);   // per-ray scalars [64][8]: rgb[3], g_z3[3], acc, lit
    const long long r0 = (long long)blockIdx.x * SB_RAYS;
    const float* w1t = a.mlp + a.m.w1t; const float* b1 = a.mlp + a.m.b1;
    const float* w2t = a.mlp + a.m.w2t; const float* b2 = a.mlp + a.m.b2;
    const float* w3 = a.mlp + a.m.w3;   const float* b3 = a.mlp + a.m.b3;
    const float* w1n = a.mlp + a.m.w1n; const float* w2n = a.mlp + a.m.w2n;

    // ================= forward recompute (same as shade_fwd_kernel) =================
    for (int i = tid; i < a.app_dim * ta; i += SB_THREADS) {
        const int j = i / ta, c = i - j * ta;
        sB[c * BS + j] = __ldg(a.basis + i);
    }
    for (int i = tid; i < (32 - a.app_dim) * ta; i += SB_THREADS) {
        const int c = i / (32 - a.app_dim), j = a.app_dim + i - c * (32 - a.app_dim);
        sB[c * BS + j] = 0.f;
    }
    for (int i = tid; i < SB_RAYS * (ta >> 2); i += SB_THREADS) {
        const int ray = i / (ta >> 2), c4 = i - ray * (ta >> 2);
        const long long r = r0 + ray;
        const float4 v = (r < a.n_rays) ? __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * ta) + c4)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sF + ray * fs + c4 * 4) = v;
    }
    __syncthreads();
    {   // feat = B . F  (RAYS/32 rays x 4 cols per thread)
        constexpr int RB = RAYS / 32;
        const int tx = tid & 7, ty = tid >> 3;
        float o[RB][4];
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) o[rr][0] = o[rr][1] = o[rr][2] = o[rr][3] = 0.f;
        for (int c = 0; c < ta; c += 4) {
            const float4 q0 = *reinterpret_cast<const float4*>(sB + (c + 0) * BS + tx * 4);
            const float4 q1 = *reinterpret_cast<const float4*>(sB + (c + 1) * BS + tx * 4);
            const float4 q2 = *reinterpret_cast<const float4*>(sB + (c + 2) * BS + tx * 4);
            const float4 q3 = *reinterpret_cast<const float4*>(sB + (c + 3) * BS + tx * 4);
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
                const float4 x = *reinterpret_cast<const float4*>(sF + (ty * RB + rr) * fs + c);
                o[rr][0] = fmaf(x.x, q0.x, fmaf(x.y, q1.x, fmaf(x.z, q2.x, fmaf(x.w, q3.x, o[rr][0]))));
                o[rr][1] = fmaf(x.x, q0.y, fmaf(x.y, q1.y, fmaf(x.z, q2.y, fmaf(x.w, q3.y, o[rr][1]))));
                o[rr][2] = fmaf(x.x, q0.z, fmaf(x.y, q1.z, fmaf(x.z, q2.z, fmaf(x.w, q3.z, o[rr][2]))));
                o[rr][3] = fmaf(x.x, q0.w, fmaf(x.y, q1.w, fmaf(x.z, q2.w, fmaf(x.w, q3.w, o[rr][3]))));
            }
        }
#pragma unroll
        for (int rr = 0; rr < RB; ++rr)
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (tx * 4 + e < a.app_dim) sX[(ty * RB + rr) * k1 + tx * 4 + e] = o[rr][e];
        if (tid < SB_RAYS * 3) {
            const int vr = tid / 3, c = tid - vr * 3;
            const long long r = r0 + vr;
            sX[vr * k1 + a.app_dim + c] = (r < a.n_rays) ? __ldg(a.rays + r * a.ray_stride + 3 + c) : 0.f;
        }
        for (int i = tid; i < SB_RAYS * (k1 - a.m.in_c); i += SB_THREADS) {
            const int pr = i / (k1 - a.m.in_c), c = i - pr * (k1 - a.m.in_c);
            sX[pr * k1 + a.m.in_c + c] = 0.f;
        }
    }
    __syncthreads();
    const int nbase = a.app_dim + 3;
    const int sin_f = nbase, cos_f = sin_f + a.app_dim * a.fea_pe;
    const int sin_v = cos_f + a.app_dim * a.fea_pe, cos_v = sin_v + 3 * a.view_pe;
    const int chs = nbase > 32 ? 6 : 5;                           // lanes run over the channels of one ray
    for (int it = tid; it < (SB_RAYS << chs); it += SB_THREADS) {
        const int ch = it & ((1 << chs) - 1), ray = it >> chs;
        if (ch >= nbase) continue;
        const float v = sX[ray * k1 + ch];
        const bool is_feat = ch < a.app_dim;
        const int nf = is_feat ? a.fea_pe : a.view_pe;
        const int cc = is_feat ? ch : ch - a.app_dim;
        const int sb = is_feat ? sin_f : sin_v, cb = is_feat ? cos_f : cos_v;
        float scale = 1.f;
        for (int j = 0; j < nf; ++j) {
            float s, c;
            sincosf(v * scale, &s, &c);
            sX[ray * k1 + sb + cc * nf + j] = s;
            sX[ray * k1 + cb + cc * nf + j] = c;
            scale *= 2.f;
        }
    }
    __syncthreads();
    const int tx = lane, ty = warp;
    float acc[R][4];
    {
        const float4 b = __ldg(reinterpret_cast<const float4*>(b1) + tx);
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r][0] = b.x; acc[r][1] = b.y; acc[r][2] = b.z; acc[r][3] = b.w; }
    }
    rows_gemm<R>(acc, sX, k1, ty * R, k1, w1t, FC, tx * 4);
#pragma unroll
    for (int r = 0; r < R; ++r)
        *reinterpret_cast<float4*>(sH1 + (ty * R + r) * FC + tx * 4) =
            make_float4(fmaxf(acc[r][0], 0.f), fmaxf(acc[r][1], 0.f), fmaxf(acc[r][2], 0.f), fmaxf(acc[r][3], 0.f));
    __syncthreads();
    {
        const float4 b = __ldg(reinterpret_cast<const float4*>(b2) + tx);
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r][0] = b.x; acc[r][1] = b.y; acc[r][2] = b.z; acc[r][3] = b.w; }
    }
    rows_gemm<R>(acc, sH1, FC, ty * R, FC, w2t, FC, tx * 4);
#pragma unroll
    for (int r = 0; r < R; ++r)
        *reinterpret_cast<float4*>(sH2 + (ty * R + r) * FC + tx * 4) =
            make_float4(fmaxf(acc[r][0], 0.f), fmaxf(acc[r][1], 0.f), fmaxf(acc[r][2], 0.f), fmaxf(acc[r][3], 0.f));
    __syncthreads();
    {   // layer 3 + sigmoid + blend, and the head of the backward: g_z3, d_acc
        const float4 wr = __ldg(reinterpret_cast<const float4*>(w3) + lane);
        const float4 wg = __ldg(reinterpret_cast<const float4*>(w3 + FC) + lane);
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w3 + 2 * FC) + lane);
        for (int rr = 0; rr < R; ++rr) {
            const int ray = warp * R + rr;
            const long long r = r0 + ray;
            const float4 h = *reinterpret_cast<const float4*>(sH2 + ray * FC + lane * 4);
            float vr = h.x * wr.x + h.y * wr.y + h.z * wr.z + h.w * wr.w;
            float vg = h.x * wg.x + h.y * wg.y + h.z * wg.z + h.w * wg.w;
            float vb = h.x * wb.x + h.y * wb.y + h.z * wb.z + h.w * wb.w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vr += __shfl_xor_sync(FULL, vr, o);
                vg += __shfl_xor_sync(FULL, vg, o);
                vb += __shfl_xor_sync(FULL, vb, o);
            }
            float gz = 0.f, dacc_c = 0.f;
            if (lane < 3 && r < a.n_rays) {
                const float v = (lane == 0 ? vr : (lane == 1 ? vg : vb)) + __ldg(b3 + lane);
                const bool lit = __ldg(a.app_count + r) > 0;
                const float c = lit ? 1.f / (1.f + expf(-v)) : 0.f;
                const float ac = __ldg(a.acc + r), bgc = __ldg(a.bg + lane);
                const float out = c * ac + bgc * (1.f - ac);
                const float go = (out >= 0.f && out <= 1.f) ? __ldg(a.d_rgb + r * 3 + lane) : 0.f;   // clamp backward
                dacc_c = go * (c - bgc);
                gz = lit ? go * ac * c * (1.f - c) : 0.f;
            }
            // d_acc = sum over the 3 colour lanes
            float dsum = dacc_c + __shfl_down_sync(FULL, dacc_c, 1) + __shfl_down_sync(FULL, dacc_c, 2);
            if (lane < 3) sS[ray * 8 + 3 + lane] = gz;
            if (lane == 0 && r < a.n_rays)
                a.d_acc[r] = dsum + (a.d_acc_in ? __ldg(a.d_acc_in + r) : 0.f);
        }
    }
    __syncthreads();

    // ================= backward =================
    // g_h2 = (g_z3 . W3) * [h2 > 0]  -> sG (stride FC)
    {
        const float4 wr = __ldg(reinterpret_cast<const float4*>(w3) + tx);
        const float4 wg = __ldg(reinterpret_cast<const float4*>(w3 + FC) + tx);
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w3 + 2 * FC) + tx);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ray = ty * R + r;
            const float g0 = sS[ray * 8 + 3], g1 = sS[ray * 8 + 4], g2 = sS[ray * 8 + 5];
            const float4 h = *reinterpret_cast<const float4*>(sH2 + ray * FC + tx * 4);
            float4 g;
            g.x = h.x > 0.f ? g0 * wr.x + g1 * wg.x + g2 * wb.x : 0.f;
            g.y = h.y > 0.f ? g0 * wr.y + g1 * wg.y + g2 * wb.y : 0.f;
            g.z = h.z > 0.f ? g0 * wr.z + g1 * wg.z + g2 * wb.z : 0.f;
            g.w = h.w > 0.f ? g0 * wr.w + g1 * wg.w + g2 * wb.w : 0.f;
            *reinterpret_cast<float4*>(sG + ray * FC + tx * 4) = g;
        }
    }
    if (a.g_mlp) {   // gW3[c][n] = sum_ray g_z3[ray][c] * h2[ray][n];  gb3
        float* gw3 = a.g_mlp + a.m.w3;
        for (int o = tid; o < 3 * FC; o += SB_THREADS) {
            const int c = o / FC, n = o - c * FC;
            float s = 0.f;
            for (int ray = 0; ray < SB_RAYS; ++ray) s = fmaf(sS[ray * 8 + 3 + c], sH2[ray * FC + n], s);
            atomicAdd(gw3 + o, s);
        }
        if (tid < 3) {
            float s = 0.f;
            for (int ray = 0; ray < SB_RAYS; ++ray) s += sS[ray * 8 + 3 + tid];
            atomicAdd(a.g_mlp + a.m.b3 + tid, s);
        }
    }
    __syncthreads();
    if (a.g_mlp) {   // gW2^T[k][n] += sum_ray h1[ray][k] * g_h2[ray][n]   (packed layout is transposed: [k][n]);  gb2
        const int n0 = (tid & 15) * 8, kq = (tid >> 4) * 8;
        outer8x8_flush<RAYS>(sG, FC, n0, sH1, FC, kq, FC, a.g_mlp + a.m.w2t, FC);
        if (tid < FC) {
            float s = 0.f;
            for (int ray = 0; ray < SB_RAYS; ++ray) s += sG[ray * FC + tid];
            atomicAdd(a.g_mlp + a.m.b2 + tid, s);
        }
    }
    // g_h1 = (g_h2 . W2) * [h1 > 0]   (W2 in torch orientation [n][k]: reduction over n)
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
    rows_gemm<R>(acc, sG, FC, ty * R, FC, w2n, FC, tx * 4);
    __syncthreads();                             // every read of g_h2 (outer product + gemm) is done
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float4 h = *reinterpret_cast<const float4*>(sH1 + (ty * R + r) * FC + tx * 4);
        *reinterpret_cast<float4*>(sG + (ty * R + r) * FC + tx * 4) =
            make_float4(h.x > 0.f ? acc[r][0] : 0.f, h.y > 0.f ? acc[r][1] : 0.f, h.z > 0.f ? acc[r][2] : 0.f,
                        h.w > 0.f ? acc[r][3] : 0.f);
    }
    __syncthreads();
    if (a.g_mlp) {   // gW1^T[k][n] += sum_ray x[ray][k] * g_h1[ray][n];  gb1
        const int n0 = (tid & 15) * 8;
        for (int kq = (tid >> 4) * 8; kq < k1; kq += 16 * 8)
            outer8x8_flush<RAYS>(sG, FC, n0, sX, k1, kq, k1, a.g_mlp + a.m.w1t, FC);
        if (tid < FC) {
            float s = 0.f;
            for (int ray = 0; ray < SB_RAYS; ++ray) s += sG[ray * FC + tid];
            atomicAdd(a.g_mlp + a.m.b1 + tid, s);
        }
    }
    // g_x = g_h1 . W1  ([64][k1]); computed into registers, then written over sG's successor region (sH2 is dead)
    float* sGX = sH2;                            // [64][k1] needs k1 <= 2*FC: h2 + part of sG?  -> use sH1|sH2 span
    sGX = sH1;                                   // h1 and h2 are both dead after this point (h1 read below first)
    {
        // columns 0..127 then the remainder, 8 rays x 4 cols per thread
        float gx[R][4];
        float gx2[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) { gx[r][0] = gx[r][1] = gx[r][2] = gx[r][3] = 0.f; gx2[r][0] = gx2[r][1] = gx2[r][2] = gx2[r][3] = 0.f; }
        const bool head = tx * 4 < k1;           // k1 may be smaller than FC (fea_pe = view_pe = 0 -> k1 = 32)
        if (head) rows_gemm<R>(gx, sG, FC, ty * R, FC, w1n, k1, tx * 4);
        const bool tail = FC + tx * 4 < k1;
        if (tail) rows_gemm<R>(gx2, sG, FC, ty * R, FC, w1n, k1, FC + tx * 4);
        __syncthreads();                         // all reads of h1 (outer product) and g_h1 are done
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (head)
                *reinterpret_cast<float4*>(sGX + (ty * R + r) * k1 + tx * 4) =
                    make_float4(gx[r][0], gx[r][1], gx[r][2], gx[r][3]);
            if (tail)
                *reinterpret_cast<float4*>(sGX + (ty * R + r) * k1 + FC + tx * 4) =
                    make_float4(gx2[r][0], gx2[r][1], gx2[r][2], gx2[r][3]);
        }
    }
    __syncthreads();
    // g_feat / g_view through the encodings: d sin(v 2^j) = 2^j cos, d cos(v 2^j) = -2^j sin (values are still in sX)
    float* sGF = sG;                             // g_feat [64][32] (cols >= app_dim zero), g_h1 is dead
    for (int it = tid; it < SB_RAYS * 32; it += SB_THREADS) {
        const int ch = it & 31, ray = it >> 5;
        float g = 0.f;
        if (ch < nbase) {
            const bool is_feat = ch < a.app_dim;
            const int nf = is_feat ? a.fea_pe : a.view_pe;
            const int cc = is_feat ? ch : ch - a.app_dim;
            const int sb = is_feat ? sin_f : sin_v, cb = is_feat ? cos_f : cos_v;
            g = sGX[ray * k1 + ch];
            float scale = 1.f;
            for (int j = 0; j < nf; ++j) {
                const int si = sb + cc * nf + j, ci = cb + cc * nf + j;
                g += scale * (sGX[ray * k1 + si] * sX[ray * k1 + ci] - sGX[ray * k1 + ci] * sX[ray * k1 + si]);
                scale *= 2.f;
            }
            if (!is_feat) {
                const long long r = r0 + ray;
                if (a.d_view && r < a.n_rays) a.d_view[r * 3 + cc] = g;
                g = 0.f;
            }
        }
        if (ch >= a.app_dim) g = 0.f;
        sGF[ray * 32 + ch] = g;
    }
    __syncthreads();
    // d_ray_feat[ray][c] = sum_j g_feat[ray][j] * B[j][c]   and   gB[j][c] += sum_ray g_feat[ray][j] * F[ray][c]
    for (int i = tid; i < SB_RAYS * ta; i += SB_THREADS) {
        const int ray = i / ta, c = i - ray * ta;
        const long long r = r0 + ray;
        float s = 0.f;
        for (int j = 0; j < a.app_dim; j += 4) {          // both rows are zero-padded to 32 columns
            const float4 gq = *reinterpret_cast<const float4*>(sGF + ray * 32 + j);
            const float4 bq = *reinterpret_cast<const float4*>(sB + c * BS + j);
            s = fmaf(gq.x, bq.x, fmaf(gq.y, bq.y, fmaf(gq.z, bq.z, fmaf(gq.w, bq.w, s))));
        }
        if (r < a.n_rays) a.d_ray_feat[r * ta + c] = s;
    }
    if (a.g_basis) {
        for (int o = tid; o < a.app_dim * ta; o += SB_THREADS) {
            const int j = o / ta, c = o - j * ta;
            float s = 0.f;
            for (int ray = 0; ray < SB_RAYS; ++ray) s = fmaf(sGF[ray * 32 + j], sF[ray * fs + c], s);
            atomicAdd(a.g_basis + o, s);
        }
    }
}

}  // namespace

extern "C" int tvm_shade_bwd(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                             const float* bg, const float* d_rgb, const float* d_acc_in, float* d_ray_feat,
                             float* d_acc, float* g_basis, float* g_mlp, float* d_view, const void* ws,
                             size_t ws_bytes, void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (n_rays == 0) return 0;
    if (!rays || !bg || !d_rgb || !d_ray_feat || !d_acc || !ws || !desc->basis || !desc->mlp) return TVM_E_NULL;
    if (desc->feature_c != FC || desc->app_dim > 32 || desc->app_dim <= 0) return TVM_E_SHAPE;
    if (d_view && desc->app_dim + 3 > 32) return TVM_E_SHAPE;    // the encoding backward walks 32 channels per ray
    const TvmWorkspace w = tvm_ws_layout(desc, n_rays);
    if (ws_bytes < w.total) return TVM_E_WORKSPACE;
    const char* base = (const char*)ws;
    ShadeBwdArgs a{};
    a.rays = rays; a.n_rays = n_rays; a.ray_stride = ray_stride; a.bg = bg;
    a.d_rgb = d_rgb; a.d_acc_in = d_acc_in; a.d_ray_feat = d_ray_feat; a.d_acc = d_acc; a.d_view = d_view;
    a.g_basis = g_basis; a.g_mlp = g_mlp;
    a.ray_feat = (const float*)(base + w.ray_feat);
    a.acc = (const float*)(base + w.acc);
    a.app_count = (const int*)(base + w.app_count);
    a.basis = desc->basis; a.mlp = desc->mlp;
    a.m = tvm_mlp_layout(desc);
    a.ta = tvm_total_app(desc); a.app_dim = desc->app_dim; a.fea_pe = desc->fea_pe; a.view_pe = desc->view_pe;
    if (a.m.k1 > 2 * FC || a.m.k1 % 8) return TVM_E_SHAPE;       // g_x reuses the h1|h2 span; 8-wide outer blocks
    const bool small = TVM_SHADE_BWD_SMALL_ALWAYS || (n_rays + 31) / 32 <= (long long)TVM_SM_COUNT;     // one wave of 32-ray tiles beats idle SMs
    const size_t tile = small ? 32 : 64;
    const size_t floats = (size_t)a.ta * BS + tile * (a.ta + 4) + tile * a.m.k1 +
                          tile * (a.m.k1 > FC ? a.m.k1 : FC) + tile * FC * 2 + tile * 8;
    const size_t smem = floats * sizeof(float);
    if (smem > 227 * 1024) return TVM_E_SHAPE;
    const long long ctas = (n_rays + (long long)tile - 1) / (long long)tile;
    if (small) {
        static TvmDevMemo smem_set;
        int rc_attr = tvm_ensure_dyn_smem(shade_bwd_kernel<32>, smem, smem_set);
        if (rc_attr) return rc_attr;
        tvm_count_launch(); shade_bwd_kernel<32><<<(unsigned)ctas, SB_THREADS, smem, (cudaStream_t)stream>>>(a);
    } else {
        static TvmDevMemo smem_set;
        int rc_attr = tvm_ensure_dyn_smem(shade_bwd_kernel<64>, smem, smem_set);
        if (rc_attr) return rc_attr;
        tvm_count_launch(); shade_bwd_kernel<64><<<(unsigned)ctas, SB_THREADS, smem, (cudaStream_t)stream>>>(a);
    }
    TVM_LAUNCH_CHECK();
    return 0;
}
