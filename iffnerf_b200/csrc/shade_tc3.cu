// shade_tc3.cu — tensor-core shading with bf16x3 SPLIT operands (flag TVM_F_MLP_TC3): fp32-equivalent results.
//
// Same stage as shade.cu / shade_tc.cu (basis_mat models/tensoRF.py:158,256; MLPRender_Fea
// models/tensorBase.py:165-195; blend + depth tail :898-908).  Every fp32 operand value v is split into
// hi = bf16(v) and lo = bf16(v - hi); each K-step issues THREE tcgen05 MMAs into the same fp32 TMEM accumulator,
//     A_hi.B_hi + A_hi.B_lo + A_lo.B_hi          (the dropped lo.lo term is <= 2^-18 relative)
// which reproduces the fp32 FFMA result to ~1e-6 on rgb while running on the 5th-gen tensor cores.  Operand
// footprint doubles, so only basis_mat and W3 stay resident in shared memory; W1 / W2 (hi+lo images, 80 / 64 KB)
// are streamed from L2 into one weight buffer per tile by TMA bulk copies (cp.async.bulk + mbarrier complete_tx) that
// one thread fires as soon as the MMA reading the buffer has completed: W2 under the first hidden layer's epilogue, the
// next tile's W1 under the last epilogues and the next staging; only the MMA-issuing thread ever waits for them.  One persistent CTA (16 warps) per SM, 128 rays per tile, TMEM lane == ray.
#include "tvm_tc.cuh"

#ifndef TVM_TC3_PE_UNROLL
#define TVM_TC3_PE_UNROLL 8
#endif
constexpr int PE_UNROLL = TVM_TC3_PE_UNROLL;     // copies of the encoding body (sincosf is inlined with its slow path)

namespace {
using namespace tvmtc;

struct Tc3Args {
    const float* rays;
    long long n_rays;
    int ray_stride;
    const float* bg;
    float* rgb;
    float* depth_out;
    float* acc_out;
    TvmPeers peers;
    const float* ray_feat;
    const float* acc;
    const float* depth;
    const int* app_count;
    const unsigned char* wimg;     // [b0h|b0l|w1h|w1l|w2h|w2l|w3h|w3l] bf16 canonical images (tvm_pack_mlp_tc3)
    const float* b1; const float* b2; const float* b3;
    TcDims d;
};

constexpr int TC3K_KC = 80;                   // K columns per chunk (multiple of 16)

struct Tc3kDims { int pbase, n_fpairs, n_vpairs, k1c, n_chunks; };
__host__ __device__ inline Tc3kDims tc3k_dims(const TcDims& d) {
    Tc3kDims c;
    c.pbase = (d.app_dim + 3 + 1) & ~1;
    c.n_fpairs = d.app_dim * d.fea_pe;
    c.n_vpairs = 3 * d.view_pe;
    const int cols = c.pbase + 2 * (c.n_fpairs + c.n_vpairs);
    c.n_chunks = (cols + TC3K_KC - 1) / TC3K_KC;
    c.k1c = c.n_chunks * TC3K_KC;
    return c;
}
// does the whole-K kernel (X and W1 complete in shared memory) fit?  Otherwise the K-chunked kernel serves the head.
__host__ __device__ inline bool tc3_whole_k_fits(const TcDims& d) {
    const int kmax = d.k0 > d.k1 ? (d.k0 > FC ? d.k0 : FC) : (d.k1 > FC ? d.k1 : FC);
    const long long a_half = (long long)TC_RAYS * kmax * 2, w_half = (long long)FC * (d.k1 > FC ? d.k1 : FC) * 2;
    return 2LL * N0 * d.k0 * 2 + 2LL * N3 * FC * 2 + 2 * w_half + 2 * a_half + (2 * FC + 4) * 4 + 32 <= 227 * 1024;
}

struct Tc3Layout {
    int b0, w1, w2, w3, total;      // byte offsets of the (hi|lo) image pairs inside wimg; each pair = 2 * size
    int sz_b0, sz_w1, sz_w2, sz_w3;
};
__host__ __device__ inline Tc3Layout tc3_layout(const TcDims& d) {
    Tc3Layout L;
    const int kw = tc3_whole_k_fits(d) ? d.k1 : tc3k_dims(d).k1c;         // W1 columns as packed (chunked: permuted + padded)
    L.sz_b0 = N0 * d.k0 * 2; L.sz_w1 = FC * kw * 2; L.sz_w2 = FC * FC * 2; L.sz_w3 = N3 * FC * 2;
    L.b0 = 0;
    L.w1 = L.b0 + 2 * L.sz_b0;
    L.w2 = L.w1 + 2 * L.sz_w1;
    L.w3 = L.w2 + 2 * L.sz_w2;
    L.total = L.w3 + 2 * L.sz_w3;
    return L;
}

// D[128 x N] = A . B^T with 2-term split operands: hi.hi + hi.lo + lo.hi per K-step
__device__ __forceinline__ void issue_gemm3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            int K, int N, uint32_t bar) {
    const uint64_t dah = make_desc(a_hi, K), dal = make_desc(a_lo, K);
    const uint64_t dbh = make_desc(b_hi, K), dbl = make_desc(b_lo, K);
    const uint32_t idesc = make_idesc(128, N);
    for (int j = 0; j < K / 16; ++j) {
        const uint64_t o = (uint64_t)(j * 16);
        umma_bf16(tmem_d, dah + o, dbh + o, idesc, j > 0 ? 1u : 0u);
        umma_bf16(tmem_d, dah + o, dbl + o, idesc, 1u);
        umma_bf16(tmem_d, dal + o, dbh + o, idesc, 1u);
    }
    umma_commit(bar);
}

__device__ __forceinline__ void split_store8(unsigned char* hi_base, unsigned char* lo_base, int off, const float (&v)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
        const __nv_bfloat162 hp = __halves2bfloat162(h0, h1);
        h[e] = *reinterpret_cast<const uint32_t*>(&hp);
        l[e] = pack_bf16x2(v[2 * e] - __bfloat162float(h0), v[2 * e + 1] - __bfloat162float(h1));
    }
    *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void split_store1(unsigned char* hi_base, unsigned char* lo_base, int off, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(hi_base + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo_base + off) = __float2bfloat16_rn(v - __bfloat162float(h));
}
// Asynchronous weight streaming: one thread arms the mbarrier with the byte count and fires two TMA bulk copies
// (hi and lo image) L2 -> shared memory; nobody waits until the MMA that consumes the image is about to be issued.
__device__ __forceinline__ void bulk_load_pair(uint32_t dst_hi, uint32_t dst_lo, const unsigned char* src, int bytes_each,
                                               uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2 * bytes_each) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_hi), "l"(src), "r"(bytes_each), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_lo), "l"(src + bytes_each), "r"(bytes_each), "r"(bar) : "memory");
}
__device__ __forceinline__ void coop_copy(unsigned char* dst, const unsigned char* __restrict__ src, int bytes, int tid) {
    for (int i = tid * 16; i < bytes; i += TC_THREADS * 16)
        *reinterpret_cast<uint4*>(dst + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
}

__global__ void __launch_bounds__(TC_THREADS, 1) shade_tc3_kernel(const __grid_constant__ Tc3Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const TcDims& d = a.d;
    const Tc3Layout L = tc3_layout(d);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, cg = warp >> 2;
    const int kmax = d.k0 > d.k1 ? (d.k0 > FC ? d.k0 : FC) : (d.k1 > FC ? d.k1 : FC);
    const int a_half = TC_RAYS * kmax * 2;                     // one A image (hi or lo)
    const int w_half = FC * (d.k1 > FC ? d.k1 : FC) * 2;       // one streamed weight image (hi or lo)
    unsigned char* s_b0 = smem;                                // basis hi | lo (resident)
    unsigned char* s_w3 = s_b0 + 2 * L.sz_b0;                  // W3 hi | lo (resident)
    unsigned char* s_w = s_w3 + 2 * L.sz_w3;                   // streamed W1 / W2: hi | lo
    unsigned char* s_a = s_w + 2 * w_half;                     // A operand: hi | lo
    float* s_bias = reinterpret_cast<float*>(s_a + 2 * a_half);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 2 * FC + 4);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2);        // s_bar[0]: MMA completion, s_bar[1]: weight arrival
    unsigned char* s_ah = s_a;
    unsigned char* s_al = s_a + a_half;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(s_tmem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(s_bar), 1);
        mbar_init(smem_u32(s_bar + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    coop_copy(s_b0, a.wimg + L.b0, 2 * L.sz_b0, tid);
    coop_copy(s_w3, a.wimg + L.w3, 2 * L.sz_w3, tid);
    for (int i = tid; i < 2 * FC + 4; i += TC_THREADS)
        s_bias[i] = i < FC ? __ldg(a.b1 + i) : (i < 2 * FC ? __ldg(a.b2 + i - FC) : (i - 2 * FC < 3 ? __ldg(a.b3 + i - 2 * FC) : 0.f));
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t bar = smem_u32(s_bar);
    const uint32_t sah = smem_u32(s_ah), sal = smem_u32(s_al);
    const uint32_t swh = smem_u32(s_w), swl = smem_u32(s_w + w_half);
    const uint32_t sb0h = smem_u32(s_b0), sb0l = smem_u32(s_b0 + L.sz_b0);
    const uint32_t sw3h = smem_u32(s_w3), sw3l = smem_u32(s_w3 + L.sz_w3);
    const uint32_t wbar = smem_u32(s_bar + 1);
    uint32_t phase = 0, wphase = 0;
    const long long n_tiles = (a.n_rays + TC_RAYS - 1) / TC_RAYS;
    // W1 of the first tile; afterwards the weight buffer is refilled as soon as the MMA reading it has completed
    if (tid == 0 && (long long)blockIdx.x < n_tiles) bulk_load_pair(swh, swl, a.wimg + L.w1, L.sz_w1, wbar);
    bool w1_in_flight = (long long)blockIdx.x < n_tiles;       // uniform: a W1 copy is armed and not yet consumed
    const int nbase = d.app_dim + 3;
    const int sin_f = nbase, cos_f = sin_f + d.app_dim * d.fea_pe;
    const int sin_v = cos_f + d.app_dim * d.fea_pe, cos_v = sin_v + 3 * d.view_pe;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r = tile * TC_RAYS + row;
        const bool live = r < a.n_rays;
        // ---- tiles without a single appearance sample (background: the reference skips the MLP for such rays too,
        // tensorBase.py:876-896) need no GEMM: rgb = bg * (1 - acc).  The W1 copy already in flight simply stays in the
        // weight buffer for the next tile that is shaded.
        if (!__syncthreads_or(live && __ldg(a.app_count + r) > 0)) {
            if (cg == 0 && live) {
                const float ac = __ldg(a.acc + r);
#pragma unroll
                for (int c = 0; c < 3; ++c) tvm_put_rgb(a.peers, a.rgb, r, c, fminf(fmaxf(__ldg(a.bg + c) * (1.f - ac), 0.f), 1.f));
                const float last = __ldg(a.rays + r * a.ray_stride + a.ray_stride - 1);
                tvm_put_depth(a.peers, a.depth_out, r, __ldg(a.depth + r) + (1.f - ac) * last);
                if (a.acc_out) a.acc_out[r] = ac;
            }
            continue;
        }
        if (!w1_in_flight) {          // the previous shaded tile was this CTA's last by its own count, but this one follows
            if (tid == 0) bulk_load_pair(swh, swl, a.wimg + L.w1, L.sz_w1, wbar);
            w1_in_flight = true;
        }
        // ---- stage the ray_feat tile as split A operand (K0); W1 (hi|lo) is already in flight into the weight buffer
        const bool lit_row = live && __ldg(a.app_count + r) > 0;      // unlit rows are never written by the split march
        {   // all loads of the row first (one exposed DRAM round trip instead of one per 32-byte piece), then the stores
            constexpr int MAXIT = 5;                       // k0 <= 144 + pad: at most 5 pieces of 8 columns per column group
            float4 lo[MAXIT], hi[MAXIT];
#pragma unroll
            for (int it = 0; it < MAXIT; ++it) {
                const int kc = cg + 4 * it;
                lo[it] = hi[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lit_row && kc * 8 < d.ta) {
                    lo[it] = __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * d.ta + kc * 8));
                    if (kc * 8 + 4 < d.ta) hi[it] = __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * d.ta + kc * 8 + 4));
                }
            }
#pragma unroll
            for (int it = 0; it < MAXIT; ++it) {
                const int kc = cg + 4 * it;
                if (kc < d.k0 / 8) {
                    const float v[8] = {lo[it].x, lo[it].y, lo[it].z, lo[it].w, hi[it].x, hi[it].y, hi[it].z, hi[it].w};
                    split_store8(s_ah, s_al, canon_off(row, kc * 8, d.k0), v);
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 1: feat = F . B^T
        if (tid == 0) { tc_fence_after(); issue_gemm3(tmem + COL0, sah, sal, sb0h, sb0l, d.k0, N0, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        {
            // column group cg encodes channels 8 cg .. 8 cg + 7 of this row (8 copies of the encoding code, not 32:
            // the unrolled 32-channel form was 131 KB of SASS and stalled on instruction fetch)
            float v[8];
            tmem_ld8(lane_base + COL0 + 8 * cg, v);
#pragma unroll PE_UNROLL
            for (int e = 0; e < 8; ++e) {
                const int ch = 8 * cg + e;
                if (ch >= nbase) continue;
                const bool is_feat = ch < d.app_dim;
                float x = v[e];
                if (!is_feat) x = live ? __ldg(a.rays + r * a.ray_stride + 3 + (ch - d.app_dim)) : 0.f;
                split_store1(s_ah, s_al, canon_off(row, ch, d.k1), x);
                const int nf = is_feat ? d.fea_pe : d.view_pe;
                const int cc = is_feat ? ch : ch - d.app_dim;
                const int sb = is_feat ? sin_f : sin_v, cb = is_feat ? cos_f : cos_v;
                float sn = 0.f, cs = 1.f;
                if (nf > 0) sincosf(x, &sn, &cs);
                for (int j = 0; j < nf; ++j) {          // higher octaves by angle doubling (<= 2^(nf-1) ulp)
                    split_store1(s_ah, s_al, canon_off(row, sb + cc * nf + j, d.k1), sn);
                    split_store1(s_ah, s_al, canon_off(row, cb + cc * nf + j, d.k1), cs);
                    const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
                    sn = s2; cs = c2;
                }
            }
            if (cg == 3)
                for (int k = d.in_c; k < d.k1; ++k) split_store1(s_ah, s_al, canon_off(row, k, d.k1), 0.f);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 2: h1 = X . W1^T
        if (tid == 0) {
            mbar_wait(wbar, wphase);                       // W1 has landed
            tc_fence_after();
            issue_gemm3(tmem + COL1, sah, sal, swh, swl, d.k1, FC, bar);
        }
        wphase ^= 1;
        w1_in_flight = false;
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        // W1 and X are consumed: W2 streams into the weight buffer under the epilogue
        if (tid == 0) bulk_load_pair(swh, swl, a.wimg + L.w2, L.sz_w2, wbar);
        {
            const int cb = cg * 32;
            float v[32];
            tmem_ld32(lane_base + COL1 + cb, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float u[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) u[e] = fmaxf(v[q * 8 + e] + s_bias[cb + q * 8 + e], 0.f);
                split_store8(s_ah, s_al, canon_off(row, cb + q * 8, FC), u);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 3: h2 = h1 . W2^T
        if (tid == 0) {
            mbar_wait(wbar, wphase);                       // W2 has landed
            tc_fence_after();
            issue_gemm3(tmem + COL2, sah, sal, swh, swl, FC, FC, bar);
        }
        wphase ^= 1;
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        // W2 is consumed: W1 of this CTA's next tile streams in under the remaining epilogues and the next staging
        if (tile + gridDim.x < n_tiles) {
            if (tid == 0) bulk_load_pair(swh, swl, a.wimg + L.w1, L.sz_w1, wbar);
            w1_in_flight = true;
        }
        {
            const int cb = cg * 32;
            float v[32];
            tmem_ld32(lane_base + COL2 + cb, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float u[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) u[e] = fmaxf(v[q * 8 + e] + s_bias[FC + cb + q * 8 + e], 0.f);
                split_store8(s_ah, s_al, canon_off(row, cb + q * 8, FC), u);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 4: rgb_raw = h2 . W3^T (N padded to 16)
        if (tid == 0) { tc_fence_after(); issue_gemm3(tmem + COL3, sah, sal, sw3h, sw3l, FC, N3, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        if (cg == 0) {
            float v[32];
            tmem_ld32(lane_base + COL3, v);
            if (live) {
                const bool lit = __ldg(a.app_count + r) > 0;
                const float ac = __ldg(a.acc + r);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float col = lit ? 1.f / (1.f + expf(-(v[c] + s_bias[2 * FC + c]))) : 0.f;
                    const float out = col * ac + __ldg(a.bg + c) * (1.f - ac);
                    tvm_put_rgb(a.peers, a.rgb, r, c, fminf(fmaxf(out, 0.f), 1.f));
                }
                const float last = __ldg(a.rays + r * a.ray_stride + a.ray_stride - 1);
                tvm_put_depth(a.peers, a.depth_out, r, __ldg(a.depth + r) + (1.f - ac) * last);
                if (a.acc_out) a.acc_out[r] = ac;
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (w1_in_flight && tid == 0) mbar_wait(wbar, wphase);     // never leave with a bulk copy still writing into our smem
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------------------------
// K-chunked variant for wide heads (the reference's DEFAULT fea_pe = view_pe = 6, opt.py:131-133: in_mlpC = 390,
// models/tensorBase.py:165-183).  The split X operand (128 x 400 x 2 B x hi|lo = 205 KB) and W1 (another 205 KB)
// cannot both sit in shared memory, so the first layer runs as K-chunks of TC3K_KC columns accumulating into the same
// TMEM columns: the encoding of chunk c+1 is produced while the tensor core consumes chunk c (two X buffers, two W1
// buffers refilled by TMA bulk copies the moment the MMA that read them completes).  X columns are PERMUTED so that the
// sin and the cos of one argument are neighbours (one sincosf per pair; W1's columns are packed with the same
// permutation by tvm_pack_mlp_tc3): [feat | viewdir | pad to even | (sin, cos) pairs of feat x 2^j | pairs of viewdir].
// ------------------------------------------------------------------------------------------------------------------
// permuted column k' -> column of the reference's MLP input (models/tensorBase.py:186-191), or -1 for padding
__host__ __device__ inline int tc3k_source_column(const TcDims& d, const Tc3kDims& c, int kp) {
    const int nbase = d.app_dim + 3;
    if (kp < nbase) return kp;
    if (kp < c.pbase) return -1;
    const int q = (kp - c.pbase) >> 1, is_cos = (kp - c.pbase) & 1;
    const int sin_f = nbase, cos_f = sin_f + d.app_dim * d.fea_pe;
    const int sin_v = cos_f + d.app_dim * d.fea_pe, cos_v = sin_v + 3 * d.view_pe;
    if (q < c.n_fpairs) return (is_cos ? cos_f : sin_f) + q;                   // q = ch * fea_pe + j already
    const int qv = q - c.n_fpairs;
    if (qv < c.n_vpairs) return (is_cos ? cos_v : sin_v) + qv;
    return -1;
}

// W1 [FC][in_c] -> n_chunks x (hi | lo) canonical images of [FC][TC3K_KC], columns permuted
static __global__ void pack_w1_chunked_kernel(const float* __restrict__ w1, TcDims d, Tc3kDims c, unsigned char* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= FC * c.k1c) return;
    const int r = i / c.k1c, kp = i - r * c.k1c;
    const int src = tc3k_source_column(d, c, kp);
    const float v = src >= 0 ? __ldg(w1 + r * d.in_c + src) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const int chunk = kp / TC3K_KC, kk = kp - chunk * TC3K_KC;
    const int half = FC * TC3K_KC * 2;
    unsigned char* base = dst + (size_t)chunk * 2 * half;
    *reinterpret_cast<__nv_bfloat16*>(base + canon_off(r, kk, TC3K_KC)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + half + canon_off(r, kk, TC3K_KC)) = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ void split_store2(unsigned char* hi_base, unsigned char* lo_base, int off, float v0, float v1) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    const __nv_bfloat162 hp = __halves2bfloat162(h0, h1);
    *reinterpret_cast<uint32_t*>(hi_base + off) = *reinterpret_cast<const uint32_t*>(&hp);
    *reinterpret_cast<uint32_t*>(lo_base + off) = pack_bf16x2(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
}

struct Tc3kSmem { int b0, w3, w, a, feat, bias, bar, total; };
__host__ __device__ inline Tc3kSmem tc3k_smem(const TcDims& d) {
    const Tc3Layout L = tc3_layout(d);
    Tc3kSmem m;
    const int chunk_pair = 2 * FC * TC3K_KC * 2;                 // W1 chunk (hi|lo)  ==  X chunk (hi|lo): 128 rows either way
    const int w_bytes = 2 * chunk_pair > 2 * L.sz_w2 ? 2 * chunk_pair : 2 * L.sz_w2;
    const int a0 = 2 * TC_RAYS * d.k0 * 2, ah = 2 * TC_RAYS * FC * 2;
    int a_bytes = 2 * chunk_pair;
    if (a0 > a_bytes) a_bytes = a0;
    if (ah > a_bytes) a_bytes = ah;
    m.b0 = 0;
    m.w3 = m.b0 + 2 * L.sz_b0;
    m.w = m.w3 + 2 * L.sz_w3;
    m.a = m.w + w_bytes;
    m.feat = m.a + a_bytes;
    m.bias = m.feat + TC_RAYS * 32 * 4;
    m.bar = m.bias + (2 * FC + 4) * 4;
    m.total = m.bar + 64;
    return m;
}

__global__ void __launch_bounds__(TC_THREADS, 1) shade_tc3k_kernel(const __grid_constant__ Tc3Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const TcDims& d = a.d;
    const Tc3Layout L = tc3_layout(d);
    const Tc3kDims kd = tc3k_dims(d);
    const Tc3kSmem M = tc3k_smem(d);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, cg = warp >> 2;
    constexpr int KC = TC3K_KC, CHUNK_HALF = FC * KC * 2, CHUNK_PAIR = 2 * CHUNK_HALF;
    unsigned char* s_b0 = smem + M.b0;
    unsigned char* s_w3 = smem + M.w3;
    unsigned char* s_w = smem + M.w;            // W1 chunk buffers [2][hi|lo]   /   W2 hi|lo
    unsigned char* s_a = smem + M.a;            // ray_feat hi|lo  /  X chunk buffers [2][hi|lo]  /  h hi|lo
    float* s_feat = reinterpret_cast<float*>(smem + M.feat);       // [128][32]: feat (app_dim) | viewdir (3)
    float* s_bias = reinterpret_cast<float*>(smem + M.bias);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + M.bar);   // 0,1: MMA completion (alternating)  2,3: W1 chunk
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 6);     // buffers  4: W2 arrival

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(s_tmem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(s_bar + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    coop_copy(s_b0, a.wimg + L.b0, 2 * L.sz_b0, tid);
    coop_copy(s_w3, a.wimg + L.w3, 2 * L.sz_w3, tid);
    for (int i = tid; i < 2 * FC + 4; i += TC_THREADS)
        s_bias[i] = i < FC ? __ldg(a.b1 + i) : (i < 2 * FC ? __ldg(a.b2 + i - FC) : (i - 2 * FC < 3 ? __ldg(a.b3 + i - 2 * FC) : 0.f));
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t mbar[2] = {smem_u32(s_bar), smem_u32(s_bar + 1)};
    const uint32_t wbar[2] = {smem_u32(s_bar + 2), smem_u32(s_bar + 3)};
    const uint32_t w2bar = smem_u32(s_bar + 4);
    uint32_t mph[2] = {0, 0}, wph[2] = {0, 0}, w2ph = 0;
    const uint32_t sa = smem_u32(s_a), sw = smem_u32(s_w);
    const uint32_t sb0h = smem_u32(s_b0), sb0l = smem_u32(s_b0 + L.sz_b0);
    const uint32_t sw3h = smem_u32(s_w3), sw3l = smem_u32(s_w3 + L.sz_w3);
    const int a0_half = TC_RAYS * d.k0 * 2, ah_half = TC_RAYS * FC * 2;
    const long long n_tiles = (a.n_rays + TC_RAYS - 1) / TC_RAYS;
    const unsigned char* w1img = a.wimg + L.w1;

    // issue the W1 chunk `c` copy into chunk buffer c & 1 (one thread)
    auto load_w1_chunk = [&](int c) {
        bulk_load_pair(sw + (c & 1) * CHUNK_PAIR, sw + (c & 1) * CHUNK_PAIR + CHUNK_HALF, w1img + (size_t)c * CHUNK_PAIR,
                       CHUNK_HALF, wbar[c & 1]);
    };
    bool w1_armed = false;                     // uniform: chunks 0 and 1 of the next shaded tile are in flight

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r = tile * TC_RAYS + row;
        const bool live = r < a.n_rays;
        if (!__syncthreads_or(live && __ldg(a.app_count + r) > 0)) {
            if (cg == 0 && live) {
                const float ac = __ldg(a.acc + r);
#pragma unroll
                for (int c = 0; c < 3; ++c) tvm_put_rgb(a.peers, a.rgb, r, c, fminf(fmaxf(__ldg(a.bg + c) * (1.f - ac), 0.f), 1.f));
                const float last = __ldg(a.rays + r * a.ray_stride + a.ray_stride - 1);
                tvm_put_depth(a.peers, a.depth_out, r, __ldg(a.depth + r) + (1.f - ac) * last);
                if (a.acc_out) a.acc_out[r] = ac;
            }
            continue;
        }
        if (!w1_armed) {
            if (tid == 0) { load_w1_chunk(0); if (kd.n_chunks > 1) load_w1_chunk(1); }
            w1_armed = true;
        }
        // ---- stage ray_feat as split A operand (K0)
        const bool lit_row = live && __ldg(a.app_count + r) > 0;
        {   // all loads of the row first (one exposed DRAM round trip instead of one per 32-byte piece), then the stores
            constexpr int MAXIT = 5;                       // k0 <= 144 + pad: at most 5 pieces of 8 columns per column group
            float4 lo[MAXIT], hi[MAXIT];
#pragma unroll
            for (int it = 0; it < MAXIT; ++it) {
                const int kc = cg + 4 * it;
                lo[it] = hi[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lit_row && kc * 8 < d.ta) {
                    lo[it] = __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * d.ta + kc * 8));
                    if (kc * 8 + 4 < d.ta) hi[it] = __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * d.ta + kc * 8 + 4));
                }
            }
#pragma unroll
            for (int it = 0; it < MAXIT; ++it) {
                const int kc = cg + 4 * it;
                if (kc < d.k0 / 8) {
                    const float v[8] = {lo[it].x, lo[it].y, lo[it].z, lo[it].w, hi[it].x, hi[it].y, hi[it].z, hi[it].w};
                    split_store8(s_a, s_a + a0_half, canon_off(row, kc * 8, d.k0), v);
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 1: feat = F . B^T
        if (tid == 0) { tc_fence_after(); issue_gemm3(tmem + COL0, sa, sa + a0_half, sb0h, sb0l, d.k0, N0, mbar[0]); }
        mbar_wait(mbar[0], mph[0]); mph[0] ^= 1;
        tc_fence_after();
        {   // feat / viewdir of this row -> shared memory: every column of X is a function of these 30 values
            float v[8];
            tmem_ld8(lane_base + COL0 + 8 * cg, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ch = 8 * cg + e;
                float x = v[e];
                if (ch >= d.app_dim) x = (ch < d.app_dim + 3 && live) ? __ldg(a.rays + r * a.ray_stride + 3 + (ch - d.app_dim)) : 0.f;
                s_feat[row * 32 + ch] = x;
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- MMA 2 in K-chunks: h1 = X . W1^T; encode chunk c while the tensor core works on chunk c - 1
        for (int c = 0; c < kd.n_chunks; ++c) {
            const int buf = c & 1;
            if (c >= 2) {                      // the MMA of chunk c - 2 read X / W1 buffer `buf`: wait, then refill W1
                mbar_wait(mbar[buf], mph[buf]); mph[buf] ^= 1;
                tc_fence_after();
                if (tid == 0) load_w1_chunk(c);
            }
            unsigned char* xh = s_a + buf * CHUNK_PAIR;
            unsigned char* xl = xh + CHUNK_HALF;
            const int k0p = c * KC + cg * (KC / 4);
            // a thread's pairs are consecutive frequencies of one channel most of the time: the next octave comes from
            // the angle-doubling identities (sin 2a = 2 sin a cos a, cos 2a = cos^2 a - sin^2 a; <= 2^5 ulp after the
            // five doublings of fea_pe = 6, i.e. ~2e-6) and sincosf is evaluated only at a thread's first pair or a
            // channel's first frequency
            int prev_ch = -1, prev_j = -1;
            float sn = 0.f, cs = 0.f;
#pragma unroll 2
            for (int e = 0; e < KC / 4; e += 2) {
                const int kp = k0p + e;
                float v0 = 0.f, v1 = 0.f;
                if (kp < kd.pbase) {
                    v0 = kp < d.app_dim + 3 ? s_feat[row * 32 + kp] : 0.f;
                    v1 = kp + 1 < d.app_dim + 3 ? s_feat[row * 32 + kp + 1] : 0.f;
                } else {
                    const int q = (kp - kd.pbase) >> 1;
                    int ch = -1, j = 0;
                    if (q < kd.n_fpairs) { ch = q / d.fea_pe; j = q - ch * d.fea_pe; }
                    else if (q - kd.n_fpairs < kd.n_vpairs) { const int qv = q - kd.n_fpairs; ch = qv / d.view_pe; j = qv - ch * d.view_pe; ch += d.app_dim; }
                    if (ch >= 0) {
                        if (ch == prev_ch && j == prev_j + 1) {
                            const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
                            sn = s2; cs = c2;
                        } else {
                            sincosf(s_feat[row * 32 + ch] * (float)(1 << j), &sn, &cs);
                        }
                        prev_ch = ch; prev_j = j;
                        v0 = sn; v1 = cs;
                    }
                }
                split_store2(xh, xl, canon_off(row, kp - c * KC, KC), v0, v1);
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                mbar_wait(wbar[buf], wph[buf]);               // W1 chunk c has landed
                tc_fence_after();
                const uint64_t dah = make_desc(smem_u32(xh), KC), dal = make_desc(smem_u32(xl), KC);
                const uint64_t dbh = make_desc(sw + buf * CHUNK_PAIR, KC), dbl = make_desc(sw + buf * CHUNK_PAIR + CHUNK_HALF, KC);
                const uint32_t idesc = make_idesc(128, FC);
                for (int j = 0; j < KC / 16; ++j) {
                    const uint64_t o = (uint64_t)(j * 16);
                    umma_bf16(tmem + COL1, dah + o, dbh + o, idesc, (c > 0 || j > 0) ? 1u : 0u);
                    umma_bf16(tmem + COL1, dah + o, dbl + o, idesc, 1u);
                    umma_bf16(tmem + COL1, dal + o, dbh + o, idesc, 1u);
                }
                umma_commit(mbar[buf]);
            }
            wph[buf] ^= 1;
        }
        // drain: the last one or two chunk MMAs (in issue order)
        if (kd.n_chunks >= 2) { const int b = (kd.n_chunks - 2) & 1; mbar_wait(mbar[b], mph[b]); mph[b] ^= 1; }
        { const int b = (kd.n_chunks - 1) & 1; mbar_wait(mbar[b], mph[b]); mph[b] ^= 1; }
        tc_fence_after();
        w1_armed = false;
        // W1 and X are consumed: W2 streams into the weight region under the epilogue
        if (tid == 0) bulk_load_pair(sw, sw + L.sz_w2, a.wimg + L.w2, L.sz_w2, w2bar);
        {
            const int cb = cg * 32;
            float v[32];
            tmem_ld32(lane_base + COL1 + cb, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float u[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) u[e] = fmaxf(v[q * 8 + e] + s_bias[cb + q * 8 + e], 0.f);
                split_store8(s_a, s_a + ah_half, canon_off(row, cb + q * 8, FC), u);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 3: h2 = h1 . W2^T
        if (tid == 0) {
            mbar_wait(w2bar, w2ph);
            tc_fence_after();
            issue_gemm3(tmem + COL2, sa, sa + ah_half, sw, sw + L.sz_w2, FC, FC, mbar[0]);
        }
        w2ph ^= 1;
        mbar_wait(mbar[0], mph[0]); mph[0] ^= 1;
        tc_fence_after();
        // W2 is consumed: the first W1 chunks of this CTA's next tile stream in under the remaining epilogues
        if (tile + gridDim.x < n_tiles) {
            if (tid == 0) { load_w1_chunk(0); if (kd.n_chunks > 1) load_w1_chunk(1); }
            w1_armed = true;
        }
        {
            const int cb = cg * 32;
            float v[32];
            tmem_ld32(lane_base + COL2 + cb, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float u[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) u[e] = fmaxf(v[q * 8 + e] + s_bias[FC + cb + q * 8 + e], 0.f);
                split_store8(s_a, s_a + ah_half, canon_off(row, cb + q * 8, FC), u);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 4: rgb_raw = h2 . W3^T (N padded to 16)
        if (tid == 0) { tc_fence_after(); issue_gemm3(tmem + COL3, sa, sa + ah_half, sw3h, sw3l, FC, N3, mbar[0]); }
        mbar_wait(mbar[0], mph[0]); mph[0] ^= 1;
        tc_fence_after();
        if (cg == 0) {
            float v[32];
            tmem_ld32(lane_base + COL3, v);
            if (live) {
                const bool lit = __ldg(a.app_count + r) > 0;
                const float ac = __ldg(a.acc + r);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float col = lit ? 1.f / (1.f + expf(-(v[c] + s_bias[2 * FC + c]))) : 0.f;
                    const float out = col * ac + __ldg(a.bg + c) * (1.f - ac);
                    tvm_put_rgb(a.peers, a.rgb, r, c, fminf(fmaxf(out, 0.f), 1.f));
                }
                const float last = __ldg(a.rays + r * a.ray_stride + a.ray_stride - 1);
                tvm_put_depth(a.peers, a.depth_out, r, __ldg(a.depth + r) + (1.f - ac) * last);
                if (a.acc_out) a.acc_out[r] = ac;
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (w1_armed && tid == 0) {                // never leave with bulk copies still writing into our shared memory
        mbar_wait(wbar[0], wph[0]);
        if (kd.n_chunks > 1) mbar_wait(wbar[1], wph[1]);
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

// which kernel serves this head: 0 = none, 1 = whole-K kernel, 2 = K-chunked kernel
int tc3_variant(const tvm_field_desc* desc) {
    if (!desc || desc->feature_c != FC || desc->app_dim + 3 > N0 || desc->app_dim <= 0) return 0;
    const TcDims d = tc_dims(desc);
    if (tc3_whole_k_fits(d)) return 1;
    if (tc3k_smem(d).total <= 227 * 1024 && desc->fea_pe >= 0 && desc->view_pe >= 0) return 2;
    return 0;
}

size_t tc3_smem_bytes(const TcDims& d) {
    const Tc3Layout L = tc3_layout(d);
    const int kmax = d.k0 > d.k1 ? (d.k0 > FC ? d.k0 : FC) : (d.k1 > FC ? d.k1 : FC);
    const size_t a_half = (size_t)TC_RAYS * kmax * 2;
    const size_t w_half = (size_t)FC * (d.k1 > FC ? d.k1 : FC) * 2;
    return 2 * (size_t)L.sz_b0 + 2 * (size_t)L.sz_w3 + 2 * w_half + 2 * a_half + (2 * FC + 4) * 4 + 32;
}

}  // namespace

extern "C" int tvm_mlp_tc3_supported(const tvm_field_desc* desc) { return tc3_variant(desc) != 0 ? 1 : 0; }

extern "C" size_t tvm_mlp_tc3_pack_bytes(const tvm_field_desc* desc) {
    if (!desc) return 0;
    return (size_t)tc3_layout(tc_dims(desc)).total;
}

extern "C" int tvm_pack_mlp_tc3(const tvm_field_desc* desc, const float* basis, const float* w1, const float* w2,
                                const float* w3, void* packed, void* stream) {
    if (!desc || !basis || !w1 || !w2 || !w3 || !packed) return TVM_E_NULL;
    const TcDims d = tc_dims(desc);
    const int variant = tc3_variant(desc);
    if (variant == 0) return TVM_E_SHAPE;
    const Tc3Layout L = tc3_layout(d);
    unsigned char* out = (unsigned char*)packed;
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == 2) {            // K-chunked kernel: W1 as n_chunks x (hi | lo) images with permuted columns
        const Tc3kDims kd = tc3k_dims(d);
        tvm_count_launch(); pack_w1_chunked_kernel<<<(FC * kd.k1c + 255) / 256, 256, 0, st>>>(w1, d, kd, out + L.w1);
    }
    for (int part = 0; part < 2; ++part) {
        tvm_count_launch(); pack_bf16_operand_kernel<<<(N0 * d.k0 + 255) / 256, 256, 0, st>>>(basis, d.app_dim, d.ta, N0, d.k0,
                                                                         out + L.b0 + part * L.sz_b0, part);
        if (variant == 1) {
            tvm_count_launch(); pack_bf16_operand_kernel<<<(FC * d.k1 + 255) / 256, 256, 0, st>>>(w1, FC, d.in_c, FC, d.k1,
                                                                             out + L.w1 + part * L.sz_w1, part);
        }
        tvm_count_launch(); pack_bf16_operand_kernel<<<(FC * FC + 255) / 256, 256, 0, st>>>(w2, FC, FC, FC, FC,
                                                                       out + L.w2 + part * L.sz_w2, part);
        tvm_count_launch(); pack_bf16_operand_kernel<<<(N3 * FC + 255) / 256, 256, 0, st>>>(w3, 3, FC, N3, FC,
                                                                       out + L.w3 + part * L.sz_w3, part);
    }
    TVM_LAUNCH_CHECK();
    return 0;
}

// called by tvm_shade_fwd when TVM_F_MLP_TC3 is set
int tvm_shade_tc3_launch(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                         const float* bg, float* rgb, float* depth, float* acc, const void* ws, size_t ws_bytes,
                         cudaStream_t st, const tvm_scatter_out* sc) {
    if (!desc->mlp_tc3 || !desc->mlp) return TVM_E_NULL;
    if (desc->feature_c != FC || desc->app_dim + 3 > N0 || desc->app_dim <= 0) return TVM_E_SHAPE;
    const TcDims d = tc_dims(desc);
    const int variant = tc3_variant(desc);
    if (variant == 0) return TVM_E_SHAPE;
    const size_t smem = variant == 1 ? tc3_smem_bytes(d) : (size_t)tc3k_smem(d).total;
    const TvmWorkspace w = tvm_ws_layout(desc, n_rays);
    if (ws_bytes < w.total) return TVM_E_WORKSPACE;
    const TvmMlpLayout m = tvm_mlp_layout(desc);
    const char* base = (const char*)ws;
    Tc3Args a{};
    a.rays = rays; a.n_rays = n_rays; a.ray_stride = ray_stride; a.bg = bg;
    a.rgb = rgb; a.depth_out = depth; a.acc_out = acc;
    { int rc_p = tvm_fill_peers(a.peers, sc); if (rc_p) return rc_p; }
    a.ray_feat = (const float*)(base + w.ray_feat);
    a.acc = (const float*)(base + w.acc);
    a.depth = (const float*)(base + w.depth);
    a.app_count = (const int*)(base + w.app_count);
    a.wimg = (const unsigned char*)desc->mlp_tc3;
    a.b1 = desc->mlp + m.b1; a.b2 = desc->mlp + m.b2; a.b3 = desc->mlp + m.b3;
    a.d = d;
    const long long tiles = (n_rays + TC_RAYS - 1) / TC_RAYS;
    const unsigned grid = (unsigned)(tiles < TVM_SM_COUNT ? tiles : TVM_SM_COUNT);
    if (variant == 1) {
        static TvmDevMemo smem_set;
        int rc_attr = tvm_ensure_dyn_smem(shade_tc3_kernel, smem, smem_set);
        if (rc_attr) return rc_attr;
        tvm_count_launch(); shade_tc3_kernel<<<grid, TC_THREADS, smem, st>>>(a);
    } else {
        static TvmDevMemo smem_set_k;
        int rc_attr = tvm_ensure_dyn_smem(shade_tc3k_kernel, smem, smem_set_k);
        if (rc_attr) return rc_attr;
        tvm_count_launch(); shade_tc3k_kernel<<<grid, TC_THREADS, smem, st>>>(a);
    }
    TVM_LAUNCH_CHECK();
    return 0;
}
