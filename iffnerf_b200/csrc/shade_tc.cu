// shade_tc.cu — tensor-core variant of the per-ray shading stage (flag TVM_F_MLP_BF16).
//
// Same math as shade.cu — basis_mat (models/tensoRF.py:158,256), MLPRender_Fea (models/tensorBase.py:165-195),
// background blend and depth tail (:898-908) — but the four dense contractions
//     feat = F[128x144] . B^T      h1 = relu(X[128x160] . W1^T + b1)
//     h2   = relu(h1 . W2^T + b2)  rgb = sigmoid(h2 . W3^T + b3)
// run on the 5th-gen tensor cores: bf16 operands in shared memory (K-major, no-swizzle canonical core-matrix
// layout), fp32 accumulators in TMEM, `tcgen05.mma.cta_group::1.kind::f16` issued by one thread, completion
// signalled through an mbarrier by `tcgen05.commit`, accumulators read back with `tcgen05.ld`.  One persistent
// CTA per SM owns a 128-ray tile at a time (TMEM lane == ray), keeps all four weight images resident in shared
// memory, and its 512 threads (4 warps per TMEM lane quarter) do the epilogues (positional encoding, bias, ReLU, bf16 repack) between MMAs.
// bf16 operands bound the result to the 1e-2 tolerance of BASELINE.json's "bf16 MLP mode"; the fp32 SIMT
// kernel in shade.cu stays the default (1e-4).
#include "tvm_tc.cuh"

namespace {
using namespace tvmtc;

struct TcArgs {
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
    const unsigned char* wimg;     // bf16 weight images (tvm_pack_mlp_tc)
    const float* b1; const float* b2; const float* b3;
    TcDims d;
};

// one MMA chain: D[128 x N] (TMEM col base) = A[128 x K] . B[N x K]^T, issued by the calling thread
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_saddr, uint32_t b_saddr, int K, int N,
                                           uint32_t bar) {
    const uint64_t da = make_desc(a_saddr, K), db = make_desc(b_saddr, K);
    const uint32_t idesc = make_idesc(128, N);
    for (int j = 0; j < K / 16; ++j)          // one UMMA_K = 16 step = 2 core matrices = 256 B (>>4 = 16) further along K
        umma_bf16(tmem_d, da + (uint64_t)(j * 16), db + (uint64_t)(j * 16), idesc, j > 0 ? 1u : 0u);
    umma_commit(bar);                          // implies tcgen05.fence::before_thread_sync
}

__global__ void __launch_bounds__(TC_THREADS, 1) shade_tc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const TcDims& d = a.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 16 warps: warp w works on TMEM lane quarter (w & 3) — the only lanes it may read — i.e. on tile row
    // `row`, and on column group `cg` (a quarter of every epilogue's columns / channels)
    const int row = (warp & 3) * 32 + lane, cg = warp >> 2;
    // shared memory carve-up
    unsigned char* s_img = smem;                                        // weight images
    unsigned char* s_a = s_img + ((d.img_bytes + 127) & ~127);          // F tile (K0) / h1 (K=128)
    const int a_bytes = TC_RAYS * (d.k0 > FC ? d.k0 : FC) * 2;
    unsigned char* s_x = s_a + a_bytes;                                 // MLP input (K1) / h2 (K=128)
    const int x_bytes = TC_RAYS * (d.k1 > FC ? d.k1 : FC) * 2;
    float* s_bias = reinterpret_cast<float*>(s_x + x_bytes);            // b1[128] b2[128] b3[4]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 2 * FC + 4);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(s_tmem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(s_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid * 16; i < d.img_bytes; i += TC_THREADS * 16)
        *reinterpret_cast<uint4*>(s_img + i) = __ldg(reinterpret_cast<const uint4*>(a.wimg + i));
    for (int i = tid; i < 2 * FC + 4; i += TC_THREADS)
        s_bias[i] = i < FC ? __ldg(a.b1 + i) : (i < 2 * FC ? __ldg(a.b2 + i - FC) : (i - 2 * FC < 3 ? __ldg(a.b3 + i - 2 * FC) : 0.f));
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's TMEM lane quarter
    const uint32_t bar = smem_u32(s_bar);
    const uint32_t sa = smem_u32(s_a), sx = smem_u32(s_x), simg = smem_u32(s_img);
    uint32_t phase = 0;
    const long long n_tiles = (a.n_rays + TC_RAYS - 1) / TC_RAYS;
    const int nbase = d.app_dim + 3;
    const int sin_f = nbase, cos_f = sin_f + d.app_dim * d.fea_pe;
    const int sin_v = cos_f + d.app_dim * d.fea_pe, cos_v = sin_v + 3 * d.view_pe;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r = tile * TC_RAYS + row;          // this thread's ray == its TMEM lane
        const bool live = r < a.n_rays;
        // tiles without a single appearance sample need no GEMM (the reference skips the MLP for such rays too,
        // tensorBase.py:876-896): rgb = bg * (1 - acc)
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
        // ---- stage 0: ray_feat row -> bf16 A operand (K0 columns); the 4 column groups interleave the 16-B chunks
        const bool lit_row = live && __ldg(a.app_count + r) > 0;      // unlit rows are never written by the split march
        for (int kc = cg; kc < d.k0 / 8; kc += 4) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
            if (lit_row && kc * 8 < d.ta) {                   // ta is a multiple of 4
                const float4 lo = __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * d.ta + kc * 8));
                v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
                if (kc * 8 + 4 < d.ta) {
                    const float4 hi = __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * d.ta + kc * 8 + 4));
                    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
                }
            }
            *reinterpret_cast<uint4*>(s_a + canon_off(row, kc * 8, d.k0)) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 1: feat = F . B^T
        if (tid == 0) { tc_fence_after(); issue_gemm(tmem + COL0, sa, simg + d.img0, d.k0, N0, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        {
            // MLP input row (tensorBase.py:186-191): [feat | view | sin(feat 2^j) | cos | sin(view 2^j) | cos], zero pad.
            // Column group cg encodes channels 8 cg .. 8 cg + 7 of its lanes' rows (8 copies of the encoding code
            // instead of 32: instruction-cache footprint).
            float v[8];
            tmem_ld8(lane_base + COL0 + 8 * cg, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ch = 8 * cg + e;
                if (ch >= nbase) continue;
                const bool is_feat = ch < d.app_dim;
                float x = v[e];
                if (!is_feat) x = live ? __ldg(a.rays + r * a.ray_stride + 3 + (ch - d.app_dim)) : 0.f;
                *reinterpret_cast<__nv_bfloat16*>(s_x + canon_off(row, ch, d.k1)) = __float2bfloat16_rn(x);
                const int nf = is_feat ? d.fea_pe : d.view_pe;
                const int cc = is_feat ? ch : ch - d.app_dim;
                const int sb = is_feat ? sin_f : sin_v, cb = is_feat ? cos_f : cos_v;
                float scale = 1.f;
                for (int j = 0; j < nf; ++j) {
                    float sn, cs;
                    sincosf(x * scale, &sn, &cs);
                    *reinterpret_cast<__nv_bfloat16*>(s_x + canon_off(row, sb + cc * nf + j, d.k1)) = __float2bfloat16_rn(sn);
                    *reinterpret_cast<__nv_bfloat16*>(s_x + canon_off(row, cb + cc * nf + j, d.k1)) = __float2bfloat16_rn(cs);
                    scale *= 2.f;
                }
            }
            if (cg == 3)
                for (int k = d.in_c; k < d.k1; ++k)
                    *reinterpret_cast<__nv_bfloat16*>(s_x + canon_off(row, k, d.k1)) = __float2bfloat16_rn(0.f);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 2: h1 = X . W1^T
        if (tid == 0) { tc_fence_after(); issue_gemm(tmem + COL1, sx, simg + d.img1, d.k1, FC, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        {                                                   // bias + ReLU -> bf16 A operand (K = 128) in s_a
            const int cb = cg * 32;
            float v[32];
            tmem_ld32(lane_base + COL1 + cb, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = q * 8 + e * 2;
                    w[e] = pack_bf16x2(fmaxf(v[j] + s_bias[cb + j], 0.f), fmaxf(v[j + 1] + s_bias[cb + j + 1], 0.f));
                }
                *reinterpret_cast<uint4*>(s_a + canon_off(row, cb + q * 8, FC)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 3: h2 = h1 . W2^T
        if (tid == 0) { tc_fence_after(); issue_gemm(tmem + COL2, sa, simg + d.img2, FC, FC, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        {                                                   // bias + ReLU -> bf16 A operand (K = 128) in s_x
            const int cb = cg * 32;
            float v[32];
            tmem_ld32(lane_base + COL2 + cb, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = q * 8 + e * 2;
                    w[e] = pack_bf16x2(fmaxf(v[j] + s_bias[FC + cb + j], 0.f), fmaxf(v[j + 1] + s_bias[FC + cb + j + 1], 0.f));
                }
                *reinterpret_cast<uint4*>(s_x + canon_off(row, cb + q * 8, FC)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 4: rgb_raw = h2 . W3^T (N padded to 16)
        if (tid == 0) { tc_fence_after(); issue_gemm(tmem + COL3, sx, simg + d.img3, FC, N3, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        if (cg == 0) {
            float v[32];
            tmem_ld32(lane_base + COL3, v);               // columns 3.. are padding / neighbouring accumulators
            if (live) {
                const bool lit = __ldg(a.app_count + r) > 0;          // rays_to_consider (tensorBase.py:886)
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
        __syncthreads();                                   // TMEM + smem operands free for the next tile
        tc_fence_after();
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

size_t tc_smem_bytes(const TcDims& d) {
    const size_t a_bytes = (size_t)TC_RAYS * (d.k0 > FC ? d.k0 : FC) * 2;
    const size_t x_bytes = (size_t)TC_RAYS * (d.k1 > FC ? d.k1 : FC) * 2;
    return ((d.img_bytes + 127) & ~127) + a_bytes + x_bytes + (2 * FC + 4) * 4 + 16;
}

}  // namespace

extern "C" size_t tvm_mlp_tc_pack_bytes(const tvm_field_desc* desc) {
    if (!desc) return 0;
    return (size_t)tc_dims(desc).img_bytes;
}

extern "C" int tvm_pack_mlp_tc(const tvm_field_desc* desc, const float* basis, const float* w1, const float* w2,
                               const float* w3, void* packed, void* stream) {
    if (!desc || !basis || !w1 || !w2 || !w3 || !packed) return TVM_E_NULL;
    if (desc->feature_c != FC || desc->app_dim + 3 > N0 || desc->app_dim <= 0) return TVM_E_SHAPE;
    const TcDims d = tc_dims(desc);
    if (tc_smem_bytes(d) > 227 * 1024) return TVM_E_SHAPE;
    unsigned char* out = (unsigned char*)packed;
    cudaStream_t st = (cudaStream_t)stream;
    tvm_count_launch(); pack_bf16_operand_kernel<<<(N0 * d.k0 + 255) / 256, 256, 0, st>>>(basis, d.app_dim, d.ta, N0, d.k0, out + d.img0, 0);
    tvm_count_launch(); pack_bf16_operand_kernel<<<(FC * d.k1 + 255) / 256, 256, 0, st>>>(w1, FC, d.in_c, FC, d.k1, out + d.img1, 0);
    tvm_count_launch(); pack_bf16_operand_kernel<<<(FC * FC + 255) / 256, 256, 0, st>>>(w2, FC, FC, FC, FC, out + d.img2, 0);
    tvm_count_launch(); pack_bf16_operand_kernel<<<(N3 * FC + 255) / 256, 256, 0, st>>>(w3, 3, FC, N3, FC, out + d.img3, 0);
    TVM_LAUNCH_CHECK();
    return 0;
}

// called by tvm_shade_fwd when TVM_F_MLP_BF16 is set
int tvm_shade_tc_launch(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride, const float* bg,
                        float* rgb, float* depth, float* acc, const void* ws, size_t ws_bytes, cudaStream_t st, const tvm_scatter_out* sc) {
    if (!desc->mlp_tc || !desc->mlp) return TVM_E_NULL;
    if (desc->feature_c != FC || desc->app_dim + 3 > N0 || desc->app_dim <= 0) return TVM_E_SHAPE;   // feat|view fit 32 cols
    const TcDims d = tc_dims(desc);
    const size_t smem = tc_smem_bytes(d);
    if (smem > 227 * 1024) return TVM_E_SHAPE;
    const TvmWorkspace w = tvm_ws_layout(desc, n_rays);
    if (ws_bytes < w.total) return TVM_E_WORKSPACE;
    const TvmMlpLayout m = tvm_mlp_layout(desc);
    const char* base = (const char*)ws;
    TcArgs a{};
    a.rays = rays; a.n_rays = n_rays; a.ray_stride = ray_stride; a.bg = bg;
    a.rgb = rgb; a.depth_out = depth; a.acc_out = acc;
    { int rc_p = tvm_fill_peers(a.peers, sc); if (rc_p) return rc_p; }
    a.ray_feat = (const float*)(base + w.ray_feat);
    a.acc = (const float*)(base + w.acc);
    a.depth = (const float*)(base + w.depth);
    a.app_count = (const int*)(base + w.app_count);
    a.wimg = (const unsigned char*)desc->mlp_tc;
    a.b1 = desc->mlp + m.b1; a.b2 = desc->mlp + m.b2; a.b3 = desc->mlp + m.b3;
    a.d = d;
    {
        static TvmDevMemo smem_set;
        int rc_attr = tvm_ensure_dyn_smem(shade_tc_kernel, smem, smem_set);
        if (rc_attr) return rc_attr;
    }
    const long long tiles = (n_rays + TC_RAYS - 1) / TC_RAYS;
    const unsigned grid = (unsigned)(tiles < TVM_SM_COUNT ? tiles : TVM_SM_COUNT);
    tvm_count_launch(); shade_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
    TVM_LAUNCH_CHECK();
    return 0;
}
