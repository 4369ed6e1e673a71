// shade.cu — per-ray shading stage (fp32 SIMT variant) + the tvm_render_fwd entry point.
//
// Replaces the reference's once-per-ray tail of TensorBase.forward:
//   basis_mat (models/tensoRF.py:158,256; bias-free Linear, hoisted past the weighted sum)
//   -> MLPRender_Fea.forward (models/tensorBase.py:185-195) with positional_encoding (:14-20)
//   -> rgb_map = clamp(rgb*acc + bg*(1-acc), 0, 1) (:898-904) and the depth tail (:906-908).
//
// One CTA shades a tile of 64 rays: the tile's ray_feat rows and basis_mat are staged in shared
// memory, the MLP input row (feat | viewdir | sin/cos encodings) is built in place, and the two
// 128-wide layers run as a register-tiled fp32 GEMM (8 rays x 4 outputs per thread) whose weight
// rows are read transposed ([K][128], one coalesced 512-B row per k, L1-resident across CTAs).
// fp32 FFMA keeps this variant inside the 1e-4 parity bound; the bf16 tensor-core variant is
// selected with TVM_F_MLP_BF16.
#include "tvm_common.cuh"

int tvm_march_fwd_launch(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride, int n_samples,
                         const float* jitter, uint32_t flags, float* alpha, float* z_vals, float* dists,
                         uint32_t* valid_bits, int32_t* valid_count, int32_t* app_count, void* ws, size_t ws_bytes,
                         cudaStream_t st);

int tvm_shade_tc_launch(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride, const float* bg,
                        float* rgb, float* depth, float* acc, const void* ws, size_t ws_bytes, cudaStream_t st, const tvm_scatter_out* sc);

int tvm_shade_tc3_launch(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                         const float* bg, float* rgb, float* depth, float* acc, const void* ws, size_t ws_bytes,
                         cudaStream_t st, const tvm_scatter_out* sc);

namespace {

constexpr int SH_RAYS = 64;
constexpr int SH_THREADS = 256;
constexpr int FC = TVM_FEATURE_C;
constexpr unsigned FULL = 0xffffffffu;

struct ShadeArgs {
    const float* rays;
    long long n_rays;
    int ray_stride;
    const float* bg;         // device [3]
    float* rgb;
    float* depth_out;
    float* acc_out;
    TvmPeers peers;
    const float* ray_feat;
    const float* acc;
    const float* depth;
    const int* app_count;
    const float* basis;      // [app_dim][ta]
    const float* w1t; const float* b1; const float* w2t; const float* b2; const float* w3; const float* b3;
    int ta, app_dim, fea_pe, view_pe, in_c, k1;
    int sF_stride;           // ta + 4
};

// acc[8][4] += X[8 rays][K] * Wt[K][4 cols]
template <int XS_IS_FC>
__device__ __forceinline__ void tile_gemm(float (&acc)[8][4], const float* __restrict__ sX, int x_stride, int K,
                                          const float* __restrict__ Wt, int tx, int ty) {
    const float4* W4 = reinterpret_cast<const float4*>(Wt) + tx;
    // register double-buffer of the weight rows: the loads for step k+4 are in flight while step k is computed
    float4 n0 = __ldg(W4), n1 = __ldg(W4 + FC / 4), n2 = __ldg(W4 + 2 * (FC / 4)), n3 = __ldg(W4 + 3 * (FC / 4));
    for (int k = 0; k < K; k += 4) {
        const float4 w0 = n0, w1 = n1, w2 = n2, w3 = n3;
        if (k + 4 < K) {
            n0 = __ldg(W4 + (k + 4) * (FC / 4));
            n1 = __ldg(W4 + (k + 5) * (FC / 4));
            n2 = __ldg(W4 + (k + 6) * (FC / 4));
            n3 = __ldg(W4 + (k + 7) * (FC / 4));
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(sX + (ty * 8 + r) * x_stride + k);
            acc[r][0] = fmaf(x.x, w0.x, fmaf(x.y, w1.x, fmaf(x.z, w2.x, fmaf(x.w, w3.x, acc[r][0]))));
            acc[r][1] = fmaf(x.x, w0.y, fmaf(x.y, w1.y, fmaf(x.z, w2.y, fmaf(x.w, w3.y, acc[r][1]))));
            acc[r][2] = fmaf(x.x, w0.z, fmaf(x.y, w1.z, fmaf(x.z, w2.z, fmaf(x.w, w3.z, acc[r][2]))));
            acc[r][3] = fmaf(x.x, w0.w, fmaf(x.y, w1.w, fmaf(x.z, w2.w, fmaf(x.w, w3.w, acc[r][3]))));
        }
    }
}

__global__ void __launch_bounds__(SH_THREADS) shade_fwd_kernel(const __grid_constant__ ShadeArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ta = a.ta, k1 = a.k1, fs = a.sF_stride;
    float* sB = smem;                                            // basis transposed + padded: [ta][32]
    float* sF = sB + ta * 32;                                    // [64][ta+4]; later aliased by sH [64][FC]
    float* sX = sF + ((max(SH_RAYS * fs, SH_RAYS * FC) + 3) & ~3);   // [64][k1]
    float* sH = sF;
    const long long r0 = (long long)blockIdx.x * SH_RAYS;

    // tiles without a single appearance sample (background) skip the MLP, as the reference does for such rays
    // (tensorBase.py:876-896): rgb = bg * (1 - acc)
    {
        const long long r = r0 + tid;
        const bool live = tid < SH_RAYS && r < a.n_rays;
        if (!__syncthreads_or(live && __ldg(a.app_count + r) > 0)) {
            if (live) {
                const float ac = __ldg(a.acc + r);
#pragma unroll
                for (int c = 0; c < 3; ++c) tvm_put_rgb(a.peers, a.rgb, r, c, fminf(fmaxf(__ldg(a.bg + c) * (1.f - ac), 0.f), 1.f));
                const float last = __ldg(a.rays + r * a.ray_stride + a.ray_stride - 1);
                tvm_put_depth(a.peers, a.depth_out, r, __ldg(a.depth + r) + (1.f - ac) * last);
                if (a.acc_out) a.acc_out[r] = ac;
            }
            return;
        }
    }

    for (int i = tid; i < a.app_dim * ta; i += SH_THREADS) {
        const int j = i / ta, c = i - j * ta;
        sB[c * 32 + j] = __ldg(a.basis + i);
    }
    for (int i = tid; i < (32 - a.app_dim) * ta; i += SH_THREADS) {
        const int c = i / (32 - a.app_dim), j = a.app_dim + i - c * (32 - a.app_dim);
        sB[c * 32 + j] = 0.f;
    }
    for (int i = tid; i < SH_RAYS * (ta >> 2); i += SH_THREADS) {
        const int ray = i / (ta >> 2), c4 = i - ray * (ta >> 2);
        const long long r = r0 + ray;
        // rows of rays without appearance samples are never written by the split march (and are not shaded)
        const float4 v = (r < a.n_rays && __ldg(a.app_count + r) > 0)
                             ? __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * ta) + c4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sF + ray * fs + c4 * 4) = v;
    }
    __syncthreads();

    // ---- basis_mat: feat[j] = sum_c B[j][c] * F[c]  -> X[:, 0:app_dim]; viewdirs -> X[:, app_dim:app_dim+3]
    {
        // register tile: 2 rays x 4 output columns per thread (tx = column group, ty = ray pair)
        const int tx = tid & 7, ty = tid >> 3;
        float o[2][4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[rr][e] = 0.f;
        for (int c = 0; c < ta; c += 4) {
            const float4 b0 = *reinterpret_cast<const float4*>(sB + (c + 0) * 32 + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(sB + (c + 1) * 32 + tx * 4);
            const float4 b2 = *reinterpret_cast<const float4*>(sB + (c + 2) * 32 + tx * 4);
            const float4 b3 = *reinterpret_cast<const float4*>(sB + (c + 3) * 32 + tx * 4);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const float4 x = *reinterpret_cast<const float4*>(sF + (ty * 2 + rr) * fs + c);
                o[rr][0] = fmaf(x.x, b0.x, fmaf(x.y, b1.x, fmaf(x.z, b2.x, fmaf(x.w, b3.x, o[rr][0]))));
                o[rr][1] = fmaf(x.x, b0.y, fmaf(x.y, b1.y, fmaf(x.z, b2.y, fmaf(x.w, b3.y, o[rr][1]))));
                o[rr][2] = fmaf(x.x, b0.z, fmaf(x.y, b1.z, fmaf(x.z, b2.z, fmaf(x.w, b3.z, o[rr][2]))));
                o[rr][3] = fmaf(x.x, b0.w, fmaf(x.y, b1.w, fmaf(x.z, b2.w, fmaf(x.w, b3.w, o[rr][3]))));
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = tx * 4 + e;
                if (j < a.app_dim) sX[(ty * 2 + rr) * k1 + j] = o[rr][e];
            }
        if (tid < SH_RAYS * 3) {
            const int vr = tid / 3, c = tid - vr * 3;
            const long long r = r0 + vr;
            sX[vr * k1 + a.app_dim + c] = (r < a.n_rays) ? __ldg(a.rays + r * a.ray_stride + 3 + c) : 0.f;
        }
        for (int i = tid; i < SH_RAYS * (k1 - a.in_c); i += SH_THREADS) {
            const int pr = i / (k1 - a.in_c), c = i - pr * (k1 - a.in_c);
            sX[pr * k1 + a.in_c + c] = 0.f;
        }
    }
    __syncthreads();

    // ---- positional encodings (tensorBase.py:14-20): index = channel*freqs + j, sin block then cos block
    {
        const int nbase = a.app_dim + 3;
        const int sin_f = nbase, cos_f = sin_f + a.app_dim * a.fea_pe;
        const int sin_v = cos_f + a.app_dim * a.fea_pe, cos_v = sin_v + 3 * a.view_pe;
        const int chs = nbase > 32 ? 6 : 5;                           // lanes run over the channels of one ray
        for (int it = tid; it < (SH_RAYS << chs); it += SH_THREADS) { // (bank-conflict-free; nbase <= 35)
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
    }
    __syncthreads();

    // ---- layer 1 + ReLU, layer 2 + ReLU (register tile: 8 rays x 4 outputs per thread)
    const int tx = lane, ty = warp;
    float acc[8][4];
    {
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.b1) + tx);
#pragma unroll
        for (int r = 0; r < 8; ++r) { acc[r][0] = b.x; acc[r][1] = b.y; acc[r][2] = b.z; acc[r][3] = b.w; }
    }
    tile_gemm<0>(acc, sX, k1, k1, a.w1t, tx, ty);
    // sF is dead (basis done, barrier passed) -> sH aliases it
#pragma unroll
    for (int r = 0; r < 8; ++r)
        *reinterpret_cast<float4*>(sH + (ty * 8 + r) * FC + tx * 4) =
            make_float4(fmaxf(acc[r][0], 0.f), fmaxf(acc[r][1], 0.f), fmaxf(acc[r][2], 0.f), fmaxf(acc[r][3], 0.f));
    __syncthreads();
    {
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.b2) + tx);
#pragma unroll
        for (int r = 0; r < 8; ++r) { acc[r][0] = b.x; acc[r][1] = b.y; acc[r][2] = b.z; acc[r][3] = b.w; }
    }
    tile_gemm<1>(acc, sH, FC, FC, a.w2t, tx, ty);
    __syncthreads();                                            // all reads of sH done before overwrite
#pragma unroll
    for (int r = 0; r < 8; ++r)
        *reinterpret_cast<float4*>(sH + (ty * 8 + r) * FC + tx * 4) =
            make_float4(fmaxf(acc[r][0], 0.f), fmaxf(acc[r][1], 0.f), fmaxf(acc[r][2], 0.f), fmaxf(acc[r][3], 0.f));
    __syncthreads();

    // ---- layer 3 + sigmoid + background blend + depth tail: warp w owns rays 8w..8w+7
    {
        const float4 wr = __ldg(reinterpret_cast<const float4*>(a.w3) + lane);
        const float4 wg = __ldg(reinterpret_cast<const float4*>(a.w3 + FC) + lane);
        const float4 wb = __ldg(reinterpret_cast<const float4*>(a.w3 + 2 * FC) + lane);
        for (int rr = 0; rr < 8; ++rr) {
            const int ray = warp * 8 + rr;
            const long long r = r0 + ray;
            if (r >= a.n_rays) break;                            // warp-uniform
            const float4 h = *reinterpret_cast<const float4*>(sH + ray * FC + lane * 4);
            float vr = h.x * wr.x + h.y * wr.y + h.z * wr.z + h.w * wr.w;
            float vg = h.x * wg.x + h.y * wg.y + h.z * wg.z + h.w * wg.w;
            float vb = h.x * wb.x + h.y * wb.y + h.z * wb.z + h.w * wb.w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vr += __shfl_xor_sync(FULL, vr, o);
                vg += __shfl_xor_sync(FULL, vg, o);
                vb += __shfl_xor_sync(FULL, vb, o);
            }
            if (lane < 3) {
                const float v = (lane == 0 ? vr : (lane == 1 ? vg : vb)) + __ldg(a.b3 + lane);
                const bool lit = __ldg(a.app_count + r) > 0;     // rays_to_consider (tensorBase.py:886)
                const float c = lit ? 1.f / (1.f + expf(-v)) : 0.f;
                const float ac = __ldg(a.acc + r);
                float out = c * ac + __ldg(a.bg + lane) * (1.f - ac);
                out = fminf(fmaxf(out, 0.f), 1.f);
                tvm_put_rgb(a.peers, a.rgb, r, lane, out);
                if (lane == 0) {
                    const float last = __ldg(a.rays + r * a.ray_stride + a.ray_stride - 1);
                    tvm_put_depth(a.peers, a.depth_out, r, __ldg(a.depth + r) + (1.f - ac) * last);
                    if (a.acc_out) a.acc_out[r] = ac;
                }
            }
        }
    }
}

}  // namespace

static int shade_fwd_core(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                          const float* bg, uint32_t flags, float* rgb, float* depth, float* acc, const void* ws,
                          size_t ws_bytes, void* stream, const tvm_scatter_out* sc) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (n_rays == 0) return 0;
    if (!rays || (!rgb && !sc) || !ws || !desc->basis || !desc->mlp || !bg) return TVM_E_NULL;
    if (desc->feature_c != FC) return TVM_E_SHAPE;
    if (desc->app_dim > 32 || desc->app_dim <= 0) return TVM_E_SHAPE;
    if (flags & TVM_F_MLP_TC3)                        // tcgen05 bf16x3 split variant (shade_tc3.cu), fp32-equivalent
        return tvm_shade_tc3_launch(desc, rays, n_rays, ray_stride, bg, rgb, depth, acc, ws, ws_bytes,
                                    (cudaStream_t)stream, sc);
    if (flags & TVM_F_MLP_BF16)                       // tcgen05 bf16 variant (shade_tc.cu), tolerance 1e-2
        return tvm_shade_tc_launch(desc, rays, n_rays, ray_stride, bg, rgb, depth, acc, ws, ws_bytes,
                                   (cudaStream_t)stream, sc);
    const TvmWorkspace w = tvm_ws_layout(desc, n_rays);
    if (ws_bytes < w.total) return TVM_E_WORKSPACE;
    const TvmMlpLayout m = tvm_mlp_layout(desc);
    const char* base = (const char*)ws;
    ShadeArgs a{};
    a.rays = rays; a.n_rays = n_rays; a.ray_stride = ray_stride;
    a.bg = bg;
    a.rgb = rgb; a.depth_out = depth; a.acc_out = acc;
    rc = tvm_fill_peers(a.peers, sc);
    if (rc) return rc;
    a.ray_feat = (const float*)(base + w.ray_feat);
    a.acc = (const float*)(base + w.acc);
    a.depth = (const float*)(base + w.depth);
    a.app_count = (const int*)(base + w.app_count);
    a.basis = desc->basis;
    a.w1t = desc->mlp + m.w1t; a.b1 = desc->mlp + m.b1; a.w2t = desc->mlp + m.w2t; a.b2 = desc->mlp + m.b2;
    a.w3 = desc->mlp + m.w3; a.b3 = desc->mlp + m.b3;
    a.ta = tvm_total_app(desc); a.app_dim = desc->app_dim; a.fea_pe = desc->fea_pe; a.view_pe = desc->view_pe;
    a.in_c = m.in_c; a.k1 = m.k1; a.sF_stride = a.ta + 4;     // +4: conflict-free float4 rows
    const size_t nB = (size_t)a.ta * 32;
    const size_t nF = (size_t)((max(SH_RAYS * a.sF_stride, SH_RAYS * FC) + 3) & ~3);
    const size_t nX = (size_t)SH_RAYS * a.k1;
    const size_t smem = (nB + nF + nX) * sizeof(float);
    if (smem > 227 * 1024) return TVM_E_SHAPE;
    {
        static TvmDevMemo smem_set;
        int rc_attr = tvm_ensure_dyn_smem(shade_fwd_kernel, smem, smem_set);
        if (rc_attr) return rc_attr;
    }
    const long long ctas = (n_rays + SH_RAYS - 1) / SH_RAYS;
    tvm_count_launch(); shade_fwd_kernel<<<(unsigned)ctas, SH_THREADS, smem, (cudaStream_t)stream>>>(a);
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int tvm_shade_fwd(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                             const float* bg, uint32_t flags, float* rgb, float* depth, float* acc, const void* ws,
                             size_t ws_bytes, void* stream) {
    return shade_fwd_core(desc, rays, n_rays, ray_stride, bg, flags, rgb, depth, acc, ws, ws_bytes, stream, nullptr);
}

extern "C" int tvm_shade_fwd_scatter(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                                     const float* bg, uint32_t flags, const tvm_scatter_out* out, float* acc,
                                     const void* ws, size_t ws_bytes, void* stream) {
    if (!out) return TVM_E_NULL;
    return shade_fwd_core(desc, rays, n_rays, ray_stride, bg, flags, nullptr, nullptr, acc, ws, ws_bytes, stream, out);
}

extern "C" int tvm_render_fwd(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                              int n_samples, const float* jitter, const float* bg, uint32_t flags, float* rgb,
                              float* depth, float* acc, float* alpha, float* z_vals, float* dists,
                              uint32_t* valid_bits, int32_t* valid_count, int32_t* app_count, void* ws,
                              size_t ws_bytes, void* stream) {
    int rc = tvm_march_fwd_launch(desc, rays, n_rays, ray_stride, n_samples, jitter, flags, alpha, z_vals, dists,
                                  valid_bits, valid_count, app_count, ws, ws_bytes, (cudaStream_t)stream);
    if (rc) return rc;
    if (flags & TVM_F_NO_SHADE) return 0;
    return tvm_shade_fwd(desc, rays, n_rays, ray_stride, bg, flags, rgb, depth, acc, ws, ws_bytes, stream);
}
