// shade_ref.cu — fused per-ray tail for the `Ref` shading head, forward / eval.
//
// What configs/lego.txt:25 and truck.txt:26 select: models/ref.py:48-155 evaluated once per ray on the accumulated
// appearance feature (models/tensorBase.py:886-896), followed by the composite tail (:898-908):
//   feat      = basis_mat . ray_feat                                     (tensoRF.py:158, hoisted past the sum)
//   normals   = -normalize(Wn feat + bn)                                 (ref.py normal_mlp: Linear, UnitNorm, Scale(-1))
//   tint      = sigmoid(Wt feat + bt);  rough = softplus(Wr feat + br - 1);  diffuse = sigmoid(Wd feat + bd - ln 3)
//   refdirs   = reflect(-v, n) = 2 (n.(-v)) n + v                        (ref_utils.py:6-19)
//   enc       = IDE(refdirs, rough): (x+iy)^m * sum_k z^k mat[k][j] * exp(-l(l+1)/2 * rough), re/im interleaved (:22-112)
//   specular  = tint * sigmoid(premul * (Ws [bottleneck(feat), enc, n.v] + bs) + bias)
//   rgb       = clip(linear_to_srgb(specular + diffuse), 0, 1) * (1 + 2 pad) - pad        (image.py:6-13)
//   rgb_map   = clamp([app samples > 0] rgb * acc + bg (1 - acc), 0, 1);  depth = sum w z + (1 - acc) rays[:, -1]
// One thread owns one ray; every weight matrix sits in shared memory and is read as warp-wide broadcasts, the
// 128-wide bottleneck is folded into the 3 specular dot products on the fly (never materialised).  ~8 kFMA per ray.
#include "tvm_common.cuh"

namespace {

constexpr int REF_THREADS = 128;
constexpr int REF_MAX_IN = 32;
constexpr int REF_MAX_PAIRS = 32;
constexpr int REF_MAX_L = 16;

struct RefLayout {
    int in4;          // in_c rounded up to 4 (row stride of the in_c-input matrices)
    int n_dir;        // 2 * n_pairs + 1
    int small_w;      // 10 rows x in4: normal(3) tint(3) rough(1) diffuse(3)
    int bott_w;       // feature_c rows x in4
    int small_b;      // 10 (+2 pad)
    int bott_b;       // feature_c
    int spec_w;       // 3 x (feature_c + n_dir)
    int spec_b;       // 3 (+1 pad)
    int ide_mat;      // (l_max + 1) x n_pairs
    int total;
};
__host__ __device__ inline RefLayout ref_layout(const tvm_ref_head& h) {
    RefLayout L;
    L.in4 = (h.in_c + 3) & ~3;
    L.n_dir = 2 * h.n_pairs + 1;
    int off = 0;
    L.small_w = off; off += 10 * L.in4;
    L.bott_w = off;  off += h.feature_c * L.in4;
    L.small_b = off; off += 12;
    L.bott_b = off;  off += h.feature_c;
    L.spec_w = off;  off += 3 * (h.feature_c + L.n_dir); off = (off + 3) & ~3;
    L.spec_b = off;  off += 4;
    L.ide_mat = off; off += (h.l_max + 1) * h.n_pairs; off = (off + 3) & ~3;
    L.total = off;
    return L;
}

struct RefArgs {
    tvm_field_desc f;
    tvm_ref_head h;
    const float* rays;
    long long n;
    int ray_stride;
    const float* bg;
    float* rgb;
    float* depth;
    float* acc_out;
    TvmPeers peers;
    const float* ray_feat;
    const float* acc;
    const float* depth_part;
    const int* app_count;
    int ta;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float to_srgb(float v) {
    const float eps = 1.1920928955078125e-07f;                     // torch.finfo(float32).eps
    const float low = (323.0f / 25.0f) * v;
    const float high = (211.0f * powf(fmaxf(v, eps), 5.0f / 12.0f) - 11.0f) / 200.0f;
    return v <= 0.0031308f ? low : high;
}

template <int IN_C>
__global__ void __launch_bounds__(REF_THREADS) shade_ref_kernel(const __grid_constant__ RefArgs a) {
    extern __shared__ __align__(16) float smem[];
    const tvm_ref_head& h = a.h;
    const RefLayout L = ref_layout(h);
    constexpr int IN4 = (IN_C + 3) & ~3;
    float* s_par = smem;                         // packed head parameters
    float* s_basis = smem + L.total;             // [IN_C][ta]
    for (int i = threadIdx.x; i < L.total; i += REF_THREADS) s_par[i] = __ldg(h.params + i);
    for (int i = threadIdx.x; i < IN_C * a.ta; i += REF_THREADS) s_basis[i] = __ldg(a.f.basis + i);
    __syncthreads();

    // persistent CTAs: the 34 KB of weights are staged once, then the CTA walks 128-ray tiles
    for (long long r = (long long)blockIdx.x * REF_THREADS + threadIdx.x; r < a.n; r += (long long)gridDim.x * REF_THREADS) {

    const float* rp = a.rays + r * a.ray_stride;
    if (a.app_count[r] <= 0) {       // no appearance sample: the head is not evaluated (tensorBase.py:876-896), rgb = bg (1 - acc)
        const float acc0 = a.acc[r];
#pragma unroll
        for (int o = 0; o < 3; ++o) tvm_put_rgb(a.peers, a.rgb, r, o, fminf(fmaxf(__ldg(a.bg + o) * (1.0f - acc0), 0.f), 1.f));
        tvm_put_depth(a.peers, a.depth, r, a.depth_part[r] + (1.0f - acc0) * __ldg(rp + a.ray_stride - 1));
        if (a.acc_out) a.acc_out[r] = acc0;
        continue;
    }
    // ---- feat = basis_mat . ray_feat
    float F[IN4];
#pragma unroll
    for (int i = 0; i < IN4; ++i) F[i] = 0.f;
    const float4* rf = reinterpret_cast<const float4*>(a.ray_feat + r * a.ta);
    for (int c4 = 0; c4 < (a.ta >> 2); ++c4) {
        const float4 x = __ldg(rf + c4);
#pragma unroll
        for (int i = 0; i < IN_C; ++i) {
            const float4 b = *reinterpret_cast<const float4*>(s_basis + i * a.ta + 4 * c4);
            F[i] = fmaf(b.x, x.x, fmaf(b.y, x.y, fmaf(b.z, x.z, fmaf(b.w, x.w, F[i]))));
        }
    }
    const float v[3] = {__ldg(rp + 3), __ldg(rp + 4), __ldg(rp + 5)};

    // ---- the ten scalar heads
    float small[10];
#pragma unroll
    for (int o = 0; o < 10; ++o) {
        float s = s_par[L.small_b + o];
#pragma unroll
        for (int i = 0; i < IN_C; ++i) s = fmaf(s_par[L.small_w + o * IN4 + i], F[i], s);
        small[o] = s;
    }
    float nrm[3];
    {
        const float len = fmaxf(sqrtf(small[0] * small[0] + small[1] * small[1] + small[2] * small[2]), 1e-12f);
#pragma unroll
        for (int c = 0; c < 3; ++c) nrm[c] = -(small[c] / len);
    }
    const float tint[3] = {sigmoidf_(small[3]), sigmoidf_(small[4]), sigmoidf_(small[5])};
    const float rough = softplusf_(small[6] + h.rough_shift);
    const float diffuse[3] = {sigmoidf_(small[7] + h.diffuse_shift), sigmoidf_(small[8] + h.diffuse_shift),
                              sigmoidf_(small[9] + h.diffuse_shift)};
    // reflect(-v, n): 2 (n . (-v)) n - (-v)
    const float ndv = nrm[0] * v[0] + nrm[1] * v[1] + nrm[2] * v[2];
    float rd[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) rd[c] = (2.0f * -ndv) * nrm[c] + v[c];

    // ---- specular pre-activation: bottleneck columns, then the directional encoding, then n.v
    const int FC = h.feature_c, SW = FC + L.n_dir;
    float sp[3] = {s_par[L.spec_b], s_par[L.spec_b + 1], s_par[L.spec_b + 2]};
    for (int j = 0; j < FC; ++j) {
        float b = s_par[L.bott_b + j];
        const float* wrow = s_par + L.bott_w + j * IN4;
#pragma unroll
        for (int i = 0; i < IN_C; ++i) b = fmaf(wrow[i], F[i], b);
#pragma unroll
        for (int o = 0; o < 3; ++o) sp[o] = fmaf(s_par[L.spec_w + o * SW + j], b, sp[o]);
    }
    {
        float zp[REF_MAX_L + 1], cre[REF_MAX_L + 1], cim[REF_MAX_L + 1];       // z^k, (x+iy)^m
        zp[0] = 1.f; cre[0] = 1.f; cim[0] = 0.f;
#pragma unroll
        for (int k = 1; k <= REF_MAX_L; ++k) {
            if (k <= h.l_max) {
                zp[k] = zp[k - 1] * rd[2];
                cre[k] = cre[k - 1] * rd[0] - cim[k - 1] * rd[1];
                cim[k] = cre[k - 1] * rd[1] + cim[k - 1] * rd[0];
            }
        }
        for (int p = 0; p < h.n_pairs; ++p) {
            const int m = h.m[p], l = h.l[p];
            float poly = 0.f;
            for (int k = 0; k <= h.l_max; ++k) poly = fmaf(zp[k], s_par[L.ide_mat + k * h.n_pairs + p], poly);
            const float att = expf(-(0.5f * (float)l * (float)(l + 1)) * rough);
            // dynamic index into the small power tables (local memory; 19 pairs per ray)
            const float re = cre[m] * poly * att, im = cim[m] * poly * att;
#pragma unroll
            for (int o = 0; o < 3; ++o)
                sp[o] = fmaf(s_par[L.spec_w + o * SW + FC + 2 * p], re, fmaf(s_par[L.spec_w + o * SW + FC + 2 * p + 1], im, sp[o]));
        }
    }
    float rgb[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        sp[o] = fmaf(s_par[L.spec_w + o * SW + SW - 1], ndv, sp[o]);
        const float specular = tint[o] * sigmoidf_(sp[o] * h.rgb_premultiplier + h.rgb_bias);
        const float lin = specular + diffuse[o];
        rgb[o] = fminf(fmaxf(to_srgb(lin), 0.f), 1.f) * (1.0f + 2.0f * h.rgb_padding) - h.rgb_padding;
    }
    // ---- composite tail (tensorBase.py:886-908)
    const float acc = a.acc[r];
    const float lit = a.app_count[r] > 0 ? 1.f : 0.f;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        const float c = (rgb[o] * lit) * acc + __ldg(a.bg + o) * (1.0f - acc);
        tvm_put_rgb(a.peers, a.rgb, r, o, fminf(fmaxf(c, 0.f), 1.f));
    }
    tvm_put_depth(a.peers, a.depth, r, a.depth_part[r] + (1.0f - acc) * __ldg(rp + a.ray_stride - 1));
    if (a.acc_out) a.acc_out[r] = acc;
    }
}

}  // namespace

extern "C" size_t tvm_ref_head_floats(const tvm_ref_head* head) { return head ? (size_t)ref_layout(*head).total : 0; }

extern "C" int tvm_ref_head_layout(const tvm_ref_head* head, int32_t offs[8]) {
    if (!head || !offs) return TVM_E_NULL;
    const RefLayout L = ref_layout(*head);
    offs[0] = L.small_w; offs[1] = L.bott_w; offs[2] = L.small_b; offs[3] = L.bott_b;
    offs[4] = L.spec_w; offs[5] = L.spec_b; offs[6] = L.ide_mat; offs[7] = L.in4;
    return 0;
}

static int shade_ref_fwd_core(const tvm_field_desc* desc, const tvm_ref_head* head, const float* rays,
                              int64_t n_rays, int ray_stride, const float* bg, float* rgb, float* depth,
                              float* acc, const void* ws, size_t ws_bytes, void* stream, const tvm_scatter_out* sc) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (!head) return TVM_E_NULL;
    if (n_rays == 0) return 0;
    if (!rays || !bg || ((!rgb || !depth) && !sc) || !ws || !desc->basis || !head->params) return TVM_E_NULL;
    if (ray_stride < 6 || head->in_c != desc->app_dim || head->in_c > REF_MAX_IN || head->n_pairs <= 0 ||
        head->n_pairs > REF_MAX_PAIRS || head->l_max <= 0 || head->l_max > REF_MAX_L || head->feature_c <= 0)
        return TVM_E_SHAPE;
    for (int p = 0; p < head->n_pairs; ++p)
        if (head->m[p] < 0 || head->m[p] > head->l_max || head->l[p] < 0) return TVM_E_SHAPE;
    const TvmWorkspace w = tvm_ws_layout(desc, n_rays);
    if (ws_bytes < w.total) return TVM_E_WORKSPACE;
    RefArgs a{};
    a.f = *desc; a.h = *head; a.rays = rays; a.n = n_rays; a.ray_stride = ray_stride; a.bg = bg;
    a.rgb = rgb; a.depth = depth; a.acc_out = acc;
    rc = tvm_fill_peers(a.peers, sc);
    if (rc) return rc;
    const char* base = (const char*)ws;
    a.ray_feat = (const float*)(base + w.ray_feat);
    a.acc = (const float*)(base + w.acc);
    a.depth_part = (const float*)(base + w.depth);
    a.app_count = (const int*)(base + w.app_count);
    a.ta = tvm_total_app(desc);
    const size_t smem = ((size_t)ref_layout(*head).total + (size_t)head->in_c * a.ta) * sizeof(float);
    if (smem > 200 * 1024) return TVM_E_SHAPE;
    long long tiles = (n_rays + REF_THREADS - 1) / REF_THREADS;
    const unsigned ctas = (unsigned)(tiles < TVM_SM_COUNT * 6 ? tiles : TVM_SM_COUNT * 6);
    cudaStream_t st = (cudaStream_t)stream;
    if (head->in_c == 27) {
        static TvmDevMemo smem_set;
        int rc_attr = tvm_ensure_dyn_smem(shade_ref_kernel<27>, smem, smem_set);
        if (rc_attr) return rc_attr;
        tvm_count_launch(); shade_ref_kernel<27><<<ctas, REF_THREADS, smem, st>>>(a);
    } else {
        return TVM_E_SHAPE;          // app_dim = 27 is what every reference config uses (configs/*.txt)
    }
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int tvm_shade_ref_fwd(const tvm_field_desc* desc, const tvm_ref_head* head, const float* rays,
                                 int64_t n_rays, int ray_stride, const float* bg, float* rgb, float* depth,
                                 float* acc, const void* ws, size_t ws_bytes, void* stream) {
    return shade_ref_fwd_core(desc, head, rays, n_rays, ray_stride, bg, rgb, depth, acc, ws, ws_bytes, stream, nullptr);
}

extern "C" int tvm_shade_ref_fwd_scatter(const tvm_field_desc* desc, const tvm_ref_head* head, const float* rays,
                                         int64_t n_rays, int ray_stride, const float* bg, const tvm_scatter_out* out,
                                         float* acc, const void* ws, size_t ws_bytes, void* stream) {
    if (!out) return TVM_E_NULL;
    return shade_ref_fwd_core(desc, head, rays, n_rays, ray_stride, bg, nullptr, nullptr, acc, ws, ws_bytes, stream, out);
}

// =====================================================================================================================
// Backward of the `Ref` tail (training with shadingMode="Ref", configs/lego.txt:25): replaces torch autograd through
// models/ref.py:103-155 + the composite tail (tensorBase.py:886-904).  d(rgb_map) [n][3] (+ optional upstream on the
// acc_map output) -> d(ray_feat) [n][ta], d(acc) [n], optional d(view) [n][3], and ACCUMULATED parameter gradients:
// g_basis [in_c][ta] and g_params in the packed layout of tvm_ref_head (the ide_mat section stays untouched).
// One CTA owns 64 rays.  Threads 0..63 walk one ray each through the forward (recomputed) and the per-ray backward,
// leaving F, g_F, g_small, g_sp, the encoding tail, g_bottleneck and the ray_feat row in shared memory; then all 128
// threads form the weight gradients as [out x 64 rays] . [64 rays x in] products and flush them with atomics.  The
// specular layer's bottleneck columns need no per-ray bottleneck storage: g_Ws[o][j] = Wb[j] . (sum_r g_sp[r][o] F_r)
// + bb[j] sum_r g_sp[r][o].
// =====================================================================================================================
namespace {

#ifndef TVM_REF_BWD_RAYS
#define TVM_REF_BWD_RAYS 64
#endif
constexpr int RB_RAYS = TVM_REF_BWD_RAYS;     // rays per CTA (one thread per ray in the per-ray phase)
constexpr int RB_THREADS = 128;
constexpr int RB_XT = 40;            // encoding tail (2 n_pairs + 1 <= 39) padded

struct RefBwdArgs {
    tvm_field_desc f;
    tvm_ref_head h;
    const float* rays;
    long long n;
    int ray_stride;
    const float* bg;
    const float* ray_feat;
    const float* acc;
    const int* app_count;
    const float* d_rgb;
    const float* d_acc_in;
    float* d_ray_feat;
    float* d_acc;
    float* d_view;
    float* g_basis;
    float* g_params;
    int ta;
};

template <int IN_C>
__global__ void __launch_bounds__(RB_THREADS) shade_ref_bwd_kernel(const __grid_constant__ RefBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const tvm_ref_head& h = a.h;
    const RefLayout L = ref_layout(h);
    constexpr int IN4 = (IN_C + 3) & ~3;
    const int ta = a.ta, FC = h.feature_c, NP = h.n_pairs, SW = FC + L.n_dir;
    float* s_par = smem;
    float* s_basis = s_par + L.total;                 // [IN_C][ta]
    float* sF = s_basis + IN_C * ta;                  // [64][IN4]
    // per-ray rows are written by one thread per ray: odd row strides keep those scalar accesses on distinct banks,
    // and ta + 4 does the same for the float4 reads of the ray_feat rows
    constexpr int SFS = IN4 + 1, SMS = 13, SPS = 5, XTS = RB_XT + 1;
    const int GBS = FC + 1, RFS = ta + 4;
    float* sGF = sF + RB_RAYS * SFS;                  // [64][IN4]
    float* sGsm = sGF + RB_RAYS * SFS;                // [64][10]   g_small
    float* sGsp = sGsm + RB_RAYS * SMS;               // [64][3]    g_sp
    float* sXt = sGsp + RB_RAYS * SPS;                // [64][RB_XT] encoding tail: enc (2 NP), n.v
    float* sGb = sXt + RB_RAYS * XTS;                 // [64][FC]   g_bottleneck
    float* sRf = sGb + RB_RAYS * GBS;                 // [64][ta]   ray_feat rows (16-byte aligned: 64 x anything)
    const int tid = threadIdx.x;
    for (int i = tid; i < L.total; i += RB_THREADS) s_par[i] = __ldg(h.params + i);
    for (int i = tid; i < IN_C * ta; i += RB_THREADS) s_basis[i] = __ldg(a.f.basis + i);
    const long long r0 = (long long)blockIdx.x * RB_RAYS;
    for (int i = tid; i < RB_RAYS * (ta >> 2); i += RB_THREADS) {
        const int ray = i / (ta >> 2), c4 = i - ray * (ta >> 2);
        const long long r = r0 + ray;
        reinterpret_cast<float4*>(sRf + ray * RFS)[c4] =
            r < a.n ? __ldg(reinterpret_cast<const float4*>(a.ray_feat + r * ta) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    if (tid < RB_RAYS) {
        const int ray = tid;
        const long long r = r0 + ray;
        const bool live = r < a.n;
        float F[IN4], gF[IN4];
#pragma unroll
        for (int i = 0; i < IN4; ++i) { F[i] = 0.f; gF[i] = 0.f; }
        float gsm[10], gsp[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 10; ++o) gsm[o] = 0.f;
        for (int j = 0; j < RB_XT; ++j) sXt[ray * XTS + j] = 0.f;
        for (int j = 0; j < FC; ++j) sGb[ray * GBS + j] = 0.f;
        float d_acc = 0.f, gv[3] = {0.f, 0.f, 0.f};
        if (live) {
            const float* rp = a.rays + r * a.ray_stride;
            const float v[3] = {__ldg(rp + 3), __ldg(rp + 4), __ldg(rp + 5)};
            const float acc = __ldg(a.acc + r);
            const bool lit = __ldg(a.app_count + r) > 0;
            const float go[3] = {__ldg(a.d_rgb + r * 3), __ldg(a.d_rgb + r * 3 + 1), __ldg(a.d_rgb + r * 3 + 2)};
            d_acc = a.d_acc_in ? __ldg(a.d_acc_in + r) : 0.f;
            // ---------------- forward (same arithmetic as shade_ref_kernel) ----------------
            for (int c4 = 0; c4 < (ta >> 2); ++c4) {
                const float4 x = *reinterpret_cast<const float4*>(sRf + ray * RFS + 4 * c4);
#pragma unroll
                for (int i = 0; i < IN_C; ++i) {
                    const float4 b = *reinterpret_cast<const float4*>(s_basis + i * ta + 4 * c4);
                    F[i] = fmaf(b.x, x.x, fmaf(b.y, x.y, fmaf(b.z, x.z, fmaf(b.w, x.w, F[i]))));
                }
            }
            float small[10];
#pragma unroll
            for (int o = 0; o < 10; ++o) {
                float s = s_par[L.small_b + o];
#pragma unroll
                for (int i = 0; i < IN_C; ++i) s = fmaf(s_par[L.small_w + o * IN4 + i], F[i], s);
                small[o] = s;
            }
            const float slen = sqrtf(small[0] * small[0] + small[1] * small[1] + small[2] * small[2]);
            const float len = fmaxf(slen, 1e-12f);
            const float nh[3] = {small[0] / len, small[1] / len, small[2] / len};
            const float nrm[3] = {-nh[0], -nh[1], -nh[2]};
            const float tint[3] = {sigmoidf_(small[3]), sigmoidf_(small[4]), sigmoidf_(small[5])};
            const float rough = softplusf_(small[6] + h.rough_shift);
            const float diffuse[3] = {sigmoidf_(small[7] + h.diffuse_shift), sigmoidf_(small[8] + h.diffuse_shift),
                                      sigmoidf_(small[9] + h.diffuse_shift)};
            const float ndv = nrm[0] * v[0] + nrm[1] * v[1] + nrm[2] * v[2];
            float rd[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) rd[c] = (2.0f * -ndv) * nrm[c] + v[c];
            float sp[3] = {s_par[L.spec_b], s_par[L.spec_b + 1], s_par[L.spec_b + 2]};
            for (int j = 0; j < FC; ++j) {
                float b = s_par[L.bott_b + j];
                const float* wrow = s_par + L.bott_w + j * IN4;
#pragma unroll
                for (int i = 0; i < IN_C; ++i) b = fmaf(wrow[i], F[i], b);
#pragma unroll
                for (int o = 0; o < 3; ++o) sp[o] = fmaf(s_par[L.spec_w + o * SW + j], b, sp[o]);
            }
            float zp[REF_MAX_L + 1], cre[REF_MAX_L + 1], cim[REF_MAX_L + 1];
            zp[0] = 1.f; cre[0] = 1.f; cim[0] = 0.f;
            for (int k = 1; k <= h.l_max; ++k) {
                zp[k] = zp[k - 1] * rd[2];
                cre[k] = cre[k - 1] * rd[0] - cim[k - 1] * rd[1];
                cim[k] = cre[k - 1] * rd[1] + cim[k - 1] * rd[0];
            }
            for (int p = 0; p < NP; ++p) {
                const int m = h.m[p], l = h.l[p];
                float poly = 0.f;
                for (int k = 0; k <= h.l_max; ++k) poly = fmaf(zp[k], s_par[L.ide_mat + k * NP + p], poly);
                const float att = expf(-(0.5f * (float)l * (float)(l + 1)) * rough);
                const float re = cre[m] * poly * att, im = cim[m] * poly * att;
                sXt[ray * XTS + 2 * p] = re;
                sXt[ray * XTS + 2 * p + 1] = im;
#pragma unroll
                for (int o = 0; o < 3; ++o)
                    sp[o] = fmaf(s_par[L.spec_w + o * SW + FC + 2 * p], re, fmaf(s_par[L.spec_w + o * SW + FC + 2 * p + 1], im, sp[o]));
            }
            sXt[ray * XTS + 2 * NP] = ndv;
            float S[3], lin[3], srgb[3], rgb[3];
            const float eps = 1.1920928955078125e-07f;
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                sp[o] = fmaf(s_par[L.spec_w + o * SW + SW - 1], ndv, sp[o]);
                S[o] = sigmoidf_(sp[o] * h.rgb_premultiplier + h.rgb_bias);
                lin[o] = tint[o] * S[o] + diffuse[o];
                srgb[o] = to_srgb(lin[o]);
                rgb[o] = fminf(fmaxf(srgb[o], 0.f), 1.f) * (1.0f + 2.0f * h.rgb_padding) - h.rgb_padding;
            }
            // ---------------- backward ----------------
            const float litf = lit ? 1.f : 0.f;
            float g_lin[3];
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                const float bgc = __ldg(a.bg + o);
                const float c = (rgb[o] * litf) * acc + bgc * (1.0f - acc);
                const float g_c = (c >= 0.f && c <= 1.f) ? go[o] : 0.f;             // clamp(0,1)
                d_acc = fmaf(g_c, rgb[o] * litf - bgc, d_acc);
                const float g_rgb = g_c * litf * acc;
                const float g_srgb = (srgb[o] >= 0.f && srgb[o] <= 1.f) ? g_rgb * (1.0f + 2.0f * h.rgb_padding) : 0.f;
                float dl;
                if (lin[o] <= 0.0031308f) dl = 323.0f / 25.0f;
                else dl = (lin[o] >= eps) ? (211.0f / 200.0f) * (5.0f / 12.0f) * powf(lin[o], -7.0f / 12.0f) : 0.f;
                g_lin[o] = g_srgb * dl;
            }
            if (lit) {
                float g_ndv = 0.f, g_rough = 0.f;
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    gsm[7 + o] = g_lin[o] * diffuse[o] * (1.f - diffuse[o]);
                    gsm[3 + o] = (g_lin[o] * S[o]) * tint[o] * (1.f - tint[o]);
                    gsp[o] = (g_lin[o] * tint[o]) * S[o] * (1.f - S[o]) * h.rgb_premultiplier;
                }
                // bottleneck columns: g_b[j] = sum_o g_sp[o] Ws[o][j];  g_F += Wb^T g_b
                for (int j = 0; j < FC; ++j) {
                    const float gb = gsp[0] * s_par[L.spec_w + j] + gsp[1] * s_par[L.spec_w + SW + j] + gsp[2] * s_par[L.spec_w + 2 * SW + j];
                    sGb[ray * GBS + j] = gb;
                    const float* wrow = s_par + L.bott_w + j * IN4;
#pragma unroll
                    for (int i = 0; i < IN_C; ++i) gF[i] = fmaf(wrow[i], gb, gF[i]);
                }
                // integrated directional encoding
                float g_zp[REF_MAX_L + 1], g_cre[REF_MAX_L + 1], g_cim[REF_MAX_L + 1];
                for (int k = 0; k <= h.l_max; ++k) { g_zp[k] = 0.f; g_cre[k] = 0.f; g_cim[k] = 0.f; }
                for (int p = 0; p < NP; ++p) {
                    const int m = h.m[p], l = h.l[p];
                    float g_re = 0.f, g_im = 0.f;
#pragma unroll
                    for (int o = 0; o < 3; ++o) {
                        g_re = fmaf(gsp[o], s_par[L.spec_w + o * SW + FC + 2 * p], g_re);
                        g_im = fmaf(gsp[o], s_par[L.spec_w + o * SW + FC + 2 * p + 1], g_im);
                    }
                    float poly = 0.f;
                    for (int k = 0; k <= h.l_max; ++k) poly = fmaf(zp[k], s_par[L.ide_mat + k * NP + p], poly);
                    const float sig = 0.5f * (float)l * (float)(l + 1);
                    const float att = expf(-sig * rough);
                    const float gc = g_re * cre[m] + g_im * cim[m];
                    g_rough = fmaf(gc * poly * att, -sig, g_rough);
                    const float g_poly = gc * att;
                    for (int k = 0; k <= h.l_max; ++k) g_zp[k] = fmaf(g_poly, s_par[L.ide_mat + k * NP + p], g_zp[k]);
                    g_cre[m] = fmaf(g_re, poly * att, g_cre[m]);
                    g_cim[m] = fmaf(g_im, poly * att, g_cim[m]);
                }
                float g_rd[3] = {0.f, 0.f, 0.f};
                for (int k = 1; k <= h.l_max; ++k) {
                    const float kf = (float)k;
                    g_rd[2] = fmaf(g_zp[k] * kf, zp[k - 1], g_rd[2]);
                    g_rd[0] = fmaf(kf, g_cre[k] * cre[k - 1] + g_cim[k] * cim[k - 1], g_rd[0]);
                    g_rd[1] = fmaf(kf, -g_cre[k] * cim[k - 1] + g_cim[k] * cre[k - 1], g_rd[1]);
                }
                // n.v column of the specular layer, reflect, normalise
#pragma unroll
                for (int o = 0; o < 3; ++o) g_ndv = fmaf(gsp[o], s_par[L.spec_w + o * SW + SW - 1], g_ndv);
                float g_n[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    g_ndv = fmaf(g_rd[c], -2.0f * nrm[c], g_ndv);
                    g_n[c] = g_rd[c] * (-2.0f * ndv);
                    gv[c] = g_rd[c];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) { g_n[c] = fmaf(g_ndv, v[c], g_n[c]); gv[c] = fmaf(g_ndv, nrm[c], gv[c]); }
                // nrm = -s/len
                const float gnh[3] = {-g_n[0], -g_n[1], -g_n[2]};
                if (slen > 1e-12f) {
                    const float dotn = nh[0] * gnh[0] + nh[1] * gnh[1] + nh[2] * gnh[2];
#pragma unroll
                    for (int c = 0; c < 3; ++c) gsm[c] = (gnh[c] - nh[c] * dotn) / len;
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) gsm[c] = gnh[c] / len;
                }
                const float xr = small[6] + h.rough_shift;
                gsm[6] = g_rough * (xr > 20.f ? 1.f : sigmoidf_(xr));
                // g_F += Wsm^T g_small
#pragma unroll
                for (int o = 0; o < 10; ++o)
#pragma unroll
                    for (int i = 0; i < IN_C; ++i) gF[i] = fmaf(s_par[L.small_w + o * IN4 + i], gsm[o], gF[i]);
            }
            // d_ray_feat = g_F . B
            float4* drf = reinterpret_cast<float4*>(a.d_ray_feat + r * ta);
            for (int c4 = 0; c4 < (ta >> 2); ++c4) {
                float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < IN_C; ++i) {
                    const float4 b = *reinterpret_cast<const float4*>(s_basis + i * ta + 4 * c4);
                    o4.x = fmaf(gF[i], b.x, o4.x); o4.y = fmaf(gF[i], b.y, o4.y);
                    o4.z = fmaf(gF[i], b.z, o4.z); o4.w = fmaf(gF[i], b.w, o4.w);
                }
                drf[c4] = o4;
            }
            a.d_acc[r] = d_acc;
            if (a.d_view) { a.d_view[r * 3] = gv[0]; a.d_view[r * 3 + 1] = gv[1]; a.d_view[r * 3 + 2] = gv[2]; }
        }
#pragma unroll
        for (int i = 0; i < IN4; ++i) { sF[ray * SFS + i] = F[i]; sGF[ray * SFS + i] = gF[i]; }
#pragma unroll
        for (int o = 0; o < 10; ++o) sGsm[ray * SMS + o] = gsm[o];
#pragma unroll
        for (int o = 0; o < 3; ++o) sGsp[ray * SPS + o] = gsp[o];
    }
    __syncthreads();

    // ---------------- weight gradients: [out x 64] . [64 x in], flushed with atomics ----------------
    if (a.g_params) {
        float* G = a.g_params;
        // ten scalar heads + bottleneck: rows of width IN_C
        for (int row = tid; row < 10 + FC; row += RB_THREADS) {
            float accw[IN4], accb = 0.f;
#pragma unroll
            for (int i = 0; i < IN4; ++i) accw[i] = 0.f;
            for (int ray = 0; ray < RB_RAYS; ++ray) {
                const float g = row < 10 ? sGsm[ray * SMS + row] : sGb[ray * GBS + (row - 10)];
                accb += g;
#pragma unroll
                for (int i = 0; i < IN_C; ++i) accw[i] = fmaf(g, sF[ray * SFS + i], accw[i]);
            }
            const int woff = row < 10 ? L.small_w + row * IN4 : L.bott_w + (row - 10) * IN4;
            const int boff = row < 10 ? L.small_b + row : L.bott_b + (row - 10);
#pragma unroll
            for (int i = 0; i < IN_C; ++i) atomicAdd(G + woff + i, accw[i]);
            atomicAdd(G + boff, accb);
        }
        // specular layer: M[o][i] = sum_r g_sp[r][o] F[r][i], s[o] = sum_r g_sp[r][o]   (kept in sGF's tail? no: registers)
        __shared__ float sM[3][IN4 + 1];
        if (tid < 3 * (IN_C + 1)) {
            const int o = tid / (IN_C + 1), i = tid - o * (IN_C + 1);
            float s = 0.f;
            for (int ray = 0; ray < RB_RAYS; ++ray) s = fmaf(sGsp[ray * SPS + o], i < IN_C ? sF[ray * SFS + i] : 1.f, s);
            sM[o][i] = s;                       // column IN_C holds s[o]
        }
        __syncthreads();
        for (int e = tid; e < 3 * SW; e += RB_THREADS) {
            const int o = e / SW, j = e - o * SW;
            float g = 0.f;
            if (j < FC) {
                g = s_par[L.bott_b + j] * sM[o][IN_C];
                for (int i = 0; i < IN_C; ++i) g = fmaf(s_par[L.bott_w + j * IN4 + i], sM[o][i], g);
            } else {
                for (int ray = 0; ray < RB_RAYS; ++ray) g = fmaf(sGsp[ray * SPS + o], sXt[ray * XTS + (j - FC)], g);
            }
            atomicAdd(G + L.spec_w + e, g);
        }
        if (tid < 3) atomicAdd(G + L.spec_b + tid, sM[tid][IN_C]);
    }
    if (a.g_basis) {
        for (int e = tid; e < IN_C * (ta >> 2); e += RB_THREADS) {
            const int i = e / (ta >> 2), c4 = e - i * (ta >> 2);
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int ray = 0; ray < RB_RAYS; ++ray) {
                const float g = sGF[ray * SFS + i];
                const float4 x = *reinterpret_cast<const float4*>(sRf + ray * RFS + 4 * c4);
                s.x = fmaf(g, x.x, s.x); s.y = fmaf(g, x.y, s.y); s.z = fmaf(g, x.z, s.z); s.w = fmaf(g, x.w, s.w);
            }
            atomicAdd(reinterpret_cast<float4*>(a.g_basis + (size_t)i * ta) + c4, s);
        }
    }
}

}  // namespace

extern "C" int tvm_shade_ref_bwd(const tvm_field_desc* desc, const tvm_ref_head* head, const float* rays, int64_t n_rays,
                                 int ray_stride, const float* bg, const float* ray_feat, const float* acc,
                                 const int32_t* app_count, const float* d_rgb, const float* d_acc_in, float* d_ray_feat,
                                 float* d_acc, float* d_view, float* g_basis, float* g_params, void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (!head) return TVM_E_NULL;
    if (n_rays == 0) return 0;
    if (!rays || !bg || !ray_feat || !acc || !app_count || !d_rgb || !d_ray_feat || !d_acc || !desc->basis || !head->params)
        return TVM_E_NULL;
    if (ray_stride < 6 || head->in_c != desc->app_dim || head->in_c != 27 || head->n_pairs <= 0 ||
        2 * head->n_pairs + 1 > RB_XT || head->l_max <= 0 || head->l_max > REF_MAX_L || head->feature_c <= 0)
        return TVM_E_SHAPE;
    for (int p = 0; p < head->n_pairs; ++p)
        if (head->m[p] < 0 || head->m[p] > head->l_max || head->l[p] < 0) return TVM_E_SHAPE;
    RefBwdArgs a{};
    a.f = *desc; a.h = *head; a.rays = rays; a.n = n_rays; a.ray_stride = ray_stride; a.bg = bg;
    a.ray_feat = ray_feat; a.acc = acc; a.app_count = app_count; a.d_rgb = d_rgb; a.d_acc_in = d_acc_in;
    a.d_ray_feat = d_ray_feat; a.d_acc = d_acc; a.d_view = d_view; a.g_basis = g_basis; a.g_params = g_params;
    a.ta = tvm_total_app(desc);
    const RefLayout L = ref_layout(*head);
    const int in4 = L.in4;
    const size_t floats = (size_t)L.total + (size_t)head->in_c * a.ta + (size_t)RB_RAYS * (2 * (in4 + 1) + 13 + 5 + (RB_XT + 1) + (head->feature_c + 1) + (a.ta + 4));
    const size_t smem = floats * sizeof(float);
    if (smem > 220 * 1024) return TVM_E_SHAPE;
    static TvmDevMemo smem_set;
    int rc_attr = tvm_ensure_dyn_smem(shade_ref_bwd_kernel<27>, smem, smem_set);
    if (rc_attr) return rc_attr;
    const unsigned ctas = (unsigned)((n_rays + RB_RAYS - 1) / RB_RAYS);
    tvm_count_launch(); shade_ref_bwd_kernel<27><<<ctas, RB_THREADS, smem, (cudaStream_t)stream>>>(a);
    TVM_LAUNCH_CHECK();
    return 0;
}
