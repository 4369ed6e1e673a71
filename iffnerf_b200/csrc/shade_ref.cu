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
        for (int o = 0; o < 3; ++o) a.rgb[r * 3 + o] = fminf(fmaxf(__ldg(a.bg + o) * (1.0f - acc0), 0.f), 1.f);
        a.depth[r] = a.depth_part[r] + (1.0f - acc0) * __ldg(rp + a.ray_stride - 1);
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
        a.rgb[r * 3 + o] = fminf(fmaxf(c, 0.f), 1.f);
    }
    a.depth[r] = a.depth_part[r] + (1.0f - acc) * __ldg(rp + a.ray_stride - 1);
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

extern "C" int tvm_shade_ref_fwd(const tvm_field_desc* desc, const tvm_ref_head* head, const float* rays,
                                 int64_t n_rays, int ray_stride, const float* bg, float* rgb, float* depth,
                                 float* acc, const void* ws, size_t ws_bytes, void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (!head) return TVM_E_NULL;
    if (n_rays == 0) return 0;
    if (!rays || !bg || !rgb || !depth || !ws || !desc->basis || !head->params) return TVM_E_NULL;
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
        static std::atomic<int> smem_set{0};
        int rc_attr = tvm_ensure_dyn_smem(shade_ref_kernel<27>, smem, smem_set);
        if (rc_attr) return rc_attr;
        shade_ref_kernel<27><<<ctas, REF_THREADS, smem, st>>>(a);
    } else {
        return TVM_E_SHAPE;          // app_dim = 27 is what every reference config uses (configs/*.txt)
    }
    TVM_LAUNCH_CHECK();
    return 0;
}
