// tvm_common.cuh — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include "tvm_math.cuh"

#define TVM_CUDA_OK(expr)                         \
    do {                                          \
        cudaError_t e__ = (expr);                 \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

#define TVM_LAUNCH_CHECK()                        \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when the requested size grows past what this process already
// set for the kernel.  The attribute is per-process CUDA state anyway; skipping the redundant calls keeps a warmed-up
// launch path free of non-stream API calls, so the entry points can be captured into CUDA graphs.
// The attribute is per DEVICE, so the high-water mark is kept per device (a process may render on several GPUs).
#include <atomic>
struct TvmDevMemo {
    std::atomic<int> v[16];
    TvmDevMemo() { for (auto& x : v) x.store(0, std::memory_order_relaxed); }
};
template <typename K>
static inline int tvm_ensure_dyn_smem(K kernel, size_t bytes, TvmDevMemo& memo) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<int>& high_water = memo.v[dev & 15];
    if ((int)bytes <= high_water.load(std::memory_order_relaxed)) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    high_water.store((int)bytes, std::memory_order_relaxed);
    return 0;
}

// kernels launched by the library since it was loaded (tvm_launch_count(): bench.py reports the launches of its timed
// region from this counter instead of a hand-maintained constant)
extern std::atomic<unsigned long long> g_tvm_launch_count;
static inline void tvm_count_launch() { g_tvm_launch_count.fetch_add(1, std::memory_order_relaxed); }

// Output placement of the forward shading kernels (device side of tvm_scatter_out): with n == 0 results go to the
// kernel's own rgb / depth arrays at the ray's index; otherwise ray r is written at offset dst_index[r] (identity when
// NULL) of EVERY destination — the image buffers of all ranks of a ray-sharded render, mapped through NVLink peer
// memory, so the shading epilogue is also the all-gather (iffnerf_b200/sharding.py).
struct TvmPeers {
    const long long* dst_index;
    int n;
    float* rgb[TVM_MAX_PEERS];
    float* depth[TVM_MAX_PEERS];
};
#if defined(__CUDACC__)
__device__ __forceinline__ void tvm_put_rgb(const TvmPeers& p, float* rgb, long long r, int c, float v) {
    if (p.n == 0) { rgb[r * 3 + c] = v; return; }
    const long long o = p.dst_index ? __ldg(p.dst_index + r) : r;
    for (int i = 0; i < p.n; ++i) p.rgb[i][o * 3 + c] = v;
}
__device__ __forceinline__ void tvm_put_depth(const TvmPeers& p, float* depth, long long r, float v) {
    if (p.n == 0) { if (depth) depth[r] = v; return; }
    const long long o = p.dst_index ? __ldg(p.dst_index + r) : r;
    for (int i = 0; i < p.n; ++i) p.depth[i][o] = v;
}
#endif
static inline int tvm_fill_peers(TvmPeers& p, const tvm_scatter_out* sc) {
    p = TvmPeers{};
    if (!sc) return 0;
    if (sc->n_dst <= 0 || sc->n_dst > TVM_MAX_PEERS) return TVM_E_SHAPE;
    p.dst_index = (const long long*)sc->dst_index;
    p.n = sc->n_dst;
    for (int i = 0; i < sc->n_dst; ++i) {
        if (!sc->rgb[i] || !sc->depth[i]) return TVM_E_NULL;
        p.rgb[i] = sc->rgb[i];
        p.depth[i] = sc->depth[i];
    }
    return 0;
}

constexpr int TVM_SM_COUNT = 148;        // B200: 2 dies x 74 SMs
constexpr int TVM_MAX_SIGMA_C = 16;
constexpr int TVM_MAX_APP_C = 48;
constexpr int TVM_FEATURE_C = 128;       // hidden width the shade kernels are specialised for

static inline int tvm_total_app(const tvm_field_desc* d) { return d->n_app[0] + d->n_app[1] + d->n_app[2]; }
static inline int tvm_mlp_in(const tvm_field_desc* d) {
    // MLPRender_Fea.in_mlpC (models/tensorBase.py:169)
    return 2 * d->view_pe * 3 + 2 * d->fea_pe * d->app_dim + 3 + d->app_dim;
}
static inline int tvm_round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline size_t tvm_align(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Workspace layout for n rays (all sections 256-B aligned):
//   ray_feat [n][TA] f32 | acc [n] f32 | depth [n] f32 | sigma_count [n] i32 (density samples evaluated)
//   | app_count [n] i32 (appearance samples evaluated) | occ_count [n] i32 (occupancy tests = in-aabb samples visited)
//   with TVM_F_SPLIT_APP additionally (at the end, so the common offsets do not move):
//   | app_list [n][TVM_APP_CAP] (int sample index, float weight)   (per-ray appearance sample lists, 8 B per entry)
struct TvmWorkspace {
    size_t ray_feat, acc, depth, sigma_count, app_count, occ_count, app_list, total;
};
static inline TvmWorkspace tvm_ws_layout(const tvm_field_desc* d, int64_t n, uint32_t flags = 0) {
    TvmWorkspace w;
    size_t off = 0;
    w.ray_feat = off;    off = tvm_align(off + (size_t)n * tvm_total_app(d) * sizeof(float));
    w.acc = off;         off = tvm_align(off + (size_t)n * sizeof(float));
    w.depth = off;       off = tvm_align(off + (size_t)n * sizeof(float));
    w.sigma_count = off; off = tvm_align(off + (size_t)n * sizeof(int32_t));
    w.app_count = off;   off = tvm_align(off + (size_t)n * sizeof(int32_t));
    w.occ_count = off;   off = tvm_align(off + (size_t)n * sizeof(int32_t));
    w.app_list = off;
    if (flags & TVM_F_SPLIT_APP) off = tvm_align(off + (size_t)n * TVM_APP_CAP * 2 * sizeof(float));
    w.total = off;
    return w;
}

// Packed MLP layout (floats): W1^T [k1][FC] (rows >= in_c are zero) | b1 [FC] | W2^T [FC][FC] | b2 [FC] | W3 [3][FC] | b3 [4]
//   | W1 [FC][k1] (cols >= in_c zero) | W2 [FC][FC]
struct TvmMlpLayout {
    int in_c, k1;
    size_t w1t, b1, w2t, b2, w3, b3, grad_total, w1n, w2n, total;
};
static inline TvmMlpLayout tvm_mlp_layout(const tvm_field_desc* d) {
    TvmMlpLayout m;
    m.in_c = tvm_mlp_in(d);
    m.k1 = tvm_round_up(m.in_c, 4);
    const size_t FC = TVM_FEATURE_C;
    size_t off = 0;
    m.w1t = off; off += (size_t)m.k1 * FC;
    m.b1 = off;  off += FC;
    m.w2t = off; off += FC * FC;
    m.b2 = off;  off += FC;
    m.w3 = off;  off += 3 * FC;
    m.b3 = off;  off += 4;
    m.grad_total = off;                 // the gradient buffer of tvm_shade_bwd covers [0, grad_total)
    m.w1n = off; off += FC * (size_t)m.k1;   // W1 [FC][k1] and W2 [FC][FC] in torch orientation (backward GEMMs)
    m.w2n = off; off += FC * FC;
    m.total = off;
    return m;
}

static inline int tvm_check_desc(const tvm_field_desc* d) {
    if (!d) return TVM_E_NULL;
    for (int k = 0; k < 3; ++k) {
        if (d->n_sigma[k] <= 0 || d->n_sigma[k] % 4 || d->n_sigma[k] > TVM_MAX_SIGMA_C) return TVM_E_SHAPE;
        if (d->n_app[k] <= 0 || d->n_app[k] % 4 || d->n_app[k] > TVM_MAX_APP_C) return TVM_E_SHAPE;
        if (d->grid[k] < 2) return TVM_E_SHAPE;
        if ((d->dplane_off[k] | d->dline_off[k] | d->aplane_off[k] | d->aline_off[k]) & 3) return TVM_E_SHAPE;
    }
    if (d->n_factor_floats < 0 || d->n_factor_floats > (int64_t(1) << 33)) return TVM_E_SHAPE;   // 32-bit float4 indices
    if (d->act != 0 && d->act != 1) return TVM_E_MODE;
    return 0;
}
