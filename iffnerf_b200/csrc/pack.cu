// pack.cu — layout conversion between the reference's parameter layout and the kernel layout,
// plus the small ABI utilities (version, error strings, workspace sizes).
//
//   * factors: reference stores planes as [1,C,H,W] and lines as [1,C,L,1] (models/tensoRF.py:160-170),
//     i.e. each (texel, channel) in its own 32-B sector.  The kernels want channel-last [H*W][C] so a
//     texel is one contiguous 64-B / 192-B run.  Both directions are a tiled [C][P] <-> [P][C] transpose
//     through shared memory (coalesced on both sides, HBM-bound: 2 x 69 MB at 300^3).
//   * occupancy: AlphaGridMask's {0,1} fp32 volume (models/tensorBase.py:50-64) -> one byte per cell
//     holding the occupancy of its 8 corners (see tvm_occupancy_keep in tvm_math.cuh).
//   * MLP: torch Linear weights [out][in] -> transposed [in][out] rows for the shade kernels.
#include "tvm_common.cuh"

namespace {

constexpr int TP = 128;  // pixels per tile

// The 12 factor tensors are converted by ONE launch: blockIdx.y selects the tensor, blockIdx.x the pixel tile.
// P pixels in rows of W; the channel-last side pads each row to `pitch` texels (tvm_plane_pitch; lines: W = pitch = P)
struct TransposeJob { const float* src; float* dst; int C; long long P; int W; int pitch; };
// position of pixel p0 + dp in the padded layout, given the (row, column) of the tile's first pixel: one division per
// thread per tile instead of one per element (the unpack runs in every training step over 17 M elements)
struct TileOrigin { long long y0; int x0; };
__device__ __forceinline__ TileOrigin tile_origin(long long p0, int W) {
    TileOrigin t;
    t.y0 = p0 / W;
    t.x0 = (int)(p0 - t.y0 * W);
    return t;
}
__device__ __forceinline__ long long padded_pixel(const TileOrigin& t, long long p0, int dp, int W, int pitch) {
    if (pitch == W) return p0 + dp;
    int x = t.x0 + dp;
    long long y = t.y0;
    while (x >= W) { x -= W; ++y; }
    return y * pitch + x;
}
struct TransposeJobs { TransposeJob j[12]; int accumulate; float scale; int rezero; };

// src [C][P] -> dst [P][C].  The channel-last side moves as float4s (C is a multiple of 4 and every texel starts on a
// 16-byte boundary), the [C][P] side as 128-pixel rows: 512 contiguous bytes per channel.
__global__ void __launch_bounds__(256) cp_to_pc_kernel(const __grid_constant__ TransposeJobs jobs) {
    __shared__ float tile[TVM_MAX_APP_C][TP + 1];
    const float* __restrict__ src = jobs.j[blockIdx.y].src;
    float* __restrict__ dst = jobs.j[blockIdx.y].dst;
    const int C = jobs.j[blockIdx.y].C;
    const long long P = jobs.j[blockIdx.y].P;
    const long long p0 = (long long)blockIdx.x * TP;
    if (p0 >= P) return;
    const int npix = (int)min((long long)TP, P - p0);
    for (int i = threadIdx.x; i < C * TP; i += 256) {
        const int c = i / TP, px = i - c * TP;
        if (px < npix) tile[c][px] = __ldg(src + (long long)c * P + p0 + px);
    }
    __syncthreads();
    const int W = jobs.j[blockIdx.y].W, pitch = jobs.j[blockIdx.y].pitch, C4 = C >> 2;
    const TileOrigin org = tile_origin(p0, W);
    for (int i = threadIdx.x; i < npix * C4; i += 256) {
        const int px = i / C4, c = (i - px * C4) * 4;
        reinterpret_cast<float4*>(dst + padded_pixel(org, p0, px, W, pitch) * C)[c >> 2] =
            make_float4(tile[c][px], tile[c + 1][px], tile[c + 2][px], tile[c + 3][px]);
    }
}

// src [P][C] -> dst [C][P]  (dst = scale * src^T, or dst += ...; optionally leaves src zeroed)
__global__ void __launch_bounds__(256) pc_to_cp_kernel(const __grid_constant__ TransposeJobs jobs) {
    __shared__ float tile[TVM_MAX_APP_C][TP + 1];
    const float* __restrict__ src = jobs.j[blockIdx.y].src;
    float* __restrict__ dst = jobs.j[blockIdx.y].dst;
    if (dst == nullptr) return;
    const int C = jobs.j[blockIdx.y].C, accumulate = jobs.accumulate;
    const long long P = jobs.j[blockIdx.y].P;
    const long long p0 = (long long)blockIdx.x * TP;
    if (p0 >= P) return;
    const int npix = (int)min((long long)TP, P - p0);
    const int W = jobs.j[blockIdx.y].W, pitch = jobs.j[blockIdx.y].pitch, C4 = C >> 2;
    const TileOrigin org = tile_origin(p0, W);
    for (int i = threadIdx.x; i < npix * C4; i += 256) {
        const int px = i / C4, c = (i - px * C4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + padded_pixel(org, p0, px, W, pitch) * C) + (c >> 2));
        tile[c][px] = jobs.scale * v.x; tile[c + 1][px] = jobs.scale * v.y;
        tile[c + 2][px] = jobs.scale * v.z; tile[c + 3][px] = jobs.scale * v.w;
    }
    __syncthreads();
    if (jobs.rezero) {            // leave the scatter buffer zeroed for the next step (stores only: a separate loop keeps
        float* z = const_cast<float*>(src);     // the loads above independent of them, several in flight per thread)
        for (int i = threadIdx.x; i < npix * C4; i += 256) {
            const int px = i / C4, c4 = i - px * C4;
            reinterpret_cast<float4*>(z + padded_pixel(org, p0, px, W, pitch) * C)[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (int i = threadIdx.x; i < C * TP; i += 256) {
        const int c = i / TP, px = i - c * TP;
        if (px < npix) {
            float* d = dst + (long long)c * P + p0 + px;
            *d = accumulate ? (*d + tile[c][px]) : tile[c][px];
        }
    }
}

__global__ void occupancy_cells_kernel(const float* __restrict__ vol, int dx, int dy, int dz,
                                       uint8_t* __restrict__ cells) {
    const long long n = (long long)dx * dy * dz;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % dx), y = (int)((i / dx) % dy), z = (int)(i / ((long long)dx * dy));
    unsigned code = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int xx = x + (b & 1), yy = y + ((b >> 1) & 1), zz = z + (b >> 2);
        if (xx < dx && yy < dy && zz < dz && __ldg(vol + ((long long)zz * dy + yy) * dx + xx) > 0.f) code |= 1u << b;
    }
    cells[i] = (uint8_t)code;
}

// one warp per 16^3 super-cell: 1 if any cell code inside is non-zero
__global__ void occupancy_coarse_kernel(const uint8_t* __restrict__ cells, int dx, int dy, int dz, int cx, int cy,
                                        int cz, uint8_t* __restrict__ coarse) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= cx * cy * cz) return;
    const int X = warp % cx, Y = (warp / cx) % cy, Z = warp / (cx * cy);
    bool any = false;
    for (int i = lane; i < 4096; i += 32) {
        const int x = X * 16 + (i & 15), y = Y * 16 + ((i >> 4) & 15), z = Z * 16 + (i >> 8);
        if (x < dx && y < dy && z < dz && cells[((size_t)z * dy + y) * dx + x]) any = true;
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) coarse[warp] = any ? 1 : 0;
}

// dst [cols][rows_padded...]: generic small transpose  src [R][Cc] -> dst [Cc (padded to kpad)][R]
__global__ void transpose_small_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc, int kpad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kpad * R) return;
    const int k = i / R, r = i - k * R;
    dst[i] = (k < Cc) ? __ldg(src + r * Cc + k) : 0.f;
}
// inverse: packed [kpad][R] -> torch [R][Cc], optional accumulate
__global__ void untranspose_small_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc,
                                         int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * Cc) return;
    const int r = i / Cc, k = i - r * Cc;
    const float v = __ldg(src + k * R + r);
    dst[i] = accumulate ? dst[i] + v : v;
}
// src [R][Cc] -> dst [R][kpad] (zero padded columns)
__global__ void pad_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc, int kpad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * kpad) return;
    const int r = i / kpad, k = i - r * kpad;
    dst[i] = (k < Cc) ? __ldg(src + r * Cc + k) : 0.f;
}
__global__ void copy_small_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int npad,
                                  int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    if (i < n) dst[i] = accumulate ? dst[i] + __ldg(src + i) : __ldg(src + i);
    else if (!accumulate) dst[i] = 0.f;
}

struct FactorView { long long off; int C; long long P; int W; int pitch; };

void factor_views(const tvm_field_desc* d, FactorView planes[6], FactorView lines[6]) {
    for (int k = 0; k < 3; ++k) {
        const long long hw = (long long)d->grid[TVM_M0(k)] * d->grid[TVM_M1(k)];
        const long long l = d->grid[TVM_V(k)];
        const int W = d->grid[TVM_M0(k)], pitch = tvm_plane_pitch(W);
        planes[k] = {d->dplane_off[k], d->n_sigma[k], hw, W, pitch};
        planes[3 + k] = {d->aplane_off[k], d->n_app[k], hw, W, pitch};
        lines[k] = {d->dline_off[k], d->n_sigma[k], l, (int)l, (int)l};
        lines[3 + k] = {d->aline_off[k], d->n_app[k], l, (int)l, (int)l};
    }
}

}  // namespace

std::atomic<unsigned long long> g_tvm_launch_count{0};
extern "C" unsigned long long tvm_launch_count(void) { return g_tvm_launch_count.load(std::memory_order_relaxed); }
extern "C" int tvm_abi_version(void) { return TVM_ABI_VERSION; }

extern "C" const char* tvm_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case TVM_E_NULL: return "tvm: required pointer is NULL";
        case TVM_E_SHAPE: return "tvm: unsupported shape (channels must be multiples of 4, sigma<=16, app<=48, featureC==128)";
        case TVM_E_WORKSPACE: return "tvm: workspace too small";
        case TVM_E_MODE: return "tvm: unsupported mode";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "tvm: unknown error";
    }
}

extern "C" int tvm_pack_factors(const tvm_field_desc* desc, const float* const planes[6], const float* const lines[6],
                                float* packed, void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (!planes || !lines || !packed) return TVM_E_NULL;
    FactorView pv[6], lv[6];
    factor_views(desc, pv, lv);
    TransposeJobs jobs{};
    jobs.scale = 1.0f;
    long long maxP = 0;
    for (int i = 0; i < 6; ++i) {
        if (!planes[i] || !lines[i]) return TVM_E_NULL;
        jobs.j[i] = {planes[i], packed + pv[i].off, pv[i].C, pv[i].P, pv[i].W, pv[i].pitch};
        jobs.j[6 + i] = {lines[i], packed + lv[i].off, lv[i].C, lv[i].P, lv[i].W, lv[i].pitch};
        maxP = max(maxP, max(pv[i].P, lv[i].P));
    }
    tvm_count_launch(); cp_to_pc_kernel<<<dim3((unsigned)((maxP + TP - 1) / TP), 12), 256, 0, (cudaStream_t)stream>>>(jobs);
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int tvm_unpack_factor_grads(const tvm_field_desc* desc, const float* packed_grad, float* const planes[6],
                                       float* const lines[6], int accumulate, void* stream) {
    return tvm_unpack_factor_grads_scaled(desc, const_cast<float*>(packed_grad), planes, lines, accumulate, 1.0f, 0, stream);
}

extern "C" int tvm_unpack_factor_grads_scaled(const tvm_field_desc* desc, float* packed_grad, float* const planes[6],
                                              float* const lines[6], int accumulate, float scale, int rezero,
                                              void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (!planes || !lines || !packed_grad) return TVM_E_NULL;
    FactorView pv[6], lv[6];
    factor_views(desc, pv, lv);
    TransposeJobs jobs{};
    jobs.accumulate = accumulate;
    jobs.scale = scale;
    jobs.rezero = rezero;
    long long maxP = 0;
    for (int i = 0; i < 6; ++i) {
        jobs.j[i] = {packed_grad + pv[i].off, planes[i], pv[i].C, pv[i].P, pv[i].W, pv[i].pitch};
        jobs.j[6 + i] = {packed_grad + lv[i].off, lines[i], lv[i].C, lv[i].P, lv[i].W, lv[i].pitch};
        maxP = max(maxP, max(pv[i].P, lv[i].P));
    }
    tvm_count_launch(); pc_to_cp_kernel<<<dim3((unsigned)((maxP + TP - 1) / TP), 12), 256, 0, (cudaStream_t)stream>>>(jobs);
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t tvm_occupancy_coarse_offset(int dx, int dy, int dz) {
    return tvm_align((size_t)dx * dy * dz, 16);
}
extern "C" size_t tvm_occupancy_bytes(int dx, int dy, int dz) {
    return tvm_occupancy_coarse_offset(dx, dy, dz) + (size_t)((dx + 15) / 16) * ((dy + 15) / 16) * ((dz + 15) / 16);
}

extern "C" int tvm_pack_occupancy(const float* volume, int dx, int dy, int dz, uint8_t* cells, void* stream) {
    if (!volume || !cells) return TVM_E_NULL;
    if (dx < 1 || dy < 1 || dz < 1) return TVM_E_SHAPE;
    const long long n = (long long)dx * dy * dz;
    tvm_count_launch(); occupancy_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(volume, dx, dy, dz, cells);
    const int cx = (dx + 15) / 16, cy = (dy + 15) / 16, cz = (dz + 15) / 16;
    const int warps = cx * cy * cz;
    tvm_count_launch(); occupancy_coarse_kernel<<<(warps * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        cells, dx, dy, dz, cx, cy, cz, cells + tvm_occupancy_coarse_offset(dx, dy, dz));
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t tvm_mlp_pack_floats(const tvm_field_desc* desc) {
    if (!desc) return 0;
    return tvm_mlp_layout(desc).total;
}

extern "C" int tvm_pack_mlp(const tvm_field_desc* desc, const float* w1, const float* b1, const float* w2,
                            const float* b2, const float* w3, const float* b3, float* packed, void* stream) {
    if (!desc || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !packed) return TVM_E_NULL;
    if (desc->feature_c != TVM_FEATURE_C) return TVM_E_SHAPE;
    const TvmMlpLayout m = tvm_mlp_layout(desc);
    const int FC = TVM_FEATURE_C;
    cudaStream_t st = (cudaStream_t)stream;
    tvm_count_launch(); transpose_small_kernel<<<(m.k1 * FC + 255) / 256, 256, 0, st>>>(w1, packed + m.w1t, FC, m.in_c, m.k1);
    tvm_count_launch(); copy_small_kernel<<<1, 256, 0, st>>>(b1, packed + m.b1, FC, FC, 0);
    tvm_count_launch(); transpose_small_kernel<<<(FC * FC + 255) / 256, 256, 0, st>>>(w2, packed + m.w2t, FC, FC, FC);
    tvm_count_launch(); copy_small_kernel<<<1, 256, 0, st>>>(b2, packed + m.b2, FC, FC, 0);
    tvm_count_launch(); copy_small_kernel<<<2, 256, 0, st>>>(w3, packed + m.w3, 3 * FC, 3 * FC, 0);
    tvm_count_launch(); copy_small_kernel<<<1, 32, 0, st>>>(b3, packed + m.b3, 3, 4, 0);
    tvm_count_launch(); pad_cols_kernel<<<(FC * m.k1 + 255) / 256, 256, 0, st>>>(w1, packed + m.w1n, FC, m.in_c, m.k1);
    tvm_count_launch(); copy_small_kernel<<<(FC * FC + 255) / 256, 256, 0, st>>>(w2, packed + m.w2n, FC * FC, FC * FC, 0);
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t tvm_mlp_grad_floats(const tvm_field_desc* desc) {
    if (!desc) return 0;
    return tvm_mlp_layout(desc).grad_total;
}

extern "C" int tvm_unpack_mlp_grads(const tvm_field_desc* desc, const float* packed_grad, float* w1, float* b1,
                                    float* w2, float* b2, float* w3, float* b3, int accumulate, void* stream) {
    if (!desc || !packed_grad) return TVM_E_NULL;
    if (desc->feature_c != TVM_FEATURE_C) return TVM_E_SHAPE;
    const TvmMlpLayout m = tvm_mlp_layout(desc);
    const int FC = TVM_FEATURE_C;
    cudaStream_t st = (cudaStream_t)stream;
    if (w1) { tvm_count_launch(); untranspose_small_kernel<<<(FC * m.in_c + 255) / 256, 256, 0, st>>>(packed_grad + m.w1t, w1, FC, m.in_c, accumulate); }
    if (b1) { tvm_count_launch(); copy_small_kernel<<<1, 256, 0, st>>>(packed_grad + m.b1, b1, FC, FC, accumulate); }
    if (w2) { tvm_count_launch(); untranspose_small_kernel<<<(FC * FC + 255) / 256, 256, 0, st>>>(packed_grad + m.w2t, w2, FC, FC, accumulate); }
    if (b2) { tvm_count_launch(); copy_small_kernel<<<1, 256, 0, st>>>(packed_grad + m.b2, b2, FC, FC, accumulate); }
    if (w3) { tvm_count_launch(); copy_small_kernel<<<2, 256, 0, st>>>(packed_grad + m.w3, w3, 3 * FC, 3 * FC, accumulate); }
    if (b3) { tvm_count_launch(); copy_small_kernel<<<1, 32, 0, st>>>(packed_grad + m.b3, b3, 3, 3, accumulate); }
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int tvm_workspace_bytes(const tvm_field_desc* desc, int64_t n_rays, uint32_t flags, size_t* out) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (!out) return TVM_E_NULL;
    *out = tvm_ws_layout(desc, n_rays, flags).total;
    return 0;
}

extern "C" int tvm_workspace_layout(const tvm_field_desc* desc, int64_t n_rays, size_t* ray_feat_off, size_t* acc_off,
                                    size_t* depth_off, size_t* sigma_count_off, size_t* app_count_off,
                                    size_t* occ_count_off) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    const TvmWorkspace w = tvm_ws_layout(desc, n_rays);
    if (ray_feat_off) *ray_feat_off = w.ray_feat;
    if (acc_off) *acc_off = w.acc;
    if (depth_off) *depth_off = w.depth;
    if (sigma_count_off) *sigma_count_off = w.sigma_count;
    if (app_count_off) *app_count_off = w.app_count;
    if (occ_count_off) *occ_count_off = w.occ_count;
    return 0;
}
