// gridops.cu — grid maintenance of the TensoRF-VM field, the steps train.py:384-415 runs a few
// times per training between render steps:
//   getDenseAlpha / updateAlphaMask  (models/tensorBase.py:643-696)  -> tvm_dense_alpha_mask
//   up_sampling_VM / shrink          (models/tensoRF.py:258-316)     -> tvm_resize_factor
//   filtering_rays(bbox_only=True)   (models/tensorBase.py:716-726)  -> tvm_rays_hit_box
// (filtering_rays(bbox_only=False) is tvm_sample_mask with TVM_F_MASK_ANYWHERE, csrc/march.cu.)
#include "tvm_common.cuh"
#include "tvm_gather.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct DenseArgs {
    tvm_field_desc f;
    const float* lin[3];       // torch.linspace(0, 1, g[c]) per axis (host-generated: bit-identical to the reference's)
    int g[3];
    float length, thres;
    float* alpha;              // [gz][gy][gx] scratch: clamp(alpha, 0, 1) per lattice point
    float* volume;             // [gz][gy][gx] out: {0,1} after 3x3x3 max-pool + threshold
    int* box;                  // [6] ordered-int keys: min xyz | max xyz of the occupied lattice points
    unsigned long long* count; // occupied voxels
};

// order-preserving float <-> int map (atomicMin/Max on floats of either sign)
__device__ __forceinline__ int f2key(float v) { const int b = __float_as_int(v); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// dense_xyz = aabb[0] * (1 - samples) + aabb[1] * samples   (tensorBase.py:656), separate roundings
__device__ __forceinline__ float lattice_coord(const DenseArgs& a, int c, int i) {
    const float s = __ldg(a.lin[c] + i);
    return rn_add(rn_mul(a.f.aabb[c], rn_sub(1.0f, s)), rn_mul(a.f.aabb[3 + c], s));
}

// same arithmetic as query.cu's general (zero-padded) density gather
__device__ __forceinline__ float dense_density_partial(const tvm_field_desc& f, const float n[3], int sub) {
    float tot = 0.f;
    TvmTap t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = tvm_axis_tap(n[c], f.grid[c]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int C4 = f.n_sigma[k] >> 2;
        if (sub < C4) {
            const TvmTap& tx = t[TVM_M0(k)];
            const TvmTap& ty = t[TVM_M1(k)];
            const TvmTap& tl = t[TVM_V(k)];
            const int W = tvm_plane_pitch(f.grid[TVM_M0(k)]);
            const float4* P = reinterpret_cast<const float4*>(f.factors + f.dplane_off[k]) + sub;
            const float4* L = reinterpret_cast<const float4*>(f.factors + f.dline_off[k]) + sub;
            const float4 p00 = __ldg(P + (ty.i0 * W + tx.i0) * C4), p01 = __ldg(P + (ty.i0 * W + tx.i1) * C4);
            const float4 p10 = __ldg(P + (ty.i1 * W + tx.i0) * C4), p11 = __ldg(P + (ty.i1 * W + tx.i1) * C4);
            const float4 l0 = __ldg(L + tl.i0 * C4), l1 = __ldg(L + tl.i1 * C4);
            float4 pl = f4_scale(tx.w0 * ty.w0, p00);
            pl = f4_fma(tx.w1 * ty.w0, p01, pl); pl = f4_fma(tx.w0 * ty.w1, p10, pl); pl = f4_fma(tx.w1 * ty.w1, p11, pl);
            float4 ln = f4_scale(tl.w0, l0);
            ln = f4_fma(tl.w1, l1, ln);
            tot += f4_dot(pl, ln);
        }
    }
    return tot;
}

// stage 1: alpha = compute_alpha(dense_xyz, length).clamp(0, 1) on the lattice, stored [z][y][x] (the transpose(0,2) of
// tensorBase.py:670-671).  One quad per lattice point, x fastest.
__global__ void __launch_bounds__(256) dense_alpha_kernel(const __grid_constant__ DenseArgs a) {
    const tvm_field_desc& f = a.f;
    const int lane = threadIdx.x & 31, sub = lane & 3;
    const long long total = (long long)a.g[0] * a.g[1] * a.g[2];
    const long long quad0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 2;
    for (long long base = quad0 - (lane >> 2); base < total; base += stride) {       // warp-uniform trip count
        const long long i = base + (lane >> 2);
        const bool live = i < total;
        float p[3] = {0.f, 0.f, 0.f}, n[3];
        if (live) {
            const int x = (int)(i % a.g[0]), y = (int)((i / a.g[0]) % a.g[1]), z = (int)(i / ((long long)a.g[0] * a.g[1]));
            p[0] = lattice_coord(a, 0, x); p[1] = lattice_coord(a, 1, y); p[2] = lattice_coord(a, 2, z);
        }
        bool keep = live;
        if (keep && f.occ_cells != nullptr) keep = tvm_occupancy_keep(f, p);         // old mask gates the query (:757-761)
        tvm_normalize(f, p, n);
        float part = keep ? dense_density_partial(f, n, sub) : 0.f;
        part += __shfl_xor_sync(FULL, part, 1);
        part += __shfl_xor_sync(FULL, part, 2);
        if (live && sub == 0) {
            const float v = keep ? 1.f - expf(-tvm_density(f, part) * a.length) : 0.f;
            a.alpha[i] = fminf(fmaxf(v, 0.f), 1.f);
        }
    }
}

// stage 2: max_pool3d(k=3, pad=1, stride=1) + threshold (:676-680) + tight box of the occupied lattice points (:686-691)
__global__ void __launch_bounds__(256) dense_pool_kernel(const __grid_constant__ DenseArgs a) {
    const long long total = (long long)a.g[0] * a.g[1] * a.g[2];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int kmin[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, kmax[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    bool occ = false;
    if (i < total) {
        const int gx = a.g[0], gy = a.g[1], gz = a.g[2];
        const int x = (int)(i % gx), y = (int)((i / gx) % gy), z = (int)(i / ((long long)gx * gy));
        float m = -INFINITY;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || yy < 0 || zz < 0 || xx >= gx || yy >= gy || zz >= gz) continue;
                    m = fmaxf(m, __ldg(a.alpha + ((long long)zz * gy + yy) * gx + xx));
                }
        occ = m >= a.thres;
        a.volume[i] = occ ? 1.f : 0.f;
        if (occ) {
            const int c[3] = {x, y, z};
#pragma unroll
            for (int k = 0; k < 3; ++k) kmin[k] = kmax[k] = f2key(lattice_coord(a, k, c[k]));
        }
    }
    const unsigned any = __ballot_sync(FULL, occ);
    if (any) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int lo = kmin[k], hi = kmax[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(FULL, lo, o));
                hi = max(hi, __shfl_xor_sync(FULL, hi, o));
            }
            if ((threadIdx.x & 31) == 0) { atomicMin(a.box + k, lo); atomicMax(a.box + 3 + k, hi); }
        }
        if ((threadIdx.x & 31) == 0) atomicAdd(a.count, (unsigned long long)__popc(any));
    }
}

__global__ void dense_box_init_kernel(int* box, unsigned long long* count) {
    if (threadIdx.x < 3) box[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) box[threadIdx.x] = (int)0x80000000;
    if (threadIdx.x == 0) *count = 0ull;
}
__global__ void dense_box_final_kernel(const int* box, const unsigned long long* count, float* out) {
    if (threadIdx.x < 6) out[threadIdx.x] = key2f(box[threadIdx.x]);
    if (threadIdx.x == 6) out[6] = (float)(*count);
}

struct ResizeArgs {
    const float* src;
    float* dst;
    int C, H, W, H2, W2, mode, y_off, x_off;
};

// up_sampling_VM (F.interpolate bilinear, align_corners=True; ATen UpSample.h area_pixel_compute_source_index /
// compute_source_index_and_lambda) and the crop of shrink(), NCHW in -> NCHW out (fresh Parameter storage)
__global__ void __launch_bounds__(256) resize_kernel(const __grid_constant__ ResizeArgs a) {
    const long long total = (long long)a.C * a.H2 * a.W2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % a.W2), y = (int)((i / a.W2) % a.H2), c = (int)(i / ((long long)a.W2 * a.H2));
    const float* s = a.src + (long long)c * a.H * a.W;
    if (a.mode == 1) {
        a.dst[i] = __ldg(s + (long long)(y + a.y_off) * a.W + (x + a.x_off));
        return;
    }
    const float sh = a.H2 > 1 ? (float)(a.H - 1) / (float)(a.H2 - 1) : 0.f;
    const float sw = a.W2 > 1 ? (float)(a.W - 1) / (float)(a.W2 - 1) : 0.f;
    const float ry = sh * (float)y, rx = sw * (float)x;
    const int y0 = (int)ry, x0 = (int)rx;
    const int y1 = y0 + (y0 < a.H - 1 ? 1 : 0), x1 = x0 + (x0 < a.W - 1 ? 1 : 0);
    const float ly1 = fminf(fmaxf(ry - (float)y0, 0.f), 1.f), lx1 = fminf(fmaxf(rx - (float)x0, 0.f), 1.f);
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float v00 = __ldg(s + (long long)y0 * a.W + x0), v01 = __ldg(s + (long long)y0 * a.W + x1);
    const float v10 = __ldg(s + (long long)y1 * a.W + x0), v11 = __ldg(s + (long long)y1 * a.W + x1);
    a.dst[i] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
}

// filtering_rays(bbox_only=True): t_max > t_min of the slab test without the near/far clamp (tensorBase.py:716-726)
__global__ void __launch_bounds__(256) rays_hit_box_kernel(tvm_field_desc f, const float* __restrict__ rays, long long n,
                                                           int stride, uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float tmin = -INFINITY, tmax = INFINITY;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float o = __ldg(rays + i * stride + c), d = __ldg(rays + i * stride + 3 + c);
        const float v = (d == 0.0f) ? 1e-6f : d;
        const float ra = rn_div(rn_sub(f.aabb[3 + c], o), v), rb = rn_div(rn_sub(f.aabb[c], o), v);
        tmin = fmaxf(tmin, fminf(ra, rb));
        tmax = fminf(tmax, fmaxf(ra, rb));
    }
    out[i] = tmax > tmin ? 1 : 0;
}

}  // namespace

extern "C" size_t tvm_dense_alpha_workspace_bytes(int gx, int gy, int gz) {
    return tvm_align((size_t)gx * gy * gz * sizeof(float)) + 256;
}

extern "C" int tvm_dense_alpha_mask(const tvm_field_desc* desc, const float* lin_x, const float* lin_y, const float* lin_z,
                                    int gx, int gy, int gz, float length, float thres, float* volume, float* box_out,
                                    void* ws, size_t ws_bytes, void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (gx <= 0 || gy <= 0 || gz <= 0) return TVM_E_SHAPE;
    if (!lin_x || !lin_y || !lin_z || !volume || !box_out || !ws || !desc->factors) return TVM_E_NULL;
    if (ws_bytes < tvm_dense_alpha_workspace_bytes(gx, gy, gz)) return TVM_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    DenseArgs a{};
    a.f = *desc;
    a.lin[0] = lin_x; a.lin[1] = lin_y; a.lin[2] = lin_z;
    a.g[0] = gx; a.g[1] = gy; a.g[2] = gz;
    a.length = length; a.thres = thres;
    a.alpha = (float*)ws;
    a.volume = volume;
    char* tail = (char*)ws + tvm_align((size_t)gx * gy * gz * sizeof(float));
    a.box = (int*)tail;
    a.count = (unsigned long long*)(tail + 32);
    const long long total = (long long)gx * gy * gz;
    tvm_count_launch(); dense_box_init_kernel<<<1, 32, 0, st>>>(a.box, a.count);
    long long ctas = (total + 63) / 64;
    if (ctas > TVM_SM_COUNT * 16) ctas = TVM_SM_COUNT * 16;
    tvm_count_launch(); dense_alpha_kernel<<<(unsigned)ctas, 256, 0, st>>>(a);
    tvm_count_launch(); dense_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    tvm_count_launch(); dense_box_final_kernel<<<1, 32, 0, st>>>(a.box, a.count, box_out);
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int tvm_resize_factor(const float* src, int channels, int h, int w, float* dst, int h2, int w2, int mode,
                                 int y_off, int x_off, void* stream) {
    if (!src || !dst) return TVM_E_NULL;
    if (channels <= 0 || h <= 0 || w <= 0 || h2 <= 0 || w2 <= 0 || (mode != 0 && mode != 1)) return TVM_E_SHAPE;
    if (mode == 1 && (y_off < 0 || x_off < 0 || y_off + h2 > h || x_off + w2 > w)) return TVM_E_SHAPE;
    ResizeArgs a{src, dst, channels, h, w, h2, w2, mode, y_off, x_off};
    const long long total = (long long)channels * h2 * w2;
    tvm_count_launch(); resize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int tvm_rays_hit_box(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                                uint8_t* out, void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (n_rays == 0) return 0;
    if (!rays || !out) return TVM_E_NULL;
    if (ray_stride < 6) return TVM_E_SHAPE;
    tvm_count_launch(); rays_hit_box_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*desc, rays, n_rays,
                                                                                         ray_stride, out);
    TVM_LAUNCH_CHECK();
    return 0;
}
