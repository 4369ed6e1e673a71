// raygen.cu — fused pixel -> ray generation with gradients to the camera pose.
//
// Replaces, for the pixels a step actually renders, the reference's full-image ray build:
//   get_ray_directions_Ks (ray_utils.py:28-60): K^-1 [x+.5, y+.5, 1] for the pixel and its +1 x / +1 y neighbours
//   directions = ori / |ori|                                   (inerf/estimate_pose_inerf.py:96-99)
//   get_rays (ray_utils.py:63-100): rays_d = R directions, rays_o = t,
//       radii = 0.5 (|R dx - R ori| + |R dy - R ori|) * 2/sqrt(12)
//   rays_d = F.normalize(rays_d[pixels])                        (estimate_pose_inerf.py:156-164)
//   rays_chunk = cat(rays_o, rays_d, radii)                     (:161)
// The reference materialises five [H,W,3] grids per optimisation step and then indexes 1024 pixels; here one
// thread builds one 7-column ray, and the backward reduces d(rays) into the 3x4 pose gradient per candidate pose
// (the autograd edge of `pose = cam_transf()` -> get_rays, :149-156).
#include "tvm_common.cuh"

namespace {

struct RayGenArgs {
    const float* c2w;        // [P][pose_stride] row-major, rows 0..2 of a [3|4][4] matrix are read
    int pose_stride;         // floats between poses (12 or 16)
    float kinv[9];           // K^-1 row-major
    const int* pixels;       // [n][2] (x, y) or NULL: pixel i = (i % width, i / width)
    const int* pose_index;   // [n] or NULL (pose 0)
    int width;
    long long n;
    int normalize_dirs;      // bit 0: viewdirs = ori/|ori| before the rotation (blender.py:70-72, estimate_pose_inerf.py:97);
                             // bit 1: F.normalize(rays_d) after it (estimate_pose_inerf.py:159)
};

struct PixelGeom {
    float v[3];      // camera-space direction handed to the rotation (normalised or not)
    float ori[3];    // K^-1 [x+.5, y+.5, 1]
    float ex[3];     // K^-1 [x+1.5, y+.5, 1]  (the +1 x neighbour's direction, absolute)
    float ey[3];     // K^-1 [x+.5, y+1.5, 1]
    float inv_len;   // 1/|ori| (1 when not normalising)
};

__device__ __forceinline__ void kinv_apply(const float* k, float x, float y, float out[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)      // un-fused like the reference's matmul + the 1.0 homogeneous coordinate
        out[r] = rn_add(rn_add(rn_mul(k[3 * r], x), rn_mul(k[3 * r + 1], y)), k[3 * r + 2]);
}

__device__ __forceinline__ PixelGeom pixel_geom(const RayGenArgs& a, long long i) {
    int px, py;
    if (a.pixels) { px = a.pixels[2 * i]; py = a.pixels[2 * i + 1]; }
    else { px = (int)(i % a.width); py = (int)(i / a.width); }
    const float x = (float)px + 0.5f, y = (float)py + 0.5f;
    PixelGeom g;
    float dx[3], dy[3];
    kinv_apply(a.kinv, x, y, g.ori);
    kinv_apply(a.kinv, x + 1.0f, y, dx);
    kinv_apply(a.kinv, x, y + 1.0f, dy);
    const float len = sqrtf(g.ori[0] * g.ori[0] + g.ori[1] * g.ori[1] + g.ori[2] * g.ori[2]);
    g.inv_len = (a.normalize_dirs & 1) ? 1.0f / len : 1.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g.v[c] = (a.normalize_dirs & 1) ? g.ori[c] / len : g.ori[c];
        g.ex[c] = dx[c];
        g.ey[c] = dy[c];
    }
    return g;
}

__device__ __forceinline__ void rotate(const float* R /* row-major, row stride 4 */, const float v[3], float out[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)      // (v[..., None, :] * R).sum(-1), ray_utils.py:76 — separate roundings (radii are a
                                     // difference of nearly equal vectors, so contraction noise would show there)
        out[r] = rn_add(rn_add(rn_mul(v[0], R[4 * r]), rn_mul(v[1], R[4 * r + 1])), rn_mul(v[2], R[4 * r + 2]));
}

constexpr float RADII_SCALE = 0.5f * 0.57735026918962576f;      // 0.5 * 2/sqrt(12)

__global__ void __launch_bounds__(256) pixel_rays_fwd_kernel(const __grid_constant__ RayGenArgs a, float* __restrict__ rays) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float* M = a.c2w + (long long)(a.pose_index ? a.pose_index[i] : 0) * a.pose_stride;
    const PixelGeom g = pixel_geom(a, i);
    float u[3], w0[3], wx[3], wy[3];
    rotate(M, g.v, u);
    rotate(M, g.ori, w0);
    rotate(M, g.ex, wx);
    rotate(M, g.ey, wy);
    const float un = (a.normalize_dirs & 2) ? fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f) : 1.0f;  // F.normalize eps
    float ax = 0.f, ay = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float dxc = wx[c] - w0[c], dyc = wy[c] - w0[c];
        ax += dxc * dxc;
        ay += dyc * dyc;
    }
    float* out = rays + i * 7;
    out[0] = M[3]; out[1] = M[7]; out[2] = M[11];
    out[3] = u[0] / un; out[4] = u[1] / un; out[5] = u[2] / un;
    out[6] = (sqrtf(ax) + sqrtf(ay)) * RADII_SCALE;
}

// d(rays) [n][g_stride] -> g_c2w [P][3][4] (accumulated).  Per ray the 12 contributions are
//   d t      = g_o
//   d R[r][c] = g_u[r] v[c] + g_rad * RADII_SCALE * (ax_hat[r] (ex-ori)[c] + ay_hat[r] (ey-ori)[c])
// with g_u = (g_d - d_hat (d_hat . g_d)) / |u| (backward of F.normalize).  Warps whose lanes share a pose reduce
// with shuffles and issue 12 atomics; mixed warps fall back to per-lane atomics.
__global__ void __launch_bounds__(256) pixel_rays_bwd_kernel(const __grid_constant__ RayGenArgs a, const float* __restrict__ g_rays,
                                                             int g_stride, float* __restrict__ g_c2w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.n;
    const int pose = (live && a.pose_index) ? a.pose_index[i] : 0;
    float c[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) c[k] = 0.f;
    if (live) {
        const float* M = a.c2w + (long long)pose * a.pose_stride;
        const PixelGeom g = pixel_geom(a, i);
        float u[3], w0[3], wx[3], wy[3];
        rotate(M, g.v, u);
        rotate(M, g.ori, w0);
        rotate(M, g.ex, wx);
        rotate(M, g.ey, wy);
        const float* gr = g_rays + i * g_stride;
        const float len = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        const float un = fmaxf(len, 1e-12f);
        const float dh[3] = {u[0] / un, u[1] / un, u[2] / un};
        const float dot = dh[0] * gr[3] + dh[1] * gr[4] + dh[2] * gr[5];
        float gu[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            gu[r] = !(a.normalize_dirs & 2) ? gr[3 + r] : ((len > 1e-12f) ? (gr[3 + r] - dh[r] * dot) / un : gr[3 + r] / un);
        const float grad_rad = (g_stride > 6 ? gr[6] : 0.f) * RADII_SCALE;
        float dxv[3], dyv[3], ax = 0.f, ay = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            dxv[r] = wx[r] - w0[r]; dyv[r] = wy[r] - w0[r];
            ax += dxv[r] * dxv[r]; ay += dyv[r] * dyv[r];
        }
        const float ix = ax > 0.f ? grad_rad / sqrtf(ax) : 0.f, iy = ay > 0.f ? grad_rad / sqrtf(ay) : 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                c[4 * r + cc] = gu[r] * g.v[cc] + ix * dxv[r] * (g.ex[cc] - g.ori[cc]) + iy * dyv[r] * (g.ey[cc] - g.ori[cc]);
            c[4 * r + 3] = gr[r];
        }
    }
    const unsigned FULL = 0xffffffffu;
    const int pose0 = __shfl_sync(FULL, pose, 0);
    const bool uniform = __all_sync(FULL, !live || pose == pose0);
    if (uniform) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            float v = c[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(g_c2w + (long long)pose0 * 12 + k, v);
        }
    } else if (live) {
#pragma unroll
        for (int k = 0; k < 12; ++k) atomicAdd(g_c2w + (long long)pose * 12 + k, c[k]);
    }
}

int fill(RayGenArgs& a, const float* c2w, int pose_stride, const float* kinv_host, const int32_t* pixels,
         const int32_t* pose_index, int width, int64_t n, int normalize_dirs) {
    if (!c2w || !kinv_host) return TVM_E_NULL;
    if (pose_stride < 12 || n < 0 || (!pixels && width <= 0)) return TVM_E_SHAPE;
    a.c2w = c2w; a.pose_stride = pose_stride;
    for (int k = 0; k < 9; ++k) a.kinv[k] = kinv_host[k];
    a.pixels = pixels; a.pose_index = pose_index; a.width = width; a.n = n; a.normalize_dirs = normalize_dirs;
    return 0;
}

}  // namespace

extern "C" int tvm_pixel_rays_fwd(const float* c2w, int pose_stride, const float* kinv_host, const int32_t* pixels,
                                  const int32_t* pose_index, int width, int64_t n, int normalize_dirs, float* rays,
                                  void* stream) {
    RayGenArgs a;
    int rc = fill(a, c2w, pose_stride, kinv_host, pixels, pose_index, width, n, normalize_dirs);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!rays) return TVM_E_NULL;
    tvm_count_launch(); pixel_rays_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, rays);
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int tvm_pixel_rays_bwd(const float* c2w, int pose_stride, const float* kinv_host, const int32_t* pixels,
                                  const int32_t* pose_index, int width, int64_t n, int normalize_dirs,
                                  const float* g_rays, int g_stride, float* g_c2w, void* stream) {
    RayGenArgs a;
    int rc = fill(a, c2w, pose_stride, kinv_host, pixels, pose_index, width, n, normalize_dirs);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!g_rays || !g_c2w) return TVM_E_NULL;
    if (g_stride < 6) return TVM_E_SHAPE;
    tvm_count_launch(); pixel_rays_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, g_rays, g_stride, g_c2w);
    TVM_LAUNCH_CHECK();
    return 0;
}
