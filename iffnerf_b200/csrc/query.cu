// query.cu — point queries of the density field (no rays): the kernels behind
//   TensorBase.compute_alpha            (models/tensorBase.py:756-773)  alpha = 1 - exp(-sigma * length), gated by the alphaMask
//   TensorVMSplit.compute_densityfeature (models/tensoRF.py:216-235)    raw sigma feature at normalised coordinates
// used by the pose pipeline's surface sampling (pose_estimation/sampling.py:138,172) and by the occupancy-grid
// rebuild getDenseAlpha/updateAlphaMask (tensorBase.py:643-696; 8 M lattice points per rebuild).
// Points are arbitrary (not restricted to the box), so the taps here are the general zero-padded ones.
// Mapping: one quad per point (lane = float4 channel slice), 8 points per warp pass, grid-stride over points.
#include "tvm_common.cuh"
#include "tvm_gather.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct QueryArgs {
    tvm_field_desc f;
    const float* pts;      // [m][3]
    long long m;
    int mode;              // 0: pts are normalised coords -> raw feature; 1: pts are world coords -> alpha(length)
    float length;
    float* out;            // [m]
};

// zero-padded (general) version of density_partial: out-of-range taps contribute 0 (F.grid_sample padding_mode=zeros)
__device__ __forceinline__ float density_partial_general(const tvm_field_desc& f, const float n[3], int sub) {
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
            const float4 a = __ldg(P + (ty.i0 * W + tx.i0) * C4), b = __ldg(P + (ty.i0 * W + tx.i1) * C4);
            const float4 c = __ldg(P + (ty.i1 * W + tx.i0) * C4), d = __ldg(P + (ty.i1 * W + tx.i1) * C4);
            const float4 l0 = __ldg(L + tl.i0 * C4), l1 = __ldg(L + tl.i1 * C4);
            float4 pl = f4_scale(tx.w0 * ty.w0, a);
            pl = f4_fma(tx.w1 * ty.w0, b, pl); pl = f4_fma(tx.w0 * ty.w1, c, pl); pl = f4_fma(tx.w1 * ty.w1, d, pl);
            float4 ln = f4_scale(tl.w0, l0);
            ln = f4_fma(tl.w1, l1, ln);
            tot += f4_dot(pl, ln);
        }
    }
    return tot;
}

__global__ void __launch_bounds__(256) point_density_kernel(const __grid_constant__ QueryArgs a) {
    const tvm_field_desc& f = a.f;
    const int lane = threadIdx.x & 31, sub = lane & 3;
    const long long quad0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 2;
    for (long long base = quad0 - (lane >> 2); base < a.m; base += stride) {     // warp-uniform trip count
        const long long i = base + (lane >> 2);
        const bool live = i < a.m;
        float p[3] = {0.f, 0.f, 0.f}, n[3];
        if (live) { p[0] = __ldg(a.pts + i * 3); p[1] = __ldg(a.pts + i * 3 + 1); p[2] = __ldg(a.pts + i * 3 + 2); }
        bool keep = live;
        if (a.mode == 1) {
            if (keep && f.occ_cells != nullptr) keep = tvm_occupancy_keep(f, p);
            tvm_normalize(f, p, n);
        } else {
            n[0] = p[0]; n[1] = p[1]; n[2] = p[2];
        }
        float part = keep ? density_partial_general(f, n, sub) : 0.f;
        part += __shfl_xor_sync(FULL, part, 1);
        part += __shfl_xor_sync(FULL, part, 2);
        if (live && sub == 0) {
            float v = part;
            if (a.mode == 1) v = keep ? 1.f - expf(-tvm_density(f, part) * a.length) : 0.f;
            a.out[i] = v;
        }
    }
}

// compute_appfeature (models/tensoRF.py:237-256) at arbitrary normalised points: zero-padded taps, the 3 x n_app
// plane (x) line products stay in the quad's registers and basis_mat ([app_dim][sum n_app], row-major) is applied
// in place: each lane dots its own channels with every row, two xor-shuffles finish the row.
struct AppQueryArgs {
    tvm_field_desc f;
    const float* pts;      // [m][3] normalised coordinates
    long long m;
    float* out;            // [m][app_dim]
    int ta;
    int app_off[3];
};

__global__ void __launch_bounds__(128) point_appfeature_kernel(const __grid_constant__ AppQueryArgs a) {
    const tvm_field_desc& f = a.f;
    const int lane = threadIdx.x & 31, sub = lane & 3;
    const long long quad0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 2;
    for (long long base = quad0 - (lane >> 2); base < a.m; base += stride) {     // warp-uniform trip count
        const long long i = base + (lane >> 2);
        const bool live = i < a.m;
        float n[3] = {0.f, 0.f, 0.f};
        if (live) { n[0] = __ldg(a.pts + i * 3); n[1] = __ldg(a.pts + i * 3 + 1); n[2] = __ldg(a.pts + i * 3 + 2); }
        TvmTap t[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] = tvm_axis_tap(n[c], f.grid[c]);
        float4 v[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int C4 = f.n_app[k] >> 2;
            const TvmTap& tx = t[TVM_M0(k)];
            const TvmTap& ty = t[TVM_M1(k)];
            const TvmTap& tl = t[TVM_V(k)];
            const int W = tvm_plane_pitch(f.grid[TVM_M0(k)]);
            const float4* P = reinterpret_cast<const float4*>(f.factors + f.aplane_off[k]);
            const float4* L = reinterpret_cast<const float4*>(f.factors + f.aline_off[k]);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const int j = sub + 4 * g;
                v[k][g] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && j < C4) {
                    const float4 q00 = __ldg(P + (ty.i0 * W + tx.i0) * C4 + j), q01 = __ldg(P + (ty.i0 * W + tx.i1) * C4 + j);
                    const float4 q10 = __ldg(P + (ty.i1 * W + tx.i0) * C4 + j), q11 = __ldg(P + (ty.i1 * W + tx.i1) * C4 + j);
                    const float4 l0 = __ldg(L + tl.i0 * C4 + j), l1 = __ldg(L + tl.i1 * C4 + j);
                    float4 pl = f4_scale(tx.w0 * ty.w0, q00);
                    pl = f4_fma(tx.w1 * ty.w0, q01, pl); pl = f4_fma(tx.w0 * ty.w1, q10, pl); pl = f4_fma(tx.w1 * ty.w1, q11, pl);
                    float4 ln = f4_scale(tl.w0, l0);
                    ln = f4_fma(tl.w1, l1, ln);
                    v[k][g] = f4_mul(pl, ln);
                }
            }
        }
        for (int r = 0; r < f.app_dim; ++r) {
            const float* row = f.basis + (long long)r * a.ta;
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const int j = sub + 4 * g;
                    if (j < (f.n_app[k] >> 2))
                        dot += f4_dot(v[k][g], __ldg(reinterpret_cast<const float4*>(row + a.app_off[k]) + j));
                }
            dot += __shfl_xor_sync(FULL, dot, 1);
            dot += __shfl_xor_sync(FULL, dot, 2);
            if (live && sub == 0) a.out[i * f.app_dim + r] = dot;
        }
    }
}

}  // namespace

extern "C" int tvm_point_appfeature(const tvm_field_desc* desc, const float* points, int64_t n_points, float* out,
                                    void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (n_points == 0) return 0;
    if (!points || !out || !desc->factors || !desc->basis) return TVM_E_NULL;
    if (desc->app_dim <= 0) return TVM_E_SHAPE;
    AppQueryArgs a{};
    a.f = *desc; a.pts = points; a.m = n_points; a.out = out;
    a.ta = tvm_total_app(desc);
    a.app_off[0] = 0; a.app_off[1] = desc->n_app[0]; a.app_off[2] = desc->n_app[0] + desc->n_app[1];
    const long long quads_per_cta = 128 / 4;
    long long ctas = (n_points + quads_per_cta - 1) / quads_per_cta;
    if (ctas > TVM_SM_COUNT * 16) ctas = TVM_SM_COUNT * 16;
    tvm_count_launch(); point_appfeature_kernel<<<(unsigned)ctas, 128, 0, (cudaStream_t)stream>>>(a);
    TVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int tvm_point_density(const tvm_field_desc* desc, const float* points, int64_t n_points, int mode,
                                 float length, float* out, void* stream) {
    int rc = tvm_check_desc(desc);
    if (rc) return rc;
    if (n_points == 0) return 0;
    if (!points || !out || !desc->factors) return TVM_E_NULL;
    if (mode != 0 && mode != 1) return TVM_E_MODE;
    QueryArgs a{};
    a.f = *desc; a.pts = points; a.m = n_points; a.mode = mode; a.length = length; a.out = out;
    const long long quads_per_cta = 256 / 4;
    long long ctas = (n_points + quads_per_cta - 1) / quads_per_cta;
    if (ctas > TVM_SM_COUNT * 16) ctas = TVM_SM_COUNT * 16;
    tvm_count_launch(); point_density_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(a);
    TVM_LAUNCH_CHECK();
    return 0;
}
