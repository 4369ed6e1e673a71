// TEST INFRASTRUCTURE ONLY — host build of the kernels' per-sample math.
//
// Compiles iffnerf_b200/csrc/tvm_math.cuh + tvm_gather.cuh (the SAME source the CUDA kernels
// include) with g++ -ffp-contract=off and walks rays sequentially, so the bit-exact sample mask,
// the tap/layout math and the compositing recurrences can be compared against the oracle on a
// machine without a GPU.  It is not linked into libtvm_b200.so and nothing in the product loads it.
#include <cstring>
#include <vector>
#include "../../iffnerf_b200/csrc/tvm_gather.cuh"

extern "C" void hc_pack_occupancy(const float* vol, int dx, int dy, int dz, uint8_t* cells) {
    for (int z = 0; z < dz; ++z)
        for (int y = 0; y < dy; ++y)
            for (int x = 0; x < dx; ++x) {
                unsigned code = 0;
                for (int b = 0; b < 8; ++b) {
                    int xx = x + (b & 1), yy = y + ((b >> 1) & 1), zz = z + (b >> 2);
                    if (xx < dx && yy < dy && zz < dz && vol[((size_t)zz * dy + yy) * dx + xx] > 0.f) code |= 1u << b;
                }
                cells[((size_t)z * dy + y) * dx + x] = (uint8_t)code;
            }
}

extern "C" void hc_sample_mask(const tvm_field_desc* f, const float* rays, long long n, int stride, int S,
                               const float* jitter, uint32_t* bits, int32_t* counts) {
    const int words = (S + 31) / 32;
    for (long long r = 0; r < n; ++r) {
        TvmRay ray;
        for (int c = 0; c < 3; ++c) { ray.o[c] = rays[r * stride + c]; ray.d[c] = rays[r * stride + 3 + c]; }
        ray.t0 = tvm_ray_entry(*f, ray.o, ray.d);
        ray.jit = jitter ? jitter[r] : 0.f;
        int cnt = 0;
        if (bits) memset(bits + r * words, 0, words * sizeof(uint32_t));
        for (int i = 0; i < S; ++i) {
            float p[3];
            const float z = tvm_sample_z(*f, ray, i);
            bool keep = tvm_sample_point(*f, ray, z, p);
            if (keep && f->occ_cells) keep = tvm_occupancy_keep(*f, p);
            if (keep) { ++cnt; if (bits) bits[r * words + (i >> 5)] |= 1u << (i & 31); }
        }
        if (counts) counts[r] = cnt;
    }
}

// sequential restatement of march_fwd_kernel's per-ray recurrence (no early termination)
extern "C" void hc_march(const tvm_field_desc* f, const float* rays, long long n, int stride, int S,
                         const float* jitter, float* ray_feat, float* acc_out, float* depth_out, float* alpha_out,
                         int32_t* app_count) {
    const int ta = f->n_app[0] + f->n_app[1] + f->n_app[2];
    const int off[3] = {0, f->n_app[0], f->n_app[0] + f->n_app[1]};
    for (long long r = 0; r < n; ++r) {
        TvmRay ray;
        for (int c = 0; c < 3; ++c) { ray.o[c] = rays[r * stride + c]; ray.d[c] = rays[r * stride + 3 + c]; }
        ray.t0 = tvm_ray_entry(*f, ray.o, ray.d);
        ray.jit = jitter ? jitter[r] : 0.f;
        float T = 1.f, acc = 0.f, dep = 0.f;
        int napp = 0;
        float4 A[4][3][3];
        memset(A, 0, sizeof(A));
        for (int i = 0; i < S; ++i) {
            float p[3], nrm[3];
            const float z = tvm_sample_z(*f, ray, i);
            bool keep = tvm_sample_point(*f, ray, z, p);
            if (keep && f->occ_cells) keep = tvm_occupancy_keep(*f, p);
            float alpha = 0.f;
            if (keep) {
                tvm_normalize(*f, p, nrm);
                float part[4];
                for (int sub = 0; sub < 4; ++sub) part[sub] = density_partial(*f, nrm, sub);
                const float feat = (part[0] + part[1]) + (part[2] + part[3]);
                const float sigma = tvm_density(*f, feat);
                const float dist = (i < S - 1) ? rn_sub(tvm_sample_z(*f, ray, i + 1), z) : 0.f;
                alpha = 1.f - expf(-sigma * rn_mul(dist, f->distance_scale));
                const float w = alpha * T;
                acc += w;
                dep += w * z;
                if (w > f->weight_thres) {
                    ++napp;
                    for (int sub = 0; sub < 4; ++sub) app_accumulate<3>(*f, nrm, w, sub, A[sub]);
                }
                T *= (1.f - alpha + 1e-10f);
            }
            if (alpha_out) alpha_out[r * S + i] = alpha;
        }
        for (int k = 0; k < 3; ++k)
            for (int sub = 0; sub < 4; ++sub)
                for (int g = 0; g < 3; ++g) {
                    const int j = sub + 4 * g;
                    if (j < (f->n_app[k] >> 2)) memcpy(ray_feat + r * ta + off[k] + 4 * j, &A[sub][k][g], 16);
                }
        acc_out[r] = acc; depth_out[r] = dep;
        if (app_count) app_count[r] = napp;
    }
}
