// TEST INFRASTRUCTURE ONLY — host build of the kernels' per-sample math.
//
// Compiles iffnerf_b200/csrc/tvm_math.cuh + tvm_gather.cuh (the SAME source the CUDA kernels
// include) with g++ -ffp-contract=off and walks rays sequentially, so the bit-exact sample mask,
// the tap/layout math and the compositing recurrences can be compared against the oracle on a
// machine without a GPU.  It is not linked into libtvm_b200.so and nothing in the product loads it.
#include <cstring>
#include <vector>
#include "../../iffnerf_b200/csrc/tvm_gather.cuh"

static int g_point_samples = 0;     // 1: sample_point_color sampler (TVM_F_POINT_SAMPLES)
extern "C" void hc_set_point_samples(int on) { g_point_samples = on; }
static int g_app_octets = 0;        // 1: appearance accumulated like app_gather_kernel does (quad runs over the ray's list, register texel cache)
extern "C" void hc_set_app_octets(int on) { g_app_octets = on; }

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

// coarse[cz][cy][cx]: any non-zero cell code inside the 16^3 super-cell (mirror of occupancy_coarse_kernel)
extern "C" void hc_pack_coarse(const uint8_t* cells, int dx, int dy, int dz, uint8_t* coarse) {
    const int cx = (dx + 15) / 16, cy = (dy + 15) / 16, cz = (dz + 15) / 16;
    memset(coarse, 0, (size_t)cx * cy * cz);
    for (int z = 0; z < dz; ++z)
        for (int y = 0; y < dy; ++y)
            for (int x = 0; x < dx; ++x)
                if (cells[((size_t)z * dy + y) * dx + x]) coarse[((size_t)(z >> 4) * cy + (y >> 4)) * cx + (x >> 4)] = 1;
}

// flags[r][b] = tvm_block_may_be_valid for 32-sample block b of ray r (the kernels' empty-space skip test)
extern "C" void hc_block_flags(const tvm_field_desc* f, const float* rays, long long n, int stride, int S,
                               const float* jitter, uint8_t* flags) {
    const int nblk = (S + 31) / 32;
    for (long long r = 0; r < n; ++r) {
        TvmRay ray;
        for (int c = 0; c < 3; ++c) { ray.o[c] = rays[r * stride + c]; ray.d[c] = rays[r * stride + 3 + c]; }
        tvm_init_ray(*f, ray, jitter ? jitter[r] : 0.f, S, g_point_samples != 0);
        for (int b = 0; b < nblk; ++b)
            flags[r * nblk + b] = tvm_block_may_be_valid(*f, ray, b * 32, (b * 32 + 31 < S - 1) ? b * 32 + 31 : S - 1);
    }
}

extern "C" void hc_sample_mask(const tvm_field_desc* f, const float* rays, long long n, int stride, int S,
                               const float* jitter, uint32_t* bits, int32_t* counts) {
    const int words = (S + 31) / 32;
    for (long long r = 0; r < n; ++r) {
        TvmRay ray;
        for (int c = 0; c < 3; ++c) { ray.o[c] = rays[r * stride + c]; ray.d[c] = rays[r * stride + 3 + c]; }
        tvm_init_ray(*f, ray, jitter ? jitter[r] : 0.f, S, g_point_samples != 0);
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
    bool lego = true;      // exercise the same template specialisation the CUDA dispatch picks
    for (int k = 0; k < 3; ++k) lego = lego && f->n_sigma[k] == 16 && f->n_app[k] == 48;
    for (long long r = 0; r < n; ++r) {
        TvmRay ray;
        for (int c = 0; c < 3; ++c) { ray.o[c] = rays[r * stride + c]; ray.d[c] = rays[r * stride + 3 + c]; }
        tvm_init_ray(*f, ray, jitter ? jitter[r] : 0.f, S, g_point_samples != 0);
        float T = 1.f, acc = 0.f, dep = 0.f;
        int napp = 0;
        float4 A[4][3][3];
        memset(A, 0, sizeof(A));
        // run mode: per-lane accumulators of the 32 emulated lanes + the ray's compacted appearance list
        // (what the march kernel emits and app_gather_kernel walks: quad q takes the contiguous eighth q of the list)
        float4 AL[32][3][3];
        memset(AL, 0, sizeof(AL));
        std::vector<float4> sw(S);
        std::vector<unsigned> si(S);
        int na = 0;
        const TvmSections sec = tvm_sections(*f);
        for (int i = 0; i < S; ++i) {
            float p[3], nrm[3];
            const float z = tvm_sample_z(*f, ray, i);
            bool keep = tvm_sample_point(*f, ray, z, p);
            if (keep && f->occ_cells) keep = tvm_occupancy_keep(*f, p);
            float alpha = 0.f;
            if (keep) {
                tvm_normalize(*f, p, nrm);
                float part[4];
                for (int sub = 0; sub < 4; ++sub) part[sub] = lego ? density_partial<4>(*f, nrm, sub) : density_partial<0>(*f, nrm, sub);
                const float feat = (part[0] + part[1]) + (part[2] + part[3]);
                const float sigma = tvm_density(*f, feat);
                const float dist = (i < S - 1) ? rn_sub(tvm_sample_z(*f, ray, i + 1), z) : 0.f;
                alpha = 1.f - expf(-sigma * rn_mul(dist, f->distance_scale));
                const float w = alpha * T;
                acc += w;
                dep += w * z;
                if (w > f->weight_thres) {
                    ++napp;
                    if (g_app_octets) {
                        float idx[3];
                        for (int c = 0; c < 3; ++c) idx[c] = tvm_unnormalize(nrm[c], f->grid[c]);
                        tvm_slot_from_idx(*f, idx, w, sw[na], si[na]);
                        ++na;
                    } else {
                        for (int sub = 0; sub < 4; ++sub) { if (lego) app_accumulate<3, 12>(*f, nrm, w, sub, A[sub]); else app_accumulate<3, 0>(*f, nrm, w, sub, A[sub]); }
                    }
                }
                T *= (1.f - alpha + 1e-10f);
            }
            if (alpha_out) alpha_out[r * S + i] = alpha;
        }
        if (g_app_octets && na > 0) {
            const int R = (na + 7) >> 3;
            for (int lane = 0; lane < 32; ++lane) {
                const int quad = lane >> 2, sub = lane & 3;
                const int b = quad * R, e = (b + R < na) ? b + R : na;
                for (int k = 0; k < 3; ++k) {
                    if (lego) app_run_plane<3, 12>(*f, sec, k, sw.data(), si.data(), b, e, sub, AL[lane][k]);
                    else app_run_plane<3, 0>(*f, sec, k, sw.data(), si.data(), b, e, sub, AL[lane][k]);
                }
            }
        }
        if (g_app_octets)
            for (int lane = 0; lane < 32; ++lane)
                for (int k = 0; k < 3; ++k)
                    for (int g = 0; g < 3; ++g) {
                        float4& d = A[lane & 3][k][g];
                        const float4 v = AL[lane][k][g];
                        d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w;
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

// sequential restatement of march_bwd_kernel: gradients of (ray_feat, acc, alpha) -> packed factor grads, d(rays)
extern "C" void hc_march_bwd(const tvm_field_desc* f, const float* rays, long long n, int stride, int S,
                             const float* jitter, const float* ray_feat, const float* acc_fwd,
                             const float* d_ray_feat, const float* d_acc, const float* d_alpha, float* g_factors,
                             float* g_rays) {
    const int ta = f->n_app[0] + f->n_app[1] + f->n_app[2];
    const int off[3] = {0, f->n_app[0], f->n_app[0] + f->n_app[1]};
    for (long long r = 0; r < n; ++r) {
        TvmRay ray;
        for (int c = 0; c < 3; ++c) { ray.o[c] = rays[r * stride + c]; ray.d[c] = rays[r * stride + 3 + c]; }
        tvm_init_ray(*f, ray, jitter ? jitter[r] : 0.f, S, g_point_samples != 0);
        float4 gF[4][3][3];
        memset(gF, 0, sizeof(gF));
        float total = 0.f;
        for (int k = 0; k < 3; ++k)
            for (int sub = 0; sub < 4; ++sub)
                for (int g = 0; g < 3; ++g) {
                    const int j = sub + 4 * g;
                    if (d_ray_feat && j < (f->n_app[k] >> 2)) {
                        memcpy(&gF[sub][k][g], d_ray_feat + r * ta + off[k] + 4 * j, 16);
                        float4 Fv;
                        memcpy(&Fv, ray_feat + r * ta + off[k] + 4 * j, 16);
                        total += f4_dot(gF[sub][k][g], Fv);
                    }
                }
        const float g_acc = d_acc ? d_acc[r] : 0.f;
        total += g_acc * acc_fwd[r];
        float T = 1.f, run = 0.f;
        float go[3] = {0, 0, 0}, gd[3] = {0, 0, 0};
        for (int i = 0; i < S; ++i) {
            float p[3], nrm[3];
            const float z = tvm_sample_z(*f, ray, i);
            bool keep = tvm_sample_point(*f, ray, z, p);
            if (keep && f->occ_cells) keep = tvm_occupancy_keep(*f, p);
            if (!keep) continue;
            tvm_normalize(*f, p, nrm);
            float part[4];
            for (int sub = 0; sub < 4; ++sub) part[sub] = density_partial(*f, nrm, sub);
            const float feat = (part[0] + part[1]) + (part[2] + part[3]);
            const float sigma = tvm_density(*f, feat);
            const float dist = (i < S - 1) ? rn_sub(tvm_sample_z(*f, ray, i + 1), z) : 0.f;
            const float delta = rn_mul(dist, f->distance_scale);
            const float alpha = 1.f - expf(-sigma * delta);
            const float one_m = 1.f - alpha + 1e-10f;
            const float w = alpha * T;
            float c = g_acc;
            if (w > f->weight_thres && d_ray_feat) {
                float dn[3] = {0, 0, 0};
                for (int sub = 0; sub < 4; ++sub) {
                    if (g_factors && g_rays) c += app_bwd<3, true, true>(*f, nrm, w, sub, gF[sub], g_factors, dn);
                    else if (g_factors) c += app_bwd<3, true, false>(*f, nrm, w, sub, gF[sub], g_factors, dn);
                    else c += app_bwd<3, false, true>(*f, nrm, w, sub, gF[sub], g_factors, dn);
                }
                for (int cc = 0; cc < 3; ++cc) { const float dp = dn[cc] * f->inv_aabb[cc]; go[cc] += dp; gd[cc] += dp * z; }
            }
            run += w * c;
            const float suffix = total - run;
            float dalpha = T * c - suffix / one_m;
            if (d_alpha) dalpha += d_alpha[r * S + i];
            const float dfeat = dalpha * delta * (1.f - alpha) * tvm_density_grad(*f, feat);
            T *= one_m;
            if (dfeat != 0.f) {
                float dn[3] = {0, 0, 0};
                for (int sub = 0; sub < 4; ++sub) {
                    if (g_factors && g_rays) density_bwd<true, true>(*f, nrm, dfeat, sub, g_factors, dn);
                    else if (g_factors) density_bwd<true, false>(*f, nrm, dfeat, sub, g_factors, dn);
                    else density_bwd<false, true>(*f, nrm, dfeat, sub, g_factors, dn);
                }
                for (int cc = 0; cc < 3; ++cc) { const float dp = dn[cc] * f->inv_aabb[cc]; go[cc] += dp; gd[cc] += dp * z; }
            }
        }
        if (g_rays) {
            const float gt0 = go[0] * ray.d[0] + go[1] * ray.d[1] + go[2] * ray.d[2];
            float best = -INFINITY, bv = 1.f; int bc = 0; bool bzero = false;
            for (int cc = 0; cc < 3; ++cc) {
                const bool zero = ray.d[cc] == 0.0f;
                const float v = zero ? 1e-6f : ray.d[cc];
                const float m = fminf(rn_div(rn_sub(f->aabb[3 + cc], ray.o[cc]), v), rn_div(rn_sub(f->aabb[cc], ray.o[cc]), v));
                if (m > best) { best = m; bc = cc; bv = v; bzero = zero; }
            }
            if (!g_point_samples && best >= f->near_t && best <= f->far_t) {
                go[bc] -= gt0 / bv;
                if (!bzero) gd[bc] -= gt0 * best / bv;
            }
            for (int cc = 0; cc < 3; ++cc) { g_rays[r * 6 + cc] = go[cc]; g_rays[r * 6 + 3 + cc] = gd[cc]; }
        }
    }
}
