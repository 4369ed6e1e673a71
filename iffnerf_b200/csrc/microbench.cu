// microbench.cu — measured ceilings for the gather stages ("fraction of a MEASURED gather peak:
// random 64 B and 192 B granule reads over the resident factor set").
// Each quad (4 lanes x 16 B, LDG.128 like the march kernel) fetches pseudo-random granules of `granule_bytes`
// (64 = density texel, 192 = appearance texel) from `bytes` of `buf`:
//   * bytes = the packed factor set (69 MB): L2-resident, random => defeats L1: the L2 -> SM gather ceiling;
//   * bytes = a few KB: every SM re-reads an L1-resident set: the L1 data-pipe ceiling for this access shape
//     (8 different texels per warp request, ~6 wavefronts per LDG.128 including bank conflicts).
// The loop is load-dense (about 8 instructions per LDG.128: 32-bit LCG, multiply-high range reduction, one
// IMAD.WIDE address, 4 FADDs) and keeps 8 independent loads in flight per lane, so the memory pipe, not the
// issue slots, is what saturates.
#include "tvm_common.cuh"

namespace {

__global__ void __launch_bounds__(256) gather_bench_kernel(const float4* __restrict__ buf, unsigned n_granules,
                                                           int f4_per_granule, int iters, float* __restrict__ sink) {
    const unsigned sub = threadIdx.x & 3;
    const unsigned quad = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    unsigned state = quad * 0x9E3779B9u + 0x1234567u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            state = state * 1664525u + 1013904223u;
            const unsigned g = __umulhi(state, n_granules);
            // one slice per lane per granule: a 64-B granule is read once by the quad, a 192-B granule in 3 rounds
            const unsigned j = sub + 4u * ((unsigned)u % (unsigned)(f4_per_granule >> 2));
            v[u] = __ldg(buf + (g * (unsigned)f4_per_granule + j));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;      // keep the loads alive
}

}  // namespace

// Launches the gather micro-benchmark over `bytes` of `buf`; returns the bytes it requested in *bytes_moved.
extern "C" int tvm_gather_microbench(const void* buf, size_t bytes, int granule_bytes, int iters, float* sink,
                                     unsigned long long* bytes_moved, void* stream) {
    if (!buf || !sink) return TVM_E_NULL;
    if (granule_bytes % 16 || granule_bytes <= 0 || bytes < (size_t)granule_bytes) return TVM_E_SHAPE;
    if (granule_bytes % 64 || bytes / granule_bytes > 0xffffffffull / (granule_bytes / 16)) return TVM_E_SHAPE;
    const unsigned n_granules = (unsigned)(bytes / granule_bytes);
    const int ctas = TVM_SM_COUNT * 16, threads = 256;
    gather_bench_kernel<<<ctas, threads, 0, (cudaStream_t)stream>>>((const float4*)buf, n_granules, granule_bytes / 16,
                                                                   iters, sink);
    TVM_LAUNCH_CHECK();
    // every lane requests 16 B per load, 8 loads per iteration
    if (bytes_moved) *bytes_moved = (unsigned long long)ctas * threads * (unsigned long long)iters * 8ull * 16ull;
    return 0;
}
