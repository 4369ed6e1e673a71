// microbench.cu — measured ceilings for the gather stages ("fraction of a MEASURED L2 gather
// peak: random 64 B and 192 B granule reads over the resident factor set").
// Each quad (4 lanes x 16 B) fetches pseudo-random granules of `granule_bytes` (64 or 192) from a buffer with the
// size of the packed factor set, which stays L2-resident (69 MB < 126 MB L2) but defeats L1 (random, 28 MB of L1
// in total).  The result is the L2->SM random-gather bandwidth the march kernel would be bound by if it had no L1
// reuse at all; its achieved algorithmic rate is reported against it next to the HBM figure.
#include "tvm_common.cuh"

namespace {

__global__ void __launch_bounds__(256) gather_bench_kernel(const float4* __restrict__ buf, unsigned long long n_granules,
                                                           int f4_per_granule, int iters, float* __restrict__ sink) {
    const int lane = threadIdx.x & 31, sub = lane & 3;
    unsigned long long quad = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    unsigned long long state = quad * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        // 4 independent granules per iteration (memory-level parallelism like the kernel's 4 bilinear corners)
        unsigned long long g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            g[u] = (state >> 20) % n_granules;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            for (int j = sub; j < f4_per_granule; j += 4) {
                const float4 v = __ldg(buf + g[u] * f4_per_granule + j);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;      // keep the loads alive
}

}  // namespace

// Launches the gather micro-benchmark over `bytes` of `buf`; returns the bytes it requested in *bytes_moved.
extern "C" int tvm_gather_microbench(const void* buf, size_t bytes, int granule_bytes, int iters, float* sink,
                                     unsigned long long* bytes_moved, void* stream) {
    if (!buf || !sink) return TVM_E_NULL;
    if (granule_bytes % 16 || granule_bytes <= 0 || bytes < (size_t)granule_bytes) return TVM_E_SHAPE;
    const unsigned long long n_granules = bytes / granule_bytes;
    const int ctas = TVM_SM_COUNT * 16, threads = 256;
    gather_bench_kernel<<<ctas, threads, 0, (cudaStream_t)stream>>>((const float4*)buf, n_granules, granule_bytes / 16,
                                                                   iters, sink);
    TVM_LAUNCH_CHECK();
    if (bytes_moved) *bytes_moved = (unsigned long long)ctas * threads / 4 * iters * 4ull * granule_bytes;
    return 0;
}
