// collective.cu — the one exchange step of the path: SUM all-reduce of the flat fp32 gradient workspace of
// data-parallel training, written against NVLink peer memory instead of calling NCCL.
//
// Every rank's workspace lives in symmetric memory (torch.distributed._symmetric_memory), so each GPU can address all of
// them.  Two-shot, one kernel: rank r owns the r-th slice of the buffer; it sums that slice over all ranks and stores
// the result into the slice of EVERY rank's buffer.
//   * NVSwitch multicast (NVLS) available: one multimem.ld_reduce per 16 bytes (the switch adds the N copies in flight)
//     and one multimem.st (the switch replicates the store) — per GPU ~1/N of the buffer in and out;
//   * otherwise: N peer loads + N peer stores per 16 bytes over NVLink.
// The caller brackets the launch with the symmetric-memory barrier (all scatters done before / all slices written after).
#include "tvm_common.cuh"

namespace {

struct AllReduceArgs {
    float4* peer[TVM_MAX_PEERS];
    float4* mc;               // multicast address of the same buffer, or nullptr
    long long begin4, end4;   // this rank's slice in float4 units
    int n_ranks;
};

__device__ __forceinline__ float4 ld_cg(const float4* p) { return __ldcg(p); }       // L2 only: peers rewrite these bytes

__global__ void __launch_bounds__(256) allreduce_slice_kernel(const __grid_constant__ AllReduceArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = a.begin4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.end4; i += stride) {
        if (a.mc != nullptr) {
            float4 v;
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a.mc + i) : "memory");
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                         :: "l"(a.mc + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        } else {
            float4 s = ld_cg(a.peer[0] + i);
#pragma unroll 1
            for (int p = 1; p < a.n_ranks; ++p) {
                const float4 v = ld_cg(a.peer[p] + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
#pragma unroll 1
            for (int p = 0; p < a.n_ranks; ++p) __stcg(a.peer[p] + i, s);
        }
    }
}

}  // namespace

extern "C" int tvm_allreduce_sum_peer(void* const* peer_bufs, int n_ranks, int rank, int64_t n_floats, void* multicast,
                                      void* stream) {
    if (!peer_bufs) return TVM_E_NULL;
    if (n_ranks < 1 || n_ranks > TVM_MAX_PEERS || rank < 0 || rank >= n_ranks || n_floats < 0 || (n_floats & 3))
        return TVM_E_SHAPE;
    if (n_floats == 0 || n_ranks == 1) return 0;
    AllReduceArgs a{};
    for (int p = 0; p < n_ranks; ++p) {
        if (!peer_bufs[p]) return TVM_E_NULL;
        // the sum must be bit-identical on every rank: all ranks add in the same (rank) order
        a.peer[p] = (float4*)peer_bufs[p];
    }
    a.mc = (float4*)multicast;
    a.n_ranks = n_ranks;
    const long long n4 = n_floats >> 2, per = (n4 + n_ranks - 1) / n_ranks;
    a.begin4 = per * rank;
    a.end4 = a.begin4 + per < n4 ? a.begin4 + per : n4;
    if (a.begin4 >= a.end4) return 0;
    long long ctas = (a.end4 - a.begin4 + 255) / 256;
    if (ctas > TVM_SM_COUNT * 8) ctas = TVM_SM_COUNT * 8;
    tvm_count_launch(); allreduce_slice_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(a);
    TVM_LAUNCH_CHECK();
    return 0;
}
