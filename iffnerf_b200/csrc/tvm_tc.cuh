// tvm_tc.cuh — tcgen05 / TMEM / mbarrier PTX wrappers and the operand layout shared by the tensor-core shade
// kernels (shade_tc.cu: bf16 mode; shade_tc3.cu: bf16x3 split mode).  Device-only.
#pragma once
#include <cuda_bf16.h>
#include "tvm_common.cuh"

namespace tvmtc {

constexpr int FC = TVM_FEATURE_C;
constexpr int TC_RAYS = 128;
constexpr int TC_THREADS = 512;     // 16 warps: 4 per TMEM lane quarter, each a quarter of the columns
constexpr int N0 = 32;                 // basis rows padded (app_dim <= 32)
constexpr int N3 = 16;                 // rgb rows padded
constexpr int COL0 = 0, COL1 = 32, COL2 = 160, COL3 = 288;   // TMEM column bases of the four accumulators
constexpr int TMEM_COLS = 512;

struct TcDims {
    int ta, k0;        // sum(n_app), padded to 16
    int in_c, k1;      // MLP input width, padded to 16
    int app_dim, fea_pe, view_pe;
    // byte offsets inside the weight image / shared memory
    int img0, img1, img2, img3, img_bytes;
};
__host__ __device__ inline TcDims tc_dims(const tvm_field_desc* d) {
    TcDims t;
    t.ta = d->n_app[0] + d->n_app[1] + d->n_app[2];
    t.k0 = (t.ta + 15) / 16 * 16;
    t.in_c = 2 * d->view_pe * 3 + 2 * d->fea_pe * d->app_dim + 3 + d->app_dim;
    t.k1 = (t.in_c + 15) / 16 * 16;
    t.app_dim = d->app_dim; t.fea_pe = d->fea_pe; t.view_pe = d->view_pe;
    t.img0 = 0;
    t.img1 = t.img0 + N0 * t.k0 * 2;
    t.img2 = t.img1 + FC * t.k1 * 2;
    t.img3 = t.img2 + FC * FC * 2;
    t.img_bytes = t.img3 + N3 * FC * 2;
    return t;
}

// byte offset of element (row, k) of a K-major bf16 operand with K columns in the no-swizzle canonical layout:
// 8x8 core matrices of 128 contiguous bytes (row stride 16 B); K-adjacent cores 128 B apart (LBO), 8-row groups
// K*16 B apart (SBO).  (cute::UMMA canonical INTERLEAVE layout ((8,n),(8,2)):((8,SBO),(1,LBO)) in elements.)
__host__ __device__ inline int canon_off(int row, int k, int K) {
    return (row >> 3) * (K * 16) + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

// fp32 [R][Cc] (row-major, torch Linear weight) -> bf16 canonical image with R_pad rows and K_pad columns;
// part 0 = bf16(v), part 1 = bf16(v - bf16(v)) (low term of the 2-term split used by the bf16x3 mode)
static __global__ void pack_bf16_operand_kernel(const float* __restrict__ src, int R, int Cc, int R_pad, int K_pad,
                                                unsigned char* __restrict__ dst, int part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R_pad * K_pad) return;
    const int r = i / K_pad, k = i - r * K_pad;
    const float v = (r < R && k < Cc) ? __ldg(src + r * Cc + k) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(dst + canon_off(r, k, K_pad)) =
        part == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---- raw PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int K) {
    // start address [0,14) (>>4), LBO [16,30) = 128 B, SBO [32,46) = K*16 B, version [46,48) = 1, layout [61,64) = 0
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((K * 16) >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    // c_format F32 [4,6)=1, a_format BF16 [7,10)=1, b_format BF16 [10,13)=1, K-major A/B, n_dim [17,23)=N>>3, m_dim [24,29)=M>>4
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}


}  // namespace tvmtc
