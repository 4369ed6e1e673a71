// tvm_warp.cuh — warp-level helpers shared by the forward and backward march kernels (device only).
#pragma once
#include "tvm_math.cuh"

#define TVM_FULL_MASK 0xffffffffu
constexpr int TVM_MAX_BLOCKS = 128;      // 32-sample blocks per ray the skip mask covers (n_samples <= 4096)

// Per-ray empty-space skip mask: bit b of word w says block 32w+b MAY contain valid samples
// (tvm_block_may_be_valid: exact/conservative, so skipping never changes ray_valid).  Lane == block.
struct TvmBlockMask {
    unsigned w[TVM_MAX_BLOCKS / 32];
    __device__ __forceinline__ bool test(int block) const {
        const int i = block >> 5;
        const unsigned word = i == 0 ? w[0] : (i == 1 ? w[1] : (i == 2 ? w[2] : w[3]));
        return (word >> (block & 31)) & 1u;
    }
    __device__ __forceinline__ unsigned word(int i) const {
        return i == 0 ? w[0] : (i == 1 ? w[1] : (i == 2 ? w[2] : w[3]));
    }
};
// all blocks of word i that exist for nblk blocks (used when every sample index must be visited)
__device__ __forceinline__ unsigned tvm_all_blocks_word(int i, int nblk) {
    const int nb = nblk - (i << 5);
    return nb >= 32 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << nb) - 1u));
}
__device__ __forceinline__ TvmBlockMask tvm_block_prepass(const tvm_field_desc& f, const TvmRay& ray, int S, int lane) {
    TvmBlockMask m;
    const int nblk = (S + 31) >> 5;
#pragma unroll
    for (int i = 0; i < TVM_MAX_BLOCKS / 32; ++i) {
        const int b = i * 32 + lane;
        const int nb = nblk - i * 32;                 // blocks in this word
        if (nb <= 0) { m.w[i] = 0u; continue; }
        // a word with one or two blocks (S = 1036/1039 -> 33 blocks): testing them costs more than simply
        // visiting them, so flag them untested (conservative, like every other "maybe")
        if (nb <= 2) { m.w[i] = (1u << nb) - 1u; continue; }
        bool maybe = false;
        if (b < nblk) maybe = tvm_block_may_be_valid(f, ray, b * 32, min(b * 32 + 31, S - 1));
        m.w[i] = __ballot_sync(TVM_FULL_MASK, maybe);
    }
    return m;
}
