/*
 * tvm_bench.h — measurement aids (libtvm_bench.so).  NOT part of the product library: bench.py and scripts/ load it to
 * measure the gather ceilings the march kernels are reported against.
 */
#ifndef TVM_BENCH_H
#define TVM_BENCH_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Quads of lanes gather random 64-B pieces (LDG.128 per lane, the march kernels' access shape) from
 * `granule_bytes`-sized granules (64 = density texel, 192 = appearance texel) inside `bytes` of `buf`: the factor set
 * size gives the L2 -> SM gather ceiling, a few KB the L1-resident one.  The caller times the launch with CUDA events;
 * *bytes_moved = bytes requested by the lanes. */
int tvm_gather_microbench(const void* buf, size_t bytes, int granule_bytes, int iters, float* sink,
                          unsigned long long* bytes_moved, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TVM_BENCH_H */
