// Deterministic block / grid reductions shared by the vector kernels.
#pragma once
#include <stdint.h>

namespace hdg {

constexpr int RB = 256;   // threads per block of the vector kernels

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double ws[RB / 32];
    __shared__ double total;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();   // protect ws/total reuse across calls
    if (lane == 0) ws[w] = v;
    __syncthreads();
    if (w == 0) {
        double s = lane < RB / 32 ? ws[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) total = s;
    }
    __syncthreads();
    return total;
}

// fixed-order reduction of `np` partials by the whole block; result broadcast to all threads
__device__ __forceinline__ double reduce_partials(const double* __restrict__ part, int np) {
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += RB) s += part[i];
    return block_sum(s);
}


static __global__ void final_sum(const double* __restrict__ part, int np, double* __restrict__ out) {
    double s = reduce_partials(part, np);
    if (threadIdx.x == 0) *out = s;
}

}  // namespace hdg
