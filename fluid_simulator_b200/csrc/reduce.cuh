// reduce.cuh -- deterministic grid-wide reductions for the solver kernels (blocks of RED_THREADS threads).
#pragma once
#include <cuda_runtime.h>

constexpr int RED_THREADS = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// block reduction of NS sums followed by NM maxima; the last block to arrive folds all per-block partials in a fixed
// order (bitwise reproducible) and its thread 0 returns true with the grid-wide totals in out[].
template <int NS, int NM>
__device__ __forceinline__ bool grid_reduce(double* vals, double* partials, unsigned int* counter, double* out,
                                            unsigned bid = blockIdx.x, unsigned nblocks = gridDim.x) {
    constexpr int PT = RED_THREADS;
    const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    __shared__ double sh[NS + NM][PT / 32];
    __shared__ bool last;
    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int k = 0; k < NS + NM; k++) {
        const double v = k < NS ? warp_sum(vals[k]) : warp_max(vals[k]);
        if (lane == 0) sh[k][w] = v;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NS + NM; k++) {
            double v = sh[k][0];
            for (int i = 1; i < PT / 32; i++) v = k < NS ? v + sh[k][i] : fmax(v, sh[k][i]);
            partials[(size_t)k * nblocks + bid] = v;
        }
        __threadfence();
        const unsigned t = atomicInc(counter, nblocks - 1);  // wraps back to 0 => self-resetting
        last = (t == nblocks - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    double tot[NS + NM];
#pragma unroll
    for (int k = 0; k < NS + NM; k++) {
        double v = 0.0;
#pragma unroll 8
        for (unsigned i = tid; i < nblocks; i += PT) {
            const double x = __ldcg(partials + (size_t)k * nblocks + i);
            v = k < NS ? v + x : fmax(v, x);
        }
        tot[k] = k < NS ? warp_sum(v) : warp_max(v);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NS + NM; k++) sh[k][w] = tot[k];
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NS + NM; k++) {
            double v = sh[k][0];
            for (int i = 1; i < PT / 32; i++) v = k < NS ? v + sh[k][i] : fmax(v, sh[k][i]);
            out[k] = v;
        }
        return true;
    }
    return false;
}

