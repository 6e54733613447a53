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

// Deterministic grid-wide reduction of NS sums followed by NM maxima.  Two-level arrival counters keep the same-address
// atomic traffic low for grids of tens of thousands of blocks: blocks arrive in groups of RED_GROUP; the last block of a
// group folds the group's per-block partials (fixed order) into one group partial and arrives at the top counter; the
// last group folds the group partials (fixed order).  Exactly one thread of the grid returns true with the totals in out[].
// partials: [(NS+NM)][nblocks + ngroups] doubles; counters: [1 + ngroups] unsigned, zero-initialised, self-resetting.
constexpr unsigned RED_GROUP = 32;

// grid_reduce_w: the same, returning 2 in that thread, 1 in the other lanes of its warp (warp 0 of the last CTA; a
// warp-collective epilogue such as the fused all-rank reduction of the slab mode follows) and 0 everywhere else.
template <int NS, int NM>
__device__ __forceinline__ int grid_reduce_w(double* vals, double* partials, unsigned int* counters, double* out,
                                             unsigned bid = blockIdx.x, unsigned nblocks = gridDim.x) {
    constexpr int PT = RED_THREADS;
    constexpr int NV = NS + NM;
    __shared__ double sh[NV][PT / 32];
    __shared__ int role;  // 0: done, 1: last of its group, 2: last of the grid
    const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int lane = tid & 31, w = tid >> 5;
    const unsigned ngroups = (nblocks + RED_GROUP - 1) / RED_GROUP;
    const unsigned stride = nblocks + ngroups;  // per value: block partials followed by group partials
    const unsigned grp = bid / RED_GROUP;
    const unsigned gsize = min(RED_GROUP, nblocks - grp * RED_GROUP);
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const double v = k < NS ? warp_sum(vals[k]) : warp_max(vals[k]);
        if (lane == 0) sh[k][w] = v;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double v = sh[k][0];
            for (int i = 1; i < PT / 32; i++) v = k < NS ? v + sh[k][i] : fmax(v, sh[k][i]);
            partials[(size_t)k * stride + bid] = v;
        }
        __threadfence();
        role = (atomicInc(counters + 1 + grp, gsize - 1) == gsize - 1) ? 1 : 0;  // wraps to 0 => self-resetting
    }
    __syncthreads();
    if (role == 0) return 0;
    __threadfence();
    // last block of the group: fold the group's block partials in block order (one warp is plenty for <= 32 values)
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double v = ((unsigned)lane < gsize) ? __ldcg(partials + (size_t)k * stride + grp * RED_GROUP + lane) : 0.0;
            // fixed-order serial fold by lane 0 through shuffles: tree order is fixed => reproducible
            v = k < NS ? warp_sum(v) : warp_max(v);
            if (lane == 0) partials[(size_t)k * stride + nblocks + grp] = v;
        }
        if (lane == 0) {
            __threadfence();
            role = (atomicInc(counters, ngroups - 1) == ngroups - 1) ? 2 : 0;
        }
    }
    __syncthreads();
    if (role != 2) return 0;
    __threadfence();
    double tot[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double v = 0.0;
#pragma unroll 4
        for (unsigned i = tid; i < ngroups; i += PT) {
            const double x = __ldcg(partials + (size_t)k * stride + nblocks + i);
            v = k < NS ? v + x : fmax(v, x);
        }
        tot[k] = k < NS ? warp_sum(v) : warp_max(v);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) sh[k][w] = tot[k];
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double v = sh[k][0];
            for (int i = 1; i < PT / 32; i++) v = k < NS ? v + sh[k][i] : fmax(v, sh[k][i]);
            out[k] = v;
        }
        return 2;
    }
    return w == 0 ? 1 : 0;
}
template <int NS, int NM>
__device__ __forceinline__ bool grid_reduce(double* vals, double* partials, unsigned int* counters, double* out,
                                            unsigned bid = blockIdx.x, unsigned nblocks = gridDim.x) {
    return grid_reduce_w<NS, NM>(vals, partials, counters, out, bid, nblocks) == 2;
}
