// fexch.cuh -- ghost-plane exchange fused into the solver kernels of the hybrid slab projection.
//
// Inside the PCG loop every level-0 array that a neighbour reads across the slab boundary (search direction s, multigrid
// iterates xa / xb) used to travel by a kernel of its own (push_kernel, dist.cu): ~12 us per exchange on NVLink 5, of which
// ~7 us are the copy, two system-scope fences and the flag, plus a launch gap -- four of them per iteration.  Here the
// PRODUCER of the array stores the groups of its boundary planes a second time, straight into the neighbour's ghost plane
// (posted NVLink stores), and every CTA that touched a boundary plane then adds 1 to an arrival counter in the neighbour's
// memory (fence.sys + red.release.sys).  The CONSUMER's CTAs that read a ghost plane spin on that counter before their first
// load; all other CTAs never wait.  Producers walk their boundary planes first, consumers last, so the NVLink flight time
// hides behind the interior of the slab.
//
// Counting: a producer launch advances the local expectation by the number of boundary CTAs of ITS launch geometry
// (fx_exp[array][side] += nblk[side]); the neighbour runs the same kernel with the same geometry on the same schedule, so
// its CTAs add exactly that many arrivals to fx_cnt[array][side] here.  Kernels that return early because the solve is done
// do so on all ranks alike (the decision is an all-rank reduction), so both numbers skip together.
//
// Write-after-read: a ghost plane is rewritten only after the neighbour passed an all-rank reduction (or the all-rank push
// of the coarse right-hand side) that this rank entered AFTER its last read of the old contents -- the same argument as
// for push_kernel, with the stores moved from "after the producer" to "inside the producer"; the schedule is model-checked
// in tests/test_slab_protocol_model.py.
// Groups without WATER cells are not stored by the producers, locally or remotely: no consumer reads them (every neighbour
// load is masked by the stencil code).
#pragma once
#include <stdint.h>

struct FxPush {
    float* peer[2];          // the neighbour's copy of the array this kernel writes ([0] lower, [1] upper), same cell indexing; nullptr: none
    uint32_t* peer_cnt[2];   // the neighbour's arrival counter for this array and my side
    uint32_t* my_exp;        // [2] local expectations for this array
    uint32_t nblk[2];        // CTAs of this launch that signal towards each side
    int zb[2];               // my boundary planes (global index): first and last owned plane
};

struct FxWait {
    const uint32_t* cnt;     // [2] arrivals written by the neighbours
    const uint32_t* exp;     // [2] local expectations
    uint32_t* error;         // DistComm::error
    uint32_t* err_host;      // mapped host words: which wait gave up
    unsigned long long* stat;    // FSIM_WAIT_FUSED entry of the wait statistics: waited ns; `stat_n` the number of waiting CTAs
    unsigned long long* stat_n;
    unsigned long long timeout_ns;
    int has[2];              // a neighbour exists on that side and pushes into this array
    int zb[2];
};

// All-rank form (the level-1 right-hand side after the restriction: every rank owns some coarse planes, every rank needs all of
// them): the producer stores its results into every other rank's array as well and each of its CTAs adds 1 to ONE counter on
// every rank; a consumer CTA waits until the counter holds the CTAs of all other ranks' producer launches so far.
// The peer tables live in device memory (not in the parameter struct: a run-time index into a parameter array costs a
// local-memory copy of the struct per thread).
struct FxAllTable {
    float* peer[16];         // the other ranks' arrays (same indexing), n entries
    uint32_t* peer_cnt[16];  // their arrival counters
    int n;
};
struct FxAllPush {
    const FxAllTable* tab;   // nullptr: not in use
    uint32_t* my_exp;        // local expectation
    uint32_t add;            // CTAs of all OTHER ranks' launches of the same kernel (they differ with the planes a rank owns)
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long fx_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// NOTE: these structs are kernel parameters.  Never index their arrays with a run-time value: the compiler would copy the
// whole struct into local memory in every thread (measured: +13 us on a 35 us sweep).  Use the selects below.
// side of the boundary plane z, or -1
__device__ __forceinline__ int fx_side(const int zb[2], int z) { return z == zb[0] ? 0 : (z == zb[1] ? 1 : -1); }
__device__ __forceinline__ float* fx_peer(const FxPush& f, int side) { return side == 0 ? f.peer[0] : (side == 1 ? f.peer[1] : nullptr); }
__device__ __forceinline__ int fx_has(const FxWait& w, int side) { return side == 0 ? w.has[0] : (side == 1 ? w.has[1] : 0); }

// once per producer launch (one thread of the grid): both expectations advance
__device__ __forceinline__ void fx_expect(const FxPush& f) {
    if (f.peer[0]) f.my_exp[0] += f.nblk[0];
    if (f.peer[1]) f.my_exp[1] += f.nblk[1];
}

// all threads of a CTA whose stores into the neighbour's ghost plane `side` have been issued
__device__ __forceinline__ void fx_signal(const FxPush& f, int side) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        // (the release is cumulative over the CTA's stores ordered before it by the barrier: no separate fence.sys -- one
        // MEMBAR.SYS round trip less on the critical path of a one-wave kernel)
        uint32_t* cnt = side == 0 ? f.peer_cnt[0] : f.peer_cnt[1];
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(1u) : "memory");
    }
}

// all threads of a producer CTA of the all-rank form, after their stores (fx_all_store)
__device__ __forceinline__ void fx_all_signal(const FxAllPush& f) {
    if (!f.tab) return;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) *f.my_exp += f.add;
    __syncthreads();
    const int t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    if (t < f.tab->n) {  // lane r signals rank r: its release orders the CTA's stores (barrier above) before the count
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(f.tab->peer_cnt[t]), "r"(1u) : "memory");
    }
}
__device__ __forceinline__ void fx_all_store2(const FxAllPush& f, int64_t idx, float2 v) {
    if (!f.tab) return;
    const int n = f.tab->n;
    for (int r = 0; r < n; r++) *reinterpret_cast<float2*>(f.tab->peer[r] + idx) = v;
}

// all threads of a CTA that is about to read ghost plane `side`
__device__ __forceinline__ void fx_wait(const FxWait& w, int side) {
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0 && *(volatile uint32_t*)w.error != 1u) {
        const uint32_t want = *(volatile const uint32_t*)&w.exp[side];
        const unsigned long long t0 = fx_now_ns();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w.cnt + side) : "memory");
            if ((int32_t)(v - want) >= 0) break;
            if (fx_now_ns() - t0 > w.timeout_ns) {
                if (atomicCAS(w.error, 0u, 1u) == 0u) {
                    volatile uint32_t* eh = w.err_host;
                    eh[1] = 0x90u + side; eh[2] = want; eh[3] = v;
                    eh[0] = 1u;
                }
                *(volatile uint32_t*)w.error = 1u;
                __threadfence_system();
                break;
            }
            __nanosleep(32);
        }
        atomicAdd(w.stat, fx_now_ns() - t0);
        atomicAdd(w.stat_n, 1ull);
    }
    __syncthreads();
}
#endif
