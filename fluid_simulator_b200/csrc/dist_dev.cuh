// dist_dev.cuh -- device-side definitions of the z-slab exchange layer shared by dist.cu and the solver kernels that carry a
// fused all-rank reduction (pcg.cu, mg.cu): the per-rank communication block peers write into, the flag primitives and the
// bounded wait, and the warp-level all-rank reduction.
#pragma once
#include <stdint.h>

#include "fsim_internal.h"
#include "pcg_finish.cuh"

#define DIST_MAX_RANKS 16

struct DistSlot { double v[4]; uint32_t epoch; uint32_t pad[7]; };

struct DistComm {
    // written by the z-neighbours ([0]: by the lower one, [1]: by the upper one)
    uint32_t arrive[2], done[2];
    uint32_t mig_epoch[2], mig_count[2];
    // written by every rank: slot[parity][source rank]
    DistSlot slot[2][DIST_MAX_RANKS];
    uint32_t arrive_all[DIST_MAX_RANKS], done_all[DIST_MAX_RANKS];  // all-rank handshake of the solver-input gather
    uint32_t push_flag[2];                 // "your ghost plane holds my boundary plane of exchange #epoch" (push halos of the PCG loop)
    uint32_t gpush_flag[DIST_MAX_RANKS];   // the same for the all-rank push of the coarse right-hand side
    uint32_t fx_cnt[3][2];                 // fused exchanges (fexch.cuh): arrivals [array: s, xa, xb][from the lower / upper neighbour]
    uint32_t gx_cnt;                       // fused all-rank push of the level-1 right-hand side: arrivals from all other ranks
    // local
    uint32_t fx_exp[3][2];                 // ... and how many this rank expects by now
    uint32_t gx_exp;
    uint32_t halo_epoch, ar_epoch, mig_ep, blocks_done, mig_blocks_done, gather_epoch, gather_blocks_done, push_epoch, push_blocks_done, gpush_epoch, gpush_blocks_done, pad1;
    uint32_t error;   // 1 wait timed out, 2 emigrant list full, 3 immigrant outside the owned planes
    uint32_t n_src;   // particles the next reorder reads: locals + immigrants
    uint32_t pad0;
    unsigned long long timeout_ns;
    // time spent spinning in wait_ge and the number of waits, by class (FSIM_WAIT_*, fsim.h): the share of an exchange that is
    // waiting for the peer (its skew + flight time of the flag) rather than this rank's own launch / copy
    unsigned long long wait_ns[FSIM_WAIT_CLASSES], waits[FSIM_WAIT_CLASSES], kern_ns[FSIM_WAIT_CLASSES];
};


__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_peer_u4(const void* p) {  // never served from a stale L1 line
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag >= ep (epochs only grow); bounded: on time-out the error word is set and every later wait returns at once
// (err_host[1..3] record which wait gave up: 1 halo arrive, 2 halo done, 3 all-rank reduction, 4 migration; the epoch; the flag)
__device__ __forceinline__ bool wait_ge(const uint32_t* flag, uint32_t ep, DistComm* c, uint32_t* err_host, uint32_t where) {
    if (*(volatile uint32_t*)&c->error == 1u) return false;
    const unsigned long long t0 = now_ns();
    while ((int32_t)(ld_acquire_sys(flag) - ep) < 0) {
        if (now_ns() - t0 > c->timeout_ns) {
            if (atomicCAS(&c->error, 0u, 1u) == 0u) {
                volatile uint32_t* eh = err_host;
                eh[1] = where; eh[2] = ep; eh[3] = ld_acquire_sys(flag);
                eh[0] = 1u;
            }
            *(volatile uint32_t*)&c->error = 1u;
            __threadfence_system();
            return false;
        }
        __nanosleep(64);
    }
    return true;
}

// elapsed waiting time of ONE designated thread per kernel (block 0 / the last block; waits of the other blocks and of the
// other lanes run concurrently with it), so that the sums are time on this rank's stream
__device__ __forceinline__ void wait_account(DistComm* c, int cls, unsigned long long t0) {
    atomicAdd(&c->wait_ns[cls], now_ns() - t0);
    atomicAdd(&c->waits[cls], 1ull);
}


// ---- all-rank reduction inside the reducing kernel --------------------------------------------------------------------------
// The three reductions of a PCG iteration (s.As, max |r|, z.r) used to be a 32-thread kernel each behind the kernel that
// produced this rank's partial result (allreduce_kernel, dist.cu).  ar_warp is the same protocol executed by warp 0 of the LAST
// CTA of the reducing kernel itself, right after the grid-wide reduction: lane r stores this rank's four doubles + the epoch
// into its slot on rank r (release), waits for rank r's slot here (acquire); lane 0 sums the slots in rank order -- the
// order, and so the bits, of allreduce_kernel -- and takes the PCG decision (pcg_finish.cuh).  One kernel hop less per
// reduction.  The table lives in device memory (a run-time index into a kernel-parameter array would cost a local copy).
struct ArDev {
    DistComm* comm;
    DistComm* all[DIST_MAX_RANKS];
    uint32_t* err_host;
    int rank, nranks;
};

// all 32 lanes of one warp; `mine` (4 values: sum, sum, max, sum) is read from lane 0
__device__ __forceinline__ void ar_warp(const ArDev* t, PcgScalars* sc, PcgHostStatus* status, int kind, const double mine[4]) {
    __syncwarp();
    const int lane = threadIdx.x & 31;
    double m[4];
#pragma unroll
    for (int k = 0; k < 4; k++) m[k] = __shfl_sync(0xffffffffu, mine[k], 0);
    DistComm* c = t->comm;
    const int nranks = t->nranks;
    const uint32_t ep = *(volatile uint32_t*)&c->ar_epoch + 1u;
    const int par = ep & 1u;
    __shared__ double ar_sv[DIST_MAX_RANKS][4];
    const unsigned long long t0 = now_ns();
    if (lane < nranks) {
        DistSlot* s = &t->all[lane]->slot[par][t->rank];
        volatile double* v = s->v;
        v[0] = m[0]; v[1] = m[1]; v[2] = m[2]; v[3] = m[3];
        st_release_sys(&s->epoch, ep);
        DistSlot* in = &c->slot[par][lane];
        wait_ge(&in->epoch, ep, c, t->err_host, 0x300u + (kind << 4) + lane);
        volatile double* w = in->v;
        ar_sv[lane][0] = w[0]; ar_sv[lane][1] = w[1]; ar_sv[lane][2] = w[2]; ar_sv[lane][3] = w[3];
    }
    __syncwarp();
    if (lane == 0) {
        wait_account(c, FSIM_WAIT_ALLREDUCE, t0);
        double s0 = 0.0, s1 = 0.0, mx = 0.0, s3 = 0.0;
        for (int r = 0; r < nranks; r++) { s0 += ar_sv[r][0]; s1 += ar_sv[r][1]; mx = fmax(mx, ar_sv[r][2]); s3 += ar_sv[r][3]; }
        switch (kind) {
            case AR_SPMV: sc->sq = s0; break;
            case AR_UPDATE: pcg_finish_update(sc, status, 0.0, mx); break;
            case AR_DOTZR: sc->sigma_new = s0; break;
            default: break;
        }
        (void)s1; (void)s3;
        sc->loc[0] = sc->loc[1] = sc->loc[2] = sc->loc[3] = 0.0;
        c->ar_epoch = ep;
        __threadfence();
        atomicAdd(&c->kern_ns[FSIM_WAIT_ALLREDUCE], now_ns() - t0);
    }
}
