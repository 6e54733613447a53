// dist.cu -- z-slab decomposition of the hot path over NVLink peer memory.
//
// One handle (= one process, one GPU) holds the global planes [zoff, zoff + gz) of the grid: the planes it owns plus one
// ghost plane towards each z-neighbour.  Positions stay global, so every particle-side expression is the one the
// single-GPU path evaluates.  All inter-GPU traffic is hand-written loads / stores on peer memory (cudaIpc mappings of
// the neighbours' arrays, or plain peer pointers when the handles live in one process):
//   * halo_kernel       two-sided handshake + pull of the neighbours' boundary planes into the local ghost planes
//                       (epochs live in device memory, so the kernel replays unchanged inside the PCG loop graph);
//   * allreduce_kernel  every rank pushes its partial reduction into a slot of every rank, then sums the slots in rank
//                       order: deterministic, bit-identical on all ranks, ~one NVLink round trip; it then takes the PCG
//                       decision (pcg_finish.cuh) that the last block takes in the single-GPU path;
//   * mig_send / mig_recv  particles whose cell left the owned planes are pushed into the neighbour's receive buffer and
//                       binned there before the sort;
//   * push_kernel       the in-loop exchanges of the hybrid projection (search direction, multigrid iterate, coarse
//                       right-hand side): posted stores into the peers' arrays + a flag, no handshake (see the kernel);
//   * gather_kernel     all-rank handshake + pull of the planes every rank owns (solver inputs, coarse right-hand side).
// The projection of a slab handle runs on a full-grid solver context (fsim::solver) whose arrays sit at global plane
// indices on every rank (DESIGN.md §7 "hybrid"): halos there need no index translation.
// Every wait is bounded (FSIM_DIST_TIMEOUT_MS, default 20 s): a lost peer becomes FSIM_ERR_COMM, never a hung GPU.
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fsim_internal.h"
#include "dist_dev.cuh"
#include "fexch.cuh"
#include "launch.cuh"
#include "pcg_finish.cuh"

#define DIST_NARR 24
#define DIST_NCH 15  // particle channels of the widest (APIC) layout

enum { ARR_U = 0, ARR_U2 = 3, ARR_W = 6, ARR_DENS = 9, ARR_CNT = 10, ARR_FLAGS = 11, ARR_P = 12, ARR_S = 13, ARR_COMM = 14, ARR_RECV = 15,
       // arrays of the full-grid solver context (hybrid projection): same global plane index on every rank
       ARR_HS_S = 16, ARR_HS_P = 17, ARR_HS_XA = 18, ARR_HS_XB = 19, ARR_HS_B1 = 20,
       ARR_PA = 21 };  // push-apart: positions of the particles in my two boundary planes, staged for the neighbours

static bool fused_enabled();

struct DistState {
    int rank, nranks;
    int own_lo, own_hi;  // global planes owned
    DistComm* comm;
    uint32_t* err_host;  // pinned + mapped
    uint32_t* err_dev;
    MigDev* mig;
    uint32_t* mig_idx[2];
    int64_t mig_cap;
    float* recv;         // [2 sides][DIST_NCH + 1 (ids)][mig_cap]
    float* stage;        // P2G staging: [2 sides][2 (ghost, boundary)][7 channels][plane]
    void* local_arr[DIST_NARR];       // allocation bases
    size_t local_off[DIST_NARR];      // byte offset of the array inside its allocation
    void* peer_arr[2][DIST_NARR];
    DistComm* peer_comm[2];
    int peer_ghost[2], peer_bnd[2];  // the neighbour's ghost plane towards us / its boundary plane (its local indices)
    DistComm* all_comm[DIST_MAX_RANKS];
    void* all_arr[DIST_MAX_RANKS][DIST_NARR];     // every rank's arrays (the gather pulls the planes each rank owns)
    int all_zown0[DIST_MAX_RANKS], all_zown1[DIST_MAX_RANKS], all_zoff[DIST_MAX_RANKS];
    bool connected, nsrc_valid;
    bool peer_in_process;  // some other rank of the group lives in this process (they would share libc's rand() stream)
    FxAllTable* gx_table;  // device copy of the peer table of the fused all-rank push (built on first use)
    ArDev* ar_table;       // device copy of the rank table of the fused all-rank reduction (dist_dev.cuh ar_warp)
    // push-apart across slabs (pushParticlesApart, hashedParticles.cpp:64-107): per side one block of words
    //   [0] particle count | [1 .. sz+1] per-cell starts of the plane (relative) | x[pa_capa] | y[pa_capa] | z[pa_capa]
    // pa_stage: my boundary planes (read by the neighbours), pa_ghost: the neighbours' boundary planes (pulled)
    uint32_t *pa_stage, *pa_ghost;
    int64_t pa_capa, pa_head, pa_words;  // particles per plane (multiple of 4), header words (multiple of 4), words per side
    std::vector<void*> ipc_opened;
};

namespace {

// mask >= 0: a plane of floats of which only the cells with (x + y) & 1 == mask are copied (row length gx): the red-black sweeps of
// the BasicMacGrid solver on slabs, where each rank owns the updates of one colour of a shared face plane
// limit != nullptr: copy only the first round_up(*limit, 4) 4-byte elements (a count that lives in the neighbour's memory)
struct HaloCopy { void* dst; const void* src; uint32_t bytes; int side; int mask, gx; const uint32_t* limit; };
struct HaloArgs {
    DistComm* comm;
    DistComm* peer[2];
    uint32_t* err_host;
    const PcgScalars* sc;  // non-null inside the PCG loop: skip once the solve is done (the flag is the same on all ranks)
    int ncopy;
    HaloCopy cp[32];
};

__global__ void __launch_bounds__(256) halo_kernel(const __grid_constant__ HaloArgs a) {
    if (a.sc && a.sc->done) return;
    DistComm* c = a.comm;
    __shared__ uint32_t ep_s;
    if (threadIdx.x == 0) {
        const uint32_t ep = *(volatile uint32_t*)&c->halo_epoch + 1u;  // advanced by the last block on its way out
        ep_s = ep;
        if (blockIdx.x == 0) {  // everything this rank launched before is complete: its planes are final, its ghosts free
            __threadfence_system();
            for (int side = 0; side < 2; side++)
                if (a.peer[side]) st_release_sys(&a.peer[side]->arrive[1 - side], ep);
        }
        const unsigned long long t0 = now_ns();
        for (int side = 0; side < 2; side++)
            if (a.peer[side]) wait_ge(&c->arrive[side], ep, c, a.err_host, 0x10u + side);
        if (blockIdx.x == 0) wait_account(c, FSIM_WAIT_HALO, t0);
    }
    __syncthreads();
    const uint32_t ep = ep_s;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < a.ncopy; k++) {
        const HaloCopy& cp = a.cp[k];
        if (!a.peer[cp.side]) continue;
        if (cp.mask >= 0) {
            const size_t n = cp.bytes >> 2;
            for (size_t i = gtid; i < n; i += gsz) {
                const int x = (int)(i % cp.gx), y = (int)(i / cp.gx);
                if (((x + y) & 1) == cp.mask) reinterpret_cast<float*>(cp.dst)[i] = reinterpret_cast<const volatile float*>(cp.src)[i];
            }
        } else if ((((size_t)cp.dst | (size_t)cp.src | (size_t)cp.bytes) & 15) == 0) {
            size_t n = cp.bytes >> 4;
            if (cp.limit) n = min(n, (size_t)((*(const volatile uint32_t*)cp.limit + 3u) >> 2));
            for (size_t i = gtid; i < n; i += gsz) reinterpret_cast<uint4*>(cp.dst)[i] = ld_peer_u4(reinterpret_cast<const uint4*>(cp.src) + i);
        } else {
            for (size_t i = gtid; i < cp.bytes; i += gsz) reinterpret_cast<uint8_t*>(cp.dst)[i] = reinterpret_cast<const volatile uint8_t*>(cp.src)[i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t old = atomicAdd(&c->blocks_done, 1u);
        if (old == gridDim.x - 1) {  // last block: tell the neighbours we are done reading their planes, wait for the same
            c->blocks_done = 0;
            __threadfence_system();
            for (int side = 0; side < 2; side++)
                if (a.peer[side]) st_release_sys(&a.peer[side]->done[1 - side], ep);
            const unsigned long long t0 = now_ns();
            for (int side = 0; side < 2; side++)
                if (a.peer[side]) wait_ge(&c->done[side], ep, c, a.err_host, 0x20u + side);
            wait_account(c, FSIM_WAIT_HALO, t0);
            c->halo_epoch = ep;
            __threadfence();
        }
    }
}

// Push form of the exchange, for the halos INSIDE the PCG loop (search direction, multigrid iterate, coarse right-hand side):
// every rank stores its planes straight into the neighbours' arrays (posted NVLink writes, no read round trip) and raises a
// flag; the kernel ends when the neighbours' flags have arrived.  No handshake guards the destination: between two uses of a
// ghost plane the loop always passes an all-rank reduction or gather, so its previous reader has finished (DESIGN.md §7;
// the schedule is model-checked in tests/test_slab_protocol_model.py).
struct PushCopy { void* dst; const void* src; size_t bytes; int to; };
struct PushArgs {
    DistComm* comm;
    DistComm* peer[DIST_MAX_RANKS];  // halo: [0] lower, [1] upper neighbour; gather: every rank
    uint32_t* err_host;
    const PcgScalars* sc;
    int npeer, self, all_ranks, ncopy;
    PushCopy cp[DIST_MAX_RANKS];
};

__global__ void __launch_bounds__(256) push_kernel(const __grid_constant__ PushArgs a) {
    // launched as a programmatic dependent (launch.cuh): resident while the producer of the planes drains.  It never triggers
    // ITS dependents early: a grid parked behind a kernel that spins on another rank's flag could starve that rank of CTA
    // slots when several ranks share one device
    pdl_wait();
    if (a.sc && a.sc->done) return;
    const unsigned long long t_in = now_ns();
    DistComm* c = a.comm;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < a.ncopy; k++) {
        const PushCopy& cp = a.cp[k];
        const size_t n = cp.bytes >> 4;  // planes are 16-byte multiples on this path (checked by the host)
        const uint4* src = reinterpret_cast<const uint4*>(cp.src);
        uint4* dst = reinterpret_cast<uint4*>(cp.dst);
        for (size_t i = gtid; i < n; i += gsz) dst[i] = src[i];
    }
    __syncthreads();
    __shared__ int last_s;
    uint32_t* my_epoch = a.all_ranks ? &c->gpush_epoch : &c->push_epoch;
    uint32_t* my_count = a.all_ranks ? &c->gpush_blocks_done : &c->push_blocks_done;
    if (threadIdx.x == 0) {
        __threadfence_system();
        last_s = atomicAdd(my_count, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last_s) return;
    const uint32_t ep = *(volatile uint32_t*)my_epoch + 1u;
    const unsigned long long t0 = now_ns();
    if (threadIdx.x < a.npeer && a.peer[threadIdx.x] && !(a.all_ranks && (int)threadIdx.x == a.self)) {
        const int r = threadIdx.x;
        __threadfence_system();
        if (a.all_ranks) {
            st_release_sys(&a.peer[r]->gpush_flag[a.self], ep);
            wait_ge(&c->gpush_flag[r], ep, c, a.err_host, 0x70u + r);
        } else {  // I am the upper neighbour of my lower neighbour: its flag [1], and the other way round
            st_release_sys(&a.peer[r]->push_flag[1 - r], ep);
            wait_ge(&c->push_flag[r], ep, c, a.err_host, 0x80u + r);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        wait_account(c, a.all_ranks ? FSIM_WAIT_GPUSH : FSIM_WAIT_PUSH, t0);
        atomicAdd(&c->kern_ns[a.all_ranks ? FSIM_WAIT_GPUSH : FSIM_WAIT_PUSH], now_ns() - t_in);  // entry of the last block -> exit
    }
    if (threadIdx.x == 0) { *my_count = 0; *my_epoch = ep; __threadfence(); }
}

// P2G ghost-plane sum: boundary plane += what the neighbour scattered into its ghost plane; ghost plane += the neighbour's
// own boundary plane (fp32 addition commutes, so both ranks hold bit-identical totals)
struct HaloAddArgs {
    float* dst[2][2][7];        // [side][0: my boundary plane, 1: my ghost plane][channel]
    const float* stg[2][2][7];  // [side][0: neighbour's ghost plane, 1: neighbour's boundary plane][channel]
    int has[2];
    int plane;
};
__global__ void __launch_bounds__(256) halo_add_kernel(const __grid_constant__ HaloAddArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.plane) return;
    const int side = blockIdx.y >> 1, which = blockIdx.y & 1;
    if (!a.has[side]) return;
#pragma unroll
    for (int ch = 0; ch < 7; ch++) a.dst[side][which][ch][i] += a.stg[side][which][ch][i];
}

struct ArArgs {
    DistComm* comm;
    DistComm* all[DIST_MAX_RANKS];
    uint32_t* err_host;
    PcgScalars* sc;
    PcgHostStatus* status;
    int rank, nranks, kind, check_done;
};

__global__ void __launch_bounds__(32) allreduce_kernel(const __grid_constant__ ArArgs a) {
    pdl_wait();  // (see push_kernel)
    if (a.check_done && a.sc->done) return;
    const unsigned long long t_in = now_ns();
    DistComm* c = a.comm;
    const int lane = threadIdx.x;
    const uint32_t ep = *(volatile uint32_t*)&c->ar_epoch + 1u;
    const int par = ep & 1u;
    __shared__ double sv[DIST_MAX_RANKS][4];
    const unsigned long long t0 = now_ns();
    if (lane < a.nranks) {
        DistSlot* s = &a.all[lane]->slot[par][a.rank];
        volatile double* v = s->v;
        v[0] = a.sc->loc[0]; v[1] = a.sc->loc[1]; v[2] = a.sc->loc[2]; v[3] = a.sc->loc[3];
        st_release_sys(&s->epoch, ep);  // (orders this lane's four stores before the epoch: no separate fence.sys)
        DistSlot* m = &c->slot[par][lane];
        wait_ge(&m->epoch, ep, c, a.err_host, 0x300u + (a.kind << 4) + lane);
        volatile double* w = m->v;
        sv[lane][0] = w[0]; sv[lane][1] = w[1]; sv[lane][2] = w[2]; sv[lane][3] = w[3];
    }
    __syncwarp();
    if (lane == 0) {
        wait_account(c, FSIM_WAIT_ALLREDUCE, t0);  // (slot stores to every rank + the wait for theirs)
        double s0 = 0.0, s1 = 0.0, mx = 0.0, s3 = 0.0;
        for (int r = 0; r < a.nranks; r++) { s0 += sv[r][0]; s1 += sv[r][1]; mx = fmax(mx, sv[r][2]); s3 += sv[r][3]; }
        PcgScalars* sc = a.sc;
        switch (a.kind) {
            case AR_RHS: pcg_finish_rhs(sc, a.status, s0, s3); break;
            case AR_RESIDUAL: pcg_finish_residual(sc, a.status, mx); break;
            case AR_START: sc->sigma = s0; break;
            case AR_SPMV: sc->sq = s0; break;
            case AR_UPDATE: pcg_finish_update(sc, a.status, 0.0, mx); break;
            case AR_UPDATE_JACOBI: pcg_finish_update(sc, a.status, s0, mx); break;
            case AR_DOTZR: sc->sigma_new = s0; break;
        }
        sc->loc[0] = sc->loc[1] = sc->loc[2] = sc->loc[3] = 0.0;
        c->ar_epoch = ep;
        __threadfence();
        atomicAdd(&c->kern_ns[FSIM_WAIT_ALLREDUCE], now_ns() - t_in);
    }
}

struct MigSendArgs {
    DistComm* comm;
    DistComm* peer[2];
    MigDev* mig;
    const float* src[DIST_NCH];
    const uint32_t* src_id;
    int nch;
    float* peer_recv[2];  // the receive area of neighbour `dir` that belongs to us (its side 1 - dir)
    int64_t cap;
};

__global__ void __launch_bounds__(256) mig_send_kernel(const __grid_constant__ MigSendArgs a) {
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    for (int dir = 0; dir < 2; dir++) {
        if (!a.peer[dir]) continue;
        const uint32_t n = min(a.mig->count[dir], a.mig->cap);
        float* out = a.peer_recv[dir];
        for (size_t j = gtid; j < n; j += gsz) {
            const uint32_t i = a.mig->idx[dir][j];
            for (int ch = 0; ch < a.nch; ch++) out[(size_t)ch * a.cap + j] = a.src[ch][i];
            reinterpret_cast<uint32_t*>(out)[(size_t)DIST_NCH * a.cap + j] = a.src_id ? a.src_id[i] : 0u;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        DistComm* c = a.comm;
        const uint32_t old = atomicAdd(&c->mig_blocks_done, 1u);
        if (old == gridDim.x - 1) {
            c->mig_blocks_done = 0;
            const uint32_t ep = c->mig_ep + 1u;
            for (int dir = 0; dir < 2; dir++) {
                if (!a.peer[dir]) continue;
                *(volatile uint32_t*)&a.peer[dir]->mig_count[1 - dir] = min(a.mig->count[dir], a.mig->cap);
                __threadfence_system();
                st_release_sys(&a.peer[dir]->mig_epoch[1 - dir], ep);
            }
            if (a.mig->overflow) { *(volatile uint32_t*)&c->error = 2u; }
            c->mig_ep = ep;
            __threadfence();
        }
    }
}

struct MigRecvArgs {
    DistComm* comm;
    int has[2];
    uint32_t* err_host;
    MigDev* mig;
    const float* recv;  // [2][DIST_NCH + 1][cap]
    int64_t cap;
    float* dst[DIST_NCH];
    uint32_t* dst_id;
    int nch;
    int64_t np;  // immigrants are appended behind the locals
    GridDims g;
    uint32_t *cnt, *key, *rank;
};

__global__ void __launch_bounds__(256) mig_recv_kernel(const __grid_constant__ MigRecvArgs a) {
    DistComm* c = a.comm;
    if (threadIdx.x == 0) {
        const uint32_t ep = c->mig_ep;  // advanced by this step's mig_send_kernel
        const unsigned long long t0 = now_ns();
        for (int side = 0; side < 2; side++)
            if (a.has[side]) wait_ge(&c->mig_epoch[side], ep, c, a.err_host, 0x40u + side);
        if (blockIdx.x == 0) wait_account(c, FSIM_WAIT_MIGRATE, t0);
    }
    __syncthreads();
    const uint32_t n0 = a.has[0] ? *(volatile uint32_t*)&c->mig_count[0] : 0u, n1 = a.has[1] ? *(volatile uint32_t*)&c->mig_count[1] : 0u;
    const uint32_t total = n0 + n1;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    const GridDims& g = a.g;
    for (size_t j = gtid; j < total; j += gsz) {
        const int side = j < n0 ? 0 : 1;
        const size_t jj = side ? j - n0 : j;
        const float* in = a.recv + (size_t)side * (DIST_NCH + 1) * a.cap;
        const size_t d = (size_t)a.np + j;
        float x = 0.f, y = 0.f, z = 0.f;
        for (int ch = 0; ch < a.nch; ch++) {
            const float v = in[(size_t)ch * a.cap + jj];
            a.dst[ch][d] = v;
            if (ch == 0) x = v; else if (ch == 1) y = v; else if (ch == 2) z = v;
        }
        if (a.dst_id) a.dst_id[d] = reinterpret_cast<const uint32_t*>(in)[(size_t)DIST_NCH * a.cap + jj];
        // same key expression as the advect kernels (particles.cu cell_key)
        int ix = (int)((double)x * g.dihx), iy = (int)((double)y * g.dihy), iz = (int)((double)z * g.dihz) - g.zoff;
        ix = min(max(ix, 0), g.gx - 1); iy = min(max(iy, 0), g.gy - 1);
        if (iz < g.zown0 || iz >= g.zown1) {  // crossed a whole slab in one step: not supported, flagged; the particle is dropped
            *(volatile uint32_t*)&c->error = 3u;
            *(volatile uint32_t*)a.err_host = 3u;
            a.rank[d] = 0xFFFFFFFFu;
            continue;
        }
        const uint32_t k = (uint32_t)((iz * g.gy + iy) * g.gx + ix);
        a.rank[d] = atomicAdd(&a.cnt[k], 1u);  // (the reorder recomputes the key from the position it moves)
    }
    if (gtid == 0) {
        c->n_src = (uint32_t)a.np + total;
        if (*(volatile uint32_t*)&c->error == 2u) *(volatile uint32_t*)a.err_host = 2u;
        a.mig->count[0] = 0; a.mig->count[1] = 0; a.mig->overflow = 0;
    }
}

// All-rank gather of the solver inputs (replicated projection): after an all-to-all handshake every rank pulls, from every
// other rank, the planes that rank owns of flags / v2 / avgPNum into its full-grid arrays; a second handshake keeps the
// sources untouched until everybody has read them.
struct GatherCopy { void* dst; const void* src; size_t bytes; };
struct GatherArgs {
    DistComm* comm;
    DistComm* all[DIST_MAX_RANKS];
    uint32_t* err_host;
    const PcgScalars* sc;  // non-null inside the PCG loop: skip once the solve is done
    int rank, nranks, ncopy;
    GatherCopy cp[DIST_MAX_RANKS * 5];
};

__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ GatherArgs a) {
    if (a.sc && a.sc->done) return;
    DistComm* c = a.comm;
    __shared__ uint32_t ep_s;
    if (threadIdx.x == 0) ep_s = *(volatile uint32_t*)&c->gather_epoch + 1u;
    __syncthreads();
    const uint32_t ep = ep_s;
    const unsigned long long t0 = now_ns();
    if (threadIdx.x < a.nranks && threadIdx.x != a.rank) {
        const int r = threadIdx.x;
        if (blockIdx.x == 0) { __threadfence_system(); st_release_sys(&a.all[r]->arrive_all[a.rank], ep); }
        wait_ge(&c->arrive_all[r], ep, c, a.err_host, 0x50u + r);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) wait_account(c, FSIM_WAIT_GATHER, t0);
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < a.ncopy; k++) {
        const GatherCopy& cp = a.cp[k];
        if ((((size_t)cp.dst | (size_t)cp.src | cp.bytes) & 15) == 0) {
            const size_t n = cp.bytes >> 4;
            const uint4* src = reinterpret_cast<const uint4*>(cp.src);
            uint4* dst = reinterpret_cast<uint4*>(cp.dst);
            size_t i = gtid;
            for (; i + 3 * gsz < n; i += 4 * gsz) {  // four loads in flight per thread: the pull is NVLink-latency bound
                const uint4 v0 = ld_peer_u4(src + i), v1 = ld_peer_u4(src + i + gsz), v2 = ld_peer_u4(src + i + 2 * gsz), v3 = ld_peer_u4(src + i + 3 * gsz);
                dst[i] = v0; dst[i + gsz] = v1; dst[i + 2 * gsz] = v2; dst[i + 3 * gsz] = v3;
            }
            for (; i < n; i += gsz) dst[i] = ld_peer_u4(src + i);
        } else {
            for (size_t i = gtid; i < cp.bytes; i += gsz) reinterpret_cast<uint8_t*>(cp.dst)[i] = reinterpret_cast<const volatile uint8_t*>(cp.src)[i];
        }
    }
    __syncthreads();
    __shared__ int last_s;
    if (threadIdx.x == 0) {
        __threadfence();
        last_s = atomicAdd(&c->gather_blocks_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_s) {  // last block: nobody may overwrite its planes before every rank has read them
        const unsigned long long t0_done = now_ns();
        if (threadIdx.x < a.nranks && threadIdx.x != a.rank) {
            const int r = threadIdx.x;
            __threadfence_system();
            st_release_sys(&a.all[r]->done_all[a.rank], ep);
            wait_ge(&c->done_all[r], ep, c, a.err_host, 0x60u + r);
        }
        __syncthreads();
        if (threadIdx.x == 0) { wait_account(c, FSIM_WAIT_GATHER, t0_done); c->gather_blocks_done = 0; c->gather_epoch = ep; __threadfence(); }
    }
}

template <typename T>
int dalloc(fsim* h, T** p, size_t n) {
    *p = nullptr;
    FSIM_CUDA(h, cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    FSIM_CUDA(h, cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    return FSIM_OK;
}

size_t elem_size(int arr) { return arr == ARR_FLAGS ? 1 : (arr == ARR_P ? 8 : 4); }

int check_err(fsim* h) {
    DistState* d = h->dist;
    const uint32_t e = *(volatile uint32_t*)d->err_host;
    if (!e) return FSIM_OK;
    h->sticky = FSIM_ERR_COMM;
    const uint32_t* eh = d->err_host;
    if (e == 1)
        return fsim_fail(h, FSIM_ERR_COMM, "slab exchange timed out waiting for a neighbour (rank %d of %d; wait 0x%x, epoch %u, flag %u)", d->rank,
                         d->nranks, eh[1], eh[2], eh[3]);
    return fsim_fail(h, FSIM_ERR_COMM, e == 2 ? "emigrant list overflow (rank %d of %d): more particles crossed a slab boundary in one step than the list holds"
                                              : "a particle crossed a whole slab in one step (rank %d of %d)", d->rank, d->nranks);
}

}  // namespace

static DistState* dist_of(const fsim* h) { return h->dist ? h->dist : (h->parent ? h->parent->dist : nullptr); }

MigDev* dist_mig_dev(const fsim* h) { return h->dist ? h->dist->mig : nullptr; }
int64_t dist_mig_capacity(const fsim* h) { return h->dist ? h->dist->mig_cap : 0; }
const uint32_t* dist_nsrc_dev(fsim* h) {
    DistState* d = h->dist;
    if (!d || !d->nsrc_valid) return nullptr;
    d->nsrc_valid = false;
    return &d->comm->n_src;
}

void dist_rank(const fsim* h, int* rank, int* nranks) { if (h->dist) { *rank = h->dist->rank; *nranks = h->dist->nranks; } }

bool dist_peer_in_process(const fsim* h) { const DistState* d = dist_of(h); return d && d->peer_in_process; }

int dist_wait_stats(fsim* h, FsimDistWaitStats* out, int reset) {
    DistState* d = dist_of(h);
    if (!d) return fsim_fail(h, FSIM_ERR_INVALID, "not a slab handle (fsim_create_slab with nranks > 1)");
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    unsigned long long buf[3 * FSIM_WAIT_CLASSES];
    char* base = (char*)d->comm + offsetof(DistComm, wait_ns);
    FSIM_CUDA(h, cudaMemcpy(buf, base, sizeof(buf), cudaMemcpyDeviceToHost));
    for (int k = 0; k < FSIM_WAIT_CLASSES; k++) { out->wait_ns[k] = buf[k]; out->waits[k] = buf[FSIM_WAIT_CLASSES + k]; out->kernel_ns[k] = buf[2 * FSIM_WAIT_CLASSES + k]; }
    if (reset) FSIM_CUDA(h, cudaMemset(base, 0, sizeof(buf)));
    return FSIM_OK;
}

int dist_check(fsim* h) { return (h->dist && h->dist->connected) ? check_err(h) : FSIM_OK; }

// the slab geometry of rank r of n over gzg global planes
void dist_partition(int gzg, int rank, int nranks, int* own_lo, int* own_hi, int* zoff, int* gz_local) {
    // slab boundaries are even planes: the 2x2x2 aggregates of the multigrid never straddle two ranks
    const int half = gzg / 2;
    const int lo = rank == 0 ? 0 : 2 * (int)((int64_t)half * rank / nranks);
    const int hi = rank == nranks - 1 ? gzg : 2 * (int)((int64_t)half * (rank + 1) / nranks);
    const int z0 = std::max(lo - 1, 0), z1 = std::min(hi + 1, gzg);
    *own_lo = lo; *own_hi = hi; *zoff = z0; *gz_local = z1 - z0;
}

int dist_init(fsim* h, int rank, int nranks, int own_lo, int own_hi) {
    DistState* d = new DistState();  // value-initialised: every POD member is zero
    h->dist = d;
    d->rank = rank; d->nranks = nranks; d->own_lo = own_lo; d->own_hi = own_hi;
    const GridDims& g = h->g;
    int rc = dalloc(h, &d->comm, 1);
    if (rc) return rc;
    DistComm c0;
    memset(&c0, 0, sizeof(c0));
    const char* e = getenv("FSIM_DIST_TIMEOUT_MS");
    c0.timeout_ns = (unsigned long long)(e ? atoll(e) : 20000) * 1000000ull;
    FSIM_CUDA(h, cudaMemcpyAsync(d->comm, &c0, sizeof(c0), cudaMemcpyHostToDevice, h->stream));
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    FSIM_CUDA(h, cudaHostAlloc((void**)&d->err_host, 4 * sizeof(uint32_t), cudaHostAllocMapped));
    memset(d->err_host, 0, 4 * sizeof(uint32_t));
    FSIM_CUDA(h, cudaHostGetDevicePointer((void**)&d->err_dev, d->err_host, 0));
    // up to two planes of 8 particles per cell may change owner in one step before the lists overflow (FSIM_DIST_MIG_CAP overrides)
    const char* mc = getenv("FSIM_DIST_MIG_CAP");
    d->mig_cap = mc ? atoll(mc) : std::max<int64_t>(4096, (int64_t)g.gx * g.gy * 16);
    for (int k = 0; k < 2; k++) if ((rc = dalloc(h, &d->mig_idx[k], (size_t)d->mig_cap))) return rc;
    if ((rc = dalloc(h, &d->mig, 1))) return rc;
    MigDev m;
    memset(&m, 0, sizeof(m));
    m.cap = (uint32_t)d->mig_cap; m.idx[0] = d->mig_idx[0]; m.idx[1] = d->mig_idx[1]; m.own_lo = own_lo; m.own_hi = own_hi;
    FSIM_CUDA(h, cudaMemcpyAsync(d->mig, &m, sizeof(m), cudaMemcpyHostToDevice, h->stream));
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    if ((rc = dalloc(h, &d->recv, (size_t)2 * (DIST_NCH + 1) * d->mig_cap))) return rc;
    if ((rc = dalloc(h, &d->stage, (size_t)2 * 2 * 7 * g.sz))) return rc;
    d->pa_capa = ((std::max<int64_t>(4096, (int64_t)g.sz * 16) + 3) / 4) * 4;
    d->pa_head = (((int64_t)g.sz + 2 + 3) / 4) * 4;
    d->pa_words = d->pa_head + 3 * d->pa_capa;
    if ((rc = dalloc(h, &d->pa_stage, (size_t)2 * d->pa_words))) return rc;
    if ((rc = dalloc(h, &d->pa_ghost, (size_t)2 * d->pa_words))) return rc;
    for (int a = 0; a < 3; a++) { d->local_arr[ARR_U + a] = h->u[a]; d->local_arr[ARR_U2 + a] = h->u2[a]; d->local_arr[ARR_W + a] = h->wsum[a]; }
    d->local_arr[ARR_DENS] = h->dens; d->local_arr[ARR_CNT] = h->cnt; d->local_arr[ARR_FLAGS] = h->flags;
    d->local_arr[ARR_P] = h->p; d->local_arr[ARR_S] = h->s; d->local_arr[ARR_COMM] = d->comm; d->local_arr[ARR_RECV] = d->recv;
    d->local_arr[ARR_PA] = d->pa_stage;
    d->all_comm[rank] = d->comm;
    return FSIM_OK;
}

void dist_free(fsim* h) {
    if (h->dist && h->dist->gx_table) { cudaFree(h->dist->gx_table); h->dist->gx_table = nullptr; }
    if (h->dist && h->dist->ar_table) { cudaFree(h->dist->ar_table); h->dist->ar_table = nullptr; }
    DistState* d = h->dist;
    if (!d) return;
    for (void* p : d->ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(d->pa_stage); cudaFree(d->pa_ghost);
    cudaFree(d->comm); cudaFree(d->mig); cudaFree(d->mig_idx[0]); cudaFree(d->mig_idx[1]); cudaFree(d->recv); cudaFree(d->stage);
    if (d->err_host) cudaFreeHost(d->err_host);
    delete d;
    h->dist = nullptr;
}

int dist_export(fsim* h, FsimDistExport* out) {
    DistState* d = h->dist;
    memset(out, 0, sizeof(*out));
    out->rank = d->rank; out->nranks = d->nranks; out->device = h->device; out->pid = (int64_t)getpid();
    out->zown0 = h->g.zown0; out->zown1 = h->g.zown1; out->gz_local = h->g.gz; out->z_offset = h->g.zoff;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    for (int k = 0; k < DIST_NARR; k++) {
        if (!d->local_arr[k]) continue;
        out->raw[k] = (uint64_t)(uintptr_t)d->local_arr[k];
        out->offset[k] = (uint64_t)d->local_off[k];
        cudaIpcMemHandle_t hd;
        const cudaError_t e = cudaIpcGetMemHandle(&hd, d->local_arr[k]);
        if (e == cudaSuccess) memcpy(out->ipc[k], &hd, 64);
        else { cudaGetLastError(); out->ipc_missing = 1; }  // same-process connections do not need it
    }
    return FSIM_OK;
}

int dist_connect(fsim* h, const FsimDistExport* all, int n) {
    DistState* d = h->dist;
    if (n != d->nranks) return fsim_fail(h, FSIM_ERR_INVALID, "expected %d exports, got %d", d->nranks, n);
    const int64_t pid = (int64_t)getpid();
    auto map_base = [&](const FsimDistExport& ex, int arr, void** out) -> int {
        if (ex.pid == pid) {  // same process: plain pointers (peer access when the handle lives on another device)
            if (ex.device != h->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(ex.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fsim_fail(h, FSIM_ERR_COMM, "cudaDeviceEnablePeerAccess(%d): %s", ex.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            *out = (void*)(uintptr_t)ex.raw[arr];
            return FSIM_OK;
        }
        if (ex.ipc_missing) return fsim_fail(h, FSIM_ERR_COMM, "rank %d exported no IPC handles", ex.rank);
        cudaIpcMemHandle_t hd;
        memcpy(&hd, ex.ipc[arr], 64);
        const cudaError_t e = cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fsim_fail(h, FSIM_ERR_COMM, "cudaIpcOpenMemHandle(rank %d, array %d): %s", ex.rank, arr, cudaGetErrorString(e));
        d->ipc_opened.push_back(*out);
        return FSIM_OK;
    };
    auto map = [&](const FsimDistExport& ex, int arr, void** out) -> int {
        *out = nullptr;
        if (!ex.raw[arr]) return FSIM_OK;  // not published by that rank (e.g. no hybrid solver context)
        int rc = map_base(ex, arr, out);
        if (rc) return rc;
        *out = (char*)*out + ex.offset[arr];
        return FSIM_OK;
    };
    for (int r = 0; r < n; r++) {
        const FsimDistExport& ex = all[r];
        if (ex.rank != r || ex.nranks != n) return fsim_fail(h, FSIM_ERR_INVALID, "export %d is from rank %d of %d", r, ex.rank, ex.nranks);
        d->all_zown0[r] = ex.zown0; d->all_zown1[r] = ex.zown1; d->all_zoff[r] = ex.z_offset;
        if (r == d->rank) {
            for (int k = 0; k < DIST_NARR; k++) d->all_arr[r][k] = d->local_arr[k] ? (char*)d->local_arr[k] + d->local_off[k] : nullptr;
            continue;
        }
        if (ex.pid == pid) d->peer_in_process = true;
        const int side = r == d->rank - 1 ? 0 : (r == d->rank + 1 ? 1 : -1);
        for (int k = 0; k < DIST_NARR; k++) {
            if (k == ARR_RECV && side < 0) continue;  // only neighbours push particles
            int rc = map(ex, k, &d->all_arr[r][k]);
            if (rc) return rc;
        }
        d->all_comm[r] = (DistComm*)d->all_arr[r][ARR_COMM];
        if (side >= 0) {
            for (int k = 0; k < DIST_NARR; k++) d->peer_arr[side][k] = d->all_arr[r][k];
            d->peer_comm[side] = d->all_comm[r];
            // the lower neighbour's ghost plane towards us is its plane zown1, its boundary plane zown1 - 1; the upper one's: zown0 - 1, zown0
            d->peer_ghost[side] = side == 0 ? ex.zown1 : ex.zown0 - 1;
            d->peer_bnd[side] = side == 0 ? ex.zown1 - 1 : ex.zown0;
        }
    }
    if (d->peer_in_process) {
        // Ranks that share a process spin on each other's flags from kernels of ONE CUDA context: with lazy module loading the
        // first launch of a kernel waits for the device to go idle, which never happens while a peer's kernel spins -> dead wait
        // (profiles/r1_concurrency_probe_lazy_loading.txt).  Refuse the connection instead of hanging in the first step.
        const char* e = getenv("CUDA_MODULE_LOADING");
        if (!(e && (strcmp(e, "EAGER") == 0 || strcmp(e, "eager") == 0)))
            return fsim_fail(h, FSIM_ERR_INVALID, "slab ranks that share a process need CUDA_MODULE_LOADING=EAGER in the environment before CUDA is "
                                                   "initialised (one process per GPU has no such requirement)");
    }
    // peer table of the fused all-rank push (fexch.cuh); allocated here, not in the solve: a cudaMalloc may synchronise the device
    if (h->solver && n <= 16) {
        FxAllTable t;
        memset(&t, 0, sizeof(t));
        bool ok = true;
        for (int r = 0; r < n && ok; r++) {
            if (r == d->rank) continue;
            if (!d->all_arr[r][ARR_HS_B1] || !d->all_comm[r]) { ok = false; break; }
            t.peer[t.n] = (float*)d->all_arr[r][ARR_HS_B1];
            t.peer_cnt[t.n] = &d->all_comm[r]->gx_cnt;
            t.n++;
        }
        if (ok && t.n > 0) {
            if (d->gx_table) { cudaFree(d->gx_table); d->gx_table = nullptr; }
            FSIM_CUDA(h, cudaMalloc((void**)&d->gx_table, sizeof(t)));
            FSIM_CUDA(h, cudaMemcpy(d->gx_table, &t, sizeof(t), cudaMemcpyHostToDevice));
        }
    }
    if (h->solver) {  // rank table of the all-rank reduction fused into the reducing kernels
        ArDev t;
        memset(&t, 0, sizeof(t));
        t.comm = d->comm; t.err_host = d->err_dev; t.rank = d->rank; t.nranks = n;
        for (int r = 0; r < n; r++) t.all[r] = d->all_comm[r];
        if (d->ar_table) { cudaFree(d->ar_table); d->ar_table = nullptr; }
        FSIM_CUDA(h, cudaMalloc((void**)&d->ar_table, sizeof(t)));
        FSIM_CUDA(h, cudaMemcpy(d->ar_table, &t, sizeof(t), cudaMemcpyHostToDevice));
    }
    d->connected = true;
    return FSIM_OK;
}

// the rank table for ar_warp, or nullptr when the reductions stay kernels of their own (FSIM_SLAB_FUSED_AR=0, FSIM_SLAB_FUSED=0)
const ArDev* dist_ar_dev(const fsim* hs) {
    const DistState* d = dist_of(hs);
    if (!d || !d->connected || !hs->hybrid || !fused_enabled()) return nullptr;
    const char* e = getenv("FSIM_SLAB_FUSED_AR");
    if (e && e[0] == '0') return nullptr;
    return d->ar_table;
}

int dist_halo(fsim* h, int what, bool in_pcg_loop) {
    DistState* d = h->dist;
    if (!d || !d->connected) return fsim_fail(h, FSIM_ERR_COMM, "slab handle is not connected to its neighbours (fsim_dist_connect)");
    const GridDims& g = h->g;
    HaloArgs a;
    memset(&a, 0, sizeof(a));
    a.comm = d->comm; a.peer[0] = d->peer_comm[0]; a.peer[1] = d->peer_comm[1]; a.err_host = d->err_dev;
    a.sc = in_pcg_loop ? h->scal : nullptr;
    const size_t plane = (size_t)g.sz;
    const int my_ghost[2] = {g.zown0 - 1, g.zown1}, my_bnd[2] = {g.zown0, g.zown1 - 1};
    auto pull = [&](int arr) {  // my ghost plane <- the neighbour's boundary plane
        const size_t es = elem_size(arr);
        for (int side = 0; side < 2; side++) {
            if (!d->peer_comm[side]) continue;
            HaloCopy& c = a.cp[a.ncopy++];
            c.dst = (char*)d->local_arr[arr] + (size_t)my_ghost[side] * plane * es;
            c.src = (const char*)d->peer_arr[side][arr] + (size_t)d->peer_bnd[side] * plane * es;
            c.bytes = (uint32_t)(plane * es);
            c.side = side;
            c.mask = -1;
        }
    };
    size_t bytes = 0;
    switch (what) {
        case HALO_P: pull(ARR_P); break;
        case HALO_S: pull(ARR_S); break;
        case HALO_U2: for (int ax = 0; ax < 3; ax++) pull(ARR_U2 + ax); break;
        case HALO_U2_FLAGS: for (int ax = 0; ax < 3; ax++) pull(ARR_U2 + ax); pull(ARR_FLAGS); break;
        case HALO_P2G: {
            pull(ARR_CNT);
            const int chan[7] = {ARR_U, ARR_U + 1, ARR_U + 2, ARR_W, ARR_W + 1, ARR_W + 2, ARR_DENS};
            for (int side = 0; side < 2; side++) {
                if (!d->peer_comm[side]) continue;
                for (int which = 0; which < 2; which++)  // 0: the neighbour's ghost plane (its scatter into my boundary plane), 1: its boundary plane
                    for (int ch = 0; ch < 7; ch++) {
                        HaloCopy& c = a.cp[a.ncopy++];
                        c.dst = d->stage + ((size_t)(side * 2 + which) * 7 + ch) * plane;
                        c.src = (const float*)d->peer_arr[side][chan[ch]] + (size_t)(which == 0 ? d->peer_ghost[side] : d->peer_bnd[side]) * plane;
                        c.bytes = (uint32_t)(plane * sizeof(float));
                        c.side = side;
                        c.mask = -1;
                    }
            }
            break;
        }
        case HALO_BASIC_ODD: case HALO_BASIC_EVEN: {
            // BasicMacGrid sweep of one colour (grid.cu basic_sor_kernel) just ran on the owned planes.  The z faces between two
            // slabs live in the lower rank's top plane: cells of this colour there were updated by their owner, the others by the
            // upper rank's cells through their -z face (its ghost plane).  Each rank takes the half the other one wrote.
            const int colour = what == HALO_BASIC_ODD ? 1 : 0;
            const int zg_lo = g.zoff + g.zown0 - 1, zg_hi = g.zoff + g.zown1;  // global planes of my ghost planes
            if (d->peer_comm[0]) {  // my ghost plane below <- the lower rank's top plane: the cells it owns and swept (x + y + zg_lo of this colour)
                HaloCopy& c = a.cp[a.ncopy++];
                c.dst = (float*)d->local_arr[ARR_U2 + 2] + (size_t)my_ghost[0] * plane;
                c.src = (const float*)d->peer_arr[0][ARR_U2 + 2] + (size_t)d->peer_bnd[0] * plane;
                c.bytes = (uint32_t)(plane * sizeof(float)); c.side = 0; c.gx = g.gx;
                c.mask = colour ^ (zg_lo & 1);
            }
            if (d->peer_comm[1]) {  // my top plane <- the upper rank's ghost plane: the faces its cells of this colour (plane zg_hi) pushed on
                HaloCopy& c = a.cp[a.ncopy++];
                c.dst = (float*)d->local_arr[ARR_U2 + 2] + (size_t)my_bnd[1] * plane;
                c.src = (const float*)d->peer_arr[1][ARR_U2 + 2] + (size_t)d->peer_ghost[1] * plane;
                c.bytes = (uint32_t)(plane * sizeof(float)); c.side = 1; c.gx = g.gx;
                c.mask = colour ^ (zg_hi & 1);
            }
            break;
        }
        default: return fsim_fail(h, FSIM_ERR_INVALID, "unknown halo %d", what);
    }
    for (int k = 0; k < a.ncopy; k++) bytes += a.cp[k].bytes;
    const int blocks = (int)std::min<size_t>(128, std::max<size_t>(4, bytes / (16 * 256 * 4)));
    { KScope ks(h, K_HALO); halo_kernel<<<blocks, 256, 0, h->stream>>>(a); }
    if (what == HALO_P2G) {
        HaloAddArgs b;
        memset(&b, 0, sizeof(b));
        const int chan[7] = {ARR_U, ARR_U + 1, ARR_U + 2, ARR_W, ARR_W + 1, ARR_W + 2, ARR_DENS};
        for (int side = 0; side < 2; side++) {
            b.has[side] = d->peer_comm[side] != nullptr;
            for (int ch = 0; ch < 7; ch++) {
                float* base = (float*)d->local_arr[chan[ch]];
                b.dst[side][0][ch] = base + (size_t)my_bnd[side] * plane;
                b.dst[side][1][ch] = base + (size_t)(b.has[side] ? my_ghost[side] : my_bnd[side]) * plane;
                b.stg[side][0][ch] = d->stage + ((size_t)(side * 2 + 0) * 7 + ch) * plane;
                b.stg[side][1][ch] = d->stage + ((size_t)(side * 2 + 1) * 7 + ch) * plane;
            }
        }
        b.plane = (int)plane;
        KScope ks(h, K_HALO);
        halo_add_kernel<<<dim3(div_up(plane, 256), 4), 256, 0, h->stream>>>(b);
    }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

// ---- push-apart across slabs: the neighbour search of a particle in a boundary plane reaches into the neighbour's boundary
// plane (search radius one cell).  After the sort the particles of a plane are one contiguous run, so each rank stages the
// positions + per-cell starts of its two boundary planes and pulls the neighbours' (two-sided handshake, halo_kernel).
struct PaPackArgs {
    const float *px, *py, *pz;
    const uint32_t* cell_start;
    uint32_t* out[2];       // stage block per side
    int64_t cell0[2];       // first cell of my boundary plane on that side
    int has[2];
    int sz;
    uint32_t capa, head;
};
__global__ void __launch_bounds__(256) pa_pack_kernel(const __grid_constant__ PaPackArgs a) {
    const int side = blockIdx.y;
    if (side == 0 ? !a.has[0] : !a.has[1]) return;
    uint32_t* out = side == 0 ? a.out[0] : a.out[1];
    const int64_t c0 = side == 0 ? a.cell0[0] : a.cell0[1];
    const uint32_t base = a.cell_start[c0], end = a.cell_start[c0 + a.sz];
    const uint32_t n = min(end - base, a.capa);  // (a plane denser than 16 particles per cell on average is truncated)
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    if (gtid == 0) out[0] = n;
    for (size_t i = gtid; i <= (size_t)a.sz; i += gsz) out[1 + i] = min(a.cell_start[c0 + i] - base, n);
    float* x = reinterpret_cast<float*>(out + a.head);
    float *y = x + a.capa, *z = y + a.capa;
    for (size_t i = gtid; i < n; i += gsz) { x[i] = a.px[base + i]; y[i] = a.py[base + i]; z[i] = a.pz[base + i]; }
}

int dist_push_apart_ghosts(fsim* h) {
    DistState* d = h->dist;
    if (!d || !d->connected) return fsim_fail(h, FSIM_ERR_COMM, "slab handle is not connected to its neighbours (fsim_dist_connect)");
    const GridDims& g = h->g;
    const ParticleSet& p = h->ps[h->cur];
    PaPackArgs k;
    memset(&k, 0, sizeof(k));
    k.px = p.pos[0]; k.py = p.pos[1]; k.pz = p.pos[2]; k.cell_start = h->cell_start;
    k.sz = g.sz; k.capa = (uint32_t)d->pa_capa; k.head = (uint32_t)d->pa_head;
    const int bnd[2] = {g.zown0, g.zown1 - 1};
    HaloArgs a;
    memset(&a, 0, sizeof(a));
    a.comm = d->comm; a.peer[0] = d->peer_comm[0]; a.peer[1] = d->peer_comm[1]; a.err_host = d->err_dev;
    for (int side = 0; side < 2; side++) {
        k.has[side] = d->peer_comm[side] != nullptr;
        k.out[side] = d->pa_stage + (size_t)side * d->pa_words;
        k.cell0[side] = (int64_t)bnd[side] * g.sz;
        if (!d->peer_comm[side]) continue;
        if (!d->peer_arr[side][ARR_PA]) return fsim_fail(h, FSIM_ERR_COMM, "neighbour did not publish its push-apart staging buffer");
        // the neighbour's block that faces me: its side 1 - side
        const uint32_t* src = (const uint32_t*)d->peer_arr[side][ARR_PA] + (size_t)(1 - side) * d->pa_words;
        uint32_t* dst = d->pa_ghost + (size_t)side * d->pa_words;
        HaloCopy& c0 = a.cp[a.ncopy++];
        c0.dst = dst; c0.src = src; c0.bytes = (uint32_t)(d->pa_head * 4); c0.side = side; c0.mask = -1;
        for (int ax = 0; ax < 3; ax++) {
            HaloCopy& c = a.cp[a.ncopy++];
            c.dst = dst + d->pa_head + (size_t)ax * d->pa_capa;
            c.src = src + d->pa_head + (size_t)ax * d->pa_capa;
            c.bytes = (uint32_t)(d->pa_capa * 4); c.side = side; c.mask = -1;
            c.limit = src;  // word 0 of the neighbour's block: its particle count
        }
    }
    {
        KScope ks(h, K_HALO, 2);
        if (h->np > 0) pa_pack_kernel<<<dim3(64, 2), 256, 0, h->stream>>>(k);
        else FSIM_CUDA(h, cudaMemsetAsync(d->pa_stage, 0, sizeof(uint32_t) * 2 * d->pa_words, h->stream));  // an empty slab: counts and starts 0
        halo_kernel<<<64, 256, 0, h->stream>>>(a);
    }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

// the pulled block of neighbour `side` (nullptr: no neighbour): per-cell starts of its boundary plane and the positions
bool dist_pa_ghost(const fsim* h, int side, const uint32_t** starts, const float** x, const float** y, const float** z) {
    const DistState* d = h->dist;
    if (!d || !d->peer_comm[side]) return false;
    const uint32_t* blk = d->pa_ghost + (size_t)side * d->pa_words;
    *starts = blk + 1;
    *x = reinterpret_cast<const float*>(blk + d->pa_head); *y = *x + d->pa_capa; *z = *y + d->pa_capa;
    return true;
}

int dist_allreduce(fsim* h, int kind, bool in_pcg_loop) {
    DistState* d = dist_of(h);
    if (!d || !d->connected) return fsim_fail(h, FSIM_ERR_COMM, "slab handle is not connected to its neighbours (fsim_dist_connect)");
    ArArgs a;
    memset(&a, 0, sizeof(a));
    a.comm = d->comm;
    for (int r = 0; r < d->nranks; r++) a.all[r] = d->all_comm[r];
    a.err_host = d->err_dev; a.sc = h->scal; a.status = h->status_dev;
    a.rank = d->rank; a.nranks = d->nranks; a.kind = kind; a.check_done = in_pcg_loop ? 1 : 0;
    { KScope ks(h, K_ALLREDUCE); launch_k(h, allreduce_kernel, dim3(1), dim3(32), 0, a); }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

int dist_gather_solver_inputs(fsim* h) {
    DistState* d = h->dist;
    fsim* hs = h->solver;
    if (!d || !d->connected || !hs) return fsim_fail(h, FSIM_ERR_COMM, "slab handle is not connected to its neighbours (fsim_dist_connect)");
    GatherArgs a;
    memset(&a, 0, sizeof(a));
    a.comm = d->comm; a.err_host = d->err_dev; a.rank = d->rank; a.nranks = d->nranks;
    for (int r = 0; r < d->nranks; r++) a.all[r] = d->all_comm[r];
    const size_t plane = (size_t)h->g.sz;
    const int arr[5] = {ARR_FLAGS, ARR_U2, ARR_U2 + 1, ARR_U2 + 2, ARR_DENS};
    void* dst[5] = {hs->flags, hs->u2[0], hs->u2[1], hs->u2[2], hs->dens};
    size_t bytes = 0;
    // hybrid: the right-hand side only exists on the owned planes, whose inputs (and the plane below) the slab already holds;
    // only the cell types of the other ranks are needed (the coarse multigrid operators are global)
    const int narr = hs->hybrid ? 1 : 5;
    if (hs->hybrid) {
        float* src4[4] = {h->u2[0], h->u2[1], h->u2[2], h->dens};
        float* dst4[4] = {hs->u2[0], hs->u2[1], hs->u2[2], hs->dens};
        for (int k = 0; k < 4; k++)
            FSIM_CUDA(h, cudaMemcpyAsync(dst4[k] + (size_t)h->g.zoff * plane, src4[k], sizeof(float) * plane * h->g.gz, cudaMemcpyDeviceToDevice, h->stream));
    }
    for (int r = 0; r < d->nranks; r++) {
        const int own_lo = d->all_zoff[r] + d->all_zown0[r], planes = d->all_zown1[r] - d->all_zown0[r];
        for (int k = 0; k < narr; k++) {
            const size_t es = elem_size(arr[k]);
            GatherCopy& c = a.cp[a.ncopy++];
            c.dst = (char*)dst[k] + (size_t)own_lo * plane * es;
            c.src = (const char*)d->all_arr[r][arr[k]] + (size_t)d->all_zown0[r] * plane * es;
            c.bytes = (size_t)planes * plane * es;
            bytes += c.bytes;
        }
    }
    int blocks = (int)std::min<size_t>((size_t)h->sm_count * 4, std::max<size_t>(4, bytes / (16 * 256 * 8)));
    // several ranks sharing ONE device (tests): all their spinning gather blocks must be resident at the same time
    if (const char* e = getenv("FSIM_DIST_GATHER_BLOCKS")) blocks = std::max(1, std::min(blocks, atoi(e)));
    { KScope ks(h, K_HALO); gather_kernel<<<blocks, 256, 0, h->stream>>>(a); }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

// ---- hybrid projection: the solver context stores every array at global plane indices, so a neighbour's plane z lives at
// the same offset in the neighbour's array -------------------------------------------------------------------------------
int dist_register_solver(fsim* h) {
    DistState* d = h->dist;
    fsim* hs = h->solver;
    if (!d || !hs) return FSIM_OK;
    d->local_arr[ARR_HS_S] = hs->s; d->local_arr[ARR_HS_P] = hs->p;
    float* base = nullptr;
    size_t pad = 0;
    mg_level_array(hs, 0, 0, &base, &pad); d->local_arr[ARR_HS_XA] = base; d->local_off[ARR_HS_XA] = pad * sizeof(float);
    mg_level_array(hs, 0, 1, &base, &pad); d->local_arr[ARR_HS_XB] = base; d->local_off[ARR_HS_XB] = pad * sizeof(float);
    mg_level_array(hs, 1, 2, &base, &pad); d->local_arr[ARR_HS_B1] = base; d->local_off[ARR_HS_B1] = pad * sizeof(float);
    return FSIM_OK;
}

static bool push_enabled() { const char* e = getenv("FSIM_SLAB_PUSH"); return !(e && e[0] == '0'); }

// push form of dist_halo_sym / dist_gather_coarse (see push_kernel for when it is safe)
static int halo_sym_push(fsim* hs, DistState* d, int arr, size_t es) {
    const size_t plane = (size_t)hs->g.sz;
    PushArgs a;
    memset(&a, 0, sizeof(a));
    a.comm = d->comm; a.peer[0] = d->peer_comm[0]; a.peer[1] = d->peer_comm[1]; a.err_host = d->err_dev; a.sc = hs->scal;
    a.npeer = 2; a.self = d->rank; a.all_ranks = 0;
    const int zpl[2] = {d->own_lo, d->own_hi - 1};  // my boundary planes: the lower one goes to the lower neighbour, ...
    const char* mine = (const char*)d->local_arr[arr] + d->local_off[arr];
    size_t bytes = 0;
    for (int side = 0; side < 2; side++) {
        if (!d->peer_comm[side]) continue;
        if (!d->peer_arr[side][arr]) return fsim_fail(hs, FSIM_ERR_COMM, "neighbour did not publish its solver arrays (mixed FSIM_SLAB_SOLVER settings?)");
        PushCopy& c = a.cp[a.ncopy++];
        c.src = mine + (size_t)zpl[side] * plane * es;
        c.dst = (char*)d->peer_arr[side][arr] + (size_t)zpl[side] * plane * es;  // same global plane index in the neighbour's array
        c.bytes = plane * es;
        c.to = side;
        bytes += c.bytes;
    }
    const int blocks = (int)std::min<size_t>(128, std::max<size_t>(2, bytes / (16 * 256 * 2)));
    { KScope ks(hs, K_HALO); launch_k(hs, push_kernel, dim3(blocks), dim3(256), 0, a); }
    FSIM_CHECK_LAUNCH(hs);
    return FSIM_OK;
}

// fused form of the in-loop exchange of one level-0 array (fexch.cuh): arguments for the kernels that write it (FxPush) and
// for those that read its ghost planes (FxWait).  false: not available (then both are inert and dist_halo_sym does the job)
static bool fused_enabled() { const char* e = getenv("FSIM_SLAB_FUSED"); return !(e && e[0] == '0'); }

bool dist_fx(fsim* hs, int which, const void* ptr, FxPush* p, FxWait* w) {
    memset(p, 0, sizeof(*p));
    memset(w, 0, sizeof(*w));
    DistState* d = dist_of(hs);
    if (!d || !d->connected || !hs->hybrid || !push_enabled() || !fused_enabled()) return false;
    if (hs->g.gx % 4 != 0 || hs->g.gy % 32 != 0 || (d->own_hi - d->own_lo) % 2 != 0 || d->own_hi - d->own_lo < 2) return false;
    int arr = -1, ix = -1;
    if (which == SYM_S) { arr = ARR_HS_S; ix = 0; }
    if (which == SYM_X) {
        const void* xa = (char*)d->local_arr[ARR_HS_XA] + d->local_off[ARR_HS_XA];
        const void* xb = (char*)d->local_arr[ARR_HS_XB] + d->local_off[ARR_HS_XB];
        if (ptr == xa) { arr = ARR_HS_XA; ix = 1; }
        if (ptr == xb) { arr = ARR_HS_XB; ix = 2; }
    }
    if (arr < 0) return false;
    for (int side = 0; side < 2; side++)
        if (d->peer_comm[side] && !d->peer_arr[side][arr]) return false;
    p->my_exp = &d->comm->fx_exp[ix][0];
    w->cnt = &d->comm->fx_cnt[ix][0];
    w->exp = &d->comm->fx_exp[ix][0];
    w->error = &d->comm->error;
    w->err_host = d->err_dev;
    w->stat = &d->comm->wait_ns[FSIM_WAIT_FUSED];
    w->stat_n = &d->comm->waits[FSIM_WAIT_FUSED];
    const char* e = getenv("FSIM_DIST_TIMEOUT_MS");
    w->timeout_ns = (unsigned long long)(e ? atoll(e) : 20000) * 1000000ull;
    p->zb[0] = w->zb[0] = d->own_lo;
    p->zb[1] = w->zb[1] = d->own_hi - 1;
    for (int side = 0; side < 2; side++) {
        if (!d->peer_comm[side]) continue;
        p->peer[side] = (float*)d->peer_arr[side][arr];
        p->peer_cnt[side] = &d->peer_comm[side]->fx_cnt[ix][1 - side];  // I am the upper neighbour of my lower neighbour
        w->has[side] = 1;
    }
    return true;
}

// fused form of dist_gather_coarse (fexch.cuh, all-rank form): arguments for mg_restrict4_kernel (producer) and for the first
// level-1 kernel (consumer).  (cgx, cgy, cgz): the level-1 grid -- the producer's launch geometry on every rank follows from it
bool dist_gx(fsim* hs, int cgx, int cgy, int cgz, FxAllPush* p, FxWait* w) {
    memset(p, 0, sizeof(*p));
    memset(w, 0, sizeof(*w));
    DistState* d = dist_of(hs);
    if (!d || !d->connected || !hs->hybrid || !push_enabled() || !fused_enabled()) return false;
    // measured (256^3, one box each, --steps 20): 8 GPUs 4.051 -> 3.994 ms/step, 2 GPUs 6.45 -> 6.48 -- the same bytes leave through
    // the restriction's stores instead of a copy kernel; it pays once several peers are written in parallel.  Default: on from 4
    // ranks (FSIM_SLAB_FUSED_GATHER=1 / =0 force it)
    { const char* e = getenv("FSIM_SLAB_FUSED_GATHER"); if (e ? e[0] == '0' : d->nranks < 4) return false; }
    if (d->nranks > 16 || cgx % 2 != 0 || d->local_off[ARR_HS_B1] % 8 != 0) return false;
    if (!d->gx_table) return false;  // built by dist_connect
    uint32_t add = 0;
    for (int r = 0; r < d->nranks; r++) {
        if (r == d->rank) continue;
        const int lo = d->all_zoff[r] + d->all_zown0[r], hi = d->all_zoff[r] + d->all_zown1[r];
        const int clo = lo / 2, chi = r == d->nranks - 1 ? cgz : hi / 2;  // (as in dist_gather_coarse / the launch in mg.cu cycle())
        add += (uint32_t)(((cgx + 63) / 64) * ((cgy + 3) / 4) * ((chi - clo + 1) / 2));
    }
    p->tab = d->gx_table; p->my_exp = &d->comm->gx_exp; p->add = add;
    w->cnt = &d->comm->gx_cnt; w->exp = &d->comm->gx_exp;
    w->error = &d->comm->error; w->err_host = d->err_dev;
    w->stat = &d->comm->wait_ns[FSIM_WAIT_FUSED]; w->stat_n = &d->comm->waits[FSIM_WAIT_FUSED];
    const char* e = getenv("FSIM_DIST_TIMEOUT_MS");
    w->timeout_ns = (unsigned long long)(e ? atoll(e) : 20000) * 1000000ull;
    w->has[0] = 1;
    return true;
}

int dist_halo_sym(fsim* hs, int which, const void* ptr, bool in_pcg_loop) {
    DistState* d = dist_of(hs);
    if (!d || !d->connected) return fsim_fail(hs, FSIM_ERR_COMM, "slab handle is not connected to its neighbours (fsim_dist_connect)");
    int arr = which == SYM_S ? ARR_HS_S : (which == SYM_P ? ARR_HS_P : -1);
    if (which == SYM_X) {  // the multigrid ping-pongs between two level-0 arrays
        const void* xa = (char*)d->local_arr[ARR_HS_XA] + d->local_off[ARR_HS_XA];
        const void* xb = (char*)d->local_arr[ARR_HS_XB] + d->local_off[ARR_HS_XB];
        arr = ptr == xa ? ARR_HS_XA : (ptr == xb ? ARR_HS_XB : -1);
    }
    if (arr < 0) return fsim_fail(hs, FSIM_ERR_INVALID, "halo of an array that is not published");
    const size_t es = arr == ARR_HS_P ? 8 : 4, plane = (size_t)hs->g.sz;
    // in-loop halos of the search direction and of the float4 multigrid path (whose kernels never write a ghost plane) are pushed
    if (in_pcg_loop && which != SYM_P && push_enabled() && (plane * es) % 16 == 0 && (which == SYM_S || hs->g.gx % 4 == 0))
        return halo_sym_push(hs, d, arr, es);
    HaloArgs a;
    memset(&a, 0, sizeof(a));
    a.comm = d->comm; a.peer[0] = d->peer_comm[0]; a.peer[1] = d->peer_comm[1]; a.err_host = d->err_dev;
    a.sc = in_pcg_loop ? hs->scal : nullptr;
    const int zpl[2] = {d->own_lo - 1, d->own_hi};  // the planes just outside the owned range
    char* mine = (char*)d->local_arr[arr] + d->local_off[arr];
    for (int side = 0; side < 2; side++) {
        if (!d->peer_comm[side]) continue;
        if (!d->peer_arr[side][arr]) return fsim_fail(hs, FSIM_ERR_COMM, "neighbour did not publish its solver arrays (mixed FSIM_SLAB_SOLVER settings?)");
        HaloCopy& c = a.cp[a.ncopy++];
        c.dst = mine + (size_t)zpl[side] * plane * es;
        c.src = (const char*)d->peer_arr[side][arr] + (size_t)zpl[side] * plane * es;
        c.bytes = (uint32_t)(plane * es);
        c.side = side;
        c.mask = -1;
    }
    size_t bytes = 0;
    for (int k = 0; k < a.ncopy; k++) bytes += a.cp[k].bytes;
    const int blocks = (int)std::min<size_t>(128, std::max<size_t>(4, bytes / (16 * 256 * 4)));
    { KScope ks(hs, K_HALO); halo_kernel<<<blocks, 256, 0, hs->stream>>>(a); }
    FSIM_CHECK_LAUNCH(hs);
    return FSIM_OK;
}

int dist_gather_coarse(fsim* hs, bool in_pcg_loop) {
    DistState* d = dist_of(hs);
    if (!d || !d->connected) return fsim_fail(hs, FSIM_ERR_COMM, "slab handle is not connected to its neighbours (fsim_dist_connect)");
    GatherArgs a;
    memset(&a, 0, sizeof(a));
    a.comm = d->comm; a.err_host = d->err_dev; a.rank = d->rank; a.nranks = d->nranks;
    a.sc = in_pcg_loop ? hs->scal : nullptr;
    for (int r = 0; r < d->nranks; r++) a.all[r] = d->all_comm[r];
    // level 1 has (gx+1)/2 x (gy+1)/2 cells per plane; rank r owns the coarse planes [own_lo_r / 2, own_hi_r / 2) (boundaries are even)
    const size_t cplane = (size_t)((hs->g.gx + 1) / 2) * ((hs->g.gy + 1) / 2);
    char* mine = (char*)d->local_arr[ARR_HS_B1] + d->local_off[ARR_HS_B1];
    if (in_pcg_loop && push_enabled() && hs->g.gx % 4 == 0 && (cplane * sizeof(float)) % 16 == 0 && d->local_off[ARR_HS_B1] % 16 == 0) {
        // push my coarse planes to every rank (the float4 restriction only writes the planes this rank owns)
        PushArgs p;
        memset(&p, 0, sizeof(p));
        p.comm = d->comm; p.err_host = d->err_dev; p.sc = hs->scal; p.npeer = d->nranks; p.self = d->rank; p.all_ranks = 1;
        const int clo = d->own_lo / 2, chi = d->rank == d->nranks - 1 ? (hs->g.gz + 1) / 2 : d->own_hi / 2;
        size_t pbytes = 0;
        for (int r = 0; r < d->nranks; r++) {
            p.peer[r] = d->all_comm[r];
            if (r == d->rank) continue;
            if (!d->all_arr[r][ARR_HS_B1]) return fsim_fail(hs, FSIM_ERR_COMM, "rank %d did not publish its solver arrays", r);
            PushCopy& c = p.cp[p.ncopy++];
            c.src = mine + (size_t)clo * cplane * sizeof(float);
            c.dst = (char*)d->all_arr[r][ARR_HS_B1] + (size_t)clo * cplane * sizeof(float);
            c.bytes = (size_t)(chi - clo) * cplane * sizeof(float);
            c.to = r;
            pbytes += c.bytes;
        }
        int pblocks = (int)std::min<size_t>((size_t)hs->sm_count * 2, std::max<size_t>(2, pbytes / (16 * 256 * 4)));
        if (const char* e = getenv("FSIM_DIST_GATHER_BLOCKS")) pblocks = std::max(1, std::min(pblocks, atoi(e)));
        { KScope ks(hs, K_HALO); launch_k(hs, push_kernel, dim3(pblocks), dim3(256), 0, p); }
        FSIM_CHECK_LAUNCH(hs);
        return FSIM_OK;
    }
    size_t bytes = 0;
    for (int r = 0; r < d->nranks; r++) {
        if (r == d->rank) continue;
        if (!d->all_arr[r][ARR_HS_B1]) return fsim_fail(hs, FSIM_ERR_COMM, "rank %d did not publish its solver arrays", r);
        const int lo = d->all_zoff[r] + d->all_zown0[r], hi = d->all_zoff[r] + d->all_zown1[r];
        const int clo = lo / 2, chi = r == d->nranks - 1 ? (hs->g.gz + 1) / 2 : hi / 2;
        GatherCopy& c = a.cp[a.ncopy++];
        c.dst = mine + (size_t)clo * cplane * sizeof(float);
        c.src = (const char*)d->all_arr[r][ARR_HS_B1] + (size_t)clo * cplane * sizeof(float);
        c.bytes = (size_t)(chi - clo) * cplane * sizeof(float);
        bytes += c.bytes;
    }
    int blocks = (int)std::min<size_t>((size_t)hs->sm_count * 2, std::max<size_t>(4, bytes / (16 * 256 * 4)));
    if (const char* e = getenv("FSIM_DIST_GATHER_BLOCKS")) blocks = std::max(1, std::min(blocks, atoi(e)));
    { KScope ks(hs, K_HALO); gather_kernel<<<blocks, 256, 0, hs->stream>>>(a); }
    FSIM_CHECK_LAUNCH(hs);
    return FSIM_OK;
}

int fsim_ensure_capacity(fsim* h, int64_t n);  // fsim_api.cu

int dist_migrate(fsim* h) {
    DistState* d = h->dist;
    if (!d || !d->connected) return fsim_fail(h, FSIM_ERR_COMM, "slab handle is not connected to its neighbours (fsim_dist_connect)");
    if (h->cap < h->np + 2 * d->mig_cap) return fsim_fail(h, FSIM_ERR_NOMEM, "no room for immigrants (capacity %lld, %lld particles)", (long long)h->cap, (long long)h->np);
    const ParticleSet& p = h->ps[h->cur];
    const int nch = h->have_c ? 15 : 6;
    MigSendArgs s;
    memset(&s, 0, sizeof(s));
    s.comm = d->comm; s.peer[0] = d->peer_comm[0]; s.peer[1] = d->peer_comm[1]; s.mig = d->mig;
    for (int c = 0; c < 3; c++) { s.src[c] = p.pos[c]; s.src[3 + c] = p.vel[c]; }
    for (int c = 0; c < 9; c++) s.src[6 + c] = p.c[c];
    s.src_id = h->track_ids ? p.id : nullptr;
    s.nch = nch; s.cap = d->mig_cap;
    for (int dir = 0; dir < 2; dir++)
        if (d->peer_comm[dir]) s.peer_recv[dir] = (float*)d->peer_arr[dir][ARR_RECV] + (size_t)(1 - dir) * (DIST_NCH + 1) * d->mig_cap;
    MigRecvArgs r;
    memset(&r, 0, sizeof(r));
    r.comm = d->comm; r.has[0] = d->peer_comm[0] != nullptr; r.has[1] = d->peer_comm[1] != nullptr; r.err_host = d->err_dev; r.mig = d->mig;
    r.recv = d->recv; r.cap = d->mig_cap;
    for (int c = 0; c < 3; c++) { r.dst[c] = p.pos[c]; r.dst[3 + c] = p.vel[c]; }
    for (int c = 0; c < 9; c++) r.dst[6 + c] = p.c[c];
    r.dst_id = h->track_ids ? p.id : nullptr;
    r.nch = nch; r.np = h->np; r.g = h->g; r.cnt = h->cnt; r.key = h->key; r.rank = h->rank;
    {
        KScope ks(h, K_MIGRATE, 2);
        mig_send_kernel<<<32, 256, 0, h->stream>>>(s);
        mig_recv_kernel<<<32, 256, 0, h->stream>>>(r);
    }
    FSIM_CHECK_LAUNCH(h);
    d->nsrc_valid = true;
    h->kill_pending = true;  // the particle count changes: k_sort reads the new total back
    return FSIM_OK;
}
