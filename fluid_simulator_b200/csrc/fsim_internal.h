// fsim_internal.h -- internal state of libfsim_b200 (not part of the ABI; see include/fsim.h).
//
// Device layout:
//   grid   : fp32 SoA, one array per channel, cell (x,y,z) at (z*gy + y)*gx + x  (x fastest, z slowest so that
//            z-slab halos are contiguous planes); u8 flags; fp64 pressure / PCG vectors.
//   particles : fp32 SoA (px,py,pz,vx,vy,vz[,c00..c22]), kept in cell-binned order (sorted by device cell
//            index every step); two buffer sets (ping-pong) for the reorder pass.
// The ABI transposes to/from the reference's order (x-major, z fastest; AoS fp64 particles) on upload/download.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/fsim.h"

#define FSIM_MAX_OBS 31  // 5 bits; FSIM_MAX_OBSTACLES in the ABI is 32 -> the 32nd is rejected with FSIM_ERR_INVALID

// flags byte
#define FL_TYPE_MASK 0x03  // FSIM_CELL_* (reference enum order)
#define FL_HASPART 0x04    // >= 1 particle has ivec3(pos*cellDInv) == this cell (before obstacles/borders)
#define FL_VALID_SHIFT 3   // 2 bits: extrapolation validity 0 (WATER), 1, 2, 3 (= invalid, the reference's 100)
#define FL_VALID_MASK 0x18

// 16-bit stencil code of a cell, computed once per pressure solve (pcg.cu rhs_kernel): bits 0-5 WATER neighbours
// (-x,+x,-y,+y,-z,+z), bits 6-8 number of non-solid neighbours, bit 15 the cell itself is WATER; 0 otherwise
#define CODE_ACTIVE 0x8000u

struct GridDims {
    int gx, gy, gz;     // cells
    int sy, sz;         // strides (sx = 1): sy = gx, sz = gx*gy
    int64_t nc;
    float hx, hy, hz;   // cellD
    float ihx, ihy, ihz;  // cellDInv
    double dhx, dhy, dhz, dihx, dihy, dihz;
    int twoD;
    // z-slab decomposition (dist.cu): this handle holds global planes [zoff, zoff + gz) of a grid with gzg planes and owns
    // local planes [zown0, zown1); the others are ghost planes.  Single-GPU handles: zoff = 0, gzg = gz, zown = [0, gz).
    int zoff, gzg, zown0, zown1;
};

// obstacle as the kernels see it (constant memory); integer raster bounds are computed on the host with the
// reference's exact fp64 expressions (macGrid.cpp:232-240, 268-270)
struct DevObstacle {
    int kind;
    int mn[3], mx[3];   // raster bounds: sphere inclusive [mn,mx], box half-open [mn,mx)
    double center[3];   // pos * cellDInv
    double r2;          // (cellDInv.x * r)^2
    double pos[3], speed[3], size[3], r;
    double pmin[3], pmax[3];  // getMinMaxRect(pos,size) -/+ particleR  (push-out box extents, simulator.cpp:281-283)
};

struct ParticleSet {
    float* pos[3];
    float* vel[3];
    float* c[9];  // APIC matrix rows c[a][b] at c[3*a+b]; allocated lazily
    uint32_t* id; // optional persistent ids (tests / facade)
};

struct PcgScalars {  // device-resident scalars of the solve (kernels read their parameters here so that the captured
                     // CUDA graph of one PCG iteration never has to be re-instantiated)
    double sigma, sigma_new, sq, rmax, rhs_sumsq, r0max;
    double loc[4];                  // slab mode: this rank's reduction results ([0],[1],[3] summed, [2] maxed over the ranks)
    int dist;                       // slab mode: reductions are finished by dist_allreduce_kernel (dist.cu)
    double scale, inv_scale, tol;   // dt/(rho h^2), its inverse, residualTolerance
    int max_it, it;                 // iteration cap, index of the iteration in flight
    int iterations, done, early_out, nan_break;
    long long fluid_cells;
};
struct PcgHostStatus {  // pinned, device-mapped: written by the device, polled by the host without a stream sync
    volatile int done;
    volatile int it_done;
};

struct MgLevel;
struct DistState;  // dist.cu
// slab mode: emigrant lists filled by the binning code of the advect kernels (particles.cu), consumed by dist.cu
struct MigDev {
    uint32_t count[2];   // particles leaving towards the lower / upper z-neighbour this step
    uint32_t overflow;   // emigrants dropped because a list was full (reported as FSIM_ERR_COMM)
    uint32_t cap;        // entries per list
    uint32_t* idx[2];    // their indices in the current particle set
    int own_lo, own_hi;  // global z-planes this rank owns: [own_lo, own_hi)
};

// kernel classes for launch counting and the optional per-launch CUDA-event profiling (fsim_profile_*)
enum KernelId {
    K_ADVECT = 0, K_BIN, K_SCAN, K_REORDER, K_P2G, K_CLASSIFY, K_FINALIZE, K_RHS, K_PCG_INIT, K_SPMV, K_UPDATE,
    K_DIRECTION, K_MG, K_MG1, K_MG2, K_APPLY, K_EXTRAP, K_G2P, K_GFX, K_MEMSET, K_PUSH, K_HALO, K_ALLREDUCE, K_MIGRATE, K_COUNT
};
struct ProfRec { int kid; cudaEvent_t e0, e1; };

struct fsim {
    int device;
    cudaStream_t stream;
    GridDims g;
    FsimGridDesc desc;
    FsimGridInfo info;
    FsimParams par;
    FsimObstacle obs[FSIM_MAX_OBS];
    int nobs;
    DevObstacle* d_obs;  // device copy (global memory, one per handle so handles never share state)
    double particle_r, zval;
    int zconst;

    // particles
    int64_t np, cap;
    int cur;  // which ParticleSet is current
    ParticleSet ps[2];
    bool have_c, track_ids;
    uint32_t next_id;         // next persistent particle id to hand out
    FsimParticleGfx* gfx;     // device buffer of the gfx export, [gfx_cap]
    int64_t gfx_cap;
    // async gfx export: two device staging buffers so that the export kernel of step k+1 never waits for the D2H copy of step k
    cudaStream_t copy_stream;
    FsimParticleGfx* gfx_async[2];
    int64_t gfx_async_cap[2];
    cudaEvent_t gfx_ready[2], gfx_copied[2];
    bool gfx_inflight[2];
    int gfx_slot;              // slot the next async export uses
    uint32_t *key, *rank;     // [cap+1] each
    uint8_t* kill;            // [cap] sink-capture flags written by the advect kernel
    bool kill_pending;        // kill[] holds flags the next sort must honour
    bool sorted;  // particles are in cell-binned order consistent with cell_start
    bool lazy_g2p;     // fsim_step defers its G2P into the next step's fused G2P + advect kernel (FSIM_NO_LAZY_G2P=1: off)
    bool g2p_pending;  // the particle velocities still lack the G2P of the last step (u / u2 hold its grid)
    bool binned;  // key / rank / cnt are current for the particle arrays (the fused advect kernel produced them)

    // grid
    uint32_t *cnt, *cell_start;  // [nc], [nc+1]
    uint32_t* scan_block;        // scan scratch
    uint8_t* flags;
    uint16_t* code;                        // stencil codes of the current solve
    void* hot;                             // one allocation: [codes | level-0 multigrid rhs], the solver's L2-persisting window
    size_t hot_bytes, hot_code_bytes;
    float* mg_b0;                          // level-0 multigrid right-hand side (inside hot)
    bool l2_persist;                       // FSIM_L2_PERSIST=0 switches the L2 persistence window on the codes off
    float *u[3], *u2[3], *wsum[3], *dens;  // u: post-P2G v / accumulators; u2: working v2
    struct alignas(64) TensorMapStorage { unsigned char b[128]; } tmap_u[3], tmap_u2[3];  // CUtensorMap of u / u2 (TMA tile staging of the G2P, g2p.cu)
    bool g2p_tma;                          // the maps are valid (FSIM_G2P_TMA=0, or a grid whose rows are not 16-byte multiples: scalar staging)
    double *p, *rhs, *r, *q, *z;           // pressure + PCG vectors (fp64)
    float* s;                              // PCG search direction (fp32 storage, see pcg.cu)
    float* mg_z32;                         // result of the last multigrid cycle (fp32, 0 outside WATER)
    bool use_mg;                           // multigrid (default) or diagonal preconditioner (FSIM_PRECOND=jacobi)
    double mg_inv_scale;                   // 1 / (dt / (rho h^2)): the hierarchy works on the integer-weight Laplacian
    std::vector<MgLevel*> mg;
    bool fx_on;                            // hybrid slab projection: in-loop exchanges fused into the solver kernels (fexch.cuh)
    bool gx_on;                            // ... including the all-rank push of the level-1 right-hand side
    const struct ArDev* ar_dev;            // ... and the all-rank reductions (rank table for ar_warp, dist_dev.cuh); nullptr: allreduce_kernel
    int mg_tail_first;                     // levels >= this run inside the single-cluster tail kernel
    int mg_tail_cluster;                   // CTAs of that cluster (0: not probed yet)
    int mg_tail2;                          // tail kernel with block-local coarse levels (mg_tail2_kernel): -1 not probed, 0 off, 1 on (opt-in: FSIM_MG_TAIL2=1; measured slower, see mg.cu)
    int mg_tail_smem;                      // shared-memory-resident tail kernel: -1 not probed, 0 off, 1 on (opt-in: FSIM_MG_TAIL_SMEM=1; measured slower, see mg.cu)
    PcgScalars* scal;                      // device
    PcgScalars* scal_host;                 // pinned (parameter upload)
    PcgScalars* result_host;               // pinned + mapped: the solve's scalars, written by publish_kernel
    PcgScalars* result_dev;                // device alias of result_host
    PcgHostStatus* status_host;            // pinned + mapped
    PcgHostStatus* status_dev;             // device alias of status_host
    cudaGraphExec_t pcg_graph;             // device-side WHILE loop around one PCG iteration (SpMV, update, multigrid cycle, direction)
    bool pcg_graph_failed;                 // conditional graph nodes unavailable: host-driven loop
    int pcg_graph_launches;                // kernels per graph launch
    int pcg_graph_class[K_COUNT];          // ... per kernel class
    bool use_graph, warm_start, warm_extrapolate;
    bool pdl;           // solver kernels are launched as programmatic dependents (launch.cuh; FSIM_PDL=0: off)
    int warm_history;   // consecutive solves whose pressure is available for the warm start
    float* p_prev;      // pressure of the step before the last one (fp32)
    double* partials;                      // reduction partials [3][max_blocks]
    unsigned int* red_counter;
    int red_blocks;
    bool pressure_valid;

    // staging
    void* stage;
    size_t stage_bytes;
    void* pinned;
    size_t pinned_bytes;

    // timing
    cudaEvent_t ev[16];
    bool ev_valid;
    bool push_timed;  // the last step ran the push-apart pass (ev[9]..ev[15] bracket it)
    FsimTimings timings;
    FsimSolveInfo solve;
    double last_step_ms;
    int64_t launches, last_step_launches;
    int sm_count;

    // per-kernel-class statistics
    uint32_t prof_mask;                 // classes whose launches are bracketed by CUDA events
    std::vector<cudaEvent_t> prof_free;
    std::vector<ProfRec> prof_recs;
    double prof_ms[K_COUNT];
    int64_t prof_n[K_COUNT];
    int64_t launch_n[K_COUNT];

    // z-slab multi-GPU state (nullptr for a single-GPU handle)
    DistState* dist;
    fsim* solver;        // slab mode, replicated projection: a full-grid solver context sharing this handle's stream (nullptr otherwise)
    bool stream_shared;  // this handle is such a context: its stream belongs to the slab handle
    bool skip_apply;     // ... and k_project leaves the pressure gradient to the slab handle
    fsim* parent;        // ... the slab handle it belongs to
    int hybrid;          // ... 1: fine level and CG vectors restricted to the planes this rank owns (halos + all-rank reductions),
                         //        coarse multigrid levels replicated; 0: the whole solve is replicated
    uint32_t slab_spawn_id;  // slab handles: next id of the spawn sequence that all ranks advance together
    uint16_t* code_full; // hybrid: stencil codes of ALL fluid cells (the coarse operators are global); h->code is owned-only
    uint16_t* code_mg;  // stencil codes the multigrid preconditioner sees (slab mode: links into ghost planes cut); == code otherwise

    // error
    mutable std::string err;
    int sticky;
};

int fsim_fail(const fsim* h, int code, const char* fmt, ...);

#define FSIM_CUDA(h, call)                                                                            \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            (h)->sticky = FSIM_ERR_CUDA;                                                              \
            return fsim_fail((h), FSIM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                                     \
        }                                                                                             \
    } while (0)

#define FSIM_CHECK_LAUNCH(h) FSIM_CUDA(h, cudaGetLastError())

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// counts the launches of a kernel class and, when that class is being profiled, brackets them with CUDA events on the
// launching stream (bench.py reads the per-class totals for the live roofline figure)
struct KScope {
    fsim* h;
    int kid;
    cudaEvent_t e0, e1;
    bool on;
    static cudaEvent_t get(fsim* h) {
        cudaEvent_t e;
        if (!h->prof_free.empty()) { e = h->prof_free.back(); h->prof_free.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    KScope(fsim* h_, int kid_, int n = 1) : h(h_), kid(kid_) {
        h->launches += n;
        h->launch_n[kid] += n;
        on = (h->prof_mask >> kid) & 1u;
        if (on) { e0 = get(h); e1 = get(h); cudaEventRecord(e0, h->stream); }
    }
    ~KScope() {
        if (on) { cudaEventRecord(e1, h->stream); h->prof_recs.push_back({kid, e0, e1}); }
    }
};

// ---- kernels' host launchers (one per stage file) ----------------------------------------------------
int k_upload_obstacles(fsim* h);
int k_advect(fsim* h, double dt, bool do_advect, bool do_pushout, bool do_stop, bool do_bin = false, bool fuse_g2p = false);
int k_sort(fsim* h);
int k_push_apart(fsim* h);  // key/count -> scan -> reorder; updates np when particles were removed
int k_p2g(fsim* h);
int k_classify(fsim* h, double dt);
int k_post_p2g_only(fsim* h, double gravity_increment);
int k_pressure_apply(fsim* h, double dt);
int k_project_basic(fsim* h, int* iterations);
int k_project(fsim* h, double dt, int* iterations);
int k_extrapolate(fsim* h);
int k_g2p(fsim* h);
int g2p_init_tma(fsim* h);
int k_export_gfx(fsim* h, FsimParticleGfx* dev_out, int64_t stride = 1);
int k_particles_aos_to_soa(fsim* h, const double* dev_aos, int64_t first, int64_t n);
int k_particles_soa_to_aos(fsim* h, double* dev_aos, int64_t first, int64_t n);
int k_particles_f32_to_soa(fsim* h, const float* pos, const float* vel, const float* c, int64_t first, int64_t n);
int k_particles_soa_to_f32(fsim* h, float* pos, float* vel, float* c, int64_t first, int64_t n);
int k_particle_cells(fsim* h, int32_t* dev_out);
int k_grid_download(fsim* h, int field, void* dev_out);
int k_grid_upload(fsim* h, int field, const void* dev_in);
int k_iota_ids(fsim* h, int64_t first, int64_t n, uint32_t base);
int k_compact_remove(fsim* h, const int32_t* dev_sorted_ids, int64_t n);
int mg_build(fsim* h);
void mg_free(fsim* h);

// ---- z-slab decomposition over peer memory (dist.cu) ----------------------------------------------------
enum { AR_RHS = 0, AR_RESIDUAL, AR_START, AR_SPMV, AR_UPDATE, AR_UPDATE_JACOBI, AR_DOTZR };
enum { HALO_P2G = 0, HALO_P, HALO_S, HALO_U2, HALO_U2_FLAGS, HALO_BASIC_ODD, HALO_BASIC_EVEN };
int dist_halo(fsim* h, int what, bool in_pcg_loop);      // ghost-plane exchange with both z-neighbours (pull over peer memory)
int dist_allreduce(fsim* h, int kind, bool in_pcg_loop);  // finishes a PCG reduction across the ranks
int dist_migrate(fsim* h);
int dist_push_apart_ghosts(fsim* h);  // stage my boundary planes' particle positions, pull the neighbours' (push-apart on slabs)
bool dist_pa_ghost(const fsim* h, int side, const uint32_t** starts, const float** x, const float** y, const float** z);
int dist_gather_solver_inputs(fsim* h);
// hybrid projection (solver context hs of a slab handle): arrays live at global plane indices on every rank
enum { SYM_S = 0, SYM_P, SYM_X };
int dist_halo_sym(fsim* hs, int which, const void* ptr, bool in_pcg_loop);  // planes own_lo-1 / own_hi <- the neighbours' same planes
int dist_gather_coarse(fsim* hs, bool in_pcg_loop);                          // level-1 right-hand side: every rank's coarse planes
int dist_register_solver(fsim* h);                                           // publishes the solver context's exchanged arrays
int mg_alloc(fsim* h);
float* mg_level_array(fsim* h, int level, int which, float** base, size_t* pad);  // which: 0 xa, 1 xb, 2 b                  // all ranks' owned planes of flags / v2 / avgPNum -> the full-grid solver context                                // emigrants -> neighbours, immigrants appended + binned
void dist_free(fsim* h);
void dist_partition(int gzg, int rank, int nranks, int* own_lo, int* own_hi, int* zoff, int* gz_local);
int dist_init(fsim* h, int rank, int nranks, int own_lo, int own_hi);
int dist_export(fsim* h, FsimDistExport* out);
int dist_connect(fsim* h, const FsimDistExport* all, int n);
void dist_rank(const fsim* h, int* rank, int* nranks);
bool dist_peer_in_process(const fsim* h);  // another rank of the group lives in this process
int dist_wait_stats(fsim* h, FsimDistWaitStats* out, int reset);
struct FxPush;
struct FxWait;
struct ArDev;
const ArDev* dist_ar_dev(const fsim* hs);  // rank table of the all-rank reduction fused into the reducing kernels (dist_dev.cuh), or nullptr
struct FxAllPush;
bool dist_gx(fsim* hs, int cgx, int cgy, int cgz, FxAllPush* p, FxWait* w);  // fused all-rank push of the level-1 right-hand side
bool dist_fx(fsim* hs, int which, const void* ptr, FxPush* p, FxWait* w);  // fused in-loop exchange of a level-0 array (fexch.cuh)
int dist_check(fsim* h);  // FSIM_ERR_COMM once an exchange timed out / a migration list overflowed
int fsim_ensure_capacity(fsim* h, int64_t n);  // fsim_api.cu
MigDev* dist_mig_dev(const fsim* h);
const uint32_t* dist_nsrc_dev(fsim* h);                   // device count of source particles for the reorder (locals + immigrants); valid once after dist_migrate
int64_t dist_mig_capacity(const fsim* h);
