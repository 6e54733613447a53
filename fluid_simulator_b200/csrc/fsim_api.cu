// fsim_api.cu -- the C ABI of libfsim_b200 (include/fsim.h): handle life cycle, parameter / obstacle hand-off,
// particle and grid transfer in the reference's host layouts, Simulator::simulate orchestration and stage timing.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fsim_internal.h"

static thread_local std::string g_create_error;

int fsim_fail(const fsim* h, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

namespace {

constexpr size_t STAGE_BYTES = size_t(192) << 20;  // device staging buffer for layout conversion (chunked transfers)

template <typename T>
int dev_alloc(fsim* h, T** p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    FSIM_CUDA(h, cudaMalloc((void**)p, n * sizeof(T)));
    FSIM_CUDA(h, cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
    return FSIM_OK;
}

int bind(const fsim* h) {
    if (!h) return FSIM_ERR_INVALID;
    if (h->sticky) return fsim_fail(h, h->sticky, "handle is in a failed state: %s", h->err.c_str());
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fsim_fail(h, FSIM_ERR_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
    return FSIM_OK;
}
#define BIND(h)                      \
    do {                             \
        int rc__ = bind(h);          \
        if (rc__) return rc__;       \
    } while (0)
#define TRY(x)                       \
    do {                             \
        int rc__ = (x);              \
        if (rc__) return rc__;       \
    } while (0)

int alloc_particle_set(fsim* h, ParticleSet& s, int64_t cap, bool with_c, bool with_id) {
    for (int c = 0; c < 3; c++) { TRY(dev_alloc(h, &s.pos[c], cap)); TRY(dev_alloc(h, &s.vel[c], cap)); }
    for (int c = 0; c < 9; c++) { s.c[c] = nullptr; if (with_c) TRY(dev_alloc(h, &s.c[c], cap)); }
    s.id = nullptr;
    if (with_id) TRY(dev_alloc(h, &s.id, cap));
    return FSIM_OK;
}
void free_particle_set(ParticleSet& s) {
    for (int c = 0; c < 3; c++) { cudaFree(s.pos[c]); cudaFree(s.vel[c]); s.pos[c] = s.vel[c] = nullptr; }
    for (int c = 0; c < 9; c++) { cudaFree(s.c[c]); s.c[c] = nullptr; }
    cudaFree(s.id); s.id = nullptr;
}

int ensure_capacity(fsim* h, int64_t n) {
    if (h->dist) n += 2 * dist_mig_capacity(h);  // slab mode: immigrants are appended behind the locals before the sort
    if (n <= h->cap) return FSIM_OK;
    int64_t ncap = h->cap > 0 ? h->cap : 1024;
    while (ncap < n) ncap += ncap / 2 + 1024;
    ParticleSet ns[2];
    for (int k = 0; k < 2; k++) TRY(alloc_particle_set(h, ns[k], ncap, h->have_c, h->track_ids));
    const ParticleSet& o = h->ps[h->cur];
    if (h->np > 0) {
        for (int c = 0; c < 3; c++) {
            FSIM_CUDA(h, cudaMemcpyAsync(ns[0].pos[c], o.pos[c], sizeof(float) * h->np, cudaMemcpyDeviceToDevice, h->stream));
            FSIM_CUDA(h, cudaMemcpyAsync(ns[0].vel[c], o.vel[c], sizeof(float) * h->np, cudaMemcpyDeviceToDevice, h->stream));
        }
        if (h->have_c)
            for (int c = 0; c < 9; c++)
                FSIM_CUDA(h, cudaMemcpyAsync(ns[0].c[c], o.c[c], sizeof(float) * h->np, cudaMemcpyDeviceToDevice, h->stream));
        if (h->track_ids)
            FSIM_CUDA(h, cudaMemcpyAsync(ns[0].id, o.id, sizeof(uint32_t) * h->np, cudaMemcpyDeviceToDevice, h->stream));
    }
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    free_particle_set(h->ps[0]);
    free_particle_set(h->ps[1]);
    h->ps[0] = ns[0];
    h->ps[1] = ns[1];
    h->cur = 0;
    cudaFree(h->key); cudaFree(h->rank); cudaFree(h->kill);
    TRY(dev_alloc(h, &h->key, ncap + 1));
    TRY(dev_alloc(h, &h->rank, ncap + 1));
    TRY(dev_alloc(h, &h->kill, ncap + 1));
    // the particle-side scan scratch must cover ceil(cap / tile) blocks as well as the grid's
    h->cap = ncap;
    return FSIM_OK;
}

}  // namespace
int fsim_ensure_capacity(fsim* h, int64_t n) { return ensure_capacity(h, n); }
namespace {

int ensure_c(fsim* h) {  // APIC affine matrices are allocated the first time they are needed
    if (h->have_c) return FSIM_OK;
    for (int k = 0; k < 2; k++)
        for (int c = 0; c < 9; c++) TRY(dev_alloc(h, &h->ps[k].c[c], h->cap));
    h->have_c = true;
    return FSIM_OK;
}

// host-side mirror of MacGrid::getMinMaxRect (macGrid.h:139-145)
void min_max_rect(const FsimGridInfo& gi, const double* pos, const double* size, double* lo, double* hi) {
    for (int a = 0; a < 3; a++) {
        const double center = pos[a] * gi.cell_d_inv[a];
        const double d = size[a] * gi.cell_d_inv[a] * 0.5;
        const int mn = (int)std::fmax(std::round(center - d), 1.0);
        const int mx = (int)std::fmin(std::round(center + d), gi.grid_size[a] - 1.0);
        lo[a] = (double)mn * gi.cell_d[a];
        hi[a] = (double)mx * gi.cell_d[a];
    }
}

void fold_timings(fsim* h) {
    if (!h->ev_valid) return;
    cudaEventSynchronize(h->ev[8]);
    float ms[8];
    for (int i = 0; i < 8; i++) cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]);
    // ev: 0 start | 1 advect(+spawn,+pushout) | 2 sort | 3 p2g | 4 classify | 5 project | 6 extrapolate | 7 g2p | 8 end
    const double us_adv = ms[0] * 1e3, us_sort = ms[1] * 1e3, us_p2g = ms[2] * 1e3, us_prep = ms[3] * 1e3,
                 us_proj = ms[4] * 1e3, us_ext = ms[5] * 1e3, us_g2p = ms[6] * 1e3;
    FsimTimings& t = h->timings;
    const double f = 0.9;  // slidingAvgFactor, simulator.cpp:52
    // with push-apart on, ev[0]..ev[1] spans advect | sort + push-apart (ev[9]..ev[15]) | push-out pass (ev[15]..ev[1]): each goes
    // to the reference's own key (simulator.cpp:57-68); without it advect and push-out are one kernel, reported as SimulateParticles
    double us_push = 0, us_pushout = 0, us_sim = us_adv;
    if (h->push_timed) {
        float pm = 0, po = 0, pa = 0;
        cudaEventElapsedTime(&pa, h->ev[0], h->ev[9]);
        cudaEventElapsedTime(&pm, h->ev[9], h->ev[15]);
        cudaEventElapsedTime(&po, h->ev[15], h->ev[1]);
        us_sim = pa * 1e3; us_push = pm * 1e3; us_pushout = po * 1e3;
    }
    t.simulate_particles = (int64_t)(t.simulate_particles * f + us_sim * (1 - f));
    t.push_particles_apart = (int64_t)(t.push_particles_apart * f + us_push * (1 - f));
    t.push_particles_out_of_obstacles = (int64_t)(t.push_particles_out_of_obstacles * f + us_pushout * (1 - f));
    t.p2g_transfer = (int64_t)(t.p2g_transfer * f + (us_sort + us_p2g) * (1 - f));
    t.incompressibility_prep = (int64_t)(t.incompressibility_prep * f + us_prep * (1 - f));
    t.incompressibility = (int64_t)(t.incompressibility * f + us_proj * (1 - f));
    t.velocity_extrapolation = (int64_t)(t.velocity_extrapolation * f + us_ext * (1 - f));
    t.g2p_transfer = (int64_t)(t.g2p_transfer * f + us_g2p * (1 - f));
    t.incompressibility_it_count = h->solve.iterations;
    t.last_raw_us[0] = us_sim; t.last_raw_us[1] = us_push; t.last_raw_us[2] = us_pushout; t.last_raw_us[3] = us_p2g;
    t.last_raw_us[4] = us_prep; t.last_raw_us[5] = us_proj; t.last_raw_us[6] = us_ext; t.last_raw_us[7] = us_g2p;
    t.last_sort_us = us_sort;
    float total;
    cudaEventElapsedTime(&total, h->ev[0], h->ev[8]);
    h->last_step_ms = total;
    h->ev_valid = false;
}

int ensure_sorted(fsim* h) { return h->sorted ? FSIM_OK : k_sort(h); }

// fsim_step leaves its G2P to the next step's fused kernel; everything else that reads or writes particle velocities, or
// writes the grid velocities, first brings the particles up to date
int flush_g2p(fsim* h) { return h->g2p_pending ? k_g2p(h) : FSIM_OK; }
#define BIND_FLUSH(h)  \
    do {               \
        BIND(h);       \
        TRY(flush_g2p(h)); \
    } while (0)

int ensure_gfx(fsim* h) {
    if (h->gfx && h->gfx_cap >= h->np) return FSIM_OK;
    cudaFree(h->gfx);
    h->gfx = nullptr;
    h->gfx_cap = h->cap;
    FSIM_CUDA(h, cudaMalloc((void**)&h->gfx, sizeof(FsimParticleGfx) * (size_t)(h->gfx_cap > 0 ? h->gfx_cap : 1)));
    return FSIM_OK;
}

}  // namespace

int k_upload_obstacles(fsim* h) {
    std::vector<DevObstacle> d(h->nobs > 0 ? h->nobs : 1);
    const FsimGridInfo& gi = h->info;
    for (int k = 0; k < h->nobs; k++) {
        const FsimObstacle& o = h->obs[k];
        DevObstacle& D = d[k];
        memset(&D, 0, sizeof(D));
        D.kind = o.kind;
        D.r = o.r;
        for (int a = 0; a < 3; a++) {
            D.pos[a] = o.pos[a]; D.speed[a] = o.speed[a]; D.size[a] = o.size[a];
            D.center[a] = o.pos[a] * gi.cell_d_inv[a];  // macGrid.cpp:232
        }
        if (o.kind == FSIM_OBSTACLE_BOX) {
            for (int a = 0; a < 3; a++) {  // macGrid.cpp:268-270
                const double dd = o.size[a] * gi.cell_d_inv[a] * 0.5;
                D.mn[a] = (int)std::fmax(std::round(D.center[a] - dd), 1.0);
                D.mx[a] = (int)std::fmin(std::round(D.center[a] + dd), gi.grid_size[a] - 1.0);
            }
            double lo[3], hi[3];
            min_max_rect(gi, o.pos, o.size, lo, hi);
            for (int a = 0; a < 3; a++) { D.pmin[a] = lo[a] - h->particle_r; D.pmax[a] = hi[a] + h->particle_r; }
        } else {
            const double r = gi.cell_d_inv[0] * o.r;  // macGrid.cpp:237-240
            D.r2 = r * r;
            for (int a = 0; a < 3; a++) {
                D.mn[a] = (int)std::fmax(D.center[a] - r, 1.0);
                D.mx[a] = (int)std::fmin(D.center[a] + r, gi.grid_size[a] - 2.0);
            }
        }
    }
    FSIM_CUDA(h, cudaMemcpyAsync(h->d_obs, d.data(), sizeof(DevObstacle) * d.size(), cudaMemcpyHostToDevice, h->stream));
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    return FSIM_OK;
}

extern "C" {

int fsim_abi_version(void) { return FSIM_ABI_VERSION; }

const char* fsim_last_error(const fsim_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int create_impl(const FsimGridDesc* desc, int rank, int nranks, fsim_t** out) {
    if (!desc || !out) return fsim_fail(nullptr, FSIM_ERR_INVALID, "null argument");
    if (nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) return fsim_fail(nullptr, FSIM_ERR_INVALID, "bad rank %d of %d", rank, nranks);
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fsim_fail(nullptr, FSIM_ERR_CUDA, "no CUDA device available (%s); libfsim_b200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (desc->device < 0 || desc->device >= ndev) return fsim_fail(nullptr, FSIM_ERR_INVALID, "bad device ordinal %d", desc->device);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, desc->device)) != cudaSuccess)
        return fsim_fail(nullptr, FSIM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fsim_fail(nullptr, FSIM_ERR_CUDA, "device %d is sm_%d%d; libfsim_b200 is built for sm_100a only", desc->device,
                         prop.major, prop.minor);
    if (!(desc->resolution > 0) || !(desc->particle_radius > 0)) return fsim_fail(nullptr, FSIM_ERR_INVALID, "bad resolution / radius");

    fsim* h = new fsim();
    h->device = desc->device;
    h->desc = *desc;
    h->sticky = 0;
    cudaSetDevice(h->device);
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return fsim_fail(nullptr, FSIM_ERR_CUDA, "cudaStreamCreate failed");
    }
    // MacGrid ctor (macGrid.cpp:9-13); BridsonSolverGrid narrows the resolution to float (bridsonSolverGrid.cpp:8)
    const double res = (double)(float)desc->resolution;
    FsimGridInfo& gi = h->info;
    const int twoD = desc->two_d != 0;
    gi.cell_d[0] = 1 / res; gi.cell_d[1] = 1 / res; gi.cell_d[2] = twoD ? desc->target_dims[2] / 3 : 1 / res;
    for (int a = 0; a < 3; a++) gi.cell_d_inv[a] = 1.0 / gi.cell_d[a];
    gi.grid_size[0] = (int)(desc->target_dims[0] / gi.cell_d[0]);
    gi.grid_size[1] = (int)(desc->target_dims[1] / gi.cell_d[1]);
    gi.grid_size[2] = twoD ? 3 : (int)(desc->target_dims[2] / gi.cell_d[2]);
    gi.dimensions[0] = gi.grid_size[0] * gi.cell_d[0];
    gi.dimensions[1] = gi.grid_size[1] * gi.cell_d[1];
    gi.dimensions[2] = twoD ? desc->target_dims[2] : gi.grid_size[2] * gi.cell_d[2];
    gi.two_d = twoD;
    gi.cell_count = (int64_t)gi.grid_size[0] * gi.grid_size[1] * gi.grid_size[2];
    if (gi.grid_size[0] < 3 || gi.grid_size[1] < 3 || gi.grid_size[2] < 3 || gi.cell_count >= (int64_t(1) << 31)) {
        delete h;
        return fsim_fail(nullptr, FSIM_ERR_INVALID, "grid %dx%dx%d out of range", gi.grid_size[0], gi.grid_size[1], gi.grid_size[2]);
    }
    GridDims& g = h->g;
    g.gx = gi.grid_size[0]; g.gy = gi.grid_size[1]; g.gz = gi.grid_size[2];
    g.sy = g.gx; g.sz = g.gx * g.gy; g.nc = gi.cell_count;
    g.zoff = 0; g.gzg = g.gz; g.zown0 = 0; g.zown1 = g.gz;
    int own_lo = 0, own_hi = g.gz;
    if (nranks > 1) {  // z-slab: this handle stores the planes it owns + one ghost plane towards each neighbour
        if (twoD || g.gzg < 2 * nranks) {
            delete h;
            return fsim_fail(nullptr, FSIM_ERR_INVALID, "cannot cut %d z-planes%s into %d slabs", g.gzg, twoD ? " (2D)" : "", nranks);
        }
        int zoff, gzl;
        dist_partition(g.gzg, rank, nranks, &own_lo, &own_hi, &zoff, &gzl);
        g.zoff = zoff; g.gz = gzl; g.zown0 = own_lo - zoff; g.zown1 = own_hi - zoff;
        g.nc = (int64_t)g.gx * g.gy * g.gz;
    }
    g.dhx = gi.cell_d[0]; g.dhy = gi.cell_d[1]; g.dhz = gi.cell_d[2];
    g.dihx = gi.cell_d_inv[0]; g.dihy = gi.cell_d_inv[1]; g.dihz = gi.cell_d_inv[2];
    g.hx = (float)g.dhx; g.hy = (float)g.dhy; g.hz = (float)g.dhz;
    g.ihx = (float)g.dihx; g.ihy = (float)g.dihy; g.ihz = (float)g.dihz;
    g.twoD = twoD;
    h->particle_r = desc->particle_radius;
    h->zconst = twoD;
    h->zval = desc->target_dims[2] / 2;  // HashedParticles z = dimensions.z / 2 (simulationManager.cpp:27)
    // defaults: SimulatorConfig (simulator.h:30-39) + MacGrid params (macGrid.h:173-179)
    FsimParams& p = h->par;
    memset(&p, 0, sizeof(p));
    p.transfer_type = FSIM_TRANSFER_FLIP; p.flip_ratio = 0.99f; p.gravity = 150.0f; p.gravity_enabled = 1;
    p.push_apart_enabled = 1; p.pressure_enabled = 1; p.max_iterations = 80; p.pressure_k = 2.0;
    p.average_pressure = 2.0; p.fluid_density = 1.0; p.residual_tolerance = 1e-6;
    if (nranks > 1) p.push_apart_enabled = 0;  // (slab handles: off unless asked for -- it costs a second migration per step)
    h->nobs = 0;
    h->np = 0; h->cap = 0; h->cur = 0; h->have_c = false; h->track_ids = true; h->sorted = false; h->binned = false; h->kill_pending = false;
    memset(h->ps, 0, sizeof(h->ps));
    h->key = h->rank = nullptr; h->kill = nullptr;
    h->next_id = 0; h->gfx = nullptr; h->gfx_cap = 0;
    h->copy_stream = nullptr; h->gfx_slot = 0;
    { const char* e = getenv("FSIM_NO_LAZY_G2P"); h->lazy_g2p = !(e && e[0] == '1'); h->g2p_pending = false; }
    for (int k = 0; k < 2; k++) { h->gfx_async[k] = nullptr; h->gfx_async_cap[k] = 0; h->gfx_ready[k] = h->gfx_copied[k] = nullptr; h->gfx_inflight[k] = false; }
    h->pressure_valid = false; h->ev_valid = false;
    memset(&h->timings, 0, sizeof(h->timings));
    memset(&h->solve, 0, sizeof(h->solve));
    h->launches = h->last_step_launches = 0; h->last_step_ms = 0;
    h->prof_mask = 0;
    h->mg_z32 = nullptr; h->mg_inv_scale = 1.0;
    h->pcg_graph = nullptr; h->pcg_graph_failed = false; h->pcg_graph_launches = 0; memset(h->pcg_graph_class, 0, sizeof(h->pcg_graph_class));
    { const char* e = getenv("FSIM_NO_GRAPH"); h->use_graph = !(e && e[0] == '1'); }
    { const char* e = getenv("FSIM_NO_WARM_START"); h->warm_start = !(e && e[0] == '1'); }
    { const char* e = getenv("FSIM_PDL"); h->pdl = !(e && e[0] == '0'); }
    { const char* e = getenv("FSIM_WARM_EXTRAPOLATE"); h->warm_extrapolate = !(e && e[0] == '0'); }
    h->warm_history = 0; h->p_prev = nullptr; h->mg_tail_cluster = 0; h->mg_tail_smem = -1; h->mg_tail2 = -1; h->fx_on = false; h->gx_on = false; h->ar_dev = nullptr;
    h->status_host = nullptr; h->status_dev = nullptr;
    h->dist = nullptr; h->code_mg = nullptr; h->solver = nullptr; h->stream_shared = false; h->skip_apply = false;
    h->parent = nullptr; h->hybrid = 0; h->code_full = nullptr; h->slab_spawn_id = 0x80000000u;
    { const char* e = getenv("FSIM_PRECOND"); h->use_mg = !(e && strcmp(e, "jacobi") == 0); }
    memset(h->prof_ms, 0, sizeof(h->prof_ms)); memset(h->prof_n, 0, sizeof(h->prof_n)); memset(h->launch_n, 0, sizeof(h->launch_n));

    int rc = FSIM_OK;
    auto A = [&](int r) { if (!rc) rc = r; };
    A(dev_alloc(h, &h->cnt, g.nc));
    A(dev_alloc(h, &h->cell_start, g.nc + 1));
    A(dev_alloc(h, &h->flags, g.nc));
    {   // stencil codes and the level-0 multigrid right-hand side share one allocation: the range the solver pins in L2
        const size_t code_bytes = (((size_t)(g.nc + 8) * sizeof(uint16_t)) + 255) & ~(size_t)255;
        h->hot_code_bytes = code_bytes;
        h->hot_bytes = code_bytes + (size_t)g.nc * sizeof(float);
        h->hot = nullptr;
        if (!rc && cudaMalloc(&h->hot, h->hot_bytes) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "solver buffer alloc failed");
        if (!rc) cudaMemsetAsync(h->hot, 0, h->hot_bytes, h->stream);
        h->code = (uint16_t*)h->hot;
        h->mg_b0 = h->hot ? (float*)((char*)h->hot + code_bytes) : nullptr;
        // Every level-0 solver kernel starts with a dependent load of the stencil codes (2 B/cell, re-read by 8 kernels per PCG
        // iteration): an L2 persistence window keeps them resident (measured -0.24 ms/step at 256^3: the first round trip of
        // those kernels becomes an L2 hit).  The carve-out is taken from every other kernel's L2, so it is limited to 40 MB
        // (pinning the cycle's rhs as well, with the maximum carve-out, made the particle kernels 2x slower).
        { const char* e = getenv("FSIM_L2_PERSIST"); h->l2_persist = !(e && e[0] == '0'); }
        const size_t cap = std::min((size_t)40 << 20, (size_t)prop.persistingL2CacheMaxSize);
        if (h->l2_persist && !rc && cap > 0 && code_bytes <= (size_t)prop.accessPolicyMaxWindowSize) {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(code_bytes, cap));
            cudaStreamAttrValue v;
            memset(&v, 0, sizeof(v));
            v.accessPolicyWindow.base_ptr = h->code;
            v.accessPolicyWindow.num_bytes = code_bytes;
            v.accessPolicyWindow.hitRatio = code_bytes <= cap ? 1.0f : (float)((double)cap / (double)code_bytes);
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &v);
        }
        cudaGetLastError();
    }
    if (nranks > 1) A(dev_alloc(h, &h->code_mg, g.nc + 8)); else h->code_mg = h->code;
    for (int a = 0; a < 3; a++) { A(dev_alloc(h, &h->u[a], g.nc)); A(dev_alloc(h, &h->u2[a], g.nc)); A(dev_alloc(h, &h->wsum[a], g.nc)); }
    A(dev_alloc(h, &h->dens, g.nc));
    h->g2p_tma = false;
    if (!rc) A(g2p_init_tma(h));
    A(dev_alloc(h, &h->p, g.nc)); A(dev_alloc(h, &h->rhs, g.nc)); A(dev_alloc(h, &h->r, g.nc));
    A(dev_alloc(h, &h->s, g.nc)); A(dev_alloc(h, &h->q, g.nc)); A(dev_alloc(h, &h->z, g.nc));
    A(dev_alloc(h, &h->p_prev, g.nc));
    A(dev_alloc(h, &h->d_obs, FSIM_MAX_OBS));
    A(dev_alloc(h, &h->scal, 1));
    h->red_blocks = (int)(g.nc / 256 + 2);  // one partial per 256-thread block of the widest solver launch
    A(dev_alloc(h, &h->partials, (size_t)3 * (h->red_blocks + h->red_blocks / 32 + 2)));
    A(dev_alloc(h, &h->red_counter, (size_t)h->red_blocks / 32 + 4));
    const int64_t cap0 = desc->particle_capacity > 0 ? desc->particle_capacity : 0;
    // scan scratch: one entry per 4096-item tile of the larger of (cells, particle capacity); grown with capacity
    h->scan_block = nullptr;
    A(dev_alloc(h, &h->scan_block, (size_t)((std::max<int64_t>(g.nc, int64_t(1) << 31)) / 4096 + 2)));
    h->stage = nullptr; h->stage_bytes = STAGE_BYTES;
    if (!rc && cudaMalloc(&h->stage, h->stage_bytes) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "staging alloc failed");
    h->scal_host = nullptr; h->result_host = nullptr; h->result_dev = nullptr;
    if (!rc && cudaMallocHost((void**)&h->scal_host, sizeof(PcgScalars)) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "pinned alloc failed");
    if (!rc && (cudaHostAlloc((void**)&h->status_host, sizeof(PcgHostStatus), cudaHostAllocMapped) != cudaSuccess ||
                cudaHostGetDevicePointer((void**)&h->status_dev, (void*)h->status_host, 0) != cudaSuccess))
        rc = fsim_fail(h, FSIM_ERR_CUDA, "mapped status alloc failed");
    if (!rc && (cudaHostAlloc((void**)&h->result_host, sizeof(PcgScalars), cudaHostAllocMapped) != cudaSuccess ||
                cudaHostGetDevicePointer((void**)&h->result_dev, (void*)h->result_host, 0) != cudaSuccess))
        rc = fsim_fail(h, FSIM_ERR_CUDA, "mapped result alloc failed");
    if (!rc) { h->status_host->done = 0; h->status_host->it_done = 0; }
    for (int i = 0; i < 16 && !rc; i++)
        if (cudaEventCreate(&h->ev[i]) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "event create failed");
    if (!rc && nranks > 1) rc = dist_init(h, rank, nranks, own_lo, own_hi);
    if (!rc && nranks > 1) {
        // Projection of a slab handle: by default every rank solves the FULL pressure system redundantly on a
        // full-grid context fed by an all-rank gather of flags / v2 / avgPNum -- same multigrid, same iteration counts and the
        // same pressure as the single-GPU path.  FSIM_SLAB_SOLVER=distributed keeps the solve on the slab instead (halo of the
        // search direction, all-rank reductions, block-local multigrid: 3-4x the iterations, measured).
        const char* e = getenv("FSIM_SLAB_SOLVER");
        const bool hybrid = !(e && e[0] == 'r');  // default: hybrid
        if (!(e && e[0] == 'd')) {
            FsimGridDesc d2 = *desc;
            d2.particle_capacity = 0;
            fsim* hs = nullptr;
            rc = create_impl(&d2, 0, 1, &hs);
            if (rc) h->err = g_create_error;
            else {
                cudaSetDevice(h->device);
                cudaStreamSynchronize(hs->stream);
                cudaStreamDestroy(hs->stream);
                hs->stream = h->stream; hs->stream_shared = true; hs->skip_apply = true;
                hs->parent = h;
                h->solver = hs;
                if (hybrid) {
                    // HYBRID projection: the CG vectors and the fine multigrid level only work on the planes this rank owns (ghost
                    // planes of the search direction / iterate pulled from the neighbours, reductions over all ranks); the coarse
                    // levels run replicated on the all-rank gather of the level-1 right-hand side.  Same operator, same cycle, same
                    // iteration counts as one GPU.  FSIM_SLAB_SOLVER=replicated: every rank solves the whole system instead.
                    hs->hybrid = 1;
                    hs->g.zown0 = own_lo; hs->g.zown1 = own_hi;
                    if (cudaMalloc((void**)&hs->code_full, (size_t)(hs->g.nc + 8) * sizeof(uint16_t)) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "solver code alloc failed");
                    if (!rc) { cudaMemsetAsync(hs->code_full, 0, (size_t)(hs->g.nc + 8) * sizeof(uint16_t), h->stream); rc = mg_alloc(hs); if (rc) h->err = hs->err; }
                    if (!rc) rc = dist_register_solver(h);
                }
                // the L2 persistence window of this stream follows the codes the solver actually reads
                const size_t capb = std::min((size_t)40 << 20, (size_t)prop.persistingL2CacheMaxSize);
                if (hs->l2_persist && capb > 0 && hs->hot_code_bytes <= (size_t)prop.accessPolicyMaxWindowSize) {
                    cudaStreamAttrValue v;
                    memset(&v, 0, sizeof(v));
                    v.accessPolicyWindow.base_ptr = hs->code;
                    v.accessPolicyWindow.num_bytes = hs->hot_code_bytes;
                    v.accessPolicyWindow.hitRatio = hs->hot_code_bytes <= capb ? 1.0f : (float)((double)capb / (double)hs->hot_code_bytes);
                    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &v);
                    cudaGetLastError();
                }
            }
        }
    }
    if (!rc && cap0 > 0) rc = ensure_capacity(h, cap0);
    if (!rc) {
        // MacGrid ctor leaves every cell AIR with the solid border shell (macGrid.cpp:15-16)
        rc = k_upload_obstacles(h);
        if (!rc) rc = k_sort(h);  // cnt = 0, cell_start = 0
        if (!rc) rc = k_classify(h, 0.0);
    }
    if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "init sync failed");
    if (rc) {
        g_create_error = h->err;
        fsim_destroy(h);
        return rc;
    }
    *out = h;
    return FSIM_OK;
}

int fsim_create(const FsimGridDesc* desc, fsim_t** out) { return create_impl(desc, 0, 1, out); }
int fsim_create_slab(const FsimGridDesc* desc, int rank, int nranks, fsim_t** out) { return create_impl(desc, rank, nranks, out); }

int fsim_slab_partition(int global_gz, int rank, int nranks, FsimSlabInfo* out) {
    if (!out || nranks < 1 || rank < 0 || rank >= nranks || global_gz < 1) return FSIM_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    int lo, hi, zoff, gzl;
    dist_partition(global_gz, rank, nranks, &lo, &hi, &zoff, &gzl);
    if (nranks == 1) { lo = 0; hi = global_gz; zoff = 0; gzl = global_gz; }
    out->rank = rank; out->nranks = nranks; out->z_offset = zoff; out->gz_local = gzl; out->own_lo = lo; out->own_hi = hi;
    return FSIM_OK;
}

int fsim_get_slab_info(const fsim_t* h, FsimSlabInfo* out) {
    if (!h || !out) return FSIM_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    out->rank = 0; out->nranks = 1;
    dist_rank(h, &out->rank, &out->nranks);
    out->z_offset = h->g.zoff; out->gz_local = h->g.gz; out->own_lo = h->g.zoff + h->g.zown0; out->own_hi = h->g.zoff + h->g.zown1;
    return FSIM_OK;
}

int fsim_dist_export(fsim_t* h, FsimDistExport* out) {
    BIND(h);
    if (!out) return fsim_fail(h, FSIM_ERR_INVALID, "null export");
    if (!h->dist) return fsim_fail(h, FSIM_ERR_INVALID, "not a slab handle (fsim_create_slab with nranks > 1)");
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    return dist_export(h, out);
}

int fsim_dist_connect(fsim_t* h, const FsimDistExport* all, int n) {
    BIND(h);
    if (!all) return fsim_fail(h, FSIM_ERR_INVALID, "null exports");
    if (!h->dist) return fsim_fail(h, FSIM_ERR_INVALID, "not a slab handle (fsim_create_slab with nranks > 1)");
    return dist_connect(h, all, n);
}

int fsim_dist_wait_stats(fsim_t* h, FsimDistWaitStats* out, int reset) {
    BIND(h);
    if (!out) return fsim_fail(h, FSIM_ERR_INVALID, "null stats");
    memset(out, 0, sizeof(*out));
    return dist_wait_stats(h, out, reset);
}

int fsim_destroy(fsim_t* h) {
    if (!h) return FSIM_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->solver) { fsim_destroy(h->solver); h->solver = nullptr; }
    dist_free(h);
    if (h->code_mg && h->code_mg != h->code) cudaFree(h->code_mg);
    cudaFree(h->code_full);
    mg_free(h);
    free_particle_set(h->ps[0]);
    free_particle_set(h->ps[1]);
    cudaFree(h->key); cudaFree(h->rank); cudaFree(h->kill);
    cudaFree(h->cnt); cudaFree(h->cell_start); cudaFree(h->scan_block); cudaFree(h->flags); cudaFree(h->hot);
    for (int a = 0; a < 3; a++) { cudaFree(h->u[a]); cudaFree(h->u2[a]); cudaFree(h->wsum[a]); }
    cudaFree(h->dens);
    cudaFree(h->p); cudaFree(h->rhs); cudaFree(h->r); cudaFree(h->s); cudaFree(h->q); cudaFree(h->z); cudaFree(h->p_prev);
    cudaFree(h->d_obs); cudaFree(h->scal); cudaFree(h->partials); cudaFree(h->red_counter);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    for (int k = 0; k < 2; k++) {
        if (h->gfx_ready[k]) cudaEventDestroy(h->gfx_ready[k]);
        if (h->gfx_copied[k]) cudaEventDestroy(h->gfx_copied[k]);
        cudaFree(h->gfx_async[k]);
    }
    cudaFree(h->stage); cudaFree(h->gfx);
    for (const ProfRec& r : h->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (cudaEvent_t e : h->prof_free) cudaEventDestroy(e);
    if (h->scal_host) cudaFreeHost(h->scal_host);
    if (h->result_host) cudaFreeHost(h->result_host);
    if (h->status_host) cudaFreeHost((void*)h->status_host);
    if (h->pcg_graph) cudaGraphExecDestroy(h->pcg_graph);
    for (int i = 0; i < 16; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (!h->stream_shared) cudaStreamDestroy(h->stream);
    delete h;
    return FSIM_OK;
}

int fsim_get_grid_info(const fsim_t* h, FsimGridInfo* out) {
    if (!h || !out) return FSIM_ERR_INVALID;
    *out = h->info;
    return FSIM_OK;
}

int fsim_set_params(fsim_t* h, const FsimParams* p) {
    BIND(h);
    if (!p) return fsim_fail(h, FSIM_ERR_INVALID, "null params");
    if (p->transfer_type < 0 || p->transfer_type > 2) return fsim_fail(h, FSIM_ERR_INVALID, "bad transfer type %d", p->transfer_type);
    if (!(p->fluid_density > 0)) return fsim_fail(h, FSIM_ERR_INVALID, "fluid density must be > 0");
    if (p->solver_type != FSIM_SOLVER_BRIDSON && p->solver_type != FSIM_SOLVER_BASIC) return fsim_fail(h, FSIM_ERR_INVALID, "bad solver type %d", p->solver_type);
    if (p->transfer_type != h->par.transfer_type || p->flip_ratio != h->par.flip_ratio) TRY(flush_g2p(h));  // a pending G2P uses the old blend
    h->par = *p;
    if (p->transfer_type == FSIM_TRANSFER_APIC) TRY(ensure_c(h));
    return FSIM_OK;
}

int fsim_set_obstacles(fsim_t* h, const FsimObstacle* obs, int n) {
    BIND(h);
    if (n < 0 || n > FSIM_MAX_OBS) return fsim_fail(h, FSIM_ERR_INVALID, "obstacle count %d exceeds %d", n, FSIM_MAX_OBS);
    if (n > 0 && !obs) return fsim_fail(h, FSIM_ERR_INVALID, "null obstacles");
    for (int i = 0; i < n; i++)
        if (obs[i].kind < 0 || obs[i].kind > 3) return fsim_fail(h, FSIM_ERR_INVALID, "bad obstacle kind %d", obs[i].kind);
    h->nobs = n;
    if (n) memcpy(h->obs, obs, sizeof(FsimObstacle) * n);
    return k_upload_obstacles(h);
}

int fsim_set_particle_radius(fsim_t* h, double r) {
    BIND(h);
    if (!(r > 0)) return fsim_fail(h, FSIM_ERR_INVALID, "particle radius must be > 0");
    h->particle_r = r;
    return k_upload_obstacles(h);  // push-out box extents include the radius
}

int fsim_get_obstacles(const fsim_t* h, FsimObstacle* out, int cap, int* n) {
    if (!h) return FSIM_ERR_INVALID;
    const int m = cap < h->nobs ? cap : h->nobs;
    if (out && m > 0) memcpy(out, h->obs, sizeof(FsimObstacle) * m);
    if (n) *n = h->nobs;
    return FSIM_OK;
}

static int upload_aos(fsim* h, const double* aos15, int64_t first, int64_t n) {
    const int64_t chunk = (int64_t)(h->stage_bytes / (15 * sizeof(double)));
    for (int64_t o = 0; o < n; o += chunk) {
        const int64_t m = std::min(chunk, n - o);
        FSIM_CUDA(h, cudaMemcpyAsync(h->stage, aos15 + 15 * o, sizeof(double) * 15 * m, cudaMemcpyHostToDevice, h->stream));
        TRY(k_particles_aos_to_soa(h, (const double*)h->stage, first + o, m));
        FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return FSIM_OK;
}

int fsim_upload_particles(fsim_t* h, const double* aos15, int64_t n) {
    BIND_FLUSH(h);
    if (n < 0 || (n > 0 && !aos15)) return fsim_fail(h, FSIM_ERR_INVALID, "bad particle buffer");
    if (n >= (int64_t(1) << 31)) return fsim_fail(h, FSIM_ERR_NOMEM, "too many particles");
    h->np = 0;
    TRY(ensure_capacity(h, n));
    TRY(upload_aos(h, aos15, 0, n));
    h->np = n;
    TRY(k_iota_ids(h, 0, n, 0));
    h->next_id = (uint32_t)n;
    h->sorted = false; h->binned = false; h->kill_pending = false;
    return FSIM_OK;
}

int fsim_append_particles(fsim_t* h, const double* aos15, int64_t n) {
    BIND_FLUSH(h);
    if (n < 0 || (n > 0 && !aos15)) return fsim_fail(h, FSIM_ERR_INVALID, "bad particle buffer");
    if (n == 0) return FSIM_OK;
    TRY(ensure_capacity(h, h->np + n));
    TRY(upload_aos(h, aos15, h->np, n));
    TRY(k_iota_ids(h, h->np, n, h->next_id));  // ids continue after the largest id ever handed out
    h->next_id += (uint32_t)n;
    h->np += n;
    h->sorted = false; h->binned = false;
    return FSIM_OK;
}

int fsim_remove_particles(fsim_t* h, const int32_t* ids, int64_t n) {
    BIND_FLUSH(h);
    if (n < 0 || (n > 0 && !ids)) return fsim_fail(h, FSIM_ERR_INVALID, "bad id buffer");
    if (n == 0) return FSIM_OK;
    if ((size_t)n * sizeof(int32_t) > h->stage_bytes) return fsim_fail(h, FSIM_ERR_NOMEM, "too many ids in one call");
    FSIM_CUDA(h, cudaMemcpyAsync(h->stage, ids, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    return k_compact_remove(h, (const int32_t*)h->stage, n);
}

int fsim_particle_count(const fsim_t* h, int64_t* n) {
    if (!h || !n) return FSIM_ERR_INVALID;
    *n = h->np;
    return FSIM_OK;
}

int fsim_download_particles(fsim_t* h, double* aos15, int64_t cap, int64_t* n) {
    BIND_FLUSH(h);
    if (n) *n = h->np;
    if (!aos15) return FSIM_OK;
    const int64_t total = std::min(cap, h->np);
    const int64_t chunk = (int64_t)(h->stage_bytes / (15 * sizeof(double)));
    for (int64_t o = 0; o < total; o += chunk) {
        const int64_t m = std::min(chunk, total - o);
        TRY(k_particles_soa_to_aos(h, (double*)h->stage, o, m));
        FSIM_CUDA(h, cudaMemcpyAsync(aos15 + 15 * o, h->stage, sizeof(double) * 15 * m, cudaMemcpyDeviceToHost, h->stream));
        FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return FSIM_OK;
}

int fsim_download_particle_ids(fsim_t* h, uint32_t* out, int64_t cap) {
    BIND(h);
    if (!h->track_ids) return fsim_fail(h, FSIM_ERR_INVALID, "id tracking is off");
    const int64_t m = std::min(cap, h->np);
    if (m > 0) FSIM_CUDA(h, cudaMemcpyAsync(out, h->ps[h->cur].id, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, h->stream));
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    return FSIM_OK;
}

int fsim_upload_particle_ids(fsim_t* h, const uint32_t* ids, int64_t n) {
    BIND(h);
    if (!h->track_ids) return fsim_fail(h, FSIM_ERR_INVALID, "id tracking is off");
    if (n != h->np || (n > 0 && !ids)) return fsim_fail(h, FSIM_ERR_INVALID, "expected %lld ids", (long long)h->np);
    if (n > 0) FSIM_CUDA(h, cudaMemcpyAsync(h->ps[h->cur].id, ids, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, h->stream));
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    return FSIM_OK;
}

int fsim_set_id_tracking(fsim_t* h, int on) {
    BIND(h);
    if (on && !h->track_ids) {
        for (int k = 0; k < 2; k++) TRY(dev_alloc(h, &h->ps[k].id, h->cap));
        h->track_ids = true;
        TRY(k_iota_ids(h, 0, h->np, 0));
        h->next_id = (uint32_t)h->np;
    } else if (!on && h->track_ids) {
        for (int k = 0; k < 2; k++) { cudaFree(h->ps[k].id); h->ps[k].id = nullptr; }
        h->track_ids = false;
    }
    return FSIM_OK;
}

int fsim_upload_particles_f32(fsim_t* h, const float* pos, const float* vel, const float* c, int64_t n) {
    BIND_FLUSH(h);
    if (n < 0 || (n > 0 && !pos)) return fsim_fail(h, FSIM_ERR_INVALID, "bad particle buffer");
    h->np = 0;
    TRY(ensure_capacity(h, n));
    if (c) TRY(ensure_c(h));
    const int64_t per = 3 + 3 + (c ? 9 : 0);
    const int64_t chunk = (int64_t)(h->stage_bytes / (per * sizeof(float)));
    float* st = (float*)h->stage;
    for (int64_t o = 0; o < n; o += chunk) {
        const int64_t m = std::min(chunk, n - o);
        float *dp = st, *dv = st + 3 * m, *dc = st + 6 * m;
        FSIM_CUDA(h, cudaMemcpyAsync(dp, pos + 3 * o, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, h->stream));
        if (vel) FSIM_CUDA(h, cudaMemcpyAsync(dv, vel + 3 * o, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, h->stream));
        if (c) FSIM_CUDA(h, cudaMemcpyAsync(dc, c + 9 * o, sizeof(float) * 9 * m, cudaMemcpyHostToDevice, h->stream));
        TRY(k_particles_f32_to_soa(h, dp, vel ? dv : nullptr, c ? dc : nullptr, o, m));
        FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    h->np = n;
    TRY(k_iota_ids(h, 0, n, 0));
    h->next_id = (uint32_t)n;
    h->sorted = false; h->binned = false; h->kill_pending = false;
    return FSIM_OK;
}

int fsim_download_particles_f32(fsim_t* h, float* pos, float* vel, float* c, int64_t cap, int64_t* n) {
    BIND_FLUSH(h);
    if (n) *n = h->np;
    const int64_t total = std::min(cap, h->np);
    const int64_t per = 3 + 3 + 9;
    const int64_t chunk = (int64_t)(h->stage_bytes / (per * sizeof(float)));
    float* st = (float*)h->stage;
    for (int64_t o = 0; o < total; o += chunk) {
        const int64_t m = std::min(chunk, total - o);
        float *dp = st, *dv = st + 3 * m, *dc = st + 6 * m;
        TRY(k_particles_soa_to_f32(h, pos ? dp : nullptr, vel ? dv : nullptr, c ? dc : nullptr, o, m));
        if (pos) FSIM_CUDA(h, cudaMemcpyAsync(pos + 3 * o, dp, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, h->stream));
        if (vel) FSIM_CUDA(h, cudaMemcpyAsync(vel + 3 * o, dv, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, h->stream));
        if (c) FSIM_CUDA(h, cudaMemcpyAsync(c + 9 * o, dc, sizeof(float) * 9 * m, cudaMemcpyDeviceToHost, h->stream));
        FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return FSIM_OK;
}

// Simulator::spawnParticles (simulator.cpp:106-125): stays on the host so that the libc rand() sequence, and with it
// the spawned set, is identical to the reference's (util/random.h:13-26)
int fsim_stage_spawn(fsim_t* h, double dt) {
    BIND(h);
    // slab handles: every rank draws the WHOLE spawn set from libc rand() and keeps its planes' share, which is only the
    // single-handle sequence when each rank owns a rand() stream of its own, i.e. one process per rank.  Ranks that share a
    // process would interleave one stream (particles lost or duplicated, colliding ids): refused, not silently wrong.
    if (h->dist && dist_peer_in_process(h)) {
        bool any_source = false;
        for (int k = 0; k < h->nobs; k++) any_source |= h->obs[k].kind == FSIM_OBSTACLE_SOURCE && h->obs[k].spawn_rate * dt + h->obs[k].last_spawn_fraction >= 1.0;
        if (any_source)
            return fsim_fail(h, FSIM_ERR_INVALID, "particle spawning on a slab group needs one process per rank (libc rand() is process-global); "
                                                   "ranks of this group share a process");
    }
    std::vector<double> fresh;
    for (int k = 0; k < h->nobs; k++) {
        FsimObstacle& ob = h->obs[k];
        if (ob.kind != FSIM_OBSTACLE_SOURCE) continue;
        const double r = ob.r + h->particle_r;
        const double numD = ob.spawn_rate * dt + ob.last_spawn_fraction;
        const int num = (int)numD;
        ob.last_spawn_fraction = numD - num;
        for (int i = 0; i < num; i++) {
            const double theta = 0.0 + (2.0 * M_PI - 0.0) * ((double)std::rand() / RAND_MAX);
            const double phi = 0.0 + (M_PI - 0.0) * ((double)std::rand() / RAND_MAX);
            const double nx = r * sin(phi) * cos(theta), ny = r * sin(phi) * sin(theta), nz = r * cos(phi);
            const double inv = 1.0 / sqrt((nx * nx + ny * ny) + nz * nz);
            double q[15] = {ob.pos[0] + nx, ob.pos[1] + ny, ob.pos[2] + nz,
                            nx * inv * ob.spawn_speed, ny * inv * ob.spawn_speed, nz * inv * ob.spawn_speed,
                            0, 0, 0, 0, 0, 0, 0, 0, 0};
            fresh.insert(fresh.end(), q, q + 15);
        }
    }
    if (fresh.empty()) return FSIM_OK;
    if (h->dist) {
        // slab handle (one PROCESS per rank, libc rand() seeded identically everywhere): every rank draws the whole spawn set
        // -- so the rand() streams stay in step with the single-handle run -- and keeps the particles that start in its planes
        // (same plane expression as the device binning); ids continue a sequence all ranks advance together
        const int64_t total = (int64_t)fresh.size() / 15;
        std::vector<double> mine;
        std::vector<uint32_t> ids;
        for (int64_t i = 0; i < total; i++) {
            int iz = (int)((double)(float)fresh[15 * i + 2] * h->g.dihz);
            iz = std::min(std::max(iz, 0), h->g.gzg - 1);
            if (iz < h->g.zoff + h->g.zown0 || iz >= h->g.zoff + h->g.zown1) continue;
            mine.insert(mine.end(), fresh.begin() + 15 * i, fresh.begin() + 15 * (i + 1));
            ids.push_back(h->slab_spawn_id + (uint32_t)i);
        }
        h->slab_spawn_id += (uint32_t)total;
        if (mine.empty()) return FSIM_OK;
        const int64_t first = h->np;
        TRY(fsim_append_particles(h, mine.data(), (int64_t)ids.size()));
        if (h->track_ids) {
            FSIM_CUDA(h, cudaMemcpyAsync(h->ps[h->cur].id + first, ids.data(), sizeof(uint32_t) * ids.size(), cudaMemcpyHostToDevice, h->stream));
            FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
        }
        return FSIM_OK;
    }
    return fsim_append_particles(h, fresh.data(), (int64_t)fresh.size() / 15);
}

static int project_slab(fsim* h, double dt, int* its_out);

#define NO_SLAB(h, what)                                                                                                   \
    do {                                                                                                                   \
        if ((h)->dist) return fsim_fail((h), FSIM_ERR_INVALID, what " is not available on a slab handle: its exchanges only run inside fsim_step"); \
    } while (0)

int fsim_stage_advect(fsim_t* h, double dt) {
    BIND_FLUSH(h);
    NO_SLAB(h, "fsim_stage_advect");
    TRY(k_advect(h, dt, true, false, false));
    if (h->kill_pending) TRY(k_sort(h));  // removeParticles at the end of advectParticles (simulator.cpp:250)
    return FSIM_OK;
}
int fsim_stage_push_apart(fsim_t* h) {  /* updateParticleIntersectionHash + pushParticlesApart, simulator.cpp:61-64 */
    BIND_FLUSH(h);
    NO_SLAB(h, "fsim_stage_push_apart");
    TRY(ensure_sorted(h));
    return k_push_apart(h);
}
int fsim_stage_push_out(fsim_t* h) { BIND_FLUSH(h); return k_advect(h, 0.0, false, true, false); }
int fsim_stage_p2g(fsim_t* h) {
    BIND_FLUSH(h);
    NO_SLAB(h, "fsim_stage_p2g");
    if (h->par.stop_particles) TRY(k_advect(h, 0.0, false, false, true));
    TRY(ensure_sorted(h));
    return k_p2g(h);
}
int fsim_stage_classify(fsim_t* h, double dt) { BIND_FLUSH(h); NO_SLAB(h, "fsim_stage_classify"); TRY(ensure_sorted(h)); return k_classify(h, dt); }
int fsim_stage_post_p2g_update(fsim_t* h, double gravity_increment) { BIND_FLUSH(h); return k_post_p2g_only(h, gravity_increment); }
int fsim_stage_project(fsim_t* h, double dt, int* iterations) { BIND_FLUSH(h); return h->dist ? project_slab(h, dt, iterations) : k_project(h, dt, iterations); }
int fsim_stage_extrapolate(fsim_t* h) { BIND_FLUSH(h); NO_SLAB(h, "fsim_stage_extrapolate"); return k_extrapolate(h); }
int fsim_stage_g2p(fsim_t* h) { BIND_FLUSH(h); TRY(ensure_sorted(h)); return k_g2p(h); }

// Simulator::simulate (simulator.cpp:51-100)
// BridsonSolverGrid::solveIncompressibility on a slab handle: hybrid / replicated run on the full-grid solver context
// (DESIGN.md §7), distributed on the slab's own arrays
static int project_slab(fsim* h, double dt, int* its_out) {
    int its_local = 0;
    if (!its_out) its_out = &its_local;
    if (h->par.solver_type == FSIM_SOLVER_BASIC) {
        // BasicMacGrid::solveIncompressibility (basicMacGrid.cpp:15-102) works on the velocities in place: red-black sweeps over the
        // owned planes, the shared z faces exchanged after every colour (dist.cu HALO_BASIC_*)
        TRY(k_project(h, dt, its_out));
        return FSIM_OK;
    }
    if (h->solver) {
        // replicated projection: gather every rank's planes of the solver inputs, solve the whole system here, keep our planes
        fsim* hs = h->solver;
        TRY(dist_gather_solver_inputs(h));
        hs->par = h->par;
        hs->prof_mask = h->prof_mask;
        const int64_t ls0 = hs->launches;
        int64_t c0[K_COUNT];
        memcpy(c0, hs->launch_n, sizeof(c0));
        const int rc = k_project(hs, dt, its_out);
        if (rc) { h->err = hs->err; if (hs->sticky) h->sticky = hs->sticky; return rc; }
        h->launches += hs->launches - ls0;
        for (int k = 0; k < K_COUNT; k++) h->launch_n[k] += hs->launch_n[k] - c0[k];
        for (const ProfRec& r : hs->prof_recs) h->prof_recs.push_back(r);  // event pairs of the profiled launches
        hs->prof_recs.clear();
        h->solve = hs->solve;
        // (hybrid: hs->p holds this rank's planes and, after the last halo, the two planes around them -- exactly the local block)
        const size_t plane = (size_t)h->g.sz;
        FSIM_CUDA(h, cudaMemcpyAsync(h->p, hs->p + (size_t)h->g.zoff * plane, sizeof(double) * plane * h->g.gz, cudaMemcpyDeviceToDevice, h->stream));
        FSIM_CUDA(h, cudaMemcpyAsync(h->rhs, hs->rhs + (size_t)h->g.zoff * plane, sizeof(double) * plane * h->g.gz, cudaMemcpyDeviceToDevice, h->stream));
        if (!hs->solve.early_out) TRY(k_pressure_apply(h, dt));  // ghost planes included: the pressure around them is known
    } else {
        TRY(k_project(h, dt, its_out));  // halos of the search direction / pressure and the all-rank reductions inside
    }
    return FSIM_OK;
}

// the same step on one z-slab of the grid (dist.cu): collective over the ranks
static int step_slab(fsim* h, double dt, int* pcg_iterations) {
    TRY(dist_check(h));
    fold_timings(h);
    const int64_t l0 = h->launches;
    FSIM_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    if (h->par.spawning_enabled) TRY(fsim_stage_spawn(h, dt));
    TRY(ensure_capacity(h, h->np));  // room for this step's immigrants; must happen before the advect kernel bins
    const bool fuse = h->g2p_pending && h->sorted && h->np > 0;
    if (!fuse) TRY(flush_g2p(h));
    h->push_timed = false;
    if (h->par.push_apart_enabled) {
        // simulate() order (simulator.cpp:57-74): advect -> push apart -> push out of obstacles -> stop.  On slabs the advected
        // particles change owner first, the pair search then also sees the neighbours' boundary planes, and whatever the two
        // position passes moved across a slab boundary migrates a second time
        TRY(k_advect(h, dt, true, false, false, /*do_bin=*/true, fuse));
        TRY(dist_migrate(h));
        TRY(k_sort(h));
        TRY(ensure_capacity(h, h->np));
        TRY(dist_push_apart_ghosts(h));
        TRY(k_push_apart(h));
        TRY(k_advect(h, dt, false, true, h->par.stop_particles != 0, /*do_bin=*/true, false));
    } else {
        TRY(k_advect(h, dt, true, true, h->par.stop_particles != 0, /*do_bin=*/true, fuse));
    }
    if (h->np > 0 && !h->binned) return fsim_fail(h, FSIM_ERR_INVALID, "slab step: particles were not binned");
    TRY(dist_migrate(h));  // emigrants -> neighbours; immigrants appended behind the locals and binned
    FSIM_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
    TRY(k_sort(h));        // also reads the new particle count back
    FSIM_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
    TRY(k_p2g(h));
    TRY(dist_halo(h, HALO_P2G, false));  // ghost-plane sums of the 7 scatter channels + the neighbours' particle counts
    FSIM_CUDA(h, cudaEventRecord(h->ev[3], h->stream));
    TRY(k_classify(h, dt));
    FSIM_CUDA(h, cudaEventRecord(h->ev[4], h->stream));
    int its = 0;
    TRY(project_slab(h, dt, &its));
    FSIM_CUDA(h, cudaEventRecord(h->ev[5], h->stream));
    TRY(dist_halo(h, HALO_U2, false));
    TRY(k_extrapolate(h));        // exchanges u2 + validity between and after its two sweeps
    FSIM_CUDA(h, cudaEventRecord(h->ev[6], h->stream));
    if (h->lazy_g2p) h->g2p_pending = true;
    else TRY(k_g2p(h));
    FSIM_CUDA(h, cudaEventRecord(h->ev[7], h->stream));
    FSIM_CUDA(h, cudaEventRecord(h->ev[8], h->stream));
    h->ev_valid = true;
    h->last_step_launches = h->launches - l0;
    if (pcg_iterations) *pcg_iterations = its;
    return FSIM_OK;
}

int fsim_step(fsim_t* h, double dt, int* pcg_iterations) {
    BIND(h);
    if (h->dist) return step_slab(h, dt, pcg_iterations);
    fold_timings(h);
    const int64_t l0 = h->launches;
    FSIM_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    if (h->par.spawning_enabled) TRY(fsim_stage_spawn(h, dt));
    // the G2P the previous step deferred runs inside this step's first particle pass (same arithmetic, one HBM round trip less)
    const bool fuse = h->g2p_pending && h->sorted && h->np > 0;
    if (!fuse) TRY(flush_g2p(h));
    if (h->par.push_apart_enabled) {
        // simulate() order (simulator.cpp:57-74): advect -> push apart -> push out of obstacles -> stop
        TRY(k_advect(h, dt, true, false, false, /*do_bin=*/true, fuse));
        FSIM_CUDA(h, cudaEventRecord(h->ev[9], h->stream));
        TRY(k_sort(h));
        TRY(k_push_apart(h));
        FSIM_CUDA(h, cudaEventRecord(h->ev[10 + 5], h->stream));
        TRY(k_advect(h, dt, false, true, h->par.stop_particles != 0, /*do_bin=*/true));
        h->push_timed = true;
    } else {
        // advect + obstacle push-out + stopParticles are one pass over the particles
        TRY(k_advect(h, dt, true, true, h->par.stop_particles != 0, /*do_bin=*/true, fuse));
        h->push_timed = false;
    }
    FSIM_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
    TRY(k_sort(h));
    FSIM_CUDA(h, cudaEventRecord(h->ev[2], h->stream));
    TRY(k_p2g(h));
    FSIM_CUDA(h, cudaEventRecord(h->ev[3], h->stream));
    TRY(k_classify(h, dt));
    FSIM_CUDA(h, cudaEventRecord(h->ev[4], h->stream));
    int its = 0;
    TRY(k_project(h, dt, &its));
    FSIM_CUDA(h, cudaEventRecord(h->ev[5], h->stream));
    TRY(k_extrapolate(h));
    FSIM_CUDA(h, cudaEventRecord(h->ev[6], h->stream));
    if (h->lazy_g2p) h->g2p_pending = true;
    else TRY(k_g2p(h));
    FSIM_CUDA(h, cudaEventRecord(h->ev[7], h->stream));
    FSIM_CUDA(h, cudaEventRecord(h->ev[8], h->stream));
    h->ev_valid = true;
    h->last_step_launches = h->launches - l0;
    if (pcg_iterations) *pcg_iterations = its;
    return FSIM_OK;
}

int fsim_download_grid(fsim_t* h, int field, void* out, int64_t out_bytes) {
    BIND(h);
    const int64_t nc = h->g.nc;
    int64_t need = 0;
    switch (field) {
        case FSIM_FIELD_TYPE: need = nc; break;
        case FSIM_FIELD_V: case FSIM_FIELD_V2: case FSIM_FIELD_WSUM: need = nc * 3 * 8; break;
        case FSIM_FIELD_AVGPNUM: case FSIM_FIELD_PRESSURE: case FSIM_FIELD_RHS: need = nc * 8; break;
        case FSIM_FIELD_PCOUNT: need = nc * 4; break;
        default: return fsim_fail(h, FSIM_ERR_INVALID, "unknown field %d", field);
    }
    if (!out || out_bytes < need) return fsim_fail(h, FSIM_ERR_INVALID, "output buffer too small (%lld < %lld)", (long long)out_bytes, (long long)need);
    if (field == FSIM_FIELD_PCOUNT) TRY(ensure_sorted(h));
    void* tmp = nullptr;
    FSIM_CUDA(h, cudaMalloc(&tmp, need));
    int rc = k_grid_download(h, field, tmp);
    if (!rc && cudaMemcpyAsync(out, tmp, need, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "grid D2H failed");
    if (cudaStreamSynchronize(h->stream) != cudaSuccess && !rc) rc = fsim_fail(h, FSIM_ERR_CUDA, "grid download sync failed");
    cudaFree(tmp);
    return rc;
}

int fsim_upload_grid(fsim_t* h, int field, const void* in, int64_t in_bytes) {
    BIND_FLUSH(h);
    const int64_t nc = h->g.nc;
    int64_t need = 0;
    switch (field) {
        case FSIM_FIELD_TYPE: need = nc; break;
        case FSIM_FIELD_V: case FSIM_FIELD_V2: case FSIM_FIELD_WSUM: need = nc * 3 * 8; break;
        case FSIM_FIELD_AVGPNUM: need = nc * 8; break;
        default: return fsim_fail(h, FSIM_ERR_INVALID, "field %d cannot be uploaded", field);
    }
    if (!in || in_bytes < need) return fsim_fail(h, FSIM_ERR_INVALID, "input buffer too small");
    void* tmp = nullptr;
    FSIM_CUDA(h, cudaMalloc(&tmp, need));
    int rc = FSIM_OK;
    if (cudaMemcpyAsync(tmp, in, need, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "grid H2D failed");
    if (!rc) rc = k_grid_upload(h, field, tmp);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess && !rc) rc = fsim_fail(h, FSIM_ERR_CUDA, "grid upload sync failed");
    cudaFree(tmp);
    return rc;
}

int fsim_download_particle_cells(fsim_t* h, int32_t* out, int64_t cap) {
    BIND(h);
    const int64_t m = std::min(cap, h->np);
    if (m <= 0) return FSIM_OK;
    int32_t* tmp = nullptr;
    FSIM_CUDA(h, cudaMalloc((void**)&tmp, sizeof(int32_t) * h->np));
    int rc = k_particle_cells(h, tmp);
    if (!rc && cudaMemcpyAsync(out, tmp, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) rc = fsim_fail(h, FSIM_ERR_CUDA, "D2H failed");
    if (cudaStreamSynchronize(h->stream) != cudaSuccess && !rc) rc = fsim_fail(h, FSIM_ERR_CUDA, "sync failed");
    cudaFree(tmp);
    return rc;
}

int fsim_export_gfx(fsim_t* h, FsimParticleGfx* out, int64_t cap, int64_t* n) {
    BIND_FLUSH(h);
    if (n) *n = h->np;
    const int64_t m = std::min(cap, h->np);
    if (m <= 0 || !out) return FSIM_OK;
    TRY(ensure_gfx(h));
    TRY(k_export_gfx(h, h->gfx));
    FSIM_CUDA(h, cudaMemcpyAsync(out, h->gfx, sizeof(FsimParticleGfx) * m, cudaMemcpyDeviceToHost, h->stream));
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    return FSIM_OK;
}

static int export_gfx_async_impl(fsim_t* h, FsimParticleGfx* out, int64_t cap, int64_t stride, int64_t* n) {
    BIND_FLUSH(h);
    if (stride < 1) return fsim_fail(h, FSIM_ERR_INVALID, "stride must be >= 1");
    const int64_t nout = (h->np + stride - 1) / stride;
    if (n) *n = nout;
    const int64_t m = std::min(cap, nout);
    if (m <= 0 || !out) return FSIM_OK;
    if (!h->copy_stream) {
        FSIM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            FSIM_CUDA(h, cudaEventCreateWithFlags(&h->gfx_ready[k], cudaEventDisableTiming));
            FSIM_CUDA(h, cudaEventCreateWithFlags(&h->gfx_copied[k], cudaEventDisableTiming));
        }
    }
    const int k = h->gfx_slot;
    h->gfx_slot ^= 1;
    // a staging buffer is reused every second export: its previous copy must have left before the kernel overwrites it
    if (h->gfx_inflight[k]) FSIM_CUDA(h, cudaStreamWaitEvent(h->stream, h->gfx_copied[k], 0));
    if (!(h->gfx_async[k] && h->gfx_async_cap[k] >= h->np)) {
        if (h->gfx_inflight[k]) FSIM_CUDA(h, cudaEventSynchronize(h->gfx_copied[k]));
        cudaFree(h->gfx_async[k]);
        h->gfx_async[k] = nullptr;
        h->gfx_async_cap[k] = h->cap;
        FSIM_CUDA(h, cudaMalloc((void**)&h->gfx_async[k], sizeof(FsimParticleGfx) * (size_t)(h->cap > 0 ? h->cap : 1)));
    }
    TRY(k_export_gfx(h, h->gfx_async[k], stride));
    FSIM_CUDA(h, cudaEventRecord(h->gfx_ready[k], h->stream));
    FSIM_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->gfx_ready[k], 0));
    FSIM_CUDA(h, cudaMemcpyAsync(out, h->gfx_async[k], sizeof(FsimParticleGfx) * m, cudaMemcpyDeviceToHost, h->copy_stream));
    FSIM_CUDA(h, cudaEventRecord(h->gfx_copied[k], h->copy_stream));
    h->gfx_inflight[k] = true;
    return FSIM_OK;
}

int fsim_export_gfx_async(fsim_t* h, FsimParticleGfx* out, int64_t cap, int64_t* n) { return export_gfx_async_impl(h, out, cap, 1, n); }
int fsim_export_gfx_strided_async(fsim_t* h, FsimParticleGfx* out, int64_t cap, int64_t stride, int64_t* n) { return export_gfx_async_impl(h, out, cap, stride, n); }

int fsim_export_gfx_wait(fsim_t* h) {
    BIND(h);
    for (int k = 0; k < 2; k++)
        if (h->gfx_inflight[k]) {
            FSIM_CUDA(h, cudaEventSynchronize(h->gfx_copied[k]));
            h->gfx_inflight[k] = false;
        }
    return FSIM_OK;
}

int fsim_export_gfx_wait_previous(fsim_t* h) {
    BIND(h);
    const int k = h->gfx_slot;  // the slot of the export before the most recent one
    if (h->gfx_inflight[k]) {
        FSIM_CUDA(h, cudaEventSynchronize(h->gfx_copied[k]));
        h->gfx_inflight[k] = false;
    }
    return FSIM_OK;
}

int fsim_get_step_durations(const fsim_t* hc, FsimTimings* out) {
    fsim* h = const_cast<fsim*>(hc);
    BIND(h);
    if (!out) return FSIM_ERR_INVALID;
    fold_timings(h);
    *out = h->timings;
    out->incompressibility_it_count = h->solve.iterations;
    return FSIM_OK;
}

int fsim_get_solve_info(const fsim_t* h, FsimSolveInfo* out) {
    if (!h || !out) return FSIM_ERR_INVALID;
    *out = h->solve;
    return FSIM_OK;
}

int fsim_get_last_step_stats(const fsim_t* hc, double* device_ms, int64_t* kernel_launches) {
    fsim* h = const_cast<fsim*>(hc);
    BIND(h);
    fold_timings(h);
    if (device_ms) *device_ms = h->last_step_ms;
    if (kernel_launches) *kernel_launches = h->last_step_launches;
    return FSIM_OK;
}

static const char* const kKernelNames[K_COUNT] = {"advect", "bin", "scan", "reorder", "p2g", "classify", "finalize", "rhs",
                                                    "pcg_init", "spmv", "pcg_update", "pcg_direction", "mg", "mg_level1", "mg_coarse", "pressure_apply",
                                                    "extrapolate", "g2p", "gfx", "memset", "push_apart", "halo", "allreduce", "migrate"};

int fsim_kernel_class_count(void) { return K_COUNT; }
const char* fsim_kernel_class_name(int kid) { return (kid >= 0 && kid < K_COUNT) ? kKernelNames[kid] : ""; }

int fsim_profile_enable(fsim_t* h, uint32_t class_mask) {
    BIND(h);
    h->prof_mask = class_mask;
    return FSIM_OK;
}

int fsim_profile_read(fsim_t* h, double* total_ms, int64_t* profiled_launches, int64_t* launches, int reset) {
    BIND(h);
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    for (const ProfRec& r : h->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { h->prof_ms[r.kid] += ms; h->prof_n[r.kid]++; }
        h->prof_free.push_back(r.e0);
        h->prof_free.push_back(r.e1);
    }
    h->prof_recs.clear();
    for (int k = 0; k < K_COUNT; k++) {
        if (total_ms) total_ms[k] = h->prof_ms[k];
        if (profiled_launches) profiled_launches[k] = h->prof_n[k];
        if (launches) launches[k] = h->launch_n[k];
    }
    if (reset) { memset(h->prof_ms, 0, sizeof(h->prof_ms)); memset(h->prof_n, 0, sizeof(h->prof_n)); memset(h->launch_n, 0, sizeof(h->launch_n)); }
    return FSIM_OK;
}

int fsim_timer_record(fsim_t* h, int slot) {
    BIND(h);
    if (slot < 0 || slot > 4) return fsim_fail(h, FSIM_ERR_INVALID, "timer slot out of range");
    FSIM_CUDA(h, cudaEventRecord(h->ev[10 + slot], h->stream));
    return FSIM_OK;
}

int fsim_timer_elapsed_ms(fsim_t* h, int slot_begin, int slot_end, double* ms) {
    BIND(h);
    if (slot_begin < 0 || slot_begin > 4 || slot_end < 0 || slot_end > 4 || !ms) return fsim_fail(h, FSIM_ERR_INVALID, "bad timer slots");
    FSIM_CUDA(h, cudaEventSynchronize(h->ev[10 + slot_end]));
    float f = 0.f;
    FSIM_CUDA(h, cudaEventElapsedTime(&f, h->ev[10 + slot_begin], h->ev[10 + slot_end]));
    *ms = f;
    return FSIM_OK;
}

int fsim_synchronize(fsim_t* h) {
    BIND(h);
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    TRY(dist_check(h));
    return FSIM_OK;
}

}  // extern "C"
