// pcg.cu -- pressure projection: BridsonSolverGrid::solveIncompressibility (bridsonSolverGrid.cpp:244-293).
//
// Same linear system and stopping rule as the reference (RHS :79-89, 7-point matrix :40-77, early-out when
// sum(rhs^2) < 1e-7 :254-258, stop when ||r||_inf < residualTolerance :276-281, NaN guard :271), but
//   * matrix-free: the row of a cell is a 16-bit code (6-bit WATER-neighbour mask + #non-solid neighbours) computed
//     once per solve from the cell flags; the reference's compacted fluid-cell list / AMatrixRow array / cell.id
//     indirection do not exist;
//   * the sequential MIC(0) preconditioner (:91-163) is replaced by a parallel one (multigrid cycle, mg.cu; or diagonal
//     scaling with FSIM_PRECOND=jacobi) -- equivalence is on the converged pressure field;
//   * CG vectors (p, r, s, q) are fp64: with |p| ~ 1e4 and a 1e-6 absolute tolerance an fp32 recurrence stalls
//     (SURVEY.md §7 "hard parts"); the preconditioner works in fp32;
//   * SpMV is fused with the s.q reduction, the x/r update with ||r||_inf and (diagonal case) z.r; every thread owns two
//     consecutive cells (16-byte loads of the fp64 vectors); reductions are warp shuffle -> block -> fixed-order
//     last-block sum => deterministic, no fp64 atomics;
//   * all scalars (sigma, alpha, beta, done flag, iteration count) stay on the device; the host only polls the done flag.
#include <cstdio>

#include "fsim_internal.h"
#include "reduce.cuh"
#include "pcg_finish.cuh"
#include "dist_dev.cuh"
#include "fexch.cuh"
#include "launch.cuh"
#include "tile4.cuh"

int mg_apply(fsim* h, bool first_done, bool with_dot);  // z = M^-1 r   (mg.cu)
int mg_update_first(fsim* h);
bool mg_can_fuse(const fsim* h);
bool mg_enabled(const fsim* h);

namespace {

constexpr int PT = RED_THREADS;  // threads per block
constexpr int CV = 2;       // consecutive cells per thread and trip (16-byte fp64 loads)
constexpr int CHUNK = 8192; // cells per block of the reducing kernels: few partials => short fixed-order final sum

struct PcgArgs {
    GridDims g;
    const uint8_t* flags;
    uint16_t* code;
    uint16_t* code_full;  // hybrid solver context: codes of ALL fluid cells (code is then owned-only); nullptr otherwise
    uint16_t* code_mg;  // slab mode: the codes the multigrid sees (links into ghost planes cut); nullptr otherwise
    const float* u2[3];
    const float* dens;
    double *p, *rhs, *r, *q, *z;
    float* s;  // search direction, stored in fp32: A s and p += alpha s both use the stored value, so r = b - A p keeps holding to fp64 rounding
    float* p_prev;     // last step's converged pressure (fp32 copy) for the extrapolated warm start
    const float* z32;  // multigrid result (fp32) when the multigrid preconditioner is active, else nullptr
    PcgScalars* sc;
    double* partials;  // [3][gridDim.x]
    unsigned int* counter;
    PcgHostStatus* status;  // host-mapped progress flags
    double inv_h;           // 1/h
    double avg_pressure, pressure_k;
    int pressure_enabled, warm;
    int cut_neumann;  // slab mode: how code_mg treats a cut link
    int64_t cb, ce;   // cell range of the chunked kernels
    int64_t rc0, rc1; // cell range rhs_kernel visits (hybrid: the owned planes; the other planes only need codes, codes_full4_kernel)
    Tile4 t4;         // the same range as 2-D blocks (tile4.cuh): iteration space of the four-cells-per-thread kernels
    // hybrid slab projection with fused exchanges (fexch.cuh): direction4 stores the boundary planes of s into the neighbours'
    // ghost planes (t4p: boundary planes first), spmv4's boundary CTAs wait for the neighbours' (t4: boundary planes last)
    Tile4 t4p;
    FxPush fxs;
    FxWait fws;
    const ArDev* ar;  // all-rank reductions finished inside the reducing kernels (dist_dev.cuh ar_warp); nullptr: allreduce_kernel
};

// iteration space of the chunked kernels: block b owns cells [b*CHUNK, (b+1)*CHUNK), a thread visits CV cells per trip
// (a.cb, a.ce): the cell range the kernel visits -- the whole grid, or the planes this rank owns (hybrid slab projection)
#define FOR_CHUNK(c0, nc)                                                                                      \
    for (int64_t c0 = a.cb + (int64_t)blockIdx.x * CHUNK + (int64_t)threadIdx.x * CV,                           \
                 cend__ = min(a.cb + (int64_t)(blockIdx.x + 1) * CHUNK, a.ce);                                  \
         c0 < cend__; c0 += PT * CV)

// stencil code of a cell (calculateAMatrix, bridsonSolverGrid.cpp:40-77): bits 0-5 WATER neighbours (-x,+x,-y,+y,-z,+z),
// bits 6-8 number of non-solid neighbours (the diagonal / scale), bit 15 the cell itself is WATER.  0 for other cells.
__device__ __forceinline__ int code_ns(unsigned c) { return (c >> 6) & 7; }

// calculateRHS (:79-89) + p = 0, r = rhs + sum(rhs^2) (:254-258) + stencil codes; one thread per cell
__global__ void __launch_bounds__(PT) rhs_kernel(PcgArgs a) {
    double acc[2] = {0.0, 0.0};
    const int64_t c = a.rc0 + (int64_t)blockIdx.x * PT + threadIdx.x;
    if (c < a.rc1) {
        double rhs = 0.0;
        unsigned code = 0;
        const int zc = (int)c / a.g.sz;
        const bool owned = zc >= a.g.zown0 && zc < a.g.zown1;  // slab modes: only the planes this rank owns become active
        unsigned full = 0;
        if ((owned || a.code_full) && (a.flags[c] & FL_TYPE_MASK) == FSIM_CELL_WATER) {  // WATER cells are interior: all six neighbours exist
            const int64_t nb[6] = {c - 1, c + 1, c - a.g.sy, c + a.g.sy, c - a.g.sz, c + a.g.sz};
            unsigned wm = 0, ns = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const int t = a.flags[nb[k]] & FL_TYPE_MASK;
                ns += (t != FSIM_CELL_SOLID);
                wm |= (t == FSIM_CELL_WATER) ? (1u << k) : 0u;
            }
            full = CODE_ACTIVE | (ns << 6) | wm;
            if (owned) {
                code = full;
                const double div = ((double)a.u2[0][c] + (double)a.u2[1][c] + (double)a.u2[2][c]) - (double)a.u2[0][c - 1] -
                                   (double)a.u2[1][c - a.g.sy] - (double)a.u2[2][c - a.g.sz];
                rhs = -a.inv_h * div + (a.pressure_enabled ? ((double)a.dens[c] - a.avg_pressure) * a.pressure_k : 0.0);
                acc[0] = rhs * rhs;
                acc[1] = 1.0;
            }
        }
        a.code[c] = (uint16_t)code;
        if (a.code_full) a.code_full[c] = (uint16_t)full;
        // hybrid slab projection: of the planes of other ranks only the codes are needed (the replicated coarse operators); their
        // CG vectors are never read -- kernels walk the owned planes, ghost planes are filled by the exchanges (p before the
        // warm-start residual and before the pressure is applied, s every iteration)
        if (a.code_full && !owned) goto reduce;
        if (a.code_mg) {
            // the preconditioner is block-local: a link into a ghost plane is cut, either leaving the neighbour on the diagonal
            // (Dirichlet: the exact diagonal block of A) or dropping it from the diagonal as well (Neumann: A = M + a positive
            // semi-definite interface term); measured: both need 3-5x the iterations of the global multigrid (FSIM_SLAB_CUT)
            unsigned cm = code;
            if (zc - 1 < a.g.zown0 && (cm & 16u)) { cm &= ~16u; if (a.cut_neumann) cm -= 1u << 6; }
            if (zc + 1 >= a.g.zown1 && (cm & 32u)) { cm &= ~32u; if (a.cut_neumann) cm -= 1u << 6; }
            a.code_mg[c] = (uint16_t)cm;
        }
        a.rhs[c] = rhs;
        if (!a.warm) { a.r[c] = rhs; a.p[c] = 0.0; if (a.p_prev) a.p_prev[c] = 0.f; }
        else {
            // warm start from last step's pressure, linearly extrapolated in time where two previous solutions exist
            // (initial guess only: the converged field does not depend on it)
            const double p1 = (code & CODE_ACTIVE) ? a.p[c] : 0.0;
            double guess = p1;
            if (a.p_prev) {
                const double p0 = (double)a.p_prev[c];
                if (a.warm > 1 && p1 != 0.0 && p0 != 0.0) guess = 2.0 * p1 - p0;
                a.p_prev[c] = (float)p1;
            }
            a.p[c] = (code & CODE_ACTIVE) ? guess : 0.0;
        }
    }
reduce:
    double out[2];
    if (grid_reduce<2, 0>(acc, a.partials, a.counter, out)) {
        if (a.sc->dist) { a.sc->loc[0] = out[0]; a.sc->loc[3] = out[1]; a.sc->done = 0; }
        else pcg_finish_rhs(a.sc, a.status, out[0], out[1]);
    }
}

// Hybrid slab projection: the replicated coarse operators need the stencil codes of EVERY plane, the right-hand side only
// exists on the owned ones.  The planes of the other ranks get their codes here, four cells per thread (flags as 4-byte
// loads; x neighbours of the inner cells from registers): 14.7 M cells at N = 8 in ~30 us where rhs_kernel, one cell per
// thread with seven byte loads each, took 190 us over the whole grid.
__global__ void __launch_bounds__(256) codes_full4_kernel(GridDims g, const uint8_t* __restrict__ flags, uint16_t* __restrict__ code_full, int nplanes) {
    const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int pz = blockIdx.z;  // index among the planes this rank does not own: below the slab first, then above it
    if (x >= g.gx || y >= g.gy || pz >= nplanes) return;
    const int z = pz < g.zown0 ? pz : pz + (g.zown1 - g.zown0);
    const int64_t c = ((int64_t)z * g.gy + y) * g.gx + x;
    const uchar4 f4 = *reinterpret_cast<const uchar4*>(flags + c);
    const unsigned t[4] = {f4.x & FL_TYPE_MASK, f4.y & FL_TYPE_MASK, f4.z & FL_TYPE_MASK, f4.w & FL_TYPE_MASK};
    unsigned out[4] = {0, 0, 0, 0};
    if (t[0] == FSIM_CELL_WATER || t[1] == FSIM_CELL_WATER || t[2] == FSIM_CELL_WATER || t[3] == FSIM_CELL_WATER) {
        // a WATER cell is interior (border shell), so the rows y +- 1, z +- 1 of this group exist
        const uchar4 ym = *reinterpret_cast<const uchar4*>(flags + c - g.sy), yp = *reinterpret_cast<const uchar4*>(flags + c + g.sy);
        const uchar4 zm = *reinterpret_cast<const uchar4*>(flags + c - g.sz), zp = *reinterpret_cast<const uchar4*>(flags + c + g.sz);
        const unsigned xl = t[0] == FSIM_CELL_WATER ? (flags[c - 1] & FL_TYPE_MASK) : (unsigned)FSIM_CELL_SOLID;
        const unsigned xr = t[3] == FSIM_CELL_WATER ? (flags[c + 4] & FL_TYPE_MASK) : (unsigned)FSIM_CELL_SOLID;
        const unsigned ymv[4] = {ym.x, ym.y, ym.z, ym.w}, ypv[4] = {yp.x, yp.y, yp.z, yp.w}, zmv[4] = {zm.x, zm.y, zm.z, zm.w}, zpv[4] = {zp.x, zp.y, zp.z, zp.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (t[i] != FSIM_CELL_WATER) continue;
            // the same neighbour order as rhs_kernel: -x, +x, -y, +y, -z, +z
            const unsigned nb[6] = {i == 0 ? xl : t[i - 1], i == 3 ? xr : t[i + 1], ymv[i] & FL_TYPE_MASK, ypv[i] & FL_TYPE_MASK, zmv[i] & FL_TYPE_MASK, zpv[i] & FL_TYPE_MASK};
            unsigned wm = 0, ns = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                ns += (nb[k] != FSIM_CELL_SOLID);
                wm |= (nb[k] == FSIM_CELL_WATER) ? (1u << k) : 0u;
            }
            out[i] = CODE_ACTIVE | (ns << 6) | wm;
        }
    }
    *reinterpret_cast<ushort4*>(code_full + c) = make_ushort4((unsigned short)out[0], (unsigned short)out[1], (unsigned short)out[2], (unsigned short)out[3]);
}

// warm start: r = rhs - A p for the pressure kept from the previous step; also max |r| (already converged => done)
__global__ void __launch_bounds__(PT) residual_kernel(PcgArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.sc->done) return;
    double acc[1] = {0.0};
    const double scale = a.sc->scale;
    FOR_CHUNK(c0, a.g.nc)
#pragma unroll
    for (int k = 0; k < CV; k++) {
        const int64_t c = c0 + k;
        if (c >= a.g.nc) break;
        const unsigned cd = a.code[c];
        double r = 0.0;
        if (cd & CODE_ACTIVE) {
            double n = 0.0;
            if (cd & 1u) n += a.p[c - 1];
            if (cd & 2u) n += a.p[c + 1];
            if (cd & 4u) n += a.p[c - a.g.sy];
            if (cd & 8u) n += a.p[c + a.g.sy];
            if (cd & 16u) n += a.p[c - a.g.sz];
            if (cd & 32u) n += a.p[c + a.g.sz];
            r = a.rhs[c] - scale * ((double)code_ns(cd) * a.p[c] - n);
            acc[0] = fmax(acc[0], fabs(r));
        }
        a.r[c] = r;
    }
    double out[1];
    if (grid_reduce<0, 1>(acc, a.partials, a.counter, out)) {
        if (a.sc->dist) a.sc->loc[2] = out[0];
        else pcg_finish_residual(a.sc, a.status, out[0]);
    }
}

__device__ __forceinline__ double jacobi_z(double scale, unsigned code, double r) {
    const int ns = code_ns(code);
    return ns > 0 ? r / (scale * ns) : r;
}

// first preconditioner application (:262-265): s = z ; sigma = z.r, with z = r / A_ii (diagonal) or the multigrid result
template <bool JACOBI>
__global__ void __launch_bounds__(PT) start_kernel(PcgArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.sc->done) return;
    double acc[1] = {0.0};
    FOR_CHUNK(c0, a.g.nc)
#pragma unroll
    for (int k = 0; k < CV; k++) {
        const int64_t c = c0 + k;
        if (c >= a.g.nc) break;
        const unsigned code = a.code[c];
        double z = 0.0;
        if (code & CODE_ACTIVE) {
            const double r = a.r[c];
            z = JACOBI ? jacobi_z(a.sc->scale, code, r) : (double)a.z32[c];
            acc[0] += z * r;
        }
        if (JACOBI) a.z[c] = z;
        a.s[c] = (float)z;
    }
    double out[1];
    if (grid_reduce<1, 0>(acc, a.partials, a.counter, out)) { if (a.sc->dist) a.sc->loc[0] = out[0]; else a.sc->sigma = out[0]; }
}

// q = A s fused with s.q  (applyAMatrix :165-198 + dotProduct :200-214); VEC: 16-byte loads, two cells per thread
template <bool VEC>
__global__ void __launch_bounds__(PT) spmv_kernel(PcgArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.sc->done) return;
    double acc[1] = {0.0};
    const int64_t sy = a.g.sy, sz = a.g.sz;
    const double scale = a.sc->scale;
    FOR_CHUNK(c, a.g.nc)
    if (VEC) {
        {  // nc is even in this path
            const ushort2 cd = *reinterpret_cast<const ushort2*>(a.code + c);
            const unsigned c0 = cd.x, c1 = cd.y;
            if ((c0 | c1) & CODE_ACTIVE) {
                auto ld2 = [&](const float* src) { const float2 v = *reinterpret_cast<const float2*>(src); return make_double2((double)v.x, (double)v.y); };
                const double2 sc = ld2(a.s + c);
                const unsigned any = c0 | c1;
                double2 ym = {0, 0}, yp = {0, 0}, zm = {0, 0}, zp = {0, 0};
                if (any & 4u) ym = ld2(a.s + c - sy);
                if (any & 8u) yp = ld2(a.s + c + sy);
                if (any & 16u) zm = ld2(a.s + c - sz);
                if (any & 32u) zp = ld2(a.s + c + sz);
                const double xl = (c0 & 1u) ? (double)a.s[c - 1] : 0.0, xr = (c1 & 2u) ? (double)a.s[c + 2] : 0.0;
                double2 q = {0, 0};
                if (c0 & CODE_ACTIVE) {
                    double n = xl;
                    if (c0 & 2u) n += sc.y;
                    if (c0 & 4u) n += ym.x;
                    if (c0 & 8u) n += yp.x;
                    if (c0 & 16u) n += zm.x;
                    if (c0 & 32u) n += zp.x;
                    q.x = scale * ((double)code_ns(c0) * sc.x - n);
                    acc[0] += sc.x * q.x;
                }
                if (c1 & CODE_ACTIVE) {
                    double n = xr;
                    if (c1 & 1u) n += sc.x;
                    if (c1 & 4u) n += ym.y;
                    if (c1 & 8u) n += yp.y;
                    if (c1 & 16u) n += zm.y;
                    if (c1 & 32u) n += zp.y;
                    q.y = scale * ((double)code_ns(c1) * sc.y - n);
                    acc[0] += sc.y * q.y;
                }
                *reinterpret_cast<double2*>(a.q + c) = q;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < CV; k++) {
            const int64_t cc = c + k;
            if (cc >= a.g.nc) break;
            const unsigned cd = a.code[cc];
            if (!(cd & CODE_ACTIVE)) continue;
            const double sc = (double)a.s[cc];
            double n = 0.0;
            if (cd & 1u) n += (double)a.s[cc - 1];
            if (cd & 2u) n += (double)a.s[cc + 1];
            if (cd & 4u) n += (double)a.s[cc - sy];
            if (cd & 8u) n += (double)a.s[cc + sy];
            if (cd & 16u) n += (double)a.s[cc - sz];
            if (cd & 32u) n += (double)a.s[cc + sz];
            const double q = scale * ((double)code_ns(cd) * sc - n);
            a.q[cc] = q;
            acc[0] += sc * q;
        }
    }
    double out[1];
    if (grid_reduce<1, 0>(acc, a.partials, a.counter, out)) { if (a.sc->dist) a.sc->loc[0] = out[0]; else a.sc->sq = out[0]; }
}

// same as spmv_kernel<true> with four cells per thread and trip (gx % 4 == 0): 13 independent loads in flight per thread
// (issuing all neighbour loads unconditionally, in parallel with the code load, was measured slower: +50 % traffic over AIR)
__global__ void __launch_bounds__(PT, 4) spmv4_kernel(PcgArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.sc->done) return;
    if (a.fws.has[0] | a.fws.has[1]) {  // fused exchange of s: a CTA on a boundary plane reads the neighbour's plane
        const int side = fx_side(a.fws.zb, tile4_plane(a.t4));
        if (fx_has(a.fws, side)) fx_wait(a.fws, side);
    }
    double acc[1] = {0.0};
    const int64_t sy = a.g.sy, sz = a.g.sz;
    const double scale = a.sc->scale;
    // A thread visits four groups of four cells (tile4.cuh); every trip is two dependent round trips (stencil codes, then the
    // stencil itself).  The codes of the NEXT trip are requested before this trip's stencil loads.
    int64_t c = 0, cn = 0;
    ushort4 t = make_ushort4(0, 0, 0, 0);
    if (tile4_cell(a.t4, 0, c)) t = *reinterpret_cast<const ushort4*>(a.code + c);
    for (int trip = 0; trip < T4_TRIPS; trip++, c = cn) {
        ushort4 tn = make_ushort4(0, 0, 0, 0);
        if (trip + 1 < T4_TRIPS && tile4_cell(a.t4, trip + 1, cn)) tn = *reinterpret_cast<const ushort4*>(a.code + cn);
        const unsigned cd[4] = {t.x, t.y, t.z, t.w};
        const unsigned any = cd[0] | cd[1] | cd[2] | cd[3];
        if (any & CODE_ACTIVE) {
            double sc[4], ym[4] = {0, 0, 0, 0}, yp[4] = {0, 0, 0, 0}, zm[4] = {0, 0, 0, 0}, zp[4] = {0, 0, 0, 0};
            double xl = 0.0, xr = 0.0;
            auto ld = [&](double* dst, const float* src) {
                const float4 u = *reinterpret_cast<const float4*>(src);
                dst[0] = (double)u.x; dst[1] = (double)u.y; dst[2] = (double)u.z; dst[3] = (double)u.w;
            };
            ld(sc, a.s + c);
            if (any & 4u) ld(ym, a.s + c - sy);
            if (any & 8u) ld(yp, a.s + c + sy);
            if (any & 16u) ld(zm, a.s + c - sz);
            if (any & 32u) ld(zp, a.s + c + sz);
            xl = (cd[0] & 1u) ? (double)a.s[c - 1] : 0.0; xr = (cd[3] & 2u) ? (double)a.s[c + 4] : 0.0;
            double q[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (cd[i] & CODE_ACTIVE) {
                    double n = 0.0;
                    if (cd[i] & 1u) n += i == 0 ? xl : sc[i - 1];
                    if (cd[i] & 2u) n += i == 3 ? xr : sc[i + 1];
                    if (cd[i] & 4u) n += ym[i];
                    if (cd[i] & 8u) n += yp[i];
                    if (cd[i] & 16u) n += zm[i];
                    if (cd[i] & 32u) n += zp[i];
                    q[i] = scale * ((double)code_ns(cd[i]) * sc[i] - n);
                    acc[0] += sc[i] * q[i];
                }
            *reinterpret_cast<double2*>(a.q + c) = make_double2(q[0], q[1]);
            *reinterpret_cast<double2*>(a.q + c + 2) = make_double2(q[2], q[3]);
        }
        t = tn;
    }
    double out[1] = {0.0};
    const int fr = grid_reduce_w<1, 0>(acc, a.partials, a.counter, out);
    if (fr && a.ar) {  // slab mode, fused reduction: warp 0 of the last CTA exchanges the partial sums with all ranks
        const double m[4] = {out[0], 0.0, 0.0, 0.0};
        ar_warp(a.ar, a.sc, a.status, AR_SPMV, m);
    } else if (fr == 2) { if (a.sc->dist) a.sc->loc[0] = out[0]; else a.sc->sq = out[0]; }
}

// alpha = sigma / s.q ; p += alpha s ; r -= alpha q ; ||r||_inf ; (JACOBI: z = r / A_ii ; sigma' = z.r)   (:270-284)
template <bool JACOBI>
__global__ void __launch_bounds__(PT) update_kernel(PcgArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.sc->done) return;
    const double alpha = a.sc->sigma / a.sc->sq;
    const bool bad = alpha != alpha;  // NaN => the reference breaks before touching p (:271-272)
    const double scale = a.sc->scale;
    double acc[2] = {0.0, 0.0};      // [0] = z.r (sum), [1] = max |r|
    if (!bad) {
        FOR_CHUNK(c0, a.g.nc)
#pragma unroll
        for (int k = 0; k < CV; k++) {
            const int64_t c = c0 + k;
            if (c >= a.g.nc) break;
            const unsigned code = a.code[c];
            if (!(code & CODE_ACTIVE)) continue;
            a.p[c] += alpha * (double)a.s[c];
            const double r = a.r[c] - alpha * a.q[c];
            a.r[c] = r;
            acc[1] = fmax(acc[1], fabs(r));
            if (JACOBI) {
                const double z = jacobi_z(scale, code, r);
                a.z[c] = z;
                acc[0] += z * r;
            }
        }
    }
    double out[2];
    if (grid_reduce<1, 1>(acc, a.partials, a.counter, out)) {
        if (a.sc->dist) { a.sc->loc[0] = out[0]; a.sc->loc[2] = out[1]; }
        else pcg_finish_update(a.sc, a.status, out[0], out[1]);
    }
}

// sigma' = z.r for an externally computed z (multigrid)
__global__ void __launch_bounds__(PT) dot_zr_kernel(PcgArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.sc->done) return;
    double acc[1] = {0.0};
    FOR_CHUNK(c0, a.g.nc)
#pragma unroll
    for (int k = 0; k < CV; k++) {
        const int64_t c = c0 + k;
        if (c >= a.g.nc) break;
        if (a.code[c] & CODE_ACTIVE) acc[0] += (double)a.z32[c] * a.r[c];
    }
    double out[1];
    if (grid_reduce<1, 0>(acc, a.partials, a.counter, out)) { if (a.sc->dist) a.sc->loc[0] = out[0]; else a.sc->sigma_new = out[0]; }
}

// beta = sigma'/sigma ; s = z + beta s   (:284-289); sigma <- sigma' is published by sigma_kernel afterwards
__global__ void __launch_bounds__(PT) direction_kernel(PcgArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.sc->done) return;
    const double beta = a.sc->sigma_new / a.sc->sigma;
    FOR_CHUNK(c0, a.g.nc)
#pragma unroll
    for (int k = 0; k < CV; k++) {
        const int64_t c = c0 + k;
        if (c >= a.g.nc) break;
        if (!(a.code[c] & CODE_ACTIVE)) continue;
        a.s[c] = (float)((double)a.s[c] * beta + (a.z32 ? (double)a.z32[c] : a.z[c]));
    }
}
// the same update with four cells per trip (gx % 4 == 0, multigrid preconditioner): 16-byte loads, stencil codes of the next
// trip prefetched (see spmv4_kernel)
__global__ void __launch_bounds__(PT) direction4_kernel(PcgArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.sc->done) return;
    const double beta = a.sc->sigma_new / a.sc->sigma;
    // fused exchange: a CTA on a boundary plane stores its groups into the neighbour's ghost plane as well, then signals
    float* speer = nullptr;
    int side = -1;
    if (a.fxs.peer[0] || a.fxs.peer[1]) {
        side = fx_side(a.fxs.zb, tile4_plane(a.t4p));
        speer = fx_peer(a.fxs, side);
        if (blockIdx.x == 0 && threadIdx.x == 0) fx_expect(a.fxs);
    }
    int64_t c = 0, cn = 0;
    ushort4 t = make_ushort4(0, 0, 0, 0);
    if (tile4_cell(a.t4p, 0, c)) t = *reinterpret_cast<const ushort4*>(a.code + c);
    for (int trip = 0; trip < T4_TRIPS; trip++, c = cn) {
        ushort4 tn = make_ushort4(0, 0, 0, 0);
        if (trip + 1 < T4_TRIPS && tile4_cell(a.t4p, trip + 1, cn)) tn = *reinterpret_cast<const ushort4*>(a.code + cn);
        const unsigned cd[4] = {t.x, t.y, t.z, t.w};
        t = tn;
        if (!((cd[0] | cd[1] | cd[2] | cd[3]) & CODE_ACTIVE)) continue;
        const float4 sv = *reinterpret_cast<const float4*>(a.s + c), zv = *reinterpret_cast<const float4*>(a.z32 + c);
        float so[4] = {sv.x, sv.y, sv.z, sv.w};
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (cd[i] & CODE_ACTIVE) so[i] = (float)((double)so[i] * beta + (double)zz[i]);
        *reinterpret_cast<float4*>(a.s + c) = make_float4(so[0], so[1], so[2], so[3]);
        if (speer) *reinterpret_cast<float4*>(speer + c) = make_float4(so[0], so[1], so[2], so[3]);
    }
    if (speer) fx_signal(a.fxs, side);
}
// the host reads the solve's scalars from mapped pinned memory: a D2H memcpy would queue behind whatever bulk copy the
// application has in flight on the device-to-host copy engine (the async gfx export is 1.3 GB per step at 256^3)
__global__ void publish_kernel(const PcgScalars* sc, PcgScalars* out) {
    *out = *sc;
    __threadfence_system();
}

// closes iteration `it`: sigma <- sigma', it <- it + 1, progress published to the host-mapped status word; inside the
// device-side WHILE graph (use_handle) the same thread re-arms the loop condition from the done flag -- one 1-thread launch
// per iteration instead of two (sigma_kernel + loop_condition_kernel)
__global__ void close_kernel(PcgScalars* sc, PcgHostStatus* status, cudaGraphConditionalHandle handle, int use_handle) {
    pdl_wait();
    pdl_trigger();
    const int done = sc->done;
    if (!done) {
        sc->sigma = sc->sigma_new;
        sc->it = sc->it + 1;
        status->it_done = sc->it;
        __threadfence_system();
    }
    if (use_handle) cudaGraphSetConditional(handle, done ? 0u : 1u);
}

// first evaluation of the loop condition (graph root, before the WHILE node)
__global__ void loop_condition_kernel(cudaGraphConditionalHandle handle, const PcgScalars* sc) {
    cudaGraphSetConditional(handle, sc->done ? 0u : 1u);
}

}  // namespace

// one PCG iteration: SpMV -> update -> (multigrid cycle -> z.r | fused diagonal) -> direction -> close
static int enqueue_iteration(fsim* h, const PcgArgs& a, bool vec, bool use_mg, int nbv, cudaGraphConditionalHandle handle = 0, int use_handle = 0) {
    const int mode = h->dist ? 1 : (h->hybrid ? 2 : 0);  // 1: slab-local solve, 2: full-grid context restricted to the owned planes
    // (the three reductions of the fused four-cells-per-thread chain are finished inside the reducing kernels when a.ar is set)
    auto AR = [&](int kind) { return (mode && !(a.ar && (kind == AR_SPMV || kind == AR_UPDATE || kind == AR_DOTZR))) ? dist_allreduce(h, kind, true) : FSIM_OK; };
    // the neighbours' boundary planes of the search direction
    if (mode == 1) { int rc = dist_halo(h, HALO_S, true); if (rc) return rc; }
    if (mode == 2 && !h->fx_on) { int rc = dist_halo_sym(h, SYM_S, h->s, true); if (rc) return rc; }  // (fused: direction4 -> spmv4)
    {
        KScope ks(h, K_SPMV);
        if (vec && a.g.gx % 4 == 0 && a.g.nc % 4 == 0) launch_k(h, spmv4_kernel, dim3(tile4_blocks(a.t4)), dim3(PT), 0, a);
        else if (vec) launch_k(h, spmv_kernel<true>, dim3(nbv), dim3(PT), 0, a);
        else launch_k(h, spmv_kernel<false>, dim3(nbv), dim3(PT), 0, a);
    }
    { int rc = AR(AR_SPMV); if (rc) return rc; }
    if (use_mg && mg_can_fuse(h)) {
        // update fused with the cycle's first sweep, z.r fused with its last sweep: 2 passes and 2 launches fewer
        int rc = mg_update_first(h);
        if (rc) return rc;
        rc = AR(AR_UPDATE);  // (slab modes: max |r| over the ranks + the convergence decision)
        if (rc) return rc;
        rc = mg_apply(h, true, true);
        if (rc) return rc;
        if (a.z32 != h->mg_z32) return fsim_fail(h, FSIM_ERR_INVALID, "multigrid result buffer moved between cycles");
        rc = AR(AR_DOTZR);
        if (rc) return rc;
    } else if (use_mg) {
        { KScope ks(h, K_UPDATE); launch_k(h, update_kernel<false>, dim3(nbv), dim3(PT), 0, a); }
        int rc = AR(AR_UPDATE);
        if (rc) return rc;
        rc = mg_apply(h, false, false);
        if (rc) return rc;
        if (a.z32 != h->mg_z32) return fsim_fail(h, FSIM_ERR_INVALID, "multigrid result buffer moved between cycles");
        { KScope ks(h, K_UPDATE); launch_k(h, dot_zr_kernel, dim3(nbv), dim3(PT), 0, a); }
        rc = AR(AR_DOTZR);
        if (rc) return rc;
    } else {
        { KScope ks(h, K_UPDATE); launch_k(h, update_kernel<true>, dim3(nbv), dim3(PT), 0, a); }
        int rc = AR(AR_UPDATE_JACOBI);
        if (rc) return rc;
    }
    {
        KScope ks(h, K_DIRECTION, 2);
        if (vec && a.z32 && a.g.gx % 4 == 0 && a.g.nc % 4 == 0 && a.cb % 4 == 0) launch_k(h, direction4_kernel, dim3(tile4_blocks(a.t4p)), dim3(PT), 0, a);
        else launch_k(h, direction_kernel, dim3(nbv), dim3(PT), 0, a);
        launch_k(h, close_kernel, dim3(1), dim3(1), 0, h->scal, h->status_dev, handle, use_handle);
    }
    return FSIM_OK;
}

int k_project(fsim* h, double dt, int* iterations) {
    if (h->par.solver_type == FSIM_SOLVER_BASIC) return k_project_basic(h, iterations);
    const GridDims& g = h->g;
    PcgArgs a;
    a.g = g; a.flags = h->flags; a.code = h->code; a.code_mg = h->code_mg != h->code ? h->code_mg : nullptr; a.code_full = h->code_full;
    for (int ax = 0; ax < 3; ax++) a.u2[ax] = h->u2[ax];
    a.dens = h->dens;
    a.p = h->p; a.rhs = h->rhs; a.r = h->r; a.s = h->s; a.q = h->q; a.z = h->z;
    a.sc = h->scal; a.partials = h->partials; a.counter = h->red_counter;
    a.status = h->status_dev;
    const double hcell = h->info.cell_d[0];
    a.inv_h = 1.0 / hcell;
    a.avg_pressure = h->par.average_pressure; a.pressure_k = h->par.pressure_k;
    a.pressure_enabled = h->par.pressure_enabled;
    a.warm = (h->warm_start && h->pressure_valid) ? (h->warm_extrapolate && h->warm_history >= 2 ? 2 : 1) : 0;
    if (h->par.max_iterations <= 0) a.warm = 0;  // zero iterations apply p = 0 like the reference, not last step's pressure
    a.p_prev = h->warm_extrapolate ? h->p_prev : nullptr;
    { const char* e = getenv("FSIM_SLAB_CUT"); a.cut_neumann = (e && e[0] == 'n') ? 1 : 0; }
    a.z32 = nullptr;
    a.cb = h->hybrid ? (int64_t)g.zown0 * g.sz : 0;
    a.ce = h->hybrid ? (int64_t)g.zown1 * g.sz : g.nc;
    const int nbv = div_up(a.ce - a.cb, CHUNK);     // chunked kernels
    a.t4 = tile4_make(g.gx, a.cb / g.gx, (a.ce - a.cb) / g.gx, g.gy);
    a.t4p = a.t4;
    const bool vec = (g.gx % 2 == 0) && (g.nc % 2 == 0);
    const bool use_mg = mg_enabled(h);
    // hybrid slab projection: the in-loop exchanges ride inside the solver kernels when the whole four-cells-per-thread chain runs
    h->fx_on = false;
    memset(&a.fxs, 0, sizeof(a.fxs));
    memset(&a.fws, 0, sizeof(a.fws));
    if (h->hybrid && use_mg && mg_can_fuse(h) && vec && g.gx % 4 == 0 && g.nc % 4 == 0 && a.cb % 4 == 0 && dist_fx(h, SYM_S, h->s, &a.fxs, &a.fws)) {
        h->fx_on = true;
        a.t4p = tile4_boundary_first(a.t4);
        a.t4 = tile4_boundary_last(a.t4);
        a.fxs.nblk[0] = a.fxs.nblk[1] = (uint32_t)(a.t4.nbx * (g.gy / T4_ROWS));
    }
    a.ar = h->fx_on ? dist_ar_dev(h) : nullptr;
    h->ar_dev = a.ar;
    const int max_it = h->par.max_iterations;

    // solve parameters live on the device: the captured iteration graph stays valid when dt / tolerance change
    PcgScalars* sh = h->scal_host;
    memset(sh, 0, sizeof(*sh));
    sh->scale = dt / (h->par.fluid_density * hcell * hcell);
    sh->inv_scale = 1.0 / sh->scale;
    sh->tol = h->par.residual_tolerance;
    sh->max_it = max_it;
    sh->dist = (h->dist || h->hybrid) ? 1 : 0;
    h->status_host->done = 0;
    h->status_host->it_done = 0;
    FSIM_CUDA(h, cudaMemcpyAsync(h->scal, sh, sizeof(PcgScalars), cudaMemcpyHostToDevice, h->stream));

    const int mode = h->dist ? 1 : (h->hybrid ? 2 : 0);
    const bool dist = mode != 0;
    a.rc0 = 0; a.rc1 = g.nc;
    {
        // hybrid slab projection: the owned planes get everything, the planes of the other ranks their stencil codes only
        const int nother = g.gz - (g.zown1 - g.zown0);
        const bool split = h->hybrid && h->code_full && g.gx % 4 == 0;
        if (split) { a.rc0 = a.cb; a.rc1 = a.ce; }
        KScope ks(h, K_RHS, (split && nother > 0) ? 2 : 1);
        if (split && nother > 0) codes_full4_kernel<<<dim3(div_up(g.gx, 128), div_up(g.gy, 8), nother), 256, 0, h->stream>>>(g, h->flags, h->code_full, nother);
        rhs_kernel<<<div_up(a.rc1 - a.rc0, PT), PT, 0, h->stream>>>(a);  // one cell per thread (a chunked loop was 25 % slower)
    }
    if (dist) { int rc = dist_allreduce(h, AR_RHS, false); if (rc) return rc; }
    if (a.warm) {
        // the initial guess of the neighbours' boundary planes
        if (mode == 1) { int rc = dist_halo(h, HALO_P, true); if (rc) return rc; }
        if (mode == 2) { int rc = dist_halo_sym(h, SYM_P, h->p, true); if (rc) return rc; }
        { KScope ks(h, K_RHS); residual_kernel<<<nbv, PT, 0, h->stream>>>(a); }
        if (dist) { int rc = dist_allreduce(h, AR_RESIDUAL, true); if (rc) return rc; }
    }
    if (use_mg) {
        int rc = mg_build(h);
        if (rc) return rc;
        rc = mg_apply(h, false, false);
        if (rc) return rc;
        a.z32 = h->mg_z32;
        { KScope ks(h, K_PCG_INIT); start_kernel<false><<<nbv, PT, 0, h->stream>>>(a); }
    } else {
        KScope ks(h, K_PCG_INIT);
        start_kernel<true><<<nbv, PT, 0, h->stream>>>(a);
    }
    if (dist) { int rc = dist_allreduce(h, AR_START, true); if (rc) return rc; }
    // fused exchanges: the first search direction comes from start_kernel, which has no fused store -- one stand-alone push
    if (h->fx_on) { int rc = dist_halo_sym(h, SYM_S, h->s, true); if (rc) return rc; }
    FSIM_CHECK_LAUNCH(h);

    // The whole PCG loop runs on the device: a CUDA graph whose body (one iteration, captured once -- every argument is a
    // fixed device pointer and all solve parameters live in device memory) sits inside a conditional WHILE node; a
    // one-thread kernel re-arms the condition from the done flag.  The host launches it once and never polls.
    bool graph = h->use_graph && h->prof_mask == 0;
    if (graph && !h->pcg_graph && !h->pcg_graph_failed) {
        auto build = [&]() -> bool {
            const int sticky0 = h->sticky;
            cudaGraph_t gr = nullptr, body = nullptr;
            cudaGraphConditionalHandle handle;
            cudaGraphNode_t set_node, cond_node;
            bool ok = cudaGraphCreate(&gr, 0) == cudaSuccess &&
                      cudaGraphConditionalHandleCreate(&handle, gr, 1, cudaGraphCondAssignDefault) == cudaSuccess;
            if (ok) {
                void* kargs[2] = {(void*)&handle, (void*)&h->scal};
                cudaKernelNodeParams kp = {};
                kp.func = (void*)loop_condition_kernel;
                kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.sharedMemBytes = 0; kp.kernelParams = kargs; kp.extra = nullptr;
                ok = cudaGraphAddKernelNode(&set_node, gr, nullptr, 0, &kp) == cudaSuccess;
            }
            if (ok) {
                cudaGraphNodeParams cp = {};
                cp.type = cudaGraphNodeTypeConditional;
                cp.conditional.handle = handle;
                cp.conditional.type = cudaGraphCondTypeWhile;
                cp.conditional.size = 1;
                ok = cudaGraphAddNode(&cond_node, gr, &set_node, 1, &cp) == cudaSuccess;
                if (ok) body = cp.conditional.phGraph_out[0];
            }
            if (ok) {
                const int64_t l0 = h->launches;
                int64_t c0[K_COUNT];
                memcpy(c0, h->launch_n, sizeof(c0));
                ok = cudaStreamBeginCaptureToGraph(h->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                if (ok) {
                    const int rc = enqueue_iteration(h, a, vec, use_mg, nbv, handle, 1);
                    cudaGraph_t dummy = nullptr;
                    const cudaError_t e = cudaStreamEndCapture(h->stream, &dummy);
                    ok = rc == FSIM_OK && e == cudaSuccess;
                }
                h->pcg_graph_launches = (int)(h->launches - l0);
                for (int k = 0; k < K_COUNT; k++) { h->pcg_graph_class[k] = (int)(h->launch_n[k] - c0[k]); h->launch_n[k] = c0[k]; }
                h->launches = l0;  // capture does not execute
            }
            if (ok) ok = cudaGraphInstantiate(&h->pcg_graph, gr, 0) == cudaSuccess;
            if (gr) cudaGraphDestroy(gr);
            if (!ok) { h->pcg_graph = nullptr; cudaGetLastError(); h->sticky = sticky0; }
            return ok;
        };
        bool ok = build();
        if (!ok) {  // older driver / unsupported node: fall back to enqueueing iterations from the host
            fprintf(stderr, "libfsim_b200: device-side PCG loop graph unavailable; using the host-driven loop\n");
            h->pcg_graph = nullptr;
            h->pcg_graph_failed = true;
        }
    }
    graph = graph && h->pcg_graph;
    if (graph) {
        FSIM_CUDA(h, cudaGraphLaunch(h->pcg_graph, h->stream));
    } else {
        // host-driven loop: polls the host-mapped status word (no stream sync), at most two iterations in flight;
        // kernels of an iteration enqueued after convergence see done != 0 and exit immediately
        for (int it = 0; it < max_it; it++) {
            if (h->status_host->done) break;
            int rc = enqueue_iteration(h, a, vec, use_mg, nbv);
            if (rc) return rc;
            while (!h->status_host->done && it + 1 - h->status_host->it_done > 1) {
                if (cudaStreamQuery(h->stream) != cudaErrorNotReady) break;  // drained (or failed): flags are final
            }
        }
    }
    FSIM_CHECK_LAUNCH(h);
    publish_kernel<<<1, 1, 0, h->stream>>>(h->scal, h->result_dev);
    h->launches++;
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    const PcgScalars s = *h->result_host;
    if (graph) {  // the device decided how many iterations ran: account their launches now
        const int ran = s.early_out ? 0 : s.it + ((s.done == 1 || s.nan_break) ? 1 : 0);
        h->launches += (int64_t)h->pcg_graph_launches * ran + 1;
        for (int k = 0; k < K_COUNT; k++) h->launch_n[k] += (int64_t)h->pcg_graph_class[k] * ran;
    }
    h->solve.iterations = s.early_out ? 0 : (s.done ? s.iterations : max_it);
    h->solve.early_out = s.early_out;
    h->solve.rhs_sumsq = s.rhs_sumsq;
    h->solve.residual_max = s.rmax;
    h->solve.fluid_cells = s.fluid_cells;
    if (iterations) *iterations = h->solve.iterations;
    h->pressure_valid = !s.early_out;
    h->warm_history = s.early_out ? 0 : h->warm_history + 1;
    if (!s.early_out) {  // the early-out returns before applying anything (:257-258)
        if (mode == 2) { int rc = dist_halo_sym(h, SYM_P, h->p, false); if (rc) return rc; }  // pressure of the neighbours' boundary planes
        if (h->skip_apply) return FSIM_OK;  // full-grid solver context of a slab handle: the slab applies its own part
        if (mode == 1) { int rc = dist_halo(h, HALO_P, false); if (rc) return rc; }
        return k_pressure_apply(h, dt);
    }
    return FSIM_OK;
}
