// mg.cu -- parallel preconditioner of the pressure PCG: one aggregation-multigrid cycle, z = M^-1 r.
//
// Replaces the reference's strictly sequential MIC(0) factor + triangular solves
// (BridsonSolverGrid::calculatePreconditioner / applyPreconditioner, bridsonSolverGrid.cpp:91-163); BASELINE.json's
// north_star allows any SPD preconditioner because equivalence is defined on the converged pressure field.
//
//  * hierarchy: 2x2x2 cell aggregation, piecewise-constant prolongation P, restriction P^T, Galerkin coarse operators
//    A_c = P^T A P.  For the 7-point matrix of calculateAMatrix (:40-77) the Galerkin operator is again a 7-point
//    stencil whose face weight is the NUMBER of WATER-WATER fine connections across the coarse face and whose diagonal
//    adds the WATER-AIR (Dirichlet) connections -- free surfaces, solid walls and obstacles are represented exactly on
//    every level, no geometric heuristics.  Level 0 is matrix-free (1-byte flags); levels >= 1 store 4 floats per cell.
//  * smoother: Jacobi with 2-step Chebyshev weights (pre: 1.389, 0.5617; post: reversed) => symmetric cycle.
//  * plain aggregation under-corrects smooth modes by ~2x; the coarse correction is scaled by 1.8 and level 2 is
//    visited twice (W recursion there only) -- tools/mg_prototype.py: 8-10 PCG iterations at 128^3..256^3 where
//    MIC(0) needs 80..155.
//  * fp32 throughout (it only has to be an approximate inverse); the CG recurrence around it is fp64 (pcg.cu).
// Level-0 kernels process 4 cells per thread (float4 / ushort4); small levels run in one thread-block-cluster launch.
#include <algorithm>
#include <cooperative_groups.h>

#include <memory>

#include "fsim_internal.h"
#include "reduce.cuh"
#include "dist_dev.cuh"
#include "fexch.cuh"
#include "launch.cuh"
#include "tile4.cuh"

namespace cg = cooperative_groups;

struct MgLevel {
    int gx, gy, gz, sy, sz;
    int64_t nc;
    size_t pad;                     // zeroed halo on both ends of every array of this level (>= sz)
    float *wx, *wy, *wz, *diag;     // + face weights and diagonal (levels >= 1)
    float *xa, *xb, *b;             // ping-pong iterate and right-hand side
    float* base[7];                 // raw allocations (for cudaFree)
};

namespace {

constexpr float OMEGA = 0.8f;                    // coarsest-level Jacobi
constexpr float OM_A = 1.389f, OM_B = 0.5617f;   // 2-step Chebyshev weights for D^-1 A on [0.5, 2]: pre-sweeps (A, B), post-sweeps (B, A)
constexpr float OVER = 1.8f;
constexpr int PRE = 2, POST = 2;
constexpr int W_FIRST = 2, W_LAST = 2; // levels W_FIRST..W_LAST are visited twice per visit of their parent
constexpr int COARSE_SWEEPS = 30;    // damped Jacobi sweeps on the coarsest level (12 is 0.05 ms/step cheaper at 256^3 but costs iterations elsewhere: 128^3 cold 8 -> 10)
constexpr int COARSE_MAX = 1024;     // the coarsest level fits one CTA

struct Lv {  // kernel view of a level
    int gx, gy, gz, sy, sz;
    const float *wx, *wy, *wz, *diag;
    const uint8_t* flags;   // level 0 only (hierarchy build)
    const uint16_t* code;   // level 0 only: stencil codes (fsim_internal.h CODE_ACTIVE)
    int z0, z1;             // planes the per-cell kernels visit: [0, gz), or the planes this rank owns (hybrid slab projection, level 0)
};

__device__ __forceinline__ bool cell_of(const Lv& L, int& x, int& y, int& z, int64_t& c) {
    x = blockIdx.x * blockDim.x + threadIdx.x;
    y = blockIdx.y * blockDim.y + threadIdx.y;
    z = blockIdx.z * blockDim.z + threadIdx.z + L.z0;
    c = ((int64_t)z * L.gy + y) * L.gx + x;
    return x < L.gx && y < L.gy && z < L.z1;
}

// (diagonal, sum of w_nbr * x_nbr) of row c.  FINE: from the cell flags (weights 1 to WATER neighbours, diagonal =
// #non-solid neighbours); coarse: from the stored Galerkin weights (zero across the domain boundary, arrays padded).
template <bool FINE>
__device__ __forceinline__ float row(const Lv& L, int64_t c, const float* __restrict__ x, float* offsum) {
    if (FINE) {
        const unsigned cd = L.code[c];
        float s = 0.f;
        if (cd & 1u) s += x[c - 1];
        if (cd & 2u) s += x[c + 1];
        if (cd & 4u) s += x[c - L.sy];
        if (cd & 8u) s += x[c + L.sy];
        if (cd & 16u) s += x[c - L.sz];
        if (cd & 32u) s += x[c + L.sz];
        *offsum = s;
        return (float)((cd >> 6) & 7u);
    } else {
        *offsum = L.wx[c] * x[c + 1] + L.wx[c - 1] * x[c - 1] + L.wy[c] * x[c + L.sy] + L.wy[c - L.sy] * x[c - L.sy] +
                  L.wz[c] * x[c + L.sz] + L.wz[c - L.sz] * x[c - L.sz];
        return L.diag[c];
    }
}

template <bool FINE>
__device__ __forceinline__ bool active(const Lv& L, int64_t c) {
    return FINE ? ((L.code[c] & CODE_ACTIVE) != 0) : (L.diag[c] > 0.f);
}

// first sweep from a zero guess: x = omega * b / diag.  FINE also converts the fp64 CG residual: b = r / scale.
template <bool FINE>
__global__ void __launch_bounds__(256) mg_first_kernel(Lv L, const double* __restrict__ r64, const PcgScalars* __restrict__ sc,
                                                        float* __restrict__ b, float* __restrict__ xout) {
    pdl_wait();
    pdl_trigger();
    int x, y, z; int64_t c;
    if (sc->done) return;
    const double inv_scale = sc->inv_scale;
    if (!cell_of(L, x, y, z, c)) return;
    float v = 0.f;
    if (active<FINE>(L, c)) {
        float bb;
        if (FINE) { bb = (float)(r64[c] * inv_scale); b[c] = bb; } else bb = b[c];
        const float d = FINE ? (float)((L.code[c] >> 6) & 7u) : L.diag[c];
        v = d > 0.f ? OM_A * bb / d : 0.f;
    } else if (FINE) b[c] = 0.f;
    xout[c] = v;
}

// damped Jacobi sweep: xout = xin + omega * (b - A xin) / diag
template <bool FINE>
__global__ void __launch_bounds__(256) mg_jacobi_kernel(Lv L, const PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xin,
                                                         float* __restrict__ xout, float om) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    int x, y, z; int64_t c;
    if (!cell_of(L, x, y, z, c)) return;
    float v = 0.f;
    if (active<FINE>(L, c)) {
        float off;
        const float d = row<FINE>(L, c, xin, &off);
        const float xi = xin[c];
        v = d > 0.f ? xi + om * (b[c] - (d * xi - off)) / d : 0.f;
    }
    xout[c] = v;
}

// two pre-smoothing sweeps from a zero guess in one pass (coarse levels): x1 = omega b / d needs no neighbours, so
// x2 = x1 + omega (b - A x1) / d only reads b and d at the 7 stencil points.  All 20 loads are issued at once (the arrays
// are padded by a plane of zeros, a zero weight masks an absent neighbour): one L2 round trip instead of three dependent
// ones -- these levels are pure latency.
template <class I>
__device__ __forceinline__ float pre2_value(const Lv& L, const float* __restrict__ b, I c) {
    const I nb[6] = {c - 1, c + 1, c - L.sy, c + L.sy, c - L.sz, c + L.sz};
    const float w[6] = {L.wx[c - 1], L.wx[c], L.wy[c - L.sy], L.wy[c], L.wz[c - L.sz], L.wz[c]};
    const float d = L.diag[c], bb = b[c];
    float dn[6], bn[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { dn[k] = L.diag[nb[k]]; bn[k] = b[nb[k]]; }
    float off = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (w[k] > 0.f && dn[k] > 0.f) off += w[k] * (OM_A * bn[k] / dn[k]);
    if (!(d > 0.f)) return 0.f;
    const float xi = OM_A * bb / d;
    return xi + OM_B * (bb - (d * xi - off)) / d;
}
__global__ void __launch_bounds__(256) mg_pre2_kernel(Lv L, const PcgScalars* __restrict__ sc, const float* __restrict__ b,
                                                      float* __restrict__ xout, FxWait gw) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    if (gw.has[0]) fx_wait(gw, 0);  // (hybrid slab projection, fused all-rank push: every rank's planes of b have arrived)
    int x, y, z; int64_t c;
    if (!cell_of(L, x, y, z, c)) return;
    xout[c] = pre2_value(L, b, c);
}

// coarse right-hand side: b_c(I) = sum over the 2x2x2 children of (b - A x)   (restriction = P^T)
template <bool FINE>
__global__ void __launch_bounds__(256) mg_restrict_kernel(Lv L, Lv C, const PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xf,
                                                           float* __restrict__ bc) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    int X, Y, Z; int64_t cc;
    if (!cell_of(C, X, Y, Z, cc)) return;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int x = 2 * X + i, y = 2 * Y + j, z = 2 * Z + k;
                if (x >= L.gx || y >= L.gy || z >= L.gz) continue;
                const int64_t c = ((int64_t)z * L.gy + y) * L.gx + x;
                if (!active<FINE>(L, c)) continue;
                float off;
                const float d = row<FINE>(L, c, xf, &off);
                s += b[c] - (d * xf[c] - off);
            }
    bc[cc] = s;
}

// the same restriction for the stored-operator levels with one FINE cell per thread (see t_restrict below): eight lanes per
// coarse cell, three shuffles
__global__ void __launch_bounds__(256) mg_restrict8_kernel(Lv L, Lv C, const PcgScalars* __restrict__ sc, const float* __restrict__ b,
                                                            const float* __restrict__ xf, float* __restrict__ bc, int ncc) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    const int g = blockIdx.x * 256 + threadIdx.x;
    const int cc = g >> 3, sub = g & 7;
    float r = 0.f;
    if (cc < ncc) {
        const int X = cc % C.gx, Y = (cc / C.gx) % C.gy, Z = cc / (C.gx * C.gy);
        const int x = 2 * X + (sub & 1), y = 2 * Y + ((sub >> 1) & 1), z = 2 * Z + (sub >> 2);
        if (x < L.gx && y < L.gy && z < L.gz) {
            const int64_t c = ((int64_t)z * L.gy + y) * L.gx + x;
            if (active<false>(L, c)) {
                float off;
                const float d = row<false>(L, c, xf, &off);
                r = b[c] - (d * xf[c] - off);
            }
        }
    }
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    if (sub == 0 && cc < ncc) bc[cc] = r;
}

// prolongation + over-corrected update fused with the first post-smoothing sweep:
//   xc = xin + OVER * e_c[parent] ;  xout = xc + omega * (b - A xc) / diag
template <bool FINE>
__global__ void __launch_bounds__(256) mg_prolong_jacobi_kernel(Lv L, Lv C, const PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xin,
                                                                 const float* __restrict__ ec, float* __restrict__ xout) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    int x, y, z; int64_t c;
    if (!cell_of(L, x, y, z, c)) return;
    float v = 0.f;
    auto xc = [&](int xx, int yy, int zz, int64_t cn) -> float {
        const int64_t pc = ((int64_t)(zz >> 1) * C.gy + (yy >> 1)) * C.gx + (xx >> 1);
        return xin[cn] + OVER * ec[pc];
    };
    if (FINE) {
        if (active<FINE>(L, c)) {
            float off = 0.f;
            const unsigned cd = L.code[c];
            if (cd & 1u) off += xc(x - 1, y, z, c - 1);
            if (cd & 2u) off += xc(x + 1, y, z, c + 1);
            if (cd & 4u) off += xc(x, y - 1, z, c - L.sy);
            if (cd & 8u) off += xc(x, y + 1, z, c + L.sy);
            if (cd & 16u) off += xc(x, y, z - 1, c - L.sz);
            if (cd & 32u) off += xc(x, y, z + 1, c + L.sz);
            const float d = (float)((cd >> 6) & 7u);
            const float xi = xc(x, y, z, c);
            v = d > 0.f ? xi + OM_B * (b[c] - (d * xi - off)) / d : 0.f;
        }
    } else {
        // first batch: everything that does not depend on the weights (the arrays are padded: all addresses valid)
        const float d = L.diag[c], bb = b[c], xi = xc(x, y, z, c);
        const float w0 = L.wx[c - 1], w1 = L.wx[c], w2 = L.wy[c - L.sy], w3 = L.wy[c], w4 = L.wz[c - L.sz], w5 = L.wz[c];
        if (d > 0.f) {
            float off = 0.f;
            if (w0 > 0.f) off += w0 * xc(x - 1, y, z, c - 1);
            if (w1 > 0.f) off += w1 * xc(x + 1, y, z, c + 1);
            if (w2 > 0.f) off += w2 * xc(x, y - 1, z, c - L.sy);
            if (w3 > 0.f) off += w3 * xc(x, y + 1, z, c + L.sy);
            if (w4 > 0.f) off += w4 * xc(x, y, z - 1, c - L.sz);
            if (w5 > 0.f) off += w5 * xc(x, y, z + 1, c + L.sz);
            v = xi + OM_B * (bb - (d * xi - off)) / d;
        }
    }
    xout[c] = v;
}

// ---- level-0 kernels, four consecutive x cells per thread (gx % 4 == 0): 16-byte loads of the iterate, 8-byte loads of
// the codes; the x neighbours of the inner cells come from registers, only the two outer ones are extra loads ----------
struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float* p) { const float4 t = *reinterpret_cast<const float4*>(p); F4 r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }
__device__ __forceinline__ void st4(float* p, const F4& r) { *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]); }
__device__ __forceinline__ F4 zero4() { F4 r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.f; return r; }

struct Stencil4 {  // the 7-point neighbourhood of four consecutive cells
    F4 c, ym, yp, zm, zp;
    float xl, xr;
};
// off-diagonal sum of cell i (0..3) of the group
__device__ __forceinline__ float off4(const Stencil4& s, int i, unsigned cd) {
    float o = 0.f;
    if (cd & 1u) o += i == 0 ? s.xl : s.c.v[i - 1];
    if (cd & 2u) o += i == 3 ? s.xr : s.c.v[i + 1];
    if (cd & 4u) o += s.ym.v[i];
    if (cd & 8u) o += s.yp.v[i];
    if (cd & 16u) o += s.zm.v[i];
    if (cd & 32u) o += s.zp.v[i];
    return o;
}
// z index of the CTA: blockIdx.z rotated by zrot (fused slab exchanges visit the boundary planes first or last)
__device__ __forceinline__ int zblock_of(int zrot) {
    int zb = blockIdx.z + zrot;
    if (zb >= (int)gridDim.z) zb -= gridDim.z;
    return zb;
}
__device__ __forceinline__ bool group_of(const Lv& L, int64_t& c, unsigned cd[4], int zblk) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y, z = zblk * blockDim.z + threadIdx.z + L.z0;
    if (x >= L.gx || y >= L.gy || z >= L.z1) return false;
    c = ((int64_t)z * L.gy + y) * L.gx + x;
    const ushort4 t = *reinterpret_cast<const ushort4*>(L.code + c);
    cd[0] = t.x; cd[1] = t.y; cd[2] = t.z; cd[3] = t.w;
    return true;
}
// fused exchanges in the kernels whose CTAs cover `planes` planes from plane zlo on (fexch.cuh).  All threads of the CTA call these.
__device__ __forceinline__ void fx_wait_planes(const FxWait& w, int zlo, int planes) {
    if (!(w.has[0] | w.has[1])) return;
#pragma unroll
    for (int side = 0; side < 2; side++)
        if (w.has[side] && w.zb[side] >= zlo && w.zb[side] < zlo + planes) fx_wait(w, side);
}
__device__ __forceinline__ void fx_signal_planes(const FxPush& f, int zlo, int planes) {
    if (!(f.peer[0] || f.peer[1])) return;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) fx_expect(f);
#pragma unroll
    for (int side = 0; side < 2; side++)
        if (f.peer[side] && f.zb[side] >= zlo && f.zb[side] < zlo + planes) fx_signal(f, side);
}
// the neighbour's array when plane z is one of my boundary planes and a neighbour reads it
__device__ __forceinline__ float* fx_peer_of(const FxPush& f, int z) { return fx_peer(f, fx_side(f.zb, z)); }

__device__ __forceinline__ Stencil4 load_stencil4(const Lv& L, const float* __restrict__ x, int64_t c, const unsigned cd[4]) {
    Stencil4 s;
    const unsigned any = cd[0] | cd[1] | cd[2] | cd[3];
    s.c = ld4(x + c);
    s.ym = (any & 4u) ? ld4(x + c - L.sy) : zero4();
    s.yp = (any & 8u) ? ld4(x + c + L.sy) : zero4();
    s.zm = (any & 16u) ? ld4(x + c - L.sz) : zero4();
    s.zp = (any & 32u) ? ld4(x + c + L.sz) : zero4();
    s.xl = (cd[0] & 1u) ? x[c - 1] : 0.f;
    s.xr = (cd[3] & 2u) ? x[c + 4] : 0.f;
    return s;
}

__global__ void __launch_bounds__(256) mg_first4_kernel(Lv L, const double* __restrict__ r64, const PcgScalars* __restrict__ sc,
                                                        float* __restrict__ b, float* __restrict__ xout, FxPush fp, int zrot) {
    pdl_wait();
    pdl_trigger();
    int64_t c; unsigned cd[4];
    if (sc->done) return;
    const double inv_scale = sc->inv_scale;
    const int zblk = zblock_of(zrot);
    if (group_of(L, c, cd, zblk) && ((cd[0] | cd[1] | cd[2] | cd[3]) & CODE_ACTIVE)) {
        F4 bb = zero4(), xo = zero4();
        const double2 r0 = *reinterpret_cast<const double2*>(r64 + c), r1 = *reinterpret_cast<const double2*>(r64 + c + 2);
        const double rr[4] = {r0.x, r0.y, r1.x, r1.y};
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (cd[i] & CODE_ACTIVE) {
                bb.v[i] = (float)(rr[i] * inv_scale);
                const float d = (float)((cd[i] >> 6) & 7u);
                xo.v[i] = d > 0.f ? OM_A * bb.v[i] / d : 0.f;
            }
        st4(b + c, bb);     // groups without WATER cells are never read (every consumer checks the code) => not written
        st4(xout + c, xo);
        if (float* peer = fx_peer_of(fp, zblk * 2 + (int)threadIdx.z + L.z0)) st4(peer + c, xo);
    }
    fx_signal_planes(fp, zblk * 2 + L.z0, 2);
}

__global__ void __launch_bounds__(256) mg_jacobi4_kernel(Lv L, PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xin,
                                                         float* __restrict__ xout, float om, FxWait fw, FxPush fp, int zrot) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    const int zblk = zblock_of(zrot);
    fx_wait_planes(fw, zblk * 2 + L.z0, 2);  // (hybrid slab projection: the neighbours' boundary planes of xin)
    int64_t c; unsigned cd[4];
    if (group_of(L, c, cd, zblk) && ((cd[0] | cd[1] | cd[2] | cd[3]) & CODE_ACTIVE)) {
        F4 xo = zero4();
        const Stencil4 s = load_stencil4(L, xin, c, cd);
        const F4 bb = ld4(b + c);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (cd[i] & CODE_ACTIVE) {
                const float d = (float)((cd[i] >> 6) & 7u);
                const float xi = s.c.v[i];
                xo.v[i] = d > 0.f ? xi + om * (bb.v[i] - (d * xi - off4(s, i, cd[i]))) / d : 0.f;
            }
        st4(xout + c, xo);
        if (float* peer = fx_peer_of(fp, zblk * 2 + (int)threadIdx.z + L.z0)) st4(peer + c, xo);
    }
    fx_signal_planes(fp, zblk * 2 + L.z0, 2);
}

// The last sweep of the cycle also accumulates sigma' = z.r = scale * z.b (fp64 accumulation; saves a pass over z and r.
// The fp32 rounding of r inside b perturbs sigma' by ~1e-8 relative: it only enters alpha and beta, never p or r).
// Persistent form: ~8 CTAs per SM walk the (32,4,2)-thread tiles of the grid, so the grid-wide reduction sees ~1.2 K
// partials instead of 16 K -- with one CTA per tile its fence + arrival atomic per CTA cost 30 us on top of the 37 us sweep.
__global__ void __launch_bounds__(256) mg_jacobi4_dot_kernel(Lv L, PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xin,
                                                             float* __restrict__ xout, double* partials, unsigned int* counter, float om,
                                                             int ntx, int nty, int ntiles, FxWait fw, int trot, const ArDev* ar) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    double acc[1] = {0.0};
    const int tx = threadIdx.x & 31, ty = (threadIdx.x >> 5) & 3, tz = threadIdx.x >> 7;
    // tile index -> tile, rotated by trot tiles (fused slab exchanges: the layers next to the ghost planes come last)
    auto rot = [&](int t) -> int { const int u = t + trot; return u >= ntiles ? u - ntiles : u; };
    // the codes of the next tile are requested before this tile's stencil loads: one exposed round trip per tile instead of two
    auto tile_cell = [&](int t, int64_t& c) -> bool {
        t = rot(t);
        const int bx = t % ntx, by = (t / ntx) % nty, bz = t / (ntx * nty);
        const int x = (bx * 32 + tx) * 4, y = by * 4 + ty, z = bz * 2 + tz + L.z0;  // (hybrid slab projection: the owned planes only)
        c = ((int64_t)z * L.gy + y) * L.gx + x;
        return x < L.gx && y < L.gy && z < L.z1;
    };
    int64_t c = 0;
    ushort4 tc = make_ushort4(0, 0, 0, 0);
    if ((int)blockIdx.x < ntiles && tile_cell(blockIdx.x, c)) tc = *reinterpret_cast<const ushort4*>(L.code + c);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t cn = 0;
        ushort4 tcn = make_ushort4(0, 0, 0, 0);
        if (t + (int)gridDim.x < ntiles && tile_cell(t + gridDim.x, cn)) tcn = *reinterpret_cast<const ushort4*>(L.code + cn);
        const unsigned cd[4] = {tc.x, tc.y, tc.z, tc.w};
        const int64_t cc = c;
        tc = tcn; c = cn;
        if (fw.has[0] | fw.has[1]) fx_wait_planes(fw, (rot(t) / (ntx * nty)) * 2 + L.z0, 2);  // (CTA-uniform; a satisfied wait returns at once)
        if (!((cd[0] | cd[1] | cd[2] | cd[3]) & CODE_ACTIVE)) continue;  // (also every out-of-grid thread: its codes stay 0)
        F4 xo = zero4();
        const Stencil4 s = load_stencil4(L, xin, cc, cd);
        const F4 bb = ld4(b + cc);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (cd[i] & CODE_ACTIVE) {
                const float d = (float)((cd[i] >> 6) & 7u);
                const float xi = s.c.v[i];
                xo.v[i] = d > 0.f ? xi + om * (bb.v[i] - (d * xi - off4(s, i, cd[i]))) / d : 0.f;
                acc[0] += (double)xo.v[i] * (double)bb.v[i];
            }
        st4(xout + cc, xo);
    }
    double out[1] = {0.0};
    const int fr = grid_reduce_w<1, 0>(acc, partials, counter, out);
    if (fr && ar) {  // slab mode, fused reduction (dist_dev.cuh): sigma' = the all-rank sum
        const double m[4] = {out[0] * sc->scale, 0.0, 0.0, 0.0};
        ar_warp(ar, sc, nullptr, AR_DOTZR, m);
    } else if (fr == 2) {
        if (sc->dist) sc->loc[0] = out[0] * sc->scale;  // slab mode: finished by the all-rank reduction (AR_DOTZR)
        else sc->sigma_new = out[0] * sc->scale;
    }
}

constexpr int UF_MIN_BLOCKS = 4;  // 64 registers: four resident CTAs instead of three (ncu r2: 35 % active warps at 68 registers)
// CG update fused with the first smoothing sweep of the next cycle (level 0, linear chunks of 4-cell groups):
//   alpha = sigma / s.q ; p += alpha s ; r -= alpha q ; ||r||_inf ; b = r / scale ; x1 = omega b / diag
__global__ void __launch_bounds__(256, UF_MIN_BLOCKS) mg_update_first4_kernel(Lv L, Tile4 t4, PcgScalars* sc, PcgHostStatus* status, double* __restrict__ p,
                                                               const float* __restrict__ sv, double* __restrict__ r,
                                                               const double* __restrict__ q, float* __restrict__ b, float* __restrict__ xout,
                                                               double* partials, unsigned int* counter, FxPush fp, const ArDev* ar) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    // fused exchange of the first iterate: a CTA on a boundary plane stores its groups into the neighbour's ghost plane too
    float* xpeer = nullptr;
    int side = -1;
    if (fp.peer[0] || fp.peer[1]) {
        side = fx_side(fp.zb, tile4_plane(t4));
        xpeer = fx_peer(fp, side);
        if (blockIdx.x == 0 && threadIdx.x == 0) fx_expect(fp);
    }
    const double alpha = sc->sigma / sc->sq;
    const bool bad = alpha != alpha;  // NaN => the reference breaks before touching p (bridsonSolverGrid.cpp:271-272)
    const double inv_scale = sc->inv_scale;
    double acc[1] = {0.0};
    if (!bad) {
        // 2-D blocks (tile4.cuh); the stencil codes of the next trip are requested before this trip's vector loads
        int64_t c = 0, cn = 0;
        ushort4 t = make_ushort4(0, 0, 0, 0);
        if (tile4_cell(t4, 0, c)) t = *reinterpret_cast<const ushort4*>(L.code + c);
        for (int trip = 0; trip < T4_TRIPS; trip++, c = cn) {
            ushort4 tn = make_ushort4(0, 0, 0, 0);
            if (trip + 1 < T4_TRIPS && tile4_cell(t4, trip + 1, cn)) tn = *reinterpret_cast<const ushort4*>(L.code + cn);
            const unsigned cd[4] = {t.x, t.y, t.z, t.w};
            t = tn;
            if (!((cd[0] | cd[1] | cd[2] | cd[3]) & CODE_ACTIVE)) continue;
            double pp[4], ss[4], rr[4], qq[4];
            {
                const float4 sf = *reinterpret_cast<const float4*>(sv + c);
                ss[0] = (double)sf.x; ss[1] = (double)sf.y; ss[2] = (double)sf.z; ss[3] = (double)sf.w;
            }
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++) {
                const double2 a0 = *reinterpret_cast<const double2*>(p + c + 2 * h2);
                const double2 a2 = *reinterpret_cast<const double2*>(r + c + 2 * h2), a3 = *reinterpret_cast<const double2*>(q + c + 2 * h2);
                pp[2 * h2] = a0.x; pp[2 * h2 + 1] = a0.y;
                rr[2 * h2] = a2.x; rr[2 * h2 + 1] = a2.y; qq[2 * h2] = a3.x; qq[2 * h2 + 1] = a3.y;
            }
            F4 bb = zero4(), xo = zero4();
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (cd[i] & CODE_ACTIVE) {
                    pp[i] += alpha * ss[i];
                    rr[i] -= alpha * qq[i];
                    acc[0] = fmax(acc[0], fabs(rr[i]));
                    bb.v[i] = (float)(rr[i] * inv_scale);
                    const float d = (float)((cd[i] >> 6) & 7u);
                    xo.v[i] = d > 0.f ? OM_A * bb.v[i] / d : 0.f;
                }
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++) {
                *reinterpret_cast<double2*>(p + c + 2 * h2) = make_double2(pp[2 * h2], pp[2 * h2 + 1]);
                *reinterpret_cast<double2*>(r + c + 2 * h2) = make_double2(rr[2 * h2], rr[2 * h2 + 1]);
            }
            st4(b + c, bb);
            st4(xout + c, xo);
            if (xpeer) st4(xpeer + c, xo);
        }
    }
    if (xpeer) fx_signal(fp, side);
    double out[1] = {0.0};
    const int fr = grid_reduce_w<0, 1>(acc, partials, counter, out);
    if (fr && ar) {  // slab mode, fused reduction (dist_dev.cuh): max |r| over all ranks + the convergence decision
        const double m[4] = {0.0, 0.0, out[0], 0.0};
        ar_warp(ar, sc, status, AR_UPDATE, m);
    } else if (fr == 2) {
        if (sc->dist) {
            sc->loc[2] = out[0];  // slab mode: this rank's max |r|; dist_allreduce_kernel takes the decision (pcg_finish_update)
        } else {
            const int it = sc->it;
            if (bad) {
                sc->nan_break = 1; sc->done = 1; sc->iterations = it;
            } else {
                sc->rmax = out[0];
                if (out[0] < sc->tol) { sc->done = 1; sc->iterations = it; }                       // :280-281, 292
                else if (it + 1 >= sc->max_it) { sc->done = 2; sc->iterations = sc->max_it; }       // iteration cap
            }
            if (sc->done) { status->done = sc->done; __threadfence_system(); }
        }
    }
}

// one thread = two coarse cells in x = a 4x2x2 block of fine cells
__global__ void __launch_bounds__(256) mg_restrict4_kernel(Lv L, Lv C, const PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xf,
                                                           float* __restrict__ bc, int cz0, int cz1, FxWait fw, int zrot, FxAllPush gp) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    const int zblk = zblock_of(zrot);
    fx_wait_planes(fw, 2 * (zblk * 2 + cz0), 4);  // the CTA's two coarse planes have their children in four fine planes
    const int X = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y, Z = zblk * blockDim.z + threadIdx.z + cz0;
    // hybrid: only the coarse planes [cz0, cz1) whose children this rank owns; the others are written by their owners
    if (X < C.gx && Y < C.gy && Z < C.gz && Z < cz1) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int y = 2 * Y + j, z = 2 * Z + k;
                if (y >= L.gy || z >= L.gz) continue;
                const int64_t c = ((int64_t)z * L.gy + y) * L.gx + 2 * X;
                const ushort4 t = *reinterpret_cast<const ushort4*>(L.code + c);
                const unsigned cd[4] = {t.x, t.y, t.z, t.w};
                if (!((cd[0] | cd[1] | cd[2] | cd[3]) & CODE_ACTIVE)) continue;
                const Stencil4 s = load_stencil4(L, xf, c, cd);
                const F4 bb = ld4(b + c);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (cd[i] & CODE_ACTIVE) {
                        const float d = (float)((cd[i] >> 6) & 7u);
                        const float res = bb.v[i] - (d * s.c.v[i] - off4(s, i, cd[i]));
                        if (i < 2) s0 += res; else s1 += res;
                    }
            }
        const int64_t idx = ((int64_t)Z * C.gy + Y) * C.gx + X;
        *reinterpret_cast<float2*>(bc + idx) = make_float2(s0, s1);
        fx_all_store2(gp, idx, make_float2(s0, s1));  // (fused all-rank push: the same two cells in every other rank's array)
    }
    fx_all_signal(gp);
}

__global__ void __launch_bounds__(256) mg_prolong_jacobi4_kernel(Lv L, Lv C, const PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xin,
                                                                 const float* __restrict__ ec, float* __restrict__ xout, FxPush fp, int zrot) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    int64_t c; unsigned cd[4];
    const int zblk = zblock_of(zrot);
    const bool in = group_of(L, c, cd, zblk);
    F4 xo = zero4();
    const unsigned any = in ? (cd[0] | cd[1] | cd[2] | cd[3]) : 0u;
    if (any & CODE_ACTIVE) {
        const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
        const int y = blockIdx.y * blockDim.y + threadIdx.y, z = zblk * blockDim.z + threadIdx.z + L.z0;
        Stencil4 s = load_stencil4(L, xin, c, cd);
        // coarse corrections: the group's parents are (X0, X0+1) in row (y>>1, z>>1); neighbours use their own rows
        const int X0 = x >> 1;
        auto prow = [&](int yy, int zz) -> const float* { return ec + ((int64_t)(zz >> 1) * C.gy + (yy >> 1)) * C.gx; };
        auto add2 = [&](F4& v, const float* row) {
            const float2 e = *reinterpret_cast<const float2*>(row + X0);
            v.v[0] += OVER * e.x; v.v[1] += OVER * e.x; v.v[2] += OVER * e.y; v.v[3] += OVER * e.y;
        };
        const float* r0 = prow(y, z);
        add2(s.c, r0);
        if (any & 4u) add2(s.ym, prow(y - 1, z));
        if (any & 8u) add2(s.yp, prow(y + 1, z));
        if (any & 16u) add2(s.zm, prow(y, z - 1));
        if (any & 32u) add2(s.zp, prow(y, z + 1));
        if (cd[0] & 1u) s.xl += OVER * r0[X0 - 1];
        if (cd[3] & 2u) s.xr += OVER * r0[X0 + 2];
        const F4 bb = ld4(b + c);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (cd[i] & CODE_ACTIVE) {
                const float d = (float)((cd[i] >> 6) & 7u);
                const float xi = s.c.v[i];
                xo.v[i] = d > 0.f ? xi + OM_B * (bb.v[i] - (d * xi - off4(s, i, cd[i]))) / d : 0.f;
            }
        st4(xout + c, xo);
        if (float* peer = fx_peer_of(fp, z)) st4(peer + c, xo);
    }
    fx_signal_planes(fp, zblk * 2 + L.z0, 2);
}

// ---- stored-operator levels, four consecutive x cells per thread (gx % 4 == 0, level 1 of a 256^3 grid = 2 M cells) -----------
// One cell per thread these sweeps are ~7 waves of threads that each wait for one L2 round trip; with four cells per thread
// the same loads are in flight from a quarter of the threads (16-byte loads of the iterate, the right-hand side and the four
// weight arrays).  Per cell the arithmetic and its order are those of row<false> / pre2_value / mg_prolong_jacobi_kernel<false>.
struct Row4 {  // weights of four consecutive cells: w[0..5] = to x-1, x+1, y-1, y+1, z-1, z+1
    F4 d, xp, ym, yp, zm, zp;
    float xm0;  // weight of cell 0 towards x-1 (cells 1..3: xp of their left neighbour)
};
__device__ __forceinline__ Row4 load_row4(const Lv& L, int64_t c) {
    Row4 r;
    r.d = ld4(L.diag + c);
    r.xp = ld4(L.wx + c); r.xm0 = L.wx[c - 1];
    r.yp = ld4(L.wy + c); r.ym = ld4(L.wy + c - L.sy);
    r.zp = ld4(L.wz + c); r.zm = ld4(L.wz + c - L.sz);
    return r;
}
__device__ __forceinline__ bool group4_of(const Lv& L, int& x, int& y, int& z, int64_t& c) {
    x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    y = blockIdx.y * blockDim.y + threadIdx.y;
    z = blockIdx.z * blockDim.z + threadIdx.z;
    c = ((int64_t)z * L.gy + y) * L.gx + x;
    return x < L.gx && y < L.gy && z < L.gz;
}
__device__ __forceinline__ Stencil4 load_stencil4_all(const Lv& L, const float* __restrict__ x, int64_t c) {
    Stencil4 s;
    s.c = ld4(x + c);
    s.ym = ld4(x + c - L.sy); s.yp = ld4(x + c + L.sy);
    s.zm = ld4(x + c - L.sz); s.zp = ld4(x + c + L.sz);
    s.xl = x[c - 1]; s.xr = x[c + 4];
    return s;
}
// sum of w_nbr * x_nbr of cell i, in the order of row<false>
__device__ __forceinline__ float offc4(const Row4& r, const Stencil4& s, int i) {
    const float xr = i == 3 ? s.xr : s.c.v[i + 1], xl = i == 0 ? s.xl : s.c.v[i - 1];
    const float wl = i == 0 ? r.xm0 : r.xp.v[i - 1];
    return r.xp.v[i] * xr + wl * xl + r.yp.v[i] * s.yp.v[i] + r.ym.v[i] * s.ym.v[i] + r.zp.v[i] * s.zp.v[i] + r.zm.v[i] * s.zm.v[i];
}

__global__ void __launch_bounds__(256) mg_jacobic4_kernel(Lv L, const PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xin,
                                                          float* __restrict__ xout, float om) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    int x, y, z; int64_t c;
    if (!group4_of(L, x, y, z, c)) return;
    const Row4 r = load_row4(L, c);
    F4 xo = zero4();
    if (r.d.v[0] > 0.f || r.d.v[1] > 0.f || r.d.v[2] > 0.f || r.d.v[3] > 0.f) {
        const Stencil4 s = load_stencil4_all(L, xin, c);
        const F4 bb = ld4(b + c);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float d = r.d.v[i], xi = s.c.v[i];
            if (d > 0.f) xo.v[i] = xi + om * (bb.v[i] - (d * xi - offc4(r, s, i))) / d;
        }
    }
    st4(xout + c, xo);
}

__global__ void __launch_bounds__(256) mg_pre2c4_kernel(Lv L, const PcgScalars* __restrict__ sc, const float* __restrict__ b, float* __restrict__ xout,
                                                        FxWait gw) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    if (gw.has[0]) fx_wait(gw, 0);  // (see mg_pre2_kernel)
    int x, y, z; int64_t c;
    if (!group4_of(L, x, y, z, c)) return;
    const Row4 r = load_row4(L, c);
    F4 xo = zero4();
    if (r.d.v[0] > 0.f || r.d.v[1] > 0.f || r.d.v[2] > 0.f || r.d.v[3] > 0.f) {
        const Stencil4 dn = load_stencil4_all(L, L.diag, c), bn = load_stencil4_all(L, b, c);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float d = r.d.v[i], bb = bn.c.v[i];
            if (!(d > 0.f)) continue;
            // pre2_value: neighbours in the order x-1, x+1, y-1, y+1, z-1, z+1
            const float w[6] = {i == 0 ? r.xm0 : r.xp.v[i - 1], r.xp.v[i], r.ym.v[i], r.yp.v[i], r.zm.v[i], r.zp.v[i]};
            const float dk[6] = {i == 0 ? dn.xl : dn.c.v[i - 1], i == 3 ? dn.xr : dn.c.v[i + 1], dn.ym.v[i], dn.yp.v[i], dn.zm.v[i], dn.zp.v[i]};
            const float bk[6] = {i == 0 ? bn.xl : bn.c.v[i - 1], i == 3 ? bn.xr : bn.c.v[i + 1], bn.ym.v[i], bn.yp.v[i], bn.zm.v[i], bn.zp.v[i]};
            float off = 0.f;
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (w[k] > 0.f && dk[k] > 0.f) off += w[k] * (OM_A * bk[k] / dk[k]);
            const float xi = OM_A * bb / d;
            xo.v[i] = xi + OM_B * (bb - (d * xi - off)) / d;
        }
    }
    st4(xout + c, xo);
}

// restriction from a stored-operator level: one thread = two coarse cells in x = a 4x2x2 block of fine cells
__global__ void __launch_bounds__(256) mg_restrictc4_kernel(Lv L, Lv C, const PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xf,
                                                            float* __restrict__ bc) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    const int X = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y, Z = blockIdx.z * blockDim.z + threadIdx.z;
    if (X >= C.gx || Y >= C.gy || Z >= C.gz) return;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int y = 2 * Y + j, z = 2 * Z + k;
            if (y >= L.gy || z >= L.gz) continue;
            const int64_t c = ((int64_t)z * L.gy + y) * L.gx + 2 * X;
            const Row4 r = load_row4(L, c);
            if (!(r.d.v[0] > 0.f || r.d.v[1] > 0.f || r.d.v[2] > 0.f || r.d.v[3] > 0.f)) continue;
            const Stencil4 s = load_stencil4_all(L, xf, c);
            const F4 bb = ld4(b + c);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (r.d.v[i] > 0.f) {
                    const float res = bb.v[i] - (r.d.v[i] * s.c.v[i] - offc4(r, s, i));
                    if (i < 2) s0 += res; else s1 += res;
                }
        }
    *reinterpret_cast<float2*>(bc + ((int64_t)Z * C.gy + Y) * C.gx + X) = make_float2(s0, s1);
}

__global__ void __launch_bounds__(256) mg_prolong_jacobic4_kernel(Lv L, Lv C, const PcgScalars* __restrict__ sc, const float* __restrict__ b, const float* __restrict__ xin,
                                                                  const float* __restrict__ ec, float* __restrict__ xout) {
    pdl_wait();
    pdl_trigger();
    if (sc->done) return;
    int x, y, z; int64_t c;
    if (!group4_of(L, x, y, z, c)) return;
    const Row4 r = load_row4(L, c);
    F4 xo = zero4();
    if (r.d.v[0] > 0.f || r.d.v[1] > 0.f || r.d.v[2] > 0.f || r.d.v[3] > 0.f) {
        Stencil4 s = load_stencil4_all(L, xin, c);
        // coarse corrections: the group's parents are (X0, X0+1) in row (y>>1, z>>1); the neighbour rows use their own parents
        // (rows outside the grid carry zero weights; their parent row index is clamped to stay inside the padded array)
        const int X0 = x >> 1;
        auto prow = [&](int yy, int zz) -> const float* {
            const int py = min(max(yy, 0), L.gy - 1) >> 1, pz = min(max(zz, 0), L.gz - 1) >> 1;
            return ec + ((int64_t)pz * C.gy + py) * C.gx;
        };
        auto add2 = [&](F4& v, const float* row) {
            const float2 e = *reinterpret_cast<const float2*>(row + X0);
            v.v[0] += OVER * e.x; v.v[1] += OVER * e.x; v.v[2] += OVER * e.y; v.v[3] += OVER * e.y;
        };
        const float* r0 = prow(y, z);
        add2(s.c, r0);
        add2(s.ym, prow(y - 1, z)); add2(s.yp, prow(y + 1, z));
        add2(s.zm, prow(y, z - 1)); add2(s.zp, prow(y, z + 1));
        s.xl += OVER * r0[X0 - 1];   // (x = 0: the element before the row, finite, multiplied by a zero weight below)
        s.xr += OVER * r0[X0 + 2];
        const F4 bb = ld4(b + c);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float d = r.d.v[i];
            if (!(d > 0.f)) continue;
            // mg_prolong_jacobi_kernel<false>: conditional adds in the order x-1, x+1, y-1, y+1, z-1, z+1
            const float w[6] = {i == 0 ? r.xm0 : r.xp.v[i - 1], r.xp.v[i], r.ym.v[i], r.yp.v[i], r.zm.v[i], r.zp.v[i]};
            const float xn[6] = {i == 0 ? s.xl : s.c.v[i - 1], i == 3 ? s.xr : s.c.v[i + 1], s.ym.v[i], s.yp.v[i], s.zm.v[i], s.zp.v[i]};
            float off = 0.f;
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (w[k] > 0.f) off += w[k] * xn[k];
            const float xi = s.c.v[i];
            xo.v[i] = xi + OM_B * (bb.v[i] - (d * xi - off)) / d;
        }
    }
    st4(xout + c, xo);
}

// Galerkin operator of level 1 from the stencil codes: face weight = # WATER-WATER fine connections across the coarse
// face; diagonal = # connections from WATER children to non-solid cells outside the aggregate
__global__ void __launch_bounds__(256) mg_build1_kernel(Lv L, Lv C, float* __restrict__ wx, float* __restrict__ wy,
                                                         float* __restrict__ wz, float* __restrict__ diag) {
    pdl_wait();
    pdl_trigger();
    int X, Y, Z; int64_t cc;
    if (!cell_of(C, X, Y, Z, cc)) return;
    float w[3] = {0.f, 0.f, 0.f}, d = 0.f;
    for (int k = 0; k < 2; k++)
        for (int j = 0; j < 2; j++)
            for (int i = 0; i < 2; i++) {
                const int x = 2 * X + i, y = 2 * Y + j, z = 2 * Z + k;
                if (x >= L.gx || y >= L.gy || z >= L.gz) continue;
                const int64_t c = ((int64_t)z * L.gy + y) * L.gx + x;
                // from the stencil code (bits 0-5: linked WATER neighbours, bits 6-8: # non-solid neighbours); in slab mode
                // the links into ghost planes are cut in these codes, so such a neighbour counts on the diagonal only
                const unsigned cd = L.code[c];
                if (!(cd & CODE_ACTIVE)) continue;
                const int loc[3] = {i, j, k};
                d += (float)((cd >> 6) & 7u);
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const bool wm = (cd >> (2 * a)) & 1u, wp = (cd >> (2 * a + 1)) & 1u;
                    if (loc[a] == 1 && wm) d -= 1.f;  // - neighbour is a WATER sibling inside the aggregate
                    if (loc[a] == 0 && wp) d -= 1.f;  // + neighbour is a WATER sibling inside the aggregate
                    if (loc[a] == 1 && wp) w[a] += 1.f;
                }
            }
    wx[cc] = w[0]; wy[cc] = w[1]; wz[cc] = w[2]; diag[cc] = d;
}

// Galerkin operator of level l+1 from level l: outer face weights add up, inner ones cancel out of the diagonal
__global__ void __launch_bounds__(256) mg_buildn_kernel(Lv L, Lv C, float* __restrict__ wx, float* __restrict__ wy,
                                                         float* __restrict__ wz, float* __restrict__ diag) {
    pdl_wait();
    pdl_trigger();
    int X, Y, Z; int64_t cc;
    if (!cell_of(C, X, Y, Z, cc)) return;
    float w[3] = {0.f, 0.f, 0.f}, d = 0.f;
    for (int k = 0; k < 2; k++)
        for (int j = 0; j < 2; j++)
            for (int i = 0; i < 2; i++) {
                const int x = 2 * X + i, y = 2 * Y + j, z = 2 * Z + k;
                if (x >= L.gx || y >= L.gy || z >= L.gz) continue;
                const int64_t c = ((int64_t)z * L.gy + y) * L.gx + x;
                const float dj = L.diag[c];
                if (!(dj > 0.f)) continue;
                d += dj;
                const float wp[3] = {L.wx[c], L.wy[c], L.wz[c]};
                const float wm[3] = {L.wx[c - 1], L.wy[c - L.sy], L.wz[c - L.sz]};
                const int loc[3] = {i, j, k};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (loc[a] == 0) d -= wp[a];          // + neighbour is a sibling (zero weight if it does not exist)
                    else { d -= wm[a]; w[a] += wp[a]; }   // - neighbour is a sibling; + face is an outer face
                }
            }
    wx[cc] = w[0]; wy[cc] = w[1]; wz[cc] = w[2]; diag[cc] = d > 0.f ? d : 0.f;
}

// ---- the small end of the hierarchy in ONE launch -----------------------------------------------------------------
// Levels with <= TAIL_MAX_CELLS cells are latency-bound (a separate launch per sweep costs more than the sweep), so one
// thread-block cluster (16 CTAs = 16 K threads) runs their whole V(2,2) sub-cycle: all CTAs stride over the cells of a level, phases are separated
// by cluster barriers (release/acquire at cluster scope; data is exchanged through global memory, i.e. L2), and the
// coarsest level is iterated by CTA 0 in shared memory.
constexpr int TAIL_MAX_CELLS = 40000;
constexpr int TAIL_MAX_LEVELS = 8;
constexpr int TAIL_CLUSTER = 16;     // non-portable cluster size (one GPC); 8 is used when 16 cannot be scheduled
constexpr int TAIL_CLUSTER_PORTABLE = 8;
constexpr int TAIL_THREADS = 1024;
constexpr size_t TAIL_SMEM_MAX = 200 * 1024;  // dynamic shared memory the shared-memory-resident tail kernel may use per CTA

struct TailLevel { Lv L; float *xa, *xb, *b; };
struct TailArgs {
    int n;
    TailLevel lv[TAIL_MAX_LEVELS];
    const PcgScalars* sc;
    int zero_guess;
    FxWait gw;  // hybrid slab projection with the fused all-rank push, tail starting at level 1: wait for every rank's planes of b
};

__device__ __forceinline__ float t_row(const Lv& L, int c, const float* x, float* offsum) {
    *offsum = L.wx[c] * x[c + 1] + L.wx[c - 1] * x[c - 1] + L.wy[c] * x[c + L.sy] + L.wy[c - L.sy] * x[c - L.sy] +
              L.wz[c] * x[c + L.sz] + L.wz[c - L.sz] * x[c - L.sz];
    return L.diag[c];
}
__device__ void t_jacobi(const Lv& L, const float* b, const float* xin, float* xout, int t0, int nt, float om) {
    const int nc = L.gx * L.gy * L.gz;
    for (int c = t0; c < nc; c += nt) {
        float off;
        const float d = t_row(L, c, xin, &off);
        const float xi = xin[c];
        xout[c] = d > 0.f ? xi + om * (b[c] - (d * xi - off)) / d : 0.f;
    }
}
__device__ void t_pre2(const Lv& L, const float* b, float* xout, int t0, int nt) {
    const int nc = L.gx * L.gy * L.gz;
    for (int c = t0; c < nc; c += nt) xout[c] = pre2_value(L, b, c);
}
// restriction with one FINE cell per thread: lanes 8k..8k+7 hold the 2x2x2 children of one coarse cell and fold their
// residuals with three shuffles -- a thread per coarse cell walks its eight children one dependent L2 round trip after the
// other, which made this the longest stage of the tail (measured 10-11 K cycles per call, now ~3 K)
__device__ __forceinline__ float child_residual(const Lv& L, const Lv& C, const float* b, const float* xf, int cc, int sub) {
    const int X = cc % C.gx, Y = (cc / C.gx) % C.gy, Z = cc / (C.gx * C.gy);
    const int x = 2 * X + (sub & 1), y = 2 * Y + ((sub >> 1) & 1), z = 2 * Z + (sub >> 2);
    if (x >= L.gx || y >= L.gy || z >= L.gz) return 0.f;
    const int c = (z * L.gy + y) * L.gx + x;
    float off;
    const float d = t_row(L, c, xf, &off);
    return d > 0.f ? b[c] - (d * xf[c] - off) : 0.f;
}
__device__ void t_restrict(const Lv& L, const Lv& C, const float* b, const float* xf, float* bc, int t0, int nt) {
    const int ncc = C.gx * C.gy * C.gz;
    const int n8 = (ncc * 8 + 31) & ~31;  // whole warps take part in the shuffles
    for (int g = t0; g < n8; g += nt) {
        const int cc = g >> 3;
        float r = cc < ncc ? child_residual(L, C, b, xf, cc, g & 7) : 0.f;
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 4);
        if ((g & 7) == 0 && cc < ncc) bc[cc] = r;
    }
}
__device__ void t_prolong_jacobi(const Lv& L, const Lv& C, const float* b, const float* xin, const float* ec, float* xout, int t0, int nt) {
    const int nc = L.gx * L.gy * L.gz;
    for (int c = t0; c < nc; c += nt) {
        const int x = c % L.gx, y = (c / L.gx) % L.gy, z = c / (L.gx * L.gy);
        // every load up front (clamped coarse coordinates keep the addresses valid; zero weights mask what is not there)
        const int cn[7] = {c, c - 1, c + 1, c - L.sy, c + L.sy, c - L.sz, c + L.sz};
        const int xs[7] = {x, max(x - 1, 0), min(x + 1, L.gx - 1), x, x, x, x};
        const int ys[7] = {y, y, y, max(y - 1, 0), min(y + 1, L.gy - 1), y, y};
        const int zs[7] = {z, z, z, z, z, max(z - 1, 0), min(z + 1, L.gz - 1)};
        const float w[7] = {0.f, L.wx[c - 1], L.wx[c], L.wy[c - L.sy], L.wy[c], L.wz[c - L.sz], L.wz[c]};
        const float d = L.diag[c], bb = b[c];
        float xc[7];
#pragma unroll
        for (int k = 0; k < 7; k++) xc[k] = xin[cn[k]] + OVER * ec[((zs[k] >> 1) * C.gy + (ys[k] >> 1)) * C.gx + (xs[k] >> 1)];
        float off = 0.f;
#pragma unroll
        for (int k = 1; k < 7; k++)
            if (w[k] > 0.f) off += w[k] * xc[k];
        xout[c] = d > 0.f ? xc[0] + OM_B * (bb - (d * xc[0] - off)) / d : 0.f;
    }
}

// V(2,2) sub-cycle over levels lv[0..n-1] by a group of CTAs whose threads are numbered t0 (of nt) and which meet at
// bar(); the coarsest level is iterated by the group's first CTA (lead) in shared memory.  The result ends in lv[0].xa.
template <class Bar>
__device__ __forceinline__ void tail_cycle(const TailLevel* lv, int n, bool zero_guess, int t0, int nt, bool lead, float (*xs)[COARSE_MAX], Bar& bar) {
    // down
    for (int i = 0; i + 1 < n; i++) {
        const TailLevel& T = lv[i];
        if (i == 0 && !zero_guess) {
            t_jacobi(T.L, T.b, T.xa, T.xb, t0, nt, OM_A); bar();
            t_jacobi(T.L, T.b, T.xb, T.xa, t0, nt, OM_B); bar();
        } else {
            t_pre2(T.L, T.b, T.xa, t0, nt); bar();
        }
        t_restrict(T.L, lv[i + 1].L, T.b, T.xa, lv[i + 1].b, t0, nt); bar();
    }
    // coarsest level: damped Jacobi in shared memory (zero guess unless it is the only level of a revisit)
    {
        const TailLevel& T = lv[n - 1];
        const Lv& L = T.L;
        if (lead) {
            const int c = threadIdx.x;
            const int nc = L.gx * L.gy * L.gz;
            const bool in = c < nc;
            float d = 0.f, w[6] = {0, 0, 0, 0, 0, 0}, bb = 0.f;
            int nb[6] = {0, 0, 0, 0, 0, 0};
            if (in) {
                d = L.diag[c];
                bb = T.b[c];
                w[0] = L.wx[c - 1]; w[1] = L.wx[c]; w[2] = L.wy[c - L.sy]; w[3] = L.wy[c]; w[4] = L.wz[c - L.sz]; w[5] = L.wz[c];
                nb[0] = c - 1; nb[1] = c + 1; nb[2] = c - L.sy; nb[3] = c + L.sy; nb[4] = c - L.sz; nb[5] = c + L.sz;
#pragma unroll
                for (int k = 0; k < 6; k++)
                    if (!(w[k] > 0.f) || nb[k] < 0 || nb[k] >= nc) { w[k] = 0.f; nb[k] = c; }
            }
            xs[0][c] = (in && n == 1 && !zero_guess) ? T.xa[c] : 0.f;
            __syncthreads();
            int cur = 0;
            for (int s = 0; s < COARSE_SWEEPS; s++) {
                float v = 0.f;
                if (in && d > 0.f) {
                    const float* xv = xs[cur];
                    const float off = w[0] * xv[nb[0]] + w[1] * xv[nb[1]] + w[2] * xv[nb[2]] + w[3] * xv[nb[3]] + w[4] * xv[nb[4]] + w[5] * xv[nb[5]];
                    const float xi = xv[c];
                    v = xi + OMEGA * (bb - (d * xi - off)) / d;
                }
                xs[cur ^ 1][c] = v;
                __syncthreads();
                cur ^= 1;
            }
            if (in) T.xa[c] = xs[cur][c];
        }
        bar();
    }
    // up
    for (int i = n - 2; i >= 0; i--) {
        const TailLevel& T = lv[i];
        t_prolong_jacobi(T.L, lv[i + 1].L, T.b, T.xa, lv[i + 1].xa, T.xb, t0, nt); bar();
        t_jacobi(T.L, T.b, T.xb, T.xa, t0, nt, OM_A);
        if (i > 0) bar();
    }
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) mg_tail_kernel(TailArgs a) {
    pdl_wait();
    pdl_trigger();
    __shared__ float xs[2][COARSE_MAX];
    if (a.sc->done) return;  // uniform over the cluster
    if (a.gw.has[0]) fx_wait(a.gw, 0);
    cg::cluster_group cluster = cg::this_cluster();
    const int t0 = (int)cluster.block_rank() * TAIL_THREADS + threadIdx.x;
    const int nt = (int)cluster.num_blocks() * TAIL_THREADS;
    auto bar = [&]() { cluster.sync(); };
    tail_cycle(a.lv, a.n, a.zero_guess != 0, t0, nt, cluster.block_rank() == 0, xs, bar);
}

// ---- tail kernel, second form: only the FIRST tail level is shared by the cluster ------------------------------------------
// In mg_tail_kernel every phase of every level ends in a cluster barrier with release / acquire semantics: the CTA's stores
// must have reached L2 before it arrives, the loads after it miss L1 (ncu r2: 54 % of the samples in UCGABAR_WAIT, ~3 us per
// phase, 9 phases).  The levels below the first one are tiny (16^3 and 8^3 at 256^3): here EVERY CTA keeps its own copy of their
// operators and iterates in shared memory and runs their whole sub-cycle redundantly with block barriers only.  What is left
// for the cluster: pre-smoothing and restriction of the first level (its coarse right-hand side is assembled in global memory
// by all CTAs), and its prolongation + post-smoothing -- 3 cluster barriers instead of 9.  Same functions, same operand order
// as mg_tail_kernel => the same iterates.
constexpr size_t TAIL2_SMEM_MAX = 200 * 1024;
__host__ __device__ inline int tail2_pad(int sz) { return (sz + 1 + 3) & ~3; }  // halo of every local array: covers c +- sz and c +- 1

__global__ void __launch_bounds__(TAIL_THREADS, 1) mg_tail2_kernel(const __grid_constant__ TailArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float sm2[];
    __shared__ float xs[2][COARSE_MAX];
    __shared__ TailLevel loc[TAIL_MAX_LEVELS];  // levels 1 .. n-1 with every pointer redirected into shared memory
    if (a.sc->done) return;  // uniform over the cluster
    cg::cluster_group cluster = cg::this_cluster();
    const int t0 = (int)cluster.block_rank() * TAIL_THREADS + threadIdx.x;
    const int nt = (int)cluster.num_blocks() * TAIL_THREADS;
    const int n = a.n;
    // local copies of the operators of levels 1 .. n-1 (halo included: the global arrays are padded with zeros, alloc_level)
    if (threadIdx.x == 0) {
        size_t off = 0;
        for (int i = 1; i < n; i++) {
            const Lv& G = a.lv[i].L;
            const int nc = G.gx * G.gy * G.gz, pad = tail2_pad(G.sz), stride = nc + 2 * pad;
            TailLevel T = a.lv[i];
            float* base = sm2 + off + pad;
            T.L.wx = base; T.L.wy = base + stride; T.L.wz = base + 2 * stride; T.L.diag = base + 3 * stride;
            T.b = base + 4 * stride; T.xa = base + 5 * stride; T.xb = base + 6 * stride;
            loc[i - 1] = T;
            off += (size_t)7 * stride;
        }
    }
    __syncthreads();
    for (int i = 1; i < n; i++) {
        const Lv& G = a.lv[i].L;
        const TailLevel& T = loc[i - 1];
        const int nc = G.gx * G.gy * G.gz, pad = tail2_pad(G.sz);
        float* dst[4] = {const_cast<float*>(T.L.wx), const_cast<float*>(T.L.wy), const_cast<float*>(T.L.wz), const_cast<float*>(T.L.diag)};
        const float* src[4] = {G.wx, G.wy, G.wz, G.diag};
        for (int c = (int)threadIdx.x - pad; c < nc + pad; c += TAIL_THREADS) {
#pragma unroll
            for (int k = 0; k < 4; k++) dst[k][c] = src[k][c];
            T.b[c] = 0.f; T.xa[c] = 0.f; T.xb[c] = 0.f;
        }
    }
    // first level: pre-smoothing and restriction by the whole cluster, through global memory
    const TailLevel& F = a.lv[0];
    if (!a.zero_guess) {
        t_jacobi(F.L, F.b, F.xa, F.xb, t0, nt, OM_A); cluster.sync();
        t_jacobi(F.L, F.b, F.xb, F.xa, t0, nt, OM_B); cluster.sync();
    } else {
        t_pre2(F.L, F.b, F.xa, t0, nt); cluster.sync();
    }
    t_restrict(F.L, a.lv[1].L, F.b, F.xa, a.lv[1].b, t0, nt); cluster.sync();
    // every CTA: the coarse right-hand side, then the whole sub-cycle of levels 1 .. n-1 in its own shared memory
    {
        const Lv& G = a.lv[1].L;
        const int nc = G.gx * G.gy * G.gz;
        for (int c = threadIdx.x; c < nc; c += TAIL_THREADS) loc[0].b[c] = a.lv[1].b[c];
    }
    __syncthreads();
    auto bbar = [&]() { __syncthreads(); };
    tail_cycle(loc, n - 1, true, (int)threadIdx.x, TAIL_THREADS, true, xs, bbar);
    __syncthreads();
    // first level: prolongation + post-smoothing (the coarse correction is this CTA's own copy)
    t_prolong_jacobi(F.L, loc[0].L, F.b, F.xa, loc[0].xa, F.xb, t0, nt); cluster.sync();
    t_jacobi(F.L, F.b, F.xb, F.xa, t0, nt, OM_A);
}

// ---- the same sub-cycle with every tail level RESIDENT IN (DISTRIBUTED) SHARED MEMORY -------------------------------------
// mg_tail_kernel exchanges its iterates through global memory: every phase is at least one dependent L2 round trip plus the
// store drain in front of the cluster barrier (measured 42 us per call at 256^3, two calls per cycle = the largest single
// item of the coarse levels).  Here CTA r of the cluster owns a block of z-planes of every level and keeps their operator
// rows (six face weights + diagonal), right-hand side and both iterates in its shared memory for the whole call; in-plane
// neighbours are local shared-memory reads, z-neighbours and parent / child cells of other blocks are read (or written)
// through distributed shared memory (mapa + ld/st.shared::cluster).  A phase then costs a shared-memory round trip instead
// of an L2 one.  Plane blocks start at even planes, so the eight children of a coarse cell always share an owner; the
// coarsest level lives in CTA 0.  Same arithmetic, same order of operations as tail_cycle => the same iterates.
enum { SA_WXM = 0, SA_WXP, SA_WYM, SA_WYP, SA_WZM, SA_WZP, SA_D, SA_B, SA_XA, SA_XB, SA_COUNT };

struct SLevelArg {
    int gx, gy, gz;
    int cap;   // cells of the largest block of this level (uniform shared-memory layout across the cluster)
    int off;   // float offset of the level's arrays in dynamic shared memory
    const float *wx, *wy, *wz, *diag;
};
struct STailArgs {
    int n;  // levels; the last one is the coarsest (<= COARSE_MAX cells, owned by CTA 0)
    SLevelArg lv[TAIL_MAX_LEVELS];
    const float* b_in;   // right-hand side of the first level (global)
    const float* xa_in;  // its iterate when the visit does not start from zero
    float* xa_out;       // result
    const PcgScalars* sc;
    int zero_guess;
};

constexpr int STAIL_MAX_PLANES = 64;  // z-planes of a level the owner tables hold (larger levels: global-memory kernel)
struct SLv {
    int gx, gy, gz, plane, cap, hz, coarsest, zb, ze;
    float* base;
    const uint8_t* own;   // [gz] cluster rank that owns plane z          (tables in shared memory: the hot loops must not
    const int16_t* z0;    // [gz] first plane of that rank's block          spend their time in integer divisions)
};
// blocks of plane PAIRS: rank r owns pairs [r*hz/nb, (r+1)*hz/nb), hz = ceil(gz/2)
__device__ __forceinline__ int s_z0(const SLv& L, int r, int nb) {
    if (L.coarsest) return r == 0 ? 0 : L.gz;
    return min(2 * ((r * L.hz) / nb), L.gz);
}
__device__ __forceinline__ int s_owner(const SLv& L, int z, int nb) {
    if (L.coarsest) return 0;
    return (((z >> 1) + 1) * nb + L.hz - 1) / L.hz - 1;
}
__device__ __forceinline__ float* s_arr(const SLv& L, int arr) { return L.base + arr * L.cap; }
// address of element (plane z, in-plane index j) of array arr in the shared memory of the CTA that owns plane z
__device__ __forceinline__ float* s_remote(cg::cluster_group& cl, const SLv& L, int arr, int z, int j) {
    float* p = L.base + arr * L.cap + (z - (int)L.z0[z]) * L.plane + j;
    return cl.map_shared_rank(p, (int)L.own[z]);
}

// x-, x+, y-, y+, z-, z+ neighbour values of array arr around local cell li = (z - zb) * plane + j (0 where the weight is 0)
struct Nb6 { float v[6]; };
__device__ __forceinline__ Nb6 s_nb6(cg::cluster_group& cl, const SLv& L, int arr, int li, int z, int j, const float w[6]) {
    Nb6 r;
    const float* a = s_arr(L, arr);
    r.v[0] = w[0] > 0.f ? a[li - 1] : 0.f;
    r.v[1] = w[1] > 0.f ? a[li + 1] : 0.f;
    r.v[2] = w[2] > 0.f ? a[li - L.gx] : 0.f;
    r.v[3] = w[3] > 0.f ? a[li + L.gx] : 0.f;
    r.v[4] = w[4] > 0.f ? (z - 1 >= L.zb ? a[li - L.plane] : *s_remote(cl, L, arr, z - 1, j)) : 0.f;
    r.v[5] = w[5] > 0.f ? (z + 1 < L.ze ? a[li + L.plane] : *s_remote(cl, L, arr, z + 1, j)) : 0.f;
    return r;
}
__device__ __forceinline__ void s_w6(const SLv& L, int li, float w[6]) {
#pragma unroll
    for (int k = 0; k < 6; k++) w[k] = s_arr(L, SA_WXM + k)[li];
}
// t_row: (diag, sum w x) with the operand order of the global-memory kernels: x+, x-, y+, y-, z+, z-
__device__ __forceinline__ float s_off(const float w[6], const Nb6& x) {
    return w[1] * x.v[1] + w[0] * x.v[0] + w[3] * x.v[3] + w[2] * x.v[2] + w[5] * x.v[5] + w[4] * x.v[4];
}
// the cells of this CTA's block: planes outside, in-plane index inside (no division per cell)
#define S_FOR_CELLS(L, z, j, li)                      \
    for (int z = (L).zb; z < (L).ze; z++)             \
        for (int j = threadIdx.x, li = (z - (L).zb) * (L).plane + threadIdx.x; j < (L).plane; j += TAIL_THREADS, li += TAIL_THREADS)

__device__ void s_jacobi(cg::cluster_group& cl, const SLv& L, int in, int out, float om) {
    S_FOR_CELLS(L, z, j, li) {
        float w[6];
        s_w6(L, li, w);
        const Nb6 x = s_nb6(cl, L, in, li, z, j, w);
        const float d = s_arr(L, SA_D)[li], xi = s_arr(L, in)[li];
        s_arr(L, out)[li] = d > 0.f ? xi + om * (s_arr(L, SA_B)[li] - (d * xi - s_off(w, x))) / d : 0.f;
    }
}
// two pre-smoothing sweeps from a zero guess in one pass (pre2_value)
__device__ void s_pre2(cg::cluster_group& cl, const SLv& L) {
    S_FOR_CELLS(L, z, j, li) {
        float w[6];
        s_w6(L, li, w);
        const Nb6 dn = s_nb6(cl, L, SA_D, li, z, j, w), bn = s_nb6(cl, L, SA_B, li, z, j, w);
        const float d = s_arr(L, SA_D)[li], bb = s_arr(L, SA_B)[li];
        float off = 0.f;
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (w[k] > 0.f && dn.v[k] > 0.f) off += w[k] * (OM_A * bn.v[k] / dn.v[k]);
        float v = 0.f;
        if (d > 0.f) {
            const float xi = OM_A * bb / d;
            v = xi + OM_B * (bb - (d * xi - off)) / d;
        }
        s_arr(L, SA_XA)[li] = v;
    }
}
// restriction: the eight children of a coarse cell sit in eight neighbouring lanes of the CTA that owns them; the sum goes
// to the coarse cell's owner
__device__ void s_restrict(cg::cluster_group& cl, const SLv& L, const SLv& C) {
    const int Z0 = L.zb >> 1, Z1 = (L.ze + 1) >> 1;  // coarse planes whose children this CTA owns (blocks start at even planes)
    const int n8 = (C.plane * 8 + 31) & ~31;         // whole warps take part in the shuffles
    for (int Z = Z0; Z < Z1; Z++)
        for (int g = threadIdx.x; g < n8; g += TAIL_THREADS) {
            const int J = g >> 3, sub = g & 7;
            float r = 0.f;
            if (J < C.plane) {
                const int X = J % C.gx, Y = J / C.gx;
                const int x = 2 * X + (sub & 1), y = 2 * Y + ((sub >> 1) & 1), z = 2 * Z + (sub >> 2);
                if (x < L.gx && y < L.gy && z < L.ze) {
                    const int j = y * L.gx + x, li = (z - L.zb) * L.plane + j;
                    const float d = s_arr(L, SA_D)[li];
                    float w[6];
                    s_w6(L, li, w);
                    const Nb6 xn = s_nb6(cl, L, SA_XA, li, z, j, w);
                    if (d > 0.f) r = s_arr(L, SA_B)[li] - (d * s_arr(L, SA_XA)[li] - s_off(w, xn));
                }
            }
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            r += __shfl_xor_sync(0xffffffffu, r, 2);
            r += __shfl_xor_sync(0xffffffffu, r, 4);
            if (sub == 0 && J < C.plane) *s_remote(cl, C, SA_B, Z, J) = r;
        }
}
// prolongation + over-corrected update fused with the first post-smoothing sweep (t_prolong_jacobi): XA -> XB
__device__ void s_prolong_jacobi(cg::cluster_group& cl, const SLv& L, const SLv& C) {
    S_FOR_CELLS(L, z, j, li) {
        const int y = j / L.gx, x = j - y * L.gx;
        float w[6];
        s_w6(L, li, w);
        const Nb6 xn = s_nb6(cl, L, SA_XA, li, z, j, w);
        // parents: own row (y>>1, z>>1) and the rows of the y / z neighbours
        const int jp = (y >> 1) * C.gx + (x >> 1);
        const float* e_own = s_remote(cl, C, SA_XA, z >> 1, 0);
        const float d = s_arr(L, SA_D)[li], bb = s_arr(L, SA_B)[li];
        const float x0 = s_arr(L, SA_XA)[li] + OVER * e_own[jp];
        float off = 0.f;
        if (w[0] > 0.f) off += w[0] * (xn.v[0] + OVER * e_own[(y >> 1) * C.gx + ((x - 1) >> 1)]);
        if (w[1] > 0.f) off += w[1] * (xn.v[1] + OVER * e_own[(y >> 1) * C.gx + ((x + 1) >> 1)]);
        if (w[2] > 0.f) off += w[2] * (xn.v[2] + OVER * e_own[((y - 1) >> 1) * C.gx + (x >> 1)]);
        if (w[3] > 0.f) off += w[3] * (xn.v[3] + OVER * e_own[((y + 1) >> 1) * C.gx + (x >> 1)]);
        if (w[4] > 0.f) off += w[4] * (xn.v[4] + OVER * *s_remote(cl, C, SA_XA, (z - 1) >> 1, jp));
        if (w[5] > 0.f) off += w[5] * (xn.v[5] + OVER * *s_remote(cl, C, SA_XA, (z + 1) >> 1, jp));
        s_arr(L, SA_XB)[li] = d > 0.f ? x0 + OM_B * (bb - (d * x0 - off)) / d : 0.f;
    }
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) mg_tail_smem_kernel(const __grid_constant__ STailArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float s_dyn[];
    if (a.sc->done) return;  // uniform over the cluster
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank(), nb = (int)cl.num_blocks();
    __shared__ SLv lv[TAIL_MAX_LEVELS];
    __shared__ uint8_t t_own[TAIL_MAX_LEVELS][STAIL_MAX_PLANES];
    __shared__ int16_t t_z0[TAIL_MAX_LEVELS][STAIL_MAX_PLANES];
    if (threadIdx.x < a.n) {
        const int i = threadIdx.x;
        SLv L;
        const SLevelArg& A = a.lv[i];
        L.gx = A.gx; L.gy = A.gy; L.gz = A.gz; L.plane = A.gx * A.gy; L.cap = A.cap; L.hz = (A.gz + 1) >> 1;
        L.coarsest = i == a.n - 1;
        L.base = s_dyn + A.off;
        L.zb = s_z0(L, rank, nb);
        L.ze = (L.coarsest || rank == nb - 1) ? L.gz : s_z0(L, rank + 1, nb);
        if (L.coarsest && rank != 0) L.zb = L.ze = L.gz;
        L.own = t_own[i]; L.z0 = t_z0[i];
        lv[i] = L;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < a.n * STAIL_MAX_PLANES; t += TAIL_THREADS) {  // owner tables: one entry per (level, plane)
        const int i = t / STAIL_MAX_PLANES, z = t % STAIL_MAX_PLANES;
        if (z < lv[i].gz) {
            const int o = s_owner(lv[i], z, nb);
            t_own[i][z] = (uint8_t)o;
            t_z0[i][z] = (int16_t)s_z0(lv[i], o, nb);
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < a.n; i++) {
        const SLv& L = lv[i];
        const SLevelArg& A = a.lv[i];
        // operator rows of the planes this CTA owns (one pass over global memory, all loads independent)
        S_FOR_CELLS(L, z, j, li) {
            const int y = j / L.gx, x = j - y * L.gx;
            const int c = z * L.plane + j;
            s_arr(L, SA_WXM)[li] = x > 0 ? A.wx[c - 1] : 0.f;
            s_arr(L, SA_WXP)[li] = x + 1 < L.gx ? A.wx[c] : 0.f;
            s_arr(L, SA_WYM)[li] = y > 0 ? A.wy[c - L.gx] : 0.f;
            s_arr(L, SA_WYP)[li] = y + 1 < L.gy ? A.wy[c] : 0.f;
            s_arr(L, SA_WZM)[li] = z > 0 ? A.wz[c - L.plane] : 0.f;
            s_arr(L, SA_WZP)[li] = z + 1 < L.gz ? A.wz[c] : 0.f;
            s_arr(L, SA_D)[li] = A.diag[c];
            if (i == 0) {
                s_arr(L, SA_B)[li] = a.b_in[c];
                if (!a.zero_guess) s_arr(L, SA_XA)[li] = a.xa_in[c];
            }
        }
    }
    cl.sync();
    const int n = a.n;
    // down
    for (int i = 0; i + 1 < n; i++) {
        if (i == 0 && !a.zero_guess) {
            s_jacobi(cl, lv[0], SA_XA, SA_XB, OM_A); cl.sync();
            s_jacobi(cl, lv[0], SA_XB, SA_XA, OM_B); cl.sync();
        } else {
            s_pre2(cl, lv[i]); cl.sync();
        }
        s_restrict(cl, lv[i], lv[i + 1]); cl.sync();
    }
    // coarsest level: damped Jacobi by CTA 0, one cell per thread, iterates ping-pong between XA and XB
    if (rank == 0) {
        const SLv& L = lv[n - 1];
        const int c = threadIdx.x, nc = L.gz * L.plane;
        const bool in = c < nc;
        float d = 0.f, w[6] = {0, 0, 0, 0, 0, 0}, bb = 0.f;
        int nbi[6] = {0, 0, 0, 0, 0, 0};
        if (in) {
            d = s_arr(L, SA_D)[c]; bb = s_arr(L, SA_B)[c];
            s_w6(L, c, w);
            nbi[0] = c - 1; nbi[1] = c + 1; nbi[2] = c - L.gx; nbi[3] = c + L.gx; nbi[4] = c - L.plane; nbi[5] = c + L.plane;
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (!(w[k] > 0.f)) { w[k] = 0.f; nbi[k] = c; }
            if (n != 1 || a.zero_guess) s_arr(L, SA_XA)[c] = 0.f;
        }
        __syncthreads();
        int cur = SA_XA;
        for (int sw = 0; sw < COARSE_SWEEPS; sw++) {
            float v = 0.f;
            if (in && d > 0.f) {
                const float* xv = s_arr(L, cur);
                const float off = w[0] * xv[nbi[0]] + w[1] * xv[nbi[1]] + w[2] * xv[nbi[2]] + w[3] * xv[nbi[3]] + w[4] * xv[nbi[4]] + w[5] * xv[nbi[5]];
                const float xi = xv[c];
                v = xi + OMEGA * (bb - (d * xi - off)) / d;
            }
            const int nxt = cur == SA_XA ? SA_XB : SA_XA;
            if (in) s_arr(L, nxt)[c] = v;
            __syncthreads();
            cur = nxt;
        }
        if (in && cur != SA_XA) s_arr(L, SA_XA)[c] = s_arr(L, SA_XB)[c];  // (COARSE_SWEEPS is even: not taken)
    }
    cl.sync();
    // up
    for (int i = n - 2; i >= 0; i--) {
        s_prolong_jacobi(cl, lv[i], lv[i + 1]); cl.sync();
        s_jacobi(cl, lv[i], SA_XB, SA_XA, OM_A);
        if (i > 0) cl.sync();
    }
    // result of the first level -> global (only this CTA's threads wrote these entries: a block barrier orders them)
    __syncthreads();
    {
        const SLv& L = lv[0];
        const int nn = (L.ze - L.zb) * L.plane;
        for (int li = threadIdx.x; li < nn; li += TAIL_THREADS) a.xa_out[L.zb * L.plane + li] = s_arr(L, SA_XA)[li];
    }
    cl.sync();  // no CTA may exit while a peer can still read its shared memory
}

Lv view(const fsim* h, const MgLevel* m, int level) {
    Lv v;
    v.gx = m->gx; v.gy = m->gy; v.gz = m->gz; v.sy = m->sy; v.sz = m->sz;
    v.wx = m->wx; v.wy = m->wy; v.wz = m->wz; v.diag = m->diag;
    v.flags = level == 0 ? h->flags : nullptr;
    v.code = level == 0 ? h->code_mg : nullptr;
    v.z0 = 0; v.z1 = m->gz;
    if (level == 0 && h->hybrid) { v.z0 = h->g.zown0; v.z1 = h->g.zown1; }
    return v;
}

dim3 grid_of(const MgLevel* m, dim3 blk) { return dim3(div_up(m->gx, blk.x), div_up(m->gy, blk.y), div_up(m->gz, blk.z)); }

int alloc_level(fsim* h, MgLevel* m, bool fine) {
    m->sy = m->gx; m->sz = m->gx * m->gy; m->nc = (int64_t)m->gx * m->gy * m->gz;
    m->pad = fine ? 0 : (size_t)m->sz + 32;
    const size_t n = (size_t)m->nc + 2 * m->pad;
    float** dst[7] = {&m->wx, &m->wy, &m->wz, &m->diag, &m->xa, &m->xb, &m->b};
    for (int k = 0; k < 7; k++) {
        m->base[k] = nullptr;
        *dst[k] = nullptr;
        if (fine && k < 4) continue;  // level 0 is matrix-free
        if (fine && k == 6) { *dst[k] = h->mg_b0; continue; }  // lives next to the codes (fsim_create)
        FSIM_CUDA(h, cudaMalloc((void**)&m->base[k], n * sizeof(float)));
        FSIM_CUDA(h, cudaMemsetAsync(m->base[k], 0, n * sizeof(float), h->stream));
        *dst[k] = m->base[k] + m->pad;
    }
    return FSIM_OK;
}

// one multigrid cycle on level l for the right-hand side m->b; returns the array holding the result
int cycle(fsim* h, int l, bool zero_guess, float** result, bool first_done = false, bool with_dot = false) {
    MgLevel* m = h->mg[l];
    const dim3 blk(32, 4, 2);
    const Lv L = view(h, m, l);
    const PcgScalars* sc = h->scal;
    const bool fine = l == 0;
    const int kid = l == 0 ? K_MG : (l == 1 ? K_MG1 : K_MG2);
    const bool v4 = fine && (m->gx % 4 == 0);  // float4 path; then the coarse gx is even (float2 stores / loads)
    const dim3 blk4(32, 4, 2);
    const dim3 grd4(div_up(m->gx, 4 * 32), div_up(m->gy, 4), div_up(L.z1 - L.z0, 2));
    const dim3 grdL(div_up(m->gx, blk.x), div_up(m->gy, blk.y), div_up(L.z1 - L.z0, blk.z));  // per-cell kernels of this level
    if (l == h->mg_tail_first) {  // this level and everything below it: one cluster launch
        TailArgs ta;
        ta.n = (int)h->mg.size() - l;
        for (int i = 0; i < ta.n; i++) {
            MgLevel* t = h->mg[l + i];
            ta.lv[i].L = view(h, t, l + i);
            ta.lv[i].xa = t->xa; ta.lv[i].xb = t->xb; ta.lv[i].b = t->b;
        }
        ta.sc = sc;
        ta.zero_guess = zero_guess ? 1 : 0;
        memset(&ta.gw, 0, sizeof(ta.gw));
        if (l == 1 && zero_guess && h->hybrid && h->gx_on) {  // its right-hand side arrives through the fused all-rank push
            FxAllPush gp;
            if (!dist_gx(h, m->gx, m->gy, m->gz, &gp, &ta.gw)) return fsim_fail(h, FSIM_ERR_COMM, "fused all-rank push: peer table unavailable");
        }
        cudaLaunchConfig_t cfg = {};
        if (h->mg_tail_cluster == 0) {  // first use: can a 16-CTA cluster of 1024-thread CTAs be co-scheduled on this device?
            h->mg_tail_cluster = TAIL_CLUSTER_PORTABLE;
            if (cudaFuncSetAttribute(mg_tail_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t q = {};
                q.gridDim = dim3(TAIL_CLUSTER); q.blockDim = dim3(TAIL_THREADS);
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = TAIL_CLUSTER; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                q.attrs = qa; q.numAttrs = 1;
                int nclusters = 0;
                if (cudaOccupancyMaxActiveClusters(&nclusters, mg_tail_kernel, &q) == cudaSuccess && nclusters >= 1) h->mg_tail_cluster = TAIL_CLUSTER;
            }
            cudaGetLastError();
            // FSIM_MG_TAIL_CLUSTER=<1..16>: several slab handles sharing ONE device (tests) must not need a whole GPC each
            if (const char* e = getenv("FSIM_MG_TAIL_CLUSTER")) { const int v = atoi(e); if (v >= 1 && v <= h->mg_tail_cluster) h->mg_tail_cluster = v; }
        }
        const int tail_cluster = h->mg_tail_cluster;
        // shared-memory-resident variant (mg_tail_smem_kernel) when every level's block fits the CTA's shared memory
        if (h->mg_tail_smem != 0) {
            STailArgs sa;
            sa.n = ta.n; sa.sc = sc; sa.zero_guess = ta.zero_guess;
            sa.b_in = m->b; sa.xa_in = m->xa; sa.xa_out = m->xa;
            size_t floats = 0;
            for (int i = 0; i < sa.n; i++) {
                MgLevel* t = h->mg[l + i];
                SLevelArg& A = sa.lv[i];
                A.gx = t->gx; A.gy = t->gy; A.gz = t->gz; A.wx = t->wx; A.wy = t->wy; A.wz = t->wz; A.diag = t->diag;
                const int plane = t->gx * t->gy, hz = (t->gz + 1) / 2;
                int planes = 0;
                if (i == sa.n - 1) planes = t->gz;  // coarsest: CTA 0 holds all of it
                else
                    for (int r = 0; r < tail_cluster; r++) {
                        const int z0 = std::min(2 * ((r * hz) / tail_cluster), t->gz);
                        const int z1 = r == tail_cluster - 1 ? t->gz : std::min(2 * (((r + 1) * hz) / tail_cluster), t->gz);
                        planes = std::max(planes, z1 - z0);
                    }
                A.cap = std::max(planes * plane, 1);
                A.off = (int)floats;
                floats += (size_t)SA_COUNT * A.cap;
            }
            const size_t bytes = floats * sizeof(float);
            bool coarse_ok = h->mg[l + sa.n - 1]->nc <= TAIL_THREADS;
            for (int i = 0; i < sa.n; i++) coarse_ok = coarse_ok && h->mg[l + i]->gz <= STAIL_MAX_PLANES;
            if (h->mg_tail_smem < 0) {  // first use: opt in to the large dynamic shared memory, make sure the cluster still fits
                h->mg_tail_smem = 0;
                // opt-in (FSIM_MG_TAIL_SMEM=1): measured SLOWER than the global-memory kernel on B200 (60 us vs 33 us per call at
                // 256^3 under ncu, +0.4 ms per step; profiles/r2_mg_tail_ab.md) -- both spend their time in cluster barriers,
                // and this one executes 1.8x the instructions (operator rows re-staged on every call, generic loads)
                if ((getenv("FSIM_MG_TAIL_SMEM") && getenv("FSIM_MG_TAIL_SMEM")[0] == '1') &&
                    cudaFuncSetAttribute(mg_tail_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL_SMEM_MAX) == cudaSuccess &&
                    cudaFuncSetAttribute(mg_tail_smem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                    cudaLaunchConfig_t q = {};
                    q.gridDim = dim3(tail_cluster); q.blockDim = dim3(TAIL_THREADS); q.dynamicSmemBytes = std::min(bytes, TAIL_SMEM_MAX);
                    cudaLaunchAttribute qa[1];
                    qa[0].id = cudaLaunchAttributeClusterDimension;
                    qa[0].val.clusterDim.x = tail_cluster; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                    q.attrs = qa; q.numAttrs = 1;
                    int nclusters = 0;
                    if (cudaOccupancyMaxActiveClusters(&nclusters, mg_tail_smem_kernel, &q) == cudaSuccess && nclusters >= 1) h->mg_tail_smem = 1;
                }
                cudaGetLastError();
            }
            if (h->mg_tail_smem == 1 && bytes <= TAIL_SMEM_MAX && coarse_ok) {
                cfg.gridDim = dim3(tail_cluster);
                cfg.blockDim = dim3(TAIL_THREADS);
                cfg.dynamicSmemBytes = bytes;
                cfg.stream = h->stream;
                cudaLaunchAttribute at[2];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = tail_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[1].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = h->pdl ? 2 : 1;
                KScope ks(h, K_MG2);
                FSIM_CUDA(h, cudaLaunchKernelEx(&cfg, mg_tail_smem_kernel, sa));
                *result = m->xa;
                return FSIM_OK;
            }
        }
        cfg.gridDim = dim3(tail_cluster);
        cfg.blockDim = dim3(TAIL_THREADS);
        cfg.stream = h->stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = tail_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = h->pdl ? 2 : 1;
        // second form (mg_tail2_kernel): the levels below the first one run redundantly in every CTA's shared memory
        size_t bytes2 = 0;
        for (int i = 1; i < ta.n; i++) {
            const MgLevel* t = h->mg[l + i];
            bytes2 += (size_t)7 * (t->nc + 2 * tail2_pad(t->sz)) * sizeof(float);
        }
        if (h->mg_tail2 < 0) {  // first use: opt in to the dynamic shared memory, make sure the cluster can still be scheduled
            h->mg_tail2 = 0;
            // opt-in (FSIM_MG_TAIL2=1): measured SLOWER on B200 (+15 us per call at 256^3: profiles/r2_mg_tail_ab.md) -- re-staging the
            // operators of the local levels on every call and running them on one CTA's 1024 threads costs more than the six
            // cluster barriers it removes
            const char* e = getenv("FSIM_MG_TAIL2");
            if ((e && e[0] == '1') && cudaFuncSetAttribute(mg_tail2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL2_SMEM_MAX) == cudaSuccess &&
                cudaFuncSetAttribute(mg_tail2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t q = cfg;
                q.numAttrs = 1; q.dynamicSmemBytes = TAIL2_SMEM_MAX;
                int nclusters = 0;
                if (cudaOccupancyMaxActiveClusters(&nclusters, mg_tail2_kernel, &q) == cudaSuccess && nclusters >= 1) h->mg_tail2 = 1;
            }
            cudaGetLastError();
        }
        KScope ks(h, K_MG2);
        if (h->mg_tail2 == 1 && ta.n >= 2 && bytes2 <= TAIL2_SMEM_MAX) {
            cfg.dynamicSmemBytes = bytes2;
            FSIM_CUDA(h, cudaLaunchKernelEx(&cfg, mg_tail2_kernel, ta));
        } else
            FSIM_CUDA(h, cudaLaunchKernelEx(&cfg, mg_tail_kernel, ta));
        *result = m->xa;
        return FSIM_OK;
    }
    MgLevel* mc = h->mg[l + 1];
    const Lv C = view(h, mc, l + 1);
    float *cur = m->xa, *oth = m->xb;
    // hybrid slab projection with fused exchanges (fexch.cuh): the kernel that writes an iterate stores its boundary planes into
    // the neighbours' ghost planes, the kernel that reads it waits for theirs -- no exchange kernels between the sweeps
    // stored-operator levels with >= 1 M cells (level 1 of a 256^3 grid): four cells per thread (mg_*c4_kernel).  Measured at 256^3: level 1
    // 0.65 -> 0.53 ms/step, whole step -0.09 ms; also on level 2 (262 k cells) +0.06 ms -- there the launch is the cost, not the waves
    const int64_t c4_min = getenv("FSIM_MG_C4_MIN") ? atoll(getenv("FSIM_MG_C4_MIN")) : 1000000;
    const bool c4 = !fine && m->gx % 4 == 0 && m->pad % 4 == 0 && m->nc >= c4_min;
    const dim3 grdC4(div_up(m->gx, 4 * 32), div_up(m->gy, 4), div_up(m->gz, 2));
    const bool fx = fine && h->hybrid && h->fx_on && v4;
    const int zrot_first = fx ? (int)grd4.z - 1 : 0;  // producers: the CTAs of the top boundary plane first, then the bottom one, ...
    FxPush no_push; FxWait no_wait;
    memset(&no_push, 0, sizeof(no_push));
    memset(&no_wait, 0, sizeof(no_wait));
    int fx_rc = FSIM_OK;
    auto push_of = [&](const float* arr) -> FxPush {
        FxPush p = no_push; FxWait w;
        if (fx) { if (!dist_fx(h, SYM_X, arr, &p, &w)) fx_rc = FSIM_ERR_COMM; p.nblk[0] = p.nblk[1] = grd4.x * grd4.y; }
        return p;
    };
    auto wait_of = [&](const float* arr) -> FxWait {
        FxPush p; FxWait w = no_wait;
        if (fx && !dist_fx(h, SYM_X, arr, &p, &w)) fx_rc = FSIM_ERR_COMM;
        return w;
    };
    if (!fine && zero_guess && PRE == 2) {
        KScope ks(h, kid);
        FxWait gw;
        memset(&gw, 0, sizeof(gw));
        if (l == 1 && h->hybrid && h->gx_on) {  // the right-hand side arrives through the fused all-rank push
            FxAllPush gp;
            if (!dist_gx(h, m->gx, m->gy, m->gz, &gp, &gw)) return fsim_fail(h, FSIM_ERR_COMM, "fused all-rank push: peer table unavailable");
        }
        if (c4) launch_k(h, mg_pre2c4_kernel, grdC4, blk4, 0, L, sc, m->b, cur, gw);
        else launch_k(h, mg_pre2_kernel, grdL, blk, 0, L, sc, m->b, cur, gw);
    } else {
        std::unique_ptr<KScope> ks(new KScope(h, kid, PRE - ((zero_guess && first_done) ? 1 : 0)));
        for (int s = 0; s < PRE; s++) {
            if (s == 0 && zero_guess && first_done) continue;  // x1 and b were written by the fused CG update
            if (s == 0 && zero_guess) {
                if (v4) launch_k(h, mg_first4_kernel, grd4, blk4, 0, L, h->r, sc, m->b, cur, push_of(cur), zrot_first);
                else if (fine) launch_k(h, mg_first_kernel<true>, grdL, blk, 0, L, h->r, sc, m->b, cur);
                else launch_k(h, mg_first_kernel<false>, grdL, blk, 0, L, nullptr, sc, m->b, cur);
            } else {
                const float om = s == 0 ? OM_A : OM_B;
                if (fine && h->hybrid && !fx) {  // the neighbours' planes of the iterate (timed as an exchange, not as a sweep)
                    ks.reset();
                    int rc = dist_halo_sym(h, SYM_X, cur, true);
                    if (rc) return rc;
                    ks.reset(new KScope(h, kid, 0));
                }
                if (v4) launch_k(h, mg_jacobi4_kernel, grd4, blk4, 0, L, h->scal, m->b, cur, oth, om, wait_of(cur), push_of(oth), zrot_first);
                else if (fine) launch_k(h, mg_jacobi_kernel<true>, grdL, blk, 0, L, sc, m->b, cur, oth, om);
                else if (c4) launch_k(h, mg_jacobic4_kernel, grdC4, blk4, 0, L, sc, m->b, cur, oth, om);
                else launch_k(h, mg_jacobi_kernel<false>, grdL, blk, 0, L, sc, m->b, cur, oth, om);
                float* t = cur; cur = oth; oth = t;
            }
        }
    }
    if (fine && h->hybrid && !fx) { int rc = dist_halo_sym(h, SYM_X, cur, true); if (rc) return rc; }
    // hybrid: the coarse planes whose children this rank owns (slab boundaries are even); otherwise all of them
    const int cz0 = (fine && h->hybrid) ? h->g.zown0 / 2 : 0;
    const int cz1 = (fine && h->hybrid && h->g.zown1 < h->g.gz) ? h->g.zown1 / 2 : (1 << 30);
    {
        KScope ks(h, kid);
        if (v4) {
            const dim3 grdR(div_up(mc->gx, 2 * 32), div_up(mc->gy, 4), div_up(std::min(cz1, mc->gz) - cz0, 2));
            // (consumer of the iterate's ghost planes: the layers next to them come last)
            // fused all-rank push of the coarse right-hand side (fexch.cuh): needs the default tail kernel when the tail starts at level 1
            FxAllPush gp;
            FxWait gw_unused;
            memset(&gp, 0, sizeof(gp));
            // (ranks sharing one process and device -- the test harness -- keep the gather kernel when the consumer is a full-grid
            // kernel: all of its CTAs spin on the counter and could starve the other ranks' producers of CTA slots)
            h->gx_on = fx && (h->mg_tail_first == 1 || !dist_peer_in_process(h)) && (h->mg_tail_first > 1 || (h->mg_tail_smem != 1 && h->mg_tail2 != 1 && !(getenv("FSIM_MG_TAIL_SMEM") || getenv("FSIM_MG_TAIL2")))) &&
                       (PRE == 2) && dist_gx(h, mc->gx, mc->gy, mc->gz, &gp, &gw_unused);
            launch_k(h, mg_restrict4_kernel, grdR, blk4, 0, L, C, sc, m->b, cur, mc->b, cz0, cz1, wait_of(cur), (fx && grdR.z > 1) ? 1 : 0, gp);
        }
        else if (fine) launch_k(h, mg_restrict_kernel<true>, grid_of(mc, blk), blk, 0, L, C, sc, m->b, cur, mc->b);
        else if (c4 && mc->gx % 2 == 0 && mc->pad % 2 == 0)
            launch_k(h, mg_restrictc4_kernel, dim3(div_up(mc->gx, 2 * 32), div_up(mc->gy, 4), div_up(mc->gz, 2)), blk4, 0, L, C, sc, m->b, cur, mc->b);
        else if (mc->nc > 100000) launch_k(h, mg_restrict_kernel<false>, grid_of(mc, blk), blk, 0, L, C, sc, m->b, cur, mc->b);  // enough threads as it is
        else launch_k(h, mg_restrict8_kernel, div_up(mc->nc * 8, 256), 256, 0, L, C, sc, m->b, cur, mc->b, (int)mc->nc);
    }
    // hybrid: every rank restricted the planes it owns; with all ranks' coarse planes gathered, levels >= 1 run replicated
    if (fine && h->hybrid && !(v4 && h->gx_on)) { int rc = dist_gather_coarse(h, true); if (rc) return rc; }
    float* ec = nullptr;
    // experiment switches (default: the constants above): FSIM_MG_W_FIRST / FSIM_MG_W_LAST move the doubly visited levels
    static const int w_first = getenv("FSIM_MG_W_FIRST") ? atoi(getenv("FSIM_MG_W_FIRST")) : W_FIRST;
    static const int w_last = getenv("FSIM_MG_W_LAST") ? atoi(getenv("FSIM_MG_W_LAST")) : W_LAST;
    const int visits = (l + 1 >= w_first && l + 1 <= w_last && l + 1 < (int)h->mg.size() - 1) ? 2 : 1;
    for (int v = 0; v < visits; v++) {
        int rc = cycle(h, l + 1, v == 0, &ec);
        if (rc) return rc;
        if (ec != mc->xa) {  // keep the running coarse iterate in xa so that a second visit continues from it
            float* t = mc->xa; mc->xa = mc->xb; mc->xb = t;
        }
    }
    {
        std::unique_ptr<KScope> ks(new KScope(h, kid, POST));
        if (v4) launch_k(h, mg_prolong_jacobi4_kernel, grd4, blk4, 0, L, C, sc, m->b, cur, mc->xa, oth, push_of(oth), zrot_first);
        else if (fine) launch_k(h, mg_prolong_jacobi_kernel<true>, grdL, blk, 0, L, C, sc, m->b, cur, mc->xa, oth);
        else if (c4 && mc->gx % 2 == 0 && mc->pad % 2 == 0) launch_k(h, mg_prolong_jacobic4_kernel, grdC4, blk4, 0, L, C, sc, m->b, cur, mc->xa, oth);
        else launch_k(h, mg_prolong_jacobi_kernel<false>, grdL, blk, 0, L, C, sc, m->b, cur, mc->xa, oth);
        { float* t = cur; cur = oth; oth = t; }
        for (int s = 1; s < POST; s++) {
            const float om = OM_A;  // post-sweeps run the pre-sweep weights in reverse order (B in the fused prolongation sweep, then A)
            if (fine && h->hybrid && !fx) {
                ks.reset();
                int rc = dist_halo_sym(h, SYM_X, cur, true);
                if (rc) return rc;
                ks.reset(new KScope(h, kid, 0));
            }
            if (v4 && with_dot && s == POST - 1) {
                const int ntiles = (int)(grd4.x * grd4.y * grd4.z);
                // persistent CTAs walk the tiles with stride gridDim: the stride must not share a factor with the tile grid's x
                // extent, or a CTA only ever sees tiles of one x-column -- in a dam-break scene half the CTAs then got nothing but
                // air tiles and the other half all the work (64 us against 37 us for the same sweep without the dot product)
                int nper = std::min(ntiles, h->sm_count * 8);
                while (nper > 1 && (int)grd4.x > 1 && nper % (int)grd4.x == 0) nper--;
                const int layer = (int)(grd4.x * grd4.y);
                launch_k(h, mg_jacobi4_dot_kernel, nper, 256, 0, L, h->scal, m->b, cur, oth, h->partials, h->red_counter, om,
                                                                                               (int)grd4.x, (int)grd4.y, ntiles, wait_of(cur), (fx && ntiles > layer) ? layer : 0, h->ar_dev);
            } else if (v4) launch_k(h, mg_jacobi4_kernel, grd4, blk4, 0, L, h->scal, m->b, cur, oth, om, wait_of(cur), no_push, 0);
            else if (fine) launch_k(h, mg_jacobi_kernel<true>, grdL, blk, 0, L, sc, m->b, cur, oth, om);
            else if (c4) launch_k(h, mg_jacobic4_kernel, grdC4, blk4, 0, L, sc, m->b, cur, oth, om);
            else launch_k(h, mg_jacobi_kernel<false>, grdL, blk, 0, L, sc, m->b, cur, oth, om);
            float* t = cur; cur = oth; oth = t;
        }
    }
    *result = cur;
    if (fx_rc) return fsim_fail(h, FSIM_ERR_COMM, "fused slab exchange: a level-0 multigrid array is not published to the neighbours");
    return FSIM_OK;
}

}  // namespace

bool mg_enabled(const fsim* h) { return h->use_mg; }

void mg_free(fsim* h) {
    for (MgLevel* m : h->mg) {
        for (int k = 0; k < 7; k++) cudaFree(m->base[k]);
        delete m;
    }
    h->mg.clear();
}

// (re)builds the Galerkin hierarchy for the current cell flags; arrays are allocated once per grid
float* mg_level_array(fsim* h, int level, int which, float** base, size_t* pad) {
    if (level >= (int)h->mg.size()) { if (base) *base = nullptr; if (pad) *pad = 0; return nullptr; }
    MgLevel* m = h->mg[level];
    float* p = which == 0 ? m->xa : (which == 1 ? m->xb : m->b);
    // xa / xb may have been swapped since allocation: find the allocation that holds p
    for (int k = 4; k < 7; k++)
        if (m->base[k] && p == m->base[k] + m->pad) { if (base) *base = m->base[k]; if (pad) *pad = m->pad; return p; }
    if (base) *base = p;
    if (pad) *pad = 0;
    return p;
}

int mg_alloc(fsim* h) {
    if (h->mg.empty()) {
        int gx = h->g.gx, gy = h->g.gy, gz = h->g.gz;
        for (int l = 0;; l++) {
            MgLevel* m = new MgLevel();
            m->gx = gx; m->gy = gy; m->gz = gz;
            h->mg.push_back(m);
            int rc = alloc_level(h, m, l == 0);
            if (rc) return rc;
            if (l > 0 && m->nc <= COARSE_MAX) break;
            gx = (gx + 1) / 2; gy = (gy + 1) / 2; gz = (gz + 1) / 2;
        }
        // first level (>= 1) that is small enough for the single-cluster tail kernel
        const int nl = (int)h->mg.size();
        h->mg_tail_first = nl - 1;
        for (int l = 1; l < nl; l++)
            if (h->mg[l]->nc <= TAIL_MAX_CELLS && nl - l <= TAIL_MAX_LEVELS) { h->mg_tail_first = l; break; }
    }
    return FSIM_OK;
}

int mg_build(fsim* h) {
    const dim3 blk(32, 4, 2);
    int rc0 = mg_alloc(h);
    if (rc0) return rc0;
    for (size_t l = 1; l < h->mg.size(); l++) {
        MgLevel *f = h->mg[l - 1], *c = h->mg[l];
        KScope ks(h, K_MG);
        Lv f0 = view(h, f, 0);
        if (h->code_full) f0.code = h->code_full;  // hybrid: the coarse operators are global, h->code only marks this rank's planes
        if (l == 1) launch_k(h, mg_build1_kernel, grid_of(c, blk), blk, 0, f0, view(h, c, 1), c->wx, c->wy, c->wz, c->diag);
        else launch_k(h, mg_buildn_kernel, grid_of(c, blk), blk, 0, view(h, f, (int)l - 1), view(h, c, (int)l), c->wx, c->wy, c->wz, c->diag);
    }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

// the fused paths need the float4 level-0 kernels
// (hybrid slab projection included: the fused kernels then leave their partial reductions in PcgScalars::loc for the all-rank
// reduction; the slab-local solve of FSIM_SLAB_SOLVER=distributed keeps the separate kernels)
bool mg_can_fuse(const fsim* h) { return h->use_mg && !h->dist && !h->mg.empty() && h->mg[0]->gx % 4 == 0 && h->g.nc % 4 == 0; }

// CG update (p, r, ||r||_inf, convergence flags) fused with the first smoothing sweep of the cycle that follows
int mg_update_first(fsim* h) {
    MgLevel* m = h->mg[0];
    const Lv L = view(h, m, 0);
    KScope ks(h, K_UPDATE);
    { Tile4 t4 = h->hybrid ? tile4_make(h->g.gx, (int64_t)h->g.zown0 * h->g.gy, (int64_t)(h->g.zown1 - h->g.zown0) * h->g.gy, h->g.gy)
                         : tile4_make(h->g.gx, 0, h->g.nc / h->g.gx, h->g.gy);
    FxPush fp; FxWait fw;
    memset(&fp, 0, sizeof(fp));
    if (h->hybrid && h->fx_on) {  // fused exchange of the first iterate (fexch.cuh): boundary planes first
        if (!dist_fx(h, SYM_X, m->xa, &fp, &fw)) return fsim_fail(h, FSIM_ERR_COMM, "fused slab exchange: the multigrid iterate is not published to the neighbours");
        t4 = tile4_boundary_first(t4);
        fp.nblk[0] = fp.nblk[1] = (uint32_t)(t4.nbx * (h->g.gy / T4_ROWS));
    }
    launch_k(h, mg_update_first4_kernel, tile4_blocks(t4), 256, 0, L, t4, h->scal, h->status_dev, h->p, h->s, h->r, h->q, m->b,
                                                                              m->xa, h->partials, h->red_counter, fp, h->ar_dev); }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

// z32 = M^-1 r : one cycle from a zero guess on the fp64 CG residual h->r.  first_done: b and the first sweep were
// produced by mg_update_first; with_dot: the last sweep also writes sigma' = z.r
int mg_apply(fsim* h, bool first_done, bool with_dot) {
    float* res = nullptr;
    int rc = cycle(h, 0, true, &res, first_done, with_dot);
    if (rc) return rc;
    h->mg_z32 = res;
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}
