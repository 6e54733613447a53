// pcg.cu -- pressure projection: BridsonSolverGrid::solveIncompressibility (bridsonSolverGrid.cpp:244-293).
//
// Same linear system and stopping rule as the reference (RHS :79-89, 7-point matrix :40-77, early-out when
// sum(rhs^2) < 1e-7 :254-258, stop when ||r||_inf < residualTolerance :276-281, NaN guard :271), but
//   * matrix-free: coefficients are recomputed from the 1-byte cell flags inside the SpMV, the reference's
//     compacted fluid-cell list / AMatrixRow array / cell.id indirection do not exist;
//   * the sequential MIC(0) preconditioner (:91-163) is replaced by a parallel one (multigrid V-cycle, mg.cu;
//     or diagonal scaling) -- equivalence is on the converged pressure field;
//   * CG vectors (p, r, s, q) are fp64: with |p| ~ 1e4 and a 1e-6 absolute tolerance an fp32 recurrence stalls
//     (SURVEY.md §7 "hard parts"); the preconditioner works in fp32;
//   * SpMV is fused with the s.q reduction, the x/r update with ||r||_inf and (diagonal case) z.r; reductions are
//     two-level (warp shuffle -> block -> fixed-order last-block sum) => deterministic, no fp64 atomics;
//   * all scalars (sigma, alpha, beta, done flag, iteration count) stay on the device; the host only polls the
//     done flag every few iterations.
#include "fsim_internal.h"

int mg_apply(fsim* h);  // z = M^-1 r   (mg.cu)
bool mg_enabled(const fsim* h);

namespace {

constexpr int PT = 256;  // threads per block of the persistent solver kernels

struct PcgArgs {
    GridDims g;
    const uint8_t* flags;
    const float* u2[3];
    const float* dens;
    double *p, *rhs, *r, *s, *q, *z;
    const float* z32;  // multigrid result (fp32) when the multigrid preconditioner is active, else nullptr
    PcgScalars* sc;
    double* partials;  // [3][gridDim.x]
    unsigned int* counter;
    double inv_h, scale;  // 1/h ; dt/(rho h^2)
    double avg_pressure, pressure_k, tol;
    int pressure_enabled, max_it, it;
};

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

// block-level reduction of up to 3 values (sum, sum, max); returns true in exactly one block (the last to arrive),
// whose thread 0 then holds the grid-wide totals in out[].
template <int NS, int NM>
__device__ __forceinline__ bool grid_reduce(double* vals, double* partials, unsigned int* counter, double* out) {
    __shared__ double sh[NS + NM][PT / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS + NM; k++) {
        const double v = k < NS ? warp_sum(vals[k]) : warp_max(vals[k]);
        if (lane == 0) sh[k][w] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NS + NM; k++) {
            double v = sh[k][0];
            for (int i = 1; i < PT / 32; i++) v = k < NS ? v + sh[k][i] : fmax(v, sh[k][i]);
            partials[(size_t)k * gridDim.x + blockIdx.x] = v;
        }
        __threadfence();
        const unsigned t = atomicInc(counter, gridDim.x - 1);  // wraps back to 0 => self-resetting
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    // fixed-order final sum by one warp => bitwise reproducible
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NS + NM; k++) {
            double v = k < NS ? 0.0 : 0.0;
            for (unsigned i = lane; i < gridDim.x; i += 32) {
                const double x = ((volatile double*)partials)[(size_t)k * gridDim.x + i];
                v = k < NS ? v + x : fmax(v, x);
            }
            v = k < NS ? warp_sum(v) : warp_max(v);
            if (lane == 0) out[k] = v;
        }
    }
    return threadIdx.x == 0;
}

__device__ __forceinline__ int type_of(const uint8_t* flags, int64_t c) { return flags[c] & FL_TYPE_MASK; }

// diagonal / neighbour structure of row c (calculateAMatrix, bridsonSolverGrid.cpp:40-77): returns #non-solid nbrs and a
// 6-bit mask of WATER neighbours (-x,+x,-y,+y,-z,+z).  WATER cells are always interior, so no bounds checks.
__device__ __forceinline__ int row_structure(const PcgArgs& a, int64_t c, unsigned* water) {
    const int64_t nb[6] = {c - 1, c + 1, c - a.g.sy, c + a.g.sy, c - a.g.sz, c + a.g.sz};
    int nonsolid = 0;
    unsigned wm = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int t = type_of(a.flags, nb[k]);
        nonsolid += (t != FSIM_CELL_SOLID);
        wm |= (t == FSIM_CELL_WATER) ? (1u << k) : 0u;
    }
    *water = wm;
    return nonsolid;
}

// calculateRHS (:79-89) + p = 0, r = rhs + sum(rhs^2) (:254-258)
__global__ void __launch_bounds__(PT) rhs_kernel(PcgArgs a) {
    double acc[2] = {0.0, 0.0};
    for (int64_t c = (int64_t)blockIdx.x * PT + threadIdx.x; c < a.g.nc; c += (int64_t)gridDim.x * PT) {
        double rhs = 0.0;
        if (type_of(a.flags, c) == FSIM_CELL_WATER) {
            const double div = ((double)a.u2[0][c] + (double)a.u2[1][c] + (double)a.u2[2][c]) - (double)a.u2[0][c - 1] -
                               (double)a.u2[1][c - a.g.sy] - (double)a.u2[2][c - a.g.sz];
            rhs = -a.inv_h * div + (a.pressure_enabled ? ((double)a.dens[c] - a.avg_pressure) * a.pressure_k : 0.0);
            acc[0] += rhs * rhs;
            acc[1] += 1.0;
        }
        a.rhs[c] = rhs;
        a.r[c] = rhs;
        a.p[c] = 0.0;
    }
    double out[2];
    if (grid_reduce<2, 0>(acc, a.partials, a.counter, out)) {
        a.sc->rhs_sumsq = out[0];
        a.sc->fluid_cells = (long long)(out[1] + 0.5);
        a.sc->early_out = out[0] < 1e-7;
        a.sc->done = a.sc->early_out;
        a.sc->iterations = 0;
        a.sc->nan_break = 0;
        a.sc->rmax = 0.0;
        a.sc->sigma = 0.0;
    }
}

// diagonal preconditioner: z = r / A_ii ; s = z ; sigma = z.r   (first application, :262-265)
__global__ void __launch_bounds__(PT) jacobi_init_kernel(PcgArgs a) {
    if (a.sc->done) return;
    double acc[1] = {0.0};
    for (int64_t c = (int64_t)blockIdx.x * PT + threadIdx.x; c < a.g.nc; c += (int64_t)gridDim.x * PT) {
        double z = 0.0;
        if (type_of(a.flags, c) == FSIM_CELL_WATER) {
            unsigned wm;
            const int ns = row_structure(a, c, &wm);
            const double r = a.r[c];
            z = ns > 0 ? r / (a.scale * ns) : r;
            acc[0] += z * r;
        }
        a.z[c] = z;
        a.s[c] = z;
    }
    double out[1];
    if (grid_reduce<1, 0>(acc, a.partials, a.counter, out)) a.sc->sigma = out[0];
}

// generic first application: s = z ; sigma = z.r  (z supplied by the multigrid preconditioner)
__global__ void __launch_bounds__(PT) start_kernel(PcgArgs a) {
    if (a.sc->done) return;
    double acc[1] = {0.0};
    for (int64_t c = (int64_t)blockIdx.x * PT + threadIdx.x; c < a.g.nc; c += (int64_t)gridDim.x * PT) {
        const double z = (double)a.z32[c];
        a.s[c] = z;
        acc[0] += z * a.r[c];
    }
    double out[1];
    if (grid_reduce<1, 0>(acc, a.partials, a.counter, out)) a.sc->sigma = out[0];
}

// q = A s fused with s.q  (applyAMatrix :165-198 + dotProduct :200-214)
__global__ void __launch_bounds__(PT) spmv_kernel(PcgArgs a) {
    if (a.sc->done) return;
    double acc[1] = {0.0};
    for (int64_t c = (int64_t)blockIdx.x * PT + threadIdx.x; c < a.g.nc; c += (int64_t)gridDim.x * PT) {
        if (type_of(a.flags, c) != FSIM_CELL_WATER) continue;
        unsigned wm;
        const int ns = row_structure(a, c, &wm);
        const double sc = a.s[c];
        double nsum = 0.0;
        if (wm & 1u) nsum += a.s[c - 1];
        if (wm & 2u) nsum += a.s[c + 1];
        if (wm & 4u) nsum += a.s[c - a.g.sy];
        if (wm & 8u) nsum += a.s[c + a.g.sy];
        if (wm & 16u) nsum += a.s[c - a.g.sz];
        if (wm & 32u) nsum += a.s[c + a.g.sz];
        const double q = a.scale * ((double)ns * sc - nsum);
        a.q[c] = q;
        acc[0] += sc * q;
    }
    double out[1];
    if (grid_reduce<1, 0>(acc, a.partials, a.counter, out)) a.sc->sq = out[0];
}

// alpha = sigma / s.q ; p += alpha s ; r -= alpha q ; ||r||_inf ; (JACOBI: z = r / A_ii ; sigma' = z.r)   (:270-284)
template <bool JACOBI>
__global__ void __launch_bounds__(PT) update_kernel(PcgArgs a) {
    if (a.sc->done) return;
    const double alpha = a.sc->sigma / a.sc->sq;
    const bool bad = alpha != alpha;  // NaN => the reference breaks before touching p (:271-272)
    double acc[2] = {0.0, 0.0};      // [0] = z.r (sum), [1] = max |r|
    if (!bad) {
        for (int64_t c = (int64_t)blockIdx.x * PT + threadIdx.x; c < a.g.nc; c += (int64_t)gridDim.x * PT) {
            if (type_of(a.flags, c) != FSIM_CELL_WATER) continue;
            a.p[c] += alpha * a.s[c];
            const double r = a.r[c] - alpha * a.q[c];
            a.r[c] = r;
            acc[1] = fmax(acc[1], fabs(r));
            if (JACOBI) {
                unsigned wm;
                const int ns = row_structure(a, c, &wm);
                const double z = ns > 0 ? r / (a.scale * ns) : r;
                a.z[c] = z;
                acc[0] += z * r;
            }
        }
    }
    double out[2];
    if (grid_reduce<1, 1>(acc, a.partials, a.counter, out)) {
        if (bad) {
            a.sc->nan_break = 1;
            a.sc->done = 1;
            a.sc->iterations = a.it;
        } else {
            a.sc->rmax = out[1];
            a.sc->sigma_new = out[0];
            if (out[1] < a.tol) {  // converged inside iteration `it` => the reference returns it (:280-281, 292)
                a.sc->done = 1;
                a.sc->iterations = a.it;
            } else if (a.it + 1 >= a.max_it) {
                a.sc->done = 2;  // iteration cap; the direction update below is skipped like the loop exit would
                a.sc->iterations = a.max_it;
            }
        }
    }
}

// sigma' = z.r for an externally computed z (multigrid)
__global__ void __launch_bounds__(PT) dot_zr_kernel(PcgArgs a) {
    if (a.sc->done) return;
    double acc[1] = {0.0};
    for (int64_t c = (int64_t)blockIdx.x * PT + threadIdx.x; c < a.g.nc; c += (int64_t)gridDim.x * PT)
        acc[0] += (double)a.z32[c] * a.r[c];
    double out[1];
    if (grid_reduce<1, 0>(acc, a.partials, a.counter, out)) a.sc->sigma_new = out[0];
}

// beta = sigma'/sigma ; s = z + beta s ; sigma = sigma'   (:284-289)
__global__ void __launch_bounds__(PT) direction_kernel(PcgArgs a) {
    if (a.sc->done) return;
    const double beta = a.sc->sigma_new / a.sc->sigma;
    for (int64_t c = (int64_t)blockIdx.x * PT + threadIdx.x; c < a.g.nc; c += (int64_t)gridDim.x * PT) {
        if (type_of(a.flags, c) != FSIM_CELL_WATER) continue;
        a.s[c] = a.s[c] * beta + (a.z32 ? (double)a.z32[c] : a.z[c]);
    }
    // the last block to finish publishes sigma = sigma' (every block has read both scalars before arriving)
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicInc(a.counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) a.sc->sigma = a.sc->sigma_new;
}

}  // namespace

int k_project(fsim* h, double dt, int* iterations) {
    const GridDims& g = h->g;
    PcgArgs a;
    a.g = g; a.flags = h->flags;
    for (int ax = 0; ax < 3; ax++) a.u2[ax] = h->u2[ax];
    a.dens = h->dens;
    a.p = h->p; a.rhs = h->rhs; a.r = h->r; a.s = h->s; a.q = h->q; a.z = h->z;
    a.sc = h->scal; a.partials = h->partials; a.counter = h->red_counter;
    const double hcell = h->info.cell_d[0];
    a.inv_h = 1.0 / hcell;
    a.scale = dt / (h->par.fluid_density * hcell * hcell);
    a.avg_pressure = h->par.average_pressure; a.pressure_k = h->par.pressure_k;
    a.tol = h->par.residual_tolerance;
    a.pressure_enabled = h->par.pressure_enabled;
    a.max_it = h->par.max_iterations;
    a.it = 0;
    a.z32 = nullptr;
    h->mg_inv_scale = 1.0 / a.scale;
    const int nb = h->red_blocks;
    const bool use_mg = mg_enabled(h);

    { KScope ks(h, K_RHS); rhs_kernel<<<nb, PT, 0, h->stream>>>(a); }
    if (use_mg) {
        int rc = mg_build(h);
        if (rc) return rc;
        rc = mg_apply(h);
        if (rc) return rc;
        a.z32 = h->mg_z32;
        KScope ks(h, K_PCG_INIT);
        start_kernel<<<nb, PT, 0, h->stream>>>(a);
    } else {
        KScope ks(h, K_PCG_INIT);
        jacobi_init_kernel<<<nb, PT, 0, h->stream>>>(a);
    }
    FSIM_CHECK_LAUNCH(h);

    const int poll = use_mg ? 1 : 16;
    int done = 0;
    for (int it = 0; it < a.max_it && !done; it++) {
        a.it = it;
        { KScope ks(h, K_SPMV); spmv_kernel<<<nb, PT, 0, h->stream>>>(a); }
        if (use_mg) {
            { KScope ks(h, K_UPDATE); update_kernel<false><<<nb, PT, 0, h->stream>>>(a); }
            int rc = mg_apply(h);
            if (rc) return rc;
            a.z32 = h->mg_z32;
            KScope ks(h, K_UPDATE);
            dot_zr_kernel<<<nb, PT, 0, h->stream>>>(a);
        } else {
            KScope ks(h, K_UPDATE);
            update_kernel<true><<<nb, PT, 0, h->stream>>>(a);
        }
        { KScope ks(h, K_DIRECTION); direction_kernel<<<nb, PT, 0, h->stream>>>(a); }
        if ((it + 1) % poll == 0 || it + 1 == a.max_it) {
            FSIM_CUDA(h, cudaMemcpyAsync(h->scal_host, h->scal, sizeof(PcgScalars), cudaMemcpyDeviceToHost, h->stream));
            FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
            done = h->scal_host->done;
        }
    }
    FSIM_CHECK_LAUNCH(h);
    FSIM_CUDA(h, cudaMemcpyAsync(h->scal_host, h->scal, sizeof(PcgScalars), cudaMemcpyDeviceToHost, h->stream));
    FSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    const PcgScalars& s = *h->scal_host;
    h->solve.iterations = s.early_out ? 0 : (s.done ? s.iterations : a.max_it);
    h->solve.early_out = s.early_out;
    h->solve.rhs_sumsq = s.rhs_sumsq;
    h->solve.residual_max = s.rmax;
    h->solve.fluid_cells = s.fluid_cells;
    if (iterations) *iterations = h->solve.iterations;
    h->pressure_valid = !s.early_out;
    if (!s.early_out) return k_pressure_apply(h, dt);  // the early-out returns before applying anything (:257-258)
    return FSIM_OK;
}
