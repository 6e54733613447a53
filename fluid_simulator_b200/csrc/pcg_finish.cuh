// pcg_finish.cuh -- what the PCG does with a finished reduction (the scalar half of bridsonSolverGrid.cpp:254-292).
// Single-GPU: called by the last block of the reducing kernel (pcg.cu).  Slab mode: the reducing kernel only stores this
// rank's partial results in PcgScalars::loc and dist_allreduce_kernel (dist.cu) calls the same function with the
// all-rank values, so every rank takes the same decision from bit-identical numbers.
#pragma once
#include "fsim_internal.h"

__device__ __forceinline__ void pcg_finish_rhs(PcgScalars* sc, PcgHostStatus* st, double sumsq, double cells) {
    sc->rhs_sumsq = sumsq;
    sc->fluid_cells = (long long)(cells + 0.5);
    sc->early_out = sumsq < 1e-7;  // bridsonSolverGrid.cpp:254-258
    sc->done = sc->early_out;
    // incompressibilityMaxIterationCount <= 0: the reference's loop body never runs (bridsonSolverGrid.cpp:267) and p = 0 is
    // applied; done = 2 keeps the device-side loop from running its first iteration (k_project starts such a solve cold)
    if (!sc->done && sc->max_it <= 0) sc->done = 2;
    sc->iterations = 0;
    sc->nan_break = 0;
    sc->rmax = 0.0;
    sc->sigma = 0.0;
    sc->it = 0;
    st->done = sc->done;
    st->it_done = 0;
    __threadfence_system();
}

__device__ __forceinline__ void pcg_finish_residual(PcgScalars* sc, PcgHostStatus* st, double r0max) {
    sc->r0max = r0max;
    if (r0max < sc->tol) {  // last step's pressure already satisfies the tolerance
        sc->done = 1;
        sc->rmax = r0max;
        st->done = 1;
        __threadfence_system();
    }
}

// after p += alpha s ; r -= alpha q: zr = z.r (diagonal preconditioner only), rmax = ||r||_inf   (:270-284)
__device__ __forceinline__ void pcg_finish_update(PcgScalars* sc, PcgHostStatus* st, double zr, double rmax) {
    const double alpha = sc->sigma / sc->sq;
    const bool bad = alpha != alpha;  // NaN => the reference breaks before touching p (:271-272)
    const int it = sc->it;
    if (bad) {
        sc->nan_break = 1;
        sc->done = 1;
        sc->iterations = it;
    } else {
        sc->rmax = rmax;
        sc->sigma_new = zr;
        if (rmax < sc->tol) {  // converged inside iteration `it` => the reference returns it (:280-281, 292)
            sc->done = 1;
            sc->iterations = it;
        } else if (it + 1 >= sc->max_it) {
            sc->done = 2;  // iteration cap; the direction update is skipped like the loop exit would
            sc->iterations = sc->max_it;
        }
    }
    if (sc->done) { st->done = sc->done; __threadfence_system(); }
}
