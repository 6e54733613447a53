// launch.cuh -- programmatic dependent launch (PDL) for the latency-bound solver kernels.
//
// One PCG iteration is ~24 short, strictly dependent kernels (pcg.cu, mg.cu); between two of them the GPU drains, the
// next grid is launched and its CTAs ramp up.  With the programmatic-stream-serialization launch attribute the next grid
// may be scheduled as soon as every CTA of the running one has executed `griddepcontrol.launch_dependents` (pdl_trigger, the
// first thing each solver kernel does), and its threads block in `griddepcontrol.wait` (pdl_wait) until the running grid has
// completed and its writes are visible -- so the launch latency and the CTA ramp-up overlap the predecessor's tail, while
// the data dependency stays exactly what stream order gives.  Both instructions are no-ops in a kernel that was launched
// without the attribute; stream capture turns the attribute into programmatic edges of the PCG loop graph.
#pragma once
#include <cuda_runtime.h>

#include <utility>

#include "fsim_internal.h"

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launches `kernel` on the handle's stream, as a programmatic dependent of whatever precedes it when h->pdl is set
template <class... KArgs, class... Args>
inline cudaError_t launch_k(fsim* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
