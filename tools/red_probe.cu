// tools/red_probe.cu -- micro-benchmark behind the P2G flush design: what does a coalesced fp32 reduction
// (RED.ADD.F32, 32 consecutive floats per warp instruction) cost on sm_100a, and do the vector forms (red.global.add.v2.f32 /
// .v4.f32, sm_90+) move 2x / 4x the data per instruction for the same price?  Also plain stores as the floor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/red_probe tools/red_probe.cu && tools/_bin/red_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red1(float* p, float a) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory"); }
__device__ __forceinline__ void red2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE 0: red f32, 1: red v2, 2: red v4, 3: st f32, 4: st v2.  Every warp walks rows of `width` elements (width lanes active)
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* buf, size_t nrows, int width, int iters, int row_floats) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
    size_t row = warp;
    for (int it = 0; it < iters; it++) {
        float* p = buf + row * row_floats;
        for (int x = lane; x < width; x += 32) {
            if (MODE == 0) red1(p + x, 1.f);
            if (MODE == 1) red2(p + 2 * x, 1.f, 2.f);
            if (MODE == 2) red4(p + 4 * x, 1.f, 2.f, 3.f, 4.f);
            if (MODE == 3) p[x] = 1.f;
            if (MODE == 4) *reinterpret_cast<float2*>(p + 2 * x) = make_float2(1.f, 2.f);
        }
        row += nwarps;
        if (row >= nrows) row -= nrows;
    }
}

template <int MODE>
void run(const char* name, float* buf, size_t bytes, int width, int elem_floats) {
    const int row_floats = ((width * elem_floats + 31) / 32) * 32;
    const size_t nrows = bytes / sizeof(float) / row_floats;
    const int iters = 2048, blocks = 148 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(buf, nrows, width, 64, row_floats);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(buf, nrows, width, iters, row_floats);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)blocks * 8 * iters;
    const double lanes = rows * width;
    printf("%-14s width %3d  footprint %6.0f MB : %8.3f ms  %7.3f ns/row/SM-equivalent  %6.3f cyc/lane/SM @1.92GHz  %7.1f GB/s payload\n", name, width,
           bytes / 1e6, ms, ms * 1e6 / (rows / 148), ms * 1e-3 * 1.92e9 / (lanes / 148), lanes * elem_floats * 4 / (ms * 1e-3) / 1e9);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    float* buf;
    const size_t big = (size_t)1 << 30;
    cudaMalloc(&buf, big);
    cudaMemset(buf, 0, big);
    for (size_t bytes : {(size_t)64 << 20, big}) {
        for (int width : {32, 34}) {
            run<0>("red.f32", buf, bytes, width, 1);
            run<1>("red.v2.f32", buf, bytes, width, 2);
            run<2>("red.v4.f32", buf, bytes, width, 4);
            run<3>("st.f32", buf, bytes, width, 1);
            run<4>("st.v2.f32", buf, bytes, width, 2);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
