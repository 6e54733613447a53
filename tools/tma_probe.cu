// tools/tma_probe.cu -- stand-alone check of the 3-D TMA tile load the G2P staging uses (g2p_core.cuh): box width, negative
// start coordinates (halo of border tiles, zero fill), tensor map as a top-level __grid_constant__ parameter or nested in a
// struct.  Prints one line per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tma_probe tools/tma_probe.cu && tools/_bin/tma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Nested {
    int pad[5];
    alignas(64) CUtensorMap tm;
};

template <int BOXW>
__device__ __forceinline__ void body(const CUtensorMap* tm, int c0, int c1, int c2, float* out, int* status) {
    __shared__ alignas(128) float s[BOXW * 6 * 6 + 32];
    __shared__ alignas(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(BOXW * 6 * 6 * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(s)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0, spins = 0;
    while (!ok && spins < (1u << 22)) {
        spins++;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    if (threadIdx.x == 0) { status[0] = ok; status[1] = spins; }
    if (ok)
        for (int i = threadIdx.x; i < BOXW * 6 * 6; i += blockDim.x) out[i] = s[i];
}

template <int BOXW>
__global__ void k_top(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, float* out, int* status) { body<BOXW>(&tm, c0, c1, c2, out, status); }
template <int BOXW>
__global__ void k_nested(const __grid_constant__ Nested n, int c0, int c1, int c2, float* out, int* status) { body<BOXW>(&n.tm, c0, c1, c2, out, status); }

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// usage: tma_probe BOXW C0 C1 C2 NESTED   (one variant per process: a faulting variant poisons the context)
int main(int argc, char** argv) {
    const int boxw = argc > 1 ? atoi(argv[1]) : 36;
    const int c0 = argc > 2 ? atoi(argv[2]) : -1, c1 = argc > 3 ? atoi(argv[3]) : -1, c2 = argc > 4 ? atoi(argv[4]) : -1;
    const bool nested = argc > 5 ? atoi(argv[5]) != 0 : true;
    const int gx = 64, gy = 40, gz = 24;
    std::vector<float> h((size_t)gx * gy * gz);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i + 1);
    float *d, *out;
    int* st;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 40 * 36 * 4); cudaMalloc(&st, 8);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (!fn) { printf("no entry point: %s\n", cudaGetErrorString(e)); return 1; }
    CUtensorMap tm;
    const cuuint64_t dims[3] = {gx, gy, gz}, strides[2] = {(cuuint64_t)gx * 4, (cuuint64_t)gx * gy * 4};
    const cuuint32_t box[3] = {(cuuint32_t)boxw, 6, 6}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    Nested n;
    memset(&n, 0, sizeof(n));
    memcpy(&n.tm, &tm, sizeof(tm));
    cudaMemset(st, 0, 8); cudaMemset(out, 0, 40 * 36 * 4);
    if (boxw == 32) { if (nested) k_nested<32><<<1, 128>>>(n, c0, c1, c2, out, st); else k_top<32><<<1, 128>>>(tm, c0, c1, c2, out, st); }
    else if (boxw == 40) { if (nested) k_nested<40><<<1, 128>>>(n, c0, c1, c2, out, st); else k_top<40><<<1, 128>>>(tm, c0, c1, c2, out, st); }
    else { if (nested) k_nested<36><<<1, 128>>>(n, c0, c1, c2, out, st); else k_top<36><<<1, 128>>>(tm, c0, c1, c2, out, st); }
    e = cudaDeviceSynchronize();
    int hs[2] = {0, 0};
    std::vector<float> ho(boxw * 36, 0.f);
    if (e == cudaSuccess) { cudaMemcpy(hs, st, 8, cudaMemcpyDeviceToHost); cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost); }
    // check every element of the box against the source (0 outside the grid)
    int bad = 0;
    for (int k = 0; k < 6; k++) for (int j = 0; j < 6; j++) for (int i = 0; i < boxw; i++) {
        const int x = c0 + i, y = c1 + j, z = c2 + k;
        const float want = (x >= 0 && x < gx && y >= 0 && y < gy && z >= 0 && z < gz) ? h[((size_t)z * gy + y) * gx + x] : 0.f;
        bad += ho[(k * 6 + j) * boxw + i] != want;
    }
    printf("box %d start (%d,%d,%d) %s: encode %d, %s, completed=%d spins=%d, wrong elements=%d\n", boxw, c0, c1, c2, nested ? "nested" : "top", (int)r,
           cudaGetErrorString(e), hs[0], hs[1], e == cudaSuccess ? bad : -1);
    return e != cudaSuccess;
}
