// concurrency_probe.cu -- can a kernel on stream B run while a kernel on stream A spins on a flag B sets?  (what the
// in-process slab tests need from a single device)  Also: does cudaFuncSetAttribute / a first launch block meanwhile?
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
__global__ void spin(volatile unsigned* flag, unsigned* out, unsigned long long timeout_ns) {
    unsigned long long t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (*flag == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); if (t - t0 > timeout_ns) { *out = 2; return; } }
    *out = 1;
}
__global__ void setflag(volatile unsigned* flag) { *flag = 1; __threadfence_system(); }
__global__ void never_loaded_before(volatile unsigned* flag) { *flag = 1; __threadfence_system(); }
extern __shared__ float dyn[];
__global__ void big_smem(volatile unsigned* flag) { dyn[threadIdx.x] = 1.f; *flag = 1; __threadfence_system(); }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    unsigned *flag, *out; cudaMalloc(&flag, 4); cudaMalloc(&out, 4);
    unsigned h;
    // warm both kernels (loaded)
    cudaMemset(flag, 0, 4); setflag<<<1, 1, 0, b>>>(flag); spin<<<1, 1, 0, a>>>(flag, out, 1000000000ull); cudaDeviceSynchronize();
    for (int test = 0; test < 3; test++) {
        cudaMemset(flag, 0, 4); cudaMemset(out, 0, 4); cudaDeviceSynchronize();
        double t0 = now();
        spin<<<1, 1, 0, a>>>(flag, out, 3000000000ull);
        if (test == 0) setflag<<<1, 1, 0, b>>>(flag);                       // already loaded kernel
        if (test == 1) never_loaded_before<<<1, 1, 0, b>>>(flag);            // first launch of a kernel while another spins
        if (test == 2) { cudaFuncSetAttribute(big_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); double t1 = now();
                         printf("  cudaFuncSetAttribute returned after %.3f s\n", t1 - t0); big_smem<<<1, 32, 100 * 1024, b>>>(flag); }
        double t1 = now();
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
        printf("test %d: launch returned after %.3f s, total %.3f s, spin result %u (1 = saw the flag, 2 = timed out) err=%s\n", test, t1 - t0, now() - t0, h,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
