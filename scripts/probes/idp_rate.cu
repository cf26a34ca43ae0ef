// Probe: issue rate of dp2a (IDP.2A) against IMAD and PRMT on sm_100a — can one dp2a replace the PRMT + IMAD pair that turns
// a packed sender index into a shared-memory gather address?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o idp_rate scripts/probes/idp_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(uint32_t* out, uint32_t seed, int iters, long long* cyc) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i + 1);
    const uint32_t b = seed | 0x00a000a0u, c = seed ^ 0x280u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 2) asm volatile("prmt.b32 %0, %0, %1, 0x4440;" : "+r"(a[i]) : "r"(b));
            if (MODE == 3) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* what) {
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    for (int r = 0; r < 2; ++r) probe<MODE><<<148, 1024>>>(out, 12345u, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // 32 warps x 8 instructions x iters warp-instructions per SM
    printf("%-10s %6.2f warp-instructions per clock per SM\n", what, 32.0 * 8 * iters / (double)h[0]);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("dp2a.lo");
    run<3>("dp4a");
    run<1>("mad.lo");
    run<2>("prmt");
    return 0;
}
