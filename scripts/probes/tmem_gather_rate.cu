// Probe: per-SM throughput of single-column TMEM loads at data-dependent columns ("gathers from a TMEM-resident table")
// against LDS.32 gathers from a shared-memory table, alone and mixed.  One CTA per SM; clock64 inside the kernel.
// mode 0: 8 x tcgen05.ld.32x32b.x1 (random column) per iteration      mode 1: 8 x LDS.32 (random 128-byte row)
// mode 2: 4 + 4 interleaved                                           mode 3: tcgen05.ld.32x32b.x8 at a random column
// mode 4: 4 x LDTM.x1 + 4 x LDS + 1 x LDTM.x8 (the mix a hybrid edge kernel would issue per 4 slots)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probes/tmem_gather_rate scripts/probes/tmem_gather_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    return v;
}
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

__global__ void __launch_bounds__(1024, 1) probe(int mode, int iters, long long* cyc, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t base_s;
    float* tab = reinterpret_cast<float*>(smem);          // 256 rows x 32 floats
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) tab[i] = (float)i;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = base_s + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {
        for (int c = 0; c < 512; ++c) {
            uint32_t v = c * 128 + warp * 32 + lane;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + c), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t s = 0x9e3779b9u * (warp + 1), acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        uint32_t v[8], w[8];
        if (mode == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ld1(tm + ((s >> (3 * u + 4)) & 255u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
        } else if (mode == 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __float_as_uint(tab[((s >> (3 * u + 4)) & 255u) * 32 + lane]);
        } else if (mode == 2) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = ld1(tm + ((s >> (3 * u + 4)) & 255u));
                v[4 + u] = __float_as_uint(tab[((s >> (3 * u + 16)) & 255u) * 32 + lane]);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]));
        } else if (mode == 3) {
            ld8(tm + ((s >> 4) & 0xf8u), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
        } else {
            ld8(tm + 256u + ((s >> 4) & 0xf8u), w);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = ld1(tm + ((s >> (3 * u + 4)) & 255u));
                v[4 + u] = __float_as_uint(tab[((s >> (3 * u + 16)) & 255u) * 32 + lane]);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]));
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += w[u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "r"(512u) : "memory");
}

int main() {
    long long* cyc;
    uint32_t* sink;
    cudaMalloc(&cyc, 148 * 8);
    cudaMalloc(&sink, 148 * 1024 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 20000;
    const char* names[5] = {"LDTM.x1 gather", "LDS.32 gather", "4 LDTM.x1 + 4 LDS", "LDTM.x8", "4 LDTM.x1 + 4 LDS + LDTM.x8"};
    for (int mode = 0; mode < 5; ++mode)
        for (int nw : {4, 8, 12, 16, 24, 32}) {
            probe<<<148, nw * 32, 200 * 1024>>>(mode, iters, cyc, sink);   // 200 KB of dynamic shared memory: one CTA per SM
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double c = 0;
            for (int i = 0; i < 148; ++i) c += h[i];
            c /= 148;
            const double per_it = c / iters;   // cycles per iteration of warp 0 (all warps run the same loop concurrently)
            const double loads = mode == 3 ? 1.0 : (mode == 4 ? 9.0 : 8.0);
            const double bytes = (mode == 3 ? 8.0 : mode == 4 ? 16.0 : 8.0) * 128.0 * nw;   // bytes returned per iteration over all warps
            printf("%-30s warps %2d: %8.1f cyc/iter  %6.2f cyc per warp-load (SM-wide)  %7.1f B/clk/SM\n", names[mode], nw, per_it,
                   per_it / (loads * nw), bytes / per_it);
        }
    return 0;
}
