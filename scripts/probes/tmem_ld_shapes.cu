// Probe: which (lane, column) of TMEM does each thread's register receive for tcgen05.ld shapes .16x256b / .16x128b / .16x64b?
// TMEM is filled with value = lane * 1000 + column through the documented .32x32b shape (thread i <-> lane i).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_probe scripts/probes/tmem_ld_shapes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t* out) {
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = base_s;
    // fill 128 lanes x 32 columns: warp w writes its lane quarter
    {
        const uint32_t row = warp * 32 + lane;
        const uint32_t t0 = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 32; c += 4) {
            uint32_t v0 = row * 1000 + c, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t0 + c), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1) {   // lane quarter 1: lanes 32..63
        const uint32_t t0 = tmem + ((uint32_t)(32) << 16);
        uint32_t r[4];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) out[lane * 4 + i] = r[i];
        // second half of the quarter: lane offset 16
        const uint32_t t1 = tmem + ((uint32_t)(48) << 16) + 8;
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t1));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) out[128 + lane * 4 + i] = r[i];
        uint32_t s[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]) : "r"(t0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[256 + lane * 8 + i] = s[i];
        uint32_t u[2];
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(u[0]), "=r"(u[1]) : "r"(t0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 2; ++i) out[512 + lane * 2 + i] = u[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 1024 * 4);
    cudaMemset(d, 0xff, 1024 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    uint32_t h[1024];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("16x256b.x1 at lane 32, col 0: thread -> (lane,col) x4\n");
    for (int t = 0; t < 32; ++t) { printf("t%2d:", t); for (int i = 0; i < 4; ++i) printf(" (%u,%u)", h[t * 4 + i] / 1000, h[t * 4 + i] % 1000); printf("\n"); }
    printf("16x256b.x1 at lane 48, col 8:\n");
    for (int t = 0; t < 32; t += 5) { printf("t%2d:", t); for (int i = 0; i < 4; ++i) printf(" (%u,%u)", h[128 + t * 4 + i] / 1000, h[128 + t * 4 + i] % 1000); printf("\n"); }
    printf("16x256b.x2 at lane 32, col 0:\n");
    for (int t = 0; t < 32; t += 3) { printf("t%2d:", t); for (int i = 0; i < 8; ++i) printf(" (%u,%u)", h[256 + t * 8 + i] / 1000, h[256 + t * 8 + i] % 1000); printf("\n"); }
    printf("16x128b.x1 at lane 32, col 0:\n");
    for (int t = 0; t < 32; t += 1) { printf("t%2d:", t); for (int i = 0; i < 2; ++i) printf(" (%u,%u)", h[512 + t * 2 + i] / 1000, h[512 + t * 2 + i] % 1000); printf("\n"); }
    return 0;
}
