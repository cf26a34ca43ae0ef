// Probe: global-store rate of per-row bulk (TMA 1-D) stores from shared memory versus plain STG, in the access pattern of
// a GEMM epilogue: each warp owns a 32 x 32 fp32 tile whose rows go to global rows `ldy` floats apart.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store_probe scripts/probes/bulk_store_rows.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int WARPS = 8, ROW_B = 144;   // padded staging row (128 B of data)

__global__ void __launch_bounds__(WARPS * 32) probe(float* Y, int ldy, int tiles_per_cta, int mode) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* st = sm + warp * 32 * ROW_B;
    for (int t = 0; t < tiles_per_cta; ++t) {
        const int64_t tile = (int64_t)blockIdx.x * tiles_per_cta + t;
        // this warp's rows: tile*128 + (warp&3)*32 + lane; its columns: (warp>>2)*32 .. +32, and 4 column blocks per tile
        for (int cb = 0; cb < 4; ++cb) {
            const int64_t m = tile * 128 + (warp & 3) * 32 + lane;
            float* yrow = Y + m * ldy + cb * 64 + (warp >> 2) * 32;
            if (mode == 0) {               // plain stores, thread = row (half-used sectors per instruction)
#pragma unroll
                for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(yrow)[j] = make_float4(1.f, 2.f, 3.f, (float)j);
            } else {
                if (mode == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging free again
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(st + lane * ROW_B)[j] = make_float4(1.f, 2.f, 3.f, (float)j);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (mode == 2) {           // one 128-byte bulk store per lane (= per row)
                    const uint32_t s = (uint32_t)__cvta_generic_to_shared(st + lane * ROW_B);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(yrow), "r"(s) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                } else {                   // mode 1: staged, coalesced STG (4 rows x 128 B per instruction)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + (lane >> 3), lc = lane & 7;
                        const float4 v = *reinterpret_cast<const float4*>(st + row * ROW_B + lc * 16);
                        *reinterpret_cast<float4*>(Y + (tile * 128 + (warp & 3) * 32 + row) * (int64_t)ldy + cb * 64 + (warp >> 2) * 32 + lc * 4) = v;
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (mode == 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const int ldy = 256, tiles_per_cta = 124, grid = 148;
    const size_t rows = (size_t)grid * tiles_per_cta * 128;
    float* Y;
    cudaMalloc(&Y, rows * ldy * sizeof(float));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            probe<<<grid, WARPS * 32, 200 * 1024>>>(Y, ldy, tiles_per_cta, mode);   // big smem request: 1 CTA per SM
            cudaEventRecord(b);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d error: %s\n", mode, cudaGetErrorString(e)); return 1; }
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double bytes = (double)rows * ldy * 4;
            if (rep == 2) printf("mode %d (%s): %.3f ms, %.2f TB/s, %.1f B/clk/SM at 1.9 GHz\n", mode,
                                 mode == 0 ? "STG row-per-thread" : mode == 1 ? "staged coalesced STG" : "per-row bulk stores",
                                 ms, bytes / ms / 1e9, bytes / (ms * 1e-3) / 148 / 1.9e9);
        }
    }
    return 0;
}
