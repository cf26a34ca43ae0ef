// Probe: how fast can ONE SM pull a 255-node x 32-channel slice of the node rows (x: [M, 384] fp32, mu: [M, 3, 128] fp32 ->
// 6 segments of 128 B per node, 196 KB per CTA) into shared memory, in the configuration of edge_stream_kernel (226 KB of
// dynamic shared memory, 1 CTA per SM, 576 staging threads)?  Only a few CTAs run, so HBM is far from saturated — as in the
// real kernel, where ~9 % of the SMs are in their prologue at any time.
//   mode 0: one item (6 x LDG.128.nc) per thread in flight        (round-2 first form)
//   mode 1: two items in flight, ld.global.cg
//   mode 2: four items in flight, ld.global.cg (needs ~100 registers)
//   mode 3: 5 bulk (TMA 1-D) copies of 128 B per node + x_m through registers
//   mode 4: like 1 with 8-byte loads (twice the requests, half the bytes each)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stage_rate scripts/probes/stage_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int F = 128, SL = 32, NB = 640, N = 255, THREADS = 768, STAGE = 576;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) probe(const float* __restrict__ x, const float* __restrict__ mu, int n_struct, int reps,
                                                    long long* out, float* sink, int prefetch) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long total = 0;
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
        const int b = (blockIdx.x * reps + r) * 977 % (n_struct * 4);
        const int g = b >> 2, slice = b & 3;
        const size_t node0 = (size_t)g * N;
        if (prefetch) {          // ask L2 for the structure's rows (16 chunks, like the kernel), then idle ~30 k cycles
            if (tid < 16 && slice == (tid & 3)) {}
            if (tid < 16) {
                const uint32_t total = (uint32_t)N * 3u * F * 4u;
                const uint32_t per = ((total + 7u) / 8u + 127u) & ~127u;
                const uint32_t off = (uint32_t)(tid & 7) * per;
                const uint8_t* base = reinterpret_cast<const uint8_t*>(tid < 8 ? x : mu) + node0 * 3 * F * 4;
                const uint32_t bytes = total - off < per ? total - off : per;
                if (off < total) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(bytes) : "memory");
            }
            const long long w0 = clock64();
            while (clock64() - w0 < 30000) {}
        }
        __syncthreads();
        const long long t0 = clock64();
        if (tid < STAGE) {
            if (MODE == 0 || MODE == 1 || MODE == 2) {
                constexpr int D = MODE == 0 ? 1 : MODE == 1 ? 2 : 4;
                for (int id0 = tid; id0 < N * 8; id0 += D * STAGE) {
                    float4 v[D][6];
#pragma unroll
                    for (int h = 0; h < D; ++h) {
                        const int id = id0 + h * STAGE;
                        if (id < N * 8) {
                            const int node = id >> 3, c4 = id & 7;
                            const float4* xr = reinterpret_cast<const float4*>(x + (node0 + node) * 3 * F + slice * SL) + c4;
                            const float4* mr = reinterpret_cast<const float4*>(mu + (node0 + node) * 3 * F + slice * SL) + c4;
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                v[h][k] = MODE == 0 ? __ldg(xr + k * (F / 4)) : __ldcg(xr + k * (F / 4));
                                v[h][3 + k] = MODE == 0 ? __ldg(mr + k * (F / 4)) : __ldcg(mr + k * (F / 4));
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < D; ++h) {
                        const int id = id0 + h * STAGE;
                        if (id < N * 8) {
                            const int node = id >> 3, c4 = id & 7;
                            uint8_t* row = smem + node * NB + c4 * 16;
                            *reinterpret_cast<float4*>(row) = v[h][0];
                            *reinterpret_cast<float4*>(row + 128) = v[h][1];
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                *reinterpret_cast<float4*>(row + (2 + k) * 128) =
                                    make_float4(v[h][2].x * v[h][3 + k].x, v[h][2].y * v[h][3 + k].y, v[h][2].z * v[h][3 + k].z, v[h][2].w * v[h][3 + k].w);
                        }
                    }
                }
            } else if (MODE == 4) {
                for (int id0 = tid; id0 < N * 16; id0 += 2 * STAGE) {
                    float2 v[2][6];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int id = id0 + h * STAGE;
                        if (id < N * 16) {
                            const int node = id >> 4, c2 = id & 15;
                            const float2* xr = reinterpret_cast<const float2*>(x + (node0 + node) * 3 * F + slice * SL) + c2;
                            const float2* mr = reinterpret_cast<const float2*>(mu + (node0 + node) * 3 * F + slice * SL) + c2;
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                v[h][k] = __ldcg(xr + k * (F / 2));
                                v[h][3 + k] = __ldcg(mr + k * (F / 2));
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int id = id0 + h * STAGE;
                        if (id < N * 16) {
                            const int node = id >> 4, c2 = id & 15;
                            uint8_t* row = smem + node * NB + c2 * 8;
                            *reinterpret_cast<float2*>(row) = v[h][0];
                            *reinterpret_cast<float2*>(row + 128) = v[h][1];
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                *reinterpret_cast<float2*>(row + (2 + k) * 128) = make_float2(v[h][2].x * v[h][3 + k].x, v[h][2].y * v[h][3 + k].y);
                        }
                    }
                }
            } else {   // MODE 3
                if (tid == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(N * 5 * 128) : "memory");
                }
                // x_m through registers (4 items of one float4)
                float4 xm[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int id = tid + h * STAGE;
                    if (id < N * 8)
                        xm[h] = __ldcg(reinterpret_cast<const float4*>(x + (node0 + (id >> 3)) * 3 * F + 2 * F + slice * SL) + (id & 7));
                }
                __syncwarp();
                // bulk copies: 5 per node
                for (int c = tid; c < N * 5; c += STAGE) {
                    const int node = c / 5, seg = c - node * 5;
                    const float* src = seg < 2 ? x + (node0 + node) * 3 * F + seg * F + slice * SL
                                               : mu + (node0 + node) * 3 * F + (seg - 2) * F + slice * SL;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     s32(smem + node * NB + seg * 128)),
                                 "l"(src), "r"(128), "r"(s32(&bar))
                                 : "memory");
                }
                // wait
                uint32_t ok = 0;
                while (!ok) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                        : "=r"(ok)
                        : "r"(s32(&bar)), "r"((uint32_t)(r & 1)), "r"(20000u)
                        : "memory");
                }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int id = tid + h * STAGE;
                    if (id < N * 8) {
                        uint8_t* row = smem + (id >> 3) * NB + (id & 7) * 16;
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            float4 m = *reinterpret_cast<float4*>(row + (2 + k) * 128);
                            m.x *= xm[h].x; m.y *= xm[h].y; m.z *= xm[h].z; m.w *= xm[h].w;
                            *reinterpret_cast<float4*>(row + (2 + k) * 128) = m;
                        }
                    }
                }
            }
        }
        __syncthreads();
        total += clock64() - t0;
        acc += reinterpret_cast<float*>(smem)[(tid * 37 + r) % (N * 160)];
    }
    if (tid == 0) out[blockIdx.x] = total / reps;
    if (acc == 123.456f) sink[0] = acc;
}

template <int MODE>
void run(const float* x, const float* mu, int n_struct, int grid, long long* d_out, float* sink, const char* what, int prefetch = 0) {
    const int smem = 226 * 1024 - 64;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 24;
    for (int it = 0; it < 2; ++it) probe<MODE><<<grid, THREADS, smem>>>(x, mu, n_struct, reps, d_out, sink, prefetch);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(err)); return; }
    long long h[148];
    cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < grid; ++i) s += (double)h[i];
    printf("grid %3d  mode %d  %-58s %8.0f cycles per table (196 KB) = %5.1f B/clk\n", grid, MODE, what, s / grid, 195840.0 / (s / grid));
}

int main() {
    const int n_struct = 3072;
    const size_t elems = (size_t)n_struct * N * 3 * F;
    float *x, *mu, *sink;
    long long* d_out;
    cudaMalloc(&x, elems * 4);
    cudaMalloc(&mu, elems * 4);
    cudaMalloc(&sink, 4);
    cudaMalloc(&d_out, 148 * 8);
    cudaMemset(x, 0, elems * 4);
    cudaMemset(mu, 0, elems * 4);
    for (int grid : {8, 148}) {
        run<0>(x, mu, n_struct, grid, d_out, sink, "1 item in flight, ld.global.nc");
        run<1>(x, mu, n_struct, grid, d_out, sink, "2 items in flight, ld.global.cg");
        run<2>(x, mu, n_struct, grid, d_out, sink, "4 items in flight, ld.global.cg");
        run<4>(x, mu, n_struct, grid, d_out, sink, "2 items in flight, 8-byte loads");
        run<3>(x, mu, n_struct, grid, d_out, sink, "bulk copies (5 x 128 B per node) + x_m via registers");
    }
    run<1>(x, mu, 48, 148, d_out, sink, "2 in flight, 48 structures (L2-resident)");
    run<1>(x, mu, n_struct, 148, d_out, sink, "2 in flight, rows prefetched to L2 30 k cycles before", 1);
    run<1>(x, mu, n_struct, 8, d_out, sink, "2 in flight, rows prefetched to L2 30 k cycles before", 1);
    return 0;
}
