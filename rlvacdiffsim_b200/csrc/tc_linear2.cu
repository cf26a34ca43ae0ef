// K3b for LARGE node batches — the same split-fp16 tcgen05 GEMM as tc_linear.cu, laid out for TWO CTAs PER SM.
//
// tc_linear.cu keeps one CTA per SM busy with an intra-CTA pipeline (fp32 staging, pre-conversion of the next tile, double
// accumulators); its in-kernel timeline (profiles/timeline_tclinear_r02.txt) shows what is left: inside a CTA the phases of a
// tile — operand load / convert, first GEMM, hidden epilogue, second GEMM, three store epilogues — still run one after another
// (23 k cycles per 128-row tile of the context net against 11 k of HBM time), because one warp group does all of them.
// Here a CTA is deliberately sequential and SMALL (99 KB of shared memory, 256 TMEM columns, <= 96 registers at 320 threads),
// so that two of them share an SM and the store-heavy epilogues of one run under the loads, conversions and MMAs of the other.
//
//   warps 0-7  workers: global fp32 A tile -> registers -> fp16 hi/lo -> K-major core-matrix operand buffers (no staging
//              buffer: lane (row%8, k-chunk) reads 32 contiguous bytes and writes one conflict-free 16-byte core row);
//              later drain TMEM, bias / SiLU, and either write the hidden tile back into the operand buffers (two-layer
//              mode) or store Y straight from registers in the m16n8 fragment shape (whole 32-byte sectors)
//   warp  8    producer: bulk-TMA copies of 16 KB weight stages (128 out-rows x 32 k, hi | lo) through a 2-deep ring, read
//              from the SAME pre-formatted 32 KB stages tc_linear.cu uses (a K = 32 stage is half of each half)
//   warp  9    TMEM owner (256 columns = D0 | D1 of one 128 x 128 accumulator); one lane issues tcgen05.mma + commits
// Arithmetic, operand formats and epilogue expressions are those of tc_linear.cu: D0 += A_hi B_hi, D1 += A_hi B_lo + A_lo B_hi,
// Y = D0 + 2^-11 D1 — the two kernels produce bit-identical outputs (tests/test_gpu_parity.py::test_tc_linear2_equals_tc_linear).
// Shapes: K = 128 with N in {128, 256, 384}, or K = 256 (two phases) with N = 128; optional fused second layer (N = 128 ->
// N2 <= 384); optional mix-reduce epilogue (see tc_linear.cu).  Anything else, and small batches, stay on tc_linear.cu.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rlvd {

namespace {

constexpr int T2_M = 128;
constexpr int T2_N = 128;
constexpr int T2_KS = 32;                              // k elements per weight stage
constexpr int T2_NSTAGE = 2;
constexpr int T2_HALF = T2_N * T2_KS * 2;              // 8192 B: one of hi / lo
constexpr int T2_STAGE = 2 * T2_HALF;                  // 16384 B
constexpr int T1_STAGE = 32768;                        // tc_linear.cu's stage: [hi 128 x 64 | lo 128 x 64]
constexpr int T1_HALF = 16384;
constexpr int CORE_LBO = 2048;                         // bytes between K-adjacent core matrices
constexpr int CORE_SBO = 128;                          // bytes between 8-row groups
constexpr int T2_THREADS = 320;
constexpr int T2_SMEM_A = 65536;                       // A_hi | A_lo
constexpr int T2_SMEM = T2_SMEM_A + T2_NSTAGE * T2_STAGE + 128 + 2 * 384 * 4;
constexpr uint32_t IDESC = (1u << 4) | (16u << 17) | (8u << 24);   // fp32 accum, fp16 A/B, K-major, M = 128, N = 128

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(CORE_LBO >> 4) << 16) | ((uint64_t)(CORE_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void split8(const float4 lo4, const float4 hi4, uint4& H, uint4& L, bool& bad) {
    const float v[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
    __half h[8], l[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        bad |= !(fabsf(v[t]) < 65504.0f);
        h[t] = __float2half_rn(v[t]);
        l[t] = __float2half_rn((v[t] - __half2float(h[t])) * 2048.0f);
    }
    H = *reinterpret_cast<uint4*>(h);
    L = *reinterpret_cast<uint4*>(l);
}

struct T2Smem {
    uint64_t full[T2_NSTAGE], empty[T2_NSTAGE], acc_full, acc_empty, a_ready, a_free;
    uint32_t tmem_base;
};

__device__ __forceinline__ int64_t tile_row(const LinearArgs& a, int tile, int row, bool& live) {
    if (a.mixred) {
        const int l = row & 31;
        const int64_t m = (int64_t)tile * 120 + (row >> 5) * 30 + l;
        live = l < 30 && m < a.M;
        return m;
    }
    const int64_t m = (int64_t)tile * T2_M + row;
    live = m < a.M;
    return m;
}

__global__ void __launch_bounds__(T2_THREADS, 2) tc_linear2_kernel(const LinearArgs a, const uint8_t* __restrict__ Wfmt,
                                                                   const uint8_t* __restrict__ W2fmt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (a.gate != nullptr && *a.gate != a.gate_value) return;   // uniform over the grid
    uint8_t* A_hi = smem;                              // 32 KB  [16 core columns][16 row groups][8 rows][16 B]
    uint8_t* A_lo = smem + 32768;                      // 32 KB
    uint8_t* Bst = smem + T2_SMEM_A;                   // T2_NSTAGE x 16 KB weight stages [hi 4 core columns | lo 4 core columns]
    T2Smem* S = reinterpret_cast<T2Smem*>(Bst + T2_NSTAGE * T2_STAGE);
    float* bias_s = reinterpret_cast<float*>(Bst + T2_NSTAGE * T2_STAGE + 128);   // [384]
    float* bias2_s = bias_s + 384;                                                 // [384]
    const bool two = W2fmt != nullptr;
    const int NB2 = two ? a.N2 / T2_N : 0;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = a.mixred ? (a.n_nodes + 39) / 40 : (a.M + T2_M - 1) / T2_M;
    const int NB = a.N / T2_N, NPH = a.K / 128;

    if (tid == 0) {
        for (int s = 0; s < T2_NSTAGE; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 1); }
        mbar_init(&S->acc_full, 1);
        mbar_init(&S->acc_empty, 256);
        mbar_init(&S->a_ready, 256);
        mbar_init(&S->a_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < a.N; i += T2_THREADS) bias_s[i] = a.bias != nullptr ? __ldg(a.bias + i) : 0.f;
    for (int i = tid; i < (two ? a.N2 : 0); i += T2_THREADS) bias2_s[i] = a.bias2 != nullptr ? __ldg(a.bias2 + i) : 0.f;
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S->tmem_base)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;

    if (warp < 8) {
        // ================= workers: operand conversion + epilogues (256 threads) =================
        const int q = warp & 3, h = warp >> 2;               // epilogue role: TMEM lane quarter q, column half h
        const int er = q * 32 + lane;                        // epilogue row = TMEM lane
        const int r8 = lane & 7, kc4 = lane >> 3;            // converter role: row in the 8-row group, k-chunk in the quad
        uint32_t acc_it = 0, a_it = 0;                       // accumulator uses / operand-buffer writes so far
        bool bad = false;                                    // an operand left the fp16-split range (engine status bit 0)
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int phase = 0; phase < NPH; ++phase, ++a_it) {
                const float* base = a.A;
                int ld = a.lda, k0 = phase * 128;
                if (a.A2 != nullptr && k0 >= a.K1) { base = a.A2; ld = a.lda2; k0 -= a.K1; }
                // 8 iterations per warp: row group 2 warp + it / 4, k-chunk quad it % 4 (chunk = 8 k = 32 bytes per lane)
                float4 v[8][2];
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rg = warp * 2 + (it >> 2), kc = (it & 3) * 4 + kc4;
                    bool live;
                    const int64_t m = tile_row(a, tile, rg * 8 + r8, live);
                    const float4* src = reinterpret_cast<const float4*>(base + (live ? m * ld : 0) + k0 + kc * 8);
                    v[it][0] = live ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[it][1] = live ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(&S->a_free, (a_it & 1) ^ 1);       // MMAs that read the previous operand contents have retired
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rg = warp * 2 + (it >> 2), kc = (it & 3) * 4 + kc4;
                    uint4 H, L;
                    split8(v[it][0], v[it][1], H, L, bad);
                    const uint32_t off = kc * CORE_LBO + rg * CORE_SBO + r8 * 16;
                    *reinterpret_cast<uint4*>(A_hi + off) = H;
                    *reinterpret_cast<uint4*>(A_lo + off) = L;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> UMMA
                mbar_arrive(&S->a_ready);
            }
            if (two) {
                // ---- hidden layer: TMEM -> registers -> bias / SiLU -> fp16 hi/lo -> operand buffers
                mbar_wait(&S->acc_full, acc_it & 1);          // implies every first-layer MMA has finished reading A
                tc_fence_after();
                const uint32_t hrow = (er >> 3) * CORE_SBO + (er & 7) * 16;
#pragma unroll 1
                for (int sg = 0; sg < 4; ++sg) {
                    const int c0 = 64 * (sg >> 1) + 32 * h + 16 * (sg & 1);
                    const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + c0;
                    uint32_t d0[16], d1[16];
                    tmem_ld16(t0, d0);
                    tmem_ld16(t0 + 128, d1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (sg == 3) {
                        tc_fence_before();
                        mbar_arrive(&S->acc_empty);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 8) {
                        float y[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float w = fmaf(__uint_as_float(d1[j + u]), 1.0f / 2048.0f, __uint_as_float(d0[j + u])) + bias_s[c0 + j + u];
                            if (a.act == 1) w = __fdividef(w, 1.0f + __expf(-w));
                            y[u] = w;
                        }
                        uint4 H, L;
                        split8(make_float4(y[0], y[1], y[2], y[3]), make_float4(y[4], y[5], y[6], y[7]), H, L, bad);
                        const int kc = (c0 + j) >> 3;
                        *reinterpret_cast<uint4*>(A_hi + kc * CORE_LBO + hrow) = H;
                        *reinterpret_cast<uint4*>(A_lo + kc * CORE_LBO + hrow) = L;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&S->a_ready);
                ++acc_it;
                ++a_it;
            }
            const int NBo = two ? NB2 : NB;
            const float* bias_o = two ? bias2_s : bias_s;
            const int act_o = two ? a.act2 : a.act;
            for (int nb = 0; nb < NBo; ++nb, ++acc_it) {
                mbar_wait(&S->acc_full, acc_it & 1);
                tc_fence_after();
                if (a.mixred) {
                    // ---- this N-block holds V (columns 0-63) and W (64-127) of channels [64 nb, 64 nb + 64): nrm, vw per node
                    bool live_row;
                    const int64_t m_row = tile_row(a, tile, er, live_row);
                    const bool owner = live_row && (lane % 3) == 0;      // lane of component 0 writes the node's nrm / vw
                    const int64_t node = m_row / 3;
#pragma unroll 1
                    for (int hf = 0; hf < 2; ++hf) {
                        const int c0 = 32 * h + 16 * hf;
                        const uint32_t tV = tmem + ((uint32_t)(q * 32) << 16) + c0;
                        uint32_t v0[16], v1[16], w0[16], w1[16];
                        tmem_ld16(tV, v0);
                        tmem_ld16(tV + 128, v1);
                        tmem_ld16(tV + 64, w0);
                        tmem_ld16(tV + 64 + 128, w1);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        float nr[16], dt[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const float vv = fmaf(__uint_as_float(v1[u]), 1.0f / 2048.0f, __uint_as_float(v0[u]));
                            const float ww = fmaf(__uint_as_float(w1[u]), 1.0f / 2048.0f, __uint_as_float(w0[u]));
                            const float vy = __shfl_down_sync(0xffffffffu, vv, 1), vz = __shfl_down_sync(0xffffffffu, vv, 2);
                            const float wy = __shfl_down_sync(0xffffffffu, ww, 1), wz = __shfl_down_sync(0xffffffffu, ww, 2);
                            nr[u] = sqrtf(fmaf(vz, vz, fmaf(vy, vy, vv * vv)) + a.eps);
                            dt[u] = fmaf(vz, wz, fmaf(vy, wy, vv * ww));
                        }
                        if (owner) {
                            float4* pn = reinterpret_cast<float4*>(a.nrm + node * T2_N + 64 * nb + c0);
                            float4* pv = reinterpret_cast<float4*>(a.vw + node * T2_N + 64 * nb + c0);
#pragma unroll
                            for (int u = 0; u < 16; u += 4) {
                                pn[u >> 2] = make_float4(nr[u], nr[u + 1], nr[u + 2], nr[u + 3]);
                                pv[u >> 2] = make_float4(dt[u], dt[u + 1], dt[u + 2], dt[u + 3]);
                            }
                        }
                    }
                    if (a.Y != nullptr) {
                        const int u = lane & 3, rq = lane >> 2;
#pragma unroll 1
                        for (int hl = 0; hl < 2; ++hl) {
                            const uint32_t t1 = tmem + ((uint32_t)(q * 32 + 16 * hl) << 16) + 64 + 32 * h;
                            uint32_t d0[16], d1[16];
                            tmem_ld_16x256b_x4(t1, d0);
                            tmem_ld_16x256b_x4(t1 + 128, d1);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int rh = 0; rh < 2; ++rh) {
                                bool live;
                                const int64_t m = tile_row(a, tile, q * 32 + 16 * hl + 8 * rh + rq, live);
                                float* yrow = a.Y + m * a.ldy + 64 * nb + 32 * h + 2 * u;
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    const float w0v = fmaf(__uint_as_float(d1[4 * g + 2 * rh]), 1.0f / 2048.0f, __uint_as_float(d0[4 * g + 2 * rh]));
                                    const float w1v = fmaf(__uint_as_float(d1[4 * g + 2 * rh + 1]), 1.0f / 2048.0f, __uint_as_float(d0[4 * g + 2 * rh + 1]));
                                    if (live) *reinterpret_cast<float2*>(yrow + 8 * g) = make_float2(w0v, w1v);
                                }
                            }
                        }
                    }
                    tc_fence_before();                       // all TMEM reads of this accumulator are done
                    mbar_arrive(&S->acc_empty);
                    continue;
                }
                // ---- plain epilogue: four quarter-steps (column segment, 16-lane half), m16n8 fragment shape, direct stores
                const int u = lane & 3, rq = lane >> 2;
#pragma unroll 1
                for (int st = 0; st < 4; ++st) {
                    const int sgm = st >> 1, hl = st & 1;
                    const int c0 = 64 * sgm + 32 * h;        // first column (inside the N-block) of this warp's segment
                    const int col0 = nb * T2_N + c0;
                    const uint32_t t0 = tmem + ((uint32_t)(q * 32 + 16 * hl) << 16) + c0;
                    uint32_t d0[16], d1[16];
                    tmem_ld_16x256b_x4(t0, d0);
                    tmem_ld_16x256b_x4(t0 + 128, d1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (st == 3) {                           // all TMEM reads of this accumulator are done
                        tc_fence_before();
                        mbar_arrive(&S->acc_empty);
                    }
                    float2 bb[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) bb[g] = *reinterpret_cast<const float2*>(bias_o + col0 + 8 * g + 2 * u);
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const int64_t m = (int64_t)tile * T2_M + q * 32 + 16 * hl + 8 * rh + rq;
                        float* yrow = a.Y + m * a.ldy + col0 + 2 * u;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float v0 = fmaf(__uint_as_float(d1[4 * g + 2 * rh]), 1.0f / 2048.0f, __uint_as_float(d0[4 * g + 2 * rh])) + bb[g].x;
                            float v1 = fmaf(__uint_as_float(d1[4 * g + 2 * rh + 1]), 1.0f / 2048.0f, __uint_as_float(d0[4 * g + 2 * rh + 1])) + bb[g].y;
                            if (act_o == 1) {
                                v0 = __fdividef(v0, 1.0f + __expf(-v0));
                                v1 = __fdividef(v1, 1.0f + __expf(-v1));
                            }
                            if (m < a.M) *reinterpret_cast<float2*>(yrow + 8 * g) = make_float2(v0, v1);
                        }
                    }
                }
            }
        }
        if (bad && a.status != nullptr) atomicOr(a.status, 1);
    } else if (warp == 8) {
        // ================= weight-stage producer =================
        if (lane == 0) {
            uint32_t cnt = 0;
            const int KSN = a.K / 64;                        // tc_linear.cu's 64-k stages per N-block
            auto stage = [&](const uint8_t* W, int nb, int ksn, int k32) {   // k32: K = 32 sub-stage index inside the layer's K
                const uint32_t s = cnt % T2_NSTAGE, ph = (cnt / T2_NSTAGE) & 1;
                mbar_wait(&S->empty[s], ph ^ 1);
                mbar_expect_tx(&S->full[s], T2_STAGE);
                const uint8_t* src = W + (size_t)(nb * ksn + (k32 >> 1)) * T1_STAGE + (k32 & 1) * T2_HALF;
                bulk_g2s(Bst + s * T2_STAGE, src, T2_HALF, &S->full[s]);
                bulk_g2s(Bst + s * T2_STAGE + T2_HALF, src + T1_HALF, T2_HALF, &S->full[s]);
                ++cnt;
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int phase = 0; phase < NPH; ++phase)
                    for (int nb = 0; nb < NB; ++nb)
                        for (int k = 0; k < 4; ++k) stage(Wfmt, nb, KSN, phase * 4 + k);
                for (int nb = 0; nb < NB2; ++nb)
                    for (int k = 0; k < 4; ++k) stage(W2fmt, nb, 2, k);
            }
        }
    } else {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t cnt = 0, acc_it = 0, a_it = 0;
            const uint32_t a_hi0 = smem_u32(A_hi), a_lo0 = smem_u32(A_lo), b0 = smem_u32(Bst);
            const uint32_t d0 = tmem, d1 = tmem + 128;
            auto kstages = [&](bool first) {                 // the four K = 32 stages of one 128-k phase of one N-block
                for (int k = 0; k < 4; ++k, ++cnt) {
                    const uint32_t s = cnt % T2_NSTAGE, ph = (cnt / T2_NSTAGE) & 1;
                    mbar_wait(&S->full[s], ph);
                    tc_fence_after();
                    const uint32_t bh = b0 + s * T2_STAGE, bl = bh + T2_HALF;
#pragma unroll
                    for (int j = 0; j < T2_KS / 16; ++j) {
                        const uint32_t koff_a = (uint32_t)(k * (T2_KS / 8) + 2 * j) * CORE_LBO;
                        const uint32_t koff_b = (uint32_t)(2 * j) * CORE_LBO;
                        const uint32_t acc = (first && k == 0 && j == 0) ? 0u : 1u;
                        umma_f16(d0, make_desc(a_hi0 + koff_a), make_desc(bh + koff_b), acc);
                        umma_f16(d1, make_desc(a_hi0 + koff_a), make_desc(bl + koff_b), acc);
                        umma_f16(d1, make_desc(a_lo0 + koff_a), make_desc(bh + koff_b), 1u);
                    }
                    umma_commit(&S->empty[s]);               // frees the weight stage
                }
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int phase = 0; phase < NPH; ++phase, ++a_it) {
                    mbar_wait(&S->a_ready, a_it & 1);
                    tc_fence_after();
                    for (int nb = 0; nb < NB; ++nb) {        // NPH == 2 implies NB == 1 (one accumulator)
                        if (phase == 0) {
                            mbar_wait(&S->acc_empty, (acc_it & 1) ^ 1);
                            tc_fence_after();
                        }
                        kstages(phase == 0);
                        if (phase == NPH - 1) { umma_commit(&S->acc_full); ++acc_it; }
                    }
                    umma_commit(&S->a_free);                 // the operand buffers may be overwritten
                }
                if (two) {
                    mbar_wait(&S->a_ready, a_it & 1);
                    tc_fence_after();
                    for (int nb = 0; nb < NB2; ++nb, ++acc_it) {
                        mbar_wait(&S->acc_empty, (acc_it & 1) ^ 1);
                        tc_fence_after();
                        kstages(true);
                        umma_commit(&S->acc_full);
                    }
                    umma_commit(&S->a_free);
                    ++a_it;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

}  // namespace

bool tc_linear2_supported(const LinearArgs& a) {
    if (!tc_linear_supported(a)) return false;
    if (a.K == 256) return a.N == T2_N;                      // two phases share the single accumulator
    return a.K == 128;
}

int launch_tc_linear2(rlvd_engine* e, cudaStream_t s, const LinearArgs& a_in, const void* Wfmt) {
    LinearArgs a = a_in;
    a.status = e->status.as<int>();
    RLVD_CHECK(tc_linear2_supported(a), "tc_linear2: unsupported shape M=%d K=%d N=%d", a.M, a.K, a.N);
    if (a.M <= 0) return 0;
    if (!e->attr_tc_linear2) {
        RLVD_CUDA(cudaFuncSetAttribute(tc_linear2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
        e->attr_tc_linear2 = true;
    }
    const int n_tiles = a.mixred ? (a.n_nodes + 39) / 40 : (a.M + T2_M - 1) / T2_M;
    const int grid = n_tiles < 2 * e->sm_count ? n_tiles : 2 * e->sm_count;
    Span sp(e, s, FAM_NODE);
    tc_linear2_kernel<<<grid, T2_THREADS, T2_SMEM, s>>>(a, reinterpret_cast<const uint8_t*>(Wfmt),
                                                        reinterpret_cast<const uint8_t*>(a.W2fmt));
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
