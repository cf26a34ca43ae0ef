// K3b on the 5th-generation tensor cores — Y = act(A·Wᵀ + b) with tcgen05.mma, accumulators in TMEM.
//
// The channel-mixing linears of PaiNN (128→128, 128→384, 128→256 on the three vector
// components, 256→128, 128→384) are true dense GEMMs with M = number of atoms in the batch.  The parity
// bar is fp32 (1e-5 relative on Q), which a single fp16/bf16/tf32 pass cannot meet, so every operand is
// split into two fp16 numbers
//        a = a_hi + 2^-11 · a_lo        (a_hi = fp16(a), a_lo = fp16((a - a_hi)·2^11): 22 significant bits)
// and the product is evaluated as three tensor-core passes into two fp32 TMEM accumulators
//        D0 += A_hi·B_hiᵀ          D1 += A_hi·B_loᵀ + A_lo·B_hiᵀ          Y = D0 + 2^-11·D1
// (the dropped A_lo·B_lo term is 2^-22 relative).  Scaling the low parts by 2^11 keeps them in fp16's normal
// range.  Three kind::f16 passes cost 1.5 tf32-pass equivalents and 4 bytes of shared memory per element.
//
// Kernel shape (persistent, one CTA per SM, 576 threads, 224 KB of shared memory; EW = 16 stager / epilogue warps — four per
// TMEM lane quarter, each draining 32 of an accumulator's 128 columns; -DRLVD_TC_EPI_WARPS=8 is the round-1 shape, warps 8 / 9
// below are then the producer and the issuer):
//   warps 0-15 stage the next 128-row x 128-k fp32 A phase with fully coalesced cp.async (prefetched one phase
//              ahead, overlapping the MMAs and the epilogue), split it into fp16 hi/lo and write it in the
//              canonical no-swizzle K-major core-matrix layout; later drain TMEM (tcgen05.ld, thread = row),
//              apply bias/SiLU, transpose through a swizzled staging tile and store Y with coalesced 16-byte
//              stores
//   warp  16   producer: 1-D bulk-TMA copies (cp.async.bulk → mbarrier complete_tx) of pre-formatted 32 KB
//              weight stages [128 out-rows x 64 k, hi | lo] through a 2-deep ring
//   warp  17   allocates TMEM (512 columns = 2 x (D0,D1) x 128); one elected lane issues tcgen05.mma and the
//              tcgen05.commit arrivals that free weight stages / A buffers and publish accumulators
// K = 256 (the [q | n] concatenation of the intra-atomic net) runs as two 128-k phases through the same A
// buffers, accumulating in TMEM.  Every mbarrier wait is bounded (trap, never hang).
// Two-layer mode (LinearArgs::W2fmt): the interatomic / intraatomic context nets are Linear -> SiLU -> Linear with a
// 128-wide hidden layer.  The epilogue of the first GEMM writes SiLU(hidden) straight into the A operand buffers (fp16
// hi/lo, core-matrix layout) and the second GEMM runs from there: the hidden activations never touch HBM and are never
// staged or converted a second time.
// Mix-reduce mode (LinearArgs::mixred): mu_channel_mix maps every (node, component) row of mu to [V | W].  A 128-row
// tile is laid out as 4 warps x (10 nodes x 3 components) + 2 idle rows, so the three components of a node sit in
// adjacent TMEM lanes of ONE warp.  The weight rows are permuted on the host (tc_format_mix_weight) so that N-block nb
// holds V and W of the SAME 64 channels (columns 0-63 | 64-127 of one accumulator): the epilogue of a block forms
// nrm = sqrt(sum_c V_c^2 + eps) and vw = sum_c V_c W_c with two warp shuffles per value and stores nrm, vw and W, while
// the MMAs of the other block (or of the next tile) run — the same double-buffered schedule as the plain epilogue.
// V never reaches HBM and the separate reduction kernel (and its 8 x [M,128] of traffic) is gone.  With Y == nullptr W is
// not stored either (the last mixing block of a call whose caller wants q only: mu's update has no consumer).
// Weights are formatted once on the host (rlvd_finalize_weights).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rlvd {

namespace {

constexpr int TC_M = 128;
constexpr int TC_N = 128;
constexpr int TC_KS = 64;                              // k elements per weight stage
constexpr int TC_NSTAGE = 2;
constexpr int TC_HALF_STAGE = TC_N * TC_KS * 2;        // 16384 B: one of hi / lo
constexpr int TC_STAGE_BYTES = 2 * TC_HALF_STAGE;      // 32768 B
constexpr int CORE_LBO = 2048;                         // bytes between K-adjacent core matrices
constexpr int CORE_SBO = 128;                          // bytes between 8-row groups
#ifndef RLVD_TC_EPI_WARPS
#define RLVD_TC_EPI_WARPS 16
#endif
constexpr int EW = RLVD_TC_EPI_WARPS;                 // stager / epilogue warps: 8 or 16 (2 or 4 per TMEM lane quarter)
constexpr int ET = EW * 32;                           // ... their threads
constexpr int HQ = EW / 4;                            // column partitions of an accumulator per lane quarter
constexpr int NSEG = 4 / HQ;                          // 32-column segments of a 128-column accumulator per warp
constexpr int MIXC = 64 / HQ;                         // V (and W) columns of a mix-reduce block per warp
constexpr int CPT = 16 / (ET / 128);                  // core columns (8 k each) one converter thread writes per phase
static_assert(EW == 8 || EW == 16, "tc_linear: 8 or 16 epilogue warps");
constexpr int TC_THREADS = ET + 64;                   // + weight producer warp + MMA issuer warp
constexpr uint32_t IDESC_F16_M128_N128 = (1u << 4) | (16u << 17) | (8u << 24);   // fp32 accum, fp16 A/B, K-major

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(CORE_LBO >> 4) << 16) | ((uint64_t)(CORE_SBO >> 4) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(IDESC_F16_M128_N128), "r"(accumulate)
        : "memory");
}
// `bad` collects values outside the fp16 range (|v| >= 65504, inf, NaN): the fp16 hi/lo split cannot represent them.
// Physical inputs stay below ~1e2; unphysically dense structures make the network itself diverge to 1e20+ in fp32.
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

struct TcSmem {
    uint64_t full[TC_NSTAGE], empty[TC_NSTAGE], acc_full[2], acc_empty[2], a_ready, a_free;
    uint32_t tmem_base;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory"); }

// Global row of tile row `row`: plain tiles are 128 consecutive rows; mix-reduce tiles hold 40 nodes as
// 4 warps x 30 rows (+2 idle rows per warp), global row = 120*tile + 30*warp + lane.
__device__ __forceinline__ int64_t tile_row(const LinearArgs& a, int tile, int row, bool& live) {
    if (a.mixred) {
        const int l = row & 31;
        const int64_t m = (int64_t)tile * 120 + (row >> 5) * 30 + l;
        live = l < 30 && m < a.M;
        return m;
    }
    const int64_t m = (int64_t)tile * TC_M + row;
    live = m < a.M;
    return m;
}

// Asynchronous, fully coalesced load of one 128-row x 128-k fp32 A phase into the swizzled staging buffer:
// 16-byte chunk `lc` of row `row` lands at row*512 + ((lc ^ (row & 7)) * 16).
__device__ __forceinline__ void stage_A_async(const LinearArgs& a, int tile, int phase, uint32_t stage_addr, int t) {
    const float* base = a.A;
    int ld = a.lda, k0 = phase * 128;
    if (a.A2 != nullptr && k0 >= a.K1) { base = a.A2; ld = a.lda2; k0 -= a.K1; }
#pragma unroll
    for (int i = 0; i < 4096 / ET; ++i) {
        const int id = i * ET + t, row = id >> 5, lc = id & 31;
        bool live;
        const int64_t m = tile_row(a, tile, row, live);
        const float* src = base + (live ? m * ld : 0) + k0 + lc * 4;
        cp_async16(stage_addr + row * 512 + ((lc ^ (row & 7)) << 4), src, live ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

#ifdef RLVD_TC_TIMELINE
#define TCS(role, tag) if (a.dbg != nullptr && blockIdx.x == 0 && tl_n < 500) { a.dbg[(role) * 1024 + 2 * tl_n] = (tag); a.dbg[(role) * 1024 + 2 * tl_n + 1] = clock64(); ++tl_n; }
#else
#define TCS(role, tag)
#endif

__global__ void __launch_bounds__(TC_THREADS, 1) tc_linear_kernel(const LinearArgs a, const uint8_t* __restrict__ Wfmt,
                                                                  const uint8_t* __restrict__ W2fmt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (a.gate != nullptr && *a.gate != a.gate_value) return;   // uniform over the grid
    uint8_t* A_hi = smem;                              // 32 KB  [16 core columns][16 row groups][8 rows][16 B]
    uint8_t* A_lo = smem + 32768;                      // 32 KB
    uint8_t* Ast = smem + 65536;                       // 64 KB  fp32 staging of the next A phase (swizzled rows)
    uint8_t* Bst = smem + 131072;                      // TC_NSTAGE x 32 KB weight stages
    uint8_t* Yst = Bst + TC_NSTAGE * TC_STAGE_BYTES;   // 32 KB  [128 rows][64 cols] fp32 output staging (swizzled)
    TcSmem* S = reinterpret_cast<TcSmem*>(Yst + 32768);
    float* bias_s = reinterpret_cast<float*>(Yst + 32768 + 128);      // [N <= 384] bias staged once (zeros when absent)
    const bool two = W2fmt != nullptr;
    float* bias2_s = bias_s + 128;                                     // [N2 <= 384] second-layer bias (two-layer mode: N == 128)
    const int NB2 = two ? a.N2 / TC_N : 0;
    // small M (fewer tiles than SMs): gridDim.y CTAs share a tile's second-layer N-blocks, each recomputing the hidden tile
    const int nb2_lo = two ? (int)blockIdx.y * NB2 / (int)gridDim.y : 0, nb2_hi = two ? ((int)blockIdx.y + 1) * NB2 / (int)gridDim.y : 0;
    const int NB2c = nb2_hi - nb2_lo;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int tl_n = 0; (void)tl_n;
    const int n_tiles = a.mixred ? (a.n_nodes + 39) / 40 : (a.M + TC_M - 1) / TC_M;
    const int NB = a.N / TC_N, NPH = a.K / 128;       // K-phases of 128 (the A buffers hold one phase)

    if (tid == 0) {
        for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&S->acc_full[b], 1); mbar_init(&S->acc_empty[b], ET); }
        mbar_init(&S->a_ready, ET);
        mbar_init(&S->a_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < a.N; i += TC_THREADS) bias_s[i] = a.bias != nullptr ? __ldg(a.bias + i) : 0.f;
    for (int i = tid; i < (two ? a.N2 : 0); i += TC_THREADS) bias2_s[i] = a.bias2 != nullptr ? __ldg(a.bias2 + i) : 0.f;
    if (warp == EW + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S->tmem_base)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;

    if (warp < EW) {
        // ================= A stager/converter + epilogue (ET threads) =================
        const int t = tid;
        const int r = t & 127, khalf = t >> 7;               // converter role: row r, core columns khalf*CPT..+CPT
        const int q = warp & 3, h = warp >> 2;               // epilogue role: TMEM lane quarter q, column half h
        const int er = q * 32 + lane;                        // epilogue row = TMEM lane

        const uint32_t ast = smem_u32(Ast);
        const uint32_t row_off = (r >> 3) * CORE_SBO + (r & 7) * 16;
        uint32_t acc_it = 0, ph_it = 0;
        bool bad = false;                                    // an operand left the fp16-split range (engine status bit 0)
        if (blockIdx.x < n_tiles) stage_A_async(a, blockIdx.x, 0, ast, t);
        // One A phase: fp32 staging (cp.async, issued one phase ahead) -> fp16 hi/lo operand buffers, then prefetch the next.
        auto convert_phase = [&](int tl, int phase) {
            if (tid == 0) { TCS(0, 1) }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (tid == 0) { TCS(0, 2) }
            mbar_wait(&S->a_free, (ph_it & 1) ^ 1);          // MMAs that read the previous A contents have retired
            epi_barrier();                                   // everyone's cp.async data is visible
            if (tid == 0) { TCS(0, 3) }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int kc = khalf * CPT + c;
                const float4 v0 = *reinterpret_cast<const float4*>(Ast + r * 512 + (((2 * kc) ^ (r & 7)) << 4));
                const float4 v1 = *reinterpret_cast<const float4*>(Ast + r * 512 + (((2 * kc + 1) ^ (r & 7)) << 4));
                uint4 H, L;
                split8(v0, v1, H, L, bad);
                *reinterpret_cast<uint4*>(A_hi + kc * CORE_LBO + row_off) = H;
                *reinterpret_cast<uint4*>(A_lo + kc * CORE_LBO + row_off) = L;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores → UMMA
            mbar_arrive(&S->a_ready);
            if (tid == 0) { TCS(0, 4) }
            epi_barrier();                                   // staging fully consumed → prefetch the next phase
            if (phase + 1 < NPH) stage_A_async(a, tl, phase + 1, ast, t);
            else if (tl + (int)gridDim.x < n_tiles) stage_A_async(a, tl + gridDim.x, 0, ast, t);
            if (tid == 0) { TCS(0, 5) }
            ++ph_it;
        };
        // Schedule: the first A phase of tile t+1 is converted BEFORE the last N-block epilogue of tile t, so the MMAs of
        // tile t+1 run underneath that epilogue instead of the epilogue warps idling on them afterwards (the MMA and
        // producer warps keep their program order; every barrier still sees its events in the same sequence).
        bool pre = false;                                    // phase 0 of the current tile is already in the operand buffers
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (tid == 0) { TCS(0, 0) }
            for (int phase = pre ? 1 : 0; phase < NPH; ++phase) convert_phase(tile, phase);
            pre = false;
            if (two) {
                // ---- hidden layer: TMEM -> registers -> bias/SiLU -> fp16 hi/lo -> A operand buffers (no HBM, no staging)
                const uint32_t buf = acc_it & 1;
                if (tid == 0) { TCS(0, 10) }
                mbar_wait(&S->acc_full[buf], (acc_it >> 1) & 1);
                mbar_wait(&S->a_free, (ph_it & 1) ^ 1);           // the first layer's MMAs have finished reading A
                tc_fence_after();
                if (tid == 0) { TCS(0, 11) }
#pragma unroll 1
                for (int sgm = 0; sgm < NSEG; ++sgm) {
                    const int c0 = (sgm * HQ + h) * 32;
                    const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + c0;
                    uint32_t d0[32], d1[32];
                    tmem_ld32(t0, d0);
                    tmem_ld32(t0 + 128, d1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (sgm == NSEG - 1) {
                        tc_fence_before();
                        mbar_arrive(&S->acc_empty[buf]);
                    }
                    const uint32_t hrow = (er >> 3) * CORE_SBO + (er & 7) * 16;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float y[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float v = fmaf(__uint_as_float(d1[j + u]), 1.0f / 2048.0f, __uint_as_float(d0[j + u])) + bias_s[c0 + j + u];
                            if (a.act == 1) v = __fdividef(v, 1.0f + __expf(-v));
                            y[u] = v;
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
                if (tid == 0) { TCS(0, 12) }
                ++acc_it;
                ++ph_it;
            }
            // ---- epilogue: TMEM → registers → bias/SiLU → swizzled staging → coalesced global stores
            const int nb_first = two ? nb2_lo : 0, nb_end = two ? nb2_hi : NB;
            const float* bias_o = two ? bias2_s : bias_s;
            const int act_o = two ? a.act2 : a.act;
            for (int nb = nb_first; nb < nb_end; ++nb, ++acc_it) {
                if (nb == nb_end - 1 && tile + (int)gridDim.x < n_tiles) {
                    convert_phase(tile + gridDim.x, 0);
                    pre = true;
                }
                const uint32_t buf = acc_it & 1;
                if (tid == 0) { TCS(0, 20 + nb) }
                mbar_wait(&S->acc_full[buf], (acc_it >> 1) & 1);
                tc_fence_after();
                if (tid == 0) { TCS(0, 30 + nb) }
                if (a.mixred) {
                    // ---- this N-block holds V (columns 0-63) and W (64-127) of channels [64 nb, 64 nb + 64): nrm, vw per node
                    bool live_row;
                    const int64_t m_row = tile_row(a, tile, er, live_row);
                    const bool owner = live_row && (lane % 3) == 0;      // lane of component 0 writes the node's nrm / vw
                    const int64_t node = m_row / 3;
#pragma unroll 1
                    for (int hf = 0; hf < MIXC / 16; ++hf) {
                        const int c0 = MIXC * h + 16 * hf;
                        const uint32_t tV = tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + c0;
                        uint32_t v0[16], v1[16], w0[16], w1[16];
                        tmem_ld16(tV, v0);
                        tmem_ld16(tV + 128, v1);
                        tmem_ld16(tV + 64, w0);
                        tmem_ld16(tV + 64 + 128, w1);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        float nr[16], dt[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const float v = fmaf(__uint_as_float(v1[u]), 1.0f / 2048.0f, __uint_as_float(v0[u]));
                            const float w = fmaf(__uint_as_float(w1[u]), 1.0f / 2048.0f, __uint_as_float(w0[u]));
#ifdef RLVD_MIXRED_NOSHFL   // ceiling experiment (wrong numbers): what the four shuffles per value cost
                            const float vy = v, vz = v, wy = w, wz = w;
#else
                            const float vy = __shfl_down_sync(0xffffffffu, v, 1), vz = __shfl_down_sync(0xffffffffu, v, 2);
                            const float wy = __shfl_down_sync(0xffffffffu, w, 1), wz = __shfl_down_sync(0xffffffffu, w, 2);
#endif
                            nr[u] = sqrtf(fmaf(vz, vz, fmaf(vy, vy, v * v)) + a.eps);
                            dt[u] = fmaf(vz, wz, fmaf(vy, wy, v * w));
                        }
                        if (owner) {
                            float4* pn = reinterpret_cast<float4*>(a.nrm + node * TC_N + 64 * nb + c0);
                            float4* pv = reinterpret_cast<float4*>(a.vw + node * TC_N + 64 * nb + c0);
#pragma unroll
                            for (int u = 0; u < 16; u += 4) {
                                pn[u >> 2] = make_float4(nr[u], nr[u + 1], nr[u + 2], nr[u + 3]);
                                pv[u >> 2] = make_float4(dt[u], dt[u + 1], dt[u + 2], dt[u + 3]);
                            }
                        }
                    }
                    if (a.Y != nullptr) {
                        // W (columns 64-127 of the block) straight from TMEM in the m16n8 fragment shape: the four threads of
                        // a row write 32 contiguous bytes, whole sectors, no staging and no block barrier
                        const int u = lane & 3, rq = lane >> 2;
#pragma unroll 1
                        for (int hl = 0; hl < 2; ++hl) {
                            const uint32_t t1 = tmem + ((uint32_t)(q * 32 + 16 * hl) << 16) + buf * 256 + 64 + MIXC * h;
                            uint32_t d0[16], d1[16];
                            if (MIXC == 32) {
                                tmem_ld_16x256b_x4(t1, d0);
                                tmem_ld_16x256b_x4(t1 + 128, d1);
                            } else {
                                tmem_ld_16x256b_x2(t1, d0);
                                tmem_ld_16x256b_x2(t1 + 128, d1);
                            }
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int rh = 0; rh < 2; ++rh) {
                                bool live;
                                const int64_t m = tile_row(a, tile, q * 32 + 16 * hl + 8 * rh + rq, live);
                                float* yrow = a.Y + m * a.ldy + 64 * nb + MIXC * h + 2 * u;
#pragma unroll
                                for (int g = 0; g < MIXC / 8; ++g) {
                                    const float w0v = fmaf(__uint_as_float(d1[4 * g + 2 * rh]), 1.0f / 2048.0f, __uint_as_float(d0[4 * g + 2 * rh]));
                                    const float w1v = fmaf(__uint_as_float(d1[4 * g + 2 * rh + 1]), 1.0f / 2048.0f, __uint_as_float(d0[4 * g + 2 * rh + 1]));
                                    if (live) *reinterpret_cast<float2*>(yrow + 8 * g) = make_float2(w0v, w1v);
                                }
                            }
                        }
                    }
                    tc_fence_before();                       // all TMEM reads of this accumulator are done
                    mbar_arrive(&S->acc_empty[buf]);
                    continue;
                }
                // four quarter-steps (column segment sgm, 16-lane half hl); the TMEM loads of step i+1 are in flight while
                // step i is combined and stored (a load + wait costs 250-650 cycles, as much as the stores it feeds)
                const int u = lane & 3, rq = lane >> 2;
                uint32_t dd[2][32];                          // [ping-pong][d0 (16) | d1 (16)]
                {
                    const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + 32 * h;
                    tmem_ld_16x256b_x4(t0, dd[0]);
                    tmem_ld_16x256b_x4(t0 + 128, dd[0] + 16);
                }
#pragma unroll
                for (int st = 0; st < 2 * NSEG; ++st) {
                    const int sgm = st >> 1, hl = st & 1;
                    const int c0 = (sgm * HQ + h) * 32;      // first column (inside the N-block) of this warp's segment
                    const int col0 = nb * TC_N + c0;
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (st < 2 * NSEG - 1) {
                        const int sg2 = (st + 1) >> 1, hl2 = (st + 1) & 1;
                        const uint32_t t1 = tmem + ((uint32_t)(q * 32 + 16 * hl2) << 16) + buf * 256 + (sg2 * HQ + h) * 32;
                        tmem_ld_16x256b_x4(t1, dd[(st + 1) & 1]);
                        tmem_ld_16x256b_x4(t1 + 128, dd[(st + 1) & 1] + 16);
                    } else {                                 // all TMEM reads of this accumulator are done
                        tc_fence_before();
                        mbar_arrive(&S->acc_empty[buf]);
                    }
                    const uint32_t* d0 = dd[st & 1];
                    const uint32_t* d1 = dd[st & 1] + 16;
                    float2 bb[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) bb[g] = *reinterpret_cast<const float2*>(bias_o + col0 + 8 * g + 2 * u);
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const int64_t m = (int64_t)tile * TC_M + q * 32 + 16 * hl + 8 * rh + rq;
                        float* yrow = a.Y + m * a.ldy + col0 + 2 * u;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float v0 = fmaf(__uint_as_float(d1[4 * g + 2 * rh]), 1.0f / 2048.0f, __uint_as_float(d0[4 * g + 2 * rh])) + bb[g].x;
                            float v1 = fmaf(__uint_as_float(d1[4 * g + 2 * rh + 1]), 1.0f / 2048.0f, __uint_as_float(d0[4 * g + 2 * rh + 1])) + bb[g].y;
                            if (act_o == 1) {                // SiLU; |error| < 2e-7 of max(|v|, 1)
                                v0 = __fdividef(v0, 1.0f + __expf(-v0));
                                v1 = __fdividef(v1, 1.0f + __expf(-v1));
                            }
                            // the four threads of a row write 32 contiguous bytes: one full sector per row, no staging
                            if (m < a.M) *reinterpret_cast<float2*>(yrow + 8 * g) = make_float2(v0, v1);
                        }
                    }
                }
            }
        }
        if (bad && a.status != nullptr) atomicOr(a.status, 1);
    } else if (warp == EW) {
        // ================= weight-stage producer =================
        if (lane == 0) {
            uint32_t cnt = 0;
            const int KSN = a.K / TC_KS;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int phase = 0; phase < NPH; ++phase)
                    for (int nb = 0; nb < NB; ++nb)
                        for (int ksl = 0; ksl < 128 / TC_KS; ++ksl, ++cnt) {
                            const uint32_t s = cnt % TC_NSTAGE, ph = (cnt / TC_NSTAGE) & 1;
                            mbar_wait(&S->empty[s], ph ^ 1);
                            mbar_expect_tx(&S->full[s], TC_STAGE_BYTES);
                            const int ks = phase * (128 / TC_KS) + ksl;
                            bulk_g2s(Bst + s * TC_STAGE_BYTES, Wfmt + (size_t)(nb * KSN + ks) * TC_STAGE_BYTES,
                                     TC_STAGE_BYTES, &S->full[s]);
                        }
                for (int nb = nb2_lo; nb < nb2_hi; ++nb)                       // second layer: K = 128 hidden units
                    for (int ksl = 0; ksl < 128 / TC_KS; ++ksl, ++cnt) {
                        const uint32_t s = cnt % TC_NSTAGE, ph = (cnt / TC_NSTAGE) & 1;
                        mbar_wait(&S->empty[s], ph ^ 1);
                        mbar_expect_tx(&S->full[s], TC_STAGE_BYTES);
                        bulk_g2s(Bst + s * TC_STAGE_BYTES, W2fmt + (size_t)(nb * (128 / TC_KS) + ksl) * TC_STAGE_BYTES,
                                 TC_STAGE_BYTES, &S->full[s]);
                    }
            }
        }
    } else {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t cnt = 0, acc_base = 0, ph_it = 0;
            const uint32_t a_hi0 = smem_u32(A_hi), a_lo0 = smem_u32(A_lo), b0 = smem_u32(Bst);
            auto kstages = [&](uint32_t d0, uint32_t d1, int k_first_stage, bool first) {
                for (int ksl = 0; ksl < 128 / TC_KS; ++ksl, ++cnt) {
                    const uint32_t s = cnt % TC_NSTAGE, ph = (cnt / TC_NSTAGE) & 1;
                    mbar_wait(&S->full[s], ph);
                    tc_fence_after();
                    const uint32_t bh = b0 + s * TC_STAGE_BYTES, bl = bh + TC_HALF_STAGE;
#pragma unroll
                    for (int j = 0; j < TC_KS / 16; ++j) {
                        const uint32_t koff_a = (uint32_t)(ksl * (TC_KS / 8) + 2 * j) * CORE_LBO;
                        const uint32_t koff_b = (uint32_t)(2 * j) * CORE_LBO;
                        const uint32_t acc = (first && ksl == 0 && j == 0) ? 0u : 1u;
                        umma_f16(d0, make_desc(a_hi0 + koff_a), make_desc(bh + koff_b), acc);
                        umma_f16(d1, make_desc(a_hi0 + koff_a), make_desc(bl + koff_b), acc);
                        umma_f16(d1, make_desc(a_lo0 + koff_a), make_desc(bh + koff_b), 1u);
                    }
                    umma_commit(&S->empty[s]);                               // frees the weight stage
                }
                (void)k_first_stage;
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, acc_base += NB + NB2c) {
                for (int phase = 0; phase < NPH; ++phase, ++ph_it) {
                    TCS(1, 100)
                    mbar_wait(&S->a_ready, ph_it & 1);
                    tc_fence_after();
                    TCS(1, 101)
                    for (int nb = 0; nb < NB; ++nb) {
                        const uint32_t acc_it = acc_base + nb, buf = acc_it & 1;
                        if (phase == 0) {
                            mbar_wait(&S->acc_empty[buf], ((acc_it >> 1) & 1) ^ 1);
                            tc_fence_after();
                        }
                        const uint32_t d0 = tmem + buf * 256, d1 = d0 + 128;
                        TCS(1, 110 + nb)
                        kstages(d0, d1, phase * (128 / TC_KS), phase == 0);
                        if (phase == NPH - 1) umma_commit(&S->acc_full[buf]);  // publishes (D0, D1)
                        TCS(1, 120 + nb)
                    }
                    umma_commit(&S->a_free);                                 // A buffers may be overwritten
                }
                if (two) {                                                   // second layer from the hidden tile in A
                    TCS(1, 130)
                    mbar_wait(&S->a_ready, ph_it & 1);
                    tc_fence_after();
                    TCS(1, 131)
                    for (int nb = nb2_lo; nb < nb2_hi; ++nb) {
                        const uint32_t acc_it = acc_base + NB + (nb - nb2_lo), buf = acc_it & 1;
                        mbar_wait(&S->acc_empty[buf], ((acc_it >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t d0 = tmem + buf * 256, d1 = d0 + 128;
                        TCS(1, 140 + nb)
                        kstages(d0, d1, 0, true);
                        umma_commit(&S->acc_full[buf]);
                        TCS(1, 150 + nb)
                    }
                    umma_commit(&S->a_free);
                    ++ph_it;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EW + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

size_t tc_smem_bytes(int) { return (size_t)131072 + TC_NSTAGE * TC_STAGE_BYTES + 32768 + 128 + 512 * 4 + 64; }
static_assert(131072 + TC_NSTAGE * TC_STAGE_BYTES + 32768 + 128 + 512 * 4 + 64 <= 232448, "shared memory budget");
static_assert(sizeof(TcSmem) <= 128, "TcSmem must fit its 128-byte slot");

}  // namespace

// Host: format W[N][K] (checkpoint layout, out-major) into 32 KB stages [(nb, ks)] = [hi 128x64 | lo 128x64],
// each half in the canonical core-matrix layout the MMA descriptors describe.
void tc_format_weight(const float* W, int N, int K, std::vector<uint8_t>& out) {
    const int NB = N / TC_N, KSN = K / TC_KS;
    out.assign((size_t)NB * KSN * TC_STAGE_BYTES, 0);
    for (int nb = 0; nb < NB; ++nb)
        for (int ks = 0; ks < KSN; ++ks) {
            uint8_t* st = out.data() + (size_t)(nb * KSN + ks) * TC_STAGE_BYTES;
            for (int r = 0; r < TC_N; ++r)
                for (int k = 0; k < TC_KS; ++k) {
                    const float w = W[(size_t)(nb * TC_N + r) * K + ks * TC_KS + k];
                    const __half h = __float2half_rn(w);
                    const __half l = __float2half_rn((w - __half2float(h)) * 2048.0f);
                    const size_t off = (size_t)(k / 8) * CORE_LBO + (r / 8) * CORE_SBO + (r % 8) * 16 + (k % 8) * 2;
                    *reinterpret_cast<__half*>(st + off) = h;
                    *reinterpret_cast<__half*>(st + TC_HALF_STAGE + off) = l;
                }
        }
}

// mu_channel_mix.weight [256][128] (rows 0-127 -> V, 128-255 -> W) with its rows regrouped for the mix-reduce epilogue:
// block nb = [V rows 64 nb .. 64 nb + 63 | W rows 64 nb .. 64 nb + 63].
void tc_format_mix_weight(const float* W, std::vector<uint8_t>& out) {
    std::vector<float> P((size_t)2 * TC_N * 128);
    for (int r = 0; r < 2 * TC_N; ++r) {
        const int nb = r / TC_N, within = r % TC_N;
        const int src = within < 64 ? 64 * nb + within : TC_N + 64 * nb + (within - 64);
        for (int k = 0; k < 128; ++k) P[(size_t)r * 128 + k] = W[(size_t)src * 128 + k];
    }
    tc_format_weight(P.data(), 2 * TC_N, 128, out);
}

bool tc_linear_supported(const LinearArgs& a) {
    const bool k_ok = a.K == 128 || (a.K == 256 && a.N <= 2 * TC_N);   // K=256 keeps every N-block's accumulator live
    if (a.mixred && (a.K != 128 || a.N != 2 * TC_N || a.bias != nullptr || a.A2 != nullptr || a.W2fmt != nullptr ||
                     a.nrm == nullptr || a.vw == nullptr || a.M != 3 * a.n_nodes))
        return false;
    if (a.W2fmt != nullptr && (a.N != TC_N || a.N2 % TC_N != 0 || a.N2 <= 0 || a.N2 > 384)) return false;
    return a.N % TC_N == 0 && a.N <= 384 && k_ok && (a.A2 == nullptr || a.K1 == 128) && a.lda % 4 == 0 && a.ldy % 4 == 0;
}

int launch_tc_linear(rlvd_engine* e, cudaStream_t s, const LinearArgs& a_in, const void* Wfmt) {
    LinearArgs a = a_in;
    a.status = e->status.as<int>();
    RLVD_CHECK(tc_linear_supported(a), "tc_linear: unsupported shape M=%d K=%d N=%d", a.M, a.K, a.N);
    if (a.M <= 0) return 0;
    if (!e->attr_tc_linear) {                 // per engine = per device: the attribute is a per-device property
        RLVD_CUDA(cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc_smem_bytes(256)));
        e->attr_tc_linear = true;
    }
    const int n_tiles = a.mixred ? (a.n_nodes + 39) / 40 : (a.M + TC_M - 1) / TC_M;
    const int grid = n_tiles < e->sm_count ? n_tiles : e->sm_count;
#ifdef RLVD_TC_TIMELINE   // (tag, clock64) stamps of block 0: epilogue thread 0 and the MMA issuer (RLVD_EXTRA_NVCC_FLAGS=-DRLVD_TC_TIMELINE)
    static long long* dbg_buf = nullptr;
    static int dbg_calls[4] = {0, 0, 0, 0};
    if (dbg_buf == nullptr) cudaMalloc(&dbg_buf, 2048 * 8);
    const int kind = a.mixred ? 0 : (a.W2fmt != nullptr ? (a.K == 128 ? 1 : 2) : 3);
    const bool dump = a.M > 100000 && ++dbg_calls[kind] == 12;
    if (dump) { cudaMemsetAsync(dbg_buf, 0, 2048 * 8, s); a.dbg = dbg_buf; }
#endif
    Span sp(e, s, FAM_NODE);
    // two-layer launches with fewer tiles than half the SMs (single-replica stepping): split the second layer's N-blocks over CTAs
    const int gy = a.W2fmt != nullptr ? std::max(1, std::min(a.N2 / TC_N, e->sm_count / grid)) : 1;
    tc_linear_kernel<<<dim3(grid, gy), TC_THREADS, tc_smem_bytes(a.K), s>>>(a, reinterpret_cast<const uint8_t*>(Wfmt),
                                                                  reinterpret_cast<const uint8_t*>(a.W2fmt));
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
#ifdef RLVD_TC_TIMELINE
    if (dump) {
        cudaStreamSynchronize(s);
        static long long h[2048];
        cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
        printf("tc_linear timeline kind %d (0 mixred, 1 two-layer K128, 2 two-layer K256, 3 plain) M %d K %d N %d N2 %d\n", kind, a.M, a.K, a.N, a.N2);
        for (int role = 0; role < 2; ++role) {
            const long long t0 = h[1];
            long long prev = h[role * 1024 + 1];
            for (int i = 0; i < 500 && h[role * 1024 + 2 * i + 1] != 0; ++i) {
                const long long tag = h[role * 1024 + 2 * i], c = h[role * 1024 + 2 * i + 1];
                if (c - t0 > 120000) break;
                printf("  role %d tag %3lld t %7lld (+%5lld)\n", role, tag, c - t0, c - prev);
                prev = c;
            }
        }
    }
#endif
    return 0;
}

}  // namespace rlvd
