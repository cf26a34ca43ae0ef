// K2 + K3a + K3c + K4 fused for structures of at most 256 atoms — the BASELINE workload (255/256-atom replicas).
// "Edge stream" design: the per-edge embedding is packed ONCE per representation call into tensor-core-ready tiles
// that the three layers x four channel slices then only stream.
//
// Replaces, inside rgnn's PaiNN (reached from rlsim/drl/simulator.py:56): the Bessel / cosine-cutoff / unit-vector
// embedding, `filter_net` (nn.Linear 20 -> 1152, materialised as [E,1152] by the eager path), the index_select
// gathers of x[j] / mu[j] and torch_scatter.scatter_add (SURVEY.md §8a R2, R3, R5 edge part, K4).
//
// edge_pack_kernel (once per call, one CTA per structure)
//   Splits the structure's receivers into NPIPE contiguous ranges of equal edge-group count and writes, per range,
//   a stream of 32-slot tiles.  Slots come in receiver-aligned groups of 8 (short rows padded with zero filters).
//   A tile is 4096 B of MMA operand + 464 B of meta:
//     operand  [32 slots][64 fp16], K-major core-matrix layout:  b_hi[21] | b_lo[21] | b_hi[21] | 0
//              b = 2^10 * {phi_k(d) * fcut(d) (k < 20), fcut(d)}, split b = b_hi + b_lo in fp16 (22 significant bits)
//     meta     4*j[32] u16 (local sender index, pre-scaled for the dp2a gather address) | new_receiver[4] u8 (one per group) | pad |
//              dir_x[32] | dir_y[32] | dir_z[32]
//
// edge_stream_kernel (per layer; CTA = one structure x one 32-channel slice; 1 CTA per SM; batches of less than ~two waves put
// S <= 3 CTAs on a (structure, slice), each with NPIPE of the structure's NPIPE x S receiver ranges)
//   The slice of the structure's node table lives in shared memory (gathers never leave the SM):
//     row(j) = kappa*x_q | kappa*x_R | kappa*x_m*mu_x | kappa*x_m*mu_y | kappa*x_m*mu_z        (5 x 32 fp32)
//   Filters run on tcgen05 with the weight operand resident in TMEM (A rows = dq | dmuR | dmumu | dmumu filters of
//   the slice), pre-split on the host as a_hi[21] | a_hi[21] | a_lo[21] | 0 so that ONE K=64 fp16 MMA chain evaluates
//   a_hi.b_hi + a_hi.b_lo + a_lo.b_hi (fp32-accurate, 4 dispatches):      D[128, 32 slots] = A . Bt
//   A TMEM lane quarter can only be read by the warps with warp%4 == quarter, so each of a pipeline's four warps
//   takes the role whose filter rows sit in its quarter:
//        role 0 (q)   : dq        += D * x_q[j]                  (+ all global writes)
//        role 1 (R)   : dmuR_xyz  += D * x_R[j] * dir_xyz
//        role 2 (M)   : dmuM_xyz  += D * (x_m mu_xyz)[j]
//        role 3       : no rows; one thread feeds the pipeline (bulk copies, tcgen05.mma issue, commits)
//   warp%4 is also the SM sub-partition (issue scheduler) a warp runs on and the roles cost 0..85 issue slots per
//   group, so the operand is kept in TMEM in four row rotations and pipeline p uses rotation pipe_rot(p): every scheduler
//   then hosts a mix of roles instead of one scheduler carrying all the R warps.
//   The kernel is shared-memory-bandwidth bound (gathers 5 x 128 B per slot), so every sender row is gathered exactly
//   once.  When a receiver's row ends, warps 1-3 drop their partial sums into a per-SM scratch slot (55 KB per SM,
//   L2-resident); warp 0 adds them to mu_in and writes mu_out one tile later — the tile barriers already order
//   those accesses, so the combine needs no synchronisation of its own.  Accumulators (TMEM) and operand buffers
//   are double-buffered: tile t+1's MMAs run while tile t is consumed.
//   No atomics; sums run in CSR order (bitwise reproducible).  NPIPE such 4-warp pipelines run on disjoint receiver
//   ranges; while one waits for its next tile the others fill the issue slots.  kappa = 2^-(sA+10) undoes the fp16
//   range scaling exactly.
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace rlvd {

namespace {

constexpr int SL = 32;                    // channels per slice
constexpr int NSLICE = F / SL;            // 4
#ifndef RLVD_NPIPE
#define RLVD_NPIPE 6
#endif
constexpr int NPIPE = RLVD_NPIPE;         // 4-warp pipelines per CTA (6: 768 threads at <= 85 registers; experiments: -DRLVD_NPIPE=5)
constexpr int TE = 32;                    // slots per tile = MMA N
constexpr int GS = 8;                     // slots per receiver-aligned group
constexpr int KH = 64;                    // fp16 k-extent of the split operands
constexpr int NK = NRBF + 1;              // 21: 20 Bessel functions + the bias column (fcut)
constexpr int TILE_B = TE * KH * 2;       // 4096
constexpr int META = 464;                 // 4*j[32] u16 | newrecv[4] u8 | pad[12] | dir[3][32] f32
constexpr int META_DIR = 80;
constexpr int META_FLAG = 64;
constexpr int TILE_BYTES = TILE_B + META; // 4560
constexpr int MAXN = 256;
constexpr int ST_THREADS = NPIPE * 128;
constexpr int MAX_SPLIT = 3;              // small batches: a structure's receivers are spread over up to 3 CTAs per channel slice
constexpr int MAXR = NPIPE * MAX_SPLIT;   // receiver ranges per structure (NPIPE per CTA)
constexpr int B_SCALE_LOG2 = 10;
constexpr int TM_D = 0;                   // TMEM columns: accumulator buffer b of pipe p at (2p+b)*32
constexpr int TM_A = NPIPE * 2 * TE;      // 4 row rotations of the weight operand (64 fp16 = 32 columns each)
constexpr int NROT = 4;
// rotation of pipeline p: with role costs q 45 / R 75 / M 75 / feeder 0 issue slots per group this spreads the six
// pipelines' warps over the four schedulers as 315 / 270 / 315 / 270
__device__ __forceinline__ int pipe_rot(int p) { return p < 4 ? p : (p == 4 ? 0 : 2); }
static_assert(NPIPE >= 1 && NPIPE <= 6, "pipe_rot and the TMEM budget cover at most 6 pipelines");
constexpr int TM_COLS = 512;
constexpr int A_SLICE_WORDS = NROT * 128 * 32;   // u32 per (layer, slice): [rotation][row][32]
constexpr int NB_STAGE = 2;               // operand buffers per pipeline
constexpr int NM_STAGE = 5;               // meta buffers per pipeline (tiles t-? .. t+4 in flight)
constexpr int NP_STAGE = 4;               // partial-sum buffers per pipeline (tile t uses t % 4)
constexpr int PART_ROWS = 6;              // partial-sum rows per finished receiver: R x,y,z | M x,y | M z
constexpr int PART_SLOT = PART_ROWS * SL; // floats
constexpr int PART_PIPE = NP_STAGE * (TE / GS) * PART_SLOT;   // floats per pipeline: [tile%4][<= 4 finished receivers]
constexpr int PART_SM = NPIPE * PART_PIPE;                    // floats per SM (73,728 B)
constexpr int MAX_SMID = 256;
#ifndef RLVD_PF
#define RLVD_PF 0
#endif
#ifndef RLVD_SPLIT
#define RLVD_SPLIT 0
#endif
constexpr int PF_DIST = RLVD_PF;          // tiles between the L2 prefetch and the bulk copy of a tile (0 = off)
constexpr float PI_F = 3.14159265358979323846f;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TE >> 3) << 17) | (8u << 24);   // f16 x f16 -> f32, M=128, N=32

static_assert(TM_A + NROT * 32 <= TM_COLS, "TMEM budget");

struct StBars {
    uint64_t b_full[NPIPE][NB_STAGE], mma_done[NPIPE][2], acc_free[NPIPE][2];
    int chg[NPIPE][NP_STAGE][TE / GS];    // receivers finished in a tile (tile t uses slot t % 4)
    uint32_t tmem_base;
};

// Waits use try_wait WITH a suspend-time hint (TRYWAIT + NANOSLEEP.SYNCS: the warp sleeps in hardware until the phase
// completes).  The hint-less form spins; in this kernel a third of all issued instructions were such spins (ncu,
// profiles/ncu_stream_r01d.txt) taken from the warps that had work.
__device__ __forceinline__ void st_wait(uint64_t* bar, uint32_t parity) {
    // The whole loop is PTX: 7 SASS instructions per wake-up instead of the compiler's 11 (NANOSLEEP.SYNCS wakes on every
    // mbarrier event of the CTA, so waiting warps issue a third of the kernel's instructions).  Measured: neither this nor
    // dropping the bounded-wait counter changes the kernel time — the spins only take issue slots nobody else wanted.
    // The counter stays: a protocol bug must trap, never hang the GPU.
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b32 n;\n\t"
        "mov.b32 n, 0;\n"
        "ST_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra ST_WAIT_DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.gt.u32 q, n, 4000000;\n\t"
        "@q trap;\n\t"
        "bra ST_WAIT_LOOP;\n"
        "ST_WAIT_DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// TMEM loads carry no "memory" clobber: they touch no memory the compiler knows about, and a clobber would pin every
// shared-memory load of the next group behind this group's FMAs (no software pipelining across groups).  The wait
// takes the loaded registers as in/out operands so that their uses stay behind it.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld8(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
}

// Packed fp32 pairs (FFMA2 / FMUL2, sm_100): the per-slot arithmetic runs on (even slot, odd slot) pairs, one issue slot
// for two FMAs.  Same operations in the same order per lane as the scalar form.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// shared-memory gather with an explicit 32-bit address (the node table is read-only while the roles run)
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
// Gather address in one instruction (IDP.2A issues at the IMAD rate, scripts/probes/idp_rate.cu): the tile meta carries
// 4*j as u16 pairs; dp2a.lo adds (low half) * b.byte0, dp2a.hi adds (high half) * b.byte3 to the base.
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ float sum2(f32x2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}

// ------------------------------------------------------------------------------------------------------------------
// edge_pack_kernel
// ------------------------------------------------------------------------------------------------------------------

// 21 embedding values of one edge: phi_k(d)*fcut(d) (k < 20) and fcut(d).  sin(freqs[k] d): one accurate sincos of
// theta = freqs[0]*d (fp32 product error folded in), an angle-addition tree of depth <= 5 for k*theta, and the deviation
// freqs[k] - (k+1)*freqs[0] folded in to first order (max abs error 1.4e-6, below the 1.9e-6 argument-rounding error of
// evaluating sin(fl(freqs[k]*d)) directly in fp32).
__device__ __forceinline__ void edge_embedding(float d, float f1, const float* __restrict__ feps, float cutoff, float* v /*21*/) {
    const float inv_d = 1.0f / d;
    const float fc = d < cutoff ? 0.5f * (cosf(d * (PI_F / cutoff)) + 1.0f) : 0.f;
    const float scale = inv_d * fc;
    const float pr = f1 * d, perr = fmaf(f1, d, -pr);
    float sn, cs;
    sincosf(pr, &sn, &cs);
    float s[NRBF + 1], c[NRBF + 1];
    s[1] = fmaf(perr, cs, sn);
    c[1] = fmaf(-perr, sn, cs);
#define RLVD_ADD(a, b) { s[(a) + (b)] = fmaf(s[a], c[b], c[a] * s[b]); c[(a) + (b)] = fmaf(c[a], c[b], -(s[a] * s[b])); }
    RLVD_ADD(1, 1) RLVD_ADD(2, 1) RLVD_ADD(2, 2) RLVD_ADD(4, 1) RLVD_ADD(4, 2) RLVD_ADD(4, 3) RLVD_ADD(4, 4)
    RLVD_ADD(8, 1) RLVD_ADD(8, 2) RLVD_ADD(8, 3) RLVD_ADD(8, 4) RLVD_ADD(8, 5) RLVD_ADD(8, 6) RLVD_ADD(8, 7)
    RLVD_ADD(8, 8) RLVD_ADD(16, 1) RLVD_ADD(16, 2) RLVD_ADD(16, 3) RLVD_ADD(16, 4)
#undef RLVD_ADD
#pragma unroll
    for (int k = 0; k < NRBF; ++k) v[k] = fmaf(__ldg(feps + k) * d, c[k + 1], s[k + 1]) * scale;
    v[NRBF] = fc;
}


// Writes slot `r` of the tile at `tile`: operand row (8 x 16 B, one per core column) and meta.
__device__ __forceinline__ void write_slot(uint8_t* tile, int r, const uint4* row8, int j, float dx, float dy, float dz) {
    uint8_t* b = tile + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) *reinterpret_cast<uint4*>(b + kc * 512) = row8[kc];
    uint8_t* m = tile + TILE_B;
    reinterpret_cast<uint16_t*>(m)[r] = (uint16_t)(j * 4);   // the gather address is ONE dp2a: 4j * (NODE_BYTES / 4) + base
    reinterpret_cast<float*>(m + META_DIR)[r] = dx;
    reinterpret_cast<float*>(m + META_DIR + 128)[r] = dy;
    reinterpret_cast<float*>(m + META_DIR + 256)[r] = dz;
}

__global__ void __launch_bounds__(256, 3) edge_pack_kernel(int n_struct, const int* __restrict__ node_ptr,
                                                        const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                        const int8_t* __restrict__ shift, const float* __restrict__ pos,
                                                        const float* __restrict__ cell, const float* __restrict__ freqs,
                                                        const float* __restrict__ feps, float cutoff,
                                                        int* __restrict__ tile_counter, int tile_capacity,
                                                        uint8_t* __restrict__ tiles, int4* __restrict__ pipe_info,
                                                        int* __restrict__ err_flag, int* __restrict__ status_word, const int* __restrict__ Z,
                                                        const int* __restrict__ z2s, float* __restrict__ l0A, int mode,
                                                        int n_splits, int NR /* receiver ranges per structure: NPIPE x stream split */) {
    // mode 0: allocate tiles and fill them (one CTA per structure).  Small batches run mode 1 (allocation only) and then
    // mode 2 with grid (n_struct, n_splits): CTA (g, y) fills the receivers of warp-stride class y; every CTA repeats the
    // (cheap, deterministic) scan and takes the tile bases mode 1 published in pipe_info.
    __shared__ int goff[MAXN + 1];        // exclusive prefix of per-receiver group counts
    // per-warp species-sorted slot records for the layer-0 sums: v[32][21] and (dir_x, dir_y, dir_z, 1)[32]
    __shared__ __align__(16) float stg_v[8][32 * NK];
    __shared__ __align__(16) float4 stg_d[8][32];
    __shared__ float spos[3 * MAXN];      // the structure's positions and species: senders are gathered from here
    __shared__ int8_t ssp[MAXN];
    __shared__ int wsum[8];
    __shared__ int bnd[MAXR + 1];         // receiver range boundaries of the pipelines
    __shared__ int tbeg[MAXR + 1];        // first tile of each pipeline (absolute)
    const int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int node0 = __ldg(node_ptr + g), n = __ldg(node_ptr + g + 1) - node0;
    if (n <= 0) {
        if (tid < NR && mode != 2) pipe_info[g * NR + tid] = make_int4(0, 0, 0, 0);
        return;
    }
    for (int t = tid; t < 3 * n; t += 256) spos[t] = __ldg(pos + 3 * (int64_t)node0 + t);
    if (tid < n) {
        int sp = 0;
        if (l0A != nullptr) {
            sp = __ldg(z2s + min(max(__ldg(Z + node0 + tid), 0), 99));
            if (sp < 0) { atomicExch(err_flag + 1, 1); sp = 0; }    // not a model species: the gated fallback runs
        }
        ssp[tid] = (int8_t)sp;
    }
    // ---- group counts and their block-wide exclusive scan (its barriers also publish spos / ssp)
    int e0 = 0, deg = 0;
    if (tid < n) {
        e0 = __ldg(row_ptr + node0 + tid);
        deg = __ldg(row_ptr + node0 + tid + 1) - e0;
    }
    const int gc = tid < n ? max(1, (deg + GS - 1) / GS) : 0;   // edgeless receivers keep one (zero-filter) group
    int incl = gc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wsum[w];
    goff[tid] = wbase + incl - gc;
    if (tid == 255) goff[256] = wbase + incl;
    __syncthreads();
    const int G = goff[256];   // threads >= n contributed 0, so goff[i] == G for i >= n
    // ---- receiver ranges with (nearly) equal group counts; tile allocation
    if (tid <= NR) {
        int b;
        if (tid == 0) b = 0;
        else if (tid == NR) b = n;
        else {
            const int target = (int)(((long long)tid * G) / NR);
            int lo = 0, hi = n;                       // smallest i in [0, n] with goff[i] >= target
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (goff[mid] >= target) hi = mid; else lo = mid + 1;
            }
            b = lo;
        }
        bnd[tid] = b;
    }
    __syncthreads();
    if (tid == 0 && mode == 2) {
        int ok = 0;
        for (int p = 0; p < NR; ++p) {
            const int4 pi = pipe_info[g * NR + p];
            tbeg[p] = pi.x;
            ok |= pi.y > 0 ? 1 : 0;
        }
        tbeg[MAXR] = ok;
    }
    if (tid == 0 && mode != 2) {
        int total = 0;
        for (int p = 0; p < NR; ++p) total += (goff[bnd[p + 1]] - goff[bnd[p]] + 3) >> 2;
        int base = atomicAdd(tile_counter, total);
        bool ok = base + total <= tile_capacity;
        if (!ok) { atomicExch(err_flag, 1); atomicOr(status_word, 2); }    // engine status bit 1: results of this call are invalid
        for (int p = 0; p < NR; ++p) {
            const int tl = (goff[bnd[p + 1]] - goff[bnd[p]] + 3) >> 2;
            tbeg[p] = base;
            pipe_info[g * NR + p] = make_int4(base, ok ? tl : 0, bnd[p], bnd[p + 1]);
            base += tl;
        }
        tbeg[MAXR] = ok ? 1 : 0;
    }
    __syncthreads();
    if (tbeg[MAXR] == 0 || mode == 1) return;

    float C[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) C[t] = __ldg(cell + 9 * g + t);
    const float f1 = __ldg(freqs);
    const float bscale = (float)(1 << B_SCALE_LOG2);
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);

    // ---- one warp per receiver row, lanes over its slots
    // layer-0 species sums: lane k < 21 owns radial term k of all four blocks (H_x, H_y | H_z, G as two f32x2 pairs)
    const int l0k = min(lane, NK - 1);
    float* const wst = stg_v[warp];
    float4* const wsd = stg_d[warp];
    for (int i = warp + 8 * (int)blockIdx.y; i < n; i += 8 * n_splits) {
        int p = 0;
        while (i >= bnd[p + 1]) ++p;
        const int gi = node0 + i;
        const int re0 = __ldg(row_ptr + gi), rdeg = __ldg(row_ptr + gi + 1) - re0;
        const int nslots = max(1, (rdeg + GS - 1) / GS) * GS;
        const int slot0 = (goff[i] - goff[bnd[p]]) * GS;
        const float xi = spos[3 * i], yi = spos[3 * i + 1], zi = spos[3 * i + 2];
        f32x2 l0xy[3], l0zg[3];
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) l0xy[sp] = l0zg[sp] = pk2(0.f, 0.f);
        for (int s0 = 0; s0 < nslots; s0 += 32) {
            const int s = s0 + lane;
            const bool live = s < rdeg;
            const int jg = live ? __ldg(col + re0 + s) : 0;
            // layer-0 sums: slot records are staged sorted by sender species (CSR order within a species)
            int at = 0, c0 = 0, c1 = 0, c2 = 0;
            if (l0A != nullptr) {
                const int my_sp = live ? (int)ssp[jg - node0] : -1;     // -1 = padding slot
                const unsigned m0 = __ballot_sync(0xffffffffu, my_sp == 0), m1 = __ballot_sync(0xffffffffu, my_sp == 1),
                               m2 = __ballot_sync(0xffffffffu, my_sp == 2);
                const unsigned below = (1u << lane) - 1u;
                c0 = __popc(m0); c1 = __popc(m1); c2 = __popc(m2);
                at = my_sp == 0 ? __popc(m0 & below) : my_sp == 1 ? c0 + __popc(m1 & below) : c0 + c1 + __popc(m2 & below);
            }
            if (s < nslots) {
            const int slot = slot0 + s;
            uint8_t* tile = tiles + (size_t)(tbeg[p] + (slot >> 5)) * TILE_BYTES;
            const int r = slot & 31;
            if ((s & (GS - 1)) == 0) tile[TILE_B + META_FLAG + (r >> 3)] = (uint8_t)(s == 0 && i > bnd[p] ? 1 : 0);   // new receiver starts
            if (live) {
                const int e = re0 + s;
                const float S0 = (float)shift[3 * (int64_t)e], S1 = (float)shift[3 * (int64_t)e + 1], S2 = (float)shift[3 * (int64_t)e + 2];
                const int jl = jg - node0;
                float rx = spos[3 * jl] - xi, ry = spos[3 * jl + 1] - yi, rz = spos[3 * jl + 2] - zi;
                rx += S0 * C[0] + S1 * C[3] + S2 * C[6];
                ry += S0 * C[1] + S1 * C[4] + S2 * C[7];
                rz += S0 * C[2] + S1 * C[5] + S2 * C[8];
                const float d = sqrtf(rx * rx + ry * ry + rz * rz);
                const float inv_d = 1.0f / d;
                float v[NK];
                edge_embedding(d, f1, feps, cutoff, v);
                if (l0A != nullptr) {
                    float* rec = wst + at * NK;
#pragma unroll
                    for (int k = 0; k < NK; ++k) rec[k] = v[k];
                    wsd[at] = make_float4(rx * inv_d, ry * inv_d, rz * inv_d, 1.0f);
                }
                // k' = 0..20 hi | 21..41 lo | 42..62 hi | 63 zero, two fp16 per 32-bit word; converted pairwise (one
                // cvt.rn.f16x2.f32 per word): words 0-9 (hi,hi), 10 (hi20, lo0), 11-20 (lo,lo), 21-30 = words 0-9, 31 (hi20, 0).
                // Same values as hi = fp16(sv), lo = fp16(sv - float(hi)) computed one by one.
                float sv[NK], lf[NK];
#pragma unroll
                for (int k = 0; k < NK; ++k) sv[k] = v[k] * bscale;
                uint32_t wd[32];
#pragma unroll
                for (int w = 0; w < 10; ++w) {
                    const __half2 h2 = __floats2half2_rn(sv[2 * w], sv[2 * w + 1]);
                    const float2 f = __half22float2(h2);
                    lf[2 * w] = sv[2 * w] - f.x;
                    lf[2 * w + 1] = sv[2 * w + 1] - f.y;
                    wd[w] = *reinterpret_cast<const uint32_t*>(&h2);
                    wd[21 + w] = wd[w];
                }
                const float h20 = __half2float(__float2half_rn(sv[NK - 1]));
                lf[NK - 1] = sv[NK - 1] - h20;
                {
                    const __half2 m = __floats2half2_rn(sv[NK - 1], lf[0]);
                    const __half2 z = __floats2half2_rn(sv[NK - 1], 0.f);
                    wd[10] = *reinterpret_cast<const uint32_t*>(&m);
                    wd[31] = *reinterpret_cast<const uint32_t*>(&z);
                }
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    const __half2 l2 = __floats2half2_rn(lf[2 * i + 1], lf[2 * i + 2]);
                    wd[11 + i] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                uint4 row8[8];
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) row8[kc] = make_uint4(wd[4 * kc], wd[4 * kc + 1], wd[4 * kc + 2], wd[4 * kc + 3]);
                write_slot(tile, r, row8, jg - node0, rx * inv_d, ry * inv_d, rz * inv_d);
            } else {
                uint4 row8[8];
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) row8[kc] = z4;
                write_slot(tile, r, row8, 0, 0.f, 0.f, 0.f);
            }
            }
            if (l0A != nullptr) {
                // species-resolved radial sums of this chunk of <= 32 edges:
                //   G[sp][k] += v_k,  H_a[sp][k] += v_k dir_a      (DESIGN.md "layer 0 collapses to per-node GEMMs")
                __syncwarp();
                int u = 0;
#pragma unroll
                for (int sp = 0; sp < 3; ++sp) {
                    const int uend = sp == 0 ? c0 : sp == 1 ? c0 + c1 : c0 + c1 + c2;
#pragma unroll 4
                    for (; u < uend; ++u) {
                        const float vk = wst[u * NK + l0k];
                        const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(wsd + u);   // (dx, dy) | (dz, 1)
                        const f32x2 vv = pk2(vk, vk);
                        l0xy[sp] = fma2(vv, d4.x, l0xy[sp]);
                        l0zg[sp] = fma2(vv, d4.y, l0zg[sp]);
                    }
                }
                __syncwarp();
            }
        }
        if (l0A != nullptr) {
            float* row = l0A + (size_t)gi * L0_COLS;
            if (lane < NK) {
#pragma unroll
                for (int sp = 0; sp < 3; ++sp) {
                    float hx, hy, hz, gg;
                    unpk2(l0xy[sp], hx, hy);
                    unpk2(l0zg[sp], hz, gg);
                    row[sp * NK + lane] = hx;
                    row[64 + sp * NK + lane] = hy;
                    row[128 + sp * NK + lane] = hz;
                    row[192 + sp * NK + lane] = gg;
                }
            } else if (lane < NK + 4) {
                row[(lane - NK) * 64 + 63] = 0.f;       // pad column of each 64-wide block
            }
        }
    }
    // ---- padding groups that complete each pipeline's last tile (zero filters, harmless receiver)
    if (blockIdx.y == 0)
    for (int pk = tid; pk < NR * 24; pk += 256) {
        const int p = pk / 24, k = pk - p * 24;
        const int groups = goff[bnd[p + 1]] - goff[bnd[p]];
        const int ntile = (groups + 3) >> 2;
        const int slot = groups * GS + k;
        if (slot < ntile * TE) {
            uint8_t* tile = tiles + (size_t)(tbeg[p] + (slot >> 5)) * TILE_BYTES;
            const int r = slot & 31;
            uint4 row8[8];
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) row8[kc] = z4;
            write_slot(tile, r, row8, 0, 0.f, 0.f, 0.f);
            if ((r & (GS - 1)) == 0) tile[TILE_B + META_FLAG + (r >> 3)] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// edge_stream_kernel
// ------------------------------------------------------------------------------------------------------------------

template <bool HAS_MU>
struct Tab {
    static constexpr int NODE_BYTES = HAS_MU ? 5 * SL * 4 : 2 * SL * 4;   // 640 / 256
};

__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

struct PipeCtx {
    StBars* S;
    int pipe, lane, quarter;
    uint32_t tmem;
    const uint8_t* tab;      // node table of the slice
    uint8_t* MetaP;          // NM_STAGE meta buffers
    float* Part;             // this SM's / pipeline's partial-sum scratch: [3][4][PART_ROWS][32] (global, L2-resident)
    int n_tiles, ra, rb, node0, ch, pf_eighths;
    long long* dbg;          // RLVD_TIMELINE: clock64 stamps of block 0 / pipeline 0
};

// warp 0: adds the partial sums of the receivers finished in a tile to mu_in and writes mu_out.
template <bool HAS_MU>
__device__ __forceinline__ void combine_mu(const PipeCtx& c, int buf, int count, const float* __restrict__ mu_in,
                                           float* __restrict__ mu_out, int first = 0) {
#pragma unroll 1
    for (int k = first; k < count; ++k) {
        const int i = c.S->chg[c.pipe][buf][k];
        const float* P = c.Part + (buf * (TE / GS) + k) * PART_SLOT + c.lane;
        const int64_t o = (int64_t)(c.node0 + i) * 3 * F + c.ch;
        float v0 = __ldcg(P), v1 = __ldcg(P + SL), v2 = __ldcg(P + 2 * SL);
        if (HAS_MU) {
            v0 += __ldcg(P + 3 * SL) + __ldg(mu_in + o);
            v1 += __ldcg(P + 4 * SL) + __ldg(mu_in + o + F);
            v2 += __ldcg(P + 5 * SL) + __ldg(mu_in + o + 2 * F);
        }
        mu_out[o] = v0;
        mu_out[o + F] = v1;
        mu_out[o + 2 * F] = v2;
    }
}

struct Feed {                // what the thread doing a pipeline's producer duty needs
    const uint8_t* src;      // first tile of the pipeline
    uint8_t* Bp;             // NB_STAGE operand buffers
    uint64_t bdesc0;
    const uint8_t* pf;       // L2 prefetch duty of this feeder (see edge_stream_kernel): a chunk of the node rows a CTA that
    uint32_t pf_bytes;       // starts ~one CTA lifetime from now will stage; 0 = none
};

__device__ __forceinline__ void feed_copy(const PipeCtx& c, const Feed& fd, int k) {
    const uint8_t* sp = fd.src + (size_t)k * TILE_BYTES;
    uint64_t* bar = &c.S->b_full[c.pipe][k & 1];
    mbar_expect_tx(bar, TILE_BYTES);
    bulk_g2s(fd.Bp + (k & 1) * TILE_B, sp, TILE_B, bar);
    bulk_g2s(c.MetaP + (k % NM_STAGE) * META, sp + TILE_B, META, bar);
    if (PF_DIST > 0 && k + PF_DIST < c.n_tiles) prefetch_l2(sp + (size_t)PF_DIST * TILE_BYTES, TILE_BYTES);   // HBM latency is paid here
}
__device__ __forceinline__ void feed_mma(const PipeCtx& c, const Feed& fd, int m) {
    st_wait(&c.S->b_full[c.pipe][m & 1], ((uint32_t)m >> 1) & 1u);
    tc_fence_after();
    const uint32_t d = c.tmem + TM_D + (uint32_t)(c.pipe * 2 + (m & 1)) * TE;
    const uint64_t bd = fd.bdesc0 + (uint64_t)((m & 1) * (TILE_B >> 4));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16_ts(d, c.tmem + TM_A + pipe_rot(c.pipe) * 32 + k * 8, bd + (uint64_t)(k * 64), k);
    umma_commit(&c.S->mma_done[c.pipe][m & 1]);
}

// The pipeline's fourth warp (one thread of it) only feeds the other three.  Tile t uses operand buffer t%2, TMEM
// accumulator t%2, meta buffer t%5; barriers indexed t%2 complete their (t/2)-th phase with tile t.  Every step costs
// this thread 100-450 cycles of issue latency (mbarrier try_wait ~120, four tcgen05.mma + commit ~420, bulk copies
// ~450: measured, profiles/timeline_r01.txt), ~1300 per tile — which is why no consuming warp carries it:
//   wait acc_free(t)  (every consumer has left tile t: accumulator t%2 is free)
//   wait b_full(t+2), issue the MMAs of tile t+2, commit -> mma_done(t+2)
//   wait mma_done(t+2) (operand buffer t%2 read), bulk-copy tile t+4 into it and into meta buffer (t+4)%5 (last used
//   by tile t-1): a copy gets two tile periods to cover its HBM latency.
// PART: 0 = both halves in one thread (layers with mu: all four warps of the pipeline are taken), 1 = MMA half only,
// 2 = copy half only (layer 0 has an idle third warp, so the halves run as two threads and the ~1300-cycle duty, which
// there exceeds what the two consumers need per tile, is cut in two).
template <int PART>
__device__ __forceinline__ void run_feeder(const PipeCtx& c, const Feed& fd) {
    StBars* S = c.S;
    const int pipe = c.pipe, n_tiles = c.n_tiles;
    if (n_tiles <= 0 || c.ra >= c.rb || c.lane != 0) return;
    if (PART == 1) {
#pragma unroll 1
        for (int m = 0; m < n_tiles; ++m) {
            if (m >= 2) st_wait(&S->acc_free[pipe][m & 1], (((uint32_t)m - 2u) >> 1) & 1u);
            feed_mma(c, fd, m);
        }
        return;
    }
    if (PART == 2) {
#pragma unroll 1
        for (int k = 0; k < n_tiles; ++k) {     // mma_done(k-2) also implies every consumer has left tile k-4 >= k-5
            if (k >= 2) st_wait(&S->mma_done[pipe][k & 1], (((uint32_t)k - 2u) >> 1) & 1u);
            feed_copy(c, fd, k);
        }
        return;
    }
    feed_copy(c, fd, 0);
    if (n_tiles > 1) feed_copy(c, fd, 1);
    feed_mma(c, fd, 0);
    if (n_tiles > 1) feed_mma(c, fd, 1);
    if (n_tiles > 2) {
        st_wait(&S->mma_done[pipe][0], 0);
        feed_copy(c, fd, 2);
    }
    if (n_tiles > 3) {
        st_wait(&S->mma_done[pipe][1], 0);
        feed_copy(c, fd, 3);
    }
    const int pf_t = (n_tiles * c.pf_eighths) >> 3;
#pragma unroll 1
    for (int t = 0; t + 2 < n_tiles; ++t) {
        const uint32_t bi = (uint32_t)t & 1u;
        if (t == pf_t && fd.pf_bytes != 0) prefetch_l2(fd.pf, fd.pf_bytes);
        st_wait(&S->acc_free[pipe][bi], ((uint32_t)t >> 1) & 1u);
        feed_mma(c, fd, t + 2);
        if (t + 4 < n_tiles) {
            st_wait(&S->mma_done[pipe][bi], (((uint32_t)t + 2u) >> 1) & 1u);
            feed_copy(c, fd, t + 4);
        }
    }
}

// One consuming warp of one pipeline.  ROLE: see the file header; c.quarter = the warp's TMEM lane quarter.  Tile t
// uses TMEM accumulator t%2, meta buffer t%5, partial-sum buffer t%4; barriers indexed t%2:
//   wait mma_done(t) -> [role 0: write mu_out for the receivers finished in tile t-2] -> consume the tile's 4 groups
//   -> arrive on acc_free(t)
// mma_done(t) implies every consumer left tile t-2, so a warp runs at most one tile ahead of the slowest warp of its
// pipeline and the partial sums of tile t-2 are complete when role 0 reads them.  Every receiver of the range owns at
// least one group (edge_pack_kernel) and receivers are consecutive, so a receiver change is "drop row cur, start cur+1".
template <bool HAS_MU, int ROLE>
__device__ __forceinline__ void run_role(const PipeCtx& c, const Feed& fd, const float* __restrict__ mu_in,
                                         float* __restrict__ q, float* __restrict__ mu_out) {
    constexpr int NB = Tab<HAS_MU>::NODE_BYTES;
    constexpr uint32_t NB_SEL = (uint32_t)(NB / 4) | ((uint32_t)(NB / 4) << 24);   // dp2a.lo: h0 * b0 | dp2a.hi: h1 * b3
    static_assert(NB / 4 <= 255 && NB % 4 == 0, "gather stride as a dp2a byte");
    constexpr int NACC = ROLE == 0 ? 1 : 3;
    constexpr int T_OFF = ROLE == 0 ? 0 : ROLE == 1 ? SL * 4 : 2 * SL * 4;
    constexpr int P_ROW = ROLE == 1 ? 0 : 3;
    constexpr int NG = TE / GS;
    StBars* S = c.S;
    const int pipe = c.pipe, lane = c.lane, n_tiles = c.n_tiles;
    if (n_tiles <= 0 || c.ra >= c.rb) return;     // (a non-empty receiver range always has tiles)

    const uint32_t dcol0 = c.tmem + ((uint32_t)(c.quarter * 32) << 16) + TM_D + (uint32_t)(pipe * 2) * TE;
    // Gather base as ONE opaque 32-bit shared-window address: a gather is then PRMT (sender byte) + IMAD (j * NB + tb) + LDS.
    // (Left as a generic pointer, nvcc re-associated it into (j * NB + lane * 4) + table base in some builds of this kernel:
    // one more integer instruction per gather, 14 % more instructions per tile.)
    uint32_t tb = smem_u32(c.tab) + lane * 4 + T_OFF;
    asm("mov.b32 %0, %0;" : "+r"(tb));
    float* qp = q + ((int64_t)c.node0 * F + c.ch);
    float* mup = mu_out + ((int64_t)c.node0 * 3 * F + c.ch);

    int cur = c.ra;
    f32x2 acc[NACC];                          // (even slots, odd slots) partial sums
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = pk2(0.f, 0.f);
    float qbase = ROLE == 0 ? qp[(int64_t)cur * F] : 0.f;
    int n1 = 0, n2 = 0;                       // receivers finished in tiles t-1 / t-2 (role 0 combines them)

    int ms = 0, ps = 0;                       // t % NM_STAGE, t % 4
#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
        const uint32_t bi = (uint32_t)t & 1u, ph = ((uint32_t)t >> 1) & 1u;
#ifdef RLVD_TIMELINE
#define RLVD_STAMP(k) if (c.dbg != nullptr && t < 64) c.dbg[(t * 4 + ROLE) * 10 + (k)] = clock64();
#else
#define RLVD_STAMP(k)
#endif
        RLVD_STAMP(0)
        st_wait(&S->mma_done[pipe][bi], ph);
        tc_fence_after();
        RLVD_STAMP(1)
        // role 0: mu_out rows of the receivers finished in tile t-2.  The loads of the first one are issued here and
        // consumed after this tile's groups (their L2 / HBM latency hides behind the tile); the rare further ones follow.
        float cv[9];
        int64_t co = 0;
        const bool comb = RLVD_SPLIT && HAS_MU && ROLE == 0 && t >= 2 && n2 > 0;
        if (!RLVD_SPLIT && HAS_MU && ROLE == 0 && t >= 2 && n2 > 0) {
            if (lane == 0) st_wait(&S->acc_free[pipe][bi], ph ^ 1u);
            __syncwarp();
            combine_mu<HAS_MU>(c, (ps + 2) & 3, n2, mu_in, mu_out);
        }
        if (comb) {
            if (lane == 0) st_wait(&S->acc_free[pipe][bi], ph ^ 1u);   // already complete: acquires the partial sums of tile t-2
            __syncwarp();
            const int cb = (ps + 2) & 3;
            const int i = S->chg[pipe][cb][0];
            const float* P = c.Part + (cb * NG) * PART_SLOT + lane;
            co = (int64_t)(c.node0 + i) * 3 * F + c.ch;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                cv[r] = __ldcg(P + r * SL);
                cv[3 + r] = HAS_MU ? __ldcg(P + (3 + r) * SL) : 0.f;
                cv[6 + r] = HAS_MU ? __ldg(mu_in + co + r * F) : 0.f;
            }
        }
        RLVD_STAMP(2)
        const uint8_t* meta = c.MetaP + ms * META;
        float* part = c.Part + (ps * NG * PART_ROWS + P_ROW) * SL + lane;
        const uint32_t dcol = dcol0 + bi * TE;
        const uint32_t flagw = *reinterpret_cast<const uint32_t*>(meta + META_FLAG);
        int nchg = 0;
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) {
            uint32_t f[8];
            tmem_ld8(dcol + gg * GS, f);
            const uint4 jw = *reinterpret_cast<const uint4*>(meta + gg * GS * 2);
            float dr[3][8];
            if (ROLE == 1) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float4 da = *reinterpret_cast<const float4*>(meta + META_DIR + a * 128 + gg * 32);
                    const float4 db = *reinterpret_cast<const float4*>(meta + META_DIR + a * 128 + gg * 32 + 16);
                    dr[a][0] = da.x; dr[a][1] = da.y; dr[a][2] = da.z; dr[a][3] = da.w;
                    dr[a][4] = db.x; dr[a][5] = db.y; dr[a][6] = db.z; dr[a][7] = db.w;
                }
            }
            float t1[8], t2[8], t3[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t jp = u < 2 ? jw.x : u < 4 ? jw.y : u < 6 ? jw.z : jw.w;     // (4 j_even, 4 j_odd) as two u16
                const uint32_t row = (u & 1) ? dp2a_hi(jp, NB_SEL, tb) : dp2a_lo(jp, NB_SEL, tb);
                t1[u] = lds_f32<0>(row);
                if (ROLE == 2) {
                    t2[u] = lds_f32<SL * 4>(row);
                    t3[u] = lds_f32<2 * SL * 4>(row);
                }
            }
            if ((flagw >> (8 * gg)) & 0xffu) {                       // this group starts the next receiver
                if (ROLE == 0) {
                    qp[(int64_t)cur * F] = qbase + sum2(acc[0]);
                    qbase = qp[(int64_t)(cur + 1) * F];
                    if (HAS_MU && lane == 0) S->chg[pipe][ps][nchg] = cur;
                } else if (!HAS_MU) {                              // layer 0: mu = dmuR, role 1 owns the rows outright
#pragma unroll
                    for (int a = 0; a < NACC; ++a) mup[(int64_t)cur * 3 * F + a * F] = sum2(acc[a]);
                } else {
#pragma unroll
                    for (int a = 0; a < NACC; ++a) part[(nchg * PART_ROWS + a) * SL] = sum2(acc[a]);
                }
                ++nchg;
                ++cur;
#pragma unroll
                for (int a = 0; a < NACC; ++a) acc[a] = pk2(0.f, 0.f);
            }
            tmem_wait_ld8(f);
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const f32x2 w2 = pk2(__uint_as_float(f[u]), __uint_as_float(f[u + 1]));
                const f32x2 x2 = pk2(t1[u], t1[u + 1]);
                if (ROLE == 0) {
                    acc[0] = fma2(w2, x2, acc[0]);
                } else if (ROLE == 1) {
                    const f32x2 sR = mul2(w2, x2);
#pragma unroll
                    for (int a = 0; a < 3; ++a) acc[a] = fma2(sR, pk2(dr[a][u], dr[a][u + 1]), acc[a]);
                } else {
                    acc[0] = fma2(w2, x2, acc[0]);
                    acc[1] = fma2(w2, pk2(t2[u], t2[u + 1]), acc[1]);
                    acc[2] = fma2(w2, pk2(t3[u], t3[u + 1]), acc[2]);
                }
            }
        }
        if (comb) {
#pragma unroll
            for (int r = 0; r < 3; ++r) mu_out[co + r * F] = cv[r] + cv[3 + r] + cv[6 + r];
            if (n2 > 1) combine_mu<HAS_MU>(c, (ps + 2) & 3, n2, mu_in, mu_out, 1);
        }
        RLVD_STAMP(3)
        n2 = n1;
        n1 = nchg;
        ms = ms + 1 == NM_STAGE ? 0 : ms + 1;
        ps = (ps + 1) & 3;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->acc_free[pipe][bi]);
    }
    // ---- the receiver still in flight
    if (!HAS_MU) {                                                 // layer 0: every role owns its rows
        if (ROLE == 0) qp[(int64_t)cur * F] = qbase + sum2(acc[0]);
        else {
#pragma unroll
            for (int a = 0; a < NACC; ++a) mup[(int64_t)cur * 3 * F + a * F] = sum2(acc[a]);
        }
        return;
    }
    // once every role has left the last tile, drop its sums into slot 0 of the next partial buffer; role 0 then
    // writes the rows of the last two tiles and that one
    const uint32_t lb = (uint32_t)(n_tiles - 1) & 1u, lph = ((uint32_t)(n_tiles - 1) >> 1) & 1u;
    st_wait(&S->acc_free[pipe][lb], lph);
    if (ROLE == 0) {
        qp[(int64_t)cur * F] = qbase + sum2(acc[0]);
        if (lane == 0) S->chg[pipe][ps][0] = cur;
    } else {
        float* part = c.Part + (ps * NG * PART_ROWS + P_ROW) * SL + lane;
#pragma unroll
        for (int a = 0; a < NACC; ++a) part[a * SL] = sum2(acc[a]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S->acc_free[pipe][lb ^ 1u]);
    if (ROLE == 0) {
        st_wait(&S->acc_free[pipe][lb ^ 1u], ((uint32_t)n_tiles >> 1) & 1u);
        if (n_tiles >= 2) combine_mu<HAS_MU>(c, (ps + 2) & 3, n2, mu_in, mu_out);
        combine_mu<HAS_MU>(c, (ps + 3) & 3, n1, mu_in, mu_out);
        combine_mu<HAS_MU>(c, ps, 1, mu_in, mu_out);
    }
}

// Table staging loads bypass L1 (ld.global.cg): with 226 KB of the SM's 228 KB configured as shared memory the L1 has a
// handful of lines left, and allocating loads (ld.global.nc) queue on them.
#ifndef RLVD_STAGE_LD
#define RLVD_STAGE_LD __ldcg
#endif
template <bool HAS_MU>
__global__ void __launch_bounds__(ST_THREADS, 1) edge_stream_kernel(
    int n_max, const int* __restrict__ node_ptr, const uint32_t* __restrict__ Afmt /* [4 slices][4 rotations][128][32] */,
    float kappa, const int4* __restrict__ pipe_info, const uint8_t* __restrict__ tiles, float* __restrict__ scratch,
    const float* __restrict__ x, const float* __restrict__ mu_in, float* __restrict__ q, float* __restrict__ mu_out,
    long long* dbg, const int* __restrict__ gate, int gate_value, int n_struct, int pf_ahead, int pf_eighths, int n_split) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (gate != nullptr && *gate != gate_value) return;     // uniform over the grid: layer-0 fallback not needed
    constexpr int NB = Tab<HAS_MU>::NODE_BYTES;
    constexpr int NROLE = HAS_MU ? 3 : 2;           // consuming warps per pipeline (layer 0, mu == 0, has no dmumu term)
    uint8_t* TabS = smem;
    const uint32_t tab_bytes = ((uint32_t)n_max * NB + 127u) & ~127u;
    uint8_t* Bs = smem + tab_bytes;
    uint8_t* MetaS = Bs + NPIPE * NB_STAGE * TILE_B;
    StBars* S = reinterpret_cast<StBars*>(MetaS + NPIPE * NM_STAGE * META);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // grid = n_struct x n_split x NSLICE: CTA (g, split, slice) runs receiver ranges split * NPIPE .. + NPIPE of structure g
    // (n_split > 1 only for batches of less than ~two waves: single-replica stepping)
    const int g = blockIdx.x / (NSLICE * n_split), split = (blockIdx.x / NSLICE) % n_split, slice = blockIdx.x % NSLICE;
    const int node0 = __ldg(node_ptr + g), n = __ldg(node_ptr + g + 1) - node0;
#ifdef RLVD_TIMELINE   // CTA-level stamps of block 0 and of a mid-grid block: entry | feeders released | table staged | roles done | exit
    long long* cst = dbg == nullptr ? nullptr : blockIdx.x == 0 ? dbg + 64 * 4 * 10 : blockIdx.x == 6001 ? dbg + 64 * 4 * 10 + 8 : nullptr;
#define RLVD_CTA_STAMP(k, w) if (cst != nullptr && warp == (w) && lane == 0) cst[k] = clock64();
#else
#define RLVD_CTA_STAMP(k, w)
#endif
    RLVD_CTA_STAMP(0, 0)
    uint32_t smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid >= MAX_SMID) {
        if (tid == 0) printf("rlvd: edge_stream: smid %u exceeds the scratch table\n", smid);
        __trap();
    }

    // Prologue.  The feeder warps only need the weight operand in TMEM, the consuming warps only need the node table.
    //   feeder warps   : the feeders of pipelines 0..3 sit in TMEM lane quarters 3, 2, 1, 0 (pipe_rot) and each owns its quarter
    //                    of the operand.  They request rotations 0-1 from L2 at once; warp 3 allocates TMEM while pipeline 1's
    //                    feeder initialises the mbarriers; named barrier 1 hands the base address over; rotations 0-1 are
    //                    stored and rotations 2-3 requested.  Only then does ONE block barrier let the consuming warps start:
    //                    measured, a feeder's mbarrier.init / L2 loads issued under the 576-thread staging flood wait 2-5 k
    //                    cycles in the LSU queue.  Second store, named barrier 1 again, and the first tiles' copies and MMAs
    //                    run under the table staging (feeders released at ~4.8 k cycles, first accumulator waiting for the
    //                    consumers when the table is ready).
    //   consuming warps: stage the node table (nothing else) and meet at named barrier 2.
    // The waves of this kernel run in lockstep (equal work per CTA), so all 148 SMs stage at once and the 196 KB per CTA come
    // at the HBM rate (10-11 k cycles; 4.5 k when a single SM stages, 4.0 k from L2: scripts/probes/stage_rate.cu) unless the
    // previous wave asked L2 for them (Feed::pf below: 7 k).
    // (Round 2, first form: warps 0-3 loaded the operand before staging their share of the table: feeders released at
    // 9-12 k cycles, table staged at 15-21 k, CTA 173 k cycles; now 4.8 k / 10 k / 164 k.  profiles/timeline_stream_r02.txt)
    const int pipe_w = warp >> 2;
    const int role_w = ((warp & 3) + pipe_rot(pipe_w)) & 3;
    const bool feeds = role_w == 3 || (!HAS_MU && role_w == 2);
    constexpr int NCW = HAS_MU ? 3 : 2;                             // consuming warps per pipeline
    constexpr int N_FEED_WARPS = NPIPE * (4 - NCW);
    constexpr int FEED_THREADS = N_FEED_WARPS * 32;
    constexpr int STAGE_THREADS = NPIPE * NCW * 32;
    static_assert(NPIPE >= 4, "the feeders of pipelines 0..3 cover the four TMEM lane quarters");

    if (feeds) {
        // the operand loads (L2) are issued first: their latency hides behind barrier initialisation and TMEM allocation
        const bool loads_a = role_w == 3 && pipe_w < 4;
        const int quarter = warp & 3;
        const uint4* wrow = reinterpret_cast<const uint4*>(Afmt + ((size_t)(slice * NROT) * 128 + quarter * 32 + lane) * 32);
        uint4 w[16];
        if (loads_a) {
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) w[r2 * 8 + c8] = __ldg(wrow + (size_t)r2 * 128 * 8 + c8);
        }
        // warp 3 allocates TMEM, pipeline 1's feeder initialises the mbarriers meanwhile; named barrier 1 hands the base
        // address to the feeders
        if (role_w == 3 && pipe_w == 1 && lane == 0) {
            for (int p = 0; p < NPIPE; ++p)
                for (int k = 0; k < 2; ++k) {
                    mbar_init(&S->b_full[p][k], 1);
                    mbar_init(&S->mma_done[p][k], 1);
                    mbar_init(&S->acc_free[p][k], NROLE);
                }
            fence_mbar_init();
        }
        if (warp == 3) {
            tmem_alloc(&S->tmem_base, TM_COLS);
            tc_fence_before();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(FEED_THREADS) : "memory");
        tc_fence_after();
        RLVD_CTA_STAMP(5, 3)
        const uint32_t tmem_f = S->tmem_base;
        // ---- weight operand -> TMEM in its four row rotations (lane = filter row; 64 fp16 = 32 columns each): rotations
        // 0-1 stored, rotations 2-3 requested, all before the block barrier releases the staging loads
        if (loads_a) {
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
                const uint32_t ta = tmem_f + ((uint32_t)(quarter * 32) << 16) + TM_A + r2 * 32;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const uint4 w0 = w[r2 * 8 + 2 * c8], w1 = w[r2 * 8 + 2 * c8 + 1];
                    const uint32_t v[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    tmem_st8(ta + c8 * 8, v);
                }
            }
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) w[r2 * 8 + c8] = __ldg(wrow + (size_t)(2 + r2) * 128 * 8 + c8);
        }
        __syncthreads();          // barriers and TMEM base published to the whole CTA before the staging loads flood the LSU
        tc_fence_after();
        RLVD_CTA_STAMP(6, 3)
        if (loads_a) {
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
                const uint32_t ta = tmem_f + ((uint32_t)(quarter * 32) << 16) + TM_A + (2 + r2) * 32;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const uint4 w0 = w[r2 * 8 + 2 * c8], w1 = w[r2 * 8 + 2 * c8 + 1];
                    const uint32_t v[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    tmem_st8(ta + c8 * 8, v);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(FEED_THREADS) : "memory");
        tc_fence_after();
        RLVD_CTA_STAMP(1, 3)
    } else {
        __syncthreads();
        tc_fence_after();
        // ---- stage this slice of the structure's node table (consuming warps only); two items' loads in flight
        const int sid = (pipe_w * NCW + role_w) * 32 + lane;
        const int n_items = n * 8;
#pragma unroll 1
        for (int id0 = sid; id0 < n_items; id0 += 2 * STAGE_THREADS) {
            float4 xq[2], xR[2], xm[2], m[2][3];
            bool on[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int id = id0 + h * STAGE_THREADS;
                on[h] = id < n_items;
                if (on[h]) {
                    const int node = id >> 3, c4 = id & 7;
                    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)(node0 + node) * 3 * F + slice * SL) + c4;
                    xq[h] = RLVD_STAGE_LD(xr);
                    xR[h] = RLVD_STAGE_LD(xr + F / 4);
                    if (HAS_MU) {
                        xm[h] = RLVD_STAGE_LD(xr + 2 * F / 4);
                        const float4* mr = reinterpret_cast<const float4*>(mu_in + (size_t)(node0 + node) * 3 * F + slice * SL) + c4;
#pragma unroll
                        for (int a = 0; a < 3; ++a) m[h][a] = RLVD_STAGE_LD(mr + a * (F / 4));
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!on[h]) continue;
                const int id = id0 + h * STAGE_THREADS;
                const int node = id >> 3, c4 = id & 7;
                uint8_t* row = TabS + node * NB + c4 * 16;
                float4 a = xq[h], b = xR[h];
                a.x *= kappa; a.y *= kappa; a.z *= kappa; a.w *= kappa;
                b.x *= kappa; b.y *= kappa; b.z *= kappa; b.w *= kappa;
                *reinterpret_cast<float4*>(row) = a;
                *reinterpret_cast<float4*>(row + SL * 4) = b;
                if (HAS_MU) {
                    float4 c = xm[h];
                    c.x *= kappa; c.y *= kappa; c.z *= kappa; c.w *= kappa;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        *reinterpret_cast<float4*>(row + (2 + k) * SL * 4) =
                            make_float4(c.x * m[h][k].x, c.y * m[h][k].y, c.z * m[h][k].z, c.w * m[h][k].w);
                }
            }
        }
        asm volatile("bar.sync 2, %0;" ::"n"(STAGE_THREADS) : "memory");
        tc_fence_after();
        RLVD_CTA_STAMP(2, 0)
    }
    const uint32_t tmem = S->tmem_base;

    PipeCtx c;
    c.S = S;
    c.pipe = pipe_w;
    c.lane = lane;
    c.quarter = warp & 3;
    c.tmem = tmem;
    c.tab = TabS;
    c.MetaP = MetaS + c.pipe * NM_STAGE * META;
    c.Part = scratch + (size_t)smid * PART_SM + (size_t)c.pipe * PART_PIPE;
    const int4 pi = __ldg(pipe_info + ((size_t)g * n_split + split) * NPIPE + c.pipe);
    c.n_tiles = pi.y;
    c.ra = pi.z;
    c.rb = pi.w;
    c.node0 = node0;
    c.ch = slice * SL + lane;
    c.pf_eighths = pf_eighths;
    c.dbg = (blockIdx.x == 0 && c.pipe == 0 && lane == 0) ? dbg : nullptr;
    const int role = role_w;                          // the pipeline's rotation puts this role's filter rows in this warp's quarter
    Feed fd;
    fd.src = tiles + (size_t)pi.x * TILE_BYTES;
    fd.Bp = Bs + c.pipe * NB_STAGE * TILE_B;
    fd.bdesc0 = umma_desc(smem_u32(fd.Bp), 512, 128);
    // Table-staging prefetch: the CTAs that take over the SMs when this wave retires (structure g + pf_ahead) will pull
    // their x / mu rows from HBM with every consuming warp stalled on them (~15 k cycles of a ~175 k-cycle CTA).  Half way
    // through its tiles each of the 16 (slice, pipeline < 4) feeders asks L2 for one sixteenth of those rows.
    fd.pf = nullptr;
    fd.pf_bytes = 0;
    if (HAS_MU && pf_ahead > 0 && c.pipe < 4 && g + pf_ahead < n_struct) {
        const int na = __ldg(node_ptr + g + pf_ahead), nb = __ldg(node_ptr + g + pf_ahead + 1);
        const int chunk = c.pipe * NSLICE + slice;                       // 0..15: x chunks 0..7, mu chunks 8..15
        const uint32_t total = (uint32_t)(nb - na) * 3u * F * 4u;        // bytes of the structure's rows in either array
        const uint32_t per = ((total + 7u) / 8u + 127u) & ~127u;
        const uint32_t off = (uint32_t)(chunk & 7) * per;
        if (off < total) {
            const uint8_t* base = reinterpret_cast<const uint8_t*>(chunk < 8 ? x : mu_in) + (size_t)na * 3 * F * 4;
            fd.pf = base + off;
            fd.pf_bytes = total - off < per ? total - off : per;
        }
    }
    if (role == 0) run_role<HAS_MU, 0>(c, fd, mu_in, q, mu_out);
    else if (role == 1) run_role<HAS_MU, 1>(c, fd, mu_in, q, mu_out);
    else if (role == 2) { if (HAS_MU) run_role<HAS_MU, 2>(c, fd, mu_in, q, mu_out); else run_feeder<2>(c, fd); }
    else run_feeder<HAS_MU ? 0 : 1>(c, fd);
    RLVD_CTA_STAMP(3, 0)

    tc_fence_before();
    __syncthreads();
    RLVD_CTA_STAMP(4, 0)
    if (warp == 3) {                                  // the allocating warp
        tc_fence_after();
        tmem_dealloc(tmem, TM_COLS);
    }
}

template <bool HAS_MU>
size_t stream_smem_bytes(int n_max) {
    const size_t tab = ((size_t)n_max * Tab<HAS_MU>::NODE_BYTES + 127) & ~(size_t)127;
    return tab + (size_t)NPIPE * NB_STAGE * TILE_B + (size_t)NPIPE * NM_STAGE * META + sizeof(StBars);
}

}  // namespace

// Host: per layer, per 32-channel slice, the 128-row weight operand in the TMEM image the kernel stores verbatim:
// row = 32 u32 = 64 fp16 = a_hi[21] | a_hi[21] | a_lo[21] | 0 with a = 2^sA * {filter_net weight (k < 20), bias}.
// Row quarters of rotation r: the filters of role (quarter + r) % 4 (roles: dq | dmuR | dmumu | dmumu).
// kappa[l] = 2^-(sA_l + 10) undoes the operand scaling (applied to the node table).
void edge_stream_format(const float* W, const float* b, std::vector<uint8_t>& out, float* kappa /*[NLAYER]*/) {
    out.assign((size_t)NLAYER * NSLICE * A_SLICE_WORDS * 4, 0);
    uint32_t* o = reinterpret_cast<uint32_t*>(out.data());
    for (int l = 0; l < NLAYER; ++l) {
        float mx = 0.f;
        for (int r = l * 3 * F; r < (l + 1) * 3 * F; ++r) {
            for (int k = 0; k < NRBF; ++k) mx = std::fmax(mx, std::fabs(W[(size_t)r * NRBF + k]));
            mx = std::fmax(mx, std::fabs(b[r]));
        }
        int sA = 0;
        if (mx > 0.f) sA = (int)std::floor(std::log2(16384.0 / (double)mx));
        sA = sA < -20 ? -20 : (sA > 24 ? 24 : sA);
        const float ascale = std::ldexp(1.0f, sA);
        kappa[l] = std::ldexp(1.0f, -(sA + B_SCALE_LOG2));
        for (int sl = 0; sl < NSLICE; ++sl)
          for (int rot = 0; rot < NROT; ++rot)
            for (int quarter = 0; quarter < 4; ++quarter) {
                const int role = (quarter + rot) & 3;          // a pipeline with rotation rot: warp quarter -> role
                const int part = role < 3 ? role : 2;
                for (int chn = 0; chn < SL; ++chn) {
                    const int row = l * 3 * F + part * F + sl * SL + chn;
                    __half kk[KH];
                    for (int k = 0; k < KH; ++k) kk[k] = __float2half_rn(0.f);
                    for (int k = 0; k < NK; ++k) {
                        const float w = (k < NRBF ? W[(size_t)row * NRBF + k] : b[row]) * ascale;
                        const __half h = __float2half_rn(w);
                        const __half lo = __float2half_rn(w - __half2float(h));
                        kk[k] = h;
                        kk[NK + k] = h;
                        kk[2 * NK + k] = lo;
                    }
                    uint32_t* dst = o + ((((size_t)l * NSLICE + sl) * NROT + rot) * 128 + quarter * SL + chn) * 32;
                    for (int c = 0; c < 32; ++c) {
                        uint16_t h0, h1;
                        memcpy(&h0, &kk[2 * c], 2);
                        memcpy(&h1, &kk[2 * c + 1], 2);
                        dst[c] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                    }
                }
            }
    }
}

bool edge_stream_supported(int max_struct_nodes) { return max_struct_nodes > 0 && max_struct_nodes <= MAXN; }

size_t edge_stream_tile_capacity(int n_struct, int n_nodes, int64_t n_edges) {
    return (size_t)((n_edges / GS + n_nodes) / 4 + (int64_t)MAXR * n_struct + 16);
}

int launch_edge_pack(rlvd_engine* e, cudaStream_t s, int n_struct, int M, int64_t n_edges, const int* node_ptr,
                     const int* row_ptr, const int* col, const int8_t* shift, const float* pos, const float* cell,
                     const int* Z, float* l0A) {
    if (n_struct <= 0 || M <= 0) return 0;
    const size_t cap = edge_stream_tile_capacity(n_struct, M, n_edges);
    RLVD_CHECK(cap < (size_t)0x7fffffff, "edge stream: %zu tiles overflow int32; split the batch", cap);
    if (e->tiles.ensure(cap * TILE_BYTES) || e->pipe_info.ensure((size_t)n_struct * MAXR * sizeof(int4)) ||
        e->tile_ctr.ensure(16))
        return 1;
    RLVD_CUDA(cudaMemsetAsync(e->tile_ctr.p, 0, 16, s));   // tile counter | capacity error | species outside the model's list
    Span sp(e, s, FAM_PACK);
    // few structures (single-replica stepping): split each structure's receivers over several CTAs
    const int n_splits = n_struct >= 296 ? 1 : min(16, (592 + n_struct - 1) / n_struct);
    // ... and over several CTAs of the stream kernel: with S CTAs per (structure, slice) a CTA costs ~10 k cycles of prologue plus
    // ~154 k / S of tiles (profiles/timeline_stream_r02.txt) and the launch runs ceil(4 S n_struct / SMs) waves of them
    int best = 1;
    double best_cost = 0.0;
    for (int S = 1; S <= MAX_SPLIT; ++S) {
        const int waves = (NSLICE * S * n_struct + e->sm_count - 1) / e->sm_count;
        const double cost = waves * (10.0 + 154.0 / S);
        if (S == 1 || cost < 0.95 * best_cost) { best = S; best_cost = cost; }
    }
    e->stream_split = best;
    const int NR = NPIPE * best;
    for (int mode = n_splits > 1 ? 1 : 0; mode <= (n_splits > 1 ? 2 : 0); ++mode)
        edge_pack_kernel<<<dim3(n_struct, mode == 2 ? n_splits : 1), 256, 0, s>>>(
            n_struct, node_ptr, row_ptr, col, shift, pos, cell, e->w.freqs, e->w.feps, e->w.cutoff, e->tile_ctr.as<int>(), (int)cap,
            e->tiles.as<uint8_t>(), e->pipe_info.as<int4>(), e->tile_ctr.as<int>() + 1, e->status.as<int>(), Z, e->w.z2s, l0A, mode, n_splits, NR);
    sp.end(n_splits > 1 ? 2 : 1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_edge_stream(rlvd_engine* e, cudaStream_t s, int layer, int n_struct, int max_struct_nodes, const int* node_ptr,
                       const float* x, const float* mu_in, float* q, float* mu_out, const int* gate, int gate_value) {
    if (n_struct <= 0) return 0;
    if (!e->attr_stream) {                 // per engine = per device: the attribute is a per-device property
        RLVD_CUDA(cudaFuncSetAttribute(edge_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)stream_smem_bytes<true>(MAXN)));
        RLVD_CUDA(cudaFuncSetAttribute(edge_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)stream_smem_bytes<true>(MAXN)));
        e->attr_stream = true;
    }
    if (e->stream_scratch.ensure((size_t)MAX_SMID * PART_SM * sizeof(float))) return 1;
    float* scr = e->stream_scratch.as<float>();
    const uint32_t* af = reinterpret_cast<const uint32_t*>(e->w.filt_E) + (size_t)layer * NSLICE * A_SLICE_WORDS;
    const float kappa = e->w.filt_kappa[layer];
    // both variants request the HAS_MU footprint: one CTA per SM (each CTA owns all 512 TMEM columns)
    const size_t smem = stream_smem_bytes<true>(max_struct_nodes) < 120 * 1024 ? 120 * 1024 : stream_smem_bytes<true>(max_struct_nodes);
    // structures between a CTA and the one that follows it on the same SM (grids of more than one wave only)
    const int n_split = e->stream_split;              // chosen by launch_edge_pack for this batch
    const int pf_ahead = e->stream_prefetch > 0 && n_split == 1 && n_struct * NSLICE > e->sm_count ? (e->sm_count + NSLICE - 1) / NSLICE : 0;
    long long* dbg = nullptr;
#ifdef RLVD_TIMELINE   // in-kernel timeline of block 0 / pipeline 0 (build with RLVD_EXTRA_NVCC_FLAGS=-DRLVD_TIMELINE)
    static long long* dbg_buf = nullptr;
    static int dbg_calls = 0;
    if (dbg_buf == nullptr) { cudaMalloc(&dbg_buf, (64 * 4 * 10 + 16) * 8); cudaMemset(dbg_buf, 0, (64 * 4 * 10 + 16) * 8); }
    dbg = dbg_buf;
#endif
    Span sp(e, s, gate != nullptr ? -1 : FAM_EDGE);     // a gated launch is the (normally skipped) layer-0 fallback
    if (mu_in == nullptr)
        edge_stream_kernel<false><<<n_struct * NSLICE * n_split, ST_THREADS, smem, s>>>(
            max_struct_nodes, node_ptr, af, kappa, e->pipe_info.as<int4>(), e->tiles.as<uint8_t>(), scr, x, nullptr, q, mu_out, nullptr, gate, gate_value, n_struct, 0, 0, n_split);
    else
        edge_stream_kernel<true><<<n_struct * NSLICE * n_split, ST_THREADS, smem, s>>>(
            max_struct_nodes, node_ptr, af, kappa, e->pipe_info.as<int4>(), e->tiles.as<uint8_t>(), scr, x, mu_in, q, mu_out, dbg, gate, gate_value, n_struct, pf_ahead, e->stream_prefetch, n_split);
#ifdef RLVD_TIMELINE
    if (mu_in != nullptr && ++dbg_calls == 20) {
        cudaStreamSynchronize(s);
        static long long h[64 * 4 * 10 + 16];
        cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        for (int b = 0; b < 2; ++b) {
            const long long* w = h + 64 * 4 * 10 + 8 * b;
            printf("cta %s | TMEM base at the feeders +%lld | block barrier +%lld | feeders released +%lld | table staged +%lld | roles done +%lld | exit +%lld cycles\n",
                   b == 0 ? "0" : "6001", w[5] - w[0], w[6] - w[0], w[1] - w[0], w[2] - w[0], w[3] - w[0], w[4] - w[0]);
        }
        for (int t = 0; t < 40; ++t) {
            printf("tile %2d |", t);
            for (int r = 0; r < 4; ++r) {
                const long long* w = h + (t * 4 + r) * 10;
                printf(" R%d top %7lld wait +%5lld combine +%5lld consume +%5lld |", r, w[0] - t0, w[1] - w[0], w[2] - w[1], w[3] - w[2]);
            }
            { const long long* w = h + (t * 4 + 3) * 10; printf(" duty: accfree +%5lld bfull +%5lld mma +%5lld mmadone +%5lld copy +%5lld", w[4] - w[3], w[5] - w[4], w[6] - w[5], w[7] - w[6], w[8] - w[7]); }
            printf("\n");
        }
    }
#endif
    sp.end(gate != nullptr ? 0 : 1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
