// K2 + K3a + K3c + K4 fused for structures of at most 256 atoms — the BASELINE workload (255/256-atom replicas).
//
// Same contract as edge.cu / tc_edge.cu.  What bounds those two is the gather: every edge pulls x[j] and mu[j]
// (3 KB) from L2, 126 GB per layer at 128 replicas.  Here the gather never leaves the SM: a CTA owns ONE structure
// and ONE 32-channel slice and keeps that slice of the structure's whole node table ([n][x_q, x_R, x_m, mu_x, mu_y,
// mu_z][32] fp32 = 192 KB) in shared memory, read from HBM/L2 exactly once.  Four CTAs (slices) cover a
// structure; nothing is exchanged between them.
//
// Filters run on the tensor cores, transposed so they land in TMEM as lane = filter row, column = edge:
//      D[128, 32 edges] = Wslice[128 rows, 24] · Phi[32, 24]ᵀ          tf32 hi/lo x3, A operand resident in TMEM
// Rows 0-31 are the dq filters of the slice, 32-63 dmuR, 64-95 and 96-127 both dmumu, so the TMEM lane quarter a
// warp may read selects its share of the message:
//      warp 0 (q) :  dq          += W · x_q[j]
//      warp 1 (R) :  dmu_k       += (W · x_R[j]) · dir_k
//      warp 2 (m) :  dmu_{x,y}   += (W · x_m[j]) · mu_{x,y}[j]
//      warp 3 (m') : dmu_z       += (W · x_m[j]) · mu_z[j]            (partials are summed at the receiver flush)
// A CTA runs three such 4-warp pipelines on disjoint receiver ranges, fed by three producer warps.  The per-edge
// embedding (20 x phi_k·fcut, fcut, dir) is computed ONCE per representation call by edge_geom_kernel into a
// 128-byte record per edge, shared by the 3 layers x 4 slices; a producer streams the records of the next tile
// (prefetched one tile ahead), splits them into the tf32 hi/lo operand tile and issues the 9 MMAs.  Tiles are built
// from receiver-aligned groups of 8 edge slots (short rows padded with zero filters), so consumers test for a
// receiver change once per group, issue all shared-memory gathers of a group before the first dependent FMA, and
// address the table as [per-thread base + per-edge byte offset + immediate].  Sums run in CSR order: no atomics,
// bitwise reproducible.
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace rlvd {

namespace {

constexpr int SL = 32;                  // channels per slice
constexpr int NSLICE = F / SL;          // 4
constexpr int NE = 32;                  // edge slots per tile = MMA N
constexpr int GS = 8;                   // slots per receiver-aligned group
constexpr int KP = 24;
constexpr int NPIPE = 3;
constexpr int ES_THREADS = NPIPE * 128 + NPIPE * 32;   // 12 consumer warps + 3 producer warps
constexpr int MAXN = 256;
constexpr int NODE_BYTES = 6 * SL * 4;  // table row: x_q, x_R, x_m, mu_x, mu_y, mu_z of the slice
constexpr int A_ROW = 2 * KP;           // floats per filter row in the host-formatted slice: hi[24] | lo[24]
constexpr int A_SLICE = 128 * A_ROW * 4;// bytes per (layer, slice)
constexpr int B_HALF = 6 * 512;         // 6 core columns x 4 row groups x 128 B
constexpr int B_TILE = 2 * B_HALF;      // hi | lo = 6144
constexpr int META = 768;               // joff[32] (128 B) | dir[32] float4 (512 B) | recv[4] (16 B) | pad
constexpr int TM_D = 0;                 // TMEM columns: accumulators [pipe][buf][32]
constexpr int TM_A = NPIPE * 2 * NE;    // 192: filter operand A, hi[24] | lo[24]
constexpr int TM_COLS = 256;
constexpr float PI_F = 3.14159265358979323846f;
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NE >> 3) << 17) | (8u << 24);

struct EsBars {
    uint64_t mma_done[NPIPE][2], acc_free[NPIPE][2];
    uint32_t tmem_base;
};

// D[tmem] (+)= A[tmem] · B[smem]ᵀ : the filter rows live in TMEM (lane = row), so the MMA only reads the small
// per-tile operand from shared memory — with N = 32 an A operand in shared memory would cost 4 KB of smem reads per
// instruction, three times the gather traffic this kernel exists to keep on-chip.
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8s(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void trio_barrier(int pipe) { asm volatile("bar.sync %0, 96;" ::"r"(pipe + 1) : "memory"); }

// smem carve-up (dynamic): table | B[NPIPE] | meta[NPIPE][2] | exch[NPIPE][2][3][32] | rowp[MAXN+1] | bars
struct Layout {
    uint32_t tab, b, meta, exch, rowp, bars, total;
};
__host__ __device__ inline Layout make_layout(int n_max) {
    Layout L;
    uint32_t o = 0;
    L.tab = o; o += (uint32_t)n_max * NODE_BYTES;
    o = (o + 127u) & ~127u;
    L.b = o; o += NPIPE * B_TILE;
    L.meta = o; o += NPIPE * 2 * META;
    L.exch = o; o += NPIPE * 2 * 3 * SL * 4;
    L.rowp = o; o += (MAXN + 1 + 3) / 4 * 16;
    L.bars = o; o += 128;
    L.total = o;
    return L;
}

// Per-edge embedding record (8 x float4 = 128 B), computed once per representation call:
//   c0..c4 = phi_k * fcut (k = 0..19)     c5 = { fcut, dir.x, dir.y, dir.z }     c6 = { j - node0, i - node0, 0, 0 } (ints)
// sin(freqs[k] d): one accurate sincos of theta = freqs[0]*d (fp32 product error folded in), an angle-addition tree
// of depth <= 5 for k*theta, and the deviation freqs[k] - (k+1)*freqs[0] folded in to first order (max abs error
// 1.4e-6, below the 1.9e-6 argument-rounding error of evaluating sin(fl(freqs[k]*d)) directly in fp32).
// One warp per receiver row, lanes over its edges.
__global__ void __launch_bounds__(256) edge_geom_kernel(int M, const int* __restrict__ node_ptr,
                                                        const int* __restrict__ node_struct,
                                                        const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                        const int8_t* __restrict__ shift, const float* __restrict__ pos,
                                                        const float* __restrict__ cell, const float* __restrict__ freqs,
                                                        const float* __restrict__ feps, float cutoff,
                                                        float4* __restrict__ rec) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= M) return;
    const int gs = __ldg(node_struct + i), node0 = __ldg(node_ptr + gs);
    const float xi = __ldg(pos + 3 * (int64_t)i), yi = __ldg(pos + 3 * (int64_t)i + 1), zi = __ldg(pos + 3 * (int64_t)i + 2);
    float C[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) C[t] = __ldg(cell + 9 * gs + t);
    const float f1 = __ldg(freqs), pi_over_rc = PI_F / cutoff;
    const int e_hi = __ldg(row_ptr + i + 1);
    for (int e = __ldg(row_ptr + i) + lane; e < e_hi; e += 32) {
        const int jg = __ldg(col + e);
        const float S0 = (float)shift[3 * (int64_t)e], S1 = (float)shift[3 * (int64_t)e + 1], S2 = (float)shift[3 * (int64_t)e + 2];
        float rx = __ldg(pos + 3 * (int64_t)jg) - xi, ry = __ldg(pos + 3 * (int64_t)jg + 1) - yi,
              rz = __ldg(pos + 3 * (int64_t)jg + 2) - zi;
        rx += S0 * C[0] + S1 * C[3] + S2 * C[6];
        ry += S0 * C[1] + S1 * C[4] + S2 * C[7];
        rz += S0 * C[2] + S1 * C[5] + S2 * C[8];
        const float d = sqrtf(rx * rx + ry * ry + rz * rz);
        const float inv_d = 1.0f / d;
        const float fc = d < cutoff ? 0.5f * (cosf(d * pi_over_rc) + 1.0f) : 0.f;
        const float scale = inv_d * fc;
        const float pr = f1 * d, perr = fmaf(f1, d, -pr);
        float sn, cs;
        sincosf(pr, &sn, &cs);
        float s[NRBF + 1], c[NRBF + 1], v[NRBF];
        s[1] = fmaf(perr, cs, sn);
        c[1] = fmaf(-perr, sn, cs);
#define RLVD_ADD(a, b) { s[(a) + (b)] = fmaf(s[a], c[b], c[a] * s[b]); c[(a) + (b)] = fmaf(c[a], c[b], -(s[a] * s[b])); }
        RLVD_ADD(1, 1) RLVD_ADD(2, 1) RLVD_ADD(2, 2) RLVD_ADD(4, 1) RLVD_ADD(4, 2) RLVD_ADD(4, 3) RLVD_ADD(4, 4)
        RLVD_ADD(8, 1) RLVD_ADD(8, 2) RLVD_ADD(8, 3) RLVD_ADD(8, 4) RLVD_ADD(8, 5) RLVD_ADD(8, 6) RLVD_ADD(8, 7)
        RLVD_ADD(8, 8) RLVD_ADD(16, 1) RLVD_ADD(16, 2) RLVD_ADD(16, 3) RLVD_ADD(16, 4)
#undef RLVD_ADD
#pragma unroll
        for (int k = 0; k < NRBF; ++k) v[k] = fmaf(__ldg(feps + k) * d, c[k + 1], s[k + 1]) * scale;
        float4* o = rec + 8 * (int64_t)e;
#pragma unroll
        for (int t = 0; t < 5; ++t) o[t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
        o[5] = make_float4(fc, rx * inv_d, ry * inv_d, rz * inv_d);
        o[6] = make_float4(__int_as_float(jg - node0), __int_as_float(i - node0), 0.f, 0.f);
    }
}

struct ConsState {
    int i_cur, n_flush;
    float a0, a1, a2, q_cur;
};

// Receiver change: write the finished receiver (and any edgeless receivers before `next`).  Rare (once per ~7
// groups), kept out of line so the hot loop stays small enough for the instruction cache.
template <bool HAS_MU, int ROLE>
__device__ __noinline__ void flush_rows(ConsState& st, int next, int pipe, const uint8_t* tb, float* Ex, int rb, int node0,
                                        int ch, int lane, float* __restrict__ q, float* __restrict__ mu_out) {
    for (int i = st.i_cur; i < next; ++i, ++st.n_flush) {
        const int64_t gi = node0 + i;
        if (ROLE == 0) {
            q[gi * F + ch] = st.q_cur + st.a0;
            if (i + 1 < rb) st.q_cur = q[(gi + 1) * F + ch];      // next receiver's q: loaded now, needed one row later
        } else if (!HAS_MU) {
            mu_out[gi * 3 * F + ch] = st.a0;
            mu_out[gi * 3 * F + F + ch] = st.a1;
            mu_out[gi * 3 * F + 2 * F + ch] = st.a2;
        } else {
            float* ex = Ex + (st.n_flush & 1) * 3 * SL;
            if (ROLE == 2) { ex[lane] = st.a0; ex[SL + lane] = st.a1; }
            if (ROLE == 3) ex[2 * SL + lane] = st.a0;
            trio_barrier(pipe);
            if (ROLE == 1) {
                const float* mi = reinterpret_cast<const float*>(tb + i * NODE_BYTES + 384);
                mu_out[gi * 3 * F + ch] = mi[0] + st.a0 + ex[lane];
                mu_out[gi * 3 * F + F + ch] = mi[SL] + st.a1 + ex[SL + lane];
                mu_out[gi * 3 * F + 2 * F + ch] = mi[2 * SL] + st.a2 + ex[2 * SL + lane];
            }
        }
        st.a0 = st.a1 = st.a2 = 0.f;
    }
    st.i_cur = next;
}

// One consumer warp: ROLE is its TMEM lane quarter = its share of the message (see the file header).
template <bool HAS_MU, int ROLE>
__device__ __forceinline__ void consume_tiles(const EsBars* S, int pipe, int n_tiles, uint32_t tmem, const uint8_t* MetaP,
                                              const uint8_t* tb, float* Ex, int ra, int rb, int node0, int ch, int lane,
                                              float* __restrict__ q, float* __restrict__ mu_out, long long* dbg) {
    const uint32_t taddr0 = tmem + ((uint32_t)(ROLE * 32) << 16) + TM_D + (uint32_t)(pipe * 2) * NE;
    ConsState st;                          // ROLE 0: a0 = dq; 1: dmu_xyz (dir part); 2: dmu_xy (mu part); 3: a0 = dmu_z (mu part)
    st.i_cur = ra; st.n_flush = 0; st.a0 = st.a1 = st.a2 = 0.f; st.q_cur = 0.f;
    if (ROLE == 0 && ra < rb) st.q_cur = q[(int64_t)(node0 + ra) * F + ch];
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
#ifdef RLVD_TIMELINE
        const bool DBG = dbg != nullptr && blockIdx.x == 0 && pipe == 0 && lane == 0 && t < 48;
        if (DBG) dbg[(t * 5 + ROLE + 1) * 4 + 0] = clock64();
#endif
        mbar_wait(const_cast<uint64_t*>(&S->mma_done[pipe][buf]), (t >> 1) & 1);
#ifdef RLVD_TIMELINE
        if (DBG) dbg[(t * 5 + ROLE + 1) * 4 + 1] = clock64();
#endif
        tc_fence_after();
        const uint8_t* meta = MetaP + buf * META;
#pragma unroll 1
        for (int gg = 0; gg < NE / GS; ++gg) {
            uint32_t w[8];
            tmem_ld8s(taddr0 + buf * NE + gg * GS, w);
            const int recv = *reinterpret_cast<const int*>(meta + 640 + gg * 4);
            const int4 ja = *reinterpret_cast<const int4*>(meta + gg * 32), jb = *reinterpret_cast<const int4*>(meta + gg * 32 + 16);
            const int jo[8] = {ja.x, ja.y, ja.z, ja.w, jb.x, jb.y, jb.z, jb.w};
            float xv[8], g0[8], g1[8], g2[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint8_t* row = tb + jo[u];
                if (ROLE == 0) {
                    xv[u] = *reinterpret_cast<const float*>(row);
                } else if (ROLE == 1) {
                    xv[u] = *reinterpret_cast<const float*>(row + 128);
                    const float4 dir = *reinterpret_cast<const float4*>(meta + 128 + (gg * GS + u) * 16);
                    g0[u] = dir.x; g1[u] = dir.y; g2[u] = dir.z;
                } else if (ROLE == 2) {
                    xv[u] = *reinterpret_cast<const float*>(row + 256);
                    g0[u] = *reinterpret_cast<const float*>(row + 384);
                    g1[u] = *reinterpret_cast<const float*>(row + 512);
                } else {
                    xv[u] = *reinterpret_cast<const float*>(row + 256);
                    g0[u] = *reinterpret_cast<const float*>(row + 640);
                }
            }
            if (recv != st.i_cur) flush_rows<HAS_MU, ROLE>(st, recv, pipe, tb, Ex, rb, node0, ch, lane, q, mu_out);
            tmem_wait_ld();
            float a0 = st.a0, a1 = st.a1, a2 = st.a2;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float sx = __uint_as_float(w[u]) * xv[u];
                if (ROLE == 0) {
                    a0 += sx;
                } else if (ROLE == 1) {
                    a0 = fmaf(sx, g0[u], a0);
                    a1 = fmaf(sx, g1[u], a1);
                    a2 = fmaf(sx, g2[u], a2);
                } else if (ROLE == 2) {
                    a0 = fmaf(sx, g0[u], a0);
                    a1 = fmaf(sx, g1[u], a1);
                } else {
                    a0 = fmaf(sx, g0[u], a0);
                }
            }
            st.a0 = a0; st.a1 = a1; st.a2 = a2;
        }
        tc_fence_before();
        mbar_arrive(const_cast<uint64_t*>(&S->acc_free[pipe][buf]));
#ifdef RLVD_TIMELINE
        if (DBG) dbg[(t * 5 + ROLE + 1) * 4 + 2] = clock64();
#endif
    }
    flush_rows<HAS_MU, ROLE>(st, rb, pipe, tb, Ex, rb, node0, ch, lane, q, mu_out);
}

template <bool HAS_MU>
__global__ void __launch_bounds__(ES_THREADS, 1) tc_edge_small_kernel(
    int n_struct, int n_max, const int* __restrict__ node_ptr, const float* __restrict__ Wsl /* [4 slices][128][48] */,
    const int* __restrict__ row_ptr, const float4* __restrict__ rec, const float* __restrict__ x,
    const float* __restrict__ mu_in, float* __restrict__ q, float* __restrict__ mu_out, long long* dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Layout L = make_layout(n_max);
    uint8_t* Tab = smem + L.tab;
    int* rowp = reinterpret_cast<int*>(smem + L.rowp);
    EsBars* S = reinterpret_cast<EsBars*>(smem + L.bars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x / NSLICE, slice = blockIdx.x % NSLICE;
    const int node0 = __ldg(node_ptr + g), n = __ldg(node_ptr + g + 1) - node0;
    constexpr int NCONS = HAS_MU ? 4 : 2;                           // consumer warps per pipeline

    if (tid == 0) {
        for (int p = 0; p < NPIPE; ++p)
            for (int b = 0; b < 2; ++b) {
                mbar_init(&S->mma_done[p][b], 1);
                mbar_init(&S->acc_free[p][b], NCONS * 32);
            }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&S->tmem_base, TM_COLS);

    // ---- stage this slice of the structure's node table and the CSR row offsets
    {
        const uint32_t tab_s = smem_u32(Tab);
        const int per_node = HAS_MU ? 48 : 24;                      // 16-byte chunks: 3 (or 6) rows of 128 B
        for (int id = tid; id < n * per_node; id += ES_THREADS) {
            const int node = id / per_node, r = id - node * per_node, row = r >> 3, c16 = r & 7;
            const float* src = row < 3 ? x : mu_in;
            const size_t goff = ((size_t)(node0 + node) * 3 * F + (row % 3) * F + slice * SL) * 4 + c16 * 16;
            cp16(tab_s + node * NODE_BYTES + row * 128 + c16 * 16, reinterpret_cast<const uint8_t*>(src) + goff);
        }
        for (int id = tid; id <= n; id += ES_THREADS) rowp[id] = __ldg(row_ptr + node0 + id);
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;

    // ---- filter operand A → TMEM (lane = filter row: 0-31 dq, 32-63 dmuR, 64-95 dmumu, 96-127 dmumu again)
    if (warp < 4) {
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + TM_A;
        const float4* wrow = reinterpret_cast<const float4*>(Wsl + ((size_t)slice * 128 + warp * 32 + lane) * A_ROW);
#pragma unroll
        for (int c8 = 0; c8 < A_ROW / 8; ++c8) {
            const float4 f0 = __ldg(wrow + 2 * c8), f1 = __ldg(wrow + 2 * c8 + 1);
            const uint32_t v[8] = {__float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f0.z), __float_as_uint(f0.w),
                                   __float_as_uint(f1.x), __float_as_uint(f1.y), __float_as_uint(f1.z), __float_as_uint(f1.w)};
            tmem_st8(ta + c8 * 8, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const bool is_prod = warp >= NPIPE * 4;
    const int pipe = is_prod ? warp - NPIPE * 4 : warp >> 2;
    const int role = warp & 3;                                      // consumers: TMEM lane quarter
    const int ra = (pipe * n) / NPIPE, rb = ((pipe + 1) * n) / NPIPE;   // local receiver range of this pipeline
    int n_groups = 0;
    for (int i = ra; i < rb; ++i) n_groups += (rowp[i + 1] - rowp[i] + GS - 1) / GS;
    const int n_tiles = (n_groups + 3) / 4;
    uint8_t* Bt = smem + L.b + pipe * B_TILE;
    uint8_t* MetaP = smem + L.meta + pipe * 2 * META;
    float* Ex = reinterpret_cast<float*>(smem + L.exch) + pipe * 2 * 3 * SL;

    if (is_prod) {
        // ======================= producer + MMA issuer: lane = slot (group lane/8, position lane%8) =======================
        const uint32_t b_hi = smem_u32(Bt), b_lo = b_hi + B_HALF;
        uint64_t dbh[KP / 8], dbl[KP / 8];                              // B is single-buffered: descriptors are loop-invariant
#pragma unroll
        for (int ks = 0; ks < KP / 8; ++ks) {
            dbh[ks] = umma_desc(b_hi + ks * 1024, 512, 128);
            dbl[ks] = umma_desc(b_lo + ks * 1024, 512, 128);
        }
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int my_g = lane >> 3, my_u = lane & 7;
        int cur_i = ra, cur_e = rowp[ra];                               // group scheduler state (warp-uniform)
        float4 nx[7];
        bool nvalid = false;
        int nrecv = rb - 1;
        auto fetch = [&]() {                                            // schedule the next tile's 4 groups, load my record
            int eid = -1;
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
                while (cur_i < rb && cur_e >= rowp[cur_i + 1]) { ++cur_i; if (cur_i < rb) cur_e = rowp[cur_i]; }
                if (gg == my_g) {
                    if (cur_i < rb) {
                        nrecv = cur_i;
                        eid = cur_e + my_u < rowp[cur_i + 1] ? cur_e + my_u : -1;
                    } else {
                        nrecv = rb - 1;                                 // padding group: zero filters, harmless receiver
                    }
                }
                if (cur_i < rb) cur_e += GS;
            }
            nvalid = eid >= 0;
            if (nvalid) {
                const float4* r = rec + 8 * (int64_t)eid;
#pragma unroll
                for (int c = 0; c < 7; ++c) nx[c] = __ldg(r + c);
            }
        };
        fetch();
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t & 1;
            float4 v[6];
            int joff = 0;
            const int recv = nrecv;
            float4 dir4 = zero4;
#pragma unroll
            for (int c = 0; c < 6; ++c) v[c] = zero4;
            if (nvalid) {
#pragma unroll
                for (int c = 0; c < 5; ++c) v[c] = nx[c];
                v[5] = make_float4(nx[5].x, 0.f, 0.f, 0.f);              // k = 20: fcut (bias column); 21..23 zero
                dir4 = make_float4(nx[5].y, nx[5].z, nx[5].w, 0.f);
                joff = __float_as_int(nx[6].x) * NODE_BYTES;
            }
            if (t + 1 < n_tiles) fetch();                                // prefetch: hidden behind this tile's work
#ifdef RLVD_TIMELINE
            const bool DBG = dbg != nullptr && blockIdx.x == 0 && pipe == 0 && lane == 0 && t < 48;
            if (DBG) dbg[(t * 5) * 4 + 0] = clock64();
#endif
            mbar_wait(&S->acc_free[pipe][buf], ((t >> 1) & 1) ^ 1);     // consumers released this buffer's meta + TMEM
            if (t > 0) mbar_wait(&S->mma_done[pipe][buf ^ 1], ((t - 1) >> 1) & 1);   // previous tile's MMAs read B
#ifdef RLVD_TIMELINE
            if (DBG) dbg[(t * 5) * 4 + 1] = clock64();
#endif
            {
                uint8_t* meta = MetaP + buf * META;
                reinterpret_cast<int*>(meta)[lane] = joff;
                reinterpret_cast<float4*>(meta + 128)[lane] = dir4;
                if (my_u == 0) reinterpret_cast<int*>(meta + 640)[my_g] = recv;
                const uint32_t roff = (lane >> 3) * 128 + (lane & 7) * 16;
#pragma unroll
                for (int kc = 0; kc < 6; ++kc) {
                    float4 h, l;
                    h.x = tf32_hi(v[kc].x); h.y = tf32_hi(v[kc].y); h.z = tf32_hi(v[kc].z); h.w = tf32_hi(v[kc].w);
                    l.x = v[kc].x - h.x; l.y = v[kc].y - h.y; l.z = v[kc].z - h.z; l.w = v[kc].w - h.w;
                    *reinterpret_cast<float4*>(Bt + kc * 512 + roff) = h;
                    *reinterpret_cast<float4*>(Bt + B_HALF + kc * 512 + roff) = l;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tc_fence_after();
                const uint32_t d = tmem + TM_D + (uint32_t)(pipe * 2 + buf) * NE;
                const uint32_t a_hi = tmem + TM_A, a_lo = a_hi + KP;
#pragma unroll
                for (int ks = 0; ks < KP / 8; ++ks) {
                    umma_tf32_ts(d, a_hi + ks * 8, dbh[ks], ks == 0 ? 0u : 1u);
                    umma_tf32_ts(d, a_hi + ks * 8, dbl[ks], 1u);
                    umma_tf32_ts(d, a_lo + ks * 8, dbh[ks], 1u);
                }
                umma_commit(&S->mma_done[pipe][buf]);
#ifdef RLVD_TIMELINE
                if (DBG) dbg[(t * 5) * 4 + 2] = clock64();
#endif
            }
            __syncwarp();
        }
    } else if (role < NCONS) {
        // ======================= consumers: lane = channel of the slice, role = share of the message =======================
        // table row of sender j at Tab + j*768: x_q +0, x_R +128, x_m +256, mu_x +384, mu_y +512, mu_z +640
        const uint8_t* tb = Tab + lane * 4;
        const int ch = slice * SL + lane;
        if (role == 0)
            consume_tiles<HAS_MU, 0>(S, pipe, n_tiles, tmem, MetaP, tb, Ex, ra, rb, node0, ch, lane, q, mu_out, dbg);
        else if (role == 1)
            consume_tiles<HAS_MU, 1>(S, pipe, n_tiles, tmem, MetaP, tb, Ex, ra, rb, node0, ch, lane, q, mu_out, dbg);
        else if (role == 2)
            consume_tiles<HAS_MU, 2>(S, pipe, n_tiles, tmem, MetaP, tb, Ex, ra, rb, node0, ch, lane, q, mu_out, dbg);
        else
            consume_tiles<HAS_MU, 3>(S, pipe, n_tiles, tmem, MetaP, tb, Ex, ra, rb, node0, ch, lane, q, mu_out, dbg);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, TM_COLS);
    }
}

uint32_t tf32_round_host2(float w) {
    uint32_t u;
    memcpy(&u, &w, 4);
    u += 0xFFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    return u;
}

}  // namespace

// Host: per layer, per 32-channel slice: 128 rows (quarters: dq, dmuR, dmumu, dmumu again) of [hi[24] | lo[24]]
// tf32-split floats; k < 20 filter_net weights, k = 20 bias (meets the fcut column of the edge operand), 21..23 zero.
void tc_edge_small_format(const float* W, const float* b, std::vector<uint8_t>& out) {
    out.assign((size_t)NLAYER * NSLICE * A_SLICE, 0);
    float* o = reinterpret_cast<float*>(out.data());
    for (int l = 0; l < NLAYER; ++l)
        for (int sl = 0; sl < NSLICE; ++sl)
            for (int quarter = 0; quarter < 4; ++quarter)
                for (int chn = 0; chn < SL; ++chn)
                    for (int k = 0; k <= NRBF; ++k) {
                        const int part = quarter < 3 ? quarter : 2;
                        const int row = l * 3 * F + part * F + sl * SL + chn;
                        const float w = k < NRBF ? W[(size_t)row * NRBF + k] : b[row];
                        const uint32_t hu = tf32_round_host2(w);
                        float hf;
                        memcpy(&hf, &hu, 4);
                        const size_t base = (((size_t)l * NSLICE + sl) * 128 + quarter * SL + chn) * A_ROW;
                        o[base + k] = hf;
                        o[base + KP + k] = w - hf;
                    }
}

bool tc_edge_small_supported(int max_struct_nodes) { return max_struct_nodes > 0 && max_struct_nodes <= MAXN; }

int launch_edge_geometry(rlvd_engine* e, cudaStream_t s, int M, int64_t n_edges, const int* node_ptr,
                         const int* node_struct, const int* row_ptr, const int* col, const int8_t* shift, const float* pos,
                         const float* cell) {
    if (M <= 0 || n_edges <= 0) return 0;
    if (e->geom.ensure((size_t)n_edges * 128)) return 1;
    Span sp(e, s, FAM_PACK);
    edge_geom_kernel<<<(M + 7) / 8, 256, 0, s>>>(M, node_ptr, node_struct, row_ptr, col, shift, pos, cell, e->w.freqs,
                                                 e->w.feps, e->w.cutoff, e->geom.as<float4>());
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_tc_edge_small(rlvd_engine* e, cudaStream_t s, int layer, int n_struct, int max_struct_nodes,
                         const int* node_ptr, const int* row_ptr, const float* x, const float* mu_in, float* q,
                         float* mu_out) {
    if (n_struct <= 0) return 0;
    const Layout L = make_layout(max_struct_nodes);
    static bool attr_set = false;
    if (!attr_set) {
        const int cap = (int)make_layout(MAXN).total;
        RLVD_CUDA(cudaFuncSetAttribute(tc_edge_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        RLVD_CUDA(cudaFuncSetAttribute(tc_edge_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        attr_set = true;
    }
    const float* wf = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(e->w.filt_S) + (size_t)layer * NSLICE * A_SLICE);
    long long* dbg = nullptr;
#ifdef RLVD_TIMELINE   // in-kernel timeline of block 0 / pipeline 0 (build with RLVD_EXTRA_NVCC_FLAGS=-DRLVD_TIMELINE)
    static long long* dbg_buf = nullptr;
    static int dbg_calls = 0;
    if (dbg_buf == nullptr) { cudaMalloc(&dbg_buf, 48 * 5 * 4 * 8); cudaMemset(dbg_buf, 0, 48 * 5 * 4 * 8); }
    dbg = dbg_buf;
#endif
    Span sp(e, s, FAM_EDGE);
    if (mu_in == nullptr)
        tc_edge_small_kernel<false><<<n_struct * NSLICE, ES_THREADS, L.total, s>>>(
            n_struct, max_struct_nodes, node_ptr, wf, row_ptr, e->geom.as<float4>(), x, nullptr, q, mu_out, nullptr);
    else
        tc_edge_small_kernel<true><<<n_struct * NSLICE, ES_THREADS, L.total, s>>>(
            n_struct, max_struct_nodes, node_ptr, wf, row_ptr, e->geom.as<float4>(), x, mu_in, q, mu_out, dbg);
#ifdef RLVD_TIMELINE
    if (++dbg_calls == 20) {
        cudaStreamSynchronize(s);
        static long long h[48 * 5 * 4];
        cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        for (int t = 0; t < 40; ++t) {
            printf("tile %2d  P: ready %7lld waited %5lld issue %5lld |", t, h[t * 20] - t0, h[t * 20 + 1] - h[t * 20], h[t * 20 + 2] - h[t * 20 + 1]);
            for (int r = 0; r < 4; ++r)
                printf(" C%d: at %7lld wait %5lld work %5lld |", r, h[t * 20 + (r + 1) * 4] - t0, h[t * 20 + (r + 1) * 4 + 1] - h[t * 20 + (r + 1) * 4],
                       h[t * 20 + (r + 1) * 4 + 2] - h[t * 20 + (r + 1) * 4 + 1]);
            printf("\n");
        }
    }
#endif
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
