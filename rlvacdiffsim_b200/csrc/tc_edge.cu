// K2 + K3a + K3c + K4 fused, filters on the tensor cores (SURVEY.md §8a R2, R3, R5 edge part).
//
// Same contract as edge.cu (which stays as the fp32 SIMT cross-check): per-edge Bessel/cutoff/unit-vector
// embedding, continuous-filter evaluation W_ij = (filter_net(phi_ij))·fcut, messages W_ij ⊙ x_j and the
// deterministic segmented sum over receiver-sorted edges.  Here the filter evaluation — 58 % of the path's
// FLOPs, a [E,20(+bias)] x [21,384] product — runs as tcgen05.mma with the OUTPUT TRANSPOSED so that it lands
// in TMEM exactly as the message loop wants it: lane = channel, column = edge.
//
//   D_part[128 channels, 32 edges] = Wf_part[128, 24] · Phi[32 edges, 24]ᵀ          part ∈ {dq, dmuR, dmumu}
//
// K = 20 Bessel values phi_k·fcut, one bias column (fcut), 3 zero pads.  Operands are split into tf32 hi/lo
// (hi = rna_tf32(v), lo = v - hi) and accumulated as hi·hi + hi·lo + lo·hi into one fp32 TMEM accumulator:
// 22-bit operands, no scaling games, fp32-level filters.
//
// CTA = 64 consecutive receivers (their CSR rows are one contiguous edge range, cut into 32-edge tiles).
//   warps 0-3  consumers: thread = channel = TMEM lane.  tcgen05.ld the three filter values of 8 edges at a
//              time, gather x[j], mu[j] (coalesced 512 B rows), accumulate the row in CSR order, flush q / mu on
//              every receiver change — no atomics, bitwise reproducible
//   warps 4-7  producers: geometry of the next tile (minimum-image vector, d, fcut, dir, 20 x sin) written as
//              the tf32 hi/lo B operand in the canonical K-major core-matrix layout + per-edge metadata
//   warp  8    TMEM owner (256 columns = 2 tiles x 3 parts x 32), bulk-TMA load of the layer's pre-split filter
//              weights (72 KB, once per CTA), one elected lane issues the 27 MMAs per tile
// Producer, MMA and consumer stages are decoupled by mbarriers over double-buffered operand tiles and
// accumulators; two CTAs fit per SM (87 KB shared memory, 256 TMEM columns each).
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace rlvd {

namespace {

constexpr int NE = 32;                 // edges per tile = MMA N
constexpr int KP = 24;                 // padded K: 20 rbf + bias + 3 zeros (3 k-steps of 8 tf32)
constexpr int RC = 64;                 // receivers per CTA (small enough that the in-flight node tables stay in L2)
constexpr int TE_THREADS = 288;
constexpr int A_HALF = 6 * 2048;       // one of hi/lo of one part: 6 core columns x (16 row groups x 128 B)
constexpr int A_BYTES = 3 * 2 * A_HALF;   // 73728
constexpr int B_HALF = 6 * 512;        // 6 core columns x (4 row groups x 128 B)
constexpr int B_TILE = 2 * B_HALF;     // hi | lo
constexpr int META_BYTES = 768;        // j[32] | recv[32] | dir[32] (float4)
constexpr float PI_F = 3.14159265358979323846f;
constexpr uint32_t IDESC_TF32_M128_N32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NE >> 3) << 17) | (8u << 24);

struct TeSmem {
    uint64_t b_full[2], mma_done[2], acc_free[2], w_full;
    uint32_t tmem_base;
    float sd[NE], sscale[NE], sfc[NE];
    float freqs[NRBF];
};

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(IDESC_TF32_M128_N32), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ void prod_barrier() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

template <bool HAS_MU>
__global__ void __launch_bounds__(TE_THREADS, 2) tc_edge_kernel(
    int M, const uint8_t* __restrict__ Wfmt /* this layer: [3 parts][hi|lo][A_HALF] */,
    const float* __restrict__ freqs, const int* __restrict__ row_ptr, const int* __restrict__ col,
    const int8_t* __restrict__ shift, const float* __restrict__ pos, const float* __restrict__ cell,
    const int* __restrict__ node_struct, const float* __restrict__ x, const float* __restrict__ mu_in,
    float* __restrict__ q, float* __restrict__ mu_out, float cutoff) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Aw = smem;                               // 72 KB filter weights (tf32 hi/lo, canonical layout)
    uint8_t* Bt = smem + A_BYTES;                     // 2 x 6 KB operand tiles
    uint8_t* Meta = Bt + 2 * B_TILE;                  // 2 x 768 B
    TeSmem* S = reinterpret_cast<TeSmem*>(Meta + 2 * META_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i_begin = blockIdx.x * RC, i_end = min(M, i_begin + RC);
    const int e_begin = __ldg(row_ptr + i_begin), e_end = __ldg(row_ptr + i_end);
    const int n_tiles = (e_end - e_begin + NE - 1) / NE;
    constexpr int NPART = HAS_MU ? 3 : 2;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&S->b_full[b], 128);
            mbar_init(&S->mma_done[b], 1);
            mbar_init(&S->acc_free[b], 128);
        }
        mbar_init(&S->w_full, 1);
        fence_mbar_init();
    }
    if (tid < NRBF) S->freqs[tid] = __ldg(freqs + tid);
    if (warp == 8) tmem_alloc(&S->tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;

    if (warp < 4) {
        // ======================= consumers: thread = channel =======================
        const int c = tid;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        int i_cur = i_begin;
        float dq = 0.f, dm0 = 0.f, dm1 = 0.f, dm2 = 0.f;
        auto flush_to = [&](int next) {   // write receiver i_cur (with its sums) and the edgeless receivers up to next-1
            for (int i = i_cur; i < next; ++i) {
                const int64_t qi = (int64_t)i * F + c, mi = (int64_t)i * 3 * F + c;
                q[qi] += dq;
                if (HAS_MU) {
                    mu_out[mi] = __ldg(mu_in + mi) + dm0;
                    mu_out[mi + F] = __ldg(mu_in + mi + F) + dm1;
                    mu_out[mi + 2 * F] = __ldg(mu_in + mi + 2 * F) + dm2;
                } else {
                    mu_out[mi] = dm0;
                    mu_out[mi + F] = dm1;
                    mu_out[mi + 2 * F] = dm2;
                }
                dq = dm0 = dm1 = dm2 = 0.f;
            }
            i_cur = next;
        };
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t & 1;
            mbar_wait(&S->mma_done[buf], (t >> 1) & 1);
            tc_fence_after();
            const uint8_t* meta = Meta + buf * META_BYTES;
            const uint32_t tbase = tmem + lane_base + buf * (3 * NE);
#pragma unroll 1
            for (int g = 0; g < NE / 8; ++g) {
                uint32_t wq[8], wr[8], wm[8];
                tmem_ld8(tbase + g * 8, wq);
                tmem_ld8(tbase + NE + g * 8, wr);
                if (HAS_MU) tmem_ld8(tbase + 2 * NE + g * 8, wm);
                tmem_wait_ld();
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int e0 = g * 8 + hf * 4;
                    const int4 j4 = *reinterpret_cast<const int4*>(meta + e0 * 4);
                    const int4 r4 = *reinterpret_cast<const int4*>(meta + 128 + e0 * 4);
                    const int jj[4] = {j4.x, j4.y, j4.z, j4.w};
                    const int rr[4] = {r4.x, r4.y, r4.z, r4.w};
                    float xq[4], xr[4], xm[4], m0[4], m1[4], m2[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float* xj = x + (int64_t)jj[u] * 3 * F + c;
                        xq[u] = __ldg(xj);
                        xr[u] = __ldg(xj + F);
                        if (HAS_MU) {
                            const float* mj = mu_in + (int64_t)jj[u] * 3 * F + c;
                            xm[u] = __ldg(xj + 2 * F);
                            m0[u] = __ldg(mj);
                            m1[u] = __ldg(mj + F);
                            m2[u] = __ldg(mj + 2 * F);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (rr[u] != i_cur) flush_to(rr[u]);
                        const float4 dir = *reinterpret_cast<const float4*>(meta + 256 + (e0 + u) * 16);
                        const int w = hf * 4 + u;
                        dq = fmaf(__uint_as_float(wq[w]), xq[u], dq);
                        const float sR = __uint_as_float(wr[w]) * xr[u];
                        dm0 = fmaf(sR, dir.x, dm0);
                        dm1 = fmaf(sR, dir.y, dm1);
                        dm2 = fmaf(sR, dir.z, dm2);
                        if (HAS_MU) {
                            const float sM = __uint_as_float(wm[w]) * xm[u];
                            dm0 = fmaf(sM, m0[u], dm0);
                            dm1 = fmaf(sM, m1[u], dm1);
                            dm2 = fmaf(sM, m2[u], dm2);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&S->acc_free[buf]);
        }
        flush_to(i_end);
    } else if (warp < 8) {
        // ======================= producers: geometry → tf32 hi/lo operand tile =======================
        const int p = tid - 128;
        const float pi_over_rc = PI_F / cutoff;
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t & 1;
            mbar_wait(&S->acc_free[buf], ((t >> 1) & 1) ^ 1);   // consumers are done with this buffer's meta + TMEM
            uint8_t* meta = Meta + buf * META_BYTES;
            if (p < NE) {
                const int eid = e_begin + t * NE + p;
                int j = i_begin, recv = i_end - 1;
                float d = 1.0f, scale = 0.f, fc = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
                if (eid < e_end) {
                    int lo = i_begin, hi = i_end;              // largest i with row_ptr[i] <= eid
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (__ldg(row_ptr + mid) <= eid) lo = mid; else hi = mid;
                    }
                    recv = lo;
                    j = __ldg(col + eid);
                    const float S0 = (float)shift[3 * (int64_t)eid], S1 = (float)shift[3 * (int64_t)eid + 1],
                                S2 = (float)shift[3 * (int64_t)eid + 2];
                    const float* C = cell + 9 * __ldg(node_struct + recv);
                    float rx = __ldg(pos + 3 * (int64_t)j) - __ldg(pos + 3 * (int64_t)recv);
                    float ry = __ldg(pos + 3 * (int64_t)j + 1) - __ldg(pos + 3 * (int64_t)recv + 1);
                    float rz = __ldg(pos + 3 * (int64_t)j + 2) - __ldg(pos + 3 * (int64_t)recv + 2);
                    rx += S0 * __ldg(C + 0) + S1 * __ldg(C + 3) + S2 * __ldg(C + 6);
                    ry += S0 * __ldg(C + 1) + S1 * __ldg(C + 4) + S2 * __ldg(C + 7);
                    rz += S0 * __ldg(C + 2) + S1 * __ldg(C + 5) + S2 * __ldg(C + 8);
                    d = sqrtf(rx * rx + ry * ry + rz * rz);
                    const float inv_d = 1.0f / d;
                    fc = d < cutoff ? 0.5f * (cosf(d * pi_over_rc) + 1.0f) : 0.f;
                    scale = inv_d * fc;
                    ux = rx * inv_d; uy = ry * inv_d; uz = rz * inv_d;
                }
                reinterpret_cast<int*>(meta)[p] = j;
                reinterpret_cast<int*>(meta + 128)[p] = recv;
                reinterpret_cast<float4*>(meta + 256)[p] = make_float4(ux, uy, uz, 0.f);
                S->sd[p] = d;
                S->sscale[p] = scale;
                S->sfc[p] = fc;
            }
            prod_barrier();
            uint8_t* bh = Bt + buf * B_TILE;
#pragma unroll
            for (int it = 0; it < NE * KP / 128; ++it) {
                const int item = it * 128 + p, e = item & (NE - 1), k = item >> 5;
                float v = 0.f;
                if (k < NRBF) {
                    const float f = S->freqs[k], d = S->sd[e];
                    const float pr = f * d;
                    const float perr = fmaf(f, d, -pr);        // rounding error of the fp32 product, folded back in
                    float sn, cs;
                    sincosf(pr, &sn, &cs);
                    v = fmaf(perr, cs, sn) * S->sscale[e];
                } else if (k == NRBF) {
                    v = S->sfc[e];
                }
                const float vh = tf32_hi(v);
                const uint32_t off = (k >> 2) * 512 + (e >> 3) * 128 + (e & 7) * 16 + (k & 3) * 4;
                *reinterpret_cast<float*>(bh + off) = vh;
                *reinterpret_cast<float*>(bh + B_HALF + off) = v - vh;
            }
            fence_async_smem();
            mbar_arrive(&S->b_full[buf]);
            prod_barrier();                                       // sd/sscale/sfc reusable
        }
    } else {
        // ======================= weights loader + MMA issuer =======================
        if (lane == 0) {
            const uint32_t wbytes = NPART * 2 * A_HALF;
            mbar_expect_tx(&S->w_full, wbytes);
            for (int part = 0; part < NPART; ++part)
                bulk_g2s(Aw + part * 2 * A_HALF, Wfmt + (size_t)part * 2 * A_HALF, 2 * A_HALF, &S->w_full);
            mbar_wait(&S->w_full, 0);
            const uint32_t a0 = smem_u32(Aw), b0 = smem_u32(Bt);
            for (int t = 0; t < n_tiles; ++t) {
                const int buf = t & 1;
                mbar_wait(&S->b_full[buf], (t >> 1) & 1);
                tc_fence_after();
                const uint32_t bb = b0 + buf * B_TILE;
#pragma unroll
                for (int part = 0; part < NPART; ++part) {
                    const uint32_t d = tmem + buf * (3 * NE) + part * NE;
                    const uint32_t ah = a0 + part * 2 * A_HALF, al = ah + A_HALF;
#pragma unroll
                    for (int ks = 0; ks < KP / 8; ++ks) {
                        const uint64_t dah = umma_desc(ah + ks * 4096, 2048, 128), dal = umma_desc(al + ks * 4096, 2048, 128);
                        const uint64_t dbh = umma_desc(bb + ks * 1024, 512, 128),
                                       dbl = umma_desc(bb + B_HALF + ks * 1024, 512, 128);
                        umma_tf32(d, dah, dbh, ks == 0 ? 0u : 1u);
                        umma_tf32(d, dah, dbl, 1u);
                        umma_tf32(d, dal, dbh, 1u);
                    }
                }
                umma_commit(&S->mma_done[buf]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

size_t te_smem_bytes() { return (size_t)A_BYTES + 2 * B_TILE + 2 * META_BYTES + sizeof(TeSmem) + 64; }

uint32_t tf32_round_host(float w) {
    uint32_t u;
    memcpy(&u, &w, 4);
    u += 0xFFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    return u;
}

}  // namespace

// Host: filter_net weight [L*384, 20] + bias [L*384] → per layer [3 parts][hi | lo][6 core columns][128 rows].
void tc_edge_format_filters(const float* W, const float* b, std::vector<uint8_t>& out) {
    out.assign((size_t)NLAYER * A_BYTES, 0);
    for (int l = 0; l < NLAYER; ++l)
        for (int part = 0; part < 3; ++part)
            for (int r = 0; r < F; ++r)
                for (int k = 0; k <= NRBF; ++k) {
                    const int row = l * 3 * F + part * F + r;
                    const float w = k < NRBF ? W[(size_t)row * NRBF + k] : b[row];
                    const uint32_t hu = tf32_round_host(w);
                    float hf;
                    memcpy(&hf, &hu, 4);
                    const float lf = w - hf;
                    const size_t off = (size_t)l * A_BYTES + (size_t)part * 2 * A_HALF + (k >> 2) * 2048 + (r >> 3) * 128 +
                                       (r & 7) * 16 + (k & 3) * 4;
                    memcpy(out.data() + off, &hf, 4);
                    memcpy(out.data() + off + A_HALF, &lf, 4);
                }
}

int launch_tc_edge(rlvd_engine* e, cudaStream_t s, int layer, int M, const int* row_ptr, const int* col,
                   const int8_t* shift, const float* pos, const float* cell, const int* node_struct, const float* x,
                   const float* mu_in, float* q, float* mu_out) {
    if (M <= 0) return 0;
    if (!e->attr_tc_edge) {                 // per engine = per device: the attribute is a per-device property
        RLVD_CUDA(cudaFuncSetAttribute(tc_edge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)te_smem_bytes()));
        RLVD_CUDA(cudaFuncSetAttribute(tc_edge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)te_smem_bytes()));
        e->attr_tc_edge = true;
    }
    const uint8_t* wf = reinterpret_cast<const uint8_t*>(e->w.filt_F) + (size_t)layer * A_BYTES;
    const int grid = (M + RC - 1) / RC;
    Span sp(e, s, FAM_EDGE);
    if (mu_in == nullptr)
        tc_edge_kernel<false><<<grid, TE_THREADS, te_smem_bytes(), s>>>(M, wf, e->w.freqs, row_ptr, col, shift, pos, cell,
                                                                        node_struct, x, nullptr, q, mu_out, e->w.cutoff);
    else
        tc_edge_kernel<true><<<grid, TE_THREADS, te_smem_bytes(), s>>>(M, wf, e->w.freqs, row_ptr, col, shift, pos, cell,
                                                                       node_struct, x, mu_in, q, mu_out, e->w.cutoff);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
