// K5 — batched per-action readout.
//
// Replaces the heads of rgnn's `painn` reaction model and `dqn_v2` reached from
// rlsim/drl/simulator.py:56 / rlsim/drl/trainer.py:112,166: per-atom energy MLP (128→64→1) with the
// species scale/shift, per-atom reaction MLP (128→128→32) on q_P - q_R, sum pooling per reaction graph
// (torch_scatter / global_add_pool in the reference stack), reaction_output (33→32→32→2), Q_emb
// (33→16→16), Q_output (18→32→32→1) and the q_params combination into rl_q.
//
// One CTA per reaction graph.  Atoms are processed in tiles of 32 held in shared memory; all sums run in a
// fixed order (tile order, then atom order), so the result is bitwise reproducible.  delta_e is pooled from
// per-atom differences scale_s*(o_P - o_R): the species shifts cancel exactly instead of being subtracted
// as two ~1.6 keV totals.
#include "common.cuh"

namespace rlvd {

namespace {

constexpr int RT = 256;  // threads
constexpr int TA = 32;   // atoms per tile
constexpr int QS = F + 4;  // padded row stride of the shared q tiles
constexpr int RO_TILE_CAP = 64;   // tiles per reaction the split form keeps partials for (structures of <= 2048 atoms)

__device__ __forceinline__ float act_fn(float v, int code) {
    switch (code) {
        case RLVD_ACT_RELU: return v > 0.f ? v : 0.f;
        case RLVD_ACT_TANH: return tanhf(v);
        case RLVD_ACT_LEAKY_RELU: return v > 0.f ? v : 0.01f * v;
        default: return __fdividef(v, 1.0f + __expf(-v));    // same fast SiLU as the tc_linear epilogue
    }
}

struct ReadoutOut {
    float *rl_q, *q0, *q1, *delta_e, *E_R, *E_P;
    float* head_in;   // [n_react, 36] or null: the DQN heads' inputs (33 pooled features + delta_e, 2 extras, base), train.cu
};

// y[n] = b[n] + sum_k W[n][k] * z[k]   (W out-major as in the checkpoint); lane = n.
__device__ __forceinline__ float head_layer(const float* __restrict__ W, const float* __restrict__ b, int n_out,
                                            int n_in, const float* z, int lane) {
    float acc = 0.f;
    if (lane < n_out) {
        acc = __ldg(b + lane);
        for (int k = 0; k < n_in; ++k) acc = fmaf(__ldg(W + lane * n_in + k), z[k], acc);
    }
    return acc;
}

// DEEP: the small-batch form (one CTA per SM at most: a single-replica call).  Its per-atom GEMM loops keep 8 k-steps of weight
// loads in flight instead of 2 — with one tile per CTA and no other CTA on the SM the loops are a chain of L2 round trips —
// and it may use the registers for that.  Same operations in the same order.
template <bool DEEP>
__global__ void __launch_bounds__(RT, DEEP ? 1 : 4) readout_kernel(int n_react, const int* __restrict__ node_ptr,
                                                     const int* __restrict__ rs, const int* __restrict__ ps,
                                                     const int* __restrict__ Z, const float* __restrict__ q,
                                                     const Weights w, const rlvd_q_params qp,
                                                     const float* __restrict__ d_temp, ReadoutOut out, int n_splits,
                                                     float* __restrict__ part /* [n_react][n_splits][36] */,
                                                     unsigned int* __restrict__ arrived /* [n_react], zero */) {
    extern __shared__ __align__(16) float sm[];
    float* sR = sm;                 // [TA][QS]
    float* sP = sm + TA * QS;       // [TA][QS]
    float* red = sP + TA * QS;      // [TA][8] partial sums / scratch
    float* eo = red + TA * 8;       // [2][TA] energy-head outputs (R, P)
    float* hpart = eo + 2 * TA;     // [TA][32] per-atom reaction features of the tile
    float* pool = hpart + TA * 32;  // [32] h, then E_R, E_P, dE
    float* zbuf = pool + 36;        // [64] head scratch

    const int g = blockIdx.x;
    if (g >= n_react) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int sr = rs[g], sp_ = ps[g];
    const int r0 = node_ptr[sr], p0 = node_ptr[sp_];
    const int n = node_ptr[sr + 1] - r0;
    if (tid < 36) pool[tid] = 0.f;
    __syncthreads();

    // small batches: blockIdx.y splits a reaction's atom tiles over n_splits CTAs (tile t goes to CTA t % n_splits); the
    // last CTA to finish adds the per-tile sums in tile order and runs the heads.  n_splits == 1 is the plain kernel; so
    // is a reaction with more tiles than the scratch keeps (CTA 0 walks all of them, the others only check in).
    const int n_tiles = (n + TA - 1) / TA;
    const bool split = n_splits > 1 && n_tiles <= RO_TILE_CAP;
    const int t_first = split ? (int)blockIdx.y * TA : (blockIdx.y == 0 ? 0 : n);
    const int t_step = split ? TA * n_splits : TA;
    for (int t0 = t_first; t0 < n; t0 += t_step) {
        const int na = min(TA, n - t0);
        // ---- load tiles (rows beyond na are zero)
        for (int f = tid; f < TA * (F / 4); f += RT) {
            const int a = f / (F / 4), c4 = f % (F / 4);
            float4 vr = make_float4(0.f, 0.f, 0.f, 0.f), vp = vr;
            if (a < na) {
                vr = __ldg(reinterpret_cast<const float4*>(q + (int64_t)(r0 + t0 + a) * F) + c4);
                vp = __ldg(reinterpret_cast<const float4*>(q + (int64_t)(p0 + t0 + a) * F) + c4);
            }
            *reinterpret_cast<float4*>(sR + a * QS + 4 * c4) = vr;
            *reinterpret_cast<float4*>(sP + a * QS + 4 * c4) = vp;
        }
        __syncthreads();

        // ---- H1: energy head on R and P.  Register tile 4 hidden units x (2 atoms x 2 sides): thread = (hidden group
        //      hg = tid%16, atom group ag = tid/16), atoms ag and ag+16.  Per 4 k: 4 LDG.128 of weights + 4 LDS.128 of
        //      activations feed 64 FMAs (the 1 x 8 tile this replaces was bound by the load/store pipe, not the FMA pipe).
        {
            const int hg = tid & 15, ag = tid >> 4;
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(w.en_b0) + hg);
            const float4 w3 = __ldg(reinterpret_cast<const float4*>(w.en_w3) + hg);
            float4 acc[4];      // [side * 2 + atom half]
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = b0;
            const float* rows[4] = {sR + ag * QS, sR + (ag + 16) * QS, sP + ag * QS, sP + (ag + 16) * QS};
#pragma unroll (DEEP ? 8 : 2)      // lets the next block's weight loads (L2 latency: the weights do not stay in L1) issue above this block's FMAs
            for (int k = 0; k < F; k += 4) {
                float4 wk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) wk[i] = __ldg(reinterpret_cast<const float4*>(w.en_w0T + (k + i) * 64) + hg);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 av = *reinterpret_cast<const float4*>(rows[r] + k);
                    const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[r].x = fmaf(a4[i], wk[i].x, acc[r].x);
                        acc[r].y = fmaf(a4[i], wk[i].y, acc[r].y);
                        acc[r].z = fmaf(a4[i], wk[i].z, acc[r].z);
                        acc[r].w = fmaf(a4[i], wk[i].w, acc[r].w);
                    }
                }
            }
            // second layer: dot over the 64 hidden units = the 16 lanes of one atom group (fixed shuffle tree)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float v = act_fn(acc[r].x, qp.painn_head_act) * w3.x + act_fn(acc[r].y, qp.painn_head_act) * w3.y +
                          act_fn(acc[r].z, qp.painn_head_act) * w3.z + act_fn(acc[r].w, qp.painn_head_act) * w3.w;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (hg == 0) eo[(r >> 1) * TA + ag + 16 * (r & 1)] = v + __ldg(w.en_b3);
            }
            __syncthreads();
        }
        // ---- pooled energies (thread 0, atom order)
        //      Every pooled quantity is summed per TILE from zero and the tile sums are added in tile order: the result
        //      does not depend on which CTA computed a tile, i.e. not on the batch size (split form below).
        if (tid == 0) {
            float eR = 0.f, eP = 0.f, dE = 0.f;
            for (int a = 0; a < na; ++a) {
                const int s = __ldg(w.elem_lookup + min(max(__ldg(Z + r0 + t0 + a), 0), 99));
                const float sc = __ldg(w.sp_scales + s), sh = __ldg(w.sp_shifts + s);
                eR += fmaf(eo[a], sc, sh);
                eP += fmaf(eo[TA + a], sc, sh);
                dE += sc * (eo[TA + a] - eo[a]);
            }
            if (split) {
                float* pt = part + ((int64_t)g * RO_TILE_CAP + t0 / TA) * 36;
                pt[32] = eR; pt[33] = eP; pt[34] = dE;
            } else {
                pool[32] += eR; pool[33] += eP; pool[34] += dE;
            }
        }
        // ---- D = q_P - q_R (in place in sP)
        for (int f = tid; f < TA * F; f += RT) {
            const int a = f / F, cc = f % F;
            sP[a * QS + cc] -= sR[a * QS + cc];
        }
        __syncthreads();
        // ---- H2 layer 0: Hd = act(D·Wr0T + b0) → sR.  Register tile 4 hidden units x 4 atoms: hg = tid%32, ag = tid/32,
        //      atoms ag + 8r
        {
            const int hg = tid & 31, ag = tid >> 5;
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(w.rr_b0) + hg);
            float4 acc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = b0;
#pragma unroll (DEEP ? 8 : 2)
            for (int k = 0; k < F; k += 4) {
                float4 wk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) wk[i] = __ldg(reinterpret_cast<const float4*>(w.rr_w0T + (k + i) * F) + hg);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 av = *reinterpret_cast<const float4*>(sP + (ag + 8 * r) * QS + k);
                    const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[r].x = fmaf(a4[i], wk[i].x, acc[r].x);
                        acc[r].y = fmaf(a4[i], wk[i].y, acc[r].y);
                        acc[r].z = fmaf(a4[i], wk[i].z, acc[r].z);
                        acc[r].w = fmaf(a4[i], wk[i].w, acc[r].w);
                    }
                }
            }
            // sR is free: its last readers (the energy head) are two barriers back
#pragma unroll
            for (int r = 0; r < 4; ++r)
                *reinterpret_cast<float4*>(sR + (ag + 8 * r) * QS + 4 * hg) =
                    make_float4(act_fn(acc[r].x, qp.painn_head_act), act_fn(acc[r].y, qp.painn_head_act),
                                act_fn(acc[r].z, qp.painn_head_act), act_fn(acc[r].w, qp.painn_head_act));
        }
        __syncthreads();
        // ---- H2 layer 3 + sum pool: 4 outputs x 1 atom per thread: og = tid%8, atom a = tid/8
        {
            const int og = tid & 7, a = tid >> 3;
            float4 acc = __ldg(reinterpret_cast<const float4*>(w.rr_b3) + og);
#pragma unroll (DEEP ? 8 : 1)
            for (int k = 0; k < F; k += 4) {
                const float4 av = *reinterpret_cast<const float4*>(sR + a * QS + k);
                const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 wk = __ldg(reinterpret_cast<const float4*>(w.rr_w3T + (k + i) * 32) + og);
                    acc.x = fmaf(a4[i], wk.x, acc.x);
                    acc.y = fmaf(a4[i], wk.y, acc.y);
                    acc.z = fmaf(a4[i], wk.z, acc.z);
                    acc.w = fmaf(a4[i], wk.w, acc.w);
                }
            }
            if (a >= na) acc = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(hpart + a * 32 + 4 * og) = acc;
        }
        __syncthreads();
        if (tid < 32) {
            float h = 0.f;
            for (int a = 0; a < TA; ++a) h += hpart[a * 32 + tid];     // atom order
            if (split) part[((int64_t)g * RO_TILE_CAP + t0 / TA) * 36 + tid] = h;
            else pool[tid] += h;
        }
        __syncthreads();
    }

    if (n_splits > 1) {
        __shared__ unsigned int ticket;
        if (!split && blockIdx.y == 0 && tid < 35) part[(int64_t)g * RO_TILE_CAP * 36 + tid] = pool[tid];   // one "tile"
        __threadfence();
        __syncthreads();
        if (tid == 0) ticket = atomicAdd(arrived + g, 1u);
        __syncthreads();
        if (ticket != (unsigned int)(n_splits - 1)) return;
        __threadfence();
        if (tid < 35) {
            float v = 0.f;
            const int n_comb = split ? n_tiles : 1;
            for (int t = 0; t < n_comb; ++t) v += __ldcg(part + ((int64_t)g * RO_TILE_CAP + t) * 36 + tid);   // tile order
            pool[tid] = v;
        }
        if (tid == 0) arrived[g] = 0u;            // ready for the next launch
        __syncthreads();
    }
    // ---- H3 / H4 in warp 0
    if (wid == 0) {
        const float E_R = pool[32], E_P = pool[33], dE = pool[34];
        const float de_in = qp.delta_e_standardised ? (dE - qp.mean_delta_e) / qp.std_delta_e : dE;
        float* z = zbuf;        // [33]
        float* t1 = zbuf + 40;  // scratch
        z[lane] = pool[lane];
        if (lane == 0) z[32] = de_in;
        __syncwarp();
        // reaction_output 33 → 32 → 32 → 2
        float v = act_fn(head_layer(w.ro_w[0], w.ro_b[0], 32, 33, z, lane), qp.painn_head_act);
        t1[lane] = v;
        __syncwarp();
        v = act_fn(head_layer(w.ro_w[1], w.ro_b[1], 32, 32, t1, lane), qp.painn_head_act);
        __syncwarp();
        t1[lane] = v;
        __syncwarp();
        v = head_layer(w.ro_w[2], w.ro_b[2], 2, 32, t1, lane);
        const float bhat = __shfl_sync(0xffffffffu, v, 0), fhat = __shfl_sync(0xffffffffu, v, 1);
        const float barrier = fmaf(bhat, qp.std_barrier, qp.mean_barrier);
        const float freq = fmaf(fhat, qp.std_freq, qp.mean_freq);
        const float q0 = -barrier, q1 = freq;
        const float T = d_temp != nullptr ? __ldg(d_temp + g) : qp.temperature;
        const float kT = T * 8.617e-5f;
        float rl = qp.alpha * (q0 + kT * q1) + qp.beta * (-dE);
        if (qp.dqn) {
            __syncwarp();
            // Q_emb 33 → 16 → 16
            v = act_fn(head_layer(w.qe_w[0], w.qe_b[0], 16, 33, z, lane), qp.dqn_head_act);
            t1[lane] = v;
            __syncwarp();
            v = head_layer(w.qe_w[1], w.qe_b[1], 16, 16, t1, lane);
            __syncwarp();
            if (lane < 16) t1[lane] = v;
            if (lane == 0) {
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    t1[16 + u] = qp.q_extra[u] == RLVD_QX_KT_X10 ? kT * 10.0f : dE;
            }
            __syncwarp();
            if (out.head_in != nullptr) {         // what the trainable DQN heads see (rlvd_q_head_train_step)
                float* hi = out.head_in + (int64_t)g * RLVD_HEAD_IN;
                hi[lane] = z[lane];
                if (lane == 0) { hi[32] = z[32]; hi[33] = t1[16]; hi[34] = t1[17]; hi[35] = rl; }
            }
            // Q_output 18 → 32 → 32 → 1
            v = act_fn(head_layer(w.qo_w[0], w.qo_b[0], 32, 18, t1, lane), qp.dqn_head_act);
            __syncwarp();
            z[lane] = v;
            __syncwarp();
            v = act_fn(head_layer(w.qo_w[1], w.qo_b[1], 32, 32, z, lane), qp.dqn_head_act);
            __syncwarp();
            t1[lane] = v;
            __syncwarp();
            v = head_layer(w.qo_w[2], w.qo_b[2], 1, 32, t1, lane);
            rl += __shfl_sync(0xffffffffu, v, 0);
        }
        if (lane == 0) {
            if (out.rl_q) out.rl_q[g] = rl;
            if (out.q0) out.q0[g] = q0;
            if (out.q1) out.q1[g] = q1;
            if (out.delta_e) out.delta_e[g] = dE;
            if (out.E_R) out.E_R[g] = E_R;
            if (out.E_P) out.E_P[g] = E_P;
        }
    }
}

size_t readout_smem() { return (size_t)(2 * TA * QS + TA * 8 + 2 * TA + TA * 32 + 36 + 128) * sizeof(float); }

}  // namespace

namespace {

// T1 — t_net head (called at rlsim/cli/scripts/estimate_time.py:58 and rlsim/time/train.py:108):
// pool the node scalars of ONE structure, append the two per-frame scalars, run the 3-layer MLP
// (hidden_channels+2 -> n_feat -> n_feat -> 1) and squash to [0, tau).  One CTA per structure, thread = channel /
// hidden unit; sums run in atom order (bitwise reproducible).
constexpr int TT = 128;

__global__ void __launch_bounds__(TT) time_readout_kernel(int n_struct, const int* __restrict__ node_ptr,
                                                          const float* __restrict__ q, const float* __restrict__ T,
                                                          const float* __restrict__ defect, const rlvd_time_head h,
                                                          float* __restrict__ time_out) {
    __shared__ float z[F + 2];
    __shared__ float h0[TT], h1[TT];
    const int g = blockIdx.x, tid = threadIdx.x;
    if (g >= n_struct) return;
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += __ldg(q + (int64_t)(n0 + i) * F + tid);
    z[tid] = h.pooling == 0 ? (n > 0 ? acc / (float)n : 0.f) : acc;
    if (tid == 0) {
        z[F] = __ldg(T + g) * h.T_scale;
        z[F + 1] = __ldg(defect + g) * h.D_scale;
    }
    __syncthreads();
    const int nf = h.n_feat, k0 = F + 2;
    if (tid < nf) {
        float v = __ldg(h.d_b0 + tid);
        for (int k = 0; k < k0; ++k) v = fmaf(__ldg(h.d_w0 + tid * k0 + k), z[k], v);
        h0[tid] = act_fn(v, h.act);
    }
    __syncthreads();
    if (tid < nf) {
        float v = __ldg(h.d_b1 + tid);
        for (int k = 0; k < nf; ++k) v = fmaf(__ldg(h.d_w1 + tid * nf + k), h0[k], v);
        h1[tid] = act_fn(v, h.act);
    }
    __syncthreads();
    if (tid == 0) {
        float v = __ldg(h.d_b2);
        for (int k = 0; k < nf; ++k) v = fmaf(__ldg(h.d_w2 + k), h1[k], v);
        time_out[g] = h.output == 1 ? v : h.tau / (1.0f + expf(-v));
    }
}

}  // namespace

int launch_time_readout(rlvd_engine* e, cudaStream_t s, int n_struct, const int* node_ptr, const float* q, const float* T,
                        const float* defect, const rlvd_time_head& h, float* time_out) {
    if (n_struct <= 0) return 0;
    RLVD_CHECK(h.n_feat > 0 && h.n_feat <= TT, "rlvd_time_readout: n_feat %d not in [1,%d]", h.n_feat, TT);
    Span sp(e, s, FAM_READOUT);
    time_readout_kernel<<<n_struct, TT, 0, s>>>(n_struct, node_ptr, q, T, defect, h, time_out);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_readout(rlvd_engine* e, cudaStream_t s, int n_react, const int* node_ptr, const int* rs, const int* ps,
                   const int* Z, const float* q, const rlvd_q_params& qp, const float* d_temp, float* rl_q,
                   float* q0, float* q1, float* delta_e, float* E_R, float* E_P, float* head_in) {
    if (n_react <= 0) return 0;
    ReadoutOut out{rl_q, q0, q1, delta_e, E_R, E_P, head_in};
    Span sp(e, s, FAM_READOUT);
    // few reactions (single-replica stepping, get_max_Q batches): split each reaction's atom tiles over several CTAs
    int n_splits = 1;
    if (n_react < 296) n_splits = min(8, (592 + n_react - 1) / n_react);
    float* part = nullptr;
    unsigned int* arrived = nullptr;
    if (n_splits > 1) {
        const size_t part_bytes = (size_t)n_react * RO_TILE_CAP * 36 * sizeof(float);
        const size_t cnt_off = (part_bytes + 255) & ~(size_t)255;
        if (e->readout_scratch.ensure(cnt_off + 296 * sizeof(unsigned int))) return 1;
        part = e->readout_scratch.as<float>();
        arrived = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(e->readout_scratch.p) + cnt_off);
        RLVD_CUDA(cudaMemsetAsync(arrived, 0, 296 * sizeof(unsigned int), s));   // the kernel also re-zeroes what it used
    }
    if (n_react * n_splits <= e->sm_count)
        readout_kernel<true><<<dim3(n_react, n_splits), RT, readout_smem(), s>>>(n_react, node_ptr, rs, ps, Z, q, e->w, qp, d_temp, out,
                                                                                n_splits, part, arrived);
    else
        readout_kernel<false><<<dim3(n_react, n_splits), RT, readout_smem(), s>>>(n_react, node_ptr, rs, ps, Z, q, e->w, qp, d_temp, out,
                                                                                 n_splits, part, arrived);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
