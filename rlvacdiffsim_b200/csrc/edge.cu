// K2 + K3a + K3c + K4 fused — per-edge embedding, continuous-filter evaluation, message construction and the
// deterministic segmented reduction over receiver-sorted edges (SURVEY.md §8a R2, R3, R5 edge part).
//
// Replaces, inside rgnn's PaiNN (reached from rlsim/drl/simulator.py:56): the Bessel/cosine-cutoff/unit-vector
// embedding, `filter_net` (nn.Linear 20→1152, materialised as [E,1152] by the eager path), the
// index_select gathers of x[j] / mu[j], and torch_scatter.scatter_add (atomic, order-nondeterministic).
//
// One thread owns one of the 128 channels; a CTA walks a run of receivers.  For each receiver the CSR row
// is processed in chunks of 64 edges: the chunk's geometry record (20 x phi_k*fcut, fcut, dir) is computed
// once by the CTA into shared memory, then every thread evaluates ITS three filter values from the layer's
// filter weights held in registers (never materialising [E,1152]) and accumulates the messages of the row
// in CSR order — no atomics, bitwise reproducible.
#include "common.cuh"

namespace rlvd {

namespace {

constexpr int EDGE_THREADS = F;  // one channel per thread
constexpr int EC = 64;           // edges per shared-memory chunk
constexpr int RPC = 8;           // receivers per CTA
constexpr float PI_F = 3.14159265358979323846f;

template <bool HAS_MU>
__global__ void __launch_bounds__(EDGE_THREADS) edge_kernel(
    int M, const float* __restrict__ filt_T /* [21][384] of this layer */, const float* __restrict__ freqs,
    const int* __restrict__ row_ptr, const int* __restrict__ col, const int8_t* __restrict__ shift,
    const float* __restrict__ pos, const float* __restrict__ cell, const int* __restrict__ node_struct,
    const float* __restrict__ x, const float* __restrict__ mu_in, float* __restrict__ q,
    float* __restrict__ mu_out, float cutoff) {
    __shared__ __align__(16) float g[EC][GEOM];
    __shared__ int sj[EC];
    __shared__ float sd[EC], sscale[EC];
    __shared__ float sfreq[NRBF];

    const int c = threadIdx.x;
    float wq[NRBF], wr[NRBF], wm[NRBF];
#pragma unroll
    for (int k = 0; k < NRBF; ++k) {
        wq[k] = __ldg(filt_T + k * 3 * F + c);
        wr[k] = __ldg(filt_T + k * 3 * F + F + c);
        wm[k] = __ldg(filt_T + k * 3 * F + 2 * F + c);
    }
    const float bq = __ldg(filt_T + NRBF * 3 * F + c);
    const float br = __ldg(filt_T + NRBF * 3 * F + F + c);
    const float bm = __ldg(filt_T + NRBF * 3 * F + 2 * F + c);
    if (c < NRBF) sfreq[c] = __ldg(freqs + c);
    const float pi_over_rc = PI_F / cutoff;

    const int i_begin = blockIdx.x * RPC;
    const int i_end = min(M, i_begin + RPC);
    for (int i = i_begin; i < i_end; ++i) {
        const int e0 = __ldg(row_ptr + i), e1 = __ldg(row_ptr + i + 1);
        const int gs = __ldg(node_struct + i);
        const float xi = __ldg(pos + 3 * (int64_t)i), yi = __ldg(pos + 3 * (int64_t)i + 1),
                    zi = __ldg(pos + 3 * (int64_t)i + 2);
        float dq = 0.f, dm0 = 0.f, dm1 = 0.f, dm2 = 0.f;
        for (int eb = e0; eb < e1; eb += EC) {
            const int ne = min(EC, e1 - eb);
            __syncthreads();  // previous chunk fully consumed
            if (c < ne) {
                const int e = eb + c;
                const int j = __ldg(col + e);
                const float S0 = (float)shift[3 * (int64_t)e], S1 = (float)shift[3 * (int64_t)e + 1],
                            S2 = (float)shift[3 * (int64_t)e + 2];
                const float* C = cell + 9 * gs;
                float rx = __ldg(pos + 3 * (int64_t)j) - xi, ry = __ldg(pos + 3 * (int64_t)j + 1) - yi,
                      rz = __ldg(pos + 3 * (int64_t)j + 2) - zi;
                rx += S0 * __ldg(C + 0) + S1 * __ldg(C + 3) + S2 * __ldg(C + 6);
                ry += S0 * __ldg(C + 1) + S1 * __ldg(C + 4) + S2 * __ldg(C + 7);
                rz += S0 * __ldg(C + 2) + S1 * __ldg(C + 5) + S2 * __ldg(C + 8);
                const float d = sqrtf(rx * rx + ry * ry + rz * rz);
                const float inv_d = 1.0f / d;
                const float fc = d < cutoff ? 0.5f * (cosf(d * pi_over_rc) + 1.0f) : 0.f;
                sj[c] = j;
                sd[c] = d;
                sscale[c] = inv_d * fc;
                g[c][NRBF] = fc;
                g[c][NRBF + 1] = rx * inv_d;
                g[c][NRBF + 2] = ry * inv_d;
                g[c][NRBF + 3] = rz * inv_d;
            }
            __syncthreads();
            for (int item = c; item < ne * NRBF; item += EDGE_THREADS) {
                const int ee = item / NRBF, k = item - ee * NRBF;
                // sin(freq*d) with the rounding error of the fp32 product folded back in
                const float f = sfreq[k], d = sd[ee];
                const float p = f * d;
                const float perr = fmaf(f, d, -p);
                float sn, cs;
                sincosf(p, &sn, &cs);
                g[ee][k] = fmaf(perr, cs, sn) * sscale[ee];
            }
            __syncthreads();
            for (int ee = 0; ee < ne; ++ee) {
                const float4* gp = reinterpret_cast<const float4*>(&g[ee][0]);
                const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2], g3 = gp[3], g4 = gp[4], g5 = gp[5];
                const float ph[NRBF] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y,
                                        g2.z, g2.w, g3.x, g3.y, g3.z, g3.w, g4.x, g4.y, g4.z, g4.w};
                const float fc = g5.x;
                float fq = bq * fc, fr = br * fc, fm = bm * fc;
#pragma unroll
                for (int k = 0; k < NRBF; ++k) {
                    fq = fmaf(wq[k], ph[k], fq);
                    fr = fmaf(wr[k], ph[k], fr);
                    if (HAS_MU) fm = fmaf(wm[k], ph[k], fm);
                }
                const int j = sj[ee];
                const float* xj = x + (int64_t)j * 3 * F + c;
                dq = fmaf(fq, __ldg(xj), dq);
                const float sR = fr * __ldg(xj + F);
                dm0 = fmaf(sR, g5.y, dm0);
                dm1 = fmaf(sR, g5.z, dm1);
                dm2 = fmaf(sR, g5.w, dm2);
                if (HAS_MU) {
                    const float sM = fm * __ldg(xj + 2 * F);
                    const float* mj = mu_in + (int64_t)j * 3 * F + c;
                    dm0 = fmaf(sM, __ldg(mj), dm0);
                    dm1 = fmaf(sM, __ldg(mj + F), dm1);
                    dm2 = fmaf(sM, __ldg(mj + 2 * F), dm2);
                }
            }
        }
        const int64_t qi = (int64_t)i * F + c;
        q[qi] += dq;
        const int64_t mi = (int64_t)i * 3 * F + c;
        if (HAS_MU) {
            mu_out[mi] = __ldg(mu_in + mi) + dm0;
            mu_out[mi + F] = __ldg(mu_in + mi + F) + dm1;
            mu_out[mi + 2 * F] = __ldg(mu_in + mi + 2 * F) + dm2;
        } else {
            mu_out[mi] = dm0;
            mu_out[mi + F] = dm1;
            mu_out[mi + 2 * F] = dm2;
        }
    }
}

}  // namespace

// The fp32 SIMT kernel with a caller-supplied filter table (train_full.cu: the training forward reads the flat master
// parameters, not the engine's inference layouts).
int launch_edge_simt(rlvd_engine* e, cudaStream_t s, const float* filt_T_layer, const float* freqs, float cutoff, const EdgeArgs& a) {
    const int M = a.M;
    if (M <= 0) return 0;
    const int grid = (M + RPC - 1) / RPC;
    Span sp(e, s, FAM_EDGE);
    if (a.mu_in == nullptr) {
        edge_kernel<false><<<grid, EDGE_THREADS, 0, s>>>(M, filt_T_layer, freqs, a.row_ptr, a.col, a.shift, a.pos, a.cell, a.node_struct,
                                                         a.x, nullptr, a.q, a.mu_out, cutoff);
    } else {
        edge_kernel<true><<<grid, EDGE_THREADS, 0, s>>>(M, filt_T_layer, freqs, a.row_ptr, a.col, a.shift, a.pos, a.cell, a.node_struct,
                                                        a.x, a.mu_in, a.q, a.mu_out, cutoff);
    }
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_edge(rlvd_engine* e, cudaStream_t s, const EdgeArgs& a) {
    const int M = a.M;
    if (M <= 0) return 0;
    if (a.sliced == 2)
        return launch_edge_stream(e, s, a.layer, a.n_struct, a.max_struct_nodes, a.node_ptr, a.x, a.mu_in, a.q, a.mu_out, a.gate, a.gate_value);
    if (e->use_tensor_cores)
        return launch_tc_edge(e, s, a.layer, M, a.row_ptr, a.col, a.shift, a.pos, a.cell, a.node_struct, a.x, a.mu_in, a.q,
                              a.mu_out);
    const float* filt = e->w.filt_T + (size_t)a.layer * (NRBF + 1) * 3 * F;
    const int grid = (M + RPC - 1) / RPC;
    Span sp(e, s, FAM_EDGE);
    if (a.mu_in == nullptr) {
        edge_kernel<false><<<grid, EDGE_THREADS, 0, s>>>(M, filt, e->w.freqs, a.row_ptr, a.col, a.shift, a.pos, a.cell,
                                                         a.node_struct, a.x, nullptr, a.q, a.mu_out, e->w.cutoff);
    } else {
        edge_kernel<true><<<grid, EDGE_THREADS, 0, s>>>(M, filt, e->w.freqs, a.row_ptr, a.col, a.shift, a.pos, a.cell,
                                                        a.node_struct, a.x, a.mu_in, a.q, a.mu_out, e->w.cutoff);
    }
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
