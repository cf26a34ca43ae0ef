// K3b — per-node channel-mixing linears of PaiNN (SURVEY.md §8a R4, R5 node part, R6) — fp32 SIMT tiles.
//
//   embed        q = embedding[Z]                                         (R4)
//   linear       Y = act(A·Wt + b), A optionally the concatenation [A | A2] (interatomic / intra-atomic
//                context nets, mu_channel_mix with A = mu viewed as [3M,128])
//   mix_reduce   n = sqrt(sum_xyz mu_V^2 + eps),  vw = sum_xyz mu_V*mu_W     (R6)
//   mix_update   q += a + c*vw,  mu += b*mu_W                                (R6)
//
// The linear kernel is a classic 128 x TN register-tiled SGEMM: 256 threads, 8 x (TN/16) accumulators per
// thread, K streamed in chunks of 16 through double-buffered shared memory; A is stored transposed in
// shared memory so both operands are read with 128-bit loads.  Accumulation order over K is fixed
// (ascending), so results are run-to-run deterministic.
#include "common.cuh"

namespace rlvd {

namespace {

constexpr int TM = 128;
constexpr int KC = 16;
constexpr int LIN_THREADS = 256;

__device__ __forceinline__ float silu(float v) { return v / (1.0f + expf(-v)); }

template <int TN>
__global__ void __launch_bounds__(LIN_THREADS) linear_kernel(const LinearArgs a) {
    constexpr int RM = 8, RN = TN / 16;
    constexpr int B_F4 = KC * TN / 4;                       // float4 per B chunk
    constexpr int B_PER_THREAD = (B_F4 + LIN_THREADS - 1) / LIN_THREADS;
    __shared__ __align__(16) float As[2][KC][TM + 4];
    __shared__ __align__(16) float Bs[2][KC][TN];

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    float acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[B_PER_THREAD];
    auto load_chunk = [&](int k0) {
        const float* src = a.A;
        int ld = a.lda, kk0 = k0;
        if (a.A2 != nullptr && k0 >= a.K1) { src = a.A2; ld = a.lda2; kk0 = k0 - a.K1; }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int f = tid + LIN_THREADS * r, row = f >> 2, kq = f & 3;
            const int m = m0 + row;
            ra[r] = m < a.M ? __ldg(reinterpret_cast<const float4*>(src + (int64_t)m * ld + kk0 + 4 * kq))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < B_PER_THREAD; ++r) {
            const int f = tid + LIN_THREADS * r;
            if (f < B_F4) {
                const int kk = f / (TN / 4), c4 = f % (TN / 4);
                rb[r] = __ldg(reinterpret_cast<const float4*>(a.Wt + (int64_t)(k0 + kk) * a.N + n0 + 4 * c4));
            }
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int f = tid + LIN_THREADS * r, row = f >> 2, kq = f & 3;
            As[buf][4 * kq + 0][row] = ra[r].x;
            As[buf][4 * kq + 1][row] = ra[r].y;
            As[buf][4 * kq + 2][row] = ra[r].z;
            As[buf][4 * kq + 3][row] = ra[r].w;
        }
#pragma unroll
        for (int r = 0; r < B_PER_THREAD; ++r) {
            const int f = tid + LIN_THREADS * r;
            if (f < B_F4) {
                const int kk = f / (TN / 4), c4 = f % (TN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][kk][4 * c4]) = rb[r];
            }
        }
    };

    const int nchunks = a.K / KC;
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) load_chunk((c + 1) * KC);
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float av[RM], bv[RN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
            av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
            if constexpr (RN == 8) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][TN / 2 + tx * 4]);
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            } else if constexpr (RN == 4) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
            } else {
                const float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
                bv[0] = b0.x; bv[1] = b0.y;
            }
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (c + 1 < nchunks) store_chunk(buf ^ 1);
        __syncthreads();
    }

    // epilogue: bias + activation, vectorised stores
    constexpr int NSEG = RN == 8 ? 2 : 1;
    constexpr int SEGW = RN == 8 ? 4 : RN;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int sgm = 0; sgm < NSEG; ++sgm) {
            const int col = n0 + (RN == 8 ? sgm * (TN / 2) + tx * 4 : tx * RN);
            float v[SEGW];
#pragma unroll
            for (int j = 0; j < SEGW; ++j) {
                float t = acc[i][sgm * 4 + j];
                if (a.bias != nullptr) t += __ldg(a.bias + col + j);
                if (a.act == 1) t = silu(t);
                v[j] = t;
            }
            float* dst = a.Y + (int64_t)m * a.ldy + col;
            if constexpr (SEGW == 4) {
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
                *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
            }
        }
    }
}

// y += x: adds the layer-0 message sum (a per-node GEMM output, engine.cu) to q.
__global__ void axpy_kernel(int64_t n4, const float4* __restrict__ x, float4* __restrict__ y, const int* __restrict__ gate,
                            int gate_value) {
    if (gate != nullptr && *gate != gate_value) return;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n4) return;
    const float4 a = __ldg(x + t);
    float4 b = y[t];
    b.x += a.x; b.y += a.y; b.z += a.z; b.w += a.w;
    y[t] = b;
}

__global__ void embed_kernel(int M, const int* __restrict__ Z, const float* __restrict__ emb, float* __restrict__ q) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 per thread
    if (idx >= (int64_t)M * (F / 4)) return;
    const int i = (int)(idx / (F / 4)), c4 = (int)(idx % (F / 4));
    int z = __ldg(Z + i);
    z = z < 0 ? 0 : (z > 99 ? 99 : z);
    reinterpret_cast<float4*>(q)[idx] = __ldg(reinterpret_cast<const float4*>(emb + z * F) + c4);
}

// mix[M][3][256] (V = cols 0..127, W = cols 128..255) → nrm[M][128], vw[M][128]
__global__ void mix_reduce_kernel(int M, const float* __restrict__ mix, float* __restrict__ nrm,
                                  float* __restrict__ vw, float eps) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * (F / 4)) return;
    const int i = (int)(idx / (F / 4)), c4 = (int)(idx % (F / 4));
    float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f), dot = n2;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(mix + ((int64_t)i * 3 + k) * 2 * F) + c4);
        const float4 w = __ldg(reinterpret_cast<const float4*>(mix + ((int64_t)i * 3 + k) * 2 * F + F) + c4);
        n2.x = fmaf(v.x, v.x, n2.x); n2.y = fmaf(v.y, v.y, n2.y); n2.z = fmaf(v.z, v.z, n2.z); n2.w = fmaf(v.w, v.w, n2.w);
        dot.x = fmaf(v.x, w.x, dot.x); dot.y = fmaf(v.y, w.y, dot.y); dot.z = fmaf(v.z, w.z, dot.z); dot.w = fmaf(v.w, w.w, dot.w);
    }
    reinterpret_cast<float4*>(nrm)[idx] = make_float4(sqrtf(n2.x + eps), sqrtf(n2.y + eps), sqrtf(n2.z + eps), sqrtf(n2.w + eps));
    reinterpret_cast<float4*>(vw)[idx] = dot;
}

// ctx[M][384] = [a | b | c];  q += a + c*vw;  mu[k] += b * W[k]   (W row (i, k) at W + (3 i + k) * ldw)
__global__ void mix_update_kernel(int M, const float* __restrict__ ctx, const float* __restrict__ W, int ldw,
                                  const float* __restrict__ vw, float* __restrict__ q, float* __restrict__ mu) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * (F / 4)) return;
    const int i = (int)(idx / (F / 4)), c4 = (int)(idx % (F / 4));
    const float4 a = __ldg(reinterpret_cast<const float4*>(ctx + (int64_t)i * 3 * F) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(ctx + (int64_t)i * 3 * F + F) + c4);
    const float4 c = __ldg(reinterpret_cast<const float4*>(ctx + (int64_t)i * 3 * F + 2 * F) + c4);
    const float4 d = __ldg(reinterpret_cast<const float4*>(vw) + idx);
    float4 qv = reinterpret_cast<float4*>(q)[idx];
    qv.x += a.x + c.x * d.x; qv.y += a.y + c.y * d.y; qv.z += a.z + c.z * d.z; qv.w += a.w + c.w * d.w;
    reinterpret_cast<float4*>(q)[idx] = qv;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + ((int64_t)i * 3 + k) * ldw) + c4);
        float4* mp = reinterpret_cast<float4*>(mu + ((int64_t)i * 3 + k) * F) + c4;
        float4 m = *mp;
        m.x = fmaf(b.x, w.x, m.x); m.y = fmaf(b.y, w.y, m.y); m.z = fmaf(b.z, w.z, m.z); m.w = fmaf(b.w, w.w, m.w);
        *mp = m;
    }
}

// ctx_ac[M][256] = [a | c];  q += a + c*vw        (final mixing block when nobody reads mu afterwards)
__global__ void mix_update_q_kernel(int M, const float* __restrict__ ctx_ac, const float* __restrict__ vw, float* __restrict__ q) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * (F / 4)) return;
    const int i = (int)(idx / (F / 4)), c4 = (int)(idx % (F / 4));
    const float4 a = __ldg(reinterpret_cast<const float4*>(ctx_ac + (int64_t)i * 2 * F) + c4);
    const float4 c = __ldg(reinterpret_cast<const float4*>(ctx_ac + (int64_t)i * 2 * F + F) + c4);
    const float4 d = __ldg(reinterpret_cast<const float4*>(vw) + idx);
    float4 qv = reinterpret_cast<float4*>(q)[idx];
    qv.x += a.x + c.x * d.x; qv.y += a.y + c.y * d.y; qv.z += a.z + c.z * d.z; qv.w += a.w + c.w * d.w;
    reinterpret_cast<float4*>(q)[idx] = qv;
}

}  // namespace

int launch_linear(rlvd_engine* e, cudaStream_t s, const LinearArgs& a) {
    if (e->use_tensor_cores && a.Wfmt != nullptr && tc_linear_supported(a)) return launch_tc_linear(e, s, a, a.Wfmt);
    RLVD_CHECK(a.gate == nullptr, "linear: gated launches exist on the tcgen05 path only");
    RLVD_CHECK(a.mixred == 0, "linear: the mix-reduce form exists on the tcgen05 path only (M=%d n_nodes=%d)", a.M, a.n_nodes);
    RLVD_CHECK(a.W2fmt == nullptr, "linear: the fused two-layer form exists on the tcgen05 path only (shape M=%d K=%d N=%d N2=%d)",
               a.M, a.K, a.N, a.N2);
    RLVD_CHECK(a.Wt != nullptr, "linear: no fp32 weight for the SIMT path");
    RLVD_CHECK(a.K % KC == 0, "linear: K=%d is not a multiple of %d", a.K, KC);
    RLVD_CHECK(a.A2 == nullptr || a.K1 % KC == 0, "linear: K1=%d is not a multiple of %d", a.K1, KC);
    RLVD_CHECK(a.lda % 4 == 0 && a.ldy % 4 == 0 && a.N % 4 == 0, "linear: unaligned leading dimension");
    if (a.M <= 0) return 0;
    Span sp(e, s, FAM_NODE);
    const int gm = (a.M + TM - 1) / TM;
    if (a.N % 128 == 0) {
        linear_kernel<128><<<dim3(gm, a.N / 128), LIN_THREADS, 0, s>>>(a);
    } else if (a.N % 64 == 0) {
        linear_kernel<64><<<dim3(gm, a.N / 64), LIN_THREADS, 0, s>>>(a);
    } else if (a.N % 32 == 0) {
        linear_kernel<32><<<dim3(gm, a.N / 32), LIN_THREADS, 0, s>>>(a);
    } else {
        sp.end(0);
        RLVD_CHECK(false, "linear: N=%d is not a multiple of 32", a.N);
    }
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_embed(rlvd_engine* e, cudaStream_t s, int M, const int* Z, float* q) {
    if (M <= 0) return 0;
    Span sp(e, s, FAM_NODE);
    const int64_t n = (int64_t)M * (F / 4);
    embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(M, Z, e->w.emb, q);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_mix_reduce(rlvd_engine* e, cudaStream_t s, int M, const float* mix, float* nrm, float* vw) {
    if (M <= 0) return 0;
    Span sp(e, s, FAM_NODE);
    const int64_t n = (int64_t)M * (F / 4);
    mix_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(M, mix, nrm, vw, e->eps);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_axpy(rlvd_engine* e, cudaStream_t s, int64_t n, const float* x, float* y, const int* gate, int gate_value) {
    if (n <= 0) return 0;
    Span sp(e, s, FAM_NODE);
    axpy_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(n / 4, reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y),
                                                                gate, gate_value);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_mix_update_q(rlvd_engine* e, cudaStream_t s, int M, const float* ctx_ac, const float* vw, float* q) {
    if (M <= 0) return 0;
    Span sp(e, s, FAM_NODE);
    const int64_t n = (int64_t)M * (F / 4);
    mix_update_q_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(M, ctx_ac, vw, q);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_mix_update(rlvd_engine* e, cudaStream_t s, int M, const float* ctx, const float* W, int ldw, const float* vw,
                      float* q, float* mu) {
    if (M <= 0) return 0;
    Span sp(e, s, FAM_NODE);
    const int64_t n = (int64_t)M * (F / 4);
    mix_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(M, ctx, W, ldw, vw, q, mu);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
