// B1, the `train_all = False` form: one replay-minibatch update of the DQN heads.
//
// Replaces, for rlsim/drl/trainer.py:163-177 with the reaction model frozen (trainer.py:36-39):
//   q_pred = policy_value_net(batch, q_params)["rl_q"]; loss = mean((batch["rl_q"] - q_pred)**2); loss.backward();
//   clip_grad_norm_(parameters, 0.8); Adam.step()
// Only Q_emb (33→16→16) and Q_output (18→32→32→1) carry gradients then (2513 parameters); everything upstream of them
// (K1..K4 and the reaction-model heads) is the frozen forward that already runs on the device and hands over the heads'
// inputs per reaction graph (readout.cu, `head_in`).  One CTA does forward, backward, the global-norm clip and the Adam
// step on the engine's weights in place.  Every sum runs in a fixed order (graphs ascending, a fixed reduction tree), so
// an update is bitwise reproducible.  Train-mode dropout of the heads (rgnn's MLP = [Linear, act, Dropout], p = 0.15) is
// driven by caller-supplied uniforms, like the action draw in actions.cu; without them dropout is the identity.
#include "common.cuh"

namespace rlvd {

namespace {

constexpr int TR_THREADS = 256;
constexpr int NP = RLVD_QHEAD_PARAMS;
// parameter offsets inside the packed vector (checkpoint layout, out-major)
constexpr int O_EW0 = 0, O_EB0 = O_EW0 + 16 * 33, O_EW1 = O_EB0 + 16, O_EB1 = O_EW1 + 16 * 16, O_OW0 = O_EB1 + 16,
              O_OB0 = O_OW0 + 32 * 18, O_OW1 = O_OB0 + 32, O_OB1 = O_OW1 + 32 * 32, O_OW2 = O_OB1 + 32, O_OB2 = O_OW2 + 32;
static_assert(O_OB2 + 1 == NP, "packed Q-head parameter count");
// per-graph scratch (floats): inputs, activations, deltas
constexpr int S_Z = 0, S_E1P = 33, S_E1 = 49, S_T = 65, S_O1P = 83, S_O1 = 115, S_O2P = 147, S_O2 = 179, S_D_O2 = 211,
              S_D_O1 = 243, S_D_E2 = 275, S_D_E1 = 291, S_DOUT = 307, S_M_E1 = 308, S_M_O1 = 324, S_M_O2 = 356, S_PER = 388;

struct QHeadPtrs {
    float* p[10];   // qe_w0, qe_b0, qe_w1, qe_b1, qo_w0, qo_b0, qo_w1, qo_b1, qo_w2, qo_b2
};
__constant__ int c_off[11] = {O_EW0, O_EB0, O_EW1, O_EB1, O_OW0, O_OB0, O_OW1, O_OB1, O_OW2, O_OB2, NP};

__device__ __forceinline__ float act_f(float v, int code) {
    switch (code) {
        case RLVD_ACT_RELU: return v > 0.f ? v : 0.f;
        case RLVD_ACT_TANH: return tanhf(v);
        case RLVD_ACT_LEAKY_RELU: return v > 0.f ? v : 0.01f * v;
        default: return v / (1.0f + expf(-v));
    }
}
__device__ __forceinline__ float act_d(float v, int code) {
    switch (code) {
        case RLVD_ACT_RELU: return v > 0.f ? 1.f : 0.f;
        case RLVD_ACT_TANH: { const float t = tanhf(v); return 1.f - t * t; }
        case RLVD_ACT_LEAKY_RELU: return v > 0.f ? 1.f : 0.01f;
        default: { const float s = 1.0f / (1.0f + expf(-v)); return s * (1.0f + v * (1.0f - s)); }
    }
}

struct TrainArgs {
    int G, act, step, apply;
    float lr, beta1, beta2, eps, max_norm, p_drop;
    const float* head_in;   // [G, 36]
    const float* target;    // [G]
    const float* drop_u;    // [G, 80] uniforms (16 Q_emb hidden, 32 + 32 Q_output hidden) or null
    float *m, *v;           // Adam state [NP]
    float *loss, *pred, *gnorm, *grad_out;
};

__global__ void __launch_bounds__(TR_THREADS) q_head_train_kernel(QHeadPtrs W, TrainArgs a) {
    extern __shared__ float sm[];
    float* P = sm;                       // [NP] parameters
    float* Gd = P + NP;                  // [NP] gradients
    float* red = Gd + NP;                // [TR_THREADS]
    float* scal = red + TR_THREADS;      // [4] loss, norm, clip coefficient
    float* S = scal + 4;                 // [G][S_PER]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = 0; t < 10; ++t)
        for (int i = tid; i < c_off[t + 1] - c_off[t]; i += TR_THREADS) P[c_off[t] + i] = W.p[t][i];
    const float keep = 1.0f / (1.0f - a.p_drop);
    for (int i = tid; i < a.G * RLVD_HEAD_IN; i += TR_THREADS) {
        const int g = i / RLVD_HEAD_IN, c = i % RLVD_HEAD_IN;
        float* s = S + g * S_PER;
        const float v = a.head_in[i];
        if (c < 33) s[S_Z + c] = v;
        else if (c < 35) s[S_T + 16 + (c - 33)] = v;      // the two extra Q_output inputs
        else s[S_DOUT] = v;                                // base (alpha / beta part of rl_q), replaced by d_out later
    }
    for (int i = tid; i < a.G * 80; i += TR_THREADS) {     // dropout masks (scaled), 1 when dropout is off
        const int g = i / 80, c = i % 80;
        const float mk = a.drop_u != nullptr && a.p_drop > 0.f ? (a.drop_u[i] >= a.p_drop ? keep : 0.f) : 1.f;
        S[g * S_PER + S_M_E1 + c] = mk;                    // S_M_E1 (16) | S_M_O1 (32) | S_M_O2 (32) are contiguous
    }
    __syncthreads();
    // ---- forward, one warp per graph, lane = output unit
    for (int g = warp; g < a.G; g += TR_THREADS / 32) {
        float* s = S + g * S_PER;
        const float base = s[S_DOUT];
        if (lane < 16) {
            float acc = P[O_EB0 + lane];
            for (int k = 0; k < 33; ++k) acc = fmaf(P[O_EW0 + lane * 33 + k], s[S_Z + k], acc);
            s[S_E1P + lane] = acc;
            s[S_E1 + lane] = act_f(acc, a.act) * s[S_M_E1 + lane];
        }
        __syncwarp();
        if (lane < 16) {
            float acc = P[O_EB1 + lane];
            for (int k = 0; k < 16; ++k) acc = fmaf(P[O_EW1 + lane * 16 + k], s[S_E1 + k], acc);
            s[S_T + lane] = acc;
        }
        __syncwarp();
        {
            float acc = P[O_OB0 + lane];
            for (int k = 0; k < 18; ++k) acc = fmaf(P[O_OW0 + lane * 18 + k], s[S_T + k], acc);
            s[S_O1P + lane] = acc;
            s[S_O1 + lane] = act_f(acc, a.act) * s[S_M_O1 + lane];
        }
        __syncwarp();
        {
            float acc = P[O_OB1 + lane];
            for (int k = 0; k < 32; ++k) acc = fmaf(P[O_OW1 + lane * 32 + k], s[S_O1 + k], acc);
            s[S_O2P + lane] = acc;
            s[S_O2 + lane] = act_f(acc, a.act) * s[S_M_O2 + lane];
        }
        __syncwarp();
        if (lane == 0) {
            float acc = P[O_OB2];
            for (int k = 0; k < 32; ++k) acc = fmaf(P[O_OW2 + k], s[S_O2 + k], acc);
            const float pred = base + acc;
            if (a.pred != nullptr) a.pred[g] = pred;
            s[S_DOUT] = pred - a.target[g];               // residual, scaled to dL/dpred below
        }
    }
    __syncthreads();
    if (tid == 0) {                                       // loss = mean(residual^2), graphs ascending
        float l = 0.f;
        for (int g = 0; g < a.G; ++g) l = fmaf(S[g * S_PER + S_DOUT], S[g * S_PER + S_DOUT], l);
        scal[0] = l / (float)a.G;
        if (a.loss != nullptr) *a.loss = scal[0];
    }
    __syncthreads();
    // ---- backward per graph: deltas at every pre-activation
    for (int g = warp; g < a.G; g += TR_THREADS / 32) {
        float* s = S + g * S_PER;
        const float dout = 2.0f * s[S_DOUT] / (float)a.G;  // d loss / d pred_g
        __syncwarp();
        if (lane == 0) s[S_DOUT] = dout;
        s[S_D_O2 + lane] = dout * P[O_OW2 + lane] * s[S_M_O2 + lane] * act_d(s[S_O2P + lane], a.act);
        __syncwarp();
        {
            float acc = 0.f;
            for (int r = 0; r < 32; ++r) acc = fmaf(s[S_D_O2 + r], P[O_OW1 + r * 32 + lane], acc);
            s[S_D_O1 + lane] = acc * s[S_M_O1 + lane] * act_d(s[S_O1P + lane], a.act);
        }
        __syncwarp();
        if (lane < 16) {
            float acc = 0.f;
            for (int r = 0; r < 32; ++r) acc = fmaf(s[S_D_O1 + r], P[O_OW0 + r * 18 + lane], acc);
            s[S_D_E2 + lane] = acc;
        }
        __syncwarp();
        if (lane < 16) {
            float acc = 0.f;
            for (int r = 0; r < 16; ++r) acc = fmaf(s[S_D_E2 + r], P[O_EW1 + r * 16 + lane], acc);
            s[S_D_E1 + lane] = acc * s[S_M_E1 + lane] * act_d(s[S_E1P + lane], a.act);
        }
    }
    __syncthreads();
    // ---- parameter gradients: thread per parameter, graphs ascending
    float sq = 0.f;
    for (int p = tid; p < NP; p += TR_THREADS) {
        int t = 0;
        while (p >= c_off[t + 1]) ++t;
        const int i = p - c_off[t];
        int d_off, in_off, n_in;                            // delta vector, input vector, fan-in (0 for a bias)
        switch (t) {
            case 0: d_off = S_D_E1; in_off = S_Z; n_in = 33; break;
            case 1: d_off = S_D_E1; in_off = 0; n_in = 0; break;
            case 2: d_off = S_D_E2; in_off = S_E1; n_in = 16; break;
            case 3: d_off = S_D_E2; in_off = 0; n_in = 0; break;
            case 4: d_off = S_D_O1; in_off = S_T; n_in = 18; break;
            case 5: d_off = S_D_O1; in_off = 0; n_in = 0; break;
            case 6: d_off = S_D_O2; in_off = S_O1; n_in = 32; break;
            case 7: d_off = S_D_O2; in_off = 0; n_in = 0; break;
            case 8: d_off = S_DOUT; in_off = S_O2; n_in = 32; break;
            default: d_off = S_DOUT; in_off = 0; n_in = 0; break;
        }
        const int r = n_in > 0 ? i / n_in : i, c = n_in > 0 ? i % n_in : 0;
        float gsum = 0.f;
        for (int g = 0; g < a.G; ++g) {
            const float* s = S + g * S_PER;
            gsum = n_in > 0 ? fmaf(s[d_off + r], s[in_off + c], gsum) : gsum + s[d_off + r];
        }
        Gd[p] = gsum;
        sq = fmaf(gsum, gsum, sq);
    }
    red[tid] = sq;
    __syncthreads();
    for (int o = TR_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        const float norm = sqrtf(red[0]);
        const float coef = a.max_norm / (norm + 1e-6f);     // torch.nn.utils.clip_grad_norm_
        scal[1] = norm;
        scal[2] = coef < 1.0f ? coef : 1.0f;
        if (a.gnorm != nullptr) *a.gnorm = norm;
    }
    __syncthreads();
    // ---- clip + Adam (torch.optim.Adam defaults: no weight decay, no amsgrad), in place on the engine's weights
    const float clip = scal[2];
    const float bc1 = 1.0f - powf(a.beta1, (float)a.step), bc2 = 1.0f - powf(a.beta2, (float)a.step);
    const float step_size = a.lr / bc1, bc2_sqrt = sqrtf(bc2);
    for (int p = tid; p < NP; p += TR_THREADS) {
        const float gr = Gd[p] * clip;
        if (a.grad_out != nullptr) a.grad_out[p] = gr;
        if (!a.apply) continue;
        const float m = a.beta1 * a.m[p] + (1.0f - a.beta1) * gr;
        const float v = a.beta2 * a.v[p] + (1.0f - a.beta2) * gr * gr;
        a.m[p] = m;
        a.v[p] = v;
        const float denom = sqrtf(v) / bc2_sqrt + a.eps;
        int t = 0;
        while (p >= c_off[t + 1]) ++t;
        W.p[t][p - c_off[t]] = P[p] - step_size * (m / denom);
    }
}

}  // namespace

int launch_q_head_train(rlvd_engine* e, cudaStream_t s, int n_graphs, const float* head_in, const float* target,
                        const float* drop_u, float p_drop, int act, float lr, float beta1, float beta2, float eps,
                        float max_norm, int step, int apply, float* m, float* v, float* loss, float* pred, float* gnorm,
                        float* grad_out) {
    if (n_graphs <= 0) return 0;
    const size_t smem = (size_t)(2 * NP + TR_THREADS + 4 + (size_t)n_graphs * S_PER) * sizeof(float);
    RLVD_CHECK(smem <= 200 * 1024, "q_head_train: minibatch of %d reaction graphs does not fit one CTA (max ~120)", n_graphs);
    if (!e->attr_train) {                 // per engine = per device: the attribute is a per-device property
        RLVD_CUDA(cudaFuncSetAttribute(q_head_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        e->attr_train = true;
    }
    const Weights& w = e->w;
    QHeadPtrs W;
    const float* src[10] = {w.qe_w[0], w.qe_b[0], w.qe_w[1], w.qe_b[1], w.qo_w[0], w.qo_b[0], w.qo_w[1], w.qo_b[1], w.qo_w[2], w.qo_b[2]};
    for (int t = 0; t < 10; ++t) W.p[t] = const_cast<float*>(src[t]);
    TrainArgs a;
    a.G = n_graphs; a.act = act; a.step = step; a.apply = apply;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm; a.p_drop = p_drop;
    a.head_in = head_in; a.target = target; a.drop_u = drop_u; a.m = m; a.v = v;
    a.loss = loss; a.pred = pred; a.gnorm = gnorm; a.grad_out = grad_out;
    Span sp(e, s, FAM_READOUT);
    q_head_train_kernel<<<1, TR_THREADS, smem, s>>>(W, a);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
