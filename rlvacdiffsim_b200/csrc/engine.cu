// C ABI of librlvd_b200.so: engine lifetime, checkpoint upload, and the drivers that chain the kernels of
// neighbor.cu / node.cu / edge.cu / readout.cu into the path the reference runs at
// rlsim/drl/simulator.py:56 (`model(batch, q_params)["rl_q"]`).  See include/rlvd_b200.h.
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace rlvd {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p != nullptr) {
        RLVD_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;  // a little headroom so near-equal batches do not reallocate
    RLVD_CUDA(cudaMalloc(&p, want));
    cap = want;
    return 0;
}

void DevBuf::release() {
    if (p != nullptr) cudaFree(p);
    p = nullptr;
    cap = 0;
}

Span::Span(rlvd_engine* e_, cudaStream_t s_, int family) : e(e_), s(s_), idx(-1) {
    if (!e->profiling || family < 0) return;      // family < 0: launch kept out of the per-family timings
    ProfSpan sp;
    sp.family = family;
    for (cudaEvent_t* ev : {&sp.a, &sp.b}) {
        if (!e->event_pool.empty()) {
            *ev = e->event_pool.back();
            e->event_pool.pop_back();
        } else {
            cudaEventCreate(ev);
        }
    }
    cudaEventRecord(sp.a, s);
    e->spans.push_back(sp);
    idx = (int)e->spans.size() - 1;
}

void Span::end(int n_launches) {
    e->launches += n_launches;
    if (idx >= 0) cudaEventRecord(e->spans[idx].b, s);
}

namespace {

struct HostTensor {
    std::vector<float> f;
    std::vector<int32_t> i;
};

std::map<const rlvd_engine*, std::map<std::string, HostTensor>>& host_store() {
    static std::map<const rlvd_engine*, std::map<std::string, HostTensor>> s;
    return s;
}

int upload(rlvd_engine* e, const void* src, size_t bytes, void** out) {
    void* d = nullptr;
    RLVD_CUDA(cudaMalloc(&d, bytes));
    RLVD_CUDA(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    e->derived.push_back(d);
    *out = d;
    return 0;
}

// Upload W[out][in] transposed to Wt[in][out].
int upload_T(rlvd_engine* e, const std::vector<float>& W, int n_out, int n_in, const float** out) {
    std::vector<float> t((size_t)n_out * n_in);
    for (int o = 0; o < n_out; ++o)
        for (int k = 0; k < n_in; ++k) t[(size_t)k * n_out + o] = W[(size_t)o * n_in + k];
    void* d = nullptr;
    if (upload(e, t.data(), t.size() * sizeof(float), &d)) return 1;
    *out = reinterpret_cast<const float*>(d);
    return 0;
}

int upload_fmt(rlvd_engine* e, const std::vector<float>& W, int n_out, int n_in, const void** out) {
    std::vector<uint8_t> f;
    tc_format_weight(W.data(), n_out, n_in, f);
    void* d = nullptr;
    if (upload(e, f.data(), f.size(), &d)) return 1;
    *out = d;
    return 0;
}

int upload_raw(rlvd_engine* e, const std::vector<float>& W, const float** out);

// Layer 0 of the representation sees q = embedding[Z] and mu = 0, so its per-edge messages
//   dq_i = sum_j Wq(d_ij) * x_q[Z_j],   dmu_i = sum_j W_R(d_ij) * x_R[Z_j] * dir_ij        (x = context_net(embedding[Z]))
// depend on the sender only through its species.  With G[i][s][k] = sum_{j in N(i), species s} g_k(d_ij)  (g = phi*fcut, fcut)
// and H_a likewise weighted by dir_a, they are per-NODE GEMMs  dq = G Mq,  dmu_a = H_a MR  with constant matrices
//   Mq[(s,k)][c] = x_q[s][c] Wf_q[c][k],   MR[(s,k)][c] = x_R[s][c] Wf_R[c][k]             (k = 20: the filter bias).
// The sums come out of the edge pack kernel (edge_stream.cu), the GEMMs run on tc_linear; this builds the matrices.
template <class Store>
int build_layer0(rlvd_engine* e, Store& hs, Weights& w) {
    const std::string rep = "reaction_model.representation.";
    auto get = [&](const std::string& name) -> const std::vector<float>* {
        auto it = hs.find(name);
        return it == hs.end() ? nullptr : &it->second.f;
    };
    const std::vector<float>*emb = get(rep + "embedding.weight"), *fw = get(rep + "filter_net.weight"), *fb = get(rep + "filter_net.bias");
    const std::string pi = rep + "interactions.0.interatomic_context_net.layers.";
    const std::vector<float>*w0 = get(pi + "0.weight"), *b0 = get(pi + "0.bias"), *w3 = get(pi + "3.weight"), *b3 = get(pi + "3.bias");
    RLVD_CHECK(emb && fw && fb && w0 && b0 && w3 && b3, "layer-0 collapse: missing checkpoint tensors");
    const int S = e->n_species;
    std::vector<double> xs((size_t)S * 3 * F);
    for (int sp = 0; sp < S; ++sp) {
        const float* q = emb->data() + (size_t)e->species_z[sp] * F;
        std::vector<double> h(F);
        for (int o = 0; o < F; ++o) {
            double a = (*b0)[o];
            for (int k = 0; k < F; ++k) a += (double)(*w0)[(size_t)o * F + k] * (double)q[k];
            h[o] = a / (1.0 + std::exp(-a));
        }
        for (int o = 0; o < 3 * F; ++o) {
            double a = (*b3)[o];
            for (int k = 0; k < F; ++k) a += (double)(*w3)[(size_t)o * F + k] * h[k];
            xs[(size_t)sp * 3 * F + o] = a;
        }
    }
    // part 0 = q rows [0,F), part 1 = R rows [F,2F) of layer 0's filter and of x
    auto Mval = [&](int part, int sp, int k, int c) -> float {
        const int row = part * F + c;
        const double f = k < NRBF ? (double)(*fw)[(size_t)row * NRBF + k] : (double)(*fb)[row];
        return (float)(xs[(size_t)sp * 3 * F + row] * f);
    };
    const int NKK = NRBF + 1;
    // nn.Linear layout W[n_out][128]; the A operand is two 64-wide blocks, each (species, k) + one pad column
    std::vector<float> W1((size_t)2 * F * F, 0.f), W2((size_t)F * F, 0.f), W3((size_t)F * F, 0.f);
    for (int c = 0; c < F; ++c)
        for (int sp = 0; sp < S; ++sp)
            for (int k = 0; k < NKK; ++k) {
                const int kk = sp * NKK + k;
                const float mr = Mval(1, sp, k, c), mq = Mval(0, sp, k, c);
                W1[(size_t)c * F + kk] = mr;               // [Hx|Hy] -> dmu_x from the Hx block
                W1[(size_t)(F + c) * F + 64 + kk] = mr;    //            dmu_y from the Hy block
                W2[(size_t)c * F + kk] = mr;               // [Hz|G]  -> dmu_z from the Hz block
                W3[(size_t)c * F + 64 + kk] = mq;          // [Hz|G]  -> dq    from the G block
            }
    if (upload_fmt(e, W1, 2 * F, F, &w.l0_wF[0]) || upload_fmt(e, W2, F, F, &w.l0_wF[1]) || upload_fmt(e, W3, F, F, &w.l0_wF[2]))
        return 1;
    std::vector<int32_t> z2s(100, -1);
    for (int sp = 0; sp < S; ++sp) z2s[e->species_z[sp]] = sp;
    void* d = nullptr;
    if (upload(e, z2s.data(), z2s.size() * sizeof(int32_t), &d)) return 1;
    w.z2s = reinterpret_cast<const int32_t*>(d);
    w.l0_ready = true;
    return 0;
}

int upload_raw(rlvd_engine* e, const std::vector<float>& W, const float** out) {
    void* d = nullptr;
    if (upload(e, W.data(), W.size() * sizeof(float), &d)) return 1;
    *out = reinterpret_cast<const float*>(d);
    return 0;
}

}  // namespace

}  // namespace rlvd

using namespace rlvd;

extern "C" {

int rlvd_abi_version(void) { return RLVD_ABI_VERSION; }

const char* rlvd_last_error(void) { return g_error; }

int rlvd_engine_create(int device, rlvd_engine** out) {
    RLVD_CHECK(out != nullptr, "rlvd_engine_create: out is NULL");
    int n_dev = 0;
    RLVD_CUDA(cudaGetDeviceCount(&n_dev));
    RLVD_CHECK(device >= 0 && device < n_dev, "rlvd_engine_create: device %d not in [0,%d)", device, n_dev);
    cudaDeviceProp prop;
    RLVD_CUDA(cudaGetDeviceProperties(&prop, device));
    RLVD_CHECK(prop.major == 10, "rlvd_engine_create: device %d is sm_%d%d; this library is built for sm_100a only",
               device, prop.major, prop.minor);
    RLVD_CUDA(cudaSetDevice(device));
    rlvd_engine* e = new rlvd_engine();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    if (e->status.ensure(16)) { delete e; return 1; }
    RLVD_CUDA(cudaMemset(e->status.p, 0, 16));
    *out = e;
    return 0;
}

void rlvd_engine_destroy(rlvd_engine* e) {
    if (e == nullptr) return;
    cudaSetDevice(e->device);
    for (void* d : e->derived) cudaFree(d);
    for (DevBuf* b : {&e->x, &e->mu0, &e->mu1, &e->hid, &e->mix, &e->nrm, &e->vw, &e->ctx, &e->struct_edges,
                      &e->struct_edge_ptr, &e->scratch, &e->tiles, &e->pipe_info, &e->tile_ctr, &e->stream_scratch, &e->l0A, &e->hit_words, &e->train_ws, &e->readout_scratch, &e->status, &e->in_node_ptr, &e->in_Z, &e->in_pos, &e->in_cell,
                      &e->in_inv, &e->in_img, &e->in_rs, &e->in_ps, &e->in_temp, &e->in_qp, &e->g_row_ptr,
                      &e->g_node_struct, &e->g_col, &e->g_shift, &e->g_nedges, &e->g_q, &e->out_pack})
        b->release();
    for (auto& sp : e->spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    for (cudaEvent_t ev : e->event_pool) cudaEventDestroy(ev);
    host_store().erase(e);
    delete e;
}

int rlvd_set_weight(rlvd_engine* e, const char* name, const void* h_data, int64_t numel, int32_t is_int32) {
    RLVD_CHECK(e != nullptr && name != nullptr && h_data != nullptr && numel > 0, "rlvd_set_weight: bad argument");
    HostTensor& t = host_store()[e][name];
    if (is_int32) {
        t.i.assign(reinterpret_cast<const int32_t*>(h_data), reinterpret_cast<const int32_t*>(h_data) + numel);
        t.f.clear();
    } else {
        t.f.assign(reinterpret_cast<const float*>(h_data), reinterpret_cast<const float*>(h_data) + numel);
        t.i.clear();
    }
    e->finalized = false;
    return 0;
}

int rlvd_finalize_weights(rlvd_engine* e) {
    RLVD_CHECK(e != nullptr, "rlvd_finalize_weights: engine is NULL");
    RLVD_CUDA(cudaSetDevice(e->device));
    auto& hs = host_store()[e];
    auto need = [&](const std::string& name, size_t numel, const std::vector<float>** out) -> int {
        auto it = hs.find(name);
        RLVD_CHECK(it != hs.end(), "missing checkpoint tensor '%s'", name.c_str());
        RLVD_CHECK(it->second.f.size() == numel, "checkpoint tensor '%s' has %zu elements, expected %zu",
                   name.c_str(), it->second.f.size(), numel);
        *out = &it->second.f;
        return 0;
    };
    RLVD_CUDA(cudaDeviceSynchronize());
    for (void* d : e->derived) cudaFree(d);
    e->derived.clear();
    Weights w;
    const std::vector<float>* t = nullptr;
    const std::string rep = "reaction_model.representation.";
    if (need(rep + "cutoff_fn.cutoff", 1, &t)) return 1;
    w.cutoff = (*t)[0];
    if (need(rep + "radial_basis.freqs", NRBF, &t) || upload_raw(e, *t, &w.freqs)) return 1;
    {
        std::vector<float> eps(NRBF);
        for (int k = 0; k < NRBF; ++k) eps[k] = (float)((double)(*t)[k] - (double)(k + 1) * (double)(*t)[0]);
        if (upload_raw(e, eps, &w.feps)) return 1;
    }
    if (need(rep + "embedding.weight", 100 * F, &t) || upload_raw(e, *t, &w.emb)) return 1;
    {
        const std::vector<float>*fw = nullptr, *fb = nullptr;
        if (need(rep + "filter_net.weight", (size_t)NLAYER * 3 * F * NRBF, &fw)) return 1;
        if (need(rep + "filter_net.bias", (size_t)NLAYER * 3 * F, &fb)) return 1;
        std::vector<float> ft((size_t)NLAYER * (NRBF + 1) * 3 * F);
        for (int l = 0; l < NLAYER; ++l)
            for (int o = 0; o < 3 * F; ++o) {
                const int row = l * 3 * F + o;
                for (int k = 0; k < NRBF; ++k)
                    ft[((size_t)l * (NRBF + 1) + k) * 3 * F + o] = (*fw)[(size_t)row * NRBF + k];
                ft[((size_t)l * (NRBF + 1) + NRBF) * 3 * F + o] = (*fb)[row];
            }
        if (upload_raw(e, ft, &w.filt_T)) return 1;
        std::vector<uint8_t> ff;
        tc_edge_format_filters(fw->data(), fb->data(), ff);
        void* dff = nullptr;
        if (upload(e, ff.data(), ff.size(), &dff)) return 1;
        w.filt_F = dff;
        std::vector<uint8_t> fe;
        edge_stream_format(fw->data(), fb->data(), fe, w.filt_kappa);
        void* dfe = nullptr;
        if (upload(e, fe.data(), fe.size(), &dfe)) return 1;
        w.filt_E = dfe;
    }
    for (int l = 0; l < NLAYER; ++l) {
        const std::string pi = rep + "interactions." + std::to_string(l) + ".interatomic_context_net.layers.";
        const std::string pm = rep + "mixing." + std::to_string(l) + ".";
        if (need(pi + "0.weight", F * F, &t) || upload_T(e, *t, F, F, &w.inter_w0T[l]) ||
            upload_fmt(e, *t, F, F, &w.inter_w0F[l]))
            return 1;
        if (need(pi + "0.bias", F, &t) || upload_raw(e, *t, &w.inter_b0[l])) return 1;
        if (need(pi + "3.weight", 3 * F * F, &t) || upload_T(e, *t, 3 * F, F, &w.inter_w3T[l]) ||
            upload_fmt(e, *t, 3 * F, F, &w.inter_w3F[l]))
            return 1;
        if (need(pi + "3.bias", 3 * F, &t) || upload_raw(e, *t, &w.inter_b3[l])) return 1;
        if (need(pm + "mu_channel_mix.weight", 2 * F * F, &t) || upload_T(e, *t, 2 * F, F, &w.mix_wT[l]) ||
            upload_fmt(e, *t, 2 * F, F, &w.mix_wF[l]))
            return 1;
        {
            std::vector<uint8_t> f;
            tc_format_mix_weight(t->data(), f);
            void* d = nullptr;
            if (upload(e, f.data(), f.size(), &d)) return 1;
            w.mix_wF_red[l] = d;
        }
        if (need(pm + "intraatomic_context_net.layers.0.weight", F * 2 * F, &t) ||
            upload_T(e, *t, F, 2 * F, &w.intra_w0T[l]) || upload_fmt(e, *t, F, 2 * F, &w.intra_w0F[l]))
            return 1;
        if (need(pm + "intraatomic_context_net.layers.0.bias", F, &t) || upload_raw(e, *t, &w.intra_b0[l])) return 1;
        if (need(pm + "intraatomic_context_net.layers.3.weight", 3 * F * F, &t) ||
            upload_T(e, *t, 3 * F, F, &w.intra_w3T[l]) || upload_fmt(e, *t, 3 * F, F, &w.intra_w3F[l]))
            return 1;
        {   // rows [a | c] of the same layer (the q-only final mixing block never evaluates b)
            std::vector<float> ac((size_t)2 * F * F);
            std::copy(t->begin(), t->begin() + (size_t)F * F, ac.begin());
            std::copy(t->begin() + (size_t)2 * F * F, t->end(), ac.begin() + (size_t)F * F);
            if (upload_fmt(e, ac, 2 * F, F, &w.intra_w3F_ac[l])) return 1;
        }
        if (need(pm + "intraatomic_context_net.layers.3.bias", 3 * F, &t) || upload_raw(e, *t, &w.intra_b3[l])) return 1;
        {
            std::vector<float> ac((size_t)2 * F);
            std::copy(t->begin(), t->begin() + F, ac.begin());
            std::copy(t->begin() + 2 * F, t->end(), ac.begin() + F);
            if (upload_raw(e, ac, &w.intra_b3_ac[l])) return 1;
        }
    }
    if (e->n_species > 0 && build_layer0(e, hs, w)) return 1;
    const std::string rm = "reaction_model.";
    if (need(rm + "energy_output.layers.0.weight", 64 * F, &t) || upload_T(e, *t, 64, F, &w.en_w0T)) return 1;
    if (need(rm + "energy_output.layers.0.bias", 64, &t) || upload_raw(e, *t, &w.en_b0)) return 1;
    if (need(rm + "energy_output.layers.3.weight", 64, &t) || upload_raw(e, *t, &w.en_w3)) return 1;
    if (need(rm + "energy_output.layers.3.bias", 1, &t) || upload_raw(e, *t, &w.en_b3)) return 1;
    // species_energy_scale: one (scale, shift) per model species — the count comes from the checkpoint, not from this file
    int n_sp = 0;
    {
        auto it = hs.find(rm + "species_energy_scale.scales");
        RLVD_CHECK(it != hs.end() && !it->second.f.empty() && it->second.f.size() <= 64,
                   "missing checkpoint tensor 'reaction_model.species_energy_scale.scales' (1..64 species)");
        n_sp = (int)it->second.f.size();
    }
    if (need(rm + "species_energy_scale.scales", n_sp, &t) || upload_raw(e, *t, &w.sp_scales)) return 1;
    if (need(rm + "species_energy_scale.shifts", n_sp, &t) || upload_raw(e, *t, &w.sp_shifts)) return 1;
    e->n_scale_species = n_sp;
    {
        auto it = hs.find(rm + "species_energy_scale.elem_lookup");
        RLVD_CHECK(it != hs.end() && it->second.i.size() == 100,
                   "missing int32 checkpoint tensor 'reaction_model.species_energy_scale.elem_lookup' [100]");
        for (int32_t v : it->second.i) RLVD_CHECK(v >= 0 && v < n_sp, "elem_lookup entry %d out of range for %d species", v, n_sp);
        void* d = nullptr;
        if (upload(e, it->second.i.data(), 100 * sizeof(int32_t), &d)) return 1;
        w.elem_lookup = reinterpret_cast<const int32_t*>(d);
    }
    if (need(rm + "reaction_representation.layers.0.weight", F * F, &t) || upload_T(e, *t, F, F, &w.rr_w0T)) return 1;
    if (need(rm + "reaction_representation.layers.0.bias", F, &t) || upload_raw(e, *t, &w.rr_b0)) return 1;
    if (need(rm + "reaction_representation.layers.3.weight", 32 * F, &t) || upload_T(e, *t, 32, F, &w.rr_w3T)) return 1;
    if (need(rm + "reaction_representation.layers.3.bias", 32, &t) || upload_raw(e, *t, &w.rr_b3)) return 1;
    const int ro_out[3] = {32, 32, 2}, ro_in[3] = {33, 32, 32};
    const int qo_out[3] = {32, 32, 1}, qo_in[3] = {18, 32, 32};
    const int qe_out[2] = {16, 16}, qe_in[2] = {33, 16};
    for (int i = 0; i < 3; ++i) {
        const std::string li = std::to_string(3 * i);
        if (need(rm + "reaction_output.layers." + li + ".weight", (size_t)ro_out[i] * ro_in[i], &t) ||
            upload_raw(e, *t, &w.ro_w[i]))
            return 1;
        if (need(rm + "reaction_output.layers." + li + ".bias", ro_out[i], &t) || upload_raw(e, *t, &w.ro_b[i])) return 1;
        if (need("Q_output.layers." + li + ".weight", (size_t)qo_out[i] * qo_in[i], &t) ||
            upload_raw(e, *t, &w.qo_w[i]))
            return 1;
        if (need("Q_output.layers." + li + ".bias", qo_out[i], &t) || upload_raw(e, *t, &w.qo_b[i])) return 1;
    }
    for (int i = 0; i < 2; ++i) {
        const std::string li = std::to_string(3 * i);
        if (need("Q_emb.layers." + li + ".weight", (size_t)qe_out[i] * qe_in[i], &t) || upload_raw(e, *t, &w.qe_w[i]))
            return 1;
        if (need("Q_emb.layers." + li + ".bias", qe_out[i], &t) || upload_raw(e, *t, &w.qe_b[i])) return 1;
    }
    e->w = w;
    e->finalized = true;
    return 0;
}

int rlvd_radius_graph_count(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes,
                            const int32_t* d_node_ptr, const float* d_pos, const float* d_cell,
                            const float* d_inv_cell, const int32_t* d_img_range, float cutoff, int32_t* d_row_ptr,
                            int32_t* d_node_struct, int64_t* d_n_edges) {
    RLVD_CHECK(e != nullptr, "engine is NULL");
    RLVD_CHECK(n_struct >= 0 && n_nodes >= 0 && cutoff > 0.f, "rlvd_radius_graph_count: bad sizes");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_neighbor_count(e, (cudaStream_t)stream, n_struct, n_nodes, d_node_ptr, d_pos, d_cell, d_inv_cell,
                                 d_img_range, cutoff, d_row_ptr, d_node_struct, d_n_edges);
}

int rlvd_radius_graph_fill(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes,
                           const int32_t* d_node_ptr, const float* d_pos, const float* d_cell,
                           const float* d_inv_cell, const int32_t* d_img_range, float cutoff,
                           const int32_t* d_row_ptr, int64_t edge_capacity, int32_t* d_col, int8_t* d_shift) {
    RLVD_CHECK(e != nullptr, "engine is NULL");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_neighbor_fill(e, (cudaStream_t)stream, n_struct, n_nodes, d_node_ptr, d_pos, d_cell, d_inv_cell,
                                d_img_range, cutoff, d_row_ptr, edge_capacity, d_col, d_shift);
}

int rlvd_painn_representation(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes, int64_t n_edges,
                              int32_t max_struct_nodes, const int32_t* d_node_ptr, const int32_t* d_Z,
                              const float* d_pos, const float* d_cell, const int32_t* d_node_struct,
                              const int32_t* d_row_ptr, const int32_t* d_col, const int8_t* d_shift, float* d_q,
                              float* d_mu) {
    RLVD_CHECK(e != nullptr && e->finalized, "rlvd_painn_representation: weights are not finalized");
    RLVD_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int M = n_nodes;
    if (M == 0) return 0;
    const size_t MF = (size_t)M * F * sizeof(float);
    if (e->x.ensure(3 * MF) || e->mu0.ensure(3 * MF) || e->mu1.ensure(3 * MF) || e->hid.ensure(MF) ||
        e->mix.ensure(6 * MF) || e->nrm.ensure(MF) || e->vw.ensure(MF) || e->ctx.ensure(3 * MF))
        return 1;
    float* x = e->x.as<float>();
    float* mu_cur = e->mu0.as<float>();
    float* mu_nxt = e->mu1.as<float>();
    float* hid = e->hid.as<float>();
    float* mix = e->mix.as<float>();
    float* nrm = e->nrm.as<float>();
    float* vw = e->vw.as<float>();
    float* ctx = e->ctx.as<float>();
    const Weights& w = e->w;

    int sliced = 0;
    if (e->use_tensor_cores && e->use_sliced_edge && d_node_ptr != nullptr && n_edges > 0) {
        if (edge_stream_supported(max_struct_nodes)) sliced = 2;
    }
    const bool l0 = sliced == 2 && e->use_l0_sums && w.l0_ready;
    if (l0 && e->l0A.ensure((size_t)M * L0_COLS * sizeof(float))) return 1;
    if (sliced == 2 && launch_edge_pack(e, s, n_struct, M, n_edges, d_node_ptr, d_row_ptr, d_col, d_shift, d_pos, d_cell, d_Z,
                                        l0 ? e->l0A.as<float>() : nullptr))
        return 1;
    if (launch_embed(e, s, M, d_Z, d_q)) return 1;
    for (int l = 0; l < NLAYER; ++l) {
        if (l == 0 && l0) {
            // layer 0: messages from per-node GEMMs over the species-resolved radial sums (build_layer0)
            //          gated on the pack kernel's species flag: 0 = every sender is a model species -> GEMMs run;
            //          1 = the stock layer-0 kernels run instead (same results, no host round trip)
            const int* flag = e->tile_ctr.as<int>() + 2;
            LinearArgs f0;
            f0.A = d_q; f0.lda = F; f0.M = M; f0.K = F; f0.N = F; f0.Wt = w.inter_w0T[0]; f0.Wfmt = w.inter_w0F[0]; f0.bias = w.inter_b0[0];
            f0.act = 1; f0.W2fmt = w.inter_w3F[0]; f0.bias2 = w.inter_b3[0]; f0.N2 = 3 * F; f0.act2 = 0; f0.Y = x; f0.ldy = 3 * F;
            f0.gate = flag; f0.gate_value = 1;
            if (launch_linear(e, s, f0)) return 1;
            EdgeArgs fe;
            fe.layer = 0; fe.M = M; fe.sliced = sliced; fe.n_struct = n_struct; fe.max_struct_nodes = max_struct_nodes; fe.node_ptr = d_node_ptr;
            fe.row_ptr = d_row_ptr; fe.col = d_col; fe.shift = d_shift; fe.pos = d_pos; fe.cell = d_cell;
            fe.node_struct = d_node_struct; fe.x = x; fe.mu_in = nullptr; fe.q = d_q; fe.mu_out = mu_nxt;
            fe.gate = flag; fe.gate_value = 1;
            if (launch_edge(e, s, fe)) return 1;
            const float* A = e->l0A.as<float>();
            LinearArgs g1;
            g1.gate = flag;
            g1.A = A; g1.lda = L0_COLS; g1.M = M; g1.K = F; g1.N = 2 * F; g1.Wfmt = w.l0_wF[0]; g1.Y = mu_nxt; g1.ldy = 3 * F;
            if (launch_linear(e, s, g1)) return 1;
            LinearArgs g2;
            g2.gate = flag;
            g2.A = A + F; g2.lda = L0_COLS; g2.M = M; g2.K = F; g2.N = F; g2.Wfmt = w.l0_wF[1]; g2.Y = mu_nxt + 2 * F; g2.ldy = 3 * F;
            if (launch_linear(e, s, g2)) return 1;
            LinearArgs g3;
            g3.gate = flag;
            g3.A = A + F; g3.lda = L0_COLS; g3.M = M; g3.K = F; g3.N = F; g3.Wfmt = w.l0_wF[2]; g3.Y = hid; g3.ldy = F;
            if (launch_linear(e, s, g3)) return 1;
            if (launch_axpy(e, s, (int64_t)M * F, hid, d_q, flag, 0)) return 1;   // (q += dq in the GEMM epilogue measured slower: 0.43 vs 0.18 + 0.16 ms)
        } else {
        // R5 node part: x = Lin3(SiLU(Lin0(q)))   (one launch: the hidden tile stays on the SM)
        LinearArgs a0;
        a0.A = d_q; a0.lda = F; a0.M = M; a0.K = F; a0.N = F; a0.Wt = w.inter_w0T[l]; a0.Wfmt = w.inter_w0F[l]; a0.bias = w.inter_b0[l];
        a0.act = 1;
        if (e->use_tensor_cores && e->fuse_mlp) {
            a0.W2fmt = w.inter_w3F[l]; a0.bias2 = w.inter_b3[l]; a0.N2 = 3 * F; a0.act2 = 0; a0.Y = x; a0.ldy = 3 * F;
            if (launch_linear(e, s, a0)) return 1;
        } else {
            a0.Y = hid; a0.ldy = F;
            if (launch_linear(e, s, a0)) return 1;
            LinearArgs a1;
            a1.A = hid; a1.lda = F; a1.M = M; a1.K = F; a1.N = 3 * F; a1.Wt = w.inter_w3T[l]; a1.Wfmt = w.inter_w3F[l]; a1.bias = w.inter_b3[l];
            a1.Y = x; a1.ldy = 3 * F; a1.act = 0;
            if (launch_linear(e, s, a1)) return 1;
        }
        // R2 + R3 + R5 edge part + K4
        EdgeArgs ea;
        ea.layer = l; ea.M = M; ea.sliced = sliced; ea.n_struct = n_struct; ea.max_struct_nodes = max_struct_nodes; ea.node_ptr = d_node_ptr;
        ea.row_ptr = d_row_ptr; ea.col = d_col; ea.shift = d_shift; ea.pos = d_pos; ea.cell = d_cell;
        ea.node_struct = d_node_struct; ea.x = x; ea.mu_in = l == 0 ? nullptr : mu_cur; ea.q = d_q; ea.mu_out = mu_nxt;
        if (launch_edge(e, s, ea)) return 1;
        }
        float* tmp = mu_cur; mu_cur = mu_nxt; mu_nxt = tmp;
        // R6: mixing
        LinearArgs am;
        am.A = mu_cur; am.lda = F; am.M = 3 * M; am.K = F; am.N = 2 * F; am.Wt = w.mix_wT[l]; am.Wfmt = w.mix_wF[l]; am.bias = nullptr;
        am.act = 0;
        if (l == NLAYER - 1 && d_mu == nullptr && e->prune_final_mu && e->use_tensor_cores && e->fuse_mlp) {
            // Final mixing block of a q-only call: mu's update has no consumer, so W is not stored, the b block of the
            // intra-atomic net is not evaluated and only q += a + c * <V, W> remains.  Same q as the full block.
            am.mixred = 1; am.n_nodes = M; am.nrm = nrm; am.vw = vw; am.Y = nullptr; am.ldy = F; am.eps = e->eps; am.Wfmt = w.mix_wF_red[l];
            if (launch_linear(e, s, am)) return 1;
            LinearArgs b0;
            b0.A = d_q; b0.lda = F; b0.A2 = nrm; b0.lda2 = F; b0.K1 = F; b0.M = M; b0.K = 2 * F; b0.N = F;
            b0.Wt = w.intra_w0T[l]; b0.Wfmt = w.intra_w0F[l]; b0.bias = w.intra_b0[l]; b0.act = 1;
            b0.W2fmt = w.intra_w3F_ac[l]; b0.bias2 = w.intra_b3_ac[l]; b0.N2 = 2 * F; b0.act2 = 0; b0.Y = ctx; b0.ldy = 2 * F;
            if (launch_linear(e, s, b0)) return 1;
            if (launch_mix_update_q(e, s, M, ctx, vw, d_q)) return 1;
            continue;
        }
        const bool fused_mix = e->use_tensor_cores && e->fuse_mix;
        if (fused_mix) {                                  // nrm, vw from the GEMM epilogue; `mix` holds W only, [3M, F]
            am.mixred = 1; am.n_nodes = M; am.nrm = nrm; am.vw = vw; am.Y = mix; am.ldy = F; am.eps = e->eps; am.Wfmt = w.mix_wF_red[l];
            if (launch_linear(e, s, am)) return 1;
        } else {
            am.Y = mix; am.ldy = 2 * F;
            if (launch_linear(e, s, am)) return 1;
            if (launch_mix_reduce(e, s, M, mix, nrm, vw)) return 1;
        }
        LinearArgs b0;
        b0.A = d_q; b0.lda = F; b0.A2 = nrm; b0.lda2 = F; b0.K1 = F; b0.M = M; b0.K = 2 * F; b0.N = F;
        b0.Wt = w.intra_w0T[l]; b0.Wfmt = w.intra_w0F[l]; b0.bias = w.intra_b0[l]; b0.act = 1;
        if (e->use_tensor_cores && e->fuse_mlp) {
            b0.W2fmt = w.intra_w3F[l]; b0.bias2 = w.intra_b3[l]; b0.N2 = 3 * F; b0.act2 = 0; b0.Y = ctx; b0.ldy = 3 * F;
            if (launch_linear(e, s, b0)) return 1;
        } else {
            b0.Y = hid; b0.ldy = F;
            if (launch_linear(e, s, b0)) return 1;
            LinearArgs b1;
            b1.A = hid; b1.lda = F; b1.M = M; b1.K = F; b1.N = 3 * F; b1.Wt = w.intra_w3T[l]; b1.Wfmt = w.intra_w3F[l]; b1.bias = w.intra_b3[l];
            b1.Y = ctx; b1.ldy = 3 * F; b1.act = 0;
            if (launch_linear(e, s, b1)) return 1;
        }
        if (launch_mix_update(e, s, M, ctx, fused_mix ? mix : mix + F, fused_mix ? F : 2 * F, vw, d_q, mu_cur)) return 1;
    }
    if (d_mu != nullptr)
        RLVD_CUDA(cudaMemcpyAsync(d_mu, mu_cur, 3 * MF, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int rlvd_reaction_readout(rlvd_engine* e, void* stream, int32_t n_react, int32_t n_nodes, const int32_t* d_node_ptr,
                          const int32_t* d_react_struct, const int32_t* d_prod_struct, const int32_t* d_Z,
                          const float* d_q, const rlvd_q_params* qp, const float* d_temperature, float* d_rl_q,
                          float* d_q0, float* d_q1, float* d_delta_e, float* d_E_R, float* d_E_P) {
    (void)n_nodes;
    RLVD_CHECK(e != nullptr && e->finalized, "rlvd_reaction_readout: weights are not finalized");
    RLVD_CHECK(qp != nullptr, "rlvd_reaction_readout: q_params is NULL");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_readout(e, (cudaStream_t)stream, n_react, d_node_ptr, d_react_struct, d_prod_struct, d_Z, d_q, *qp,
                          d_temperature, d_rl_q, d_q0, d_q1, d_delta_e, d_E_R, d_E_P);
}

int rlvd_reaction_readout_ex(rlvd_engine* e, void* stream, int32_t n_react, int32_t n_nodes, const int32_t* d_node_ptr,
                             const int32_t* d_react_struct, const int32_t* d_prod_struct, const int32_t* d_Z,
                             const float* d_q, const rlvd_q_params* qp, const float* d_temperature, float* d_rl_q,
                             float* d_q0, float* d_q1, float* d_delta_e, float* d_E_R, float* d_E_P, float* d_head_in) {
    (void)n_nodes;
    RLVD_CHECK(e != nullptr && e->finalized, "rlvd_reaction_readout_ex: weights are not finalized");
    RLVD_CHECK(qp != nullptr, "rlvd_reaction_readout_ex: q_params is NULL");
    RLVD_CHECK(d_head_in == nullptr || qp->dqn != 0, "rlvd_reaction_readout_ex: the DQN-head inputs exist only with q_params.dqn");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_readout(e, (cudaStream_t)stream, n_react, d_node_ptr, d_react_struct, d_prod_struct, d_Z, d_q, *qp,
                          d_temperature, d_rl_q, d_q0, d_q1, d_delta_e, d_E_R, d_E_P, d_head_in);
}

int rlvd_q_head_train_step(rlvd_engine* e, void* stream, int32_t n_graphs, const float* d_head_in, const float* d_target,
                           const float* d_dropout_uniform, float dropout_p, int32_t head_act, float lr, float beta1,
                           float beta2, float eps, float max_norm, int32_t step, int32_t apply, float* d_exp_avg,
                           float* d_exp_avg_sq, float* d_loss, float* d_pred, float* d_grad_norm, float* d_grad) {
    RLVD_CHECK(e != nullptr && e->finalized, "rlvd_q_head_train_step: weights are not finalized");
    RLVD_CHECK(n_graphs > 0 && d_head_in && d_target, "rlvd_q_head_train_step: bad argument");
    RLVD_CHECK(!apply || (d_exp_avg && d_exp_avg_sq && step >= 1), "rlvd_q_head_train_step: Adam state / step missing");
    RLVD_CHECK(dropout_p >= 0.f && dropout_p < 1.f, "rlvd_q_head_train_step: dropout_p out of range");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_q_head_train(e, (cudaStream_t)stream, n_graphs, d_head_in, d_target, d_dropout_uniform, dropout_p, head_act,
                               lr, beta1, beta2, eps, max_norm, step, apply, d_exp_avg, d_exp_avg_sq, d_loss, d_pred,
                               d_grad_norm, d_grad);
}

int rlvd_get_weight(rlvd_engine* e, const char* name, float* h_out, int64_t numel) {
    RLVD_CHECK(e != nullptr && e->finalized && name && h_out, "rlvd_get_weight: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    const Weights& w = e->w;
    struct Ent { const char* name; const float* p; int64_t n; };
    const Ent tab[] = {
        {"Q_emb.layers.0.weight", w.qe_w[0], 16 * 33}, {"Q_emb.layers.0.bias", w.qe_b[0], 16},
        {"Q_emb.layers.3.weight", w.qe_w[1], 16 * 16}, {"Q_emb.layers.3.bias", w.qe_b[1], 16},
        {"Q_output.layers.0.weight", w.qo_w[0], 32 * 18}, {"Q_output.layers.0.bias", w.qo_b[0], 32},
        {"Q_output.layers.3.weight", w.qo_w[1], 32 * 32}, {"Q_output.layers.3.bias", w.qo_b[1], 32},
        {"Q_output.layers.6.weight", w.qo_w[2], 32}, {"Q_output.layers.6.bias", w.qo_b[2], 1}};
    for (const Ent& t : tab)
        if (strcmp(name, t.name) == 0) {
            RLVD_CHECK(numel == t.n, "rlvd_get_weight: '%s' has %lld elements, caller asked for %lld", name, (long long)t.n, (long long)numel);
            RLVD_CUDA(cudaMemcpy(h_out, t.p, (size_t)t.n * sizeof(float), cudaMemcpyDeviceToHost));
            return 0;
        }
    RLVD_CHECK(false, "rlvd_get_weight: '%s' is not a trainable tensor of this build (Q_emb.* / Q_output.* only)", name);
}

int rlvd_action_space(rlvd_engine* e, void* stream, int32_t n_struct, const int32_t* d_node_ptr, const double* d_pos,
                      const double* d_cell, double lattice_parameter, int32_t max_vacancies, int32_t max_actions,
                      int32_t* d_count, double* d_actions, int32_t* d_status) {
    RLVD_CHECK(e != nullptr && d_node_ptr && d_pos && d_cell && d_count && d_actions && d_status, "rlvd_action_space: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_action_space(e, (cudaStream_t)stream, n_struct, d_node_ptr, d_pos, d_cell, lattice_parameter, max_vacancies,
                               max_actions, d_count, d_actions, d_status);
}

int rlvd_candidate_offsets(rlvd_engine* e, void* stream, int32_t n_rep, const int32_t* d_count, int32_t* d_first,
                           int32_t* d_total) {
    RLVD_CHECK(e != nullptr && n_rep > 0 && d_count && d_first && d_total, "rlvd_candidate_offsets: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_candidate_offsets(e, (cudaStream_t)stream, n_rep, d_count, d_first, d_total);
}

int rlvd_expand_candidates(rlvd_engine* e, void* stream, int32_t n_rep, int32_t n_cand, int32_t n_atoms_per,
                           const int32_t* d_rep_ptr, const int32_t* d_first, const int32_t* d_Z, const double* d_pos,
                           const double* d_cell, const double* d_actions, int32_t max_actions, int32_t* d_node_ptr,
                           int32_t* d_Z_out, float* d_pos_out, float* d_cell_out, float* d_inv_out, int32_t* d_img_out,
                           int32_t* d_react_struct, int32_t* d_prod_struct, int32_t* d_owner) {
    RLVD_CHECK(e != nullptr && n_rep > 0 && n_cand >= 0 && n_atoms_per > 0, "rlvd_expand_candidates: bad sizes");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_expand_candidates(e, (cudaStream_t)stream, n_rep, n_cand, d_rep_ptr, d_first, d_Z, d_pos, d_cell, d_actions,
                                    max_actions, n_atoms_per, d_node_ptr, d_Z_out, d_pos_out, d_cell_out, d_inv_out, d_img_out,
                                    d_react_struct, d_prod_struct, d_owner);
}

int rlvd_select_apply(rlvd_engine* e, void* stream, int32_t n_rep, const int32_t* d_rep_ptr, const int32_t* d_first,
                      const float* d_rl_q, const float* d_temperature, const double* d_uniform, const double* d_actions,
                      int32_t max_actions, int32_t apply, double* d_pos, int32_t* d_chosen, int32_t* d_argmax, float* d_probs,
                      float* d_gamma) {
    RLVD_CHECK(e != nullptr && n_rep > 0 && d_first && d_rl_q && d_temperature && d_uniform && d_chosen && d_argmax,
               "rlvd_select_apply: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_select_apply(e, (cudaStream_t)stream, n_rep, d_rep_ptr, d_first, d_rl_q, d_temperature, d_uniform, 0ull, 0ull, 0u,
                               d_actions, max_actions, apply, d_pos, d_chosen, d_argmax, d_probs, d_gamma);
}

int rlvd_select_apply_philox(rlvd_engine* e, void* stream, int32_t n_rep, const int32_t* d_rep_ptr, const int32_t* d_first,
                             const float* d_rl_q, const float* d_temperature, uint64_t seed, uint64_t step, uint32_t replica0,
                             const double* d_actions, int32_t max_actions, int32_t apply, double* d_pos, int32_t* d_chosen,
                             int32_t* d_argmax, float* d_probs, float* d_gamma) {
    RLVD_CHECK(e != nullptr && n_rep > 0 && d_first && d_rl_q && d_temperature && d_chosen && d_argmax,
               "rlvd_select_apply_philox: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_select_apply(e, (cudaStream_t)stream, n_rep, d_rep_ptr, d_first, d_rl_q, d_temperature, nullptr, seed, step, replica0,
                               d_actions, max_actions, apply, d_pos, d_chosen, d_argmax, d_probs, d_gamma);
}

int rlvd_action_space_general(rlvd_engine* e, void* stream, int32_t n_struct, const int32_t* d_node_ptr, const double* d_pos,
                              const double* d_cell, const double* d_inv_cell, const int32_t* d_site_ptr, const double* d_sites,
                              double lattice_parameter, int32_t max_vacancies, int32_t max_actions, int32_t* d_count,
                              double* d_actions, int32_t* d_status) {
    RLVD_CHECK(e != nullptr && d_node_ptr && d_pos && d_cell && d_inv_cell && d_site_ptr && d_sites && d_count && d_actions && d_status,
               "rlvd_action_space_general: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_action_space_general(e, (cudaStream_t)stream, n_struct, d_node_ptr, d_pos, d_cell, d_inv_cell, d_site_ptr, d_sites,
                                       lattice_parameter, max_vacancies, max_actions, d_count, d_actions, d_status);
}

int rlvd_expand_candidates_ex(rlvd_engine* e, void* stream, int32_t n_rep, int32_t n_cand, int32_t n_atoms_per,
                              const int32_t* d_rep_ptr, const int32_t* d_first, const int32_t* d_Z, const double* d_pos,
                              const double* d_cell, const double* d_actions, int32_t max_actions, const float* d_inv_in,
                              const int32_t* d_img_in, int32_t* d_node_ptr, int32_t* d_Z_out, float* d_pos_out, float* d_cell_out,
                              float* d_inv_out, int32_t* d_img_out, int32_t* d_react_struct, int32_t* d_prod_struct, int32_t* d_owner) {
    RLVD_CHECK(e != nullptr && n_rep > 0 && n_cand >= 0 && n_atoms_per > 0, "rlvd_expand_candidates_ex: bad sizes");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_expand_candidates(e, (cudaStream_t)stream, n_rep, n_cand, d_rep_ptr, d_first, d_Z, d_pos, d_cell, d_actions,
                                    max_actions, n_atoms_per, d_node_ptr, d_Z_out, d_pos_out, d_cell_out, d_inv_out, d_img_out,
                                    d_react_struct, d_prod_struct, d_owner, d_inv_in, d_img_in);
}

int rlvd_pack_step_records(rlvd_engine* e, void* stream, int32_t n_rep, const int32_t* d_first, const float* d_rl_q,
                           const int32_t* d_chosen, const float* d_E_R, int32_t replica_id0, int32_t replica_id_stride,
                           int32_t max_actions, float* d_records) {
    RLVD_CHECK(e != nullptr && n_rep > 0 && d_first && d_rl_q && d_records && max_actions > 0, "rlvd_pack_step_records: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_pack_records(e, (cudaStream_t)stream, n_rep, d_first, d_rl_q, d_chosen, d_E_R, replica_id0, replica_id_stride,
                               max_actions, d_records);
}

int rlvd_sro_counts(rlvd_engine* e, void* stream, int32_t n_struct, const int32_t* d_node_ptr, const double* d_pos,
                    const double* d_cell, const int32_t* d_species_index, int32_t n_species, int64_t* d_counts,
                    int32_t* d_status) {
    RLVD_CHECK(e != nullptr && d_node_ptr && d_pos && d_cell && d_species_index && d_counts && d_status, "rlvd_sro_counts: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_sro_counts(e, (cudaStream_t)stream, n_struct, d_node_ptr, d_pos, d_cell, d_species_index, n_species,
                             reinterpret_cast<long long*>(d_counts), d_status);
}

int rlvd_time_readout(rlvd_engine* e, void* stream, int32_t n_struct, const int32_t* d_node_ptr, const float* d_q,
                      const float* d_T, const float* d_defect, const rlvd_time_head* head, float* d_time) {
    RLVD_CHECK(e != nullptr && head != nullptr, "rlvd_time_readout: bad argument");
    RLVD_CHECK(head->d_w0 && head->d_b0 && head->d_w1 && head->d_b1 && head->d_w2 && head->d_b2,
               "rlvd_time_readout: head weights missing");
    RLVD_CUDA(cudaSetDevice(e->device));
    return launch_time_readout(e, (cudaStream_t)stream, n_struct, d_node_ptr, d_q, d_T, d_defect, *head, d_time);
}

int rlvd_check_status(rlvd_engine* e, void* stream, int32_t* flags) {
    RLVD_CHECK(e != nullptr && flags != nullptr, "rlvd_check_status: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    int h = 0;
    RLVD_CUDA(cudaMemcpyAsync(&h, e->status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    RLVD_CUDA(cudaStreamSynchronize(s));
    if (h != 0) RLVD_CUDA(cudaMemsetAsync(e->status.p, 0, sizeof(int), s));
    *flags = h;
    return 0;
}

int rlvd_score_reactions_host(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes,
                              const int32_t* h_node_ptr, const int32_t* h_Z, const float* h_pos, const float* h_cell,
                              const float* h_inv_cell, const int32_t* h_img_range, float cutoff, int32_t n_react,
                              const int32_t* h_react_struct, const int32_t* h_prod_struct, const rlvd_q_params* qp,
                              const float* h_temperature, float* h_rl_q, float* h_q0, float* h_q1, float* h_delta_e,
                              float* h_E_R, float* h_E_P, int64_t* h_n_edges) {
    RLVD_CHECK(e != nullptr && e->finalized, "rlvd_score_reactions_host: weights are not finalized");
    RLVD_CHECK(qp != nullptr && n_struct > 0 && n_nodes > 0 && n_react > 0, "rlvd_score_reactions_host: bad sizes");
    RLVD_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int M = n_nodes;
    if (e->in_node_ptr.ensure(((size_t)n_struct + 1) * 4) || e->in_Z.ensure((size_t)M * 4) ||
        e->in_pos.ensure((size_t)M * 12) || e->in_cell.ensure((size_t)n_struct * 36) ||
        e->in_inv.ensure((size_t)n_struct * 36) || e->in_img.ensure((size_t)n_struct * 12) ||
        e->in_rs.ensure((size_t)n_react * 4) || e->in_ps.ensure((size_t)n_react * 4) ||
        e->in_temp.ensure((size_t)n_react * 4) || e->g_row_ptr.ensure(((size_t)M + 1) * 4) ||
        e->g_node_struct.ensure((size_t)M * 4) || e->g_nedges.ensure(8) || e->g_q.ensure((size_t)M * F * 4) ||
        e->out_pack.ensure((size_t)6 * n_react * 4))
        return 1;
    RLVD_CUDA(cudaMemcpyAsync(e->in_node_ptr.p, h_node_ptr, ((size_t)n_struct + 1) * 4, cudaMemcpyHostToDevice, s));
    RLVD_CUDA(cudaMemcpyAsync(e->in_Z.p, h_Z, (size_t)M * 4, cudaMemcpyHostToDevice, s));
    RLVD_CUDA(cudaMemcpyAsync(e->in_pos.p, h_pos, (size_t)M * 12, cudaMemcpyHostToDevice, s));
    RLVD_CUDA(cudaMemcpyAsync(e->in_cell.p, h_cell, (size_t)n_struct * 36, cudaMemcpyHostToDevice, s));
    RLVD_CUDA(cudaMemcpyAsync(e->in_inv.p, h_inv_cell, (size_t)n_struct * 36, cudaMemcpyHostToDevice, s));
    RLVD_CUDA(cudaMemcpyAsync(e->in_img.p, h_img_range, (size_t)n_struct * 12, cudaMemcpyHostToDevice, s));
    RLVD_CUDA(cudaMemcpyAsync(e->in_rs.p, h_react_struct, (size_t)n_react * 4, cudaMemcpyHostToDevice, s));
    RLVD_CUDA(cudaMemcpyAsync(e->in_ps.p, h_prod_struct, (size_t)n_react * 4, cudaMemcpyHostToDevice, s));
    if (h_temperature != nullptr)
        RLVD_CUDA(cudaMemcpyAsync(e->in_temp.p, h_temperature, (size_t)n_react * 4, cudaMemcpyHostToDevice, s));
    if (launch_neighbor_count(e, s, n_struct, M, e->in_node_ptr.as<int>(), e->in_pos.as<float>(),
                              e->in_cell.as<float>(), e->in_inv.as<float>(), e->in_img.as<int>(), cutoff,
                              e->g_row_ptr.as<int>(), e->g_node_struct.as<int>(), e->g_nedges.as<int64_t>()))
        return 1;
    // No host round trip for the edge count: the CSR is filled into a buffer sized for edge_cap edges (64 per node — FCC at
    // a = 3.528 has 54 inside 5.0 A; a denser batch is detected after the final synchronisation and the call repeated with
    // a larger bound) and the true count comes back with the results.
    const int64_t edge_cap = std::min<int64_t>((int64_t)0x7ffffff0, std::max<int64_t>(e->edge_cap_per_node, 8) * (int64_t)M);
    int64_t E = edge_cap;
    if (e->g_col.ensure((size_t)(E + 1) * 4) || e->g_shift.ensure((size_t)(E + 1) * 3)) return 1;
    if (launch_clamp_graph(e, s, M, e->g_row_ptr.as<int>(), e->g_nedges.as<int64_t>(), edge_cap)) return 1;
    if (launch_neighbor_fill(e, s, n_struct, M, e->in_node_ptr.as<int>(), e->in_pos.as<float>(),
                             e->in_cell.as<float>(), e->in_inv.as<float>(), e->in_img.as<int>(), cutoff,
                             e->g_row_ptr.as<int>(), E, e->g_col.as<int>(), e->g_shift.as<int8_t>()))
        return 1;
    int max_nodes = 0;
    for (int g = 0; g < n_struct; ++g) max_nodes = std::max(max_nodes, h_node_ptr[g + 1] - h_node_ptr[g]);
    if (rlvd_painn_representation(e, stream, n_struct, M, E, max_nodes, e->in_node_ptr.as<int>(), e->in_Z.as<int>(),
                                  e->in_pos.as<float>(),
                                  e->in_cell.as<float>(), e->g_node_struct.as<int>(), e->g_row_ptr.as<int>(),
                                  e->g_col.as<int>(), e->g_shift.as<int8_t>(), e->g_q.as<float>(), nullptr))
        return 1;
    float* op = e->out_pack.as<float>();
    if (launch_readout(e, s, n_react, e->in_node_ptr.as<int>(), e->in_rs.as<int>(), e->in_ps.as<int>(),
                       e->in_Z.as<int>(), e->g_q.as<float>(), *qp,
                       h_temperature != nullptr ? e->in_temp.as<float>() : nullptr, op, op + n_react,
                       op + 2 * (size_t)n_react, op + 3 * (size_t)n_react, op + 4 * (size_t)n_react,
                       op + 5 * (size_t)n_react))
        return 1;
    float* hosts[6] = {h_rl_q, h_q0, h_q1, h_delta_e, h_E_R, h_E_P};
    for (int k = 0; k < 6; ++k)
        if (hosts[k] != nullptr)
            RLVD_CUDA(cudaMemcpyAsync(hosts[k], op + (size_t)k * n_react, (size_t)n_react * 4, cudaMemcpyDeviceToHost, s));
    int status = 0;
    int64_t E_true = 0;
    RLVD_CUDA(cudaMemcpyAsync(&status, e->status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    RLVD_CUDA(cudaMemcpyAsync(&E_true, e->g_nedges.p, 8, cudaMemcpyDeviceToHost, s));
    RLVD_CUDA(cudaStreamSynchronize(s));
    RLVD_CHECK(E_true >= 0 && E_true < (int64_t)0x7fffffff, "edge count %lld overflows int32 offsets; split the batch", (long long)E_true);
    if (h_n_edges != nullptr) *h_n_edges = E_true;
    if (E_true > edge_cap) {
        // denser than 64 neighbours per atom: the CSR was truncated — repeat with a bound that fits
        RLVD_CUDA(cudaMemsetAsync(e->status.p, 0, sizeof(int), s));
        e->edge_cap_per_node = (int)((E_true + M - 1) / M) + 8;
        return rlvd_score_reactions_host(e, stream, n_struct, n_nodes, h_node_ptr, h_Z, h_pos, h_cell, h_inv_cell, h_img_range, cutoff,
                                         n_react, h_react_struct, h_prod_struct, qp, h_temperature, h_rl_q, h_q0, h_q1, h_delta_e, h_E_R,
                                         h_E_P, h_n_edges);
    }
    if (status & 2) {
        RLVD_CUDA(cudaMemsetAsync(e->status.p, 0, sizeof(int), s));
        RLVD_CHECK(false, "edge stream: the tile pool overflowed its capacity bound (n_edges / 8 + n_nodes) / 4 + 6 n_struct; results discarded");
    }
    if ((status & 1) && e->use_tensor_cores) {
        // an activation left the fp16-split range of the tensor-core GEMMs (|x| >= 65504: the network diverging on an
        // unphysical structure): redo the call on the fp32 SIMT kernels, which carry any finite fp32 value
        RLVD_CUDA(cudaMemsetAsync(e->status.p, 0, sizeof(int), s));
        e->use_tensor_cores = false;
        const int rc = rlvd_score_reactions_host(e, stream, n_struct, n_nodes, h_node_ptr, h_Z, h_pos, h_cell, h_inv_cell,
                                                 h_img_range, cutoff, n_react, h_react_struct, h_prod_struct, qp,
                                                 h_temperature, h_rl_q, h_q0, h_q1, h_delta_e, h_E_R, h_E_P, h_n_edges);
        e->use_tensor_cores = true;
        return rc;
    }
    return 0;
}

int rlvd_set_species(rlvd_engine* e, const int32_t* atomic_numbers, int32_t n) {
    RLVD_CHECK(e != nullptr && n >= 0 && (n == 0 || atomic_numbers != nullptr), "rlvd_set_species: bad argument");
    RLVD_CHECK(!e->finalized, "rlvd_set_species: call before rlvd_finalize_weights");
    e->n_species = 0;
    if (n == 0 || n > 3) return 0;                    // the layer-0 collapse packs 3 species x 21 radial terms into K = 64
    for (int i = 0; i < n; ++i) {
        RLVD_CHECK(atomic_numbers[i] >= 0 && atomic_numbers[i] < 100, "rlvd_set_species: atomic number %d out of range", atomic_numbers[i]);
        e->species_z[i] = atomic_numbers[i];
    }
    e->n_species = n;
    return 0;
}

int rlvd_set_hyper(rlvd_engine* e, const char* name, double value) {
    RLVD_CHECK(e != nullptr && name != nullptr, "bad argument");
    if (strcmp(name, "epsilon") == 0) {
        RLVD_CHECK(value > 0.0 && value < 1.0, "rlvd_set_hyper: epsilon %g out of range", value);
        e->eps = (float)value;
        return 0;
    }
    RLVD_CHECK(false, "unknown hyper-parameter '%s'", name);
}

int rlvd_set_option(rlvd_engine* e, const char* name, int32_t value) {
    RLVD_CHECK(e != nullptr && name != nullptr, "bad argument");
    if (strcmp(name, "layer0_sums") == 0) {
        e->use_l0_sums = value != 0;
        return 0;
    }
    if (strcmp(name, "tensor_cores") == 0) {
        e->use_tensor_cores = value != 0;
        return 0;
    }
    if (strcmp(name, "sliced_edge") == 0) {
        e->use_sliced_edge = value != 0;
        return 0;
    }
    if (strcmp(name, "fuse_mix") == 0) {
        e->fuse_mix = value != 0;
        return 0;
    }
    if (strcmp(name, "fuse_mlp") == 0) {
        e->fuse_mlp = value != 0;
        return 0;
    }
    if (strcmp(name, "hit_words") == 0) {
        e->use_hit_words = value != 0;
        e->hit_key_pos = nullptr;
        return 0;
    }
    if (strcmp(name, "stream_prefetch") == 0) {
        RLVD_CHECK(value >= 0 && value <= 7, "stream_prefetch: 0 (off) .. 7 eighths");
        e->stream_prefetch = value;
        return 0;
    }
    if (strcmp(name, "prune_final_mu") == 0) {
        e->prune_final_mu = value != 0;
        return 0;
    }
    RLVD_CHECK(false, "unknown option '%s'", name);
}

int rlvd_linear(rlvd_engine* e, void* stream, int32_t M, int32_t K, int32_t N, const float* d_A, const float* d_A2,
                int32_t K1, const float* h_W, const float* h_bias, int32_t act, float* d_Y) {
    RLVD_CHECK(e != nullptr && d_A != nullptr && h_W != nullptr && d_Y != nullptr, "rlvd_linear: bad argument");
    RLVD_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<float> W(h_W, h_W + (size_t)N * K);
    LinearArgs a;
    a.A = d_A; a.lda = d_A2 != nullptr ? K1 : K; a.A2 = d_A2; a.lda2 = K - K1; a.K1 = d_A2 != nullptr ? K1 : 0;
    a.M = M; a.K = K; a.N = N; a.Y = d_Y; a.ldy = N; a.act = act;
    void* d_bias = nullptr;
    void* d_w = nullptr;
    if (h_bias != nullptr) {
        RLVD_CUDA(cudaMalloc(&d_bias, (size_t)N * 4));
        RLVD_CUDA(cudaMemcpy(d_bias, h_bias, (size_t)N * 4, cudaMemcpyHostToDevice));
        a.bias = reinterpret_cast<const float*>(d_bias);
    }
    int rc = 0;
    if (e->use_tensor_cores && tc_linear_supported(a)) {
        std::vector<uint8_t> f;
        tc_format_weight(W.data(), N, K, f);
        RLVD_CUDA(cudaMalloc(&d_w, f.size()));
        RLVD_CUDA(cudaMemcpy(d_w, f.data(), f.size(), cudaMemcpyHostToDevice));
        a.Wfmt = d_w;
        rc = launch_linear(e, s, a);
    } else {
        std::vector<float> t((size_t)N * K);
        for (int o = 0; o < N; ++o)
            for (int k = 0; k < K; ++k) t[(size_t)k * N + o] = W[(size_t)o * K + k];
        RLVD_CUDA(cudaMalloc(&d_w, t.size() * 4));
        RLVD_CUDA(cudaMemcpy(d_w, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
        a.Wt = reinterpret_cast<const float*>(d_w);
        rc = launch_linear(e, s, a);
    }
    cudaError_t err = cudaStreamSynchronize(s);
    cudaFree(d_w);
    if (d_bias != nullptr) cudaFree(d_bias);
    if (rc != 0) return rc;
    RLVD_CHECK(err == cudaSuccess, "rlvd_linear: %s", cudaGetErrorString(err));
    return 0;
}

int64_t rlvd_launch_count(const rlvd_engine* e) { return e != nullptr ? e->launches : 0; }

int rlvd_profile_enable(rlvd_engine* e, int32_t on) {
    RLVD_CHECK(e != nullptr, "engine is NULL");
    e->profiling = on != 0;
    return 0;
}

static int drain_spans(rlvd_engine* e) {
    if (e->spans.empty()) return 0;
    RLVD_CUDA(cudaSetDevice(e->device));
    RLVD_CUDA(cudaDeviceSynchronize());
    for (auto& sp : e->spans) {
        float ms = 0.f;
        RLVD_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
        e->prof_ms[sp.family] += ms;
        e->prof_n[sp.family] += 1;
        e->event_pool.push_back(sp.a);
        e->event_pool.push_back(sp.b);
    }
    e->spans.clear();
    return 0;
}

int rlvd_profile_read(rlvd_engine* e, const char* family, double* ms_total, int64_t* launches) {
    RLVD_CHECK(e != nullptr && family != nullptr, "bad argument");
    if (drain_spans(e)) return 1;
    static const char* names[FAM_COUNT] = {"neighbor", "edge", "node", "readout", "edge_pack"};
    for (int f = 0; f < FAM_COUNT; ++f)
        if (strcmp(names[f], family) == 0) {
            if (ms_total) *ms_total = e->prof_ms[f];
            if (launches) *launches = e->prof_n[f];
            return 0;
        }
    RLVD_CHECK(false, "unknown kernel family '%s'", family);
}

int rlvd_profile_reset(rlvd_engine* e) {
    RLVD_CHECK(e != nullptr, "engine is NULL");
    if (drain_spans(e)) return 1;
    for (int f = 0; f < FAM_COUNT; ++f) {
        e->prof_ms[f] = 0.0;
        e->prof_n[f] = 0;
    }
    return 0;
}

}  // extern "C"
