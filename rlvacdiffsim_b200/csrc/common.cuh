// Shared declarations for librlvd_b200.so (sm_100a).  See include/rlvd_b200.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/rlvd_b200.h"

namespace rlvd {

constexpr int F = RLVD_F;          // 128 hidden channels
constexpr int NRBF = RLVD_NRBF;    // 20 Bessel functions
constexpr int RLVD_HEAD_IN = 36;   // DQN-head inputs per reaction graph: z[33], extras[2], base (readout.cu -> train.cu)
constexpr int RLVD_QHEAD_PARAMS = 2513;   // Q_emb [16,33]+16, [16,16]+16; Q_output [32,18]+32, [32,32]+32, [1,32]+1
constexpr int L0_COLS = 256;       // layer-0 species sums per node: 4 blocks [Hx|Hy|Hz|G] of 64 = 3 species x 21 + pad
constexpr int NLAYER = RLVD_NLAYER;
constexpr int GEOM = 24;           // per-edge geometry record of the fp32 SIMT kernel: 20 x phi*fcut, fcut, dir.xyz

void set_error(const char* fmt, ...);

#define RLVD_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t err__ = (call);                                                      \
        if (err__ != cudaSuccess) {                                                      \
            rlvd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                            cudaGetErrorString(err__));                                  \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

#define RLVD_CHECK(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            rlvd::set_error(__VA_ARGS__);                                                \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Tensor {
    void* d = nullptr;
    int64_t numel = 0;
    bool is_int = false;
};

enum Family { FAM_NEIGHBOR = 0, FAM_EDGE = 1, FAM_NODE = 2, FAM_READOUT = 3, FAM_PACK = 4, FAM_COUNT = 5 };

struct ProfSpan {
    int family;
    cudaEvent_t a, b;
};

// Kernel-side view of the checkpoint (all device pointers, fp32 unless noted).
struct Weights {
    const float* emb = nullptr;                 // [100,128]
    const float* freqs = nullptr;               // [20]
    const float* filt_T = nullptr;              // [L][21][384]  (k<20: weight^T, k=20: bias)
    const void* filt_F = nullptr;               // [L][3 parts][hi|lo] tf32 tensor-core operand (tc_edge.cu)
    const void* filt_E = nullptr;               // [L][4 slices][128][32] fp16-split TMEM image (edge_stream.cu)
    float filt_kappa[NLAYER] = {};
    // layer-0 collapse (engine.cu build_layer0): species index per atomic number and the three per-node GEMMs
    const int32_t* z2s = nullptr;               // [100], -1 = not a model species
    const void* l0_wF[3] = {};                  // [Hx|Hy] -> (dmu_x, dmu_y);  [Hz|G] -> dmu_z;  [Hz|G] -> dq
    bool l0_ready = false;              // 2^-(sA+10): undoes the fp16 range scaling of filt_E and the edge tiles
    const float* feps = nullptr;                // [20] freqs[k] - (k+1)*freqs[0]
    const float* inter_w0T[NLAYER] = {};        // [128][128]   (in-major: Wt[k][n])
    const float* inter_b0[NLAYER] = {};
    const float* inter_w3T[NLAYER] = {};        // [128][384]
    const float* inter_b3[NLAYER] = {};
    const float* mix_wT[NLAYER] = {};           // [128][256]
    const float* intra_w0T[NLAYER] = {};        // [256][128]
    const float* intra_b0[NLAYER] = {};
    const float* intra_w3T[NLAYER] = {};        // [128][384]
    const float* intra_b3[NLAYER] = {};
    // the same five linears pre-split into fp16 hi/lo tensor-core stages (tc_linear.cu)
    const void* inter_w0F[NLAYER] = {};
    const void* inter_w3F[NLAYER] = {};
    const void* mix_wF[NLAYER] = {};
    const void* mix_wF_red[NLAYER] = {};        // rows regrouped for the mix-reduce epilogue (tc_format_mix_weight)
    const void* intra_w0F[NLAYER] = {};
    const void* intra_w3F[NLAYER] = {};
    const void* intra_w3F_ac[NLAYER] = {};      // rows [a | c] of layers.3 only (256 x 128): the q-only final mixing block
    const float* intra_b3_ac[NLAYER] = {};
    const float* en_w0T = nullptr;              // [128][64]
    const float* en_b0 = nullptr;
    const float* en_w3 = nullptr;               // [64]
    const float* en_b3 = nullptr;               // [1]
    const float* sp_scales = nullptr;           // [3]
    const float* sp_shifts = nullptr;           // [3]
    const int32_t* elem_lookup = nullptr;       // [100]
    const float* rr_w0T = nullptr;              // [128][128]
    const float* rr_b0 = nullptr;
    const float* rr_w3T = nullptr;              // [128][32]
    const float* rr_b3 = nullptr;
    const float* ro_w[3] = {};                  // reaction_output [32,33] [32,32] [2,32] (out-major as stored)
    const float* ro_b[3] = {};
    const float* qe_w[2] = {};                  // Q_emb [16,33] [16,16]
    const float* qe_b[2] = {};
    const float* qo_w[3] = {};                  // Q_output [32,18] [32,32] [1,32]
    const float* qo_b[3] = {};
    float cutoff = 5.0f;
};

}  // namespace rlvd

struct rlvd_engine {
    int device = 0;
    int sm_count = 0;
    std::map<std::string, rlvd::Tensor> tensors;   // as uploaded, checkpoint layout
    std::vector<void*> derived;                    // transposed copies owned by the engine
    rlvd::Weights w;
    bool finalized = false;
    // workspaces (grown on demand, never shrunk)
    rlvd::DevBuf x, mu0, mu1, hid, mix, nrm, vw, ctx, struct_edges, struct_edge_ptr, scratch;
    rlvd::DevBuf tiles, pipe_info, tile_ctr, stream_scratch;   // edge stream (edge_stream.cu)
    rlvd::DevBuf status;                                       // int[4] device status word (rlvd_check_status)
    rlvd::DevBuf readout_scratch;                              // split-readout partial pools + arrival counters
    rlvd::DevBuf l0A;                                          // [M, 256] species-resolved radial sums of layer 0
    rlvd::DevBuf hit_words;                                    // [M, 8] hit bits of the last count pass (neighbor.cu); the key
    const void* hit_key_pos = nullptr;                         // below names the graph they belong to (consumed by its fill pass)
    const void* hit_key_ptr = nullptr;
    int hit_key_nodes = 0, hit_key_struct = 0;
    float hit_key_rc2 = 0.f;
    bool use_hit_words = true;
    rlvd::DevBuf train_ws;                                     // saved activations + backward scratch (train_full.cu)
    float eps = 1e-8f;                                         // hyper_parameters.reaction_model.epsilon (rlvd_set_hyper)
    int n_scale_species = 3;                                   // entries of species_energy_scale.{scales, shifts} in the checkpoint
    int edge_cap_per_node = 64;                                // CSR capacity bound of rlvd_score_reactions_host (grows on demand)
    int n_species = 0;                                         // rlvd_set_species
    int species_z[3] = {};
    bool use_l0_sums = true;        // layer 0 as per-node GEMMs over species-resolved radial sums (needs rlvd_set_species)
    // host-path staging
    rlvd::DevBuf in_node_ptr, in_Z, in_pos, in_cell, in_inv, in_img, in_rs, in_ps, in_temp, in_qp;
    rlvd::DevBuf g_row_ptr, g_node_struct, g_col, g_shift, g_nedges, g_q, out_pack;
    int64_t launches = 0;
    bool attr_stream = false, attr_neighbor = false, attr_tc_edge = false, attr_tc_linear = false, attr_train = false;   // dynamic-smem opt-ins done on this device
    bool use_tensor_cores = true;
    bool use_sliced_edge = true;    // structure-resident edge kernels when every structure has <= 256 atoms
    bool fuse_mix = true;           // nrm / vw reductions inside the mu_channel_mix GEMM epilogue (tc_linear.cu mixred mode): same
                                    // step time as GEMM + reduction kernel within run-to-run noise, 4.8 GB less DRAM traffic per step
    bool fuse_mlp = true;           // Linear -> SiLU -> Linear pairs as one tcgen05 launch (tc_linear.cu two-layer mode)
    int stream_split = 1;           // CTAs per (structure, channel slice) of the edge stream kernel for the batch packed last (launch_edge_pack)
    int stream_prefetch = 7;        // edge_stream_kernel: L2 prefetch of the node rows the SM's next CTA will stage, issued after this many eighths of a CTA's tiles (0 = off)
    bool prune_final_mu = true;     // callers that take q only (d_mu == NULL): the last mixing block skips everything that
                                    // only feeds mu's update (W store, the b block of the intra-atomic net, mu += b W)
    bool profiling = false;
    std::vector<rlvd::ProfSpan> spans;
    std::vector<cudaEvent_t> event_pool;
    double prof_ms[rlvd::FAM_COUNT] = {};
    int64_t prof_n[rlvd::FAM_COUNT] = {};
};

namespace rlvd {

// RAII-less span helpers: begin/end record events on `stream` when profiling is on.
struct Span {
    rlvd_engine* e;
    cudaStream_t s;
    int idx;
    Span(rlvd_engine* e_, cudaStream_t s_, int family);
    void end(int n_launches);
};

// ---- kernel launchers (each returns 0 / non-zero with set_error) ----
int launch_neighbor_count(rlvd_engine* e, cudaStream_t s, int n_struct, int n_nodes, const int* node_ptr,
                          const float* pos, const float* cell, const float* inv, const int* img, float cutoff,
                          int* row_ptr, int* node_struct, int64_t* n_edges);
int launch_neighbor_fill(rlvd_engine* e, cudaStream_t s, int n_struct, int n_nodes, const int* node_ptr,
                         const float* pos, const float* cell, const float* inv, const int* img, float cutoff,
                         const int* row_ptr, int64_t cap, int* col, int8_t* shift);

int launch_clamp_graph(rlvd_engine* e, cudaStream_t s, int n_nodes, int* row_ptr, const int64_t* n_edges, int64_t cap);

struct LinearArgs {
    const float* A = nullptr;     // [M, K1] rows of length lda
    int lda = 0;
    const float* A2 = nullptr;    // optional second source for k >= K1 (concatenation)
    int lda2 = 0;
    int K1 = 0;
    int M = 0, K = 0, N = 0;
    const float* Wt = nullptr;    // [K][N]           (fp32 SIMT path)
    const void* Wfmt = nullptr;   // tensor-core stages (tc_linear.cu); used when the engine option is on
    const float* bias = nullptr;  // [N] or null
    float* Y = nullptr;
    int ldy = 0;
    int act = 0;                  // 0 none, 1 SiLU
    // optional second layer fused behind the first (tc_linear.cu, N == 128 only): Y = act2(act(A W^T + b) W2^T + b2);
    // Y / ldy then describe the [M, N2] output and the hidden tile never leaves the SM
    const void* W2fmt = nullptr;
    const float* bias2 = nullptr;
    int N2 = 0;
    int act2 = 0;
    // mu_channel_mix with the equivariant reductions fused into its epilogue (tc_linear.cu, K = 128, N = 256, no bias):
    // A = mu as [3*n_nodes, 128] rows (node, component); the epilogue forms nrm = sqrt(sum_c V_c^2 + eps) and
    // vw = sum_c V_c W_c per node and writes only W (Y, [3*n_nodes, 128]) — V never reaches HBM
    int mixred = 0;
    int n_nodes = 0;
    float* nrm = nullptr;
    float* vw = nullptr;
    float eps = 1e-8f;
    // device-side gate: when set, the kernel runs only if *gate == gate_value (layer-0 collapse vs its fallback)
    const int* gate = nullptr;
    int gate_value = 0;
    int* status = nullptr;        // engine status word: bit 0 = an operand left the fp16-split range (tc_linear.cu)
    long long* dbg = nullptr;     // RLVD_TC_TIMELINE builds: (tag, clock64) stamps of block 0 (tc_linear.cu)
};
int launch_linear(rlvd_engine* e, cudaStream_t s, const LinearArgs& a);
int launch_tc_linear(rlvd_engine* e, cudaStream_t s, const LinearArgs& a, const void* Wfmt);
bool tc_linear_supported(const LinearArgs& a);
void tc_format_weight(const float* W, int N, int K, std::vector<uint8_t>& out);
void tc_format_mix_weight(const float* W, std::vector<uint8_t>& out);

int launch_embed(rlvd_engine* e, cudaStream_t s, int M, const int* Z, float* q);
struct EdgeArgs {
    int layer = 0, M = 0, n_struct = 0, max_struct_nodes = 0;
    int sliced = 0;                 // 0 general kernels; 2 edge tiles ready (edge_stream.cu)
    const int* node_ptr = nullptr;
    const int* row_ptr = nullptr;
    const int* col = nullptr;
    const int8_t* shift = nullptr;
    const float* pos = nullptr;
    const float* cell = nullptr;
    const int* node_struct = nullptr;
    const float* x = nullptr;
    const float* mu_in = nullptr;   // null in layer 0 (mu == 0)
    float* q = nullptr;
    float* mu_out = nullptr;
    const int* gate = nullptr;      // stream kernel only: run only if *gate == gate_value
    int gate_value = 0;
};
int launch_edge(rlvd_engine* e, cudaStream_t s, const EdgeArgs& a);
bool edge_stream_supported(int max_struct_nodes);
void edge_stream_format(const float* W, const float* b, std::vector<uint8_t>& out, float* kappa);
int launch_edge_pack(rlvd_engine* e, cudaStream_t s, int n_struct, int M, int64_t n_edges, const int* node_ptr,
                     const int* row_ptr, const int* col, const int8_t* shift, const float* pos, const float* cell,
                     const int* Z, float* l0A);
int launch_q_head_train(rlvd_engine* e, cudaStream_t s, int n_graphs, const float* head_in, const float* target,
                        const float* drop_u, float p_drop, int act, float lr, float beta1, float beta2, float eps,
                        float max_norm, int step, int apply, float* m, float* v, float* loss, float* pred, float* gnorm,
                        float* grad_out);
int launch_axpy(rlvd_engine* e, cudaStream_t s, int64_t n, const float* x, float* y, const int* gate, int gate_value);   // y += x
int launch_edge_stream(rlvd_engine* e, cudaStream_t s, int layer, int n_struct, int max_struct_nodes, const int* node_ptr,
                       const float* x, const float* mu_in, float* q, float* mu_out, const int* gate, int gate_value);
int launch_tc_edge(rlvd_engine* e, cudaStream_t s, int layer, int M, const int* row_ptr, const int* col,
                   const int8_t* shift, const float* pos, const float* cell, const int* node_struct,
                   const float* x, const float* mu_in, float* q, float* mu_out);
void tc_edge_format_filters(const float* W, const float* b, std::vector<uint8_t>& out);
int launch_mix_reduce(rlvd_engine* e, cudaStream_t s, int M, const float* mix, float* nrm, float* vw);
int launch_mix_update_q(rlvd_engine* e, cudaStream_t s, int M, const float* ctx_ac, const float* vw, float* q);
int launch_mix_update(rlvd_engine* e, cudaStream_t s, int M, const float* ctx, const float* W, int ldw,
                      const float* vw, float* q, float* mu);
int launch_readout(rlvd_engine* e, cudaStream_t s, int n_react, const int* node_ptr, const int* rs,
                   const int* ps, const int* Z, const float* q, const rlvd_q_params& qp, const float* d_temp,
                   float* rl_q, float* q0, float* q1, float* delta_e, float* E_R, float* E_P, float* head_in = nullptr);

int launch_time_readout(rlvd_engine* e, cudaStream_t s, int n_struct, const int* node_ptr, const float* q, const float* T,
                        const float* defect, const rlvd_time_head& h, float* time_out);

int launch_action_space(rlvd_engine* e, cudaStream_t s, int n_struct, const int* node_ptr, const double* pos, const double* cell,
                        double a, int max_vac, int max_actions, int* count, double* actions, int* status);

int launch_candidate_offsets(rlvd_engine* e, cudaStream_t s, int n_rep, const int* count, int* first, int* total);
int launch_expand_candidates(rlvd_engine* e, cudaStream_t s, int n_rep, int n_cand, const int* rep_ptr, const int* first, const int* Z,
                             const double* pos, const double* cell, const double* actions, int max_actions, int n_atoms_per,
                             int* node_ptr, int* Zo, float* poso, float* cello, float* invo, int* imgo, int* rs, int* ps, int* owner,
                             const float* inv_in = nullptr, const int* img_in = nullptr);
int launch_action_space_general(rlvd_engine* e, cudaStream_t s, int n_struct, const int* node_ptr, const double* pos, const double* cell,
                                const double* inv_cell, const int* site_ptr, const double* sites, double a, int max_vac, int max_actions,
                                int* count, double* actions, int* status);
int launch_pack_records(rlvd_engine* e, cudaStream_t s, int n_rep, const int* first, const float* rl_q, const int* chosen, const float* E_R,
                        int id0, int id_stride, int max_actions, float* rec);
int launch_select_apply(rlvd_engine* e, cudaStream_t s, int n_rep, const int* rep_ptr, const int* first, const float* rl_q,
                        const float* temperature, const double* u, unsigned long long seed, unsigned long long step, unsigned int replica0,
                        const double* actions, int max_actions, int apply, double* pos,
                        int* chosen, int* argmax_out, float* probs, float* gamma);

int launch_sro_counts(rlvd_engine* e, cudaStream_t s, int n_struct, const int* node_ptr, const double* pos, const double* cell,
                      const int* sid, int n_species, long long* counts, int* status);

}  // namespace rlvd
