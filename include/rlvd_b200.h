/*
 * rlvd_b200.h — C ABI of the B200-native candidate-hop scoring path of RLVacDiffSim.
 *
 * One shared library (librlvd_b200.so, sm_100a only) replaces the device work the reference reaches
 * through the un-vendored `rgnn` package at
 *     rlsim/drl/simulator.py:56     rl_q = self.model(batch, q_params=self.q_params)["rl_q"]
 *     rlsim/drl/trainer.py:112,166  model(batch, q_params=self.q_params)["rl_q"]
 *     rlsim/cli/scripts/estimate_time.py:58   model(batch, inference=True)["time"]
 * All pointers are plain; `d_` arguments are device pointers on the engine's GPU, `h_` arguments are host
 * pointers.  `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Every entry point
 * returns 0 on success, non-zero on failure with the message available from rlvd_last_error().
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef RLVD_B200_H
#define RLVD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLVD_ABI_VERSION 2
#define RLVD_F 128          /* hidden channels of the bundled painn (hyper_parameters.hidden_channels) */
#define RLVD_NRBF 20        /* n_rbf */
#define RLVD_NLAYER 3       /* n_interactions */

typedef struct rlvd_engine rlvd_engine;

/* activation codes for the head MLPs (HeadSpec in oracle/painn_oracle.py) */
enum { RLVD_ACT_SILU = 0, RLVD_ACT_RELU = 1, RLVD_ACT_TANH = 2, RLVD_ACT_LEAKY_RELU = 3 };
/* the two extra Q_output inputs */
enum { RLVD_QX_KT_X10 = 0, RLVD_QX_DELTA_E = 1 };

/* q_params of `model(batch, q_params)` (simulator.py:46, configs/*.toml [model.params]) plus the
 * swappable head assumptions of SURVEY.md Appendix C.7. */
typedef struct rlvd_q_params {
    float alpha;
    float beta;
    int32_t dqn;                 /* bool */
    float temperature;           /* K; used when d_temperature == NULL */
    int32_t delta_e_standardised;/* 0: 33rd head input is delta_e in eV; 1: (delta_e-mean)/std */
    int32_t painn_head_act;      /* RLVD_ACT_* for energy_output / reaction_representation / reaction_output */
    int32_t dqn_head_act;        /* RLVD_ACT_* for Q_emb / Q_output */
    int32_t q_extra[2];          /* RLVD_QX_* */
    float mean_barrier, std_barrier, mean_freq, std_freq, mean_delta_e, std_delta_e;
} rlvd_q_params;

/* ---- lifetime ---------------------------------------------------------------------------------- */
int  rlvd_abi_version(void);
const char* rlvd_last_error(void);
/* Creates an engine bound to CUDA device `device`.  Fails if the device is not compute capability 10.x. */
int  rlvd_engine_create(int device, rlvd_engine** out);
void rlvd_engine_destroy(rlvd_engine* e);

/* ---- weights: `registry.get_model_class("dqn_v2").load(path)` (rlsim/drl/deploy.py:18) ----------- */
/* Uploads one fp32 tensor of the checkpoint's state_dict under its checkpoint name, e.g.
 * "reaction_model.representation.filter_net.weight".  int64 `elem_lookup` is passed as int32. */
int  rlvd_set_weight(rlvd_engine* e, const char* name, const void* h_data, int64_t numel, int32_t is_int32);
/* Checks that every tensor of SURVEY.md Appendix A is present with the right size and builds the
 * kernel-side layouts (transposed weights, per-layer filter slices). */
int  rlvd_finalize_weights(rlvd_engine* e);

/* ---- K1: periodic radius graph → receiver-sorted CSR (replaces torch_cluster.radius_graph + minimum
 *      image inside rgnn) ------------------------------------------------------- */
/* n_struct disjoint structures; d_node_ptr[n_struct+1]; d_pos[M,3]; d_cell/d_inv_cell[n_struct,3,3]
 * (row-vector cells, inverse = fp32 image of the fp64 inverse); d_img_range[n_struct,3] extra periodic
 * images to scan per axis (0 for cells with heights >= 2*cutoff).
 * Phase 1 writes d_row_ptr[M+1] (global edge offsets), d_node_struct[M] and the total edge count to
 * d_n_edges[0] (int64).  Phase 2 fills d_col[E] (global sender index, ascending within a row) and
 * d_shift[E,3] (int8).  Edge order is (receiver, sender, Sx, Sy, Sz) ascending. */
int  rlvd_radius_graph_count(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes,
                             const int32_t* d_node_ptr, const float* d_pos, const float* d_cell,
                             const float* d_inv_cell, const int32_t* d_img_range, float cutoff,
                             int32_t* d_row_ptr, int32_t* d_node_struct, int64_t* d_n_edges);
int  rlvd_radius_graph_fill(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes,
                            const int32_t* d_node_ptr, const float* d_pos, const float* d_cell,
                            const float* d_inv_cell, const int32_t* d_img_range, float cutoff,
                            const int32_t* d_row_ptr, int64_t edge_capacity,
                            int32_t* d_col, int8_t* d_shift);

/* ---- K2..K4: PaiNN representation (R2..R6) on a disjoint-union batch ---------------------------- */
/* d_Z[M] atomic numbers (int32).  Output d_q[M,128] fp32 scalar features after 3 interaction+mixing
 * blocks.  d_mu (optional, may be NULL) receives the vector features [M,3,128].  max_struct_nodes is the
 * largest atom count of any structure (host-known from node_ptr); when it is <= 256 the edge stage runs the
 * structure-resident sliced kernel, otherwise the general gather kernel. */
int  rlvd_painn_representation(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes,
                               int64_t n_edges, int32_t max_struct_nodes, const int32_t* d_node_ptr,
                               const int32_t* d_Z, const float* d_pos,
                               const float* d_cell, const int32_t* d_node_struct,
                               const int32_t* d_row_ptr, const int32_t* d_col, const int8_t* d_shift,
                               float* d_q, float* d_mu);

/* ---- K5: per-action readout (H1..H4) ------------------------------------------------------------ */
/* n_react reaction graphs; graph g pairs the structure d_react_struct[g] (reactant) with
 * d_prod_struct[g] (product) of the batch the representation ran on (same atom count, same species).
 * Outputs (each [n_react], any may be NULL): rl_q, q0, q1, delta_e, E_R, E_P. */
int  rlvd_reaction_readout(rlvd_engine* e, void* stream, int32_t n_react, int32_t n_nodes,
                           const int32_t* d_node_ptr, const int32_t* d_react_struct,
                           const int32_t* d_prod_struct, const int32_t* d_Z, const float* d_q,
                           const rlvd_q_params* qp, const float* d_temperature,
                           float* d_rl_q, float* d_q0, float* d_q1, float* d_delta_e,
                           float* d_E_R, float* d_E_P);

/* ---- A1: candidate vacancy-hop enumeration (replaces rlsim/actions/action.py:12-61 `get_action_space`, which the
 *      reference runs on the host with ASE + a dense atoms x sites distance matrix per step) --------------------- */
/* n_struct replicas with orthorhombic cells (off-diagonal terms exactly zero); d_pos[M,3] and d_cell[n_struct,3,3] are
 * fp64 like ASE's.  Row w of replica g: d_actions[(g*max_actions + w)*4 + {0..3}] = {atom index, dx, dy, dz}, rows in
 * the reference's order (vacancies ascending, neighbour sites ascending); d_count[g] rows are valid.
 * d_status[g]: 0 ok, 1 cell not orthorhombic or more than 8192 sites (use the host enumerator), 2 vacancy count not
 * in 1..max_vacancies (the reference's `assert len(vacant) == 1`, action.py:35), 3 more than max_actions rows. */
int  rlvd_action_space(rlvd_engine* e, void* stream, int32_t n_struct, const int32_t* d_node_ptr, const double* d_pos,
                       const double* d_cell, double lattice_parameter, int32_t max_vacancies, int32_t max_actions,
                       int32_t* d_count, double* d_actions, int32_t* d_status);

/* ---- A2/A3/A5 on the device: a whole LSS step without a host round trip -------------------- */
/* Exclusive scan of rlvd_action_space's counts: d_first[n_rep+1], d_total[0] = number of candidates. */
int  rlvd_candidate_offsets(rlvd_engine* e, void* stream, int32_t n_rep, const int32_t* d_count, int32_t* d_first,
                            int32_t* d_total);
/* convert_to_graph_list + collate (rlsim/drl/simulator.py:439-460, :50-55) for every candidate of every replica:
 * reaction graph g = d_first[r] + c gets reactant structure 2g (the replica) and product structure 2g+1
 * (pos[atom] += disp in fp64, then fp32).  Replicas have n_atoms_per atoms each and orthorhombic cells with heights
 * >= 2*cutoff.  Outputs are the inputs of rlvd_radius_graph_* / rlvd_painn_representation / rlvd_reaction_readout. */
int  rlvd_expand_candidates(rlvd_engine* e, void* stream, int32_t n_rep, int32_t n_cand, int32_t n_atoms_per,
                            const int32_t* d_rep_ptr, const int32_t* d_first, const int32_t* d_Z, const double* d_pos,
                            const double* d_cell, const double* d_actions, int32_t max_actions, int32_t* d_node_ptr,
                            int32_t* d_Z_out, float* d_pos_out, float* d_cell_out, float* d_inv_out, int32_t* d_img_out,
                            int32_t* d_react_struct, int32_t* d_prod_struct, int32_t* d_owner);
/* simulator.py:94-98 for every replica: p = softmax(Q / (T * 8.617e-5)) in fp32, np.random.choice's inverse-cdf draw
 * with the caller's uniforms d_uniform[n_rep], argmax, and (apply != 0) pos[atom] += disp of the chosen hop in place
 * (environment.step's position update).  d_probs[n_rep, max_actions] may be NULL.  d_gamma[n_rep] (may be NULL) receives
 * the TKS total rate sum(exp(Q / kT)) of simulator.py:262 (fp32, as torch.sum(torch.exp(Q / kT))). */
int  rlvd_select_apply(rlvd_engine* e, void* stream, int32_t n_rep, const int32_t* d_rep_ptr, const int32_t* d_first,
                       const float* d_rl_q, const float* d_temperature, const double* d_uniform, const double* d_actions,
                       int32_t max_actions, int32_t apply, double* d_pos, int32_t* d_chosen, int32_t* d_argmax, float* d_probs,
                       float* d_gamma);

/* The same with a counter-based generator instead of caller uniforms: replica r of this call draws
 * u = Philox4x32-10(key = seed, counter = (replica0 + r, step)) — a 10k-step trajectory needs no host RNG traffic and
 * replays exactly whatever the sharding.  rlvd_select_apply with np.random's uniforms stays the reference-compatible mode. */
int  rlvd_select_apply_philox(rlvd_engine* e, void* stream, int32_t n_rep, const int32_t* d_rep_ptr, const int32_t* d_first,
                              const float* d_rl_q, const float* d_temperature, uint64_t seed, uint64_t step, uint32_t replica0,
                              const double* d_actions, int32_t max_actions, int32_t apply, double* d_pos, int32_t* d_chosen,
                              int32_t* d_argmax, float* d_probs, float* d_gamma);
/* rlvd_action_space for ANY periodic cell (the reference's demo/poscars/POSCAR_0 is rhombohedral): minimum image over the
 * 27 neighbouring images like ase.geometry.find_mic.  The perfect-lattice sites depend only on the cell: d_sites[d_site_ptr[g]
 * .. d_site_ptr[g+1]) (fp64 [n,3], reference site order) and d_inv_cell[g] (fp64 inverse) are computed once by the caller.
 * Status codes as rlvd_action_space (1: no sites / more than 8192). */
int  rlvd_action_space_general(rlvd_engine* e, void* stream, int32_t n_struct, const int32_t* d_node_ptr, const double* d_pos,
                               const double* d_cell, const double* d_inv_cell, const int32_t* d_site_ptr, const double* d_sites,
                               double lattice_parameter, int32_t max_vacancies, int32_t max_actions, int32_t* d_count,
                               double* d_actions, int32_t* d_status);
/* rlvd_expand_candidates for general cells: d_inv_in[n_rep,3,3] (fp32 image of the fp64 inverse) and d_img_in[n_rep,3]
 * (periodic image ranges) of the replicas are copied to their candidates' structures; NULL = orthorhombic, heights >= 2 rc. */
int  rlvd_expand_candidates_ex(rlvd_engine* e, void* stream, int32_t n_rep, int32_t n_cand, int32_t n_atoms_per,
                               const int32_t* d_rep_ptr, const int32_t* d_first, const int32_t* d_Z, const double* d_pos,
                               const double* d_cell, const double* d_actions, int32_t max_actions, const float* d_inv_in,
                               const int32_t* d_img_in, int32_t* d_node_ptr, int32_t* d_Z_out, float* d_pos_out, float* d_cell_out,
                               float* d_inv_out, int32_t* d_img_out, int32_t* d_react_struct, int32_t* d_prod_struct, int32_t* d_owner);
/* One fixed-width fp32 record per replica for the per-step gather of SURVEY.md §8e (rank-0 logger / trajectory writer):
 * d_records[n_rep, 4 + max_actions] = {replica id (replica_id0 + r * stride), n_actions, chosen, E_R, Q[...] NaN-padded}. */
int  rlvd_pack_step_records(rlvd_engine* e, void* stream, int32_t n_rep, const int32_t* d_first, const float* d_rl_q,
                            const int32_t* d_chosen, const float* d_E_R, int32_t replica_id0, int32_t replica_id_stride,
                            int32_t max_actions, float* d_records);

/* ---- replay-minibatch update of the DQN heads (rlsim/drl/trainer.py:163-177 with train_all = False, trainer.py:36-39) ---
 * rlvd_reaction_readout_ex = rlvd_reaction_readout that also writes d_head_in[n_react, 36]: what the trainable heads see
 * per reaction graph (33 Q_emb inputs, the 2 extra Q_output inputs, the alpha/beta part of rl_q).  Needs q_params.dqn. */
int  rlvd_reaction_readout_ex(rlvd_engine* e, void* stream, int32_t n_react, int32_t n_nodes, const int32_t* d_node_ptr,
                              const int32_t* d_react_struct, const int32_t* d_prod_struct, const int32_t* d_Z,
                              const float* d_q, const rlvd_q_params* qp, const float* d_temperature, float* d_rl_q,
                              float* d_q0, float* d_q1, float* d_delta_e, float* d_E_R, float* d_E_P, float* d_head_in);
/* q_pred = heads(head_in); loss = mean((target - q_pred)^2); backward; clip_grad_norm_(max_norm); Adam step (torch.optim.Adam
 * defaults) on the engine's Q_emb / Q_output weights IN PLACE.  d_exp_avg / d_exp_avg_sq: Adam state [2513] (zero before
 * step 1), packed Q_emb.layers.{0,3}.{weight,bias}, Q_output.layers.{0,3,6}.{weight,bias}; step is 1-based.
 * d_dropout_uniform[n_graphs, 80] (may be NULL = no dropout) drives the heads' train-mode dropout (keep if u >= p).
 * apply = 0 only evaluates (loss, predictions, clipped gradient).  Outputs d_loss[1], d_pred[n_graphs], d_grad_norm[1],
 * d_grad[2513] may be NULL.  Bitwise reproducible. */
int  rlvd_q_head_train_step(rlvd_engine* e, void* stream, int32_t n_graphs, const float* d_head_in, const float* d_target,
                            const float* d_dropout_uniform, float dropout_p, int32_t head_act, float lr, float beta1,
                            float beta2, float eps, float max_norm, int32_t step, int32_t apply, float* d_exp_avg,
                            float* d_exp_avg_sq, float* d_loss, float* d_pred, float* d_grad_norm, float* d_grad);
/* Copies the current device value of a trainable tensor (checkpoint name, Q_emb.* / Q_output.*) to the host:
 * policy_value_net.state_dict() after an update (trainer.py:120, train.py:102-104). */
int  rlvd_get_weight(rlvd_engine* e, const char* name, float* h_out, int64_t numel);

/* ---- replay-minibatch forward + backward of the WHOLE model (rlsim/drl/trainer.py:163-177 `dqn_update` with
 *      train_all = True, configs/train_dqn.toml:27; rlsim/drl/trainer.py:87-98 `context_bandit_update`) ---------------------
 * Parameters and gradients are flat fp32 vectors of rlvd_train_param_count() = 614,447 floats in the checkpoint's state_dict
 * order (every floating-point tensor, buffers included); rlvd_train_layout(i, ...) describes entry i (returns non-zero past
 * the end): checkpoint name, offset, element count and whether the reference optimises it (buffers `cutoff`, `freqs` do not).
 * The gradient vector is the bucket rl-train all-reduces across ranks before rlvd_adam_step. */
int64_t rlvd_train_param_count(void);
int  rlvd_train_layout(int32_t index, const char** name, int64_t* offset, int64_t* numel, int32_t* trainable);
/* One minibatch: structures [0, n_react) are the reactants of reaction graphs 0..n_react-1, structures [n_react, 2 n_react)
 * their products (n_struct = 2 n_react, equal atom counts pairwise; graph from rlvd_radius_graph_*).  loss_mode 0 (dqn):
 * loss = mean((d_target0 - rl_q)^2); loss_mode 1 (context bandit): loss = mean((q0 - d_target0)^2) + mean((q1 - d_target1)^2).
 * d_dropout_uniform[n_react, 80] (may be NULL) drives the DQN heads' train-mode dropout like rlvd_q_head_train_step.
 * Reads d_params, OVERWRITES d_grads (unclipped d loss / d parameter; zero for tensors the loss does not reach), writes
 * d_loss[1], d_pred[n_react, 3] = (rl_q, q0, q1) and optionally d_q[n_nodes, 128] (final node scalars).  fp32 SIMT kernels,
 * CSR-order sums, no atomics: bitwise reproducible.  The engine supplies only the species lookup and cutoff. */
int  rlvd_train_forward_backward(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes, int64_t n_edges,
                                 const int32_t* d_node_ptr, const int32_t* d_Z, const float* d_pos, const float* d_cell,
                                 const int32_t* d_node_struct, const int32_t* d_row_ptr, const int32_t* d_col,
                                 const int8_t* d_shift, int32_t n_react, const rlvd_q_params* qp, const float* d_temperature,
                                 int32_t loss_mode, const float* d_target0, const float* d_target1,
                                 const float* d_dropout_uniform, float dropout_p, const float* d_params, float* d_grads,
                                 float* d_loss, float* d_pred, float* d_q);
/* torch.nn.utils.clip_grad_norm_(max_norm) over the entries with d_trainable[i] != 0 (NULL = all; max_norm <= 0 = no clip)
 * followed by torch.optim.Adam's update (no weight decay, no amsgrad) in place on d_params; step is 1-based; d_grad_norm[2]
 * (may be NULL) receives the total norm and the clip coefficient. */
int  rlvd_adam_step(rlvd_engine* e, void* stream, int64_t n, float* d_params, const float* d_grads, const uint8_t* d_trainable,
                    float* d_exp_avg, float* d_exp_avg_sq, float lr, float beta1, float beta2, float eps, float max_norm,
                    int32_t step, float* d_grad_norm);

/* ---- Warren-Cowley pair counts (rlsim/utils/sro.py:49-83 `get_sro_from_atoms`, used per candidate by the sro_pixel
 *      filter at rlsim/drl/simulator.py:60-92) ------------------------------------------------------------------ */
/* d_species_index[M] in [0, n_species) (index into the sorted species list); d_counts[n_struct, n_species, n_species]
 * = ordered first-shell pair counts (i of species a, j of species b) with r_cut = (1 + sqrt2)/2 * NN_dist;
 * d_status[g] = 1 for non-orthorhombic cells (not computed).  SRO[a,b] = 1 - (n_ab / n_pairs) / c_a / c_b on the host. */
int  rlvd_sro_counts(rlvd_engine* e, void* stream, int32_t n_struct, const int32_t* d_node_ptr, const double* d_pos,
                     const double* d_cell, const int32_t* d_species_index, int32_t n_species, int64_t* d_counts,
                     int32_t* d_status);

/* ---- T1: t_net head on top of rlvd_painn_representation (replaces rgnn `t_net` forward reached from
 *      rlsim/cli/scripts/estimate_time.py:58 `model(batch, inference=True)["time"]` and rlsim/time/train.py:108,114) -- */
/* Head weights live on the device (the caller's nn.Parameters); nn.Linear layout [out, in].  No t_net checkpoint is
 * bundled with the reference, so the head wiring is a named assumption (oracle TimeHeadSpec): pooled node scalars,
 * then T * T_scale and defect * D_scale, MLP (F+2 -> n_feat -> n_feat -> 1), time = tau * sigmoid(out). */
typedef struct rlvd_time_head {
    const float* d_w0; const float* d_b0;   /* [n_feat, 130], [n_feat] */
    const float* d_w1; const float* d_b1;   /* [n_feat, n_feat], [n_feat] */
    const float* d_w2; const float* d_b2;   /* [1, n_feat], [1] */
    int32_t n_feat;                         /* <= 128 */
    int32_t pooling;                        /* 0 mean, 1 sum */
    int32_t act;                            /* RLVD_ACT_* */
    float tau, T_scale, D_scale;
    int32_t output;                         /* 0: tau * sigmoid(out) ("time"); 1: the raw output (t_net_binary's "goal" logit,
                                               binary_cross_entropy_with_logits at rlsim/time/misc.py:205) */
} rlvd_time_head;
/* d_q[M,128] from rlvd_painn_representation; d_T / d_defect / d_time are [n_struct]. */
int  rlvd_time_readout(rlvd_engine* e, void* stream, int32_t n_struct, const int32_t* d_node_ptr, const float* d_q,
                       const float* d_T, const float* d_defect, const rlvd_time_head* head, float* d_time);

/* ---- whole path from HOST buffers (the call `model(batch, q_params)` makes after `batch_to`) ---- */
/* n_struct structures (h_node_ptr, h_Z, h_pos, h_cell, h_inv_cell, h_img_range), n_react reaction
 * graphs pairing them; copies inputs H2D, runs K1..K5 on `stream`, copies rl_q/q0/q1/delta_e/E_R/E_P
 * (each [n_react], NULL to skip) back and synchronises the stream.  Host buffers should be pinned. */
int  rlvd_score_reactions_host(rlvd_engine* e, void* stream, int32_t n_struct, int32_t n_nodes,
                               const int32_t* h_node_ptr, const int32_t* h_Z, const float* h_pos,
                               const float* h_cell, const float* h_inv_cell, const int32_t* h_img_range,
                               float cutoff, int32_t n_react, const int32_t* h_react_struct,
                               const int32_t* h_prod_struct, const rlvd_q_params* qp,
                               const float* h_temperature,
                               float* h_rl_q, float* h_q0, float* h_q1, float* h_delta_e,
                               float* h_E_R, float* h_E_P, int64_t* h_n_edges);

/* ---- the channel-mixing linear as a standalone op: Y[M,N] = act([A | A2]·Wᵀ + b) -------------------- */
/* d_A [M,K1 or K], optional d_A2 [M,K-K1] (concatenated along K), host weight h_W[N,K] in nn.Linear layout,
 * optional host bias, act: 0 none / 1 SiLU.  Runs the tcgen05 split-fp16 kernel when the "tensor_cores"
 * option is on (default) and the shape qualifies (N % 128 == 0, K in {64..256} multiple of 64), else the
 * fp32 SIMT tile kernel.  Synchronises the stream.  (nn.Linear inside rgnn's PaiNN) */
int  rlvd_linear(rlvd_engine* e, void* stream, int32_t M, int32_t K, int32_t N, const float* d_A,
                 const float* d_A2, int32_t K1, const float* h_W, const float* h_bias, int32_t act, float* d_Y);
/* Atomic numbers of the model's species (hyper_parameters["reaction_model"]["species"], the list rgnn's
 * SpeciesEnergyScale and embedding are built over; deploy.py:18 loads them with the checkpoint).  Optional; call BEFORE
 * rlvd_finalize_weights.  With <= 3 species the first interaction layer (q = embedding[Z], mu = 0) is evaluated as per-node
 * GEMMs over species-resolved radial sums instead of per-edge messages; batches containing any other atomic
 * number fall back to the per-edge kernels on the device, with identical results. */
int  rlvd_set_species(rlvd_engine* e, const int32_t* atomic_numbers, int32_t n);
/* Scalar hyper-parameters of the checkpoint that kernels need: "epsilon" (hyper_parameters.reaction_model.epsilon, the
 * constant inside the mixing block's sqrt; default 1e-8). */
int  rlvd_set_hyper(rlvd_engine* e, const char* name, double value);
/* Engine options: "tensor_cores" (1 = tcgen05 GEMMs and filters, 0 = fp32 SIMT kernels; both fp32-accurate),
 * "sliced_edge" (1 = structure-resident edge stream kernel, edge_stream.cu, for structures of <= 256 atoms; 0 = general),
 * "fuse_mlp" (1 = each Linear-SiLU-Linear context net is one launch with the hidden tile kept on the SM),
 * "fuse_mix" (1 = the |mu_V| and <mu_V, mu_W> reductions run in the mu_channel_mix GEMM epilogue),
 * "layer0_sums" (1 = layer 0 from species-resolved radial sums, see rlvd_set_species; 0 = per-edge kernels),
 * "prune_final_mu" (1 = a q-only call skips the last mixing block's mu update; same q bit for bit),
 * "hit_words" (1 = the CSR fill pass walks the hit bits of the count pass; same CSR bit for bit),
 * "stream_prefetch" (0 = off, 1..7 = the edge stream kernel asks L2 for the next wave's node rows after that many
 *   eighths of a CTA's tiles; results do not depend on it). */
int  rlvd_set_option(rlvd_engine* e, const char* name, int32_t value);

/* Device status word since the last call (synchronises `stream`, then clears it).  bit 0: an operand of a tensor-core GEMM
 * left the range of the fp16 hi/lo split (|x| >= 65504, inf or NaN) — the results of that call are not valid; set option
 * "tensor_cores" to 0 and repeat it (rlvd_score_reactions_host and the Python wrappers do this by themselves).  Physical
 * structures stay orders of magnitude below the limit; unphysically dense ones make the network diverge in any precision.
 * bit 1: the edge-tile pool of the stream kernel overflowed its capacity bound — results invalid (an internal sizing error;
 * rlvd_score_reactions_host returns non-zero, the Python wrappers raise). */
int  rlvd_check_status(rlvd_engine* e, void* stream, int32_t* flags);

/* ---- introspection ------------------------------------------------------------------------------ */
/* Number of kernels this library has launched on the engine since creation (bench.py "gpu_launches"). */
int64_t rlvd_launch_count(const rlvd_engine* e);
/* Event-timed milliseconds spent in the named kernel family since the last reset
 * ("neighbor", "edge", "node", "readout", "edge_pack" = the once-per-call edge embedding pre-pass); timing is off
 * unless enabled. */
int  rlvd_profile_enable(rlvd_engine* e, int32_t on);
int  rlvd_profile_read(rlvd_engine* e, const char* family, double* ms_total, int64_t* launches);
int  rlvd_profile_reset(rlvd_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* RLVD_B200_H */
