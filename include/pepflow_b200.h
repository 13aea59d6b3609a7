/*
 * pepflow_b200.h  --  C ABI of libpepflow_b200.so (sm_100a), the drop-in boundary for the
 * PepFlow flow-matching denoising hot path.
 *
 * The reference (Ced3-han/PepFlowww) has no FFI / operator registry: its boundary is the Python
 * object API (FlowModel.forward/.sample, GAEncoder.forward, ...).  The host side of this repo
 * (the modules of pepflowww_b200/) keeps that object API and binds the entry points below with ctypes; each
 * entry point names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer into caller-owned, contiguous memory (torch tensors);
 *     fp32 unless the name says otherwise; `seqs` are int64; masks are fp32 {0,1};
 *   - the last argument is the cudaStream_t (as void*) the work is enqueued on; calls are
 *     asynchronous, re-entrant, allocate nothing and keep no mutable global state;
 *   - return value: 0 = PF_OK, otherwise a pf_status (negative: argument errors; positive: the
 *     cudaError_t of a failed launch).  pf_strerror() maps it to text.  No C++ exception crosses.
 *   - shapes use B complexes, L residues, H=8 heads, C=128 hidden, PQ=8, PV=12, c_s=128, c_z=64
 *     (configs/learn_angle.yaml:3-14).  The kernels are specialised to these constants;
 *     pf_check_config() rejects anything else.
 */
#ifndef PEPFLOW_B200_H
#define PEPFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pf_status {
  PF_OK = 0,
  PF_ERR_BAD_SHAPE = -1,
  PF_ERR_BAD_CONFIG = -2,
  PF_ERR_NULL_POINTER = -3,
  PF_ERR_MISALIGNED = -4,
  PF_ERR_WORKSPACE_TOO_SMALL = -5,
  PF_ERR_NO_DEVICE = -6,
  PF_ERR_BAD_OPTION = -7
} pf_status;

/* ---- library ---------------------------------------------------------------------------- */
int pf_version(void);                       /* ABI version (this header: 4; 4 added pf_sampler_step, pf_zero_center,
                                               pf_seq_transformer_forward and the "edge_terms" option) */
const char* pf_strerror(int status);
int pf_init(int device);                    /* opt kernels into >48 KB shared memory, query SMs */
int pf_check_config(int c_s, int c_z, int c_hidden, int no_heads, int no_qk_points, int no_v_points,
                    int tfmr_heads, int tfmr_layers);  /* configs/learn_angle.yaml:3-14        */
/* Kernel variant switches (test seams, not multi-backend dispatch: every variant is sm_100a CUDA).
 *   "edge_impl": 0 = fp32 CUDA-core kernel, 1 = 3xFP16 mma.sync kernel, 2 = 3xFP16 tcgen05 kernel (CTA pairs,
 *                operands in tensor memory; default)
 *   "gemm_impl": 0 = fp32 CUDA-core GEMM, 1 = 3xFP16 mma.sync GEMM, 2 = 3xFP16 tcgen05 GEMM for the K = 128
 *                layers that are given a workspace, mma.sync otherwise (default)
 *   "ipa_impl" : 0 = fp32 CUDA-core attention (cross-check), 3 = tensor-core attention (point-distance term folded into
 *                the Q K^T contraction, K'/V' fragments of a key tile bulk-copied into shared memory, head warps / pair
 *                warps, TMA z ring), 4 = variant 3 with the pair warps decoupled (default): the pair bias of the next key
 *                tile is computed before the o_pair accumulation of the current one (6-slot z row ring), Q' fragments
 *                parked in tensor memory so the pair warps get 88 registers
 *   "chain_impl": 1 = the K = 128 node layers between the attention kernels run as fused layer chains (one
 *                kernel per chain, activations in tensor memory; needs gemm_impl = 2 and prepacked weights;
 *                default), 0 = one GEMM / LayerNorm launch per layer
 *   "edge_terms": split-precision products the tcgen05 edge kernel issues per GEMM (error budget:
 *                profiles/r2_edge_error_budget.txt).  Bit g (0: z W1z^T, 1: z Wfz^T, 2: h1 W2^T, 3: h2 Wf^T) set =
 *                that GEMM drops its A_hi W_lo product (two passes instead of three).  Default 0: three passes everywhere.
 *   "mma_order" : issue order of the three split products in the tcgen05 GEMM kernels: 1 = all small cross products of a
 *                K chunk before its hi x hi products (the tensor-memory accumulator is truncated after every MMA; small
 *                terms first keeps most of those roundings at the 2^-11 scale: scripts/gpu_gemm_error.py), 0 = interleaved
 */
int pf_set_option(const char* name, int value);
int pf_get_option(const char* name);
/* Launch counter: number of kernels this library has enqueued since the last reset. */
int64_t pf_launch_count(void);
void pf_reset_launch_count(void);

/* In-situ kernel timing for bench.py's roofline: when enabled, CUDA events are recorded on the launch
 * stream around every IPA-attention and edge-transition kernel.  pf_profile_read() synchronises the
 * recorded events, returns the summed durations / launch counts since the last read, and resets. */
int pf_profile_enable(int on);
int pf_profile_read(double* ipa_ms, int64_t* ipa_launches, double* edge_ms, int64_t* edge_launches);
/* One category at a time: 0 = IPA attention, 1 = edge transition, 2 = the IPA operand packers (each pack launch group
 * of a block counts as one "launch"). */
int pf_profile_read_category(int category, double* ms, int64_t* launches);

/* Diagnostics: while a device buffer is registered, instrumented kernel variants write clock64() stamps of
 * their internal hand-offs into it (layout documented at the kernel).  NULL unregisters.  Not a hot-path call. */
int pf_debug_buffer(void* device_buffer, size_t bytes);

/* ---- generic node-level ops --------------------------------------------------------------- */
/* y[M,N] = act(x[M,K] W[N,K]^T + bias) (+ residual[M,N]) (* rowmask[M]).   act: 0 none, 1 ReLU.
 * Replaces models_con/ipa_pytorch.py:116-181 (Linear) and the nn.Linear layers of ga.py:22-45. */
int pf_linear(const float* x, const float* w, const float* bias, const float* residual, const float* rowmask,
              float* y, int M, int K, int N, int act, void* stream);
/* Same with a caller-owned scratch buffer (>= pf_linear_workspace_bytes(N)): with "gemm_impl" = 2 the K = 128
 * layers then run on the tcgen05 GEMM (A operand in tensor memory, pre-packed W tiles, TMA output). */
size_t pf_linear_workspace_bytes(int N);
int pf_linear_ws(const float* x, const float* w, const float* bias, const float* residual, const float* rowmask,
                 float* y, int M, int K, int N, int act, void* workspace, size_t workspace_bytes, void* stream);
/* y = LayerNorm_128(a + b) * rowmask   (b, rowmask optional).  ga.py:104, ipa_pytorch.py:203-204. */
int pf_add_layernorm(const float* a, const float* b, const float* gamma, const float* beta, const float* rowmask,
                     float* y, int M, int N, void* stream);

/* ---- K1: input feature mix (models_con/ga.py:94-95, utils.py:60-72, layers.py:104-113) ------ */
/* x[B*L,629] = node_embed(128) | seq_emb[seqs](128) | time_emb(t_b)(128) | angular_enc(angles)(245) */
int pf_mix_features(const float* node_embed, const float* seq_emb_table, const int64_t* seqs, const float* t,
                    const float* time_freqs /*[64]*/, const float* angles, const float* ang_freqs /*[24]*/,
                    float* x, int B, int L, void* stream);

/* ---- K2/K3: invariant point attention (models_con/ipa_pytorch.py:316-484) ------------------- */
/* proj[B*L,3744] = q(1024) | kv(2048, per head k(128)|v(128)) | q_pts(192) | kv_pts(480) from one
 * concatenated Linear; pf_ipa_points moves the point outputs into the global frame
 * (ipa_pytorch.py:360-387; Rigid.apply openfold/utils/rigid_utils.py:1124-1136) and writes
 * pts[B*L, H, 28, 3] = (8 q points | 8 k points | 12 v points).  rot [B*L,9] row-major. */
int pf_ipa_points(const float* proj, const float* rot, const float* trans, float* pts, int B, int L, void* stream);
/* Fused attention core: logits (scalar qk + pair bias + point distances + mask), softmax over j,
 * o, o_pt (back in the local frame, plus norms) and o_pair; writes feats[B*L,1536] in the
 * reference's concat order (ipa_pytorch.py:475).  z [B,L,L,64] is read once.
 * head_w = softplus(head_weights)*sqrt(1/108) precomputed by the caller ([8]). */
size_t pf_ipa_attention_workspace_bytes(int B, int L);   /* fp16 hi/lo operand fragments of the tensor-core variant */
int pf_ipa_attention(const float* proj, const float* pts, const float* z, const float* w_b, const float* b_b,
                     const float* w_dz, const float* b_dz, const float* head_w, const float* rot,
                     const float* trans, const float* mask, float* feats, void* workspace, size_t workspace_bytes,
                     int B, int L, void* stream);

/* ---- K5: sequence transformer attention core (torch.nn.TransformerEncoderLayer, ga.py:53-62) - */
/* qkv[B*L,384] (in_proj output) -> ctx[B*L,128]; 4 heads x 32, key-padding mask (mask==0 keys skipped). */
int pf_seq_attention(const float* qkv, const float* mask, float* ctx, int B, int L, void* stream);

/* ---- K7: backbone / rigid update (openfold/utils/rigid_utils.py:1039-1063,587-616,208-227,185-205) */
/* Block 0: quat_in == NULL and rot_in holds the input rotation matrices (rot -> quat inside).
 * upd[B*L,6]; mask[B*L]; writes quat_out[B*L,4] (unit), rot_out[B*L,9] = R(quat_out), trans_out. */
int pf_rigid_update(const float* quat_in, const float* rot_in, const float* trans_in, const float* upd,
                    const float* mask, float* quat_out, float* rot_out, float* trans_out, int n, void* stream);

/* ---- K8: edge transition (models_con/ipa_pytorch.py:233-248 + ga.py:118) --------------------- */
/* z_out = LN_64(W_f (MLP2([z,e_i,e_j]) + [z,e_i,e_j]) + b_f) * mask_i mask_j,  e = Linear_128->64(s).
 * z_out may alias z_in.  Weights are the reference tensors: w_init[64,128], w1/w2[192,192], wf[64,192]. */
size_t pf_edge_transition_workspace_bytes(int B, int L);
int pf_edge_transition(const float* s, const float* z_in, const float* w_init, const float* b_init,
                       const float* w1, const float* b1, const float* w2, const float* b2, const float* wf,
                       const float* bf, const float* ln_g, const float* ln_b, const float* mask, float* z_out,
                       void* workspace, size_t workspace_bytes, int B, int L, void* stream);

/* ---- K9: heads epilogue (ga.py:124-125) ------------------------------------------------------ */
int pf_mod_2pi(const float* x, float* y, int n, void* stream);          /* torch.remainder(x, 2*pi) */
int pf_quat_to_rot(const float* quat, float* rot, int n, void* stream); /* rigid_utils.py:185-205   */

/* ---- manifold maps (data/so3_utils.py:143-254,486-520; models_con/torus.py:5-26) -------------- */
int pf_so3_log(const float* rot, float* rotvec, int n, void* stream);
int pf_so3_exp(const float* rotvec, float* rot, int n, void* stream);
/* out = base * Exp(t * Log(base^T mat)); t has one entry per `group` consecutive matrices. */
int pf_so3_geodesic(const float* t, const float* mat, const float* base, float* out, int n, int group, void* stream);
int pf_tor_geodesic(const float* t, const float* ang1, const float* ang0, float* out, int n, int group, int d,
                    void* stream);

/* ---- K10: one Euler iteration of FlowModel.sample (models_con/flow_model.py:291-343) ----------- */
/* Post-process the denoiser output (where(generate_mask), categorical draw of residue types,
 * torsion masking; :291-303) and write the clean prediction of this step.
 *   uniforms: optional [n] injected U[0,1) for the categorical draw; NULL -> Philox(seed, counter). */
int pf_denoise_post(const float* pred_rot, const float* pred_trans, const float* pred_ang, const float* logits,
                    const float* rot1, const float* trans1, const float* ang1, const int64_t* seq1,
                    const uint8_t* gen_mask, const float* torsions_mask /*[22,5]*/, const float* uniforms,
                    uint64_t seed, uint64_t counter, float* clean_rot, float* clean_trans, float* clean_ang,
                    int64_t* clean_seq, float* clean_simplex, int n, float simplex_k, void* stream);
/* Euler update of the state on R^3 x SO(3) x T^5 x simplex (:316-333).  State tensors are updated
 * into the *_out buffers (may alias the inputs). */
int pf_euler_step(const float* rot_t, const float* trans_t, const float* ang_t, const float* simplex_t,
                  const float* clean_rot, const float* clean_trans, const float* clean_ang,
                  const int64_t* clean_seq, const float* trans0, const float* simplex0, const float* rot1,
                  const float* trans1, const float* ang1, const int64_t* seq1, const uint8_t* gen_mask,
                  const float* torsions_mask, const float* uniforms, uint64_t seed, uint64_t counter, float d_t,
                  float* rot_out, float* trans_out, float* ang_out, int64_t* seq_out, float* simplex_out, int n,
                  float simplex_k, void* stream);

/* ---- composite: GAEncoder.forward (models_con/ga.py:87-127) ----------------------------------- */
/* Weight table: device pointers to the reference state_dict tensors (fp32, contiguous), host array.
 * Slot order is pf_ga_global_slot / pf_ga_block_slot below. */
enum pf_ga_global_slot {
  PF_G_MIX0_W = 0, PF_G_MIX0_B, PF_G_MIX2_W, PF_G_MIX2_B, PF_G_SEQ_EMB, PF_G_ANG_FREQS, PF_G_TIME_FREQS,
  PF_G_SEQNET0_W, PF_G_SEQNET0_B, PF_G_SEQNET2_W, PF_G_SEQNET2_B, PF_G_SEQNET4_W, PF_G_SEQNET4_B,
  PF_G_ANGNET0_W, PF_G_ANGNET0_B, PF_G_ANGNET2_W, PF_G_ANGNET2_B, PF_G_ANGNET4_W, PF_G_ANGNET4_B,
  PF_G_NSLOTS
};
enum pf_ga_block_slot {
  PF_B_PROJ_W = 0 /* [3744,128] = cat(linear_q, linear_kv, linear_q_points, linear_kv_points) */, PF_B_PROJ_B,
  PF_B_LINB_W, PF_B_LINB_B, PF_B_DOWNZ_W, PF_B_DOWNZ_B, PF_B_HEAD_W /* softplus(head_weights)*sqrt(1/108) */,
  PF_B_OUT_W, PF_B_OUT_B, PF_B_IPA_LN_G, PF_B_IPA_LN_B,
  PF_B_T0_IN_W, PF_B_T0_IN_B, PF_B_T0_OUT_W, PF_B_T0_OUT_B, PF_B_T0_L1_W, PF_B_T0_L1_B, PF_B_T0_L2_W, PF_B_T0_L2_B,
  PF_B_T0_N1_G, PF_B_T0_N1_B, PF_B_T0_N2_G, PF_B_T0_N2_B,
  PF_B_T1_IN_W, PF_B_T1_IN_B, PF_B_T1_OUT_W, PF_B_T1_OUT_B, PF_B_T1_L1_W, PF_B_T1_L1_B, PF_B_T1_L2_W, PF_B_T1_L2_B,
  PF_B_T1_N1_G, PF_B_T1_N1_B, PF_B_T1_N2_G, PF_B_T1_N2_B,
  PF_B_POST_W, PF_B_POST_B, PF_B_NT1_W, PF_B_NT1_B, PF_B_NT2_W, PF_B_NT2_B, PF_B_NT3_W, PF_B_NT3_B,
  PF_B_NT_LN_G, PF_B_NT_LN_B, PF_B_BB_W, PF_B_BB_B,
  PF_B_ET_INIT_W, PF_B_ET_INIT_B, PF_B_ET_W1, PF_B_ET_B1, PF_B_ET_W2, PF_B_ET_B2, PF_B_ET_WF, PF_B_ET_BF,
  PF_B_ET_LN_G, PF_B_ET_LN_B,
  PF_B_NSLOTS
};
#define PF_MAX_BLOCKS 8
typedef struct pf_ga_weights {
  int32_t num_blocks;
  int32_t reserved;
  const float* g[PF_G_NSLOTS];
  const float* blk[PF_MAX_BLOCKS][PF_B_NSLOTS];
  /* Optional: tensor-core images of the weights (fp16 hi/lo tiles in shared-memory order) written once by
   * pf_ga_prepack() into a caller-owned device buffer.  NULL / 0: every call re-packs what it needs.  The
   * caller re-runs pf_ga_prepack() after changing any weight tensor. */
  const void* prepacked;
  uint64_t prepacked_bytes;
} pf_ga_weights;

size_t pf_ga_prepack_bytes(const pf_ga_weights* w);
int pf_ga_prepack(const pf_ga_weights* w, void* buffer, size_t buffer_bytes, void* stream);

size_t pf_ga_encoder_workspace_bytes(int B, int L);
/* t[B]; rot_t[B,L,9]; trans_t[B,L,3]; angles_t[B,L,5]; seqs_t[B,L] i64; node_embed[B,L,128];
 * edge_embed[B,L,L,64] (not modified); res_mask[B,L] fp32.  Outputs: pred_rot[B,L,9], pred_trans[B,L,3],
 * pred_angles[B,L,5] in [0,2pi), logits[B,L,20].  Optional node_out[B,L,128] (final node embedding). */
int pf_ga_encoder_forward(const pf_ga_weights* w, const float* t, const float* rot_t, const float* trans_t,
                          const float* angles_t, const int64_t* seqs_t, const float* node_embed,
                          const float* edge_embed, const float* res_mask, float* pred_rot, float* pred_trans,
                          float* pred_angles, float* logits, float* node_out, void* workspace,
                          size_t workspace_bytes, int B, int L, void* stream);

/* The sequence-transformer seam of one block (torch.nn.TransformerEncoder of ga.py:53-62 called at :105-106):
 * y[B,L,128] = the two post-norm encoder layers of block `block` applied to x[B,L,128] with key-padding mask
 * res_mask[B,L] - the same fused layer chains and attention kernel the composite launches.  workspace as
 * pf_ga_encoder_forward (needs w->prepacked). */
int pf_seq_transformer_forward(const pf_ga_weights* w, int block, const float* x, const float* res_mask, float* y,
                               void* workspace, size_t workspace_bytes, int B, int L, void* stream);

/* ---- one whole iteration of the FlowModel.sample loop (models_con/flow_model.py:287-343, last one :346-372) ----
 * State, trajectory and the iteration counter live on the device and every argument is the same for all iterations,
 * so one call (denoiser + post-processing + manifold Euler update, 73 kernels) can be captured in a CUDA graph once
 * and replayed num_steps times.  Iteration n = step[0] at the time the call executes: t = ts[n]; the clean prediction
 * goes to trajectory slot n; the state advances by d_t = ts[n+1] - ts[n] unless n is the last iteration; step[0]
 * becomes n + 1.  Categorical draws: uniforms[n][0 | 1][B*L] when given, else Philox4x32-10(seed, 2n | 2n+1, residue).
 * sample_bb / sample_ang / sample_seq = 0 pin that modality to the ground truth (:306-311, :336-342). */
typedef struct pf_sampler {
  const pf_ga_weights* weights;
  const float* node_embed; const float* edge_embed; const float* res_mask;     /* [B,L,128] [B,L,L,64] [B,L]       */
  void* workspace; uint64_t workspace_bytes;                                   /* pf_ga_encoder_workspace_bytes     */
  const float* rot1; const float* trans1; const float* ang1; const int64_t* seq1;   /* ground truth / context      */
  const uint8_t* gen_mask; const float* torsions_mask;                         /* [B,L] u8; [22,5]                  */
  const float* trans0; const float* simplex0;                                  /* initial noise x_0, simplex_0      */
  float* rot_t; float* trans_t; float* ang_t; int64_t* seq_t; float* simplex_t;     /* state, updated in place     */
  float* pred_rot; float* pred_trans; float* pred_ang; float* logits;          /* scratch: denoiser outputs         */
  float* traj_rot; float* traj_trans; float* traj_ang; int64_t* traj_seq; float* traj_simplex;  /* [num_steps,...]  */
  const float* ts;                                                             /* [num_steps] time grid             */
  const float* uniforms;                                                       /* [num_steps,2,B,L] or NULL         */
  int32_t* step;                                                               /* [2]: next iteration, scratch      */
  float* t_cur;                                                                /* [B] scratch                       */
  uint64_t seed;
  int32_t num_steps, B, L;
  int32_t sample_bb, sample_ang, sample_seq;
  float simplex_k;
  int32_t reserved;
} pf_sampler;
int pf_sampler_step(const pf_sampler* s, void* stream);

/* FlowModel.zero_center_part (models_con/flow_model.py:95-106), in place: pos[B,L,3] <- (pos - c_b) * res_mask with
 * c_b = sum_l pos * gen_mask / (sum_l gen_mask + 1e-8); center_out[B,3] optional. */
int pf_zero_center(float* pos, const uint8_t* gen_mask, const float* res_mask, float* center_out, int B, int L,
                   void* stream);

/* ---- once-per-sample pair embedder -------------------------------------------
 * EdgeEmbedder.forward, models_con/edge.py:39-112, fused: atom coordinates -> [N, L, L, 64] pair features in one
 * kernel.  aa[N,L] i64 (UNK already substituted where the sequence is hidden, edge.py:66-68), res_nb / chain_nb
 * [N,L] i64, pos_atoms[N,L,atoms_in,3] fp32 and mask_atoms[N,L,atoms_in] u8 (first 15 atoms used, edge.py:56-57),
 * structure_mask[N,L] u8 or NULL.  Host-prepared constants: softplus_coef[484,225] = softplus(aapair_to_distcoef);
 * t_aa[484,64] = aa_pair_embed W_o1[:, 0:64]^T; t_rel[65,64] = relpos_embed W_o1[:, 64:128]^T; the remaining
 * matrices TRANSPOSED to [in, out]: wd1_t[225,64], wd2_t[64,64] (distance_embed), wo1d_t[64,64] = W_o1[:, 128:192]^T,
 * wo1h_t[26,64] = W_o1[:, 192:218]^T, wo2_t, wo3_t [64,64] (out_mlp); biases [64].  out[N,L,L,64].            */
int pf_edge_embed(const int64_t* aa, const int64_t* res_nb, const int64_t* chain_nb, const float* pos_atoms,
                  const uint8_t* mask_atoms, const uint8_t* structure_mask, const float* softplus_coef,
                  const float* t_aa, const float* t_rel, const float* wd1_t, const float* bd1, const float* wd2_t,
                  const float* bd2, const float* wo1d_t, const float* wo1h_t, const float* bo1, const float* wo2_t,
                  const float* bo2, const float* wo3_t, const float* bo3, float* out, int N, int L, int atoms_in,
                  void* stream);

/* NodeEmbedder.forward, models_con/node.py:35-105, fused: local-frame coordinates placed
 * in the residue type's slot, backbone-dihedral encodings, 4-layer MLP, CA mask.  Inputs as pf_edge_embed.  Constants:
 * t1[22,256] = aatype_embed W1[:, 0:128]^T + b1; w1c_t[990,256] = W1[:, 128:1118]^T; w1d_t[39,256] = W1[:, 1118:1157]^T;
 * w2_t[256,128], w3_t[128,128], w4_t[128,128] transposed mlp.2 / mlp.4 / mlp.6 weights, b2 / b3 / b4 their biases.
 * out[N,L,128]. */
int pf_node_embed(const int64_t* aa, const int64_t* res_nb, const int64_t* chain_nb, const float* pos_atoms,
                  const uint8_t* mask_atoms, const uint8_t* structure_mask, const float* t1, const float* w1c_t,
                  const float* w1d_t, const float* w2_t, const float* b2, const float* w3_t, const float* b3,
                  const float* w4_t, const float* b4, float* out, int N, int L, int atoms_in, void* stream);

/* ---- post-sampling reconstruction ----------------------------------------------
 * full_atom_reconstruction, models_con/torsion.py:140-226 (with get_heavyatom_mask :126-138 as an optional second
 * output): backbone frames rot[n,3,3] / trans[n,3], torsions angles[n,5] = (psi, chi1..chi4) in radians and residue
 * types aa[n] i64 -> pos14[n,14,3]; optionally (NULL to skip) R_ret[n,6,3,3] / t_ret[n,6,3] = the backbone, psi and
 * chi1..chi4 frames, and mask_out[n,15] u8 = heavyatom_mask_table[aa] (table [22,15] u8, torsion.py:122-124).
 * Constant tables (pepflow/modules/protein/constants.py:665-749): rigid_rot[21,8,3,3], rigid_trans[21,8,3],
 * atom_group[21,14] i32, atom_pos[21,14,3].  The reference indexes its 21-row tables with aa and raises on
 * aa > 20; here such rows (PAD) get zero rigid groups - every atom lands on the backbone origin and the mask is 0 -
 * and the Python wrapper raises IndexError like the reference.                                                  */
int pf_full_atom_reconstruction(const float* rot, const float* trans, const float* angles, const int64_t* aa,
                                const float* rigid_rot, const float* rigid_trans, const int32_t* atom_group,
                                const float* atom_pos, const uint8_t* heavyatom_mask_table, float* pos14,
                                float* R_ret, float* t_ret, uint8_t* mask_out, long long n, void* stream);

/* reconstruct_backbone, pepflow/modules/common/geometry.py:446-489: rot[N,L,3,3], trans[N,L,3], aa / chain_nb /
 * res_nb [N,L] i64, mask[N,L] u8 -> pos_bb[N,L,4,3] = N, CA, C, O.  psi of residue i is measured on the rebuilt
 * backbone (N_i, CA_i, C_i, N_{i+1}) and is 0 at chain ends / gaps / masked residues (:355-390).  Tables
 * bb_coords[21,3,3], bb_oxygen[21,3] (constants.py:878-890); aa is clamped to [0, 20] (:462).                   */
int pf_reconstruct_backbone(const float* rot, const float* trans, const int64_t* aa, const int64_t* chain_nb,
                            const int64_t* res_nb, const uint8_t* mask, const float* bb_coords,
                            const float* bb_oxygen, float* pos_bb, int N, int L, void* stream);

/* get_torsion_angle, models_con/torsion.py:13-66 (the inverse of pf_full_atom_reconstruction; the dataset builder
 * stores its outputs as torsion_angle / torsion_angle_mask): pos_atoms[n,atoms_in,3] (atoms_in = 14 or 15, atom14
 * slot order), aa[n] i64 -> torsion[n,5] = (psi from N, CA, C, O; chi1..chi4) in [0, 2 pi), 0 where undefined, and
 * torsion_mask[n,5] u8 (angle exists for the residue type and its four atoms are not degenerate).  Residue types
 * outside 0..19 give zeros / false (:52-58).  chi_atoms[21,4,4] i32: atom14 slots of the defining atoms, -1 = no such
 * angle (pepflow/modules/protein/constants.py:372-400).                                                          */
int pf_torsion_angles(const float* pos_atoms, const int64_t* aa, const int32_t* chi_atoms, float* torsion,
                      uint8_t* torsion_mask, long long n, int atoms_in, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEPFLOW_B200_H */
