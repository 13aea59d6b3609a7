// Shared declarations for libpepflow_b200.so (sm_100a).  Internal header - the public ABI is
// include/pepflow_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pepflow_b200.h"

namespace pf {

// Model constants (configs/learn_angle.yaml:3-14 of the reference); kernels are specialised to them.
constexpr int H = 8;        // IPA heads
constexpr int C = 128;      // IPA hidden per head
constexpr int PQ = 8;       // query/key points
constexpr int PV = 12;      // value points
constexpr int NPT = PQ + PQ + PV;  // 28 points per (residue, head) in the pts buffer: q | k | v
constexpr int CS = 128;     // node channels
constexpr int CZ = 64;      // pair channels
constexpr int NPROJ = 3744; // q 1024 | kv 2048 | q_pts 192 | kv_pts 480
constexpr int OFF_Q = 0, OFF_KV = 1024, OFF_QP = 3072, OFF_KVP = 3264;
constexpr int NFEAT = 1536; // IPA concat width: o 1024 | o_pt xyz 288 | norms 96 | o_pair 128
constexpr int NMIX = 629;   // 128 + 128 + 128 + 245
constexpr float TWO_PI_F = 6.2831854820251465f;  // float(2*math.pi)

void count_launch();
// in-situ kernel timing (bench roofline): category 0 = IPA attention, 1 = edge transition
void profile_begin(int category, cudaStream_t st);
void profile_end(int category, cudaStream_t st);
int num_sms();
// device buffer registered with pf_debug_buffer() if it holds at least `bytes`, else nullptr
void* debug_buffer(size_t bytes);
int opt_edge_impl();
int opt_gemm_impl();
int opt_ipa_impl();
int opt_chain_impl();
int opt_edge_terms();
int opt_mma_order();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define PF_CHECK_LAUNCH()                                   \
  do {                                                      \
    pf::count_launch();                                     \
    cudaError_t pf_e_ = cudaGetLastError();                 \
    if (pf_e_ != cudaSuccess) return static_cast<int>(pf_e_); \
  } while (0)

#define PF_REQUIRE(cond, code) \
  do {                         \
    if (!(cond)) return (code); \
  } while (0)

#define PF_TRY(expr)              \
  do {                            \
    int pf_s_ = (expr);           \
    if (pf_s_ != PF_OK) return pf_s_; \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// R(q) exactly as openfold/utils/rigid_utils.py:185-205 (no renormalisation inside).
__device__ __forceinline__ void quat_to_rot_dev(float a, float b, float c, float d, float* R) {
  R[0] = a * a + b * b - c * c - d * d; R[1] = 2.f * b * c - 2.f * a * d;     R[2] = 2.f * b * d + 2.f * a * c;
  R[3] = 2.f * b * c + 2.f * a * d;     R[4] = a * a - b * b + c * c - d * d; R[5] = 2.f * c * d - 2.f * a * b;
  R[6] = 2.f * b * d - 2.f * a * c;     R[7] = 2.f * c * d + 2.f * a * b;     R[8] = a * a - b * b - c * c + d * d;
}

// torch.remainder(x, 2*pi) for fp32 (sign of the divisor).
__device__ __forceinline__ float mod_2pi(float x) {
  float m = fmodf(x, TWO_PI_F);
  if (m != 0.f && m < 0.f) m += TWO_PI_F;
  return m;
}

// ---- internal launchers shared between translation units ---------------------------------------
struct IpaArgs {
  const float* proj;    // [B*L, 3744]
  const float* pts;     // [B*L, H, 28, 3] global-frame points
  const float* z;       // [B, L, L, 64]
  const float* w_b;     // [8, 64]
  const float* b_b;     // [8]
  const float* w_dz;    // [16, 64]
  const float* b_dz;    // [16]
  const float* head_w;  // [8]  softplus(head_weights) * sqrt(1/108)
  const float* rot;     // [B*L, 9]
  const float* trans;   // [B*L, 3]
  const float* mask;    // [B*L]
  float* feats;         // [B*L, 1536]
  int B, L;
};
int launch_linear(const float* x, const float* w, const float* bias, const float* residual, const float* rowmask,
                  float* y, int M, int K, int N, int act, cudaStream_t st);
size_t gemm_umma_pack_bytes(int N);
// kvalid: columns of w that exist from w[0] on (< 128 for the ragged last chunk of a K loop; the rest packs as zero)
int launch_gemm_umma_pack(const float* w, int ldw, int N, void* wpack, cudaStream_t st, int kvalid = 128);
int launch_linear_umma(const float* x, const float* w, int ldw, const float* bias, const float* rowmask, float* y,
                       int M, int N, int act, const void* wpack, bool prepacked, cudaStream_t st);
void gemm_umma_init();
// One layer of a fused node-level chain (launch_node_chain, pf_gemm_umma.cu): y = act(A W^T + b) [+ res] -> [LayerNorm]
// -> [* rowmask]; A is the previous layer's y when that layer has next_a, else unchanged.
struct NodeChainStage {
  const void* wpack;            // tcgen05 weight image of W [N, 128] (launch_gemm_umma_pack / pf_ga_prepack)
  const float* bias;            // [N] or null
  const float* res;             // residual rows in HBM [M, 128], or null
  const float* gamma; const float* beta;   // LayerNorm, or null
  const float* rowmask;         // [M] or null
  float* out;                   // [M, N] stored result, or null
  int N;                        // layer width; > 128 streams tiles out and must not feed a next layer
  int act;                      // 1 = ReLU
  bool res_from_chain;          // add the row saved by an earlier layer (save_res)
  bool save_res;                // keep this layer's y as the residual row for a later layer
  bool next_a;                  // y becomes the A operand of the following layers
  bool mask_first;              // apply rowmask before the residual / LayerNorm instead of after
};
// kchunks > 1: x is [M, 128 * kchunks] and the first layer contracts all of it (W image: one tile per 128-column chunk)
int launch_node_chain(const float* x, int M, const NodeChainStage* stages, int n, cudaStream_t st, int kchunks = 1);
size_t linear_workspace_bytes(int N);
int launch_linear_ws(const float* x, const float* w, int ldw, const float* bias, const float* residual,
                     const float* rowmask, float* y, int M, int K, int N, int act, void* ws, size_t ws_bytes,
                     cudaStream_t st, const void* prepacked = nullptr);
bool linear_umma_eligible(int K, int N, bool has_residual);
int launch_linear_ld(const float* x, const float* w, int ldw, const float* bias, float* y, int M, int K, int N,
                     cudaStream_t st);
int launch_mix_features(const float* node, const float* emb, const int64_t* seqs, const float* t, const float* tfreq,
                        const float* angles, const float* afreq, float* x, int B, int L, cudaStream_t st,
                        int ldx = NMIX);
constexpr int NMIXP = 640;  // NMIX rounded up to whole 128-column chunks
int launch_add_layernorm(const float* a, const float* b, const float* gamma, const float* beta, const float* rowmask,
                         float* y, int M, int N, cudaStream_t st);
int launch_seq_attention(const float* qkv, const float* mask, float* ctx, int B, int L, cudaStream_t st);
int launch_rigid_update(const float* quat_in, const float* rot_in, const float* trans_in, const float* upd,
                        const float* mask, float* quat_out, float* rot_out, float* trans_out, int n, cudaStream_t st);
int launch_ipa_points(const float* proj, const float* rot, const float* trans, float* pts, int rows, cudaStream_t st);
int launch_mod_2pi(const float* x, float* y, int n, cudaStream_t st);
int launch_quat_to_rot(const float* quat, float* rot, int n, cudaStream_t st);
size_t ipa_workspace_bytes(int B, int L);
int launch_ipa_attention(const IpaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t ipa_v2_workspace_bytes(int B, int L);
int launch_ipa_attention_v3(const IpaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st,
                            bool decoupled);
void ipa_v2_kernels_init();
size_t edge_workspace_bytes(int B, int L);
int launch_edge_transition(const float* s, const float* z_in, const float* w_init, const float* b_init,
                           const float* w1, const float* b1, const float* w2, const float* b2, const float* wf,
                           const float* bf, const float* ln_g, const float* ln_b, const float* mask, float* z_out,
                           void* workspace, size_t workspace_bytes, int B, int L, cudaStream_t st,
                           const void* prepacked_weights = nullptr, bool terms_ready = false);
void edge_term_buffers(void* workspace, int B, int L, float** P, float** Q, float** U, float** V);
int launch_edge_compose_terms(const float* w_init, const float* b_init, const float* w1, const float* b1,
                              const float* wf, const float* bf, float* wc, float* bc, cudaStream_t st);
// 128-byte CUtensorMap over a pair tensor z [B, L, L, 64] fp32 viewed as (c: 64, j: L, bi: B*L), boxes of
// [32 c, 8 j, 1 bi] = 1 KB with the 128-byte swizzle (shared by the edge-transition and IPA kernels)
int encode_z_map(void* tensor_map, const float* z, int B, int L);
void edge_umma_init();
size_t edge_umma_pack_bytes(int B, int L);
int launch_edge_umma(const float* z_in, const float* P, const float* Q, const float* U, const float* V,
                     const float* w1, const float* w2, const float* wf, const float* b2, const float* ln_g,
                     const float* ln_b, const float* mask, float* z_out, void* wpack, int B, int L, cudaStream_t st,
                     const void* prepacked_weights = nullptr);
size_t edge_umma_weight_image_bytes();
int launch_edge_umma_pack_weights(const float* w1, const float* w2, const float* wf, void* image, cudaStream_t st);
void node_kernels_init();
void ipa_kernels_init();
void edge_kernels_init();
void embed_kernels_init();

}  // namespace pf
