// Library-level entry points and the composite GAEncoder.forward (models_con/ga.py:87-127):
// the whole denoiser evaluation is chained here natively so the host makes ONE call per step and the
// sequence can be captured in a CUDA graph.
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "pf_common.cuh"

namespace pf {

static std::atomic<int64_t> g_launches{0};
static int g_num_sms = 148;
static std::atomic<int> g_edge_impl{2}, g_gemm_impl{2}, g_ipa_impl{4}, g_chain_impl{1}, g_edge_terms{0}, g_mma_order{1};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_profile{0};
static std::mutex g_prof_mu;
static std::vector<cudaEvent_t> g_ev_pool;               // reusable events
constexpr int NPROF = 3;                                  // 0 IPA attention, 1 edge transition, 2 IPA operand packers
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_ev[NPROF];
static cudaEvent_t g_open[NPROF] = {nullptr, nullptr, nullptr};

static cudaEvent_t take_event() {
  if (!g_ev_pool.empty()) { cudaEvent_t e = g_ev_pool.back(); g_ev_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void profile_begin(int c, cudaStream_t st) {
  if (!g_profile.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_open[c] = take_event();
  cudaEventRecord(g_open[c], st);
}
void profile_end(int c, cudaStream_t st) {
  if (!g_profile.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_open[c]) return;
  cudaEvent_t e = take_event();
  cudaEventRecord(e, st);
  g_ev[c].emplace_back(g_open[c], e);
  g_open[c] = nullptr;
}
int num_sms() { return g_num_sms; }
static std::atomic<void*> g_dbg_ptr{nullptr};
static std::atomic<size_t> g_dbg_bytes{0};
void* debug_buffer(size_t bytes) { return g_dbg_bytes.load() >= bytes ? g_dbg_ptr.load() : nullptr; }
int opt_edge_impl() { return g_edge_impl.load(std::memory_order_relaxed); }
int opt_gemm_impl() { return g_gemm_impl.load(std::memory_order_relaxed); }
int opt_ipa_impl() { return g_ipa_impl.load(std::memory_order_relaxed); }
int opt_chain_impl() { return g_chain_impl.load(std::memory_order_relaxed); }
int opt_edge_terms() { return g_edge_terms.load(std::memory_order_relaxed); }
int opt_mma_order() { return g_mma_order.load(std::memory_order_relaxed); }
void edge_umma_resolve_driver();

static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct GaWorkspace {
  float *xmix, *s, *ta, *tb, *ya, *yb, *proj, *pts, *feats, *qkv, *ctx, *quat, *rot, *trans, *upd, *ang_raw, *zbuf;
  void* edge_ws;
  void* ipa_ws;
  void* gemm_ws;
  size_t edge_ws_bytes, ipa_ws_bytes, gemm_ws_bytes, total;
};

static GaWorkspace carve(void* base, int B, int L) {
  const size_t M = (size_t)B * L;
  unsigned char* p = static_cast<unsigned char*>(base);
  size_t off = 0;
  GaWorkspace w;
  auto take = [&](size_t bytes) {
    float* r = reinterpret_cast<float*>(p + off);
    off += al(bytes);
    return r;
  };
  w.xmix = take(M * NMIXP * 4);
  w.s = take(M * 128 * 4);
  w.ta = take(M * 128 * 4);
  w.tb = take(M * 128 * 4);
  w.ya = take(M * 128 * 4);
  w.yb = take(M * 128 * 4);
  w.proj = take(M * NPROJ * 4);
  w.pts = take(M * H * NPT * 3 * 4);
  w.feats = take(M * NFEAT * 4);
  w.qkv = take(M * 384 * 4);
  w.ctx = take(M * 128 * 4);
  w.quat = take(M * 4 * 4);
  w.rot = take(M * 9 * 4);
  w.trans = take(M * 3 * 4);
  w.upd = take(M * 6 * 4);
  w.ang_raw = take(M * 5 * 4);
  w.zbuf = take(M * L * CZ * 4);
  w.edge_ws_bytes = edge_workspace_bytes(B, L);
  w.edge_ws = take(w.edge_ws_bytes);
  w.ipa_ws_bytes = ipa_workspace_bytes(B, L);
  w.ipa_ws = take(w.ipa_ws_bytes);
  w.gemm_ws_bytes = linear_workspace_bytes(NPROJ);
  w.gemm_ws = take(w.gemm_ws_bytes);
  w.total = off;
  return w;
}

// ---- prepacked weight images (pf_ga_prepack) -----------------------------------------------------
// Every K = 128 Linear that takes the tcgen05 GEMM and every block's edge-transition MLP; the order below
// defines the offsets inside the buffer, so pf_ga_prepack and pf_ga_encoder_forward agree by construction.
// kind 0: tcgen05 tiles of a K = 128 Linear; 1: edge-transition MLP image; 2: composed per-residue edge terms
// (fp32 Wc [512,128] | bc [512] | tiles of the P, Q, U, V row groups)
struct PackItem { const float* w; int N; size_t off; bool edge; const float* w2; const float* wf; int kind; int blk; };
constexpr size_t TERMS_WC = 0, TERMS_BC = 512 * 128 * 4, TERMS_TILES = TERMS_BC + 2048;
constexpr int TERMS_ROW[4] = {0, 192, 384, 448}, TERMS_N[4] = {192, 192, 64, 64}, TERMS_TILE0[4] = {0, 2, 4, 5};

static size_t enumerate_packables(const pf_ga_weights* w, std::vector<PackItem>* items) {
  size_t off = 0;
  auto lin = [&](const float* p, int N) {
    if (!p) return;
    if (items) items->push_back(PackItem{p, N, off, false, nullptr, nullptr, 0, -1});
    off += al(gemm_umma_pack_bytes(N));
  };
  lin(w->g[PF_G_MIX2_W], 128);
  if (w->g[PF_G_MIX0_W]) {                             // feature-mix layer [128, 629]: one tile per 128-column chunk
    if (items) items->push_back(PackItem{w->g[PF_G_MIX0_W], 128, off, false, nullptr, nullptr, 4, -1});
    off += al((NMIXP / 128) * gemm_umma_pack_bytes(128));
  }
  lin(w->g[PF_G_SEQNET0_W], 128); lin(w->g[PF_G_SEQNET2_W], 128); lin(w->g[PF_G_SEQNET4_W], 20);
  lin(w->g[PF_G_ANGNET0_W], 128); lin(w->g[PF_G_ANGNET2_W], 128); lin(w->g[PF_G_ANGNET4_W], 5);
  for (int b = 0; b < w->num_blocks; ++b) {
    const float* const* W = w->blk[b];
    lin(W[PF_B_PROJ_W], NPROJ);
    for (int o : {(int)PF_B_T0_IN_W, (int)PF_B_T1_IN_W}) {
      lin(W[o], 384); lin(W[o + 2], 128); lin(W[o + 4], 128); lin(W[o + 6], 128);
    }
    lin(W[PF_B_NT1_W], 128); lin(W[PF_B_NT2_W], 128); lin(W[PF_B_NT3_W], 128);
    lin(W[PF_B_POST_W], 128); lin(W[PF_B_BB_W], 6);
    if (W[PF_B_OUT_W]) {                               // IPA linear_out [128, 1536]: one tile per 128-column chunk
      if (items) items->push_back(PackItem{W[PF_B_OUT_W], 128, off, false, nullptr, nullptr, 3, b});
      off += al((NFEAT / 128) * gemm_umma_pack_bytes(128));
    }
    if (W[PF_B_ET_W1] && W[PF_B_ET_W2] && W[PF_B_ET_WF]) {
      if (items) items->push_back(PackItem{W[PF_B_ET_W1], 0, off, true, W[PF_B_ET_W2], W[PF_B_ET_WF], 1, b});
      off += al(edge_umma_weight_image_bytes());
      if (W[PF_B_ET_INIT_W] && W[PF_B_ET_INIT_B] && W[PF_B_ET_B1] && W[PF_B_ET_BF]) {
        if (items) items->push_back(PackItem{W[PF_B_ET_INIT_W], 0, off, false, nullptr, nullptr, 2, b});
        off += al(TERMS_TILES + 6 * gemm_umma_pack_bytes(128));
      }
    }
  }
  return off;
}

}  // namespace pf

extern "C" {

int pf_version(void) { return 4; }

size_t pf_ga_prepack_bytes(const pf_ga_weights* w) {
  if (!w || w->num_blocks < 1 || w->num_blocks > PF_MAX_BLOCKS) return 0;
  return pf::enumerate_packables(w, nullptr);
}

int pf_ga_prepack(const pf_ga_weights* w, void* buffer, size_t buffer_bytes, void* stream) {
  using namespace pf;
  PF_REQUIRE(w && buffer, PF_ERR_NULL_POINTER);
  PF_REQUIRE(w->num_blocks >= 1 && w->num_blocks <= PF_MAX_BLOCKS, PF_ERR_BAD_CONFIG);
  PF_REQUIRE(aligned16(buffer), PF_ERR_MISALIGNED);
  std::vector<PackItem> items;
  const size_t total = enumerate_packables(w, &items);
  PF_REQUIRE(buffer_bytes >= total, PF_ERR_WORKSPACE_TOO_SMALL);
  unsigned char* base = static_cast<unsigned char*>(buffer);
  cudaStream_t st = as_stream(stream);
  for (const PackItem& it : items) {
    if (it.kind == 1) {
      PF_TRY(launch_edge_umma_pack_weights(it.w, it.w2, it.wf, base + it.off, st));
    } else if (it.kind == 2) {
      const float* const* W = w->blk[it.blk];
      float* wc = reinterpret_cast<float*>(base + it.off + TERMS_WC);
      float* bc = reinterpret_cast<float*>(base + it.off + TERMS_BC);
      PF_TRY(launch_edge_compose_terms(W[PF_B_ET_INIT_W], W[PF_B_ET_INIT_B], W[PF_B_ET_W1], W[PF_B_ET_B1], W[PF_B_ET_WF],
                                       W[PF_B_ET_BF], wc, bc, st));
      for (int q = 0; q < 4; ++q)
        PF_TRY(launch_gemm_umma_pack(wc + (size_t)TERMS_ROW[q] * 128, 128, TERMS_N[q],
                                     base + it.off + TERMS_TILES + TERMS_TILE0[q] * gemm_umma_pack_bytes(128), st));
    } else if (it.kind == 4) {
      for (int kc = 0; kc < NMIXP / 128; ++kc) {
        const int kvalid = NMIX - kc * 128 < 128 ? NMIX - kc * 128 : 128;
        PF_TRY(launch_gemm_umma_pack(it.w + kc * 128, NMIX, 128, base + it.off + kc * gemm_umma_pack_bytes(128), st, kvalid));
      }
    } else if (it.kind == 3) {
      for (int kc = 0; kc < NFEAT / 128; ++kc)
        PF_TRY(launch_gemm_umma_pack(it.w + kc * 128, NFEAT, 128, base + it.off + kc * gemm_umma_pack_bytes(128), st));
    } else {
      PF_TRY(launch_gemm_umma_pack(it.w, 128, it.N, base + it.off, st));
    }
  }
  return PF_OK;
}

const char* pf_strerror(int status) {
  switch (status) {
    case PF_OK: return "ok";
    case PF_ERR_BAD_SHAPE: return "bad shape";
    case PF_ERR_BAD_CONFIG: return "unsupported model configuration (kernels are specialised to learn_angle.yaml)";
    case PF_ERR_NULL_POINTER: return "null pointer";
    case PF_ERR_MISALIGNED: return "pointer not 16-byte aligned";
    case PF_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case PF_ERR_NO_DEVICE: return "no CUDA device";
    case PF_ERR_BAD_OPTION: return "unknown option";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown pf_status";
}

int pf_init(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return PF_ERR_NO_DEVICE;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return static_cast<int>(e);
  int sms = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return static_cast<int>(e);
  pf::g_num_sms = sms;
  pf::node_kernels_init();
  pf::ipa_kernels_init();
  pf::ipa_v2_kernels_init();
  pf::edge_kernels_init();
  pf::gemm_umma_init();
  pf::embed_kernels_init();
  pf::edge_umma_resolve_driver();
  e = cudaGetLastError();
  if (prev >= 0 && prev != device) cudaSetDevice(prev);   // the caller's current device is left as it was
  return e == cudaSuccess ? PF_OK : static_cast<int>(e);
}

int pf_check_config(int c_s, int c_z, int c_hidden, int no_heads, int no_qk_points, int no_v_points, int tfmr_heads,
                    int tfmr_layers) {
  const bool ok = c_s == pf::CS && c_z == pf::CZ && c_hidden == pf::C && no_heads == pf::H &&
                  no_qk_points == pf::PQ && no_v_points == pf::PV && tfmr_heads == 4 && tfmr_layers == 2;
  return ok ? PF_OK : PF_ERR_BAD_CONFIG;
}

int pf_set_option(const char* name, int value) {
  if (!name) return PF_ERR_NULL_POINTER;
  if (!std::strcmp(name, "edge_impl") && (value >= 0 && value <= 2)) { pf::g_edge_impl = value; return PF_OK; }
  if (!std::strcmp(name, "gemm_impl") && (value >= 0 && value <= 2)) { pf::g_gemm_impl = value; return PF_OK; }
  if (!std::strcmp(name, "ipa_impl") && (value == 0 || value == 3 || value == 4)) { pf::g_ipa_impl = value; return PF_OK; }
  if (!std::strcmp(name, "chain_impl") && (value >= 0 && value <= 1)) { pf::g_chain_impl = value; return PF_OK; }
  if (!std::strcmp(name, "edge_terms") && (value >= 0 && value <= 15)) { pf::g_edge_terms = value; return PF_OK; }
  if (!std::strcmp(name, "mma_order") && (value >= 0 && value <= 1)) { pf::g_mma_order = value; return PF_OK; }
  return PF_ERR_BAD_OPTION;
}

int pf_get_option(const char* name) {
  if (!name) return PF_ERR_NULL_POINTER;
  if (!std::strcmp(name, "edge_impl")) return pf::opt_edge_impl();
  if (!std::strcmp(name, "gemm_impl")) return pf::opt_gemm_impl();
  if (!std::strcmp(name, "ipa_impl")) return pf::opt_ipa_impl();
  if (!std::strcmp(name, "chain_impl")) return pf::opt_chain_impl();
  if (!std::strcmp(name, "edge_terms")) return pf::opt_edge_terms();
  if (!std::strcmp(name, "mma_order")) return pf::opt_mma_order();
  return PF_ERR_BAD_OPTION;
}

int pf_debug_buffer(void* device_buffer, size_t bytes) {
  pf::g_dbg_ptr.store(device_buffer);
  pf::g_dbg_bytes.store(device_buffer ? bytes : 0);
  return PF_OK;
}

int pf_profile_enable(int on) {
  pf::g_profile.store(on ? 1 : 0);
  return PF_OK;
}

int pf_profile_read_category(int category, double* ms_out, int64_t* launches_out) {
  if (category < 0 || category >= pf::NPROF) return PF_ERR_BAD_OPTION;
  std::lock_guard<std::mutex> lk(pf::g_prof_mu);
  double ms = 0.0;
  int64_t cnt = 0;
  for (auto& pr : pf::g_ev[category]) {
    cudaError_t e = cudaEventSynchronize(pr.second);
    if (e != cudaSuccess) return static_cast<int>(e);
    float t = 0.f;
    e = cudaEventElapsedTime(&t, pr.first, pr.second);
    if (e != cudaSuccess) return static_cast<int>(e);
    ms += t;
    ++cnt;
    pf::g_ev_pool.push_back(pr.first);
    pf::g_ev_pool.push_back(pr.second);
  }
  pf::g_ev[category].clear();
  if (ms_out) *ms_out = ms;
  if (launches_out) *launches_out = cnt;
  return PF_OK;
}

int pf_profile_read(double* ipa_ms, int64_t* ipa_launches, double* edge_ms, int64_t* edge_launches) {
  PF_TRY(pf_profile_read_category(0, ipa_ms, ipa_launches));
  return pf_profile_read_category(1, edge_ms, edge_launches);
}

int64_t pf_launch_count(void) { return pf::g_launches.load(); }
void pf_reset_launch_count(void) { pf::g_launches.store(0); }

size_t pf_ga_encoder_workspace_bytes(int B, int L) {
  if (B <= 0 || L <= 0) return 256;
  return pf::carve(nullptr, B, L).total;
}

int pf_seq_transformer_forward(const pf_ga_weights* w, int block, const float* x, const float* res_mask, float* y,
                               void* workspace, size_t workspace_bytes, int B, int L, void* stream) {
  using namespace pf;
  PF_REQUIRE(w && x && res_mask && y && workspace, PF_ERR_NULL_POINTER);
  PF_REQUIRE(B >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  PF_REQUIRE(w->num_blocks >= 1 && w->num_blocks <= PF_MAX_BLOCKS && block >= 0 && block < w->num_blocks, PF_ERR_BAD_CONFIG);
  if (B == 0 || L == 0) return PF_OK;
  PF_REQUIRE(aligned16(workspace) && aligned16(x) && aligned16(y), PF_ERR_MISALIGNED);
  const GaWorkspace ws = carve(workspace, B, L);
  PF_REQUIRE(workspace_bytes >= ws.total, PF_ERR_WORKSPACE_TOO_SMALL);
  std::vector<PackItem> packed;
  PF_REQUIRE(w->prepacked && w->prepacked_bytes >= enumerate_packables(w, nullptr), PF_ERR_BAD_CONFIG);
  enumerate_packables(w, &packed);
  const float* const* W = w->blk[block];
  for (int i = PF_B_T0_IN_W; i <= PF_B_T1_N2_B; ++i) PF_REQUIRE(W[i], PF_ERR_NULL_POINTER);
  cudaStream_t st = as_stream(stream);
  const int M = B * L;
  auto stage = [&](const float* wt, const float* bias, int N, int act) {
    NodeChainStage c{};
    for (const PackItem& it : packed)
      if (it.w == wt && !it.edge && it.kind != 2) c.wpack = static_cast<const unsigned char*>(w->prepacked) + it.off;
    c.bias = bias; c.N = N; c.act = act;
    return c;
  };
  auto run_chain = [&](const float* x0, std::vector<NodeChainStage>& c) -> int {
    for (const NodeChainStage& q : c) PF_REQUIRE(q.wpack, PF_ERR_BAD_CONFIG);
    return launch_node_chain(x0, M, c.data(), (int)c.size(), st);
  };
  const float* in = x;
  for (int l = 0; l < 2; ++l) {
    const int o = l == 0 ? PF_B_T0_IN_W : PF_B_T1_IN_W;
    {
      std::vector<NodeChainStage> c;
      NodeChainStage q = stage(W[o], W[o + 1], 384, 0); q.out = ws.qkv;                // in_proj
      c.push_back(q);
      PF_TRY(run_chain(in, c));
    }
    PF_TRY(launch_seq_attention(ws.qkv, res_mask, ws.ctx, B, L, st));
    std::vector<NodeChainStage> c;
    NodeChainStage q = stage(W[o + 2], W[o + 3], 128, 0);                              // out_proj, norm1(x + .)
    q.res = in; q.gamma = W[o + 8]; q.beta = W[o + 9]; q.save_res = q.next_a = true;
    c.push_back(q);
    q = stage(W[o + 4], W[o + 5], 128, 1); q.next_a = true;                            // linear1 + relu
    c.push_back(q);
    q = stage(W[o + 6], W[o + 7], 128, 0);                                             // linear2, norm2(y + .)
    q.res_from_chain = true; q.gamma = W[o + 10]; q.beta = W[o + 11]; q.out = l == 0 ? ws.ya : y;
    c.push_back(q);
    PF_TRY(run_chain(ws.ctx, c));
    in = ws.ya;
  }
  return PF_OK;
}

int pf_ga_encoder_forward(const pf_ga_weights* w, const float* t, const float* rot_t, const float* trans_t,
                          const float* angles_t, const int64_t* seqs_t, const float* node_embed,
                          const float* edge_embed, const float* res_mask, float* pred_rot, float* pred_trans,
                          float* pred_angles, float* logits, float* node_out, void* workspace, size_t workspace_bytes,
                          int B, int L, void* stream) {
  using namespace pf;
  PF_REQUIRE(w && t && rot_t && trans_t && angles_t && seqs_t && node_embed && edge_embed && res_mask && pred_rot &&
                 pred_trans && pred_angles && logits && workspace, PF_ERR_NULL_POINTER);
  PF_REQUIRE(B >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  PF_REQUIRE(w->num_blocks >= 1 && w->num_blocks <= PF_MAX_BLOCKS, PF_ERR_BAD_CONFIG);
  if (B == 0 || L == 0) return PF_OK;
  PF_REQUIRE(aligned16(workspace) && aligned16(edge_embed) && aligned16(node_embed), PF_ERR_MISALIGNED);
  const GaWorkspace ws = carve(workspace, B, L);
  PF_REQUIRE(workspace_bytes >= ws.total, PF_ERR_WORKSPACE_TOO_SMALL);
  for (int i = 0; i < PF_G_NSLOTS; ++i) PF_REQUIRE(w->g[i], PF_ERR_NULL_POINTER);
  cudaStream_t st = as_stream(stream);
  const int M = B * L;
  const int nb = w->num_blocks;
  // node-level Linear: the K = 128 layers take the tcgen05 GEMM through the workspace (gemm_impl = 2)
  std::vector<PackItem> packed;
  if (w->prepacked && w->prepacked_bytes >= enumerate_packables(w, nullptr)) enumerate_packables(w, &packed);
  auto find_packed = [&](const float* wt, bool edge) -> const void* {
    for (const PackItem& it : packed)
      if (it.w == wt && it.edge == edge && it.kind != 2) return static_cast<const unsigned char*>(w->prepacked) + it.off;
    return nullptr;
  };
  auto find_terms = [&](int blk) -> const unsigned char* {
    for (const PackItem& it : packed)
      if (it.kind == 2 && it.blk == blk) return static_cast<const unsigned char*>(w->prepacked) + it.off;
    return nullptr;
  };
  auto launch_linear = [&](const float* x, const float* wt, const float* bias, const float* residual,
                           const float* rowmask, float* y, int M_, int K_, int N_, int act, cudaStream_t s_) {
    const void* pre = linear_umma_eligible(K_, N_, residual != nullptr) ? find_packed(wt, false) : nullptr;
    return launch_linear_ws(x, wt, K_, bias, residual, rowmask, y, M_, K_, N_, act, ws.gemm_ws, ws.gemm_ws_bytes, s_,
                            pre);
  };

  // Fused layer chains (chain_impl = 1): need the tcgen05 GEMM and the prepacked weight images of every layer
  const bool chains = opt_chain_impl() == 1 && opt_gemm_impl() == 2 && !packed.empty();
  // K1: feature mix (ga.py:94-95); with chains the rows are zero-padded to 5 x 128 columns for the chain's K loop
  PF_TRY(launch_mix_features(node_embed, w->g[PF_G_SEQ_EMB], seqs_t, t, w->g[PF_G_TIME_FREQS], angles_t,
                             w->g[PF_G_ANG_FREQS], ws.xmix, B, L, st, chains ? NMIXP : NMIX));
  if (!chains)
    PF_TRY(launch_linear(ws.xmix, w->g[PF_G_MIX0_W], w->g[PF_G_MIX0_B], nullptr, nullptr, ws.ta, M, NMIX, 128, 1, st));
  auto stage = [&](const float* wt, const float* bias, int N, int act) {
    NodeChainStage c{};
    c.wpack = find_packed(wt, false); c.bias = bias; c.N = N; c.act = act;
    return c;
  };
  auto run_chain = [&](const float* x0, std::vector<NodeChainStage>& c) -> int {
    for (const NodeChainStage& q : c) PF_REQUIRE(q.wpack, PF_ERR_BAD_CONFIG);
    return launch_node_chain(x0, M, c.data(), (int)c.size(), st);
  };
  auto proj_stage = [&](int b) {
    NodeChainStage c = stage(w->blk[b][PF_B_PROJ_W], w->blk[b][PF_B_PROJ_B], NPROJ, 0);
    c.out = ws.proj;
    return c;
  };
  if (chains) {
    // mix0 (K = 629 as a 5-chunk K loop) + ReLU, mix2 (+ mask) -> s, then block 0's IPA projection from the same rows
    std::vector<NodeChainStage> c;
    NodeChainStage m0 = stage(w->g[PF_G_MIX0_W], w->g[PF_G_MIX0_B], 128, 1);
    m0.next_a = true;
    c.push_back(m0);
    NodeChainStage m2 = stage(w->g[PF_G_MIX2_W], w->g[PF_G_MIX2_B], 128, 0);
    m2.rowmask = res_mask; m2.out = ws.s; m2.next_a = true;
    c.push_back(m2);
    PF_REQUIRE(w->blk[0][PF_B_PROJ_W], PF_ERR_NULL_POINTER);
    c.push_back(proj_stage(0));
    for (const NodeChainStage& q : c) PF_REQUIRE(q.wpack, PF_ERR_BAD_CONFIG);
    PF_TRY(launch_node_chain(ws.xmix, M, c.data(), (int)c.size(), st, NMIXP / 128));
  } else {
    PF_TRY(launch_linear(ws.ta, w->g[PF_G_MIX2_W], w->g[PF_G_MIX2_B], nullptr, res_mask, ws.s, M, 128, 128, 0, st));
  }

  const float* rot = rot_t;      // block 0 uses the input rotation matrices as they are (ga.py:96)
  const float* trans = trans_t;
  const float* quat = nullptr;
  const float* z = edge_embed;
  for (int b = 0; b < nb; ++b) {
    const float* const* W = w->blk[b];
    const int n_need = (b < nb - 1) ? PF_B_NSLOTS : PF_B_ET_INIT_W;
    for (int i = 0; i < n_need; ++i) PF_REQUIRE(W[i], PF_ERR_NULL_POINTER);
    // IPA (ga.py:98-104); with chains the projection was produced by the previous chain
    if (!chains)
      PF_TRY(launch_linear(ws.s, W[PF_B_PROJ_W], W[PF_B_PROJ_B], nullptr, nullptr, ws.proj, M, 128, NPROJ, 0, st));
    if (opt_ipa_impl() < 3) PF_TRY(launch_ipa_points(ws.proj, rot, trans, ws.pts, M, st));   // variants 3 / 4 pack from proj
    IpaArgs ia{ws.proj, ws.pts, z, W[PF_B_LINB_W], W[PF_B_LINB_B], W[PF_B_DOWNZ_W], W[PF_B_DOWNZ_B], W[PF_B_HEAD_W],
               rot, trans, res_mask, ws.feats, B, L};
    PF_TRY(launch_ipa_attention(ia, ws.ipa_ws, ws.ipa_ws_bytes, st));
    if (chains) {
      // linear_out (K = 1536) * mask, LN(s + .) -> s, and the first transformer layer's in_proj: one chain
      std::vector<NodeChainStage> c;
      NodeChainStage q = stage(W[PF_B_OUT_W], W[PF_B_OUT_B], 128, 0);
      q.rowmask = res_mask; q.mask_first = true; q.res = ws.s; q.gamma = W[PF_B_IPA_LN_G]; q.beta = W[PF_B_IPA_LN_B];
      q.out = ws.s; q.next_a = true;
      c.push_back(q);
      q = stage(W[PF_B_T0_IN_W], W[PF_B_T0_IN_B], 384, 0); q.out = ws.qkv;
      c.push_back(q);
      for (const NodeChainStage& t : c) PF_REQUIRE(t.wpack, PF_ERR_BAD_CONFIG);
      PF_TRY(launch_node_chain(ws.feats, M, c.data(), (int)c.size(), st, NFEAT / 128));
    } else {
      PF_TRY(launch_linear(ws.feats, W[PF_B_OUT_W], W[PF_B_OUT_B], nullptr, res_mask, ws.ta, M, NFEAT, 128, 0, st));
      PF_TRY(launch_add_layernorm(ws.s, ws.ta, W[PF_B_IPA_LN_G], W[PF_B_IPA_LN_B], nullptr, ws.s, M, 128, st));
    }
    const bool last = (b == nb - 1);
    bool terms_ready = false;
    if (chains) {
      // sequence transformer (ga.py:105-106), post_tfmr (:107), node transition (:108-109), backbone update (:110) and
      // the next block's IPA projection: two attention launches and two layer chains
      const int o0 = PF_B_T0_IN_W, o1 = PF_B_T1_IN_W;
      PF_TRY(launch_seq_attention(ws.qkv, res_mask, ws.ctx, B, L, st));
      {
        std::vector<NodeChainStage> c;
        NodeChainStage q = stage(W[o0 + 2], W[o0 + 3], 128, 0);                       // out_proj, norm1(x + .)
        q.res = ws.s; q.gamma = W[o0 + 8]; q.beta = W[o0 + 9]; q.save_res = q.next_a = true;
        c.push_back(q);
        q = stage(W[o0 + 4], W[o0 + 5], 128, 1); q.next_a = true;                     // linear1 + relu
        c.push_back(q);
        q = stage(W[o0 + 6], W[o0 + 7], 128, 0);                                      // linear2, norm2(y + .)
        q.res_from_chain = true; q.gamma = W[o0 + 10]; q.beta = W[o0 + 11]; q.out = ws.ya; q.next_a = true;
        c.push_back(q);
        q = stage(W[o1], W[o1 + 1], 384, 0); q.out = ws.qkv;                          // layer 1 in_proj
        c.push_back(q);
        PF_TRY(run_chain(ws.ctx, c));
      }
      PF_TRY(launch_seq_attention(ws.qkv, res_mask, ws.ctx, B, L, st));
      {
        std::vector<NodeChainStage> c;
        NodeChainStage q = stage(W[o1 + 2], W[o1 + 3], 128, 0);
        q.res = ws.ya; q.gamma = W[o1 + 8]; q.beta = W[o1 + 9]; q.save_res = q.next_a = true;
        c.push_back(q);
        q = stage(W[o1 + 4], W[o1 + 5], 128, 1); q.next_a = true;
        c.push_back(q);
        q = stage(W[o1 + 6], W[o1 + 7], 128, 0);
        q.res_from_chain = true; q.gamma = W[o1 + 10]; q.beta = W[o1 + 11]; q.next_a = true;
        c.push_back(q);
        q = stage(W[PF_B_POST_W], W[PF_B_POST_B], 128, 0);                            // s += post_tfmr(y)
        q.res = ws.s; q.save_res = q.next_a = true;
        c.push_back(q);
        q = stage(W[PF_B_NT1_W], W[PF_B_NT1_B], 128, 1); q.next_a = true;
        c.push_back(q);
        q = stage(W[PF_B_NT2_W], W[PF_B_NT2_B], 128, 1); q.next_a = true;
        c.push_back(q);
        q = stage(W[PF_B_NT3_W], W[PF_B_NT3_B], 128, 0);                              // LN(s + .) * mask -> s
        q.res_from_chain = true; q.gamma = W[PF_B_NT_LN_G]; q.beta = W[PF_B_NT_LN_B]; q.rowmask = res_mask;
        q.out = ws.s; q.next_a = true;
        c.push_back(q);
        q = stage(W[PF_B_BB_W], W[PF_B_BB_B], 6, 0); q.out = ws.upd;
        c.push_back(q);
        if (!last) {
          // the per-residue terms of this block's edge transition, straight from the new s (composed weights)
          const unsigned char* tw = opt_edge_impl() != 0 ? find_terms(b) : nullptr;
          if (tw) {
            float* dst[4];
            edge_term_buffers(ws.edge_ws, B, L, &dst[0], &dst[1], &dst[2], &dst[3]);
            const float* bc = reinterpret_cast<const float*>(tw + TERMS_BC);
            for (int t4 = 0; t4 < 4; ++t4) {
              NodeChainStage e{};
              e.wpack = tw + TERMS_TILES + TERMS_TILE0[t4] * gemm_umma_pack_bytes(128);
              e.bias = bc + TERMS_ROW[t4]; e.N = TERMS_N[t4]; e.out = dst[t4];
              c.push_back(e);
            }
            terms_ready = true;
          }
          PF_REQUIRE(w->blk[b + 1][PF_B_PROJ_W], PF_ERR_NULL_POINTER);
          c.push_back(proj_stage(b + 1));
        }
        PF_TRY(run_chain(ws.ctx, c));
      }
    } else {
      // sequence transformer, 2 post-norm layers (ga.py:105-106)
      const float* x = ws.s;
      float* outs[2] = {ws.ya, ws.yb};
      for (int l = 0; l < 2; ++l) {
        const int o = l == 0 ? PF_B_T0_IN_W : PF_B_T1_IN_W;
        PF_TRY(launch_linear(x, W[o + 0], W[o + 1], nullptr, nullptr, ws.qkv, M, 128, 384, 0, st));
        PF_TRY(launch_seq_attention(ws.qkv, res_mask, ws.ctx, B, L, st));
        PF_TRY(launch_linear(ws.ctx, W[o + 2], W[o + 3], nullptr, nullptr, ws.ta, M, 128, 128, 0, st));
        PF_TRY(launch_add_layernorm(x, ws.ta, W[o + 8], W[o + 9], nullptr, outs[l], M, 128, st));      // norm1
        PF_TRY(launch_linear(outs[l], W[o + 4], W[o + 5], nullptr, nullptr, ws.tb, M, 128, 128, 1, st));  // linear1+relu
        PF_TRY(launch_linear(ws.tb, W[o + 6], W[o + 7], nullptr, nullptr, ws.ta, M, 128, 128, 0, st));     // linear2
        PF_TRY(launch_add_layernorm(outs[l], ws.ta, W[o + 10], W[o + 11], nullptr, outs[l], M, 128, st)); // norm2
        x = outs[l];
      }
      // s += post_tfmr(y)  (ga.py:107)
      PF_TRY(launch_linear(x, W[PF_B_POST_W], W[PF_B_POST_B], ws.s, nullptr, ws.s, M, 128, 128, 0, st));
      // node transition (ipa_pytorch.py:196-206) then mask (ga.py:108-109)
      PF_TRY(launch_linear(ws.s, W[PF_B_NT1_W], W[PF_B_NT1_B], nullptr, nullptr, ws.ta, M, 128, 128, 1, st));
      PF_TRY(launch_linear(ws.ta, W[PF_B_NT2_W], W[PF_B_NT2_B], nullptr, nullptr, ws.tb, M, 128, 128, 1, st));
      PF_TRY(launch_linear(ws.tb, W[PF_B_NT3_W], W[PF_B_NT3_B], nullptr, nullptr, ws.ta, M, 128, 128, 0, st));
      PF_TRY(launch_add_layernorm(ws.s, ws.ta, W[PF_B_NT_LN_G], W[PF_B_NT_LN_B], res_mask, ws.s, M, 128, st));
      // backbone update (ga.py:110-113)
      PF_TRY(launch_linear(ws.s, W[PF_B_BB_W], W[PF_B_BB_B], nullptr, nullptr, ws.upd, M, 128, 6, 0, st));
    }
    float* rot_o = last ? pred_rot : ws.rot;
    float* trans_o = last ? pred_trans : ws.trans;
    PF_TRY(launch_rigid_update(quat, quat ? nullptr : rot, trans, ws.upd, res_mask, ws.quat, rot_o, trans_o, M, st));
    quat = ws.quat; rot = rot_o; trans = trans_o;
    // edge transition (ga.py:115-118)
    if (!last) {
      PF_TRY(launch_edge_transition(ws.s, z, W[PF_B_ET_INIT_W], W[PF_B_ET_INIT_B], W[PF_B_ET_W1], W[PF_B_ET_B1],
                                    W[PF_B_ET_W2], W[PF_B_ET_B2], W[PF_B_ET_WF], W[PF_B_ET_BF], W[PF_B_ET_LN_G],
                                    W[PF_B_ET_LN_B], res_mask, ws.zbuf, ws.edge_ws, ws.edge_ws_bytes, B, L, st,
                                    find_packed(W[PF_B_ET_W1], true), terms_ready));
      z = ws.zbuf;
    }
  }
  // heads (ga.py:121-125): two three-layer chains from the final s, or one launch per layer
  if (chains && find_packed(w->g[PF_G_ANGNET4_W], false)) {
    const int wo[2][3] = {{PF_G_SEQNET0_W, PF_G_SEQNET2_W, PF_G_SEQNET4_W}, {PF_G_ANGNET0_W, PF_G_ANGNET2_W, PF_G_ANGNET4_W}};
    float* dst[2] = {logits, ws.ang_raw};
    const int width[2] = {20, 5};
    for (int hd = 0; hd < 2; ++hd) {
      std::vector<NodeChainStage> c;
      NodeChainStage q = stage(w->g[wo[hd][0]], w->g[wo[hd][0] + 1], 128, 1); q.next_a = true;
      c.push_back(q);
      q = stage(w->g[wo[hd][1]], w->g[wo[hd][1] + 1], 128, 1); q.next_a = true;
      c.push_back(q);
      q = stage(w->g[wo[hd][2]], w->g[wo[hd][2] + 1], width[hd], 0); q.out = dst[hd];
      c.push_back(q);
      PF_TRY(run_chain(ws.s, c));
    }
  } else {
    PF_TRY(launch_linear(ws.s, w->g[PF_G_SEQNET0_W], w->g[PF_G_SEQNET0_B], nullptr, nullptr, ws.ta, M, 128, 128, 1, st));
    PF_TRY(launch_linear(ws.ta, w->g[PF_G_SEQNET2_W], w->g[PF_G_SEQNET2_B], nullptr, nullptr, ws.tb, M, 128, 128, 1, st));
    PF_TRY(launch_linear(ws.tb, w->g[PF_G_SEQNET4_W], w->g[PF_G_SEQNET4_B], nullptr, nullptr, logits, M, 128, 20, 0, st));
    PF_TRY(launch_linear(ws.s, w->g[PF_G_ANGNET0_W], w->g[PF_G_ANGNET0_B], nullptr, nullptr, ws.ta, M, 128, 128, 1, st));
    PF_TRY(launch_linear(ws.ta, w->g[PF_G_ANGNET2_W], w->g[PF_G_ANGNET2_B], nullptr, nullptr, ws.tb, M, 128, 128, 1, st));
    PF_TRY(launch_linear(ws.tb, w->g[PF_G_ANGNET4_W], w->g[PF_G_ANGNET4_B], nullptr, nullptr, ws.ang_raw, M, 128, 5, 0, st));
  }
  PF_TRY(launch_mod_2pi(ws.ang_raw, pred_angles, M * 5, st));
  if (node_out) {
    cudaError_t e = cudaMemcpyAsync(node_out, ws.s, (size_t)M * 128 * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return PF_OK;
}

}  // extern "C"
