// Once-per-sample pair embedder: EdgeEmbedder.forward, models_con/edge.py:39-112, as ONE
// fused kernel.  The reference materialises [N, L, L, 225] distance features about five times (21 GB transient at
// N = 64, L = 271) and runs ~2,400 small kernels when chunked; here nothing of size L^2 x 225 ever exists - every
// pair goes from atom coordinates to its 64 output channels inside one thread.
//
//   per pair (i, j):  g_ab = exp(-softplus(coef[aa_i, aa_j])_ab * (|x_ia - x_jb| / 10)^2) * m_ia m_jb        (225)
//                     f_d  = relu(W_d2 relu(W_d1 g + b_d1) + b_d2) * psm_ij                                    (64)
//                     f_h  = AngularEncoding([phi_ij, psi_ij]) * psm_ij       geometry.py:393-418, layers.py:104-113 (26)
//                     y    = W_o3 relu(W_o2 relu(T_aa[aa_i, aa_j] + same_chain * T_rel[relpos] + W_o1[:, 128:192] f_d
//                                                + W_o1[:, 192:218] f_h + b_o1) + b_o2) + b_o3,  times mask_i mask_j
// T_aa = aa_pair_embed W_o1[:, 0:64]^T and T_rel = relpos_embed W_o1[:, 64:128]^T are the embedding lookups pushed
// through the first output layer (linear, so exact up to summation order); the host prepares them together with
// softplus(coef) and the transposed weight matrices (pepflowww_b200/edge.py).
//
// fp32 CUDA-core kernel: 64 k FLOP per pair, 0.3 TFLOP per sample at N = 64, L = 271 - once per 200 Euler steps,
// <0.5 % of a sampling run, so the tensor cores are not worth a split-precision pipeline here.  Persistent CTAs,
// one query row (b, i) at a time; all weights (128 KB) stay in shared memory and are read as warp-wide broadcasts,
// the per-row tables (softplus coefficients and T_aa rows of aa_i) are re-staged per row; a thread owns one pair and
// keeps its activations in registers; the output tile is staged in shared memory and written as full 128-byte lines.
#include "pf_common.cuh"
#include "pf_geom.cuh"

namespace pf {

constexpr int EE_A = 15, EE_AA = EE_A * EE_A;   // atoms per residue kept by the embedder, atom pairs
constexpr int EE_F = 64;                        // feature width
constexpr int EE_NH = 26;                       // dihedral encoding width: 2 angles x (1 + 6 sin + 6 cos)
constexpr int EE_NAA = 22;                      // amino-acid slots
constexpr int EE_NREL = 65;                     // relative positions -32..32
constexpr int EE_MAXT = 192;                    // threads per CTA (pairs per pass), upper bound
constexpr int EE_TP = EE_F + 1;                 // padded row of the per-pair-type tables and the output staging

// shared-memory layout (floats)
constexpr int EE_OFF_WD1 = 0;                                  // [225][64]
constexpr int EE_OFF_WD2 = EE_OFF_WD1 + EE_AA * EE_F;          // [64][64]
constexpr int EE_OFF_WO1D = EE_OFF_WD2 + EE_F * EE_F;          // [64][64]
constexpr int EE_OFF_WO1H = EE_OFF_WO1D + EE_F * EE_F;         // [26][64]
constexpr int EE_OFF_WO2 = EE_OFF_WO1H + EE_NH * EE_F;         // [64][64]
constexpr int EE_OFF_WO3 = EE_OFF_WO2 + EE_F * EE_F;           // [64][64]
constexpr int EE_OFF_B = EE_OFF_WO3 + EE_F * EE_F;             // b_d1 | b_d2 | b_o1 | b_o2 | b_o3
constexpr int EE_OFF_TREL = EE_OFF_B + 5 * EE_F;               // [65][65]
constexpr int EE_OFF_C = EE_OFF_TREL + EE_NREL * EE_TP;        // [22][225]   softplus coefficients of (aa_i, *)
constexpr int EE_OFF_TAA = EE_OFF_C + EE_NAA * EE_AA;          // [22][65]    T_aa rows of (aa_i, *)
constexpr int EE_OFF_PI = EE_OFF_TAA + EE_NAA * EE_TP;         // [45] x_i, [15] m_i, 4 spare
constexpr int EE_OFF_STAGE = EE_OFF_PI + 64;                   // [T][65] output staging; aliases x_j [T][45], m_j [T][15]
static_assert((EE_OFF_STAGE + EE_MAXT * EE_TP) * 4 <= 232448, "shared memory budget");

struct EdgeEmbedArgs {
  const int64_t* aa;        // [N, L] (already UNK where the sequence is hidden)
  const int64_t* res_nb;    // [N, L]
  const int64_t* chain_nb;  // [N, L]
  const float* pos;         // [N, L, A_in, 3]
  const uint8_t* mask;      // [N, L, A_in]
  const uint8_t* smask;     // [N, L] structure mask or nullptr
  const float* csp;         // [484, 225]
  const float* t_aa;        // [484, 64]
  const float* t_rel;       // [65, 64]
  const float* wd1t; const float* wd2t; const float* wo1dt; const float* wo1ht; const float* wo2t; const float* wo3t;
  const float* bd1; const float* bd2; const float* bo1; const float* bo2; const float* bo3;
  float* out;               // [N, L, L, 64]
  int N, L, A_in;
};

// y[n] += x * w[n], n = 0..63: w is one row of a [K][64] matrix in shared memory (warp-wide broadcast reads)
__device__ __forceinline__ void ee_axpy64(float (&y)[EE_F], float x, const float* __restrict__ w) {
#pragma unroll
  for (int n = 0; n < EE_F; n += 4) {
    const float4 v = *reinterpret_cast<const float4*>(w + n);
    y[n] = fmaf(x, v.x, y[n]); y[n + 1] = fmaf(x, v.y, y[n + 1]);
    y[n + 2] = fmaf(x, v.z, y[n + 2]); y[n + 3] = fmaf(x, v.w, y[n + 3]);
  }
}

__global__ void __launch_bounds__(EE_MAXT, 1) edge_embed_kernel(EdgeEmbedArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, T = blockDim.x;
  const int L = a.L;
  // ---- weights: once per CTA
  for (int i = tid; i < EE_AA * EE_F; i += T) sm[EE_OFF_WD1 + i] = a.wd1t[i];
  for (int i = tid; i < EE_F * EE_F; i += T) {
    sm[EE_OFF_WD2 + i] = a.wd2t[i]; sm[EE_OFF_WO1D + i] = a.wo1dt[i];
    sm[EE_OFF_WO2 + i] = a.wo2t[i]; sm[EE_OFF_WO3 + i] = a.wo3t[i];
  }
  for (int i = tid; i < EE_NH * EE_F; i += T) sm[EE_OFF_WO1H + i] = a.wo1ht[i];
  for (int i = tid; i < EE_F; i += T) {
    sm[EE_OFF_B + i] = a.bd1[i]; sm[EE_OFF_B + EE_F + i] = a.bd2[i]; sm[EE_OFF_B + 2 * EE_F + i] = a.bo1[i];
    sm[EE_OFF_B + 3 * EE_F + i] = a.bo2[i]; sm[EE_OFF_B + 4 * EE_F + i] = a.bo3[i];
  }
  for (int i = tid; i < EE_NREL * EE_F; i += T) sm[EE_OFF_TREL + (i >> 6) * EE_TP + (i & 63)] = a.t_rel[i];
  float* s_c = sm + EE_OFF_C;
  float* s_taa = sm + EE_OFF_TAA;
  float* s_pi = sm + EE_OFF_PI;
  float* s_mi = s_pi + 3 * EE_A;
  float* s_stage = sm + EE_OFF_STAGE;
  float* s_pj = s_stage;                      // [T][45]
  float* s_mj = s_stage + T * 3 * EE_A;       // [T][15]

  for (int row = blockIdx.x; row < a.N * L; row += gridDim.x) {
    const int b = row / L;
    const int aai = (int)a.aa[row];
    const int64_t resi = a.res_nb[row], chi = a.chain_nb[row];
    const bool smi = a.smask ? (a.smask[row] != 0) : true;
    __syncthreads();                           // previous row: every thread is done with the per-row tables
    for (int i = tid; i < EE_NAA * EE_AA; i += T) s_c[i] = a.csp[(size_t)aai * EE_NAA * EE_AA + i];
    for (int i = tid; i < EE_NAA * EE_F; i += T) s_taa[(i >> 6) * EE_TP + (i & 63)] = a.t_aa[(size_t)aai * EE_NAA * EE_F + i];
    if (tid < 3 * EE_A) s_pi[tid] = a.pos[(size_t)row * a.A_in * 3 + tid];
    if (tid < EE_A) s_mi[tid] = a.mask[(size_t)row * a.A_in + tid] ? 1.f : 0.f;
    for (int j0 = 0; j0 < L; j0 += T) {
      const int nj = min(T, L - j0);
      __syncthreads();                         // staging region free (previous pass written out), tables visible
      for (int i = tid; i < nj * 3 * EE_A; i += T) {
        const int jl = i / (3 * EE_A), e = i - jl * 3 * EE_A;
        s_pj[i] = a.pos[((size_t)b * L + j0 + jl) * a.A_in * 3 + e];
      }
      for (int i = tid; i < nj * EE_A; i += T) {
        const int jl = i / EE_A, e = i - jl * EE_A;
        s_mj[i] = a.mask[((size_t)b * L + j0 + jl) * a.A_in + e] ? 1.f : 0.f;
      }
      __syncthreads();
      const int j = j0 + tid;
      const bool act = tid < nj;
      float y[EE_F];
      if (act) {
        const size_t rj = (size_t)b * L + j;
        const int aaj = (int)a.aa[rj];
        const float psm = (smi && (a.smask ? a.smask[rj] != 0 : true)) ? 1.f : 0.f;
        const float* pj = s_pj + tid * 3 * EE_A;
        const float* mj = s_mj + tid * EE_A;
        const float* crow = s_c + aaj * EE_AA;
        // ---- distance features -> first distance layer (225 -> 64)
        float h[EE_F];
#pragma unroll
        for (int n = 0; n < EE_F; ++n) h[n] = sm[EE_OFF_B + n];
        for (int bb = 0; bb < EE_A; ++bb) {
          const float xj = pj[3 * bb], yj = pj[3 * bb + 1], zj = pj[3 * bb + 2], mjb = mj[bb];
#pragma unroll 3
          for (int aa_ = 0; aa_ < EE_A; ++aa_) {
            const int k = aa_ * EE_A + bb;
            const float dx = s_pi[3 * aa_] - xj, dy = s_pi[3 * aa_ + 1] - yj, dz = s_pi[3 * aa_ + 2] - zj;
            const float d2 = (dx * dx + dy * dy + dz * dz) * 0.01f;          // (|.| / 10)^2
            const float g = __expf(-crow[k] * d2) * (s_mi[aa_] * mjb);
            ee_axpy64(h, g, sm + EE_OFF_WD1 + k * EE_F);
          }
        }
        // ---- second distance layer (64 -> 64), relu, structure mask
#pragma unroll
        for (int n = 0; n < EE_F; ++n) y[n] = sm[EE_OFF_B + EE_F + n];
#pragma unroll
        for (int k = 0; k < EE_F; ++k) ee_axpy64(y, fmaxf(h[k], 0.f), sm + EE_OFF_WD2 + k * EE_F);
        // ---- first output layer: lookups + distance part + dihedral part
        {
          int64_t rel = resi - a.res_nb[rj];
          rel = rel < -32 ? -32 : (rel > 32 ? 32 : rel);
          const float same = (chi == a.chain_nb[rj]) ? 1.f : 0.f;
          const float* ta = s_taa + aaj * EE_TP;
          const float* tr = sm + EE_OFF_TREL + (int)(rel + 32) * EE_TP;
#pragma unroll
          for (int n = 0; n < EE_F; ++n) h[n] = sm[EE_OFF_B + 2 * EE_F + n] + ta[n] + same * tr[n];
        }
#pragma unroll
        for (int k = 0; k < EE_F; ++k) ee_axpy64(h, fmaxf(y[k], 0.f) * psm, sm + EE_OFF_WO1D + k * EE_F);
        {
          // phi = dihedral(C_i, N_j, CA_j, C_j), psi = dihedral(N_i, CA_i, C_i, N_j); atoms N 0, CA 1, C 2
          const float ang[2] = {dihedral4(s_pi + 6, pj, pj + 3, pj + 6), dihedral4(s_pi, s_pi + 3, s_pi + 6, pj)};
          const float fr[6] = {1.f, 2.f, 3.f, 1.f, 1.f / 2.f, 1.f / 3.f};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float* w = sm + EE_OFF_WO1H + q * 13 * EE_F;
            ee_axpy64(h, ang[q] * psm, w);
#pragma unroll
            for (int f = 0; f < 6; ++f) {
              float sv, cv;
              sincosf(ang[q] * fr[f], &sv, &cv);
              ee_axpy64(h, sv * psm, w + (1 + f) * EE_F);
              ee_axpy64(h, cv * psm, w + (7 + f) * EE_F);
            }
          }
        }
        // ---- output layers 2 and 3, pair mask
#pragma unroll
        for (int n = 0; n < EE_F; ++n) y[n] = sm[EE_OFF_B + 3 * EE_F + n];
#pragma unroll
        for (int k = 0; k < EE_F; ++k) ee_axpy64(y, fmaxf(h[k], 0.f), sm + EE_OFF_WO2 + k * EE_F);
#pragma unroll
        for (int n = 0; n < EE_F; ++n) h[n] = sm[EE_OFF_B + 4 * EE_F + n];
#pragma unroll
        for (int k = 0; k < EE_F; ++k) ee_axpy64(h, fmaxf(y[k], 0.f), sm + EE_OFF_WO3 + k * EE_F);
        const float mp = s_mi[1] * mj[1];                                    // CA masks of both residues
#pragma unroll
        for (int n = 0; n < EE_F; ++n) y[n] = h[n] * mp;
      }
      __syncthreads();                         // everyone is done reading x_j / m_j: the region becomes the staging tile
      if (act) {
#pragma unroll
        for (int n = 0; n < EE_F; ++n) s_stage[tid * EE_TP + n] = y[n];
      }
      __syncthreads();
      float* dst = a.out + ((size_t)row * L + j0) * EE_F;
      for (int i = tid; i < nj * EE_F; i += T) dst[i] = s_stage[(i >> 6) * EE_TP + (i & 63)];
    }
  }
}

static size_t edge_embed_smem(int T) { return (size_t)(EE_OFF_STAGE + T * EE_TP) * sizeof(float); }

void embed_kernels_init() {
  cudaFuncSetAttribute(edge_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)edge_embed_smem(EE_MAXT));
}

}  // namespace pf

extern "C" int pf_edge_embed(const int64_t* aa, const int64_t* res_nb, const int64_t* chain_nb, const float* pos_atoms,
                             const uint8_t* mask_atoms, const uint8_t* structure_mask, const float* softplus_coef,
                             const float* t_aa, const float* t_rel, const float* wd1_t, const float* bd1,
                             const float* wd2_t, const float* bd2, const float* wo1d_t, const float* wo1h_t,
                             const float* bo1, const float* wo2_t, const float* bo2, const float* wo3_t,
                             const float* bo3, float* out, int N, int L, int atoms_in, void* stream) {
  using namespace pf;
  PF_REQUIRE(aa && res_nb && chain_nb && pos_atoms && mask_atoms && softplus_coef && t_aa && t_rel && wd1_t && bd1 &&
                 wd2_t && bd2 && wo1d_t && wo1h_t && bo1 && wo2_t && bo2 && wo3_t && bo3 && out, PF_ERR_NULL_POINTER);
  PF_REQUIRE(N >= 0 && L >= 0 && atoms_in >= EE_A, PF_ERR_BAD_SHAPE);
  if (N == 0 || L == 0) return PF_OK;
  // threads per CTA: the row of L pairs in ceil(L / 192) equal passes, rounded up to whole warps
  const int passes = (L + EE_MAXT - 1) / EE_MAXT;
  int T = ((L + passes - 1) / passes + 31) & ~31;
  if (T < 64) T = 64;
  EdgeEmbedArgs a{aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, softplus_coef, t_aa, t_rel,
                  wd1_t, wd2_t, wo1d_t, wo1h_t, wo2_t, wo3_t, bd1, bd2, bo1, bo2, bo3, out, N, L, atoms_in};
  const long long rows = (long long)N * L;
  const int grid = (int)(rows < num_sms() ? rows : num_sms());
  edge_embed_kernel<<<grid, T, edge_embed_smem(T), as_stream(stream)>>>(a);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

// =================================================================================================
// Once-per-sample residue embedder: NodeEmbedder.forward, models_con/node.py:35-105.
//   crd  = R_i^T (x_ia - CA_i) of the 15 kept atoms in the residue's own frame (geometry.py:89-111, 136-155), placed in
//          the 45-wide slot of the residue's amino-acid type inside a 22 x 45 feature - i.e. only 45 columns of the
//          first Linear matter, W1[:, 128 + aa * 45 ...]
//   dih  = AngularEncoding of the backbone omega / phi / psi with their terminus masks (geometry.py:355-390)
//   y    = MLP(1157 -> 256 -> 128 -> 128 -> 128) of [aatype_embed[aa] | crd slot | dih], times the CA mask
// The embedding lookup is pushed through the first layer on the host (T1 = aatype_embed W1[:, :128]^T + b1); all
// weight matrices arrive transposed so that thread n reads column n with unit stride.  A CTA embeds NE_R residues
// so every weight element it fetches from L2 is used NE_R times.
namespace pf {

constexpr int NE_R = 8;          // residues per CTA
constexpr int NE_T = 256;        // threads = width of the first hidden layer

struct NodeEmbedArgs {
  const int64_t* aa; const int64_t* res_nb; const int64_t* chain_nb;
  const float* pos; const uint8_t* mask; const uint8_t* smask;
  const float* t1;      // [22, 256]
  const float* w1c;     // [22 * 45, 256]
  const float* w1d;     // [39, 256]
  const float* w2t; const float* b2;   // [256, 128]
  const float* w3t; const float* b3;   // [128, 128]
  const float* w4t; const float* b4;   // [128, 128]
  float* out;           // [N, L, 128]
  int N, L, A_in;
};

__global__ void __launch_bounds__(NE_T) node_embed_kernel(NodeEmbedArgs a) {
  __shared__ float s_feat[NE_R][84];      // 45 local coordinates | 39 dihedral features
  __shared__ int s_aa[NE_R];
  __shared__ float s_mres[NE_R];
  __shared__ float s_y1[NE_R][256];
  __shared__ float s_y2[NE_R][128];
  __shared__ float s_y3[NE_R][128];
  const int tid = threadIdx.x, L = a.L, M = a.N * L;
  const int r0 = blockIdx.x * NE_R;
  // ---- features: thread (r, e) for e < 45 coordinates / 3 dihedrals
  for (int i = tid; i < NE_R * 48; i += NE_T) {
    const int r = i / 48, e = i - r * 48, row = r0 + r;
    if (row >= M) {                                              // ragged last CTA: inert rows
      if (e < 45) s_feat[r][e] = 0.f;
      else for (int q = 0; q < 13; ++q) s_feat[r][45 + (e - 45) * 13 + q] = 0.f;
      if (e == 0) { s_aa[r] = 0; s_mres[r] = 0.f; }
      continue;
    }
    const int j = row % L;
    const float* P = a.pos + (size_t)row * a.A_in * 3;
    const uint8_t* mk = a.mask + (size_t)row * a.A_in;
    const bool sm = a.smask ? a.smask[row] != 0 : true;
    if (e < 45) {
      const int at = e / 3, c = e - at * 3;
      // construct_3d_basis(CA, C, N): e1 along CA->C, e2 = CA->N orthogonalised, e3 = e1 x e2  (eps 1e-6 as the reference)
      const float cx = P[3], cy = P[4], cz = P[5];
      float ax = P[6] - cx, ay = P[7] - cy, az = P[8] - cz;
      float inv = 1.0f / (sqrtf(ax * ax + ay * ay + az * az) + 1e-6f);
      ax *= inv; ay *= inv; az *= inv;
      float bx = P[0] - cx, by = P[1] - cy, bz = P[2] - cz;
      const float d = ax * bx + ay * by + az * bz;
      bx -= d * ax; by -= d * ay; bz -= d * az;
      inv = 1.0f / (sqrtf(bx * bx + by * by + bz * bz) + 1e-6f);
      bx *= inv; by *= inv; bz *= inv;
      const float ex = ay * bz - az * by, ey = az * bx - ax * bz, ez = ax * by - ay * bx;
      const float qx = P[3 * at] - cx, qy = P[3 * at + 1] - cy, qz = P[3 * at + 2] - cz;
      const float v = c == 0 ? ax * qx + ay * qy + az * qz : (c == 1 ? bx * qx + by * qy + bz * qz : ex * qx + ey * qy + ez * qz);
      s_feat[r][e] = (mk[at] && sm) ? v : 0.f;
      if (e == 0) { s_aa[r] = (int)a.aa[row]; s_mres[r] = mk[1] ? 1.f : 0.f; }
    } else {
      const int which = e - 45;                                  // 0 omega, 1 phi, 2 psi
      const int64_t rn = a.res_nb[row], cn = a.chain_nb[row];
      bool on;                                                    // terminus masks, topology.py:5-24
      float ang = 0.f;
      if (which < 2) {
        on = j > 0;
        if (on) {
          const int64_t dr = rn - a.res_nb[row - 1];
          on = (dr == 1 || dr == -1) && cn == a.chain_nb[row - 1] && a.mask[(size_t)(row - 1) * a.A_in + 1] != 0;
          const float* Q = P - (size_t)a.A_in * 3;               // residue j - 1
          ang = which == 0 ? dihedral4(Q + 3, Q + 6, P, P + 3) : dihedral4(Q + 6, P, P + 3, P + 6);
        }
      } else {
        on = j < L - 1;
        if (on) {
          const int64_t dr = a.res_nb[row + 1] - rn;
          on = (dr == 1 || dr == -1) && cn == a.chain_nb[row + 1] && mk[1] != 0;
          ang = dihedral4(P, P + 3, P + 6, P + (size_t)a.A_in * 3);
        }
      }
      // structure mask of the residue and both neighbours, with torch.roll's wrap-around (node.py:90-96)
      bool dm = true;
      if (a.smask) {
        const size_t base = (size_t)(row - j);
        dm = sm && a.smask[base + (j + L - 1) % L] != 0 && a.smask[base + (j + 1) % L] != 0;
      }
      const float keep = (on && dm) ? 1.f : 0.f;
      ang = on ? ang : 0.f;
      float* f = &s_feat[r][45 + which * 13];
      const float fr[6] = {1.f, 2.f, 3.f, 1.f, 1.f / 2.f, 1.f / 3.f};
      f[0] = ang * keep;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        float sv, cv;
        sincosf(ang * fr[q], &sv, &cv);
        f[1 + q] = sv * keep;
        f[7 + q] = cv * keep;
      }
    }
  }
  __syncthreads();
  // ---- layer 1 (212 effective inputs -> 256), ReLU: thread = output column
  {
    float acc[NE_R];
#pragma unroll
    for (int r = 0; r < NE_R; ++r) acc[r] = a.t1[s_aa[r] * 256 + tid];
    for (int k = 0; k < 45; ++k) {
#pragma unroll
      for (int r = 0; r < NE_R; ++r) acc[r] = fmaf(s_feat[r][k], a.w1c[(size_t)(s_aa[r] * 45 + k) * 256 + tid], acc[r]);
    }
    for (int k = 0; k < 39; ++k) {
      const float w = a.w1d[k * 256 + tid];
#pragma unroll
      for (int r = 0; r < NE_R; ++r) acc[r] = fmaf(s_feat[r][45 + k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < NE_R; ++r) s_y1[r][tid] = fmaxf(acc[r], 0.f);
  }
  __syncthreads();
  // ---- layers 2-4: 128 output columns; the two thread halves take four residues each
  const int col = tid & 127, rb = (tid >> 7) * (NE_R / 2);
  {
    float acc[NE_R / 2];
#pragma unroll
    for (int r = 0; r < NE_R / 2; ++r) acc[r] = a.b2[col];
    for (int k = 0; k < 256; ++k) {
      const float w = a.w2t[k * 128 + col];
#pragma unroll
      for (int r = 0; r < NE_R / 2; ++r) acc[r] = fmaf(s_y1[rb + r][k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < NE_R / 2; ++r) s_y2[rb + r][col] = fmaxf(acc[r], 0.f);
  }
  __syncthreads();
  {
    float acc[NE_R / 2];
#pragma unroll
    for (int r = 0; r < NE_R / 2; ++r) acc[r] = a.b3[col];
    for (int k = 0; k < 128; ++k) {
      const float w = a.w3t[k * 128 + col];
#pragma unroll
      for (int r = 0; r < NE_R / 2; ++r) acc[r] = fmaf(s_y2[rb + r][k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < NE_R / 2; ++r) s_y3[rb + r][col] = fmaxf(acc[r], 0.f);
  }
  __syncthreads();
  {
    float acc[NE_R / 2];
#pragma unroll
    for (int r = 0; r < NE_R / 2; ++r) acc[r] = a.b4[col];
    for (int k = 0; k < 128; ++k) {
      const float w = a.w4t[k * 128 + col];
#pragma unroll
      for (int r = 0; r < NE_R / 2; ++r) acc[r] = fmaf(s_y3[rb + r][k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < NE_R / 2; ++r) {
      const int row = r0 + rb + r;
      if (row < M) a.out[(size_t)row * 128 + col] = acc[r] * s_mres[rb + r];
    }
  }
}

}  // namespace pf

extern "C" int pf_node_embed(const int64_t* aa, const int64_t* res_nb, const int64_t* chain_nb, const float* pos_atoms,
                             const uint8_t* mask_atoms, const uint8_t* structure_mask, const float* t1,
                             const float* w1c_t, const float* w1d_t, const float* w2_t, const float* b2,
                             const float* w3_t, const float* b3, const float* w4_t, const float* b4, float* out,
                             int N, int L, int atoms_in, void* stream) {
  using namespace pf;
  PF_REQUIRE(aa && res_nb && chain_nb && pos_atoms && mask_atoms && t1 && w1c_t && w1d_t && w2_t && b2 && w3_t && b3 &&
                 w4_t && b4 && out, PF_ERR_NULL_POINTER);
  PF_REQUIRE(N >= 0 && L >= 0 && atoms_in >= EE_A, PF_ERR_BAD_SHAPE);
  if (N == 0 || L == 0) return PF_OK;
  NodeEmbedArgs a{aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, t1, w1c_t, w1d_t, w2_t, b2, w3_t, b3,
                  w4_t, b4, out, N, L, atoms_in};
  const long long rows = (long long)N * L;
  node_embed_kernel<<<(unsigned)((rows + NE_R - 1) / NE_R), NE_T, 0, as_stream(stream)>>>(a);
  PF_CHECK_LAUNCH();
  return PF_OK;
}
