// Post-sampling reconstruction: the step right after FlowModel.sample in the reference's
// sampling scripts (models_con/sample.py:46,77,105-108).
//
//   pf_full_atom_reconstruction  models_con/torsion.py:140-226 (+ _make_psi_chi_rotation_matrices :68-96,
//                                _get_rigid_group :99-113, get_heavyatom_mask :126-138): backbone frames + 5 torsions +
//                                residue types -> atom14 coordinates, the psi / chi1-4 frames, heavy-atom mask
//   pf_reconstruct_backbone      pepflow/modules/common/geometry.py:446-489: backbone frames + residue types ->
//                                N, CA, C, O (psi measured on the rebuilt backbone of residues i and i + 1)
//
// Both are O(residues) and HBM / latency bound: ~70 floats in, up to 129 floats out per residue.  A CTA takes 128
// consecutive residues: one thread composes the frame chain of one residue (the products are associated exactly as
// the reference's compose_chain does it: the last two factors first), parks the six frames in shared memory, then the
// whole CTA walks the flat output arrays so every HBM store is a run of consecutive addresses.
#include "pf_common.cuh"
#include "pf_geom.cuh"

namespace pf {

constexpr int RC_T = 128;          // threads = residues per CTA
constexpr int RC_FR = 6;           // frames kept per residue: backbone, psi, chi1..chi4
constexpr int RC_FW = 12;          // floats per frame: R (9, row-major) | t (3)
constexpr int RC_NAA = 21;         // rows of the rigid-group tables (20 residue types + UNK)

struct FullAtomArgs {
  const float* rot; const float* trans; const float* angles; const int64_t* aa;
  const float* rigid_rot; const float* rigid_trans; const int32_t* atom_group; const float* atom_pos;
  const uint8_t* mask_table;
  float* pos14; float* R_ret; float* t_ret; uint8_t* mask_out;
  long long n;
};

// (R, t) <- (Rp, tp) o (Rg, tg) o (Rx(angle), 0), associated right to left like compose_chain (geometry.py:183-189):
// first Rg Rx, then Rp (Rg Rx) and Rp tg + tp.
__device__ __forceinline__ void torsion_frame(const float* Rp, const float* tp, const float* Rg, const float* tg,
                                              float ang, float* R, float* t) {
  float s, c;
  sincosf(ang, &s, &c);
  float M[9];   // Rg Rx,  Rx = [[1,0,0],[0,c,-s],[0,s,c]]
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    M[i * 3] = Rg[i * 3];
    M[i * 3 + 1] = fmaf(Rg[i * 3 + 2], s, Rg[i * 3 + 1] * c);
    M[i * 3 + 2] = fmaf(Rg[i * 3 + 2], c, -(Rg[i * 3 + 1] * s));
  }
  mat3_mul(Rp, M, R);
  rigid_apply(Rp, tp, tg, t);
}

__global__ void __launch_bounds__(RC_T) full_atom_kernel(FullAtomArgs a) {
  __shared__ float s_fr[RC_T][RC_FR * RC_FW + 1];   // +1: residue rows on distinct banks
  __shared__ int s_aa[RC_T];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * RC_T;
  const int rows = (int)min((long long)RC_T, a.n - base);
  if (tid < rows) {
    const long long r = base + tid;
    const long long aa64 = a.aa[r];
    // rows outside the 21-row tables (PAD = 21, negatives) have no rigid groups: they collapse onto the backbone origin
    const int aa = (aa64 >= 0 && aa64 < RC_NAA) ? (int)aa64 : -1;
    s_aa[tid] = (aa64 >= 0 && aa64 < 22) ? (int)aa64 : -1;
    float* F = s_fr[tid];
#pragma unroll
    for (int k = 0; k < 9; ++k) F[k] = a.rot[r * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) F[9 + k] = a.trans[r * 3 + k];
    float ang[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ang[k] = a.angles[r * 5 + k];
    // rigid groups: 3 psi, 4..7 chi1..chi4; parents: psi and chi1 hang off the backbone, chi_k off chi_{k-1}
#pragma unroll
    for (int f = 1; f < RC_FR; ++f) {
      const int g = f + 2, parent = (f <= 2) ? 0 : f - 1;
      float Rg[9], tg[3];
      if (aa >= 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Rg[k] = __ldg(a.rigid_rot + ((size_t)aa * 8 + g) * 9 + k);
#pragma unroll
        for (int k = 0; k < 3; ++k) tg[k] = __ldg(a.rigid_trans + ((size_t)aa * 8 + g) * 3 + k);
      } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) Rg[k] = 0.f;
        tg[0] = tg[1] = tg[2] = 0.f;
      }
      float Rp[9], tp[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) Rp[k] = F[parent * RC_FW + k];
#pragma unroll
      for (int k = 0; k < 3; ++k) tp[k] = F[parent * RC_FW + 9 + k];
      torsion_frame(Rp, tp, Rg, tg, ang[f - 1], F + f * RC_FW, F + f * RC_FW + 9);
    }
  }
  __syncthreads();
  // ---- atom14 positions: one (residue, slot) per thread and trip
  for (int e = tid; e < rows * 14; e += RC_T) {
    const int r = e / 14, s = e - r * 14, aa = s_aa[r];
    float p[3] = {0.f, 0.f, 0.f};
    int f = 0;
    if (aa >= 0 && aa < RC_NAA) {
      const int g = __ldg(a.atom_group + aa * 14 + s);
      f = g < 3 ? 0 : g - 2;   // backbone / omega / phi groups all carry the backbone frame (torsion.py:214-215)
#pragma unroll
      for (int k = 0; k < 3; ++k) p[k] = __ldg(a.atom_pos + ((size_t)aa * 14 + s) * 3 + k);
    }
    float q[3];
    rigid_apply(s_fr[r] + f * RC_FW, s_fr[r] + f * RC_FW + 9, p, q);
    float* o = a.pos14 + (base * 14 + e) * 3;
    o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
  }
  if (a.R_ret)
    for (int e = tid; e < rows * RC_FR * 9; e += RC_T) {
      const int r = e / (RC_FR * 9), k = e - r * (RC_FR * 9);
      a.R_ret[base * (RC_FR * 9) + e] = s_fr[r][(k / 9) * RC_FW + k % 9];
    }
  if (a.t_ret)
    for (int e = tid; e < rows * RC_FR * 3; e += RC_T) {
      const int r = e / (RC_FR * 3), k = e - r * (RC_FR * 3);
      a.t_ret[base * (RC_FR * 3) + e] = s_fr[r][(k / 3) * RC_FW + 9 + k % 3];
    }
  if (a.mask_out)
    for (int e = tid; e < rows * 15; e += RC_T) {
      const int r = e / 15, s = e - r * 15, aa = s_aa[r];
      a.mask_out[base * 15 + e] = aa >= 0 ? __ldg(a.mask_table + aa * 15 + s) : (uint8_t)0;
    }
}

__device__ __forceinline__ int clamp_aa(int64_t v) { return v < 0 ? 0 : (v > 20 ? 20 : (int)v); }

struct BackboneArgs {
  const float* rot; const float* trans; const int64_t* aa; const int64_t* chain_nb; const int64_t* res_nb;
  const uint8_t* mask; const float* bb_coords; const float* bb_oxygen; float* pos_bb;
  int N, L;
};

__global__ void __launch_bounds__(RC_T) backbone_kernel(BackboneArgs a) {
  __shared__ float s_out[RC_T][13];
  const int tid = threadIdx.x;
  const long long n = (long long)a.N * a.L;
  const long long base = (long long)blockIdx.x * RC_T;
  const int rows = (int)min((long long)RC_T, n - base);
  if (tid < rows) {
    const long long r = base + tid;
    const int l = (int)(r % a.L);
    float R[9], t[3], bb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = a.rot[r * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = a.trans[r * 3 + k];
    const int aa = clamp_aa(a.aa[r]);   // geometry.py:462
#pragma unroll
    for (int k = 0; k < 3; ++k) rigid_apply(R, t, a.bb_coords + (size_t)aa * 9 + k * 3, bb + k * 3);
    // psi_i = dihedral(N_i, CA_i, C_i, N_{i+1}) when i+1 continues the chain (geometry.py:355-390, topology.py:5-24)
    float psi = 0.f;
    if (l + 1 < a.L) {
      long long d = a.res_nb[r + 1] - a.res_nb[r];
      d = d < 0 ? -d : d;
      if (d == 1 && a.chain_nb[r + 1] == a.chain_nb[r] && a.mask[r]) {
        const int aa1 = clamp_aa(a.aa[r + 1]);
        float R1[9], t1[3], n1[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) R1[k] = a.rot[(r + 1) * 9 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) t1[k] = a.trans[(r + 1) * 3 + k];
        rigid_apply(R1, t1, a.bb_coords + (size_t)aa1 * 9, n1);
        psi = dihedral4(bb, bb + 3, bb + 6, n1);
      }
    }
    // O = (R Rx(psi)) o + t
    float s, c;
    sincosf(psi, &s, &c);
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      M[i * 3] = R[i * 3];
      M[i * 3 + 1] = fmaf(R[i * 3 + 2], s, R[i * 3 + 1] * c);
      M[i * 3 + 2] = fmaf(R[i * 3 + 2], c, -(R[i * 3 + 1] * s));
    }
    float o[3];
    rigid_apply(M, t, a.bb_oxygen + (size_t)aa * 3, o);
#pragma unroll
    for (int k = 0; k < 9; ++k) s_out[tid][k] = bb[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) s_out[tid][9 + k] = o[k];
  }
  __syncthreads();
  for (int e = tid; e < rows * 12; e += RC_T) a.pos_bb[base * 12 + e] = s_out[e / 12][e % 12];
}

}  // namespace pf

extern "C" int pf_full_atom_reconstruction(const float* rot, const float* trans, const float* angles,
                                           const int64_t* aa, const float* rigid_rot, const float* rigid_trans,
                                           const int32_t* atom_group, const float* atom_pos,
                                           const uint8_t* heavyatom_mask_table, float* pos14, float* R_ret,
                                           float* t_ret, uint8_t* mask_out, long long n, void* stream) {
  using namespace pf;
  PF_REQUIRE(n >= 0 && n <= (long long)RC_T * 0x7fffffffLL, PF_ERR_BAD_SHAPE);
  PF_REQUIRE(rigid_rot && rigid_trans && atom_group && atom_pos, PF_ERR_NULL_POINTER);
  if (n == 0) return PF_OK;   // empty batch: the data pointers of empty tensors are NULL
  PF_REQUIRE(rot && trans && angles && aa && pos14, PF_ERR_NULL_POINTER);
  PF_REQUIRE(!mask_out || heavyatom_mask_table, PF_ERR_NULL_POINTER);
  FullAtomArgs a{rot, trans, angles, aa, rigid_rot, rigid_trans, atom_group, atom_pos, heavyatom_mask_table,
                 pos14, R_ret, t_ret, mask_out, n};
  full_atom_kernel<<<(unsigned)((n + RC_T - 1) / RC_T), RC_T, 0, as_stream(stream)>>>(a);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

extern "C" int pf_reconstruct_backbone(const float* rot, const float* trans, const int64_t* aa, const int64_t* chain_nb,
                                       const int64_t* res_nb, const uint8_t* mask, const float* bb_coords,
                                       const float* bb_oxygen, float* pos_bb, int N, int L, void* stream) {
  using namespace pf;
  PF_REQUIRE(N >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  PF_REQUIRE(bb_coords && bb_oxygen, PF_ERR_NULL_POINTER);
  if (N == 0 || L == 0) return PF_OK;
  PF_REQUIRE(rot && trans && aa && chain_nb && res_nb && mask && pos_bb, PF_ERR_NULL_POINTER);
  BackboneArgs a{rot, trans, aa, chain_nb, res_nb, mask, bb_coords, bb_oxygen, pos_bb, N, L};
  const long long n = (long long)N * L;
  backbone_kernel<<<(unsigned)((n + RC_T - 1) / RC_T), RC_T, 0, as_stream(stream)>>>(a);
  PF_CHECK_LAUNCH();
  return PF_OK;
}
