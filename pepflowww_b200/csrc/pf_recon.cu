// Post-sampling reconstruction: the step right after FlowModel.sample in the reference's
// sampling scripts (models_con/sample.py:46,77,105-108), and its inverse on the data side.
//
//   pf_full_atom_reconstruction  models_con/torsion.py:140-226 (+ _make_psi_chi_rotation_matrices :68-96,
//                                _get_rigid_group :99-113, get_heavyatom_mask :126-138): backbone frames + 5 torsions +
//                                residue types -> atom14 coordinates, the psi / chi1-4 frames, heavy-atom mask
//   pf_reconstruct_backbone      pepflow/modules/common/geometry.py:446-489: backbone frames + residue types ->
//                                N, CA, C, O (psi measured on the rebuilt backbone of residues i and i + 1)
//   pf_torsion_angles            models_con/torsion.py:13-66 (get_torsion_angle: psi + chi1-4 and their mask from
//                                atom14 coordinates) - what the dataset builder stores as torsion_angle / _mask
//
// All three are O(residues) and HBM bound: 76 B in, 183 - 471 B out per residue.  The first version issued one strided
// scalar load per thread and element (226 M L1 sectors for 8 M sectors of data: L1TEX 95 % busy, DRAM 23 %); now a CTA
// takes a tile of 128 consecutive residues, moves its inputs into shared memory with flat 16-byte coalesced copies, keeps
// the per-residue-type constant tables in shared memory, lets one thread compose the frame chain of one residue in
// registers (products associated exactly as the reference's compose_chain does it: the last two factors first) and drop
// the atoms of each frame into a staging row as soon as the frame exists, and writes the tile out with flat 16-byte
// coalesced copies.  (A middle version that parked all six frames in shared memory and computed one output float per
// thread fixed the sector counts but doubled the instruction count - 400 M warp instructions, 0.67 ms; index
// arithmetic per output float is what a kernel this light cannot afford.)
#include "pf_common.cuh"
#include "pf_geom.cuh"

namespace pf {

constexpr int RC_T = 128;          // threads = residues per tile
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

// flat, coalesced global -> shared copy of `count` floats (16-byte vectors when both sides allow it)
__device__ __forceinline__ void stage_in(float* dst, const float* __restrict__ src, int count, int tid) {
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const int v = count >> 2;
    for (int e = tid; e < v; e += RC_T) reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(src)[e];
    for (int e = (v << 2) + tid; e < count; e += RC_T) dst[e] = src[e];
  } else {
    for (int e = tid; e < count; e += RC_T) dst[e] = src[e];
  }
}
__device__ __forceinline__ void stage_out(float* __restrict__ dst, const float* src, int count, int tid) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    const int v = count >> 2;
    for (int e = tid; e < v; e += RC_T) reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(src)[e];
    for (int e = (v << 2) + tid; e < count; e += RC_T) dst[e] = src[e];
  } else {
    for (int e = tid; e < count; e += RC_T) dst[e] = src[e];
  }
}

constexpr int RC_TROWS = RC_NAA + 1;   // table rows in shared memory: 0..20 residue types, 21 = "no rigid groups"

// The atoms of rigid-group frame f (0 backbone, 1 psi, 2..5 chi1..chi4) of one residue: q = R p + t into the staging row.
__device__ __forceinline__ void emit_atoms(const uint8_t* ord, const uint8_t* start, const float* apos, int f,
                                           const float* R, const float* t, float* out_row) {
  for (int j = start[f]; j < start[f + 1]; ++j) {
    const int s = ord[j];
    float q[3];
    rigid_apply(R, t, apos + s * 3, q);
    out_row[s * 3] = q[0]; out_row[s * 3 + 1] = q[1]; out_row[s * 3 + 2] = q[2];
  }
}

// FRAMES: the six frames of every residue are also wanted (R_ret / t_ret): they are staged in 36 KB of dynamic shared
// memory ([r][f][9] and [r][f][3], flat) and leave as coalesced copies like the atoms.  (Storing them straight from
// registers - 72 scalar stores per thread at a 216-byte stride - took 1.7 ms against 0.41 ms without frames.)
template <bool FRAMES>
__global__ void __launch_bounds__(RC_T) full_atom_kernel(FullAtomArgs a) {
  extern __shared__ __align__(16) float s_dyn[];
  float* const s_R = s_dyn;                             // FRAMES only: [RC_T][6][9]
  float* const s_t = s_dyn + RC_T * RC_FR * 9;          //              [RC_T][6][3]
  __shared__ __align__(16) float s_rot[RC_T * 9];       // staged inputs, flat (odd per-thread strides: no bank conflicts)
  __shared__ __align__(16) float s_tr[RC_T * 3];
  __shared__ __align__(16) float s_ang[RC_T * 5];
  __shared__ __align__(16) float s_out[RC_T * 42];      // staged atom14 rows, flat
  __shared__ float s_grp[RC_TROWS * 5 * RC_FW];         // rigid groups 3..7 (psi, chi1..chi4) per type: R | t
  __shared__ float s_apos[RC_TROWS * 14 * 3];
  __shared__ uint8_t s_ord[RC_TROWS * 14];              // atom slots of a type sorted by frame
  __shared__ uint8_t s_start[RC_TROWS * 8];             // first entry of frame f in s_ord (7 used)
  __shared__ uint8_t s_mtab[22 * 15 + 2];
  __shared__ int s_aa[RC_T];
  const int tid = threadIdx.x;
  // ---- constant tables: once per (persistent) CTA; row 21 = zero groups, every atom at the backbone origin
  for (int e = tid; e < RC_TROWS * 5 * RC_FW; e += RC_T) {
    const int aa = e / (5 * RC_FW), k = e - aa * (5 * RC_FW), g = 3 + k / RC_FW, c = k % RC_FW;
    s_grp[e] = aa >= RC_NAA ? 0.f : (c < 9 ? a.rigid_rot[(aa * 8 + g) * 9 + c] : a.rigid_trans[(aa * 8 + g) * 3 + (c - 9)]);
  }
  for (int e = tid; e < RC_TROWS * 14 * 3; e += RC_T) s_apos[e] = e < RC_NAA * 14 * 3 ? a.atom_pos[e] : 0.f;
  if (tid < RC_TROWS) {
    int n = 0;
    for (int f = 0; f < RC_FR; ++f) {
      s_start[tid * 8 + f] = (uint8_t)n;
      for (int s = 0; s < 14; ++s) {
        const int g = tid < RC_NAA ? a.atom_group[tid * 14 + s] : 0;
        if ((g < 3 ? 0 : g - 2) == f) s_ord[tid * 14 + n++] = (uint8_t)s;   // groups 0..2 carry the backbone frame
      }
    }
    s_start[tid * 8 + RC_FR] = (uint8_t)n;
  }
  if (a.mask_out)
    for (int e = tid; e < 22 * 15; e += RC_T) s_mtab[e] = a.mask_table[e];
  const long long tiles = (a.n + RC_T - 1) / RC_T;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long base = tile * RC_T;
    const int rows = (int)min((long long)RC_T, a.n - base);
    __syncthreads();   // tables visible / previous tile's readers done
    stage_in(s_rot, a.rot + base * 9, rows * 9, tid);
    stage_in(s_tr, a.trans + base * 3, rows * 3, tid);
    stage_in(s_ang, a.angles + base * 5, rows * 5, tid);
    if (tid < rows) {
      const long long aa64 = a.aa[base + tid];
      // 0..20: table rows; 21 (PAD): no rigid groups, zero mask row; anything else: nothing at all
      s_aa[tid] = (aa64 >= 0 && aa64 < 22) ? (int)aa64 : -1;
    }
    __syncthreads();
    if (tid < rows) {
      const int aa = (s_aa[tid] >= 0 && s_aa[tid] < RC_NAA) ? s_aa[tid] : RC_NAA;
      const uint8_t* ord = s_ord + aa * 14;
      const uint8_t* start = s_start + aa * 8;
      const float* apos = s_apos + aa * 42;
      const float* grp = s_grp + aa * 5 * RC_FW;
      float* row = s_out + tid * 42;
      float R0[9], t0[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) R0[k] = s_rot[tid * 9 + k];
#pragma unroll
      for (int k = 0; k < 3; ++k) t0[k] = s_tr[tid * 3 + k];
      auto put_frame = [&](int f, const float* R, const float* t) {
        if (FRAMES) {
#pragma unroll
          for (int k = 0; k < 9; ++k) s_R[(tid * RC_FR + f) * 9 + k] = R[k];
#pragma unroll
          for (int k = 0; k < 3; ++k) s_t[(tid * RC_FR + f) * 3 + k] = t[k];
        }
      };
      emit_atoms(ord, start, apos, 0, R0, t0, row);
      put_frame(0, R0, t0);
      float Rc[9], tc[3];
      // psi (group 3) and chi1 (group 4) hang off the backbone frame, chi_k off chi_{k-1}
      torsion_frame(R0, t0, grp, grp + 9, s_ang[tid * 5], Rc, tc);
      emit_atoms(ord, start, apos, 1, Rc, tc, row);
      put_frame(1, Rc, tc);
      torsion_frame(R0, t0, grp + RC_FW, grp + RC_FW + 9, s_ang[tid * 5 + 1], Rc, tc);
      emit_atoms(ord, start, apos, 2, Rc, tc, row);
      put_frame(2, Rc, tc);
#pragma unroll
      for (int f = 3; f < RC_FR; ++f) {
        float Rn[9], tn[3];
        torsion_frame(Rc, tc, grp + (f - 1) * RC_FW, grp + (f - 1) * RC_FW + 9, s_ang[tid * 5 + f - 1], Rn, tn);
#pragma unroll
        for (int k = 0; k < 9; ++k) Rc[k] = Rn[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) tc[k] = tn[k];
        emit_atoms(ord, start, apos, f, Rc, tc, row);
        put_frame(f, Rc, tc);
      }
    }
    __syncthreads();
    stage_out(a.pos14 + base * 42, s_out, rows * 42, tid);
    if (FRAMES) {
      if (a.R_ret) stage_out(a.R_ret + base * (RC_FR * 9), s_R, rows * RC_FR * 9, tid);
      if (a.t_ret) stage_out(a.t_ret + base * (RC_FR * 3), s_t, rows * RC_FR * 3, tid);
    }
    if (a.mask_out) {
      // four mask bytes per thread and store; base * 15 is a multiple of 4 (base is a multiple of 128)
      const int words = (rows * 15 + 3) / 4;
      for (int w = tid; w < words; w += RC_T) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = w * 4 + j;
          if (e < rows * 15) {
            const int r = e / 15, aa = s_aa[r];
            v |= (uint32_t)(aa >= 0 ? s_mtab[aa * 15 + (e - r * 15)] : 0) << (8 * j);
          }
        }
        if (w * 4 + 3 < rows * 15) {
          reinterpret_cast<uint32_t*>(a.mask_out + base * 15)[w] = v;
        } else {
          for (int j = 0; j < 4 && w * 4 + j < rows * 15; ++j) a.mask_out[base * 15 + w * 4 + j] = (uint8_t)(v >> (8 * j));
        }
      }
    }
  }
}

__device__ __forceinline__ int clamp_aa(int64_t v) { return v < 0 ? 0 : (v > 20 ? 20 : (int)v); }

struct BackboneArgs {
  const float* rot; const float* trans; const int64_t* aa; const int64_t* chain_nb; const int64_t* res_nb;
  const uint8_t* mask; const float* bb_coords; const float* bb_oxygen; float* pos_bb;
  int N, L;
};

__global__ void __launch_bounds__(RC_T) backbone_kernel(BackboneArgs a) {
  __shared__ __align__(16) float s_rot[(RC_T + 1) * 9 + 3];   // the tile's residues and the one after it
  __shared__ __align__(16) float s_tr[(RC_T + 1) * 3 + 1];
  __shared__ __align__(16) float s_out[RC_T * 12];
  __shared__ float s_bb[RC_NAA * 9];                          // N, CA, C in the backbone frame per residue type
  __shared__ float s_ox[RC_NAA * 3];                          // O in the psi frame
  const int tid = threadIdx.x;
  const long long n = (long long)a.N * a.L;
  const long long base = (long long)blockIdx.x * RC_T;
  const int rows = (int)min((long long)RC_T, n - base);
  const int rows_in = (int)min((long long)RC_T + 1, n - base);
  for (int e = tid; e < RC_NAA * 9; e += RC_T) s_bb[e] = a.bb_coords[e];
  if (tid < RC_NAA * 3) s_ox[tid] = a.bb_oxygen[tid];
  stage_in(s_rot, a.rot + base * 9, rows_in * 9, tid);
  stage_in(s_tr, a.trans + base * 3, rows_in * 3, tid);
  __syncthreads();
  if (tid < rows) {
    const long long r = base + tid;
    const int l = (int)(r % a.L);
    float R[9], t[3], bb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = s_rot[tid * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = s_tr[tid * 3 + k];
    const int aa = clamp_aa(a.aa[r]);                                     // geometry.py:462
#pragma unroll
    for (int k = 0; k < 3; ++k) rigid_apply(R, t, s_bb + aa * 9 + k * 3, bb + k * 3);
    // psi_i = dihedral(N_i, CA_i, C_i, N_{i+1}) when i+1 continues the chain (geometry.py:355-390, topology.py:5-24)
    float psi = 0.f;
    if (l + 1 < a.L) {
      long long d = a.res_nb[r + 1] - a.res_nb[r];
      d = d < 0 ? -d : d;
      if (d == 1 && a.chain_nb[r + 1] == a.chain_nb[r] && a.mask[r]) {
        const int aa1 = clamp_aa(a.aa[r + 1]);
        float n1[3];
        rigid_apply(s_rot + (tid + 1) * 9, s_tr + (tid + 1) * 3, s_bb + aa1 * 9, n1);
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
    rigid_apply(M, t, s_ox + aa * 3, o);
#pragma unroll
    for (int k = 0; k < 9; ++k) s_out[tid * 12 + k] = bb[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) s_out[tid * 12 + 9 + k] = o[k];
  }
  __syncthreads();
  stage_out(a.pos_bb + base * 12, s_out, rows * 12, tid);
}

// ---- torsions from coordinates -------------------------------------------------------------------------------------
constexpr int TA_MAXA = 15;                 // atoms per residue the staging tile can hold

struct TorsionArgs {
  const float* pos; const int64_t* aa; const int32_t* chi_atoms; float* torsion; uint8_t* mask;
  long long n; int A;
};

__global__ void __launch_bounds__(RC_T) torsion_angles_kernel(TorsionArgs a) {
  __shared__ __align__(16) float s_pos[RC_T * TA_MAXA * 3];
  __shared__ int8_t s_chi[RC_NAA * 16];
  __shared__ __align__(16) float s_val[RC_T * 5];
  __shared__ uint8_t s_msk[RC_T * 5];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * RC_T;
  const int rows = (int)min((long long)RC_T, a.n - base);
  const int W = a.A * 3;
  for (int e = tid; e < RC_NAA * 16; e += RC_T) s_chi[e] = (int8_t)a.chi_atoms[e];
  stage_in(s_pos, a.pos + base * W, rows * W, tid);
  __syncthreads();
  if (tid < rows) {
    const long long aa64 = a.aa[base + tid];
    const float* P = s_pos + tid * W;
    float v[5];
    bool ok[5];
    if (aa64 >= 0 && aa64 < 20) {                                   // torsion.py:52: 0..19 only
      v[0] = dihedral4_raw(P, P + 3, P + 6, P + 9);                 // "af style psi": N, CA, C, O (:44-45)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int8_t* q = s_chi + (int)aa64 * 16 + i * 4;
        v[1 + i] = q[0] >= 0 ? dihedral4_raw(P + q[0] * 3, P + q[1] * 3, P + q[2] * 3, P + q[3] * 3)
                             : __int_as_float(0x7f800000);          // no such angle: +inf (:32)
      }
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        ok[i] = isfinite(v[i]);                                     // :56 isfinite, :60 nan_to_num(posinf=0), :63 % 2 pi
        v[i] = ok[i] ? mod_2pi(v[i]) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 5; ++i) { v[i] = 0.f; ok[i] = false; }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) { s_val[tid * 5 + i] = v[i]; s_msk[tid * 5 + i] = ok[i] ? 1 : 0; }
  }
  __syncthreads();
  stage_out(a.torsion + base * 5, s_val, rows * 5, tid);
  for (int e = tid; e < rows * 5; e += RC_T) a.mask[base * 5 + e] = s_msk[e];
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
  PF_REQUIRE(!mask_out || (reinterpret_cast<uintptr_t>(mask_out) & 3u) == 0, PF_ERR_MISALIGNED);
  FullAtomArgs a{rot, trans, angles, aa, rigid_rot, rigid_trans, atom_group, atom_pos, heavyatom_mask_table,
                 pos14, R_ret, t_ret, mask_out, n};
  // persistent CTAs (5 fit on an SM next to their 41 KB of shared memory, 2 with the frame staging): the constant
  // tables are staged once per CTA
  const long long tiles = (n + RC_T - 1) / RC_T;
  if (R_ret || t_ret) {
    const int dyn = RC_T * RC_FR * RC_FW * (int)sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(full_atom_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return static_cast<int>(e);
    const unsigned grid = (unsigned)min(tiles, (long long)num_sms() * 2);
    full_atom_kernel<true><<<grid, RC_T, dyn, as_stream(stream)>>>(a);
  } else {
    const unsigned grid = (unsigned)min(tiles, (long long)num_sms() * 5);
    full_atom_kernel<false><<<grid, RC_T, 0, as_stream(stream)>>>(a);
  }
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

extern "C" int pf_torsion_angles(const float* pos_atoms, const int64_t* aa, const int32_t* chi_atoms, float* torsion,
                                 uint8_t* torsion_mask, long long n, int atoms_in, void* stream) {
  using namespace pf;
  PF_REQUIRE(n >= 0 && n <= (long long)RC_T * 0x7fffffffLL && atoms_in >= 14 && atoms_in <= TA_MAXA, PF_ERR_BAD_SHAPE);
  PF_REQUIRE(chi_atoms, PF_ERR_NULL_POINTER);
  if (n == 0) return PF_OK;
  PF_REQUIRE(pos_atoms && aa && torsion && torsion_mask, PF_ERR_NULL_POINTER);
  TorsionArgs a{pos_atoms, aa, chi_atoms, torsion, torsion_mask, n, atoms_in};
  torsion_angles_kernel<<<(unsigned)((n + RC_T - 1) / RC_T), RC_T, 0, as_stream(stream)>>>(a);
  PF_CHECK_LAUNCH();
  return PF_OK;
}
