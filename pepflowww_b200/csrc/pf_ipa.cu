// K3: fused invariant-point-attention core (models_con/ipa_pytorch.py:393-473).
//
//   logit[h,i,j] = sqrt(1/(3C)) q_i.k_j + sqrt(1/3) (W_b z_ij + b_b)[h] - 0.5 hw_h sum_p |qp_ip - kp_jp|^2
//                  + 1e5 (m_i m_j - 1)
//   a = softmax_j(logit);  o = a v;  o_pt = R_i^T (a v_pts - t_i), |o_pt|;  o_pair = W_dz (sum_j a z_ij) + b_dz
//
// (the last identity uses sum_j a = 1: the a-weighted RAW pair row is accumulated and down_z is applied
// once per (i,h) in the epilogue instead of once per pair - SURVEY.md App. F, 6e-7.)
// Nothing of size L^2 x H is written to HBM: z is streamed, everything else stays on chip.
//
// Variant 0 (this file, "ipa_impl" = 0): CUDA-core kernel.  CTA = (TI query rows, complex); warp = head.
//   pass 1: lane = key j: logits into smem, running max;  pass 2: exp / sum;
//   pass 3: lane = channel: o (4 ch/lane), o_pt (36 values), a-weighted z (2 ch/lane); epilogue.
#include "pf_common.cuh"

namespace pf {

constexpr int TI = 4;  // query rows per CTA

__global__ void __launch_bounds__(256) ipa_attention_v0_kernel(IpaArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int L = a.L;
  const int Lp = (L + 3) & ~3;
  float* sQ = smem;                         // [TI][H][128]
  float* sQP = sQ + TI * H * C;             // [TI][H][24]
  float* sWb = sQP + TI * H * PQ * 3;       // [H][64]
  float* sAcc = sWb + H * CZ;               // [H warps][TI][64 + 36] epilogue scratch
  float* sLog = sAcc + H * TI * 100;        // [H][TI][Lp]
  const int b = blockIdx.y, i0 = blockIdx.x * TI;
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
  const size_t rowb = (size_t)b * L;

  // stage q rows, q points, W_b
  for (int idx = tid; idx < TI * H * C; idx += 256) {
    const int r = idx / (H * C), c = idx % (H * C);
    const int i = i0 + r;
    sQ[idx] = (i < L) ? a.proj[(rowb + i) * NPROJ + OFF_Q + c] : 0.f;
  }
  for (int idx = tid; idx < TI * H * PQ * 3; idx += 256) {
    const int r = idx / (H * PQ * 3), rem = idx % (H * PQ * 3);
    const int hh = rem / (PQ * 3), e = rem % (PQ * 3);
    const int i = i0 + r;
    sQP[idx] = (i < L) ? a.pts[((rowb + i) * H + hh) * (NPT * 3) + e] : 0.f;
  }
  for (int idx = tid; idx < H * CZ; idx += 256) sWb[idx] = a.w_b[idx];
  __syncthreads();

  const float sc_qk = 0.05103103630798288f;   // sqrt(1/(3*128))
  const float sc_b = 0.5773502691896257f;     // sqrt(1/3)
  const float hw = a.head_w[h];
  const float bb = a.b_b[h];
  float mi[TI];
#pragma unroll
  for (int r = 0; r < TI; ++r) mi[r] = (i0 + r < L) ? a.mask[rowb + i0 + r] : 0.f;
  float* myLog = sLog + (size_t)h * TI * Lp;

  // ---- pass 1: logits
  float mx[TI];
#pragma unroll
  for (int r = 0; r < TI; ++r) mx[r] = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    const float* krow = a.proj + (rowb + j) * NPROJ + OFF_KV + h * 2 * C;
    float qk[TI];
#pragma unroll
    for (int r = 0; r < TI; ++r) qk[r] = 0.f;
#pragma unroll 8
    for (int c = 0; c < C; c += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(krow + c);
#pragma unroll
      for (int r = 0; r < TI; ++r) {
        const float4 q = *reinterpret_cast<const float4*>(sQ + (r * H + h) * C + c);
        qk[r] = fmaf(q.x, kv.x, qk[r]);
        qk[r] = fmaf(q.y, kv.y, qk[r]);
        qk[r] = fmaf(q.z, kv.z, qk[r]);
        qk[r] = fmaf(q.w, kv.w, qk[r]);
      }
    }
    const float* kp = a.pts + ((rowb + j) * H + h) * (NPT * 3) + PQ * 3;
    float d2[TI];
#pragma unroll
    for (int r = 0; r < TI; ++r) d2[r] = 0.f;
#pragma unroll
    for (int e = 0; e < PQ * 3; ++e) {
      const float kx = kp[e];
#pragma unroll
      for (int r = 0; r < TI; ++r) {
        const float d = sQP[(r * H + h) * PQ * 3 + e] - kx;
        d2[r] = fmaf(d, d, d2[r]);
      }
    }
    const float mj = a.mask[rowb + j];
#pragma unroll
    for (int r = 0; r < TI; ++r) {
      const int i = i0 + r;
      float bias = 0.f;
      if (i < L) {
        const float* zr = a.z + ((rowb + i) * L + j) * CZ;
#pragma unroll
        for (int c = 0; c < CZ; c += 4) {
          const float4 zv = *reinterpret_cast<const float4*>(zr + c);
          const float4 w = *reinterpret_cast<const float4*>(sWb + h * CZ + c);
          bias = fmaf(zv.x, w.x, bias);
          bias = fmaf(zv.y, w.y, bias);
          bias = fmaf(zv.z, w.z, bias);
          bias = fmaf(zv.w, w.w, bias);
        }
      }
      const float lg = qk[r] * sc_qk + sc_b * (bias + bb) - 0.5f * hw * d2[r] + 1e5f * (mi[r] * mj - 1.f);
      myLog[r * Lp + j] = lg;
      mx[r] = fmaxf(mx[r], lg);
    }
  }
  // ---- pass 2: softmax numerators
  float inv_sum[TI];
#pragma unroll
  for (int r = 0; r < TI; ++r) {
    const float m = warp_max(mx[r]);
    float s = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float p = expf(myLog[r * Lp + j] - m);
      myLog[r * Lp + j] = p;
      s += p;
    }
    inv_sum[r] = 1.0f / warp_sum(s);
  }
  __syncwarp();

  // ---- pass 3: weighted sums.  lane owns o[.., 4*lane..4*lane+3], zacc[.., 2*lane..], pt slots lane, lane+32
  float o[TI][4], zacc[TI][2], pt[TI][2];
#pragma unroll
  for (int r = 0; r < TI; ++r) {
    o[r][0] = o[r][1] = o[r][2] = o[r][3] = 0.f;
    zacc[r][0] = zacc[r][1] = 0.f;
    pt[r][0] = pt[r][1] = 0.f;
  }
  for (int j = 0; j < L; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(a.proj + (rowb + j) * NPROJ + OFF_KV + h * 2 * C + C + lane * 4);
    const float* vp = a.pts + ((rowb + j) * H + h) * (NPT * 3) + 2 * PQ * 3;
    const float p0 = vp[lane];
    const float p1 = (lane < 4) ? vp[lane + 32] : 0.f;
#pragma unroll
    for (int r = 0; r < TI; ++r) {
      const float p = myLog[r * Lp + j];
      o[r][0] = fmaf(p, v.x, o[r][0]);
      o[r][1] = fmaf(p, v.y, o[r][1]);
      o[r][2] = fmaf(p, v.z, o[r][2]);
      o[r][3] = fmaf(p, v.w, o[r][3]);
      pt[r][0] = fmaf(p, p0, pt[r][0]);
      pt[r][1] = fmaf(p, p1, pt[r][1]);
      const int i = i0 + r;
      if (i < L) {
        const float2 zv = *reinterpret_cast<const float2*>(a.z + ((rowb + i) * L + j) * CZ + lane * 2);
        zacc[r][0] = fmaf(p, zv.x, zacc[r][0]);
        zacc[r][1] = fmaf(p, zv.y, zacc[r][1]);
      }
    }
  }

  // ---- epilogue
  float* scr = sAcc + (size_t)h * TI * 100;
#pragma unroll
  for (int r = 0; r < TI; ++r) {
    const int i = i0 + r;
    if (i >= L) continue;
    float* f = a.feats + (rowb + i) * NFEAT;
    const float is = inv_sum[r];
    *reinterpret_cast<float4*>(f + h * C + lane * 4) = make_float4(o[r][0] * is, o[r][1] * is, o[r][2] * is, o[r][3] * is);
    scr[r * 100 + 2 * lane] = zacc[r][0] * is;
    scr[r * 100 + 2 * lane + 1] = zacc[r][1] * is;
    scr[r * 100 + 64 + lane] = pt[r][0] * is;
    if (lane < 4) scr[r * 100 + 64 + 32 + lane] = pt[r][1] * is;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < TI; ++r) {
    const int i = i0 + r;
    if (i >= L) continue;
    float* f = a.feats + (rowb + i) * NFEAT;
    if (lane < PV) {  // o_pt: global -> local frame, norms (ipa_pytorch.py:455-460)
      const float* R = a.rot + (rowb + i) * 9;
      const float* t = a.trans + (rowb + i) * 3;
      const float gx = scr[r * 100 + 64 + lane * 3 + 0] - t[0];
      const float gy = scr[r * 100 + 64 + lane * 3 + 1] - t[1];
      const float gz = scr[r * 100 + 64 + lane * 3 + 2] - t[2];
      const float lx = R[0] * gx + R[3] * gy + R[6] * gz;
      const float ly = R[1] * gx + R[4] * gy + R[7] * gz;
      const float lz = R[2] * gx + R[5] * gy + R[8] * gz;
      f[1024 + 0 * 96 + h * PV + lane] = lx;
      f[1024 + 1 * 96 + h * PV + lane] = ly;
      f[1024 + 2 * 96 + h * PV + lane] = lz;
      f[1024 + 3 * 96 + h * PV + lane] = sqrtf(lx * lx + ly * ly + lz * lz + 1e-8f);
    }
    if (lane >= 16) {  // o_pair: down_z applied to the a-weighted pair row (ipa_pytorch.py:469-473)
      const int d = lane - 16;
      float acc = a.b_dz[d];
      for (int c = 0; c < CZ; ++c) acc = fmaf(a.w_dz[d * CZ + c], scr[r * 100 + c], acc);
      f[1024 + 384 + h * 16 + d] = acc;
    }
  }
}

size_t ipa_v0_smem(int L) {
  const int Lp = (L + 3) & ~3;
  return (size_t)(TI * H * C + TI * H * PQ * 3 + H * CZ + H * TI * 100 + (size_t)H * TI * Lp) * sizeof(float);
}

int launch_ipa_attention(const IpaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (a.B == 0 || a.L == 0) return PF_OK;
  if (opt_ipa_impl() == 4) return launch_ipa_attention_v3(a, workspace, workspace_bytes, st, true);
  if (opt_ipa_impl() == 3) return launch_ipa_attention_v3(a, workspace, workspace_bytes, st, false);
  const size_t smem = ipa_v0_smem(a.L);
  if (smem > 227 * 1024) return PF_ERR_BAD_SHAPE;  // L <= ~1600
  dim3 grid((a.L + TI - 1) / TI, a.B);
  profile_begin(0, st);
  ipa_attention_v0_kernel<<<grid, 256, smem, st>>>(a);
  profile_end(0, st);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

void ipa_kernels_init() {
  cudaFuncSetAttribute(ipa_attention_v0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

}  // namespace pf

extern "C" size_t pf_ipa_attention_workspace_bytes(int B, int L) { return pf::ipa_workspace_bytes(B, L); }

extern "C" int pf_ipa_attention(const float* proj, const float* pts, const float* z, const float* w_b,
                                const float* b_b, const float* w_dz, const float* b_dz, const float* head_w,
                                const float* rot, const float* trans, const float* mask, float* feats,
                                void* workspace, size_t workspace_bytes, int B, int L, void* stream) {
  PF_REQUIRE(proj && pts && z && w_b && b_b && w_dz && b_dz && head_w && rot && trans && mask && feats,
             PF_ERR_NULL_POINTER);
  PF_REQUIRE(B >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  PF_REQUIRE(pf::aligned16(proj) && pf::aligned16(z) && pf::aligned16(feats) && pf::aligned16(pts), PF_ERR_MISALIGNED);
  pf::IpaArgs a{proj, pts, z, w_b, b_b, w_dz, b_dz, head_w, rot, trans, mask, feats, B, L};
  return pf::launch_ipa_attention(a, workspace, workspace_bytes, pf::as_stream(stream));
}
