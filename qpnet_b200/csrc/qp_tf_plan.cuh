// Workspace plan of the teacher-forced (training) path: one bump-allocated layout shared
// by qp_forward (fills it) and qp_backward (consumes it).
#pragma once
#include "qp_common.cuh"
#include "qp_pack.cuh"

namespace qp {

struct TfPlan {
  PackedDims pd;
  int B, T, F, bl, M;
  int L0;                       // rows of the first block's input = rfA*M + rfF + bl
  int Lin[2 * QP_MAX_LAYERS];   // input rows of block l
  int shift[2 * QP_MAX_LAYERS]; // rows consumed by block l (dil or dil*M)
  int dil[2 * QP_MAX_LAYERS];
  int nmax;
  // device buffers
  int32_t* status;
  const float** tab;            // device copy of the parameter pointer table
  float** gtab;                 // device copy of the gradient pointer table
  PackedF32 W;                  // packed fp32 parameters
  float* X[2 * QP_MAX_LAYERS];  // block inputs (B, Lin[l], C)
  float* Z[2 * QP_MAX_LAYERS];  // gated outputs (B, n_l, C)
  float* G[2 * QP_MAX_LAYERS];  // (B, n_l, 2C) sigmoid/tanh values, QP_F_SAVE only
  int* pastrow[QP_MAX_LAYERS];  // adaptive blocks: (B, n_l) source row of the past tap
  float* Hup;                   // (B, L0, A)
  float* skipsum;               // (B, bl, S)
  float* H1;                    // (B, bl, S) pre-relu
  // backward scratch (QP_F_SAVE only)
  PackedF32 dW;
  float* dXa; float* dXb;       // ping-pong (B, L0, C)
  float* dgate;                 // (B, nmax, 2C)
  float* dskip;                 // (B, bl, S)
  float* dH1;                   // (B, bl, S)
  float* dHup;                  // (B, L0, A)
  float* dWp1; float* dbp1; float* dWp2; float* dbp2; float* dup;  // head + upsampler grads
  // bf16 tensor-core path (QP_F_BF16 only): packed weights and GEMM operand copies
  int Kgp;                      // gate K with the aux segment padded to 64: 2C + 64
  __nv_bfloat16* Wg_bf;         // (L, 2C, Kgp)
  __nv_bfloat16* Wrs_bf;        // (L, C+S, C)
  // the same weights as 64 x 64 swizzled blocks (tc::block_pack): K-major order for the forward GEMMs, MN-major order
  // for the backward GEMMs that contract against the un-transposed weights (MN copies with QP_F_SAVE only)
  __nv_bfloat16* WgK; __nv_bfloat16* WrsK; __nv_bfloat16* WgMN; __nv_bfloat16* WrsMN;
  __nv_bfloat16* W1_bf;         // (S, S)
  __nv_bfloat16* W2_bf;         // (Q, S)
  __nv_bfloat16* Xbf[2 * QP_MAX_LAYERS + 1];  // block inputs (B, Lin[l], C): one per block with QP_F_SAVE, else ping-pong
  __nv_bfloat16* Zbf[2 * QP_MAX_LAYERS];      // gated outputs (B, n_l, C): one per block with QP_F_SAVE, else shared
  __nv_bfloat16* Hup_bf;        // (B, L0, 64)
  __nv_bfloat16* skip_bf;       // (B, bl, S) relu(sum of skips)
  __nv_bfloat16* H1_bf;         // (B, bl, S) relu(head-1)
  int ones_col;                 // column of Hup_bf holding 1.0 (bias gradient of the gate GEMM), -1: none
  // bf16 backward operands (QP_F_BF16 | QP_F_SAVE)
  __nv_bfloat16* dgate_bf;      // (B, nmax, 2C)
  __nv_bfloat16* dX_bf;         // (B, L0, C)
  __nv_bfloat16* dskip_bf;      // (B, bl, S)
  float* dbskip;                // (S) column sum of dskip
};

// QPNET_BWD_TC bit mask (qp_backward.cu): which contractions of the bf16 backward run on tcgen05
int bwd_tc_mask();

// Fills `p`; returns total bytes.  `base` may be NULL (sizing only).
inline size_t make_tf_plan(const QpArch* a, int B, int T, int F, int bl, int M, uint32_t flags, void* base,
                           size_t cap, TfPlan* p) {
  PackedDims pd = packed_dims(a);
  p->pd = pd; p->B = B; p->T = T; p->F = F; p->bl = bl; p->M = M;
  int rfF = 0, rfA = 0;
  for (int i = 0; i < pd.nF; ++i) rfF += a->dil_fixed[i];
  for (int i = 0; i < pd.nA; ++i) rfA += a->dil_adaptive[i];
  p->L0 = rfA * M + rfF + bl;
  int L = p->L0; p->nmax = 0;
  for (int l = 0; l < pd.L; ++l) {
    p->dil[l] = l < pd.nF ? a->dil_fixed[l] : a->dil_adaptive[l - pd.nF];
    p->shift[l] = l < pd.nF ? p->dil[l] : p->dil[l] * M;
    p->Lin[l] = L;
    L -= p->shift[l];
    if (L > p->nmax) p->nmax = L;
  }
  const bool save = flags & QP_F_SAVE;
  Arena ar(base, cap);
  const int C = pd.C, S = pd.S, Q = pd.Q;
  p->status = ar.take<int32_t>(64);
  p->tab = ar.take<const float*>(tensor_map(a).count());
  p->gtab = ar.take<float*>(tensor_map(a).count());
  auto take_packed = [&](PackedF32& w) {
    w.Wg = ar.take<float>(pd.wg_elems() * pd.L);
    w.bg = ar.take<float>((size_t)2 * C * pd.L);
    w.Wrs = ar.take<float>(pd.wrs_elems() * pd.L);
    w.brs = ar.take<float>((size_t)(C + S) * pd.L);
    w.E0 = ar.take<float>((size_t)Q * C);
    w.E1 = ar.take<float>((size_t)Q * C);
  };
  take_packed(p->W);
  float* xping[2] = {nullptr, nullptr};
  float* zshare = nullptr;
  if (!save) {
    xping[0] = ar.take<float>((size_t)B * p->L0 * C);
    xping[1] = ar.take<float>((size_t)B * p->L0 * C);
    zshare = ar.take<float>((size_t)B * p->nmax * C);
  }
  for (int l = 0; l < pd.L; ++l) {
    int n = p->Lin[l] - p->shift[l];
    if (save) {
      p->X[l] = ar.take<float>((size_t)B * p->Lin[l] * C);
      p->Z[l] = ar.take<float>((size_t)B * n * C);
      p->G[l] = ar.take<float>((size_t)B * n * 2 * C);
    } else {
      p->X[l] = xping[l & 1];
      p->Z[l] = zshare;
      p->G[l] = nullptr;
    }
    if (l >= pd.nF) p->pastrow[l - pd.nF] = ar.take<int>((size_t)B * n);
  }
  p->Hup = ar.take<float>((size_t)B * p->L0 * pd.A);
  p->skipsum = ar.take<float>((size_t)B * bl * S);
  p->H1 = ar.take<float>((size_t)B * bl * S);
  if (save) {
    take_packed(p->dW);
    p->dXa = ar.take<float>((size_t)B * p->L0 * C);
    p->dXb = ar.take<float>((size_t)B * p->L0 * C);
    p->dgate = ar.take<float>((size_t)B * p->nmax * 2 * C);
    p->dskip = ar.take<float>((size_t)B * bl * S);
    p->dH1 = ar.take<float>((size_t)B * bl * S);
    p->dHup = ar.take<float>((size_t)B * p->L0 * pd.A);
    p->dWp1 = ar.take<float>((size_t)S * S);
    p->dbp1 = ar.take<float>(S);
    p->dWp2 = ar.take<float>((size_t)Q * S);
    p->dbp2 = ar.take<float>(Q);
    p->dup = ar.take<float>(pd.U + 1);
  }
  p->Kgp = 2 * C + 64;
  p->ones_col = -1;
  if (flags & QP_F_BF16) {
    p->Wg_bf = ar.take<__nv_bfloat16>((size_t)pd.L * 2 * C * p->Kgp);
    p->Wrs_bf = ar.take<__nv_bfloat16>((size_t)pd.L * (C + S) * C);
    p->WgK = ar.take<__nv_bfloat16>((size_t)pd.L * 2 * C * p->Kgp);
    p->WrsK = ar.take<__nv_bfloat16>((size_t)pd.L * (C + S) * C);
    p->WgMN = save ? ar.take<__nv_bfloat16>((size_t)pd.L * 2 * C * p->Kgp) : nullptr;
    p->WrsMN = save ? ar.take<__nv_bfloat16>((size_t)pd.L * (C + S) * C) : nullptr;
    p->W1_bf = ar.take<__nv_bfloat16>((size_t)S * S);
    p->W2_bf = ar.take<__nv_bfloat16>((size_t)Q * S);
    if (save) {
      for (int l = 0; l < pd.L; ++l) {
        p->Xbf[l] = ar.take<__nv_bfloat16>((size_t)B * p->Lin[l] * C);
        p->Zbf[l] = ar.take<__nv_bfloat16>((size_t)B * (p->Lin[l] - p->shift[l]) * C);
      }
      p->Xbf[pd.L] = p->Xbf[0];   // the (dead) residual output of the last block
      p->dgate_bf = ar.take<__nv_bfloat16>((size_t)B * p->nmax * 2 * C);
      p->dX_bf = ar.take<__nv_bfloat16>((size_t)B * p->L0 * C);
      p->dskip_bf = ar.take<__nv_bfloat16>((size_t)B * bl * S);
      p->dbskip = ar.take<float>(S);
    } else {
      __nv_bfloat16* xp0 = ar.take<__nv_bfloat16>((size_t)B * p->L0 * C);
      __nv_bfloat16* xp1 = ar.take<__nv_bfloat16>((size_t)B * p->L0 * C);
      __nv_bfloat16* zs = ar.take<__nv_bfloat16>((size_t)B * p->nmax * C);
      for (int l = 0; l <= pd.L; ++l) p->Xbf[l] = (l & 1) ? xp1 : xp0;
      for (int l = 0; l < pd.L; ++l) p->Zbf[l] = zs;
    }
    p->ones_col = pd.Ap < 64 ? pd.Ap : -1;
    p->Hup_bf = ar.take<__nv_bfloat16>((size_t)B * p->L0 * 64);
    p->skip_bf = ar.take<__nv_bfloat16>((size_t)B * bl * S);
    p->H1_bf = ar.take<__nv_bfloat16>((size_t)B * bl * S);
  }
  return align_up(ar.off, 256);
}

}  // namespace qp
