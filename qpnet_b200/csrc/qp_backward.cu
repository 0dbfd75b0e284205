// Backward of the teacher-forced stack (what autograd does for the reference around
// qpnet_train.py:526-531).  Exact fp32 SIMT contractions after an fp32 forward; after a bf16 tensor-core forward
// (QP_F_BF16) the contractions of the residual blocks run on tcgen05 (qp_tc.cu: dz / dX GEMMs against the un-transposed
// weights, MN-major weight gradients) and the two head layers on the TF32 mma.sync kernels, which are also the
// QPNET_BWD_TC = 0 fallback of every contraction.  Consumes the workspace a QP_F_SAVE forward filled.
#include <stdlib.h>

#include <algorithm>

#include "qp_gemm_f32.cuh"
#include "qp_tc.cuh"
#include "qp_tf_plan.cuh"

namespace qp {

static Seg mk(const float* base, int64_t bstride, int ld, const int* rowmap, int row_off, int src_rows, int K,
              int relu = 0) {
  Seg s; s.base = base; s.bstride = bstride; s.ld = ld; s.rowmap = rowmap; s.row_off = row_off;
  s.src_rows = src_rows; s.K = K; s.relu = relu;
  return s;
}

// dE0[x[i]] += dX0[i], dE1[x[i+1]] += dX0[i]  (transpose of the table lookup in embed_kernel)
__global__ void embed_grad_kernel(const int64_t* __restrict__ x, int T, int L0, int C, int Q,
                                  const float* __restrict__ dX0, float* __restrict__ dE0, float* __restrict__ dE1) {
  int b = blockIdx.y, i = blockIdx.x;
  const int64_t* xb = x + (int64_t)b * T + (T - L0 - 1);
  int s0 = (int)(((xb[i] % Q) + Q) % Q), s1 = (int)(((xb[i + 1] % Q) + Q) % Q);
  const float* g = dX0 + ((int64_t)b * L0 + i) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = g[c];
    atomicAdd(dE0 + (int64_t)s0 * C + c, v);
    atomicAdd(dE1 + (int64_t)s1 * C + c, v);
  }
}

// out[c] = sum over rows of M[rows][C]   (one block per 32 columns)
__global__ void colsum_kernel(const float* __restrict__ M, int64_t rows, int C, float* __restrict__ out) {
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + (threadIdx.x & 31);
  int w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int64_t r = w; r < rows; r += 8) s += M[r * C + c];
  red[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    out[c] = t;
  }
}

// d(upsampling.conv.weight)[j] = sum dHup[b][i][a]*h[b][a][f];  d(bias) = sum dHup
__global__ void upsample_grad_kernel(const float* __restrict__ dHup, const float* __restrict__ h, int B, int A, int F,
                                     int U, int L0, float* __restrict__ dw, float* __restrict__ db) {
  // one block per tap j; deterministic tree reduction inside the block
  int j = blockIdx.x;
  float sw = 0.f, sb = 0.f;
  int p0 = F * U - L0;
  for (int b = 0; b < B; ++b) {
    // rows i with (p0 + i) % U == j
    int first = ((j - p0) % U + U) % U;
    for (int i = first; i < L0; i += U) {
      int f = (p0 + i) / U;
      for (int a = threadIdx.x; a < A; a += blockDim.x) {
        float g = dHup[((int64_t)b * L0 + i) * A + a];
        sw += g * h[((int64_t)b * A + a) * F + f];
        sb += g;
      }
    }
  }
  __shared__ float r0[64], r1[64];
  r0[threadIdx.x] = sw; r1[threadIdx.x] = sb;
  __syncthreads();
  for (int o = 32; o; o >>= 1) {
    if ((int)threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { dw[j] = r0[0]; atomicAdd(db, r1[0]); }
}

// Which contractions of the bf16 training path's backward run on tcgen05 (qp_tc.cu), bit mask, default all:
// 1 weight gradients, 2 dz -> dgate GEMM, 4 dX GEMM.  QPNET_BWD_TC=0 is the TF32 mma.sync backward (A/B timing, bring-up).
int bwd_tc_mask() {
  static int m = -1;
  if (m < 0) { const char* e = getenv("QPNET_BWD_TC"); m = e ? atoi(e) & 7 : 7; }
  return m;
}

// tc: TF32 tensor-core contractions (the backward of the bf16 training path); false: exact fp32 SIMT
// Stages: 0 = the two head layers (+ the buffers the blocks accumulate into), 1 .. L = residual block L - stage (last
// block first), L + 1 = causal layer and upsampler.  [s_begin, s_end) selects a range: a data-parallel caller runs the
// backward in a few ranges and starts all-reducing the gradients of a finished range while the next one computes; every
// stage leaves its gradients in the reference's tensors.
int tf_backward_f32(const QpArch* arch, const float* const* tensors, const int64_t* x, const float* h,
                    const TfPlan& p, const float* dlogits, float* const* grads, bool tc, cudaStream_t st, int s_begin, int s_end) {
  const PackedDims& pd = p.pd;
  const TensorMap tm = tensor_map(arch);
  const int C = pd.C, S = pd.S, Q = pd.Q, A = pd.A, B = p.B, L0 = p.L0, bl = p.bl, L = pd.L;
  const int ntens = tm.count();
  for (int i = 0; i < ntens; ++i)
    if (!grads[i]) return set_error(QP_EINVAL, "backward: gradient tensor %d is NULL", i);
  if (s_begin == 0) QP_CUDA(cudaMemcpyAsync((void*)p.gtab, grads, sizeof(float*) * ntens, cudaMemcpyHostToDevice, st));
  // bf16 operands on tcgen05 for the residual blocks (the head stays on the TF32 kernels: 2 % of the work)
  const int mask = (tc && p.ones_col >= 0) ? bwd_tc_mask() : 0;
  const bool tc_w = mask & 1, tc_dz = mask & 2, tc_dx = mask & 4;
  const int Kgp = p.Kgp;

  // ---- head -------------------------------------------------------------------------
  if (s_begin <= 0 && s_end > 0) {
    WgradArgs w = {};
    w.p[0] = mk(dlogits, (int64_t)bl * Q, Q, nullptr, 0, bl, Q); w.np = 1;
    w.q[0] = mk(p.H1, (int64_t)bl * S, S, nullptr, 0, bl, S, 1); w.nq = 1;
    w.B = B; w.n_rows = bl; w.I = Q; w.J = S; w.out = grads[tm.post2_w()]; w.ldo = S; w.colsum = grads[tm.post2_b()];
    if (int e = launch_wgrad(w, st, tc)) return e;
    GemmArgs g = {};
    g.seg[0] = mk(dlogits, (int64_t)bl * Q, Q, nullptr, 0, bl, Q); g.nseg = 1;
    g.W = tensors[tm.post2_w()]; g.ldw = S; g.w_kn = 1;
    g.B = B; g.n_rows = bl; g.N = S; g.out = p.dH1; g.out_bstride = (int64_t)bl * S; g.ldo = S;
    g.mask = p.H1; g.mask_bstride = (int64_t)bl * S; g.ldmask = S;
    if (int e = launch_gemm<EPI_PLAIN>(g, st, tc)) return e;
    WgradArgs w1 = {};
    w1.p[0] = mk(p.dH1, (int64_t)bl * S, S, nullptr, 0, bl, S); w1.np = 1;
    w1.q[0] = mk(p.skipsum, (int64_t)bl * S, S, nullptr, 0, bl, S, 1); w1.nq = 1;
    w1.B = B; w1.n_rows = bl; w1.I = S; w1.J = S; w1.out = grads[tm.post1_w()]; w1.ldo = S; w1.colsum = grads[tm.post1_b()];
    if (int e = launch_wgrad(w1, st, tc)) return e;
    GemmArgs g1 = {};
    g1.seg[0] = mk(p.dH1, (int64_t)bl * S, S, nullptr, 0, bl, S); g1.nseg = 1;
    g1.W = tensors[tm.post1_w()]; g1.ldw = S; g1.w_kn = 1;
    g1.B = B; g1.n_rows = bl; g1.N = S; g1.out = p.dskip; g1.out_bstride = (int64_t)bl * S; g1.ldo = S;
    g1.mask = p.skipsum; g1.mask_bstride = (int64_t)bl * S; g1.ldmask = S;
    if (int e = launch_gemm<EPI_PLAIN>(g1, st, tc)) return e;
    QP_CUDA(cudaMemsetAsync(p.dHup, 0, sizeof(float) * (size_t)B * L0 * A, st));
    if (mask) {
      QP_CUDA(cudaMemsetAsync(p.dbskip, 0, sizeof(float) * S, st));
      if (int e = tc::f32_to_bf16_colsum(p.dskip, (long long)B * bl, S, p.dskip_bf, p.dbskip, st)) return e;
    }
    if (tc_w) {   // the tcgen05 weight-gradient kernel accumulates its row splits into zeroed outputs
      QP_CUDA(cudaMemsetAsync(p.dW.Wg, 0, sizeof(float) * pd.wg_elems() * L, st));
      QP_CUDA(cudaMemsetAsync(p.dW.bg, 0, sizeof(float) * (size_t)2 * C * L, st));
      QP_CUDA(cudaMemsetAsync(p.dW.Wrs, 0, sizeof(float) * pd.wrs_elems() * L, st));
      QP_CUDA(cudaMemsetAsync(p.dW.brs, 0, sizeof(float) * (size_t)(C + S) * L, st));
    }
  }

  // ---- residual blocks, last to first -----------------------------------------------
  float* pong[2] = {p.dXa, p.dXb};
  for (int l = L - 1; l >= 0; --l) {
    const int stage = L - l;
    if (stage < s_begin || stage >= s_end) continue;
    const float* dXnext = l == L - 1 ? nullptr : pong[(l + 1) & 1];   // dX of the block above (this block's output gradient)
    const int Lin = p.Lin[l], sh = p.shift[l], n = Lin - sh;
    const int* rowmap = l >= pd.nF ? p.pastrow[l - pd.nF] : nullptr;
    float* Wrs = p.W.Wrs + pd.wrs_elems() * l;
    // bf16 copy of dXnext (+ its column sums = the residual-projection bias gradient)
    if (mask && dXnext)
      if (int e = tc::f32_to_bf16_colsum(dXnext, (long long)B * n, C, p.dX_bf, tc_w ? p.dW.brs + (size_t)(C + S) * l : nullptr, st)) return e;
    // dz = dXnext * R + dskip * K   ->  dgate (through the gate non-linearity)
    if (tc_dz) {
      tc::Args g = {};
      int s = 0;
      if (dXnext) g.seg[s++] = tc::Seg{p.dX_bf, (long long)n * C, C, nullptr, 0, n, C};
      g.seg[s++] = tc::Seg{p.dskip_bf, (long long)bl * S, S, nullptr, -(n - bl), bl, S};
      g.nseg = s;
      g.W = p.Wrs_bf + (size_t)l * (C + S) * C + (dXnext ? 0 : (size_t)C * C); g.ldw = C; g.w_mn = 1;
      g.Wb = p.WrsMN + (size_t)l * (C + S) * C; g.wb_pitch = C / 64; g.wb_k0 = dXnext ? 0 : C / 64;
      g.B = B; g.n_rows = n; g.N = C; g.BN = C < 256 ? C : 256;
      g.gin_bf = (const __nv_bfloat16*)p.G[l]; g.dgate_bf = p.dgate_bf; g.dgate_f32 = (tc_w && tc_dx) ? nullptr : p.dgate;
      if (int e = tc::gemm_dgate(g, st)) return e;
    } else {
      GemmArgs g = {};
      int s = 0;
      if (dXnext) g.seg[s++] = mk(dXnext, (int64_t)n * C, C, nullptr, 0, n, C);
      g.seg[s++] = mk(p.dskip, (int64_t)bl * S, S, nullptr, -(n - bl), bl, S);
      g.nseg = s;
      g.W = dXnext ? Wrs : Wrs + (size_t)C * C; g.ldw = C; g.w_kn = 1;
      g.B = B; g.n_rows = n; g.N = C;
      g.gsave = p.G[l]; g.gsave_bstride = (int64_t)n * 2 * C;
      g.out = p.dgate; g.out_bstride = (int64_t)n * 2 * C; g.ldo = 2 * C;
      if (int e = launch_gemm<EPI_DGATE>(g, st, tc)) return e;
      if (mask)
        if (int e = tc::f32_to_bf16_pad(p.dgate, (long long)B * n, 2 * C, 2 * C, p.dgate_bf, 0, st)) return e;
    }
    // d[res | skip] weights and biases
    if (tc_w) {
      tc::WgradArgs w = {};
      w.p[0] = tc::Seg{dXnext ? p.dX_bf : nullptr, (long long)n * C, C, nullptr, 0, n, C};   // all-zero for the last block
      w.p[1] = tc::Seg{p.dskip_bf, (long long)bl * S, S, nullptr, -(n - bl), bl, S}; w.np = 2;
      w.q[0] = tc::Seg{p.Zbf[l], (long long)n * C, C, nullptr, 0, n, C}; w.nq = 1;
      w.B = B; w.n_rows = n; w.I = C + S; w.J = C;
      w.out = p.dW.Wrs + pd.wrs_elems() * l; w.ldo = C;
      if (int e = tc::wgrad(w, st)) return e;
      QP_CUDA(cudaMemcpyAsync(p.dW.brs + (size_t)(C + S) * l + C, p.dbskip, sizeof(float) * S, cudaMemcpyDeviceToDevice, st));
    } else {
      WgradArgs w = {};
      w.p[0] = mk(dXnext, (int64_t)n * C, C, nullptr, 0, dXnext ? n : 0, C);   // dead (all-zero) for the last block
      w.p[1] = mk(p.dskip, (int64_t)bl * S, S, nullptr, -(n - bl), bl, S); w.np = 2;
      w.q[0] = mk(p.Z[l], (int64_t)n * C, C, nullptr, 0, n, C); w.nq = 1;
      w.B = B; w.n_rows = n; w.I = C + S; w.J = C;
      w.out = p.dW.Wrs + pd.wrs_elems() * l; w.ldo = C; w.colsum = p.dW.brs + (size_t)(C + S) * l;
      if (int e = launch_wgrad(w, st, tc)) return e;
    }
    // d gate weights / biases:  dgate^T * [x_past | x_cur | h_up | 1]
    if (tc_w) {
      tc::WgradArgs w = {};
      w.p[0] = tc::Seg{p.dgate_bf, (long long)n * 2 * C, 2 * C, nullptr, 0, n, 2 * C}; w.np = 1;
      w.q[0] = tc::Seg{p.Xbf[l], (long long)Lin * C, C, rowmap, 0, Lin, C};
      w.q[1] = tc::Seg{p.Xbf[l], (long long)Lin * C, C, nullptr, sh, Lin, C};
      w.q[2] = tc::Seg{p.Hup_bf, (long long)L0 * 64, 64, nullptr, L0 - n, L0, 64}; w.nq = 3;
      w.B = B; w.n_rows = n; w.I = 2 * C; w.J = pd.Kg;
      w.out = p.dW.Wg + pd.wg_elems() * l; w.ldo = pd.Kg;
      w.ones_out = p.dW.bg + (size_t)2 * C * l; w.ones_col = 2 * C + p.ones_col;
      if (int e = tc::wgrad(w, st)) return e;
    } else {
      WgradArgs w = {};
      w.p[0] = mk(p.dgate, (int64_t)n * 2 * C, 2 * C, nullptr, 0, n, 2 * C); w.np = 1;
      w.q[0] = mk(p.X[l], (int64_t)Lin * C, C, rowmap, 0, Lin, C);
      w.q[1] = mk(p.X[l], (int64_t)Lin * C, C, nullptr, sh, Lin, C);
      w.q[2] = mk(p.Hup, (int64_t)L0 * A, A, nullptr, L0 - n, L0, A); w.nq = 3;
      w.B = B; w.n_rows = n; w.I = 2 * C; w.J = pd.Kg;
      w.out = p.dW.Wg + pd.wg_elems() * l; w.ldo = pd.Kg; w.colsum = p.dW.bg + (size_t)2 * C * l;
      if (int e = launch_wgrad(w, st, tc)) return e;
    }
    // dX[l] (scatter: past rows, current rows + residual) and dHup
    {
      float* dX = pong[l & 1];
      QP_CUDA(cudaMemsetAsync(dX, 0, sizeof(float) * (size_t)B * Lin * C, st));
      if (tc_dx) {
        tc::Args g = {};
        g.seg[0] = tc::Seg{p.dgate_bf, (long long)n * 2 * C, 2 * C, nullptr, 0, n, 2 * C}; g.nseg = 1;
        g.W = p.Wg_bf + (size_t)l * 2 * C * Kgp; g.ldw = Kgp; g.w_mn = 1;
        g.Wb = p.WgMN + (size_t)l * 2 * C * Kgp; g.wb_pitch = Kgp / 64;
        g.B = B; g.n_rows = n; g.N = Kgp; g.BN = Kgp < 256 ? Kgp : 256; g.C = C; g.A = A;
        g.dx = dX; g.dx_bstride = (long long)Lin * C; g.dx_rowmap = rowmap; g.dx_past_off = 0; g.dx_cur_off = sh; g.dx_rows = Lin;
        g.resid = dXnext; g.resid_bstride = (long long)n * C;
        g.dh = p.dHup; g.dh_bstride = (long long)L0 * A; g.dh_off = L0 - n;
        if (int e = tc::gemm_dx(g, st)) return e;
        continue;
      }
      GemmArgs g = {};
      g.seg[0] = mk(p.dgate, (int64_t)n * 2 * C, 2 * C, nullptr, 0, n, 2 * C); g.nseg = 1;
      g.W = p.W.Wg + pd.wg_elems() * l; g.ldw = pd.Kg; g.w_kn = 1;
      g.B = B; g.n_rows = n; g.N = pd.Kg; g.C = C; g.A = A;
      g.dx = dX; g.dx_bstride = (int64_t)Lin * C; g.dx_rowmap = rowmap; g.dx_past_off = 0; g.dx_cur_off = sh; g.dx_rows = Lin;
      g.resid = dXnext; g.resid_bstride = (int64_t)n * C; g.ldresid = C;
      g.dh = p.dHup; g.dh_bstride = (int64_t)L0 * A; g.dh_off = L0 - n;
      if (int e = launch_gemm<EPI_DX>(g, st, tc)) return e;
    }
  }
  {   // the blocks of this range leave their gradients in the reference's tensors: blocks [L - s_end + 1, L - s_begin] clipped
    const int l_lo = std::max(0, L - (s_end - 1)), l_hi = std::min(L - 1, L - std::max(s_begin, 1));
    if (int e = unpack_grads_layers_f32(arch, p.gtab, p.dW, l_lo, l_hi + 1, st)) return e;
  }
  // ---- front end ----------------------------------------------------------------------
  if (L + 1 < s_begin || L + 1 >= s_end) return QP_OK;
  const float* dXnext = pong[0];                                      // dX of block 0 = gradient of the causal layer's output
  QP_CUDA(cudaMemsetAsync(p.dW.E0, 0, sizeof(float) * (size_t)Q * C, st));
  QP_CUDA(cudaMemsetAsync(p.dW.E1, 0, sizeof(float) * (size_t)Q * C, st));
  embed_grad_kernel<<<dim3(L0, B), 128, 0, st>>>(x, p.T, L0, C, Q, dXnext, p.dW.E0, p.dW.E1);
  QP_LAUNCH_CHECK();
  if (mask) {
    QP_CUDA(cudaMemsetAsync(grads[tm.causal_b()], 0, sizeof(float) * C, st));
    if (int e = tc::f32_to_bf16_colsum(dXnext, (long long)B * L0, C, nullptr, grads[tm.causal_b()], st)) return e;
  } else {
    colsum_kernel<<<(C + 31) / 32, 256, 0, st>>>(dXnext, (int64_t)B * L0, C, grads[tm.causal_b()]);
    QP_LAUNCH_CHECK();
  }
  QP_CUDA(cudaMemsetAsync(grads[tm.up_b()], 0, sizeof(float), st));
  upsample_grad_kernel<<<pd.U, 64, 0, st>>>(p.dHup, h, B, A, p.F, pd.U, L0, grads[tm.up_w()], grads[tm.up_b()]);
  QP_LAUNCH_CHECK();
  return unpack_grads_front_f32(arch, p.gtab, p.dW, st);
}

}  // namespace qp

using namespace qp;

extern "C" {

int qp_backward(const QpArch* arch, const float* const* tensors_host, const int64_t* x, const float* h, const float* d,
                int32_t B, int32_t T, int32_t F, int32_t bl, int32_t M, const float* dlogits, float* const* grads_host,
                void* ws, size_t ws_bytes, uint32_t flags, void* stream) {
  if (int e = check_device()) return e;
  if (int e = check_arch(arch)) return e;
  QP_REQUIRE(tensors_host && x && h && d && dlogits && grads_host && ws, "backward: NULL pointer");
  QP_REQUIRE(flags & QP_F_SAVE, "backward: the forward pass must have run with QP_F_SAVE");
  reset_launch_count();
  TfPlan p;
  size_t need = make_tf_plan(arch, B, T, F, bl, M, flags, ws, ws_bytes, &p);
  if (need > ws_bytes) return set_error(QP_EWORKSPACE, "backward: workspace %zu < %zu bytes", ws_bytes, need);
  return tf_backward_f32(arch, tensors_host, x, h, p, dlogits, grads_host, (flags & QP_F_BF16) != 0, (cudaStream_t)stream, 0,
                         arch->n_fixed + arch->n_adaptive + 2);
}

int qp_backward_range(const QpArch* arch, const float* const* tensors_host, const int64_t* x, const float* h, const float* d,
                      int32_t B, int32_t T, int32_t F, int32_t bl, int32_t M, const float* dlogits, float* const* grads_host,
                      void* ws, size_t ws_bytes, uint32_t flags, int32_t stage_begin, int32_t stage_end, void* stream) {
  if (int e = check_device()) return e;
  if (int e = check_arch(arch)) return e;
  QP_REQUIRE(tensors_host && x && h && d && dlogits && grads_host && ws, "backward: NULL pointer");
  QP_REQUIRE(flags & QP_F_SAVE, "backward: the forward pass must have run with QP_F_SAVE");
  const int n_stages = arch->n_fixed + arch->n_adaptive + 2;
  QP_REQUIRE(stage_begin >= 0 && stage_begin < stage_end && stage_end <= n_stages, "backward: stage range [%d, %d) outside [0, %d)",
             stage_begin, stage_end, n_stages);
  reset_launch_count();
  TfPlan p;
  size_t need = make_tf_plan(arch, B, T, F, bl, M, flags, ws, ws_bytes, &p);
  if (need > ws_bytes) return set_error(QP_EWORKSPACE, "backward: workspace %zu < %zu bytes", ws_bytes, need);
  return tf_backward_f32(arch, tensors_host, x, h, p, dlogits, grads_host, (flags & QP_F_BF16) != 0, (cudaStream_t)stream,
                         stage_begin, stage_end);
}

}  // extern "C"
