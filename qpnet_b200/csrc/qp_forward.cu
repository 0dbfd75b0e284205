// Teacher-forced stack, fp32 exact path: QPNet.forward (qpnet.py:239-312).
//   embed (OneHot + causal conv as a 2-row table lookup)  -> X[0]
//   upsample (shared U-tap transposed conv)               -> Hup
//   per block: [past-row index kernel] -> gate GEMM (+sigmoid*tanh) -> res/skip GEMM
//   head: relu -> 1x1 -> relu -> 1x1
#include "qp_gemm_f32.cuh"
#include "qp_tc.cuh"
#include "qp_tf_plan.cuh"

namespace qp {

// X0[b][i][c] = E0[x[b][T-L0-1+i]][c] + E1[x[b][T-L0+i]][c] + bias[c]   (qpnet.py:76-79,131,262)
__global__ void embed_kernel(const int64_t* __restrict__ x, int T, int L0, int C, int Q, const float* __restrict__ E0,
                             const float* __restrict__ E1, const float* __restrict__ bias, float* __restrict__ X0) {
  int b = blockIdx.y, i = blockIdx.x;
  const int64_t* xb = x + (int64_t)b * T + (T - L0 - 1);
  int s0 = (int)(((xb[i] % Q) + Q) % Q), s1 = (int)(((xb[i + 1] % Q) + Q) % Q);
  float* o = X0 + ((int64_t)b * L0 + i) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) o[c] = E0[(int64_t)s0 * C + c] + E1[(int64_t)s1 * C + c] + bias[c];
}

// Hup[b][i][a] = h[b][a][f]*w[j] + bias,  p = F*U - L0 + i, f = p / U, j = p % U   (qpnet.py:143-158,264)
__global__ void upsample_kernel(const float* __restrict__ h, int A, int F, int U, int L0, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ Hup) {
  int b = blockIdx.y;
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)L0 * A) return;
  int i = (int)(e / A), a = (int)(e % A);
  int p = F * U - L0 + i;
  int f = p / U, j = p % U;
  Hup[((int64_t)b * L0 + i) * A + a] = h[((int64_t)b * A + a) * F + f] * w[j] + bias[0];
}

// pastrow[b][r] = Lin + idx,  idx from qpnet.py:594-600 on d[b][T-n+r]; flags out-of-range (qpnet.py:294)
__global__ void pastrow_kernel(const float* __restrict__ d, int T, int n, int Lin, int dil, int* __restrict__ pastrow,
                               int32_t* __restrict__ status) {
  int b = blockIdx.y;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int idx = tf_index_f32(d[(int64_t)b * T + (T - n + r)], dil, r - n);
  int src = Lin + idx;
  if (src < 0 || src >= Lin) {
    atomicExch(status, QP_ERANGE);
    src = src < 0 ? 0 : Lin - 1;
  }
  pastrow[(int64_t)b * n + r] = src;
}

static Seg make_seg(const float* base, int64_t bstride, int ld, const int* rowmap, int row_off, int src_rows, int K,
                    int relu = 0) {
  Seg s; s.base = base; s.bstride = bstride; s.ld = ld; s.rowmap = rowmap; s.row_off = row_off;
  s.src_rows = src_rows; s.K = K; s.relu = relu;
  return s;
}

int tf_forward_f32(const QpArch* arch, const float* const* tensors, const int64_t* x, const float* h, const float* d,
                   const TfPlan& p, float* logits, uint32_t flags, cudaStream_t st) {
  const PackedDims& pd = p.pd;
  const TensorMap tm = tensor_map(arch);
  const int C = pd.C, S = pd.S, Q = pd.Q, A = pd.A, B = p.B, L0 = p.L0, bl = p.bl;
  QP_CUDA(cudaMemsetAsync(p.status, 0, sizeof(int32_t), st));
  if (int e = upload_tensor_table(arch, tensors, p.tab, st)) return e;
  if (int e = pack_f32(arch, p.tab, p.W, st)) return e;
  embed_kernel<<<dim3(L0, B), 128, 0, st>>>(x, p.T, L0, C, Q, p.W.E0, p.W.E1, tensors[tm.causal_b()], p.X[0]);
  QP_LAUNCH_CHECK();
  upsample_kernel<<<dim3((unsigned)(((int64_t)L0 * A + 255) / 256), B), 256, 0, st>>>(
      h, A, p.F, pd.U, L0, tensors[tm.up_w()], tensors[tm.up_b()], p.Hup);
  QP_LAUNCH_CHECK();
  for (int l = 0; l < pd.L; ++l) {
    const int Lin = p.Lin[l], sh = p.shift[l], n = Lin - sh;
    const int* rowmap = nullptr;
    if (l >= pd.nF) {
      int* pr = p.pastrow[l - pd.nF];
      pastrow_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(d, p.T, n, Lin, p.dil[l], pr, p.status);
      QP_LAUNCH_CHECK();
      rowmap = pr;
    }
    GemmArgs g = {};
    g.seg[0] = make_seg(p.X[l], (int64_t)Lin * C, C, rowmap, 0, Lin, C);
    g.seg[1] = make_seg(p.X[l], (int64_t)Lin * C, C, nullptr, sh, Lin, C);
    g.seg[2] = make_seg(p.Hup, (int64_t)L0 * A, A, nullptr, L0 - n, L0, A);
    g.nseg = 3;
    g.W = p.W.Wg + pd.wg_elems() * l; g.ldw = pd.Kg; g.w_kn = 0;
    g.bias = p.W.bg + (size_t)2 * C * l;
    g.B = B; g.n_rows = n; g.N = 2 * C; g.n_begin = 0;
    g.out = p.Z[l]; g.out_bstride = (int64_t)n * C; g.ldo = C;
    g.gsave = p.G[l]; g.gsave_bstride = (int64_t)n * 2 * C;
    if (int e = launch_gemm<EPI_GATE>(g, st)) return e;

    GemmArgs r = {};
    r.seg[0] = make_seg(p.Z[l], (int64_t)n * C, C, nullptr, 0, n, C);
    r.nseg = 1;
    r.W = p.W.Wrs + pd.wrs_elems() * l; r.ldw = C; r.w_kn = 0;
    r.bias = p.W.brs + (size_t)(C + S) * l;
    r.B = B; r.n_rows = n; r.N = C + S;
    r.n_begin = (l == pd.L - 1) ? C : 0;  // the last block's residual projection is dead (caveat C7)
    r.out = (l + 1 < pd.L) ? p.X[l + 1] : nullptr; r.out_bstride = (int64_t)n * C; r.ldo = C;
    r.resid = p.X[l]; r.resid_bstride = (int64_t)Lin * C; r.ldresid = C; r.resid_off = sh;
    r.skip = p.skipsum; r.skip_bstride = (int64_t)bl * S; r.skip_row0 = n - bl; r.skip_accum = l > 0; r.C = C;
    if (int e = launch_gemm<EPI_RESSKIP>(r, st)) return e;
  }
  // head (qpnet.py:566-571)
  GemmArgs h1 = {};
  h1.seg[0] = make_seg(p.skipsum, (int64_t)bl * S, S, nullptr, 0, bl, S, 1);
  h1.nseg = 1; h1.W = tensors[tm.post1_w()]; h1.ldw = S; h1.bias = tensors[tm.post1_b()];
  h1.B = B; h1.n_rows = bl; h1.N = S; h1.out = p.H1; h1.out_bstride = (int64_t)bl * S; h1.ldo = S;
  if (int e = launch_gemm<EPI_PLAIN>(h1, st)) return e;
  GemmArgs h2 = {};
  h2.seg[0] = make_seg(p.H1, (int64_t)bl * S, S, nullptr, 0, bl, S, 1);
  h2.nseg = 1; h2.W = tensors[tm.post2_w()]; h2.ldw = S; h2.bias = tensors[tm.post2_b()];
  h2.B = B; h2.n_rows = bl; h2.N = Q; h2.out = logits; h2.out_bstride = (int64_t)bl * Q; h2.ldo = Q;
  if (int e = launch_gemm<EPI_PLAIN>(h2, st)) return e;
  (void)flags;
  return QP_OK;
}

// ------------------------------------------------------------------ bf16 tensor-core path
// Same dataflow as tf_forward_f32 with every contraction on tcgen05 (qp_tc.cu).  The residual
// stream X stays fp32 (read / written by the res epilogue); GEMM operands are bf16 copies.
// With QP_F_SAVE the epilogues also emit the fp32 Z and sigmoid/tanh values backward consumes.
int tf_forward_bf16(const QpArch* arch, const float* const* tensors, const int64_t* x, const float* h, const float* d,
                    const TfPlan& p, float* logits, uint32_t flags, cudaStream_t st) {
  const PackedDims& pd = p.pd;
  const TensorMap tm = tensor_map(arch);
  const int C = pd.C, S = pd.S, Q = pd.Q, A = pd.A, B = p.B, L0 = p.L0, bl = p.bl, Kgp = p.Kgp;
  const bool save = flags & QP_F_SAVE;
  // the fp32 copy of z only feeds the TF32 weight-gradient kernel; the tcgen05 one reads the bf16 operand copy
  const bool save_zf = save && !((bwd_tc_mask() & 1) && p.ones_col >= 0);
  // likewise the sigmoid / tanh values: the tcgen05 dz GEMM reads them back as bf16 (stored in the same G[l] buffer)
  const bool g_bf = (bwd_tc_mask() & 2) && p.ones_col >= 0;
  QP_CUDA(cudaMemsetAsync(p.status, 0, sizeof(int32_t), st));
  if (int e = upload_tensor_table(arch, tensors, p.tab, st)) return e;
  if (int e = pack_f32(arch, p.tab, p.W, st)) return e;
  if (int e = tc::pack_wg_bf16(p.W.Wg, (long long)pd.L * 2 * C, 2 * C, pd.Kg, Kgp, p.Wg_bf, st)) return e;
  if (int e = tc::f32_to_bf16_pad(p.W.Wrs, (long long)pd.L * (C + S), C, C, p.Wrs_bf, 0, st)) return e;
  if (int e = tc::block_pack(p.Wg_bf, pd.L, 2 * C, Kgp, 1, p.WgK, st)) return e;
  if (int e = tc::block_pack(p.Wrs_bf, pd.L, C + S, C, 1, p.WrsK, st)) return e;
  if (save) {
    if (int e = tc::block_pack(p.Wg_bf, pd.L, 2 * C, Kgp, 0, p.WgMN, st)) return e;
    if (int e = tc::block_pack(p.Wrs_bf, pd.L, C + S, C, 0, p.WrsMN, st)) return e;
  }
  if (int e = tc::f32_to_bf16_pad(tensors[tm.post1_w()], S, S, S, p.W1_bf, 0, st)) return e;
  if (int e = tc::f32_to_bf16_pad(tensors[tm.post2_w()], Q, S, S, p.W2_bf, 0, st)) return e;
  embed_kernel<<<dim3(L0, B), 128, 0, st>>>(x, p.T, L0, C, Q, p.W.E0, p.W.E1, tensors[tm.causal_b()], p.X[0]);
  QP_LAUNCH_CHECK();
  if (int e = tc::f32_to_bf16_pad(p.X[0], (long long)B * L0, C, C, p.Xbf[0], 0, st)) return e;
  upsample_kernel<<<dim3((unsigned)(((int64_t)L0 * A + 255) / 256), B), 256, 0, st>>>(
      h, A, p.F, pd.U, L0, tensors[tm.up_w()], tensors[tm.up_b()], p.Hup);
  QP_LAUNCH_CHECK();
  if (int e = tc::f32_to_bf16_pad(p.Hup, (long long)B * L0, A, 64, p.Hup_bf, 0, st, p.ones_col)) return e;
  for (int l = 0; l < pd.L; ++l) {
    const int Lin = p.Lin[l], sh = p.shift[l], n = Lin - sh;
    const int* rowmap = nullptr;
    if (l >= pd.nF) {
      int* pr = p.pastrow[l - pd.nF];
      pastrow_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(d, p.T, n, Lin, p.dil[l], pr, p.status);
      QP_LAUNCH_CHECK();
      rowmap = pr;
    }
    const __nv_bfloat16* Xin = p.Xbf[l];
    tc::Args g = {};
    g.seg[0] = tc::Seg{Xin, (long long)Lin * C, C, rowmap, 0, Lin, C};
    g.seg[1] = tc::Seg{Xin, (long long)Lin * C, C, nullptr, sh, Lin, C};
    g.seg[2] = tc::Seg{p.Hup_bf, (long long)L0 * 64, 64, nullptr, L0 - n, L0, 64};
    g.nseg = 3;
    g.W = p.Wg_bf + (size_t)l * 2 * C * Kgp; g.ldw = Kgp;
    g.Wb = p.WgK + (size_t)l * 2 * C * Kgp; g.wb_pitch = 2 * C / 64;
    g.bias = p.W.bg + (size_t)2 * C * l;
    g.B = B; g.n_rows = n; g.N = 2 * C; g.n_begin = 0; g.BN = 2 * C < 256 ? 2 * C : 256;
    g.z_bf = p.Zbf[l]; g.z_f32 = save_zf ? p.Z[l] : nullptr; g.gsave = (save && !g_bf) ? p.G[l] : nullptr;
    g.gsave_bf = (save && g_bf) ? (__nv_bfloat16*)p.G[l] : nullptr;
    if (int e = tc::gemm_gate(g, st)) return e;

    tc::Args r = {};
    r.seg[0] = tc::Seg{p.Zbf[l], (long long)n * C, C, nullptr, 0, n, C};
    r.nseg = 1;
    r.W = p.Wrs_bf + (size_t)l * (C + S) * C; r.ldw = C;
    r.Wb = p.WrsK + (size_t)l * (C + S) * C; r.wb_pitch = (C + S) / 64;
    r.bias = p.W.brs + (size_t)(C + S) * l;
    r.B = B; r.n_rows = n; r.N = C + S;
    r.n_begin = (l == pd.L - 1) ? C : 0;   // the last block's residual projection is dead (caveat C7)
    r.BN = (r.N - r.n_begin) < 256 ? (r.N - r.n_begin) : 256;
    r.C = C; r.S = S;
    r.xcur = p.X[l]; r.xcur_bstride = (long long)Lin * C; r.xcur_off = sh;
    r.xnext = (l + 1 < pd.L) ? p.X[l + 1] : nullptr; r.xnext_bf = p.Xbf[l + 1];
    r.skip = p.skipsum; r.skip_bstride = (long long)bl * S; r.skip_row0 = n - bl; r.skip_accum = l > 0;
    if (int e = tc::gemm_resskip(r, st)) return e;
  }
  // head (qpnet.py:566-571)
  if (int e = tc::f32_to_bf16_pad(p.skipsum, (long long)B * bl, S, S, p.skip_bf, 1, st)) return e;
  tc::Args h1 = {};
  h1.seg[0] = tc::Seg{p.skip_bf, (long long)bl * S, S, nullptr, 0, bl, S};
  h1.nseg = 1; h1.W = p.W1_bf; h1.ldw = S; h1.bias = tensors[tm.post1_b()];
  h1.B = B; h1.n_rows = bl; h1.N = S; h1.BN = S < 256 ? S : 256;
  h1.out = p.H1; h1.out_bstride = (long long)bl * S; h1.ldo = S; h1.out_relu_bf = p.H1_bf;
  if (int e = tc::gemm_head(h1, st)) return e;
  tc::Args h2 = {};
  h2.seg[0] = tc::Seg{p.H1_bf, (long long)bl * S, S, nullptr, 0, bl, S};
  h2.nseg = 1; h2.W = p.W2_bf; h2.ldw = S; h2.bias = tensors[tm.post2_b()];
  h2.B = B; h2.n_rows = bl; h2.N = Q; h2.BN = Q < 256 ? Q : 256;
  h2.out = logits; h2.out_bstride = (long long)bl * Q; h2.ldo = Q;
  if (int e = tc::gemm_head(h2, st)) return e;
  return QP_OK;
}

// ------------------------------------------------------------------ fused softmax-CE
// one warp per row: loss_sum += -log p[target];  dlogits = (p - onehot) * scale
__global__ void cross_entropy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t rows,
                                     int Q, float scale, float* __restrict__ loss_sum, float* __restrict__ dlogits) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* lr = logits + row * Q;
  float mx = -INFINITY;
  for (int q = lane; q < Q; q += 32) mx = fmaxf(mx, lr[q]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int q = lane; q < Q; q += 32) sum += expf(lr[q] - mx);
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  // a target outside [0, Q) (the reference asserts max(target) < n_quantize, qpnet_train.py:524) is never used as an
  // index: the loss becomes NaN, which the caller cannot miss
  const int64_t t64 = target[row];
  const bool t_ok = t64 >= 0 && t64 < Q;
  const int t = t_ok ? (int)t64 : -1;
  float lse = mx + logf(sum);
  if (lane == 0) atomicAdd(loss_sum, t_ok ? lse - lr[t] : __int_as_float(0x7fc00000));
  if (dlogits) {
    float inv = 1.f / sum;
    for (int q = lane; q < Q; q += 32) {
      float pq = expf(lr[q] - mx) * inv;
      dlogits[row * Q + q] = (pq - (q == t ? 1.f : 0.f)) * scale;
    }
  }
}

}  // namespace qp

using namespace qp;

static int validate_tf(const QpArch* arch, int B, int T, int F, int bl, int M) {
  if (int e = check_arch(arch)) return e;
  QP_REQUIRE(B >= 1 && bl >= 1 && M >= 1 && T >= 1 && F >= 1, "forward: bad shape B=%d T=%d F=%d bl=%d M=%d", B, T, F, bl, M);
  int rfF = 0, rfA = 0;
  for (int i = 0; i < arch->n_fixed; ++i) rfF += arch->dil_fixed[i];
  for (int i = 0; i < arch->n_adaptive; ++i) rfA += arch->dil_adaptive[i];
  int64_t L0 = (int64_t)rfA * M + rfF + bl;
  QP_REQUIRE(T >= L0 + 1, "forward: T=%d shorter than receptive field + batch_length = %lld", T, (long long)(L0 + 1));
  QP_REQUIRE((int64_t)F * arch->upsampling >= L0, "forward: aux covers %lld samples, need %lld",
             (long long)F * arch->upsampling, (long long)L0);
  return QP_OK;
}

extern "C" {

size_t qp_forward_workspace_bytes(const QpArch* arch, int32_t B, int32_t T, int32_t bl, int32_t M, uint32_t flags) {
  if (check_arch(arch) != QP_OK || B < 1 || bl < 1 || M < 1) return 0;
  TfPlan p;
  return make_tf_plan(arch, B, T, 0, bl, M, flags, nullptr, 0, &p);
}

int qp_forward(const QpArch* arch, const float* const* tensors_host, const int64_t* x, const float* h, const float* d,
               int32_t B, int32_t T, int32_t F, int32_t bl, int32_t M, float* logits, void* ws, size_t ws_bytes,
               uint32_t flags, void* stream) {
  if (int e = check_device()) return e;
  if (int e = validate_tf(arch, B, T, F, bl, M)) return e;
  QP_REQUIRE(tensors_host && x && h && d && logits && ws, "forward: NULL pointer");
  if (flags & QP_F_BF16)
    QP_REQUIRE(arch->n_resch % 64 == 0 && arch->n_skipch % 64 == 0 && arch->n_quantize % 32 == 0 && arch->n_aux <= 64,
               "forward: the bf16 tensor-core path needs n_resch %% 64 == 0, n_skipch %% 64 == 0, n_quantize %% 32 == 0");
  reset_launch_count();
  TfPlan p;
  size_t need = make_tf_plan(arch, B, T, F, bl, M, flags, ws, ws_bytes, &p);
  if (need > ws_bytes) return set_error(QP_EWORKSPACE, "forward: workspace %zu < %zu bytes", ws_bytes, need);
  if (flags & QP_F_BF16) return tf_forward_bf16(arch, tensors_host, x, h, d, p, logits, flags, (cudaStream_t)stream);
  return tf_forward_f32(arch, tensors_host, x, h, d, p, logits, flags, (cudaStream_t)stream);
}

int qp_cross_entropy(const float* logits, const int64_t* target, int64_t rows, int32_t Q, float scale, float* loss_sum,
                     float* dlogits, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(logits && target && loss_sum && rows >= 0 && Q > 0, "cross_entropy: bad arguments");
  reset_launch_count();
  if (rows == 0) return QP_OK;
  cross_entropy_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(logits, target, rows, Q, scale,
                                                                                    loss_sum, dlogits);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

}  // extern "C"
