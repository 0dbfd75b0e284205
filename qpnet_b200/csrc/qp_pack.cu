// Parameter packing (reference state_dict tensors -> kernel layouts) and the inverse
// scatter for gradients.  Pure data movement; see qp_pack.cuh for the layouts.
#include "qp_pack.cuh"

namespace qp {

int upload_tensor_table(const QpArch* arch, const float* const* tensors_host, const float** dev_table,
                        cudaStream_t stream) {
  int n = tensor_map(arch).count();
  for (int i = 0; i < n; ++i)
    if (!tensors_host[i]) return set_error(QP_EINVAL, "parameter tensor %d is NULL", i);
  // pageable source: the runtime stages it before returning, so the caller's array may die
  QP_CUDA(cudaMemcpyAsync((void*)dev_table, tensors_host, sizeof(float*) * n, cudaMemcpyHostToDevice, stream));
  return QP_OK;
}

// One thread per element of Wg (all layers).  GRAD=false: params -> packed.
// GRAD=true: packed grads -> param grads (aux pad columns are skipped).
template <bool GRAD>
__global__ void pack_wg_kernel(TensorMap tm, PackedDims pd, const float* const* __restrict__ tab, float* __restrict__ Wg,
                               size_t i_begin, size_t i_end) {
  size_t i = i_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per = pd.wg_elems();
  if (i >= i_end) return;
  int l = (int)(i / per);
  size_t r = i % per;
  int row = (int)(r / pd.Kg), k = (int)(r % pd.Kg);
  int c = row >> 1, g = row & 1;
  int C = pd.C, A = pd.A;
  float* dst = nullptr;  // address of the parameter element this packed element mirrors
  if (l < pd.nF) {
    if (k < C) dst = (float*)tab[tm.dilF_w(g, l)] + ((size_t)c * C + k) * 2 + 0;
    else if (k < 2 * C) dst = (float*)tab[tm.dilF_w(g, l)] + ((size_t)c * C + (k - C)) * 2 + 1;
    else if (k < 2 * C + A) dst = (float*)tab[tm.auxF_w(g, l)] + (size_t)c * A + (k - 2 * C);
  } else {
    int j = l - pd.nF;
    if (k < C) dst = (float*)tab[tm.dilA_wP(g, j)] + (size_t)c * C + k;
    else if (k < 2 * C) dst = (float*)tab[tm.dilA_wC(g, j)] + (size_t)c * C + (k - C);
    else if (k < 2 * C + A) dst = (float*)tab[tm.auxA_w(g, j)] + (size_t)c * A + (k - 2 * C);
  }
  if (GRAD) {
    if (dst) *dst = Wg[i];
  } else {
    Wg[i] = dst ? *dst : 0.f;
  }
}

template <bool GRAD>
__global__ void pack_small_kernel(TensorMap tm, PackedDims pd, const float* const* __restrict__ tab, PackedF32 P) {
  // grid-stride over the union of: bg (L*2C), Wrs (L*(C+S)*C), brs (L*(C+S)), E0/E1 (2*Q*C)
  int C = pd.C, S = pd.S, Q = pd.Q, L = pd.L;
  size_t n_bg = (size_t)L * 2 * C, n_wrs = (size_t)L * (C + S) * C, n_brs = (size_t)L * (C + S), n_e = (size_t)Q * C;
  size_t total = n_bg + n_wrs + n_brs + 2 * n_e;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n_bg) {
      int l = (int)(i / (2 * C)), row = (int)(i % (2 * C)), c = row >> 1, g = row & 1;
      if (l < pd.nF) {
        float* b0 = (float*)tab[tm.dilF_b(g, l)] + c;
        float* b1 = (float*)tab[tm.auxF_b(g, l)] + c;
        if (GRAD) { *b0 = P.bg[i]; *b1 = P.bg[i]; } else P.bg[i] = *b0 + *b1;
      } else {
        int j = l - pd.nF;
        float* b0 = (float*)tab[tm.dilA_bC(g, j)] + c;
        float* b1 = (float*)tab[tm.dilA_bP(g, j)] + c;
        float* b2 = (float*)tab[tm.auxA_b(g, j)] + c;
        if (GRAD) { *b0 = P.bg[i]; *b1 = P.bg[i]; *b2 = P.bg[i]; } else P.bg[i] = *b0 + *b1 + *b2;
      }
      continue;
    }
    size_t k = i - n_bg;
    if (k < n_wrs) {
      int l = (int)(k / ((size_t)(C + S) * C));
      size_t r = k % ((size_t)(C + S) * C);
      int row = (int)(r / C), col = (int)(r % C);
      float* p;
      if (l < pd.nF) p = row < C ? (float*)tab[tm.resF_w(l)] + (size_t)row * C + col
                                 : (float*)tab[tm.skipF_w(l)] + (size_t)(row - C) * C + col;
      else p = row < C ? (float*)tab[tm.resA_w(l - pd.nF)] + (size_t)row * C + col
                       : (float*)tab[tm.skipA_w(l - pd.nF)] + (size_t)(row - C) * C + col;
      if (GRAD) *p = P.Wrs[k]; else P.Wrs[k] = *p;
      continue;
    }
    k -= n_wrs;
    if (k < n_brs) {
      int l = (int)(k / (C + S)), row = (int)(k % (C + S));
      float* p;
      if (l < pd.nF) p = row < C ? (float*)tab[tm.resF_b(l)] + row : (float*)tab[tm.skipF_b(l)] + (row - C);
      else p = row < C ? (float*)tab[tm.resA_b(l - pd.nF)] + row : (float*)tab[tm.skipA_b(l - pd.nF)] + (row - C);
      if (GRAD) *p = P.brs[k]; else P.brs[k] = *p;
      continue;
    }
    k -= n_brs;
    {
      int tap = k >= n_e;
      size_t r = tap ? k - n_e : k;
      int q = (int)(r / C), c = (int)(r % C);
      float* p = (float*)tab[tm.causal_w()] + ((size_t)c * Q + q) * 2 + tap;  // weight (C, Q, 2)
      float* e = tap ? P.E1 : P.E0;
      if (GRAD) *p = e[r]; else e[r] = *p;
    }
  }
}

// packed gradients of the residual blocks [l_begin, l_end) -> the reference's tensors (gate biases, [res | skip] weights
// and biases)
__global__ void unpack_small_layer_kernel(TensorMap tm, PackedDims pd, const float* const* __restrict__ tab, PackedF32 P, int l_begin,
                                          int l_end) {
  const int C = pd.C, S = pd.S;
  const size_t n_bg = (size_t)2 * C, n_wrs = (size_t)(C + S) * C, n_brs = (size_t)(C + S);
  const size_t per = n_bg + n_wrs + n_brs;
  for (size_t ii = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ii < per * (l_end - l_begin); ii += (size_t)gridDim.x * blockDim.x) {
    const int l = l_begin + (int)(ii / per);
    const size_t i = ii % per;
    const bool fixed = l < pd.nF;
    const int j = fixed ? l : l - pd.nF;
    if (i < n_bg) {
      const int row = (int)i, c = row >> 1, g = row & 1;
      const float v = P.bg[(size_t)l * 2 * C + i];
      if (fixed) { ((float*)tab[tm.dilF_b(g, j)])[c] = v; ((float*)tab[tm.auxF_b(g, j)])[c] = v; }
      else { ((float*)tab[tm.dilA_bC(g, j)])[c] = v; ((float*)tab[tm.dilA_bP(g, j)])[c] = v; ((float*)tab[tm.auxA_b(g, j)])[c] = v; }
      continue;
    }
    size_t k = i - n_bg;
    if (k < n_wrs) {
      const int row = (int)(k / C), col = (int)(k % C);
      float* q;
      if (fixed) q = row < C ? (float*)tab[tm.resF_w(j)] + (size_t)row * C + col : (float*)tab[tm.skipF_w(j)] + (size_t)(row - C) * C + col;
      else q = row < C ? (float*)tab[tm.resA_w(j)] + (size_t)row * C + col : (float*)tab[tm.skipA_w(j)] + (size_t)(row - C) * C + col;
      *q = P.Wrs[(size_t)l * n_wrs + k];
      continue;
    }
    k -= n_wrs;
    const int row = (int)k;
    float* q;
    if (fixed) q = row < C ? (float*)tab[tm.resF_b(j)] + row : (float*)tab[tm.skipF_b(j)] + (row - C);
    else q = row < C ? (float*)tab[tm.resA_b(j)] + row : (float*)tab[tm.skipA_b(j)] + (row - C);
    *q = P.brs[(size_t)l * n_brs + k];
  }
}

// packed causal-layer table gradients (E0, E1) -> causal.conv.weight (C, Q, 2)
__global__ void unpack_front_kernel(TensorMap tm, PackedDims pd, const float* const* __restrict__ tab, PackedF32 P) {
  const int C = pd.C, Q = pd.Q;
  const size_t n_e = (size_t)Q * C;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < 2 * n_e; k += (size_t)gridDim.x * blockDim.x) {
    const int tap = k >= n_e;
    const size_t r = tap ? k - n_e : k;
    const int q = (int)(r / C), c = (int)(r % C);
    ((float*)tab[tm.causal_w()])[((size_t)c * Q + q) * 2 + tap] = (tap ? P.E1 : P.E0)[r];
  }
}

int unpack_grads_layers_f32(const QpArch* arch, float* const* dev_grad_table, PackedF32 g, int l_begin, int l_end, cudaStream_t stream) {
  if (l_end <= l_begin) return QP_OK;
  TensorMap tm = tensor_map(arch);
  PackedDims pd = packed_dims(arch);
  const size_t per = pd.wg_elems(), n = per * (l_end - l_begin);
  pack_wg_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tm, pd, (const float* const*)dev_grad_table, g.Wg, per * l_begin, per * l_end);
  QP_LAUNCH_CHECK();
  unpack_small_layer_kernel<<<148 * 2, 256, 0, stream>>>(tm, pd, (const float* const*)dev_grad_table, g, l_begin, l_end);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

int unpack_grads_front_f32(const QpArch* arch, float* const* dev_grad_table, PackedF32 g, cudaStream_t stream) {
  TensorMap tm = tensor_map(arch);
  PackedDims pd = packed_dims(arch);
  unpack_front_kernel<<<148, 256, 0, stream>>>(tm, pd, (const float* const*)dev_grad_table, g);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

int pack_f32(const QpArch* arch, const float* const* dev_table, PackedF32 out, cudaStream_t stream) {
  TensorMap tm = tensor_map(arch);
  PackedDims pd = packed_dims(arch);
  size_t n = pd.wg_elems() * pd.L;
  pack_wg_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tm, pd, dev_table, out.Wg, 0, n);
  QP_LAUNCH_CHECK();
  pack_small_kernel<false><<<148 * 4, 256, 0, stream>>>(tm, pd, dev_table, out);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

int unpack_grads_f32(const QpArch* arch, float* const* dev_grad_table, PackedF32 g, cudaStream_t stream) {
  TensorMap tm = tensor_map(arch);
  PackedDims pd = packed_dims(arch);
  size_t n = pd.wg_elems() * pd.L;
  pack_wg_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tm, pd, (const float* const*)dev_grad_table, g.Wg, 0, n);
  QP_LAUNCH_CHECK();
  pack_small_kernel<true><<<148 * 4, 256, 0, stream>>>(tm, pd, (const float* const*)dev_grad_table, g);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

}  // namespace qp
