// Packed weight layouts shared by the teacher-forced and the generator paths.
//
// The reference stores every projection as its own Conv1d tensor (qpnet.py:201-235).
// The kernels want, per residual block l (fixed blocks first, then adaptive):
//   Wg[l]  : (2C, Kg)  gate matrix, Kg = 2C + Ap (Ap = n_aux rounded up to 16)
//            row 2c   = sigmoid channel c,  row 2c+1 = tanh channel c
//            cols [0,C)   past tap  (dilF conv weight[..., 0]  /  dilA convP)
//            cols [C,2C)  current   (dilF conv weight[..., 1]  /  dilA convC)
//            cols [2C,2C+A) aux 1x1, remaining columns zero
//   bg[l]  : (2C)  all biases that feed the pre-activation summed
//            (fixed: conv + aux; adaptive: convC + convP + aux -- qpnet.py:95-96,630-635)
//   Wrs[l] : (C+S, C)  rows [0,C) res 1x1, rows [C,C+S) skip 1x1;  brs[l] : (C+S)
//   E0,E1  : (Q, C)  causal conv taps transposed into embedding tables (qpnet.py:76-79,131)
#pragma once
#include "qp_common.cuh"

namespace qp {

struct PackedDims {
  int C, S, Q, A, Ap, Kg, L, nF, nA, U;
  __host__ __device__ size_t wg_elems() const { return (size_t)2 * C * Kg; }
  __host__ __device__ size_t wrs_elems() const { return (size_t)(C + S) * C; }
};

inline PackedDims packed_dims(const QpArch* a) {
  PackedDims p;
  p.C = a->n_resch; p.S = a->n_skipch; p.Q = a->n_quantize; p.A = a->n_aux;
  p.Ap = (a->n_aux + 15) / 16 * 16;
  p.Kg = 2 * p.C + p.Ap;
  p.nF = a->n_fixed; p.nA = a->n_adaptive; p.L = p.nF + p.nA; p.U = a->upsampling;
  return p;
}

// fp32 packed model living in the caller's workspace
struct PackedF32 {
  float* Wg;   // L * 2C * Kg
  float* bg;   // L * 2C
  float* Wrs;  // L * (C+S) * C
  float* brs;  // L * (C+S)
  float* E0;   // Q * C
  float* E1;   // Q * C
};

// device-side table of raw parameter pointers (copied to the device once per call)
struct DevTensorTable {
  const float* p[4 + 12 * QP_MAX_LAYERS + 16 * QP_MAX_LAYERS + 4];
};

int upload_tensor_table(const QpArch* arch, const float* const* tensors_host, const float** dev_table,
                        cudaStream_t stream);
int pack_f32(const QpArch* arch, const float* const* dev_table, PackedF32 out, cudaStream_t stream);
// scatter packed gradients back into the reference's tensors (overwrites)
int unpack_grads_f32(const QpArch* arch, float* const* dev_grad_table, PackedF32 grads, cudaStream_t stream);
// the same for the residual blocks [l_begin, l_end) / for the causal layer: the backward hands finished gradients out block by block, so a
// data-parallel caller can start reducing them while the rest of the backward runs
int unpack_grads_layers_f32(const QpArch* arch, float* const* dev_grad_table, PackedF32 grads, int l_begin, int l_end, cudaStream_t stream);
int unpack_grads_front_f32(const QpArch* arch, float* const* dev_grad_table, PackedF32 grads, cudaStream_t stream);

}  // namespace qp
