// Shared host/device helpers for libqpnet_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/qpnet_b200.h"

namespace qp {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
void reset_launch_count();

#define QP_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return qp::set_error(QP_ECUDA, "%s failed: %s (%s:%d)", #call,                    \
                           cudaGetErrorString(e_), __FILE__, __LINE__);                 \
  } while (0)

#define QP_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    qp::count_launch();                                                                 \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess)                                                              \
      return qp::set_error(QP_ECUDA, "kernel launch failed: %s (%s:%d)",                \
                           cudaGetErrorString(e_), __FILE__, __LINE__);                 \
  } while (0)

#define QP_REQUIRE(cond, ...)                                                           \
  do {                                                                                  \
    if (!(cond)) return qp::set_error(QP_EINVAL, __VA_ARGS__);                          \
  } while (0)

int check_device();  // QP_OK or QP_EARCH
int check_arch(const QpArch* a);

// ---- positions of the reference's state_dict tensors in the pointer table ----------
struct TensorMap {
  int nF, nA;
  __host__ __device__ int causal_w() const { return 0; }
  __host__ __device__ int causal_b() const { return 1; }
  __host__ __device__ int up_w() const { return 2; }
  __host__ __device__ int up_b() const { return 3; }
  // fixed blocks: g = 0 sigmoid, 1 tanh
  __host__ __device__ int dilF_w(int g, int i) const { return 4 + g * 2 * nF + 2 * i; }
  __host__ __device__ int dilF_b(int g, int i) const { return dilF_w(g, i) + 1; }
  __host__ __device__ int auxF_w(int g, int i) const { return 4 + 4 * nF + g * 2 * nF + 2 * i; }
  __host__ __device__ int auxF_b(int g, int i) const { return auxF_w(g, i) + 1; }
  __host__ __device__ int skipF_w(int i) const { return 4 + 8 * nF + 2 * i; }
  __host__ __device__ int skipF_b(int i) const { return skipF_w(i) + 1; }
  __host__ __device__ int resF_w(int i) const { return 4 + 10 * nF + 2 * i; }
  __host__ __device__ int resF_b(int i) const { return resF_w(i) + 1; }
  __host__ __device__ int baseA() const { return 4 + 12 * nF; }
  // adaptive blocks: convC w,b then convP w,b
  __host__ __device__ int dilA_wC(int g, int i) const { return baseA() + g * 4 * nA + 4 * i; }
  __host__ __device__ int dilA_bC(int g, int i) const { return dilA_wC(g, i) + 1; }
  __host__ __device__ int dilA_wP(int g, int i) const { return dilA_wC(g, i) + 2; }
  __host__ __device__ int dilA_bP(int g, int i) const { return dilA_wC(g, i) + 3; }
  __host__ __device__ int auxA_w(int g, int i) const { return baseA() + 8 * nA + g * 2 * nA + 2 * i; }
  __host__ __device__ int auxA_b(int g, int i) const { return auxA_w(g, i) + 1; }
  __host__ __device__ int skipA_w(int i) const { return baseA() + 12 * nA + 2 * i; }
  __host__ __device__ int skipA_b(int i) const { return skipA_w(i) + 1; }
  __host__ __device__ int resA_w(int i) const { return baseA() + 14 * nA + 2 * i; }
  __host__ __device__ int resA_b(int i) const { return resA_w(i) + 1; }
  __host__ __device__ int post1_w() const { return baseA() + 16 * nA; }
  __host__ __device__ int post1_b() const { return post1_w() + 1; }
  __host__ __device__ int post2_w() const { return post1_w() + 2; }
  __host__ __device__ int post2_b() const { return post1_w() + 3; }
  __host__ __device__ int count() const { return post1_w() + 4; }
};

inline TensorMap tensor_map(const QpArch* a) { return TensorMap{a->n_fixed, a->n_adaptive}; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace (never allocates device memory itself).
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t c) : base((char*)p), cap(c), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return base == nullptr || off <= cap; }
};

// ---- bit-exact index arithmetic shared by the standalone builders and the fused paths --
// qpnet.py:594-600: fp32 product, fp32 sum with the (exactly representable) position,
// round-half-even.  No FMA contraction: explicit _rn intrinsics.
__device__ __forceinline__ int tf_index_f32(float d, int dil, int t_minus_n) {
  float prod = __fmul_rn(-d, (float)dil);
  float s = __fadd_rn(prod, (float)t_minus_n);
  return (int)rintf(s);
}
// qpnet.py:606-609
__device__ __forceinline__ int tf_index_f64(double d, int dil, int t_minus_n) {
  double s = __dadd_rn(__dmul_rn(-d, (double)dil), (double)t_minus_n);
  return (int)rint(s);
}
// qpnet.py:615-617
__device__ __forceinline__ int gen_index_f32(float d, int dil) {
  return (int)rintf(__fmul_rn(-d, (float)dil));
}
// qpnet.py:621-622
__device__ __forceinline__ int gen_index_f64(double d, int dil) {
  return (int)rint(__dmul_rn(-d, (double)dil));
}

}  // namespace qp
