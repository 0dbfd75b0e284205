// Optimizer step of the training path (reference: torch.optim.Adam(lr 1e-4, betas (0.9, 0.999), eps 1e-8, wd 0),
// qpnet_train.py:426-428, .step() at 531).  The trainer keeps parameters, gradients and both moments as FLAT fp32
// buffers with one layout (qpnet_b200/train.py), so the whole model is ONE element-wise pass: 5 reads + 3 writes of
// 4 bytes per parameter, HBM-bound (24.2 M parameters -> 0.77 GB per step -> ~0.12 ms at the measured 6.5 TB/s).
//
// Arithmetic follows torch's Adam in fp32 operation by operation:
//   m <- m + (1 - beta1) (g - m);  v <- beta2 v + (1 - beta2) g g
//   p <- p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// `grad_scale` multiplies the gradient first (1 / world: the data-parallel mean folded into the step, so the summed
// all-reduce result is used as it is).  The bias corrections arrive pre-computed in double precision from the host.
#include <math.h>

#include <algorithm>

#include "qp_common.cuh"

namespace qp {

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                   float4* __restrict__ v, long long n4, float grad_scale, float w1, float beta2,
                                                   float w2, float step_size, float bc2_sqrt, float eps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G[k] * grad_scale;
      M[k] = M[k] + w1 * (gr - M[k]);
      V[k] = beta2 * V[k] + w2 * gr * gr;
      const float denom = sqrtf(V[k]) / bc2_sqrt + eps;
      P[k] = P[k] - step_size * (M[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace qp

using namespace qp;

extern "C" {

int qp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                 double beta2, double eps, int32_t step, double grad_scale, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam: NULL pointer");
  QP_REQUIRE(n >= 0 && n % 4 == 0, "adam: the flat buffers hold a multiple of 4 elements (16-byte aligned pieces)");
  QP_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "adam: buffers must be 16-byte aligned");
  QP_REQUIRE(step >= 1, "adam: step counts from 1");
  reset_launch_count();
  if (n == 0) return QP_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const long long n4 = n / 4;
  const int blocks = (int)std::min<long long>((n4 + 255) / 256, 148LL * 8);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq, n4,
                                                        (float)grad_scale, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                        (float)(lr / bc1), (float)sqrt(bc2), (float)eps);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

}  // extern "C"
