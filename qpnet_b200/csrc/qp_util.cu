// Error plumbing, device checks and the small integer / fp64 kernels of the path:
// mu-law codec, F0 -> dilated factor, max-ceil reduction, dilation-index builders.
#include <stdarg.h>
#include <string.h>

#include "qp_common.cuh"

namespace qp {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(QP_EARCH, "no CUDA device: %s", cudaGetErrorString(e));
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return set_error(QP_EARCH, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (major != 10)
    return set_error(QP_EARCH, "libqpnet_b200 is built for sm_100a only (device is cc %d.x); no fallback", major);
  return QP_OK;
}

int check_arch(const QpArch* a) {
  if (!a) return set_error(QP_EINVAL, "arch is NULL");
  if (a->n_quantize < 2 || a->n_quantize > 1024) return set_error(QP_EINVAL, "n_quantize out of range");
  if (a->n_aux < 1 || a->n_aux > 64) return set_error(QP_EINVAL, "n_aux must be in [1,64]");
  if (a->n_resch < 4 || a->n_resch % 4) return set_error(QP_EINVAL, "n_resch must be a multiple of 4");
  if (a->n_skipch < 1) return set_error(QP_EINVAL, "n_skipch must be positive");
  if (a->upsampling < 1) return set_error(QP_EINVAL, "upsampling must be positive");
  if (a->n_fixed < 1 || a->n_fixed > QP_MAX_LAYERS || a->n_adaptive < 1 || a->n_adaptive > QP_MAX_LAYERS)
    return set_error(QP_EINVAL, "layer counts out of range");
  for (int i = 0; i < a->n_fixed; ++i)
    if (a->dil_fixed[i] < 1) return set_error(QP_EINVAL, "bad fixed dilation");
  for (int i = 0; i < a->n_adaptive; ++i)
    if (a->dil_adaptive[i] < 1) return set_error(QP_EINVAL, "bad adaptive dilation");
  return QP_OK;
}

// ------------------------------------------------------------------ mu-law (qpnet.py:22-45)
__global__ void mulaw_encode_kernel(const double* __restrict__ x, int64_t n, double mu, int64_t* __restrict__ y) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x[i];
  double sgn = (v > 0.0) - (v < 0.0);
  double fx = sgn * log(1.0 + mu * fabs(v)) / log(1.0 + mu);
  y[i] = (int64_t)floor((fx + 1.0) / 2.0 * mu + 0.5);
}

__global__ void mulaw_decode_kernel(const int64_t* __restrict__ y, int64_t n, double mu, double* __restrict__ x) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double fx = ((double)y[i] - 0.5) / mu * 2.0 - 1.0;
  double sgn = (fx > 0.0) - (fx < 0.0);
  x[i] = sgn / mu * (pow(1.0 + mu, fabs(fx)) - 1.0);
}

// ------------------------------------------------------------------ F0 -> d
// qpnet_train.py:147-165,178 ; qpnet_decode.py:90-108 ; utils.py:216-235
__global__ void f0_to_dilated_kernel(const double* __restrict__ f0, int B, int F, double fs, double dense,
                                     int U, double f0_floor, double* __restrict__ d64, float* __restrict__ d32) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * F * U;
  if (i >= total) return;
  int64_t frame = i / U;  // (b*F + f)
  double f = f0[frame];
  if (f0_floor >= 0.0 && f < f0_floor) f = f0_floor;
  if (f == 0.0) f = fs / dense;
  double d = __ddiv_rn(__ddiv_rn(1.0 * fs, f), dense);
  if (d64) d64[i] = d;
  if (d32) d32[i] = (float)d;
}

// ------------------------------------------------------------------ decode front end (qpnet_decode.py:163-200)
// raw features (B, Fmax, D) fp64, zero padded past n_frames[b] -> per frame f < n_frames[b]:
//   f0' = raw[f0_dim] * f0_factor (172-173); d = ((1.0*fs)/(f0' == 0 ? fs/dense : f0'))/dense, repeated U times (174-175, 90-108);
//   h[k] = float((x_k - mean_k) / scale_k) with x_f0dim = f0' (StandardScaler.transform, 268-269), transposed to (B, D, Fmax) (190).
// Padding frames give h = 0 and d = 0 (pad_list, 73-88).  One thread per (b, f, k).
__global__ void feat_prepare_kernel(const double* __restrict__ raw, const int32_t* __restrict__ n_frames, int B, int Fmax,
                                    int D, const double* __restrict__ mean, const double* __restrict__ scale,
                                    double f0_factor, int f0_dim, double fs, double dense, int U,
                                    float* __restrict__ h, double* __restrict__ d64, float* __restrict__ d32) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * Fmax * D) return;
  const int k = (int)(i % D);
  const int64_t bf = i / D;
  const int f = (int)(bf % Fmax), b = (int)(bf / Fmax);
  const bool live = f < n_frames[b];
  double x = live ? raw[i] : 0.0;
  if (k == f0_dim && live) x = __dmul_rn(x, f0_factor);
  float hv = 0.f;
  if (live) hv = (float)__ddiv_rn(__dsub_rn(x, mean[k]), scale[k]);
  h[((int64_t)b * D + k) * Fmax + f] = hv;
  if (k == f0_dim) {
    double dv = 0.0;
    if (live) {
      double f0 = x == 0.0 ? fs / dense : x;
      dv = __ddiv_rn(__ddiv_rn(1.0 * fs, f0), dense);
    }
    const int64_t o = ((int64_t)b * Fmax + f) * U;
    for (int u = 0; u < U; ++u) {
      if (d64) d64[o + u] = dv;
      if (d32) d32[o + u] = (float)dv;
    }
  }
}

// ------------------------------------------------------------------ decode back end (qpnet_decode.py:315-318)
// symbol -> decode_mu_law (qpnet.py:34-45, fp64) -> * 32768 -> clip [-32768, 32767] -> int16 (numpy astype: truncation)
__device__ __forceinline__ int16_t mulaw_pcm16_of(int y, double mu) {
  double fx = ((double)y - 0.5) / mu * 2.0 - 1.0;
  double sgn = (fx > 0.0) - (fx < 0.0);
  double w = sgn / mu * (pow(1.0 + mu, fabs(fx)) - 1.0) * 32768.0;
  w = fmin(fmax(w, -32768.0), 32767.0);
  return (int16_t)w;   // C conversion truncates toward zero, like ndarray.astype(np.int16)
}
__global__ void mulaw_pcm16_kernel(const int32_t* __restrict__ y, int64_t n, double mu, int16_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = mulaw_pcm16_of(y[i], mu);
}
// the 256-entry table the tcgen05 generator indexes in its output stage (same arithmetic: bit-identical)
__global__ void pcm_lut_kernel(int16_t* __restrict__ lut, int n, double mu) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lut[i] = mulaw_pcm16_of(i, mu);
}

// PCM output stage for the generators that do not write PCM themselves: symbols (B, ld) -> int16 (B, ld_out)
__global__ void mulaw_pcm16_rows_kernel(const int32_t* __restrict__ y, long long ld, int B, int n_steps, double mu,
                                        int16_t* __restrict__ out, long long ld_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * n_steps) return;
  const int b = (int)(i / n_steps), t = (int)(i % n_steps);
  out[(long long)b * ld_out + t] = mulaw_pcm16_of(y[(long long)b * ld + t], mu);
}

template <typename T>
__global__ void max_ceil_kernel(const T* __restrict__ d, int64_t n, int32_t* __restrict__ out) {
  int best = INT32_MIN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T v = d[i];
    if (v == v) {  // nanmax semantics (qpnet.py:350)
      int c = (int)ceil((double)v);
      best = max(best, c);
    }
  }
  for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best != INT32_MIN) atomicMax(out, best);
}

__global__ void set_int_kernel(int32_t* p, int32_t v) { *p = v; }

// ------------------------------------------------------------------ index builders
template <int FLAVOUR, typename TIn, typename TOut>
__global__ void index_kernel(const TIn* __restrict__ d, int B, int n, int64_t ld, int dil, TOut* __restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * n) return;
  int b = (int)(i / n), t = (int)(i % n);
  TIn v = d[(int64_t)b * ld + t];
  int r;
  if (FLAVOUR == 0) r = tf_index_f32((float)v, dil, t - n);
  else if (FLAVOUR == 1) r = tf_index_f64((double)v, dil, t - n);
  else if (FLAVOUR == 2) r = gen_index_f32((float)v, dil);
  else r = gen_index_f64((double)v, dil);
  idx[i] = (TOut)r;
}

}  // namespace qp

namespace qp {
int pcm_lut_fill(int16_t* lut, int n_quantize, cudaStream_t st) {
  pcm_lut_kernel<<<(n_quantize + 255) / 256, 256, 0, st>>>(lut, n_quantize, (double)(n_quantize - 1));
  QP_LAUNCH_CHECK();
  return QP_OK;
}
int pcm16_rows(const int32_t* sym, long long ld, int B, int n_steps, int n_quantize, int16_t* out, long long ld_out, cudaStream_t st) {
  const long long n = (long long)B * n_steps;
  if (n == 0) return QP_OK;
  mulaw_pcm16_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sym, ld, B, n_steps, (double)(n_quantize - 1), out, ld_out);
  QP_LAUNCH_CHECK();
  return QP_OK;
}
}  // namespace qp

using namespace qp;

extern "C" {

int qp_abi_version(void) { return QP_ABI_VERSION; }
const char* qp_last_error(void) { return g_err; }
int qp_last_launch_count(void) { return g_launches; }
int qp_device_ok(void) { return check_device(); }

int qp_num_tensors(const QpArch* arch) {
  if (check_arch(arch) != QP_OK) return QP_EINVAL;
  return tensor_map(arch).count();
}

int qp_mulaw_encode(const double* x, int64_t n, int32_t mu, int64_t* y, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(x && y && n >= 0 && mu > 1, "qp_mulaw_encode: bad arguments");
  reset_launch_count();
  if (n == 0) return QP_OK;
  mulaw_encode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, (double)(mu - 1), y);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

int qp_mulaw_decode(const int64_t* y, int64_t n, int32_t mu, double* x, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(x && y && n >= 0 && mu > 1, "qp_mulaw_decode: bad arguments");
  reset_launch_count();
  if (n == 0) return QP_OK;
  mulaw_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, n, (double)(mu - 1), x);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

int qp_f0_to_dilated(const double* f0, int32_t B, int32_t F, double fs, double dense, int32_t U,
                     double f0_floor, double* d64, float* d32, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(f0 && B >= 0 && F >= 0 && U > 0 && fs > 0 && dense > 0, "qp_f0_to_dilated: bad arguments");
  QP_REQUIRE(d64 || d32, "qp_f0_to_dilated: no output requested");
  reset_launch_count();
  int64_t total = (int64_t)B * F * U;
  if (total == 0) return QP_OK;
  f0_to_dilated_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(f0, B, F, fs, dense, U,
                                                                                         f0_floor, d64, d32);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

}  // extern "C"
template <typename T>
static int max_ceil_impl(const T* d, int64_t n, int32_t* out, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(d && out && n > 0, "qp_max_ceil: bad arguments");
  reset_launch_count();
  set_int_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out, INT32_MIN);
  QP_LAUNCH_CHECK();
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 8) blocks = 148 * 8;
  max_ceil_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(d, n, out);
  QP_LAUNCH_CHECK();
  return QP_OK;
}
extern "C" {
int qp_max_ceil_f32(const float* d, int64_t n, int32_t* out, void* stream) { return max_ceil_impl(d, n, out, stream); }
int qp_max_ceil_f64(const double* d, int64_t n, int32_t* out, void* stream) { return max_ceil_impl(d, n, out, stream); }

}  // extern "C"
template <int FL, typename TIn, typename TOut>
static int index_impl(const TIn* d, int32_t B, int32_t n, int64_t ld, int32_t dil, TOut* idx, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(B >= 0 && n >= 0 && dil >= 1 && ld >= n, "qp_index: bad shape (B=%d n=%d ld=%lld dil=%d)", B, n,
             (long long)ld, dil);
  reset_launch_count();
  int64_t total = (int64_t)B * n;
  if (total == 0) return QP_OK;  // empty input: nothing to do (ragged/empty edge case)
  QP_REQUIRE(d && idx, "qp_index: NULL pointer");
  index_kernel<FL, TIn, TOut><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, B, n, ld, dil, idx);
  QP_LAUNCH_CHECK();
  return QP_OK;
}
extern "C" {
int qp_index_tf_f32(const float* d, int32_t B, int32_t n, int64_t ld, int32_t dil, int64_t* idx, void* s) {
  return index_impl<0>(d, B, n, ld, dil, idx, s);
}
int qp_index_tf_f64(const double* d, int32_t B, int32_t n, int64_t ld, int32_t dil, int32_t* idx, void* s) {
  return index_impl<1>(d, B, n, ld, dil, idx, s);
}
int qp_index_gen_f32(const float* d, int32_t B, int32_t n, int64_t ld, int32_t dil, int64_t* idx, void* s) {
  return index_impl<2>(d, B, n, ld, dil, idx, s);
}
int qp_index_gen_f64(const double* d, int32_t B, int32_t n, int64_t ld, int32_t dil, int32_t* idx, void* s) {
  return index_impl<3>(d, B, n, ld, dil, idx, s);
}

int qp_feat_prepare(const double* raw, const int32_t* n_frames, int32_t B, int32_t Fmax, int32_t D, const double* mean,
                    const double* scale, double f0_factor, int32_t f0_dim, double fs, double dense, int32_t U, float* h,
                    double* d64, float* d32, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(B >= 0 && Fmax >= 0 && D >= 1 && U >= 1 && f0_dim >= 0 && f0_dim < D, "feat_prepare: bad shape");
  QP_REQUIRE(fs > 0 && dense > 0, "feat_prepare: fs and dense_factor must be positive");
  if ((int64_t)B * Fmax == 0) return QP_OK;
  QP_REQUIRE(raw && n_frames && mean && scale && h && (d64 || d32), "feat_prepare: NULL pointer");
  int64_t n = (int64_t)B * Fmax * D;
  feat_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(raw, n_frames, B, Fmax, D, mean, scale,
                                                                                  f0_factor, f0_dim, fs, dense, U, h, d64, d32);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

int qp_mulaw_decode_pcm16(const int32_t* y, int64_t n, int32_t mu, int16_t* pcm, void* stream) {
  if (int e = check_device()) return e;
  QP_REQUIRE(n >= 0 && mu >= 2, "mulaw_decode_pcm16: bad size");
  if (n == 0) return QP_OK;
  QP_REQUIRE(y && pcm, "mulaw_decode_pcm16: NULL pointer");
  mulaw_pcm16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, n, (double)(mu - 1), pcm);   // qpnet.py:41: mu = mu - 1
  QP_LAUNCH_CHECK();
  return QP_OK;
}

}  // extern "C"
