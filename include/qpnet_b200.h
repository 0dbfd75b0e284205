/*
 * qpnet_b200.h -- C ABI of libqpnet_b200.so (sm_100a only).
 *
 * The reference (bigpon/QPNet) has NO foreign-function interface: its hot path is the
 * Python class `QPNet` in src/nets/qpnet.py, called from four sites
 * (src/bin/qpnet_train.py:520, qpnet_update.py:485, qpnet_validate.py:421 ->
 * `forward`; qpnet_decode.py:312-314 -> `batch_fast_generate`).  This header is the
 * boundary a replacement binds underneath that class: plain pointers and sizes, no
 * torch types.  Each entry point cites the reference lines it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *  - functions return 0 on success, a negative QP_E* code otherwise, and never
 *    allocate device memory: the caller sizes a workspace with *_workspace_bytes();
 *  - qp_last_error() returns a thread-local, human readable message;
 *  - there is no CPU fallback: every compute entry fails with QP_EARCH unless the
 *    current device is compute capability 10.x.
 */
#ifndef QPNET_B200_H_
#define QPNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QP_ABI_VERSION 3
#define QP_MAX_LAYERS 64

enum {
  QP_OK = 0,
  QP_EINVAL = -1,     /* bad argument / shape */
  QP_EARCH = -2,      /* not an sm_100 device, or unsupported architecture option */
  QP_ECUDA = -3,      /* a CUDA runtime call failed; see qp_last_error() */
  QP_EWORKSPACE = -4, /* workspace too small */
  QP_ERANGE = -5,     /* gather index out of range (the reference's assert, qpnet.py:294) */
  QP_ETIMEOUT = -6    /* persistent generator watchdog fired */
};

/* Hyper-parameters of QPNet.__init__ (qpnet.py:174-199). kernel_size is fixed at 2. */
typedef struct QpArch {
  int32_t n_quantize;  /* Q, 256 */
  int32_t n_aux;       /* A, 39  */
  int32_t n_resch;     /* C, 512 */
  int32_t n_skipch;    /* S, 256 */
  int32_t upsampling;  /* U, 110 */
  int32_t n_fixed;     /* len(dilationsF) */
  int32_t n_adaptive;  /* len(dilationsA) */
  int32_t dil_fixed[QP_MAX_LAYERS];
  int32_t dil_adaptive[QP_MAX_LAYERS];
} QpArch;

/*
 * Parameters travel as a host array of device pointers to fp32 tensors in the
 * reference's state_dict() order (qpnet.py:201-235):
 *   causal.conv.{weight,bias}; upsampling.conv.{weight,bias};
 *   dilF_sigmoid.i.conv.{w,b} (all i); dilF_tanh.i.conv.{w,b};
 *   auxF_1x1_sigmoid.i.{w,b}; auxF_1x1_tanh.i.{w,b}; skipF_1x1.i.{w,b}; resF_1x1.i.{w,b};
 *   dilA_sigmoid.i.convC.{w,b}, .convP.{w,b}; dilA_tanh.i. ... ;
 *   auxA_1x1_sigmoid.i.{w,b}; auxA_1x1_tanh.i.{w,b}; skipA_1x1.i.{w,b}; resA_1x1.i.{w,b};
 *   conv_post_1.{w,b}; conv_post_2.{w,b}.
 * qp_num_tensors() returns the expected count (216 for the SI default model).
 */
int qp_abi_version(void);
const char* qp_last_error(void);
int qp_num_tensors(const QpArch* arch);
/* 0 if the current CUDA device can run this library (cc 10.x), QP_EARCH otherwise. */
int qp_device_ok(void);

/* ---- mu-law codec (qpnet.py:22-45), fp64 arithmetic on device -------------------- */
int qp_mulaw_encode(const double* x, int64_t n, int32_t mu, int64_t* y, void* stream);
int qp_mulaw_decode(const int64_t* y, int64_t n, int32_t mu, double* x, void* stream);

/* ---- caller-side F0 -> per-sample dilated factor ----------------------------------
 * qpnet_train.py:147-165,178 / qpnet_decode.py:90-108 / utils/utils.py:216-235:
 * f0 < f0_floor -> f0_floor (pass a negative floor to disable, decode side);
 * f0 == 0 -> fs/dense;  d = ((1.0*fs)/f0)/dense in fp64;  each frame repeated U times.
 * f0: (B, F) fp64.  Outputs (either may be NULL): d64 (B, F*U) fp64, d32 (B, F*U) fp32
 * (the fp64 value rounded once, like torch.from_numpy(d).float(), qpnet_train.py:298). */
int qp_f0_to_dilated(const double* f0, int32_t B, int32_t F, double fs, double dense,
                     int32_t U, double f0_floor, double* d64, float* d32, void* stream);

/* ---- decode front end (qpnet_decode.py:163-200, 268-269), one launch for a padded batch ------------
 * raw: (B, Fmax, D) fp64 unscaled acoustic features, zero padded past n_frames[b] (device, (B,) int32).
 * Per live frame: f0' = raw[f0_dim] * f0_factor (172-173); d = ((1.0*fs)/(f0' == 0 ? fs/dense : f0'))/dense
 * repeated U times (174-175, 90-108, utils.py:216-235); h = float((x - mean)/scale) with x[f0_dim] = f0'
 * (StandardScaler.transform), written transposed as (B, D, Fmax) fp32 (190).  Padding frames: h = 0, d = 0
 * (pad_list, 73-88).  d64 / d32: (B, Fmax*U); either may be NULL. */
int qp_feat_prepare(const double* raw, const int32_t* n_frames, int32_t B, int32_t Fmax, int32_t D,
                    const double* mean, const double* scale, double f0_factor, int32_t f0_dim,
                    double fs, double dense, int32_t U, float* h, double* d64, float* d32, void* stream);

/* ---- decode back end (qpnet_decode.py:315-318): symbols -> decode_mu_law (fp64) -> *32768 -> clip -> int16 */
int qp_mulaw_decode_pcm16(const int32_t* y, int64_t n, int32_t mu, int16_t* pcm, void* stream);

/* max over all elements of ceil(d)  (qpnet.py:255 / 347-350); result to *out (device). */
int qp_max_ceil_f32(const float* d, int64_t n, int32_t* out, void* stream);
int qp_max_ceil_f64(const double* d, int64_t n, int32_t* out, void* stream);

/* ---- dilation-index builders (qpnet.py:592-624), bit-exact ----------------------
 * d: (B, n) with row stride ld.  One value per (b, t); the reference's repeat over the
 * channel axis (qpnet.py:604,610,618,623) is never materialised.
 *  tf_f32 : int64( rint_f32( (-d*dil)_f32 + float(t-n) ) )       qpnet.py:594-600
 *  tf_f64 : int32( rint_f64( (-d*dil) + (t-n) ) )                qpnet.py:606-609
 *  gen_f32: int64( rint_f32( -d*dil ) )                          qpnet.py:615-617
 *  gen_f64: int32( rint_f64( -d*dil ) )                          qpnet.py:621-622 */
int qp_index_tf_f32(const float* d, int32_t B, int32_t n, int64_t ld, int32_t dil,
                    int64_t* idx, void* stream);
int qp_index_tf_f64(const double* d, int32_t B, int32_t n, int64_t ld, int32_t dil,
                    int32_t* idx, void* stream);
int qp_index_gen_f32(const float* d, int32_t B, int32_t n, int64_t ld, int32_t dil,
                     int64_t* idx, void* stream);
int qp_index_gen_f64(const double* d, int32_t B, int32_t n, int64_t ld, int32_t dil,
                     int32_t* idx, void* stream);

/* ---- teacher-forced stack: QPNet.forward (qpnet.py:239-312) + its backward --------
 * x: (B, T) int64 symbols; h: (B, A, F) fp32 frame-rate aux, F*U >= T; d: (B, T) fp32;
 * all three are consumed from their END exactly like the reference slices them.
 * M = max ceil(d) over the whole d passed (qpnet.py:255), supplied by the caller
 * (qp_max_ceil_f32).  logits: (B, bl, Q) fp32.  Unlike the reference (caveat C1 in
 * qpnet.py:250 gathers every element's past taps from batch element 0) each
 * batch element gathers from itself.
 * flags: QP_F_SAVE keeps the activations backward needs inside `ws`;
 *        QP_F_BF16 selects the bf16 tensor-core path (fp32 SIMT otherwise). */
#define QP_F_SAVE 1u
#define QP_F_BF16 2u
size_t qp_forward_workspace_bytes(const QpArch* arch, int32_t B, int32_t T, int32_t bl,
                                  int32_t M, uint32_t flags);
int qp_forward(const QpArch* arch, const float* const* tensors_host, const int64_t* x,
               const float* h, const float* d, int32_t B, int32_t T, int32_t F, int32_t bl,
               int32_t M, float* logits, void* ws, size_t ws_bytes, uint32_t flags,
               void* stream);
/* dlogits: (B, bl, Q) fp32.  grads_host: host array (same order as tensors_host) of
 * device pointers that are OVERWRITTEN with dLoss/dParam (all entries required; the dead
 * last resA_1x1 projection, caveat C7, receives zeros).  `ws` must be the
 * workspace a QP_F_SAVE forward of the same shapes filled; x/h/d are the same inputs. */
int qp_backward(const QpArch* arch, const float* const* tensors_host, const int64_t* x,
                const float* h, const float* d, int32_t B, int32_t T, int32_t F, int32_t bl,
                int32_t M, const float* dlogits, float* const* grads_host, void* ws,
                size_t ws_bytes, uint32_t flags, void* stream);

/* The same backward in stages, for data-parallel callers that overlap the gradient all-reduce with the rest of the
 * backward (the reference gets this exchange implicitly from nn.DataParallel's reduce_add, qpnet_train.py:416-423):
 * stage 0 = the two head layers, stage s in [1, L] = residual block L - s (last block first), stage L + 1 = causal layer
 * and upsampler.  Run [stage_begin, stage_end) in ascending, non-overlapping ranges starting at 0; when a call returns,
 * the launches that write the gradients of its stages are enqueued on `stream` (record an event, reduce, carry on). */
int qp_backward_range(const QpArch* arch, const float* const* tensors_host, const int64_t* x,
                      const float* h, const float* d, int32_t B, int32_t T, int32_t F, int32_t bl,
                      int32_t M, const float* dlogits, float* const* grads_host, void* ws,
                      size_t ws_bytes, uint32_t flags, int32_t stage_begin, int32_t stage_end, void* stream);

/* fused softmax cross-entropy over (rows, Q) logits (qpnet_train.py:426-430,526):
 * loss_sum[0] += sum_r -log softmax(logits[r])[target[r]];  dlogits = (p - onehot)*scale.
 * A target outside [0, Q) (the reference asserts it away, qpnet_train.py:524) turns the loss into NaN. */
int qp_cross_entropy(const float* logits, const int64_t* target, int64_t rows, int32_t Q,
                     float scale, float* loss_sum, float* dlogits, void* stream);

/* Adam step over flat fp32 buffers (torch.optim.Adam(lr 1e-4, wd 0), qpnet_train.py:426-428; .step() at 531):
 * param / grad / exp_avg / exp_avg_sq hold n elements (n % 4 == 0, 16-byte aligned), `step` counts from 1,
 * `grad_scale` multiplies the gradient first (1 / world after a summed data-parallel all-reduce).  fp32 arithmetic in
 * torch's operation order; one HBM-bound pass. */
int qp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                 double beta1, double beta2, double eps, int32_t step, double grad_scale, void* stream);

/* ---- autoregressive generator: QPNet.batch_fast_generate (qpnet.py:314-559) -------
 * One persistent cooperative kernel runs priming + every sample step for the batch. */
enum { QP_MODE_SAMPLING = 0, QP_MODE_ARGMAX = 1 };

typedef struct QpGenerateArgs {
  int32_t B;            /* utterances in this call                                   */
  int32_t F;            /* frames of the padded aux tensor                           */
  int32_t M;            /* max ceil(d) over the padded batch (qpnet.py:347-350)      */
  int32_t mode;         /* QP_MODE_*  (any other value: QP_EINVAL; the reference
                           calls sys.exit(1), qpnet.py:513-515)                      */
  int32_t max_steps;    /* number of sample steps to run (= max n_samples)           */
  int32_t d_is_f64;     /* 1: d is fp64 (extra_memory=False flavour, gen_f64 index)  */
  const int64_t* seed;  /* (B,) last symbol of the seed x (qpnet.py:356-358,443)     */
  const float* h;       /* (B, A, F) fp32, zero padded                               */
  const void* d;        /* (B, F*U) fp64 or fp32, zero padded                        */
  const int32_t* n_samples; /* (B,) symbols wanted per utterance                     */
  const float* uniforms;    /* (B, ld_uniforms) pre-drawn U[0,1) or NULL -> Philox   */
  int64_t ld_uniforms;
  uint64_t philox_seed;
  const int32_t* force;     /* optional (B, ld_force): symbols fed back instead of the
                               drawn ones (teacher-forced generator check)           */
  int64_t ld_force;
  int32_t* out;         /* (B, ld_out) generated symbols                             */
  int64_t ld_out;
  float* logits_out;    /* optional (B, max_steps, Q) per-step logits                */
  const int32_t* utt_ids;   /* optional (B,) device: caller-side index of each utterance;
                               keys the in-kernel Philox stream, so that an utterance
                               draws the same numbers whichever launch / slot it runs in
                               (NULL: the slot index)                                    */
  int16_t* out_pcm;     /* optional (B, ld_out_pcm): the generated waveform as 16-bit PCM,
                           decode_mu_law(symbol) * 32768 clipped (qpnet_decode.py:315-318),
                           written by the generator as its output stage                */
  int64_t ld_out_pcm;
} QpGenerateArgs;

size_t qp_generate_workspace_bytes(const QpArch* arch, int32_t B, int32_t M);
int qp_generate(const QpArch* arch, const float* const* tensors_host,
                const QpGenerateArgs* args, void* ws, size_t ws_bytes, void* stream);
/* Blocking read of the status word qp_forward / qp_generate leave at the start of their
 * workspace: 0, QP_ERANGE (the reference's gather assert, qpnet.py:294) or QP_ETIMEOUT. */
int qp_workspace_status(const void* ws, void* stream);
/* kernels launched by the most recent compute call on this thread (bench accounting) */
int qp_last_launch_count(void);
/* diagnostics: operand segments the tcgen05 weight-gradient launches of this process have bound to TMA descriptors
 * (cp.async.bulk.tensor) so far; 0 means every segment went through the cp.async producers */
int64_t qp_debug_tma_segments(void);

#ifdef __cplusplus
}
#endif
#endif /* QPNET_B200_H_ */
