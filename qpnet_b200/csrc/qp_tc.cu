// bf16 tensor-core path of the teacher-forced stack (QP_F_BF16): tcgen05.mma GEMMs with the
// gate / residual / skip / head math fused into the TMEM epilogue.
//
//   D[128 time rows x BN] (fp32, TMEM) = A[128 x K] (bf16, smem) * W[BN x K]^T (bf16, smem)
//
// * A rows are GATHERED: up to three K-segments [x(past row) | x(current row) | h_up] of
//   time-major bf16 activations (qpnet.py:295-298, 657-666); a producer thread owns one tile row
//   and copies its 128-byte K-slices with cp.async into the canonical K-major SWIZZLE_128B
//   layout tcgen05 reads (8-row x 128-byte atoms, 16-byte pieces XOR-ed with the row index).
// * warps 0-3: producers (cp.async ring of STAGES slots, completion published per slot through
//   an mbarrier after fence.proxy.async), afterwards the epilogue (tcgen05.ld 32x32b, one
//   TMEM lane = one time row per thread).  warp 4: TMEM allocation; its lane 0 issues the MMAs
//   and releases slots with tcgen05.commit.
// * one output tile per CTA; grid = (row tiles, column tiles, batch).
#include "qp_tc.cuh"

namespace qp {
namespace tc {

constexpr int BM = 128;      // UMMA M: time rows per tile
constexpr int BK = 64;       // bf16 elements per stage row = one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int LAG = 2;       // cp.async groups in flight per producer thread
constexpr int PRODUCERS = 128;
constexpr int THREADS = 160;
constexpr int TMEM_COLS = 256;
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000LL;   // ~2 s: a lost barrier traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  unsigned spins = 0;
  while (true) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if ((++spins & 0xFFFu) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) __trap();
    }
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO = 1 in [16,30) (unused for swizzled K-major), SBO = 1024 B >> 4
// in [32,46) (stride between 8-row atoms), version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major.
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ float fsigmoid(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) {
  float e = __expf(-2.f * fabsf(x));
  return copysignf((1.f - e) / (1.f + e), x);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(Args a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;              // SWIZZLE_128B atoms need 1024-byte alignment
  unsigned char* sm = smem_raw + (sbase - raw);
  const int stage_bytes = BM * 128 + a.BN * 128;
  const uint32_t bar0 = sbase + STAGES * stage_bytes;         // full[STAGES], empty[STAGES], accum : 8 bytes each
  uint32_t* tmem_slot = (uint32_t*)(sm + STAGES * stage_bytes + (2 * STAGES + 1) * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, r0 = blockIdx.x * BM, n0 = a.n_begin + blockIdx.y * a.BN;
  const int ncols = min(a.BN, a.N - n0);
  int nk = 0;
  for (int s = 0; s < a.nseg; ++s) nk += a.seg[s].K / BK;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(bar0 + 8 * i, PRODUCERS); mbar_init(bar0 + 8 * (STAGES + i), 1); }
    mbar_init(bar0 + 8 * 2 * STAGES, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ================================================================== producer
    const int rr = min(r0 + tid, a.n_rows - 1);
    const uint32_t row_off = (uint32_t)((tid >> 3) * 1024 + (tid & 7) * 128), sw = (uint32_t)(tid & 7);
    const int w0 = min(n0 + tid, a.N - 1), w1 = min(n0 + 128 + tid, a.N - 1);
    // a thread only copies the weight rows that exist in this tile (BN may be 32..256 in steps of 32)
    const bool ldb0 = tid < a.BN, ldb1 = 128 + tid < a.BN;
    int kc = 0, koff = 0;
    for (int s = 0; s < a.nseg; ++s) {
      const Seg sg = a.seg[s];
      int src = sg.rowmap ? sg.rowmap[(long long)b * a.n_rows + rr] : rr + sg.row_off;
      src = max(0, min(src, sg.src_rows - 1));
      const __nv_bfloat16* arow = sg.base + (long long)b * sg.bstride + (long long)src * sg.ld;
      const __nv_bfloat16* wrow0 = a.W + (long long)w0 * a.ldw + koff;
      const __nv_bfloat16* wrow1 = a.W + (long long)w1 * a.ldw + koff;
      for (int k0 = 0; k0 < sg.K; k0 += BK, ++kc) {
        const int stage = kc % STAGES, it = kc / STAGES;
        if (it > 0) mbar_wait(bar0 + 8 * (STAGES + stage), (uint32_t)(it - 1) & 1u);
        const uint32_t sA = sbase + stage * stage_bytes + row_off;
        const uint32_t sB = sA + BM * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) cp_async16(sA + ((c ^ sw) << 4), arow + k0 + c * 8);
        if (ldb0) {
#pragma unroll
          for (int c = 0; c < 8; ++c) cp_async16(sB + ((c ^ sw) << 4), wrow0 + k0 + c * 8);
        }
        if (ldb1) {
#pragma unroll
          for (int c = 0; c < 8; ++c) cp_async16(sB + 16 * 1024 + ((c ^ sw) << 4), wrow1 + k0 + c * 8);
        }
        cp_async_commit();
        if (kc >= LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          mbar_arrive(bar0 + 8 * ((kc - LAG) % STAGES));
        }
      }
      koff += sg.K;
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int j = max(0, nk - LAG); j < nk; ++j) mbar_arrive(bar0 + 8 * (j % STAGES));

    // ================================================================== epilogue
    mbar_wait(bar0 + 8 * 2 * STAGES, 0);
    tc_fence_after();
    const int r = r0 + tid;
    const bool rv = r < a.n_rows;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < ncols; g += 32) {
      uint32_t v[32];
      tmem_ld32(trow + g, v);
      const int n = n0 + g;
      if (EPI == EPI_GATE) {
        // columns (2c, 2c+1) = (sigmoid, tanh) pre-activations of channel c  (qpnet.py:665-666 / 634-635)
        float z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float sg_ = fsigmoid(__uint_as_float(v[2 * j]) + __ldg(a.bias + n + 2 * j));
          float th_ = ftanh(__uint_as_float(v[2 * j + 1]) + __ldg(a.bias + n + 2 * j + 1));
          z[j] = sg_ * th_;
          v[2 * j] = __float_as_uint(sg_);
          v[2 * j + 1] = __float_as_uint(th_);
        }
        if (rv) {
          const int halfN = a.N >> 1, c0 = n >> 1;
          uint4* zb = (uint4*)(a.z_bf + ((long long)b * a.n_rows + r) * halfN + c0);
          zb[0] = make_uint4(pack_bf16(z[0], z[1]), pack_bf16(z[2], z[3]), pack_bf16(z[4], z[5]), pack_bf16(z[6], z[7]));
          zb[1] = make_uint4(pack_bf16(z[8], z[9]), pack_bf16(z[10], z[11]), pack_bf16(z[12], z[13]), pack_bf16(z[14], z[15]));
          if (a.z_f32) {
            float4* zf = (float4*)(a.z_f32 + ((long long)b * a.n_rows + r) * halfN + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) zf[j] = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
          }
          if (a.gsave) {
            uint4* gs = (uint4*)(a.gsave + ((long long)b * a.n_rows + r) * a.N + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) gs[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
      } else if (EPI == EPI_RESSKIP) {
        if (rv) {
          if (n < a.C) {
            // residual projection + x(current row), fp32 stream + bf16 operand copy (qpnet.py:668-669)
            const float4* xc = (const float4*)(a.xcur + (long long)b * a.xcur_bstride + (long long)(r + a.xcur_off) * a.C + n);
            float4* xo = (float4*)(a.xnext + ((long long)b * a.n_rows + r) * a.C + n);
            uint4* xb = (uint4*)(a.xnext_bf + ((long long)b * a.n_rows + r) * a.C + n);
            float o[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 x4 = xc[j];
              o[4 * j] = __uint_as_float(v[4 * j]) + __ldg(a.bias + n + 4 * j) + x4.x;
              o[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + __ldg(a.bias + n + 4 * j + 1) + x4.y;
              o[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + __ldg(a.bias + n + 4 * j + 2) + x4.z;
              o[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + __ldg(a.bias + n + 4 * j + 3) + x4.w;
              xo[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              xb[j] = make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                                 pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
          } else if (r >= a.skip_row0) {
            // skip projection, accumulated over the blocks for the last bl rows only (qpnet.py:667, 283)
            float4* sk = (float4*)(a.skip + (long long)b * a.skip_bstride + (long long)(r - a.skip_row0) * a.S + (n - a.C));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = make_float4(__uint_as_float(v[4 * j]) + __ldg(a.bias + n + 4 * j),
                                     __uint_as_float(v[4 * j + 1]) + __ldg(a.bias + n + 4 * j + 1),
                                     __uint_as_float(v[4 * j + 2]) + __ldg(a.bias + n + 4 * j + 2),
                                     __uint_as_float(v[4 * j + 3]) + __ldg(a.bias + n + 4 * j + 3));
              if (a.skip_accum) { float4 p = sk[j]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
              sk[j] = o;
            }
          }
        }
      } else {  // EPI_HEAD: out = acc + bias (fp32, pre-activation) and optionally relu(out) as the next bf16 operand
        if (rv) {
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) + __ldg(a.bias + n + j);
          float4* op = (float4*)(a.out + (long long)b * a.out_bstride + (long long)r * a.ldo + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          if (a.out_relu_bf) {
            uint4* ob = (uint4*)(a.out_relu_bf + (long long)b * a.out_bstride + (long long)r * a.ldo + n);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ob[j] = make_uint4(pack_bf16(fmaxf(o[8 * j], 0.f), fmaxf(o[8 * j + 1], 0.f)),
                                 pack_bf16(fmaxf(o[8 * j + 2], 0.f), fmaxf(o[8 * j + 3], 0.f)),
                                 pack_bf16(fmaxf(o[8 * j + 4], 0.f), fmaxf(o[8 * j + 5], 0.f)),
                                 pack_bf16(fmaxf(o[8 * j + 6], 0.f), fmaxf(o[8 * j + 7], 0.f)));
          }
        }
      }
    }
  } else {
    // ================================================================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, ncols);
      for (int kc = 0; kc < nk; ++kc) {
        const int stage = kc % STAGES, it = kc / STAGES;
        mbar_wait(bar0 + 8 * stage, (uint32_t)it & 1u);
        tc_fence_after();
        const uint64_t ad = umma_desc(sbase + stage * stage_bytes);
        const uint64_t bd = umma_desc(sbase + stage * stage_bytes + BM * 128);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)          // +32 bytes per K = 16 step inside the 128-byte swizzle row
          umma(tmem, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)(kc | k));
        umma_commit(bar0 + 8 * (STAGES + stage));  // slot free once these MMAs retire
      }
      umma_commit(bar0 + 8 * 2 * STAGES);           // accumulator complete
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <int EPI>
static int launch(const Args& a, cudaStream_t st) {
  if (a.n_rows <= 0 || a.B <= 0 || a.N - a.n_begin <= 0) return QP_OK;
  int ktot = 0;
  for (int s = 0; s < a.nseg; ++s) {
    QP_REQUIRE(a.seg[s].K > 0 && a.seg[s].K % BK == 0, "tc gemm: K segment %d not a multiple of %d", a.seg[s].K, BK);
    ktot += a.seg[s].K;
  }
  QP_REQUIRE(a.BN >= 32 && a.BN <= 256 && a.BN % 32 == 0 && (a.N - a.n_begin) % 32 == 0 && a.ldw >= ktot,
             "tc gemm: bad tile shape N=%d BN=%d", a.N, a.BN);
  const size_t smem = (size_t)STAGES * (BM * 128 + a.BN * 128) + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool configured[3] = {false, false, false};
  if (!configured[EPI]) {
    QP_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[EPI] = true;
  }
  dim3 grid((a.n_rows + BM - 1) / BM, (a.N - a.n_begin + a.BN - 1) / a.BN, a.B);
  tc_gemm_kernel<EPI><<<grid, THREADS, smem, st>>>(a);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

int gemm_gate(const Args& a, cudaStream_t st) { return launch<EPI_GATE>(a, st); }
int gemm_resskip(const Args& a, cudaStream_t st) { return launch<EPI_RESSKIP>(a, st); }
int gemm_head(const Args& a, cudaStream_t st) { return launch<EPI_HEAD>(a, st); }

// ------------------------------------------------------------------ small bf16 helpers
__global__ void f32_to_bf16_pad_kernel(const float* __restrict__ src, long long rows, int K, int Kp,
                                       __nv_bfloat16* __restrict__ dst, int relu) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Kp) return;
  long long r = i / Kp;
  int k = (int)(i % Kp);
  float v = k < K ? src[r * K + k] : 0.f;
  if (relu) v = fmaxf(v, 0.f);
  dst[i] = __float2bfloat16_rn(v);
}

int f32_to_bf16_pad(const float* src, long long rows, int K, int Kp, __nv_bfloat16* dst, int relu, cudaStream_t st) {
  long long n = rows * Kp;
  if (n <= 0) return QP_OK;
  f32_to_bf16_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, rows, K, Kp, dst, relu);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

// Wg (L, 2C, Kg = 2C + Ap) fp32 -> (L, 2C, Kgp = 2C + 64) bf16, aux columns zero padded
__global__ void pack_wg_bf16_kernel(const float* __restrict__ Wg, long long rows, int twoC, int Kg, int Kgp,
                                    __nv_bfloat16* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Kgp) return;
  long long r = i / Kgp;
  int k = (int)(i % Kgp);
  dst[i] = __float2bfloat16_rn(k < Kg ? Wg[r * Kg + k] : 0.f);
  (void)twoC;
}

int pack_wg_bf16(const float* Wg, long long rows, int twoC, int Kg, int Kgp, __nv_bfloat16* dst, cudaStream_t st) {
  long long n = rows * Kgp;
  pack_wg_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Wg, rows, twoC, Kg, Kgp, dst);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

}  // namespace tc
}  // namespace qp
