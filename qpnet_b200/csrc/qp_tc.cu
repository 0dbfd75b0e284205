// bf16 tensor-core path of the teacher-forced stack (QP_F_BF16), forward and backward: tcgen05.mma GEMMs with the
// gate / residual / skip / head / dgate / dX math fused into the TMEM epilogue, and the weight-gradient contraction.
//
//   tc_gemm_kernel<EPI>:  D[128 time rows x BN] (fp32, TMEM) = A[128 x K] (bf16, smem) * B[BN x K]^T (bf16, smem)
//   tc_wgrad_kernel:      D[128 i x BJ j]                    = P[rows x 128 i]^T * Q[rows x BJ j]    (both MN-major)
//
// * A rows are GATHERED: up to three K-segments [x(past row) | x(current row) | h_up] of time-major bf16 activations
//   (qpnet.py:295-298, 657-666), copied with cp.async (eight threads per 128-byte row) into the canonical K-major
//   SWIZZLE_128B layout tcgen05 reads (8-row x 128-byte atoms, 16-byte pieces XOR-ed with the row index); rows outside
//   a segment are zero-filled.  cp.async completion arrives on the stage mbarrier by itself (mbarrier.arrive.noinc).
// * Weights: K-major (forward) or, for the backward GEMMs, the SAME matrices read un-transposed as an MN-major B
//   operand; either way one cp.async.bulk per stage from a pre-blocked, pre-swizzled copy (block_pack).
// * Weight gradients: both operands MN-major (a stage row is a time row); non-gathered segments arrive by TMA
//   (cp.async.bulk.tensor.3d, 64 x 64 swizzled boxes, zero fill outside the tensor), gathered ones by cp.async.
// * warps 0-7: producers, afterwards the epilogue (tcgen05.ld 32x32b, one TMEM lane = one output row per thread, row
//   tiles transposed through warp-private shared buffers so that global stores / loads / red.v4 are coalesced).
//   warp 8: TMEM allocation; its lane 0 issues the MMAs and releases slots with tcgen05.commit.
// * one output tile per CTA, two CTAs per SM (2-stage rings): the epilogue of one overlaps the mainloop of the other.
#include "qp_tc.cuh"

#include <stdlib.h>

#include <algorithm>

namespace qp {
namespace tc {

constexpr int BM = 128;      // UMMA M: time rows per tile
constexpr int BK = 64;       // bf16 elements per stage row = one 128-byte swizzle row
constexpr int STAGES = 2;       // two CTAs per SM: one runs its epilogue while the other feeds the tensor cores
constexpr int PRODUCERS = 256;   // warps 0-7: producers, then the epilogue; warp 8: TMEM allocation + MMA issue
constexpr int THREADS = 288;
constexpr int TMEM_COLS = 256;
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000LL;   // ~2 s: a lost barrier traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
// src_bytes = 0: the 16 destination bytes are zero-filled (rows outside a segment's source range)
__device__ __forceinline__ void cp_async16z(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once every cp.async this thread has issued so far has landed: the
// producer never blocks on its own loads, only on a free slot
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy (bytes % 16 == 0), completes on an mbarrier of this CTA
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO = 1 in [16,30) (unused for swizzled K-major), SBO = 1024 B >> 4
// in [32,46) (stride between 8-row atoms), version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major SWIZZLE_128B descriptor (cute: Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): a tile
// is stored as 64-element (128-byte) MN rows, one per K index, eight K rows to a 1024-byte swizzle atom; SBO = bytes
// between consecutive 8-row K groups, LBO = bytes between consecutive 64-wide MN blocks.
constexpr uint32_t MN_SBO = 1024, MN_LBO = 8192;   // a 64 (MN) x 64 (K) block is 8 consecutive atoms
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of the 16-byte piece c (8 MN elements) of K row k inside a 64 x 64 MN-major block
__device__ __forceinline__ uint32_t mn_piece(int k, int c) {
  return (uint32_t)((k >> 3) * 1024 + (k & 7) * 128 + ((c ^ (k & 7)) << 4));
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16; bit 15 / 16: A / B operand MN-major.
__device__ __forceinline__ uint32_t umma_idesc(int M, int N, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void red_add4(float* p, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
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

// one MUFU op per non-linearity (tanh.approx, abs error ~5e-4: below the bf16 rounding of the z operand it feeds)
__device__ __forceinline__ float ftanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fsigmoid(float x) { return fmaf(0.5f, ftanh(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
}

// ---- coalesced epilogue I/O ------------------------------------------------------------------------------------
// In the epilogue a thread owns one output row (its TMEM lane).  Storing row-per-thread makes every 16-byte store of
// a warp hit a different 128-byte line; instead the warp's 32 x NB-byte tile goes through a warp-private shared buffer
// (pitch NB + 16: conflict-free both ways) and is moved with the lanes laid along the rows: NB / 16 lanes per row, so
// an instruction touches 32 * 16 / NB rows with full sectors.  Rows [lo, hi) of the warp's 32 exist in memory.
constexpr int EPI_BUF = 32 * (128 + 16);   // bytes per warp
template <int NB>
__device__ __forceinline__ void warp_store_rows(unsigned char* buf, const uint4 (&v)[NB / 16], unsigned char* g0,
                                                long long pitch, int lo, int hi) {
  const int lane = threadIdx.x & 31;
  constexpr int P = NB / 16;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < P; ++i) *(uint4*)(buf + lane * (NB + 16) + 16 * i) = v[i];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int e = j * 32 + lane, row = e / P, piece = e % P;
    if (row >= lo && row < hi) *(uint4*)(g0 + row * pitch + 16 * piece) = *(const uint4*)(buf + row * (NB + 16) + 16 * piece);
  }
}
template <int NB>
__device__ __forceinline__ void warp_load_rows(unsigned char* buf, uint4 (&v)[NB / 16], const unsigned char* g0,
                                               long long pitch, int lo, int hi) {
  const int lane = threadIdx.x & 31;
  constexpr int P = NB / 16;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int e = j * 32 + lane, row = e / P, piece = e % P;
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (row >= lo && row < hi) t = *(const uint4*)(g0 + row * pitch + 16 * piece);
    *(uint4*)(buf + row * (NB + 16) + 16 * piece) = t;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < P; ++i) v[i] = *(const uint4*)(buf + lane * (NB + 16) + 16 * i);
}
// fp32 red.add of the warp's 32 x 128-byte tile into rows g0 + rowoff(row) (rowoff < 0: skip), lanes laid along the rows
__device__ __forceinline__ void warp_red_rows(unsigned char* buf, const uint4 (&v)[8], float* g0, long long my_rowoff,
                                              int npieces = 8) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) *(uint4*)(buf + lane * 144 + 16 * i) = v[i];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int row = j * 4 + (lane >> 3), piece = lane & 7;
    const long long off = __shfl_sync(0xffffffffu, my_rowoff, row);
    if (off >= 0 && piece < npieces) {
      const float4 t = *(const float4*)(buf + row * 144 + 16 * piece);
      red_add4(g0 + off + 4 * piece, t.x, t.y, t.z, t.w);
    }
  }
}
__device__ __forceinline__ uint4 f4_as_u4(float x, float y, float z, float w) {
  return make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(z), __float_as_uint(w));
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 2) tc_gemm_kernel(Args a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;              // SWIZZLE_128B atoms need 1024-byte alignment
  unsigned char* sm = smem_raw + (sbase - raw);
  const int stage_bytes = BM * 128 + a.BN * 128;
  const uint32_t bar0 = sbase + STAGES * stage_bytes;         // full[STAGES], empty[STAGES], accum : 8 bytes each
  uint32_t* tmem_slot = (uint32_t*)(sm + STAGES * stage_bytes + (2 * STAGES + 1) * 8);
  float* sbias = (float*)(sm + STAGES * stage_bytes + 128);   // this tile's 256 bias values
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, r0 = blockIdx.x * BM, n0 = a.n_begin + blockIdx.y * a.BN;
  const int ncols = min(a.BN, a.N - n0);
  if (tid < 256) sbias[tid] = (a.bias && tid < ncols) ? a.bias[n0 + tid] : 0.f;
  int nk = 0;
  for (int s = 0; s < a.nseg; ++s) nk += a.seg[s].K / BK;

  if (tid == 0) {
    // full: one pre-counted arrival per producer thread (+ the expect_tx arrival of the weight bulk copy)
    for (int i = 0; i < STAGES; ++i) { mbar_init(bar0 + 8 * i, PRODUCERS + (a.Wb ? 1 : 0)); mbar_init(bar0 + 8 * (STAGES + i), 1); }
    mbar_init(bar0 + 8 * 2 * STAGES, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == PRODUCERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < PRODUCERS / 32) {
    // ================================================================== producer
    // thread -> 16-byte piece c of tile rows rg + 32 j: eight threads copy one 128-byte row (one cache line per
    // quarter warp), and since 32 j is a multiple of 8 every row of a thread has the same swizzle phase: one smem
    // offset + j * 4096 serves the A rows, the K-major weight rows and the K rows of the MN-major weight blocks alike.
    const int c = tid & 7, rg = tid >> 3;
    const uint32_t off0 = (uint32_t)((rg >> 3) * 1024 + (rg & 7) * 128 + ((c ^ (rg & 7)) << 4));
    const int wblk = ncols >> 6;                 // MN-major weights: 64-column blocks in this tile
    int kc = 0, koff = 0;
    for (int s = 0; s < a.nseg; ++s) {
      const Seg sg = a.seg[s];
      const __nv_bfloat16* ap[4];
      uint32_t abytes[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = min(r0 + rg + 32 * j, a.n_rows - 1);
        const int src = sg.rowmap ? sg.rowmap[(long long)b * a.n_rows + rr] : rr + sg.row_off;
        const bool ok = sg.base != nullptr && src >= 0 && src < sg.src_rows;          // else the row reads zero
        abytes[j] = ok ? 16u : 0u;
        ap[j] = ok ? sg.base + (long long)b * sg.bstride + (long long)src * sg.ld + c * 8 : a.W;
      }
      // K-major weights: rows n0 + rg + 32 j (j < ncols / 32);  MN-major: K rows koff + k0 + rg + 32 g, blocks q
      const __nv_bfloat16* wp = a.w_mn ? a.W + (long long)(koff + rg) * a.ldw + n0 + c * 8
                                       : a.W + (long long)(n0 + rg) * a.ldw + koff + c * 8;
      const long long wstep = 32LL * a.ldw;
      for (int k0 = 0; k0 < sg.K; k0 += BK, ++kc) {
        const int stage = kc % STAGES, it = kc / STAGES;
        if (it > 0) mbar_wait(bar0 + 8 * (STAGES + stage), (uint32_t)(it - 1) & 1u);
        const uint32_t sA = sbase + stage * stage_bytes + off0;
        const uint32_t sB = sA + BM * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) cp_async16z(sA + j * 4096, abytes[j] ? ap[j] + k0 : ap[j], abytes[j]);
        if (a.Wb) {
          if (tid == 0) {      // the whole weight tile of this stage: ncols / 64 consecutive 8 KB blocks
            const uint32_t bytes = (uint32_t)(ncols >> 6) * MN_LBO;
            mbar_expect_tx(bar0 + 8 * stage, bytes);
            bulk_g2s(sbase + stage * stage_bytes + BM * 128,
                     a.Wb + ((long long)(a.wb_k0 + ((koff + k0) >> 6)) * a.wb_pitch + (n0 >> 6)) * (MN_LBO / 2), bytes, bar0 + 8 * stage);
          }
        } else if (a.w_mn) {
          const __nv_bfloat16* wk = wp + (long long)k0 * a.ldw;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < wblk) {
              cp_async16(sB + q * MN_LBO, wk + q * 64);
              cp_async16(sB + q * MN_LBO + 4096, wk + wstep + q * 64);
            }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (rg + 32 * j < ncols) cp_async16(sB + j * 4096, wp + j * wstep + k0);
        }
        cp_async_arrive_noinc(bar0 + 8 * stage);
      }
      koff += sg.K;
    }

    // ================================================================== epilogue
    mbar_wait(bar0 + 8 * 2 * STAGES, 0);
    tc_fence_after();
    // a warp reads the TMEM lanes 32 (warp % 4) ..: warps w and w + 4 share their rows and take alternate column groups.
    // Every load has landed and every MMA has retired: the stage ring is free and holds the warps' transpose buffers.
    unsigned char* buf = sm + warp * EPI_BUF;
    const int rw = r0 + 32 * (warp & 3);              // first row of this warp
    const int hi = min(32, a.n_rows - rw);            // rows [0, hi) of the warp exist
    const long long row0 = (long long)b * a.n_rows + rw;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int g = 32 * (warp >> 2); g < ncols; g += 64) {
      uint32_t v[32];
      tmem_ld32(trow + g, v);
      const int n = n0 + g;
      if (EPI == EPI_GATE) {
        // columns (2c, 2c+1) = (sigmoid, tanh) pre-activations of channel c  (qpnet.py:665-666 / 634-635)
        float z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float sg_ = fsigmoid(__uint_as_float(v[2 * j]) + sbias[g + 2 * j]);
          float th_ = ftanh(__uint_as_float(v[2 * j + 1]) + sbias[g + 2 * j + 1]);
          z[j] = sg_ * th_;
          v[2 * j] = __float_as_uint(sg_);
          v[2 * j + 1] = __float_as_uint(th_);
        }
        const int halfN = a.N >> 1, c0 = n >> 1;
        {
          uint4 o[2] = {make_uint4(pack_bf16(z[0], z[1]), pack_bf16(z[2], z[3]), pack_bf16(z[4], z[5]), pack_bf16(z[6], z[7])),
                        make_uint4(pack_bf16(z[8], z[9]), pack_bf16(z[10], z[11]), pack_bf16(z[12], z[13]), pack_bf16(z[14], z[15]))};
          warp_store_rows<32>(buf, o, (unsigned char*)(a.z_bf + row0 * halfN + c0), 2LL * halfN, 0, hi);
        }
        if (a.z_f32) {
          uint4 o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = f4_as_u4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
          warp_store_rows<64>(buf, o, (unsigned char*)(a.z_f32 + row0 * halfN + c0), 4LL * halfN, 0, hi);
        }
        if (a.gsave_bf) {
          uint4 o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = make_uint4(pack_bf16(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                              pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                              pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                              pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
          warp_store_rows<64>(buf, o, (unsigned char*)(a.gsave_bf + row0 * a.N + n), 2LL * a.N, 0, hi);
        }
        if (a.gsave) {
          uint4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          warp_store_rows<128>(buf, o, (unsigned char*)(a.gsave + row0 * a.N + n), 4LL * a.N, 0, hi);
        }
      } else if (EPI == EPI_RESSKIP) {
        if (n < a.C) {
          // residual projection + x(current row), fp32 stream + bf16 operand copy (qpnet.py:668-669)
          uint4 x[8];
          warp_load_rows<128>(buf, x, (const unsigned char*)(a.xcur + (long long)b * a.xcur_bstride + (long long)(rw + a.xcur_off) * a.C + n),
                              4LL * a.C, 0, hi);
          float o[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[4 * j] = __uint_as_float(v[4 * j]) + sbias[g + 4 * j] + __uint_as_float(x[j].x);
            o[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + sbias[g + 4 * j + 1] + __uint_as_float(x[j].y);
            o[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + sbias[g + 4 * j + 2] + __uint_as_float(x[j].z);
            o[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + sbias[g + 4 * j + 3] + __uint_as_float(x[j].w);
            x[j] = f4_as_u4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
          warp_store_rows<128>(buf, x, (unsigned char*)(a.xnext + row0 * a.C + n), 4LL * a.C, 0, hi);
          uint4 ob[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ob[j] = make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                               pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
          warp_store_rows<64>(buf, ob, (unsigned char*)(a.xnext_bf + row0 * a.C + n), 2LL * a.C, 0, hi);
        } else {
          // skip projection, accumulated over the blocks for the last bl rows only (qpnet.py:667, 283)
          const int lo = max(0, a.skip_row0 - rw);
          if (lo < hi) {
            unsigned char* sk = (unsigned char*)(a.skip + (long long)b * a.skip_bstride + (long long)(rw - a.skip_row0) * a.S + (n - a.C));
            uint4 p[8];
            if (a.skip_accum) warp_load_rows<128>(buf, p, sk, 4LL * a.S, lo, hi);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = make_float4(__uint_as_float(v[4 * j]) + sbias[g + 4 * j],
                                     __uint_as_float(v[4 * j + 1]) + sbias[g + 4 * j + 1],
                                     __uint_as_float(v[4 * j + 2]) + sbias[g + 4 * j + 2],
                                     __uint_as_float(v[4 * j + 3]) + sbias[g + 4 * j + 3]);
              if (a.skip_accum) { o.x += __uint_as_float(p[j].x); o.y += __uint_as_float(p[j].y); o.z += __uint_as_float(p[j].z); o.w += __uint_as_float(p[j].w); }
              p[j] = f4_as_u4(o.x, o.y, o.z, o.w);
            }
            warp_store_rows<128>(buf, p, sk, 4LL * a.S, lo, hi);
          }
        }
      } else if (EPI == EPI_DGATE) {
        // acc = dz[r][c], c = n .. n + 31 -> dgate columns 2c, 2c + 1 through the saved sigmoid / tanh values
        uint4 g0[8], g1[8];
        if (a.gin_bf) {
          // 64 bf16 values per row: widen into the fp32 layout the arithmetic below reads
          uint4 gb[8];
          warp_load_rows<128>(buf, gb, (const unsigned char*)(a.gin_bf + row0 * (2 * a.N) + 2 * n), 4LL * a.N, 0, hi);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 lo = make_uint4(gb[j].x << 16, gb[j].x & 0xFFFF0000u, gb[j].y << 16, gb[j].y & 0xFFFF0000u);
            const uint4 hi4 = make_uint4(gb[j].z << 16, gb[j].z & 0xFFFF0000u, gb[j].w << 16, gb[j].w & 0xFFFF0000u);
            if (j < 4) { g0[2 * j] = lo; g0[2 * j + 1] = hi4; } else { g1[2 * (j - 4)] = lo; g1[2 * (j - 4) + 1] = hi4; }
          }
        } else {
          const unsigned char* gp = (const unsigned char*)(a.gin + row0 * (2 * a.N) + 2 * n);
          warp_load_rows<128>(buf, g0, gp, 8LL * a.N, 0, hi);          // (sg, th) of channels n .. n + 15
          warp_load_rows<128>(buf, g1, gp + 128, 8LL * a.N, 0, hi);    // channels n + 16 .. n + 31
        }
        float o[64];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint4 gv = j < 8 ? g0[j] : g1[j - 8];   // (sg, th) of channels n + 2j, n + 2j + 1
          const float sg0 = __uint_as_float(gv.x), th0 = __uint_as_float(gv.y), sg1 = __uint_as_float(gv.z), th1 = __uint_as_float(gv.w);
          const float dz0 = __uint_as_float(v[2 * j]), dz1 = __uint_as_float(v[2 * j + 1]);
          o[4 * j] = dz0 * th0 * sg0 * (1.f - sg0);
          o[4 * j + 1] = dz0 * sg0 * (1.f - th0 * th0);
          o[4 * j + 2] = dz1 * th1 * sg1 * (1.f - sg1);
          o[4 * j + 3] = dz1 * sg1 * (1.f - th1 * th1);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          g0[j] = make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                             pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
        warp_store_rows<128>(buf, g0, (unsigned char*)(a.dgate_bf + row0 * (2 * a.N) + 2 * n), 4LL * a.N, 0, hi);
        if (a.dgate_f32) {
          unsigned char* of = (unsigned char*)(a.dgate_f32 + row0 * (2 * a.N) + 2 * n);
#pragma unroll
          for (int j = 0; j < 8; ++j) g0[j] = f4_as_u4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          warp_store_rows<128>(buf, g0, of, 8LL * a.N, 0, hi);
#pragma unroll
          for (int j = 0; j < 8; ++j) g0[j] = f4_as_u4(o[32 + 4 * j], o[32 + 4 * j + 1], o[32 + 4 * j + 2], o[32 + 4 * j + 3]);
          warp_store_rows<128>(buf, g0, of + 128, 8LL * a.N, 0, hi);
        }
      } else if (EPI == EPI_DX) {
        // acc = (dgate * Wg)[r][k], k = n .. n + 31 inside one of [past C | current C | aux]
        const int r = rw + lane;
        const bool rv = r < a.n_rows;
        if (n < 2 * a.C) {
          uint4 o[8];
          long long off = -1;
          if (n < a.C) {
            const int src = rv ? (a.dx_rowmap ? a.dx_rowmap[(long long)b * a.n_rows + r] : r + a.dx_past_off) : -1;
            if (src >= 0 && src < a.dx_rows) off = (long long)src * a.C + n;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            const int c = n - a.C;
            if (rv) off = (long long)(r + a.dx_cur_off) * a.C + c;
            if (a.resid) {
              warp_load_rows<128>(buf, o, (const unsigned char*)(a.resid + (long long)b * a.resid_bstride + (long long)rw * a.C + c),
                                  4LL * a.C, 0, hi);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                o[j] = f4_as_u4(__uint_as_float(v[4 * j]) + __uint_as_float(o[j].x), __uint_as_float(v[4 * j + 1]) + __uint_as_float(o[j].y),
                                __uint_as_float(v[4 * j + 2]) + __uint_as_float(o[j].z), __uint_as_float(v[4 * j + 3]) + __uint_as_float(o[j].w));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
          warp_red_rows(buf, o, a.dx + (long long)b * a.dx_bstride, off);
        } else if (rv) {
          const int ka = n - 2 * a.C;
          float* q = a.dh + (long long)b * a.dh_bstride + (long long)(r + a.dh_off) * a.A;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ka + j < a.A) q[ka + j] += __uint_as_float(v[j]);
        }
      } else {  // EPI_HEAD: out = acc + bias (fp32, pre-activation) and optionally relu(out) as the next bf16 operand
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) + sbias[g + j];
        uint4 of[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) of[j] = f4_as_u4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        warp_store_rows<128>(buf, of, (unsigned char*)(a.out + (long long)b * a.out_bstride + (long long)rw * a.ldo + n), 4LL * a.ldo, 0, hi);
        if (a.out_relu_bf) {
          uint4 ob[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ob[j] = make_uint4(pack_bf16(fmaxf(o[8 * j], 0.f), fmaxf(o[8 * j + 1], 0.f)),
                               pack_bf16(fmaxf(o[8 * j + 2], 0.f), fmaxf(o[8 * j + 3], 0.f)),
                               pack_bf16(fmaxf(o[8 * j + 4], 0.f), fmaxf(o[8 * j + 5], 0.f)),
                               pack_bf16(fmaxf(o[8 * j + 6], 0.f), fmaxf(o[8 * j + 7], 0.f)));
          warp_store_rows<64>(buf, ob, (unsigned char*)(a.out_relu_bf + (long long)b * a.out_bstride + (long long)rw * a.ldo + n), 2LL * a.ldo, 0, hi);
        }
      }
    }
  } else {
    // ================================================================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, ncols, 0, a.w_mn);
      const uint32_t bstep = a.w_mn ? (2 * MN_SBO) >> 4 : 2;   // MN-major: K = 16 is two 8-row groups; K-major: 32 bytes
      for (int kc = 0; kc < nk; ++kc) {
        const int stage = kc % STAGES, it = kc / STAGES;
        mbar_wait(bar0 + 8 * stage, (uint32_t)it & 1u);
        fence_proxy_async();     // the producers' cp.async writes (generic proxy) -> tcgen05 operand reads (async proxy)
        tc_fence_after();
        const uint64_t ad = umma_desc(sbase + stage * stage_bytes);
        const uint32_t sb = sbase + stage * stage_bytes + BM * 128;
        const uint64_t bd = a.w_mn ? umma_desc_mn(sb, a.mn_lbo, a.mn_sbo) : umma_desc(sb);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)          // +32 bytes per K = 16 step inside the 128-byte swizzle row
          umma(tmem, ad + 2 * k, bd + bstep * k, idesc, (uint32_t)(kc | k));
        umma_commit(bar0 + 8 * (STAGES + stage));  // slot free once these MMAs retire
      }
      umma_commit(bar0 + 8 * 2 * STAGES);           // accumulator complete
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PRODUCERS / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// QPNET_MN_SWAP=1 exchanges the two MN-major descriptor strides (bring-up aid; the default is the cute convention)
static void mn_strides(uint32_t* lbo, uint32_t* sbo) {
  static int swap = -1;
  if (swap < 0) { const char* e = getenv("QPNET_MN_SWAP"); swap = (e && e[0] == '1') ? 1 : 0; }
  *lbo = swap ? MN_SBO : MN_LBO;
  *sbo = swap ? MN_LBO : MN_SBO;
}

template <int EPI>
static int launch(const Args& a0, cudaStream_t st) {
  Args a = a0;
  mn_strides(&a.mn_lbo, &a.mn_sbo);
  if (a.n_rows <= 0 || a.B <= 0 || a.N - a.n_begin <= 0) return QP_OK;
  int ktot = 0;
  for (int s = 0; s < a.nseg; ++s) {
    QP_REQUIRE(a.seg[s].K > 0 && a.seg[s].K % BK == 0, "tc gemm: K segment %d not a multiple of %d", a.seg[s].K, BK);
    ktot += a.seg[s].K;
  }
  QP_REQUIRE(a.BN >= 32 && a.BN <= 256 && a.BN % 32 == 0 && (a.N - a.n_begin) % 32 == 0 && (a.w_mn || a.ldw >= ktot),
             "tc gemm: bad tile shape N=%d BN=%d", a.N, a.BN);
  if (a.Wb)
    QP_REQUIRE(a.BN % 64 == 0 && (a.N - a.n_begin) % 64 == 0 && a.n_begin % 64 == 0 && (((size_t)a.Wb) & 15) == 0,
               "tc gemm: blocked weights need 64-column tiles (N=%d BN=%d n_begin=%d)", a.N, a.BN, a.n_begin);
  if (a.w_mn)
    QP_REQUIRE(a.BN % 64 == 0 && (a.N - a.n_begin) % 64 == 0 && a.ldw % 8 == 0 && a.n_begin % 8 == 0,
               "tc gemm: MN-major weights need 64-column blocks (N=%d BN=%d ldw=%d)", a.N, a.BN, a.ldw);
  const size_t smem = (size_t)STAGES * (BM * 128 + a.BN * 128) + 128 + 1024 + 1024;   // ring, barriers, bias, alignment
  static bool configured[5] = {false, false, false, false, false};
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
int gemm_dgate(const Args& a, cudaStream_t st) {
  QP_REQUIRE(a.w_mn && (a.gin || a.gin_bf) && a.dgate_bf, "tc dgate gemm: missing operands");
  return launch<EPI_DGATE>(a, st);
}
int gemm_dx(const Args& a, cudaStream_t st) {
  QP_REQUIRE(a.w_mn && a.dx && a.dh && a.C % 32 == 0, "tc dx gemm: missing operands");
  return launch<EPI_DX>(a, st);
}

// ------------------------------------------------------------------ small bf16 helpers
__global__ void f32_to_bf16_pad_kernel(const float* __restrict__ src, long long rows, int K, int Kp,
                                       __nv_bfloat16* __restrict__ dst, int relu, int ones_col) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Kp) return;
  long long r = i / Kp;
  int k = (int)(i % Kp);
  float v = k < K ? src[r * K + k] : (k == ones_col ? 1.f : 0.f);
  if (relu) v = fmaxf(v, 0.f);
  dst[i] = __float2bfloat16_rn(v);
}

int f32_to_bf16_pad(const float* src, long long rows, int K, int Kp, __nv_bfloat16* dst, int relu, cudaStream_t st,
                    int ones_col) {
  long long n = rows * Kp;
  if (n <= 0) return QP_OK;
  f32_to_bf16_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, rows, K, Kp, dst, relu, ones_col);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

// one thread per 16-byte piece (8 consecutive columns of one row)
__global__ void block_pack_kernel(const __nv_bfloat16* __restrict__ src, long long pieces, int R, int Cc, int col_outer,
                                  __nv_bfloat16* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pieces) return;
  const int pc = Cc >> 3;                       // pieces per row
  const long long per_mat = (long long)R * pc;
  const long long m = i / per_mat;
  const int rem = (int)(i - m * per_mat), row = rem / pc, piece = rem - row * pc;
  const int rb = row >> 6, rr = row & 63, cb = piece >> 3, cp = piece & 7;
  const long long blk = col_outer ? (long long)cb * (R >> 6) + rb : (long long)rb * (Cc >> 6) + cb;
  const long long off = m * (long long)R * Cc + blk * 4096 + rr * 64 + ((cp ^ (rr & 7)) << 3);   // elements
  *(uint4*)(dst + off) = *(const uint4*)(src + m * (long long)R * Cc + (long long)row * Cc + piece * 8);
}

int block_pack(const __nv_bfloat16* src, int mats, int R, int Cc, int col_outer, __nv_bfloat16* dst, cudaStream_t st) {
  QP_REQUIRE(R % 64 == 0 && Cc % 64 == 0 && mats > 0, "block_pack: %d x %d is not made of 64 x 64 blocks", R, Cc);
  const long long pieces = (long long)mats * R * (Cc >> 3);
  block_pack_kernel<<<(unsigned)((pieces + 255) / 256), 256, 0, st>>>(src, pieces, R, Cc, col_outer, dst);
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

// ------------------------------------------------------------------ fp32 -> bf16 copy with column sums
// block: 64 rows; thread -> column quad tid % (K / 4), row lane tid / (K / 4)
constexpr int CVT_ROWS = 64;
__global__ void __launch_bounds__(256) f32_to_bf16_colsum_kernel(const float* __restrict__ src, long long rows, int K,
                                                                 __nv_bfloat16* __restrict__ dst, float* __restrict__ colsum) {
  __shared__ float4 red[256];
  const int kq = K >> 2, lanes = 256 / kq, tid = threadIdx.x;
  const int cq = tid % kq, rl = tid / kq;
  float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long r0 = (long long)blockIdx.x * CVT_ROWS, r1 = min(rows, r0 + CVT_ROWS);
  if (rl < lanes) {
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const float4 v = *(const float4*)(src + r * K + 4 * cq);
      if (dst) *(uint2*)(dst + r * K + 4 * cq) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
      s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
    }
  }
  if (!colsum) return;
  red[tid] = s4;
  __syncthreads();
  if (rl == 0) {
    for (int l = 1; l < lanes; ++l) { const float4 o = red[l * kq + cq]; s4.x += o.x; s4.y += o.y; s4.z += o.z; s4.w += o.w; }
    red_add4(colsum + 4 * cq, s4.x, s4.y, s4.z, s4.w);
  }
}

int f32_to_bf16_colsum(const float* src, long long rows, int K, __nv_bfloat16* dst, float* colsum, cudaStream_t st) {
  if (rows <= 0) return QP_OK;
  QP_REQUIRE(K > 0 && K % 4 == 0 && K <= 1024, "f32_to_bf16_colsum: K=%d", K);
  f32_to_bf16_colsum_kernel<<<(unsigned)((rows + CVT_ROWS - 1) / CVT_ROWS), 256, 0, st>>>(src, rows, K, dst, colsum);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

// ------------------------------------------------------------------ weight gradients on tcgen05
// out[128 i x BJ j] (fp32, TMEM) += P[64 rows x 128 i]^T * Q[64 rows x BJ j] per stage, both operands MN-major: a stage row
// is one time row r, so the gathered past-tap rows of Q are again plain row copies (cp.async, 128 bytes per 64-column
// block).  warps 0-3: producers, then the epilogue (TMEM lane = output row i, fp32 red.v4 into the zeroed output);
// warp 4: TMEM allocation + the MMA-issuing thread.  grid = (i tiles, j tiles, row splits).
constexpr int WG_STAGES = 2;
constexpr int WG_MAXBLK = 6;   // 64-column blocks per stage: 2 of P (128 output rows) + up to 4 of Q (256 output columns)
struct WBlk {
  const __nv_bfloat16* base;   // segment base + first column of the block; nullptr: all-zero block
  const int* rowmap;
  long long bstride;
  int ld, row_off, src_rows;
  int tma;                     // index of the segment's TMA descriptor, -1: copied with cp.async
  int col;                     // first column of the block inside its segment
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

__device__ __forceinline__ WBlk wblk_resolve(const Seg* segs, int nseg, int col, const int* use_tma, int tm0) {
  WBlk w; w.base = nullptr; w.rowmap = nullptr; w.bstride = 0; w.ld = 0; w.row_off = 0; w.src_rows = 0; w.tma = -1; w.col = 0;
  int off = 0;
  for (int s = 0; s < nseg; ++s) {
    if (col - off < segs[s].K) {
      if (segs[s].base) {
        w.base = segs[s].base + (col - off); w.rowmap = segs[s].rowmap; w.bstride = segs[s].bstride; w.ld = segs[s].ld;
        w.row_off = segs[s].row_off; w.src_rows = segs[s].src_rows;
        w.col = col - off;
        if (use_tma[tm0 + s]) w.tma = tm0 + s;
      }
      return w;
    }
    off += segs[s].K;
  }
  return w;
}

__global__ void __launch_bounds__(THREADS, 2) tc_wgrad_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ WBlk sblk[WG_MAXBLK];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (sbase - raw);
  const int stage_bytes = (2 + (a.BJ >> 6)) * (int)MN_LBO;
  const uint32_t bar0 = sbase + WG_STAGES * stage_bytes;
  uint32_t* tmem_slot = (uint32_t*)(sm + WG_STAGES * stage_bytes + (2 * WG_STAGES + 1) * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.x * BM, j0 = blockIdx.y * a.BJ;
  int jp = 0;
  for (int s = 0; s < a.nq; ++s) jp += a.q[s].K;
  const int ncols = min(a.BJ, jp - j0);            // multiple of 64
  const int nblk = 2 + (ncols >> 6);
  const int row_begin = blockIdx.z * a.chunk, row_end = min(a.n_rows, row_begin + a.chunk);
  const int steps_b = (row_end - row_begin + BK - 1) / BK;
  const int nk = a.B * steps_b;

  if (tid == 0) {
    // full: one pre-counted arrival per producer thread + the expect_tx arrival that announces the TMA blocks
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(bar0 + 8 * i, PRODUCERS + 1); mbar_init(bar0 + 8 * (WG_STAGES + i), 1); }
    mbar_init(bar0 + 8 * 2 * WG_STAGES, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (tid < nblk)
    sblk[tid] = tid < 2 ? wblk_resolve(a.p, a.np, i0 + 64 * tid, a.use_tma, 0) : wblk_resolve(a.q, a.nq, j0 + 64 * (tid - 2), a.use_tma, 2);
  if (warp == PRODUCERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < PRODUCERS / 32) {
    // ================================================================== producer
    // thread -> 16-byte piece c (8 columns) of stage rows rg and rg + 32 of every 64-column block; the block table
    // lives in registers (fully unrolled), one smem offset + g * 4096 + q * 8192 addresses every copy
    const int c = tid & 7, rg = tid >> 3;
    const uint32_t off0 = mn_piece(rg, c);
    // Per block: source row = r + off (or the gathered row), valid iff (unsigned)src < hi.  hi folds the row range of the
    // segment AND the end of this CTA's row chunk into one compare; an all-zero block has hi = 0.  Invalid rows copy
    // nothing (zero fill) from row 0 of the block, so the address is always a real one.
    const __nv_bfloat16* dummy = a.q[0].base;
    int off[WG_MAXBLK], hi[WG_MAXBLK], ld[WG_MAXBLK];
    bool gather = false, cpa[WG_MAXBLK];     // cpa: the block is copied by the threads (gathered or all-zero), else by TMA
    int ntma = 0;
#pragma unroll
    for (int q = 0; q < WG_MAXBLK; ++q) {
      const WBlk w = sblk[q < nblk ? q : 0];
      off[q] = w.row_off; ld[q] = w.base ? w.ld : 0;
      hi[q] = w.base ? (w.rowmap ? w.src_rows : max(0, min(w.src_rows, row_end + w.row_off))) : 0;
      cpa[q] = (q < nblk) && w.tma < 0;
      ntma += (q < nblk) && w.tma >= 0;
      gather |= (q < nblk) && w.base && w.rowmap;
    }
    int kc = 0;
    for (int b = 0; b < a.B; ++b) {
      // per batch element: block bases with this thread's column piece folded in; offsets inside one element fit 32 bits
      const __nv_bfloat16* pb[WG_MAXBLK];
      const int* rm[WG_MAXBLK];
#pragma unroll
      for (int q = 0; q < WG_MAXBLK; ++q) {
        const WBlk w = sblk[q < nblk ? q : 0];
        pb[q] = w.base ? w.base + (long long)b * w.bstride + c * 8 : dummy;
        rm[q] = (w.base && w.rowmap) ? w.rowmap + (long long)b * a.n_rows : nullptr;
      }
      for (int r0 = row_begin; r0 < row_end; r0 += BK, ++kc) {
        const int stage = kc % WG_STAGES, it = kc / WG_STAGES;
        if (it > 0) mbar_wait(bar0 + 8 * (WG_STAGES + stage), (uint32_t)(it - 1) & 1u);
        const uint32_t st0 = sbase + stage * stage_bytes + off0;
        if (tid == 0) {     // the non-gathered blocks: one TMA box each (64 rows x 64 columns, swizzled, zero-filled outside)
          mbar_expect_tx(bar0 + 8 * stage, (uint32_t)ntma * MN_LBO);
#pragma unroll
          for (int q = 0; q < WG_MAXBLK; ++q) {
            if (q < nblk && sblk[q].tma >= 0)
              tma_load_3d(sbase + stage * stage_bytes + q * MN_LBO, &a.tm[sblk[q].tma], sblk[q].col, r0 + sblk[q].row_off, b,
                          bar0 + 8 * stage);
          }
        }
        if (!gather) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int r = r0 + rg + 32 * g;
#pragma unroll
            for (int q = 0; q < WG_MAXBLK; ++q) {
              if (cpa[q]) {
                const int src = r + off[q];
                const bool ok = (unsigned)src < (unsigned)hi[q];
                cp_async16z(st0 + g * 4096 + q * MN_LBO, pb[q] + (unsigned)((ok ? src : 0) * ld[q]), ok ? 16u : 0u);
              }
            }
          }
        } else {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int r = r0 + rg + 32 * g;
            const bool rin = r < row_end;
#pragma unroll
            for (int q = 0; q < WG_MAXBLK; ++q) {
              if (cpa[q]) {
                int src = r + off[q];
                if (rm[q]) src = rin ? rm[q][r] : -1;
                const bool ok = (unsigned)src < (unsigned)hi[q];
                cp_async16z(st0 + g * 4096 + q * MN_LBO, pb[q] + (unsigned)((ok ? src : 0) * ld[q]), ok ? 16u : 0u);
              }
            }
          }
        }
        cp_async_arrive_noinc(bar0 + 8 * stage);
      }
    }

    // ================================================================== epilogue
    if (nk > 0) {
      mbar_wait(bar0 + 8 * 2 * WG_STAGES, 0);
      tc_fence_after();
      const int i = i0 + 32 * (warp & 3) + lane;     // warps w and w + 4 share their TMEM lanes, alternate column groups
      const bool iv = i < a.I;
      unsigned char* buf = sm + warp * EPI_BUF;      // the stage ring is free: every load landed, every MMA retired
      const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      for (int g = 32 * (warp >> 2); g < ncols; g += 64) {
        uint32_t v[32];
        tmem_ld32(trow + g, v);
        const int j = j0 + g;
        uint4 o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        warp_red_rows(buf, o, a.out, iv ? (long long)i * a.ldo + j : -1LL, min(8, (a.J - j) >> 2));
        if (iv && a.ones_out && a.ones_col >= j && a.ones_col < j + 32) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (j + e == a.ones_col) atomicAdd(a.ones_out + i, __uint_as_float(v[e]));
        }
      }
    }
  } else {
    // ================================================================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, ncols, 1, 1);
      for (int kc = 0; kc < nk; ++kc) {
        const int stage = kc % WG_STAGES, it = kc / WG_STAGES;
        mbar_wait(bar0 + 8 * stage, (uint32_t)it & 1u);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t st0 = sbase + stage * stage_bytes;
        const uint64_t ad = umma_desc_mn(st0, a.mn_lbo, a.mn_sbo);
        const uint64_t bd = umma_desc_mn(st0 + 2 * MN_LBO, a.mn_lbo, a.mn_sbo);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)          // K = 16 per instruction = two 8-row groups of the MN-major blocks
          umma(tmem, ad + ((2 * MN_SBO) >> 4) * k, bd + ((2 * MN_SBO) >> 4) * k, idesc, (uint32_t)(kc | k));
        umma_commit(bar0 + 8 * (WG_STAGES + stage));
      }
      if (nk > 0) umma_commit(bar0 + 8 * 2 * WG_STAGES);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PRODUCERS / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    const char* e = getenv("QPNET_WGRAD_TMA");
    if (e && e[0] == '0') fn = nullptr;
  }
  return fn;
}
// bf16 [B][rows][K] view of a segment (row pitch ld, batch pitch bstride), 64 x 64 boxes, 128-byte swizzle, zero fill
static bool encode_segment(EncodeTiledFn enc, const Seg& sg, int B, int rows, CUtensorMap* tm) {
  if (!enc || !sg.base || sg.rowmap || rows <= 0 || (((size_t)sg.base) & 15) || (sg.ld & 7) || (sg.bstride & 7)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)sg.K, (cuuint64_t)rows, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)sg.ld * 2, (cuuint64_t)(B > 1 ? sg.bstride : (long long)rows * sg.ld) * 2};
  const cuuint32_t box[3] = {64, 64, 1}, estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)sg.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static long long g_tma_segments = 0;
long long tma_segments_bound() { return g_tma_segments; }

int wgrad(const WgradArgs& a0, cudaStream_t st) {
  WgradArgs a = a0;
  {
    // a segment's rows beyond the last contracted row never meet a non-zero partner: clip the tensor there
    EncodeTiledFn enc = tensor_map_encoder();
    for (int s = 0; s < 5; ++s) a.use_tma[s] = 0;
    for (int s = 0; s < a.np; ++s)
      a.use_tma[s] = encode_segment(enc, a.p[s], a.B, std::min(a.p[s].src_rows, a.n_rows + a.p[s].row_off), &a.tm[s]);
    for (int s = 0; s < a.nq; ++s)
      a.use_tma[2 + s] = encode_segment(enc, a.q[s], a.B, std::min(a.q[s].src_rows, a.n_rows + a.q[s].row_off), &a.tm[2 + s]);
    for (int s = 0; s < 5; ++s) g_tma_segments += a.use_tma[s];
  }
  if (a.n_rows <= 0 || a.B <= 0 || a.I <= 0 || a.J <= 0) return QP_OK;
  int ip = 0, jp = 0;
  for (int s = 0; s < a.np; ++s) { QP_REQUIRE(a.p[s].K > 0 && a.p[s].K % 64 == 0, "tc wgrad: P segment width %d", a.p[s].K); ip += a.p[s].K; }
  for (int s = 0; s < a.nq; ++s) { QP_REQUIRE(a.q[s].K > 0 && a.q[s].K % 64 == 0 && a.q[s].base, "tc wgrad: Q segment width %d", a.q[s].K); jp += a.q[s].K; }
  QP_REQUIRE(a.I <= ip && a.J <= jp && a.ldo % 4 == 0 && a.J % 4 == 0 && (((size_t)a.out) & 15) == 0, "tc wgrad: bad output shape I=%d J=%d ldo=%d", a.I, a.J, a.ldo);
  mn_strides(&a.mn_lbo, &a.mn_sbo);
  const int jblocks = jp / 64, ntj = (jblocks + 3) / 4;
  a.BJ = 64 * ((jblocks + ntj - 1) / ntj);
  const int tiles = ((a.I + BM - 1) / BM) * ntj;
  // row splits: about two CTAs per SM in flight over the launch, at least 512 rows each
  static int waves = -1;
  if (waves < 0) { const char* e = getenv("QPNET_WGRAD_WAVES"); waves = e ? atoi(e) : 2; if (waves < 1) waves = 1; }
  int nsplit = std::max(1, std::min((148 * waves) / std::max(tiles, 1), a.n_rows / 512));
  a.chunk = ((a.n_rows + nsplit - 1) / nsplit + BK - 1) / BK * BK;
  nsplit = (a.n_rows + a.chunk - 1) / a.chunk;
  const size_t smem = (size_t)WG_STAGES * (2 + a.BJ / 64) * MN_LBO + (2 * WG_STAGES + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    QP_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // + static sblk
    configured = true;
  }
  dim3 grid((a.I + BM - 1) / BM, ntj, nsplit);
  tc_wgrad_kernel<<<grid, THREADS, smem, st>>>(a);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

}  // namespace tc
}  // namespace qp

extern "C" int64_t qp_debug_tma_segments(void) { return (int64_t)qp::tc::tma_segments_bound(); }
