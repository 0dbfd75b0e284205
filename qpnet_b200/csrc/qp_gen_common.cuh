// Device helpers shared by the persistent generators (qp_generate.cu: generic shapes,
// qp_generate_fold2.cu / qp_generate_f3.cu: cluster K-split kernels for the SI default architecture).
#pragma once
#include <cuda_fp16.h>

#include "qp_common.cuh"

namespace qp {

constexpr long long GEN_TIMEOUT_CYCLES = 6000000000LL;  // ~3 s: watchdog, not a schedule

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint4 ld_strong_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_strong_v2(const void* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_strong_u32(const void* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_strong_u32(void* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_strong_v2(void* p, unsigned a, unsigned b) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};\n" ::"l"(p), "r"(a), "r"(b) : "memory");
}

__device__ __forceinline__ void ldmatrix_x4(unsigned& a0, unsigned& a1, unsigned& a2, unsigned& a3, const void* p) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(sa));
}
__device__ __forceinline__ void ldmatrix_x2(unsigned& b0, unsigned& b1, const void* p) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(b0), "=r"(b1) : "r"(sa));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                         unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// MUFU.TANH (max relative error 2^-11, far below the bf16 rounding of z); sigmoid(x) = 0.5 tanh(x/2) + 0.5
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }

// bf16 bits of x, round-to-nearest-even
__device__ __forceinline__ unsigned bf16_rne(float x) { return (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }
// bf16 bits of x rounded to the nearest value whose mantissa LSB equals `par` (error <= 1 ulp)
__device__ __forceinline__ unsigned bf16_tagged(float x, unsigned par) {
  unsigned u = __float_as_uint(x);
  unsigned hi = u >> 16, rem = u & 0xFFFFu;
  unsigned r = hi + ((rem > 0x8000u) || (rem == 0x8000u && (hi & 1u)));
  if ((r & 1u) != par) r = (r > hi) ? r - 1u : r + 1u;
  return r & 0xFFFFu;
}
// one exchange word: two channels, epoch tag in bit 0
__device__ __forceinline__ unsigned pack_tagged(float lo, float hi, unsigned par) {
  return bf16_tagged(lo, par) | (bf16_rne(hi) << 16);
}

// Philox4x32-10, one draw per (utterance, step)
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned utt, unsigned step) {
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
  unsigned c0 = step, c1 = utt, c2 = 0x51504e45u, c3 = 0x42323030u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (float)(c0 >> 8) * (1.0f / 16777216.0f);
}

// ------------------------------------------------------------------ cluster / mbarrier helpers (qp_generate_fold2.cu, qp_generate_f3.cu)
__device__ __forceinline__ unsigned smem_u32(const void* q) { return (unsigned)__cvta_generic_to_shared(q); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void st_async_v4(unsigned dst_cluster_addr, uint4 v, unsigned mbar_cluster_addr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];\n"
               ::"r"(dst_cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ bool fresh4(uint4 v, unsigned par) {
  return (((v.x ^ par) | (v.y ^ par) | (v.z ^ par) | (v.w ^ par)) & 1u) == 0;
}
__device__ __forceinline__ unsigned pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *(unsigned*)&h;
}
__device__ __forceinline__ float2 unpack_h2(unsigned u) {
  __half2 h = *(__half2*)&u;
  return __half22float2(h);
}

struct GenArgsDev {
  const int64_t* seed; const float* h; const void* d; const int32_t* n_samples;
  const float* uniforms; long long ld_uniforms; unsigned long long philox_seed;
  const int32_t* force; long long ld_force;
  const int32_t* utt_ids;   // optional: the caller-side index of every utterance (keys the Philox stream)
  int32_t* out; long long ld_out; float* logits_out;
  int16_t* out_pcm; long long ld_out_pcm; const int16_t* pcm_lut;   // optional PCM output stage (symbol -> int16 through a 256-entry table)
  int mode, max_steps, d_is_f64;
  const float* causal_b; const float* up_w; const float* up_b;
};


}  // namespace qp
