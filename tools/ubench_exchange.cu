// Micro-benchmarks of the cross-SM exchange primitives the persistent generator can be built on.
// Every variant moves the SAME thing the generator moves per phase: 128 CTAs each produce a
// 256-byte slice (4 channels x 32 utterances, bf16) of a 32 KB activation vector and every CTA
// needs the whole vector in its shared memory before the next round can start.
//
//   flat    : tagged words in global memory, every CTA polls all 32 KB               (generator v2/v3)
//   push    : cluster of 8; a CTA polls only its 4 KB share, then st.async pushes it to the 8 CTAs
//   bulk    : as push, but the share is staged in smem and sent with cp.async.bulk smem->dsmem
//   mcast   : as push, but after the share is seen fresh ONE multicast cp.async.bulk global->8 CTAs
//   pingpong: store -> poll latency between two CTAs
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_exchange.bin tools/ubench_exchange.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int NCTA = 128, NT = 256, CL = 8;
constexpr int VEC_WORDS = 8192;           // 32 KB
constexpr long long TIMEOUT = 3000000000LL;

struct UB {
  unsigned* xbuf;      // [2][VEC_WORDS] tagged exchange vector, double buffered
  long long* out;      // [8] cycles
  int* status;
  int rounds;
  int compute;         // fake compute cycles per round
};

__device__ __forceinline__ uint4 ld_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_u32(const void* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_u32(void* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_v2(void* p, unsigned a, unsigned b) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};\n" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity, int* status) {
  long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 1023u) == 0 && (clock64() - t0 > TIMEOUT || *(volatile int*)status)) { atomicExch(status, 1); return false; }
  }
  return true;
}
__device__ __forceinline__ void st_async_v4(unsigned dst_cluster_addr, uint4 v, unsigned mbar_cluster_addr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];\n"
               ::"r"(dst_cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void bulk_s2c(unsigned dst_cluster_addr, unsigned src_cta_addr, unsigned bytes, unsigned mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void bulk_g2c_mcast(unsigned dst_cta_addr, const void* src, unsigned bytes, unsigned mbar_cta_addr, unsigned short mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n"
               ::"r"(dst_cta_addr), "l"(src), "r"(bytes), "r"(mbar_cta_addr), "h"(mask) : "memory");
}

// each CTA writes its slice of round r: 32 utterances x 2 words (word = 2 bf16 channels, tag in bit 0)
__device__ __forceinline__ void produce(const UB& u, int s, int r) {
  unsigned* buf = u.xbuf + (r & 1) * VEC_WORDS;
  unsigned par = (r >> 1) & 1u;
  if (threadIdx.x < 32) {
    int utt = threadIdx.x;
    unsigned payload = ((unsigned)r << 8) | (unsigned)s << 1;
    st_v2(buf + utt * 256 + 2 * s, (payload << 1 & ~1u) | par, ((payload + 1) << 1 & ~1u) | par);
  }
}
__device__ __forceinline__ bool fresh(uint4 v, unsigned par) { return (((v.x ^ par) | (v.y ^ par) | (v.z ^ par) | (v.w ^ par)) & 1u) == 0; }

__device__ __forceinline__ void fake_compute(int cyc) {
  if (cyc <= 0) return;
  long long t0 = clock64();
  while (clock64() - t0 < cyc) {}
}

// ------------------------------------------------------------------ flat
template <int MODE, int LAYOUT = 0>   // 0: probe-one-then-burst, 1: burst always
__global__ void __launch_bounds__(NT, 1) k_flat(UB u) {
  __shared__ uint4 sm[2048];
  const int s = blockIdx.x, t = threadIdx.x;
  long long t_begin = 0;
  for (int r = 0; r < u.rounds; ++r) {
    if (r == 16 && s == 0 && t == 0) t_begin = clock64();
    fake_compute(u.compute);
    if (LAYOUT == 0) produce(u, s, r);
    else if (t < 32) {
      unsigned par0 = (r >> 1) & 1u;
      st_v2(u.xbuf + (r & 1) * VEC_WORDS + s * 64 + t * 2, ((unsigned)r << 4) | par0, ((unsigned)r << 4) | par0);
    }
    const uint4* g = (const uint4*)(u.xbuf + (r & 1) * VEC_WORDS);
    const unsigned par = (r >> 1) & 1u;
    unsigned pend = 0xFF;
    long long t0 = clock64();
    unsigned spins = 0;
    bool burst = MODE == 1;
    int ip = t & 7;
    while (pend) {
      if (burst) {
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) if (pend & (1u << i)) v[i] = ld_v4(g + t + 256 * i);
        unsigned before = pend;
#pragma unroll
        for (int i = 0; i < 8; ++i) if ((pend & (1u << i)) && fresh(v[i], par)) { sm[t + 256 * i] = v[i]; pend &= ~(1u << i); }
        if (MODE == 0) burst = pend != before && pend != 0;
        if (pend) ip = __ffs(pend) - 1;
      } else {
        uint4 w = ld_v4(g + t + 256 * ip);
        if (fresh(w, par)) { sm[t + 256 * ip] = w; pend &= ~(1u << ip); burst = true; }
      }
      if ((++spins & 1023u) == 0 && (clock64() - t0 > TIMEOUT || *(volatile int*)u.status)) { atomicExch(u.status, 1); return; }
    }
    __syncthreads();
  }
  if (s == 0 && t == 0) { u.out[0] = clock64() - t_begin; u.out[1] = u.rounds - 16; }
}

// ------------------------------------------------------------------ cluster variants
// LAYOUT 0: vector [utt][512 ch]; a CTA's share = utterance rows 4*rank..4*rank+3 (written by ALL producers)
// LAYOUT 1: vector [producer][utt][4 ch]; a CTA's share = producers 16*rank..16*rank+15, each a contiguous,
//           fully coalesced 256-byte block written by ONE warp store of ONE producer
// POLL   0: every thread spins on its own 16-byte piece; 1: same with nanosleep back-off;
//        2: release/acquire flags (producer: stores, bar, fence, flag; consumer: 16 lanes poll 16 flags, then plain loads)
// XFER   0: st.async push to the 8 CTAs; 1: smem staging + cp.async.bulk smem->dsmem; 2: multicast bulk global->cluster;
//        3: none (share only: isolates the global hop)
template <int LAYOUT, int POLL, int XFER>
__global__ void __launch_bounds__(NT, 1) k_cluster(UB u) {
  extern __shared__ __align__(128) unsigned char dyn[];
  uint4 (*sm)[2048] = (uint4 (*)[2048])dyn;
  unsigned long long* bars = (unsigned long long*)(dyn + 2 * 2048 * 16);
  const int s = blockIdx.x, t = threadIdx.x;
  const unsigned rank = cluster_rank();
  if (t == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  long long t_begin = 0, acc_poll = 0, acc_xfer = 0;
  unsigned* flags = u.xbuf + 2 * VEC_WORDS;   // [2][128] flags (POLL 2)
  for (int r = 0; r < u.rounds; ++r) {
    if (r == 16 && s == 0 && t == 0) { t_begin = clock64(); acc_poll = acc_xfer = 0; }
    fake_compute(u.compute);
    const int b = r & 1;
    const unsigned bar = smem_u32(&bars[b]);
    const unsigned expect = XFER == 1 ? 28672u : 32768u;
    if (XFER != 3 && t == 0) mbar_expect_tx(bar, expect);
    const unsigned par = (r >> 1) & 1u;
    unsigned* buf = u.xbuf + b * VEC_WORDS;
    // ---- produce
    if (t < 32) {
      unsigned payload = (((unsigned)r << 8) | ((unsigned)s << 1)) << 1;
      unsigned* dstw = LAYOUT == 0 ? buf + t * 256 + 2 * s : buf + s * 64 + t * 2;
      st_v2(dstw, (payload & ~1u) | par, ((payload + 2) & ~1u) | par);
    }
    if (POLL == 2) {
      __syncthreads();
      if (t == 0) { __threadfence(); st_u32(flags + b * 128 + s, (unsigned)(r >> 1) + 1u); }
    }
    long long tA = clock64();
    const uint4* g = (const uint4*)buf;
    const int piece = rank * 256 + t;
    uint4 v;
    long long t0 = clock64();
    unsigned spins = 0;
    if (POLL == 2) {
      if (t < 16) {
        const unsigned want = (unsigned)(r >> 1) + 1u;
        unsigned f;
        while (true) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(f) : "l"(flags + b * 128 + (LAYOUT == 1 ? rank * 16 + t : t * 8 + (t & 7))) : "memory");
          if (f == want) break;
          if ((++spins & 1023u) == 0 && (clock64() - t0 > TIMEOUT || *(volatile int*)u.status)) { atomicExch(u.status, 1); return; }
        }
      }
      if (LAYOUT == 0) {   // every producer contributes to the share: all 128 flags must be up
        if (t < 128) {
          const unsigned want = (unsigned)(r >> 1) + 1u;
          unsigned f;
          while (true) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(f) : "l"(flags + b * 128 + t) : "memory");
            if (f == want) break;
            if ((++spins & 1023u) == 0 && (clock64() - t0 > TIMEOUT || *(volatile int*)u.status)) { atomicExch(u.status, 1); return; }
          }
        }
      }
      __syncthreads();
      v = ld_v4(g + piece);
    } else {
      while (true) {
        v = ld_v4(g + piece);
        if (fresh(v, par)) break;
        if (POLL == 1) __nanosleep(100);
        if ((++spins & 1023u) == 0 && (clock64() - t0 > TIMEOUT || *(volatile int*)u.status)) { atomicExch(u.status, 1); return; }
      }
    }
    __syncthreads();
    long long tB = clock64();
    const unsigned dst = smem_u32(&sm[b][piece]);
    if (XFER == 0) {
#pragma unroll
      for (unsigned q = 0; q < CL; ++q) st_async_v4(mapa(dst, q), v, mapa(bar, q));
    } else if (XFER == 1) {
      sm[b][piece] = v;
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncthreads();
      if (t < CL && t != rank) bulk_s2c(mapa(smem_u32(&sm[b][rank * 256]), t), smem_u32(&sm[b][rank * 256]), 4096, mapa(bar, t));
    } else if (XFER == 2) {
      if (t == 0) {
        asm volatile("fence.proxy.async;\n" ::: "memory");
        bulk_g2c_mcast(smem_u32(&sm[b][rank * 256]), g + rank * 256, 4096, bar, (unsigned short)0xFF);
      }
    } else {
      sm[b][piece] = v;
    }
    if (XFER != 3) {
      if (!mbar_wait(bar, (r >> 1) & 1u, u.status)) return;
      uint4 chk = sm[b][(t * 7) & 2047];
      if (((chk.x ^ par) & 1u) != 0) atomicExch(u.status, 3);
    }
    __syncthreads();
    long long tC = clock64();
    acc_poll += tB - tA; acc_xfer += tC - tB;
  }
  if (s == 0 && t == 0) { u.out[0] = clock64() - t_begin; u.out[1] = u.rounds - 16; u.out[2] = acc_poll; u.out[3] = acc_xfer; }
  cluster_sync();
}

// ------------------------------------------------------------------ K-split + cluster reduce skeleton
// The phase skeleton of the cluster generator: a CTA polls only its 4 KB K-share (producer-contiguous
// layout), every warp w "computes" a 32x8 partial tile destined for cluster peer w and sends it with ONE
// st.async.v4 per lane (fp16 partials: NV=1) or two (fp32: NV=2); warp 0 of every CTA waits for the 8
// partial tiles on an mbarrier, sums them and publishes the CTA's 256-byte slice for the next round.
template <int NV, int PIPE = 1>
__global__ void __launch_bounds__(NT, 1) k_creduce(UB u) {
  extern __shared__ __align__(128) unsigned char dyn[];
  uint4* smA = (uint4*)dyn;                                // [2][256] the K-share (4 KB) per buffer
  uint4* recv = (uint4*)(dyn + 2 * 4096);                  // [2][8 src][NV][32 lanes]
  unsigned long long* bars = (unsigned long long*)(dyn + 2 * 4096 + 2 * 8 * NV * 512);
  const int s = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const unsigned rank = cluster_rank();
  if (t == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  long long t_begin = 0, acc_poll = 0, acc_xfer = 0;
  // round 0 input: everybody publishes first
  for (int r = 0; r < u.rounds; ++r) {
    if (r == 16 && s == 0 && t == 0) { t_begin = clock64(); acc_poll = acc_xfer = 0; }
    const int b = r & 1;
    const unsigned par = (r >> 1) & 1u;
    const unsigned bar = smem_u32(&bars[b]);
    unsigned* buf = u.xbuf + b * VEC_WORDS;
    if (r == 0 && t < 32) st_v2(buf + s * 64 + t * 2, par, par);
    if (t == 0) mbar_expect_tx(bar, 8 * NV * 512);
    long long tA = clock64();
    // ---- (i) poll the K-share, stage it, "MMA", send partial tile to peer `warp`
    const uint4* g = (const uint4*)buf;
    const int piece = rank * 256 + t;
    uint4 v;
    long long t0 = clock64();
    unsigned spins = 0;
    if (PIPE == 1) {
      while (true) {
        v = ld_v4(g + piece);
        if (fresh(v, par)) break;
        if ((++spins & 1023u) == 0 && (clock64() - t0 > TIMEOUT || *(volatile int*)u.status)) { atomicExch(u.status, 1); return; }
      }
    } else {
      // PIPE polls in flight, issued ~RTT/PIPE apart: detection latency drops from ~RTT/2 to ~RTT/(2*PIPE)
      uint4 q[PIPE];
      bool done = false;
#pragma unroll
      for (int i = 0; i < PIPE; ++i) { q[i] = ld_v4(g + piece); if (i + 1 < PIPE) { long long c0 = clock64(); while (clock64() - c0 < 600 / PIPE) {} } }
      while (!done) {
#pragma unroll
        for (int i = 0; i < PIPE; ++i) {
          if (!done) {
            if (fresh(q[i], par)) { v = q[i]; done = true; }
            else q[i] = ld_v4(g + piece);
          }
        }
        if (!done && (++spins & 1023u) == 0 && (clock64() - t0 > TIMEOUT || *(volatile int*)u.status)) { atomicExch(u.status, 1); return; }
      }
    }
    smA[b * 256 + t] = v;
    __syncthreads();
    long long tB = clock64();
    fake_compute(u.compute);
    {
      uint4 a = smA[b * 256 + ((t * 5) & 255)];
      unsigned dst = smem_u32(&recv[((b * 8 + rank) * NV) * 32 + lane]);
#pragma unroll
      for (int j = 0; j < NV; ++j) st_async_v4(mapa(dst + j * 512, warp), a, mapa(bar, warp));
    }
    // ---- (ii) owner warp: wait for the 8 partial tiles, reduce, publish the slice of the next round
    if (warp == 0) {
      if (!mbar_wait(bar, par, u.status)) return;
      unsigned acc = 0;
#pragma unroll
      for (int src = 0; src < 8; ++src)
#pragma unroll
        for (int j = 0; j < NV; ++j) { uint4 q = recv[((b * 8 + src) * NV + j) * 32 + lane]; acc += q.x + q.y + q.z + q.w; }
      const int nb = (r + 1) & 1;
      const unsigned npar = ((r + 1) >> 1) & 1u;
      st_v2(u.xbuf + nb * VEC_WORDS + s * 64 + lane * 2, (acc & ~1u) | npar, ((acc + 2) & ~1u) | npar);
    }
    long long tC = clock64();
    acc_poll += tB - tA; acc_xfer += tC - tB;
  }
  if (s == 0 && t == 0) { u.out[0] = clock64() - t_begin; u.out[1] = u.rounds - 16; u.out[2] = acc_poll; u.out[3] = acc_xfer; }
  __syncthreads();
  cluster_sync();
}

// ------------------------------------------------------------------ ping-pong latency
__global__ void k_pingpong(UB u, int other) {
  const int s = blockIdx.x;
  if (threadIdx.x != 0 || (s != 0 && s != other)) return;
  unsigned* mine = u.xbuf + (s == 0 ? 0 : 64);
  unsigned* theirs = u.xbuf + (s == 0 ? 64 : 0);
  long long t_begin = clock64();
  for (int r = 1; r <= u.rounds; ++r) {
    if (s == 0) {
      st_u32(mine, r);
      long long t0 = clock64();
      while (ld_u32(theirs) != (unsigned)r) if (clock64() - t0 > TIMEOUT) { atomicExch(u.status, 1); return; }
    } else {
      long long t0 = clock64();
      while (ld_u32(theirs) != (unsigned)r) if (clock64() - t0 > TIMEOUT) { atomicExch(u.status, 1); return; }
      st_u32(mine, r);
    }
  }
  if (s == 0) { u.out[0] = clock64() - t_begin; u.out[1] = 2 * u.rounds; }
}

template <typename K>
static void run(const char* name, K kern, UB u, bool cluster, int compute) {
  u.compute = compute;
  const size_t dsm = cluster ? 2 * 2048 * 16 + 64 : 0;
  if (dsm) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
  CK(cudaMemset(u.xbuf, 0xFF, 2 * VEC_WORDS * 4)); CK(cudaMemset(u.xbuf + 2 * VEC_WORDS, 0, 4096));
  CK(cudaMemset(u.out, 0, 64));
  CK(cudaMemset(u.status, 0, 4));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NCTA); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = dsm; cfg.stream = 0;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (cluster) {
    at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na;
    int ncl = 0;
    cfg.attrs = at; cfg.numAttrs = na;
    CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    if (ncl < NCTA / CL) { printf("%-10s SKIP: only %d clusters of %d can be co-resident (need %d)\n", name, ncl, CL, NCTA / CL); return; }
  }
  at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na;
  cfg.attrs = at; cfg.numAttrs = na;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, u);
  if (le != cudaSuccess) { printf("%-10s launch failed: %s\n", name, cudaGetErrorString(le)); cudaGetLastError(); return; }
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long out[4]; int status;
  CK(cudaMemcpy(out, u.out, 32, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&status, u.status, 4, cudaMemcpyDeviceToHost));
  printf("%-15s compute=%4d  status=%d  cycles/round=%8.1f  poll=%7.1f xfer=%7.1f (kernel %.3f ms, %.3f us/round)\n", name, compute, status,
         out[1] ? (double)out[0] / out[1] : 0.0, out[1] ? (double)out[2] / out[1] : 0.0, out[1] ? (double)out[3] / out[1] : 0.0, ms, ms * 1e3 / u.rounds);
}

int main(int argc, char** argv) {
  UB u;
  u.rounds = argc > 1 ? atoi(argv[1]) : 4000;
  CK(cudaMalloc(&u.xbuf, 2 * VEC_WORDS * 4 + 4096));
  CK(cudaMalloc(&u.out, 64));
  CK(cudaMalloc(&u.status, 4));
  int comps[3] = {0, 200, 1000};
  for (int ci = 0; ci < 3; ++ci) {
    int c = comps[ci];
    run("flat L1 burst", k_flat<1, 1>, u, false, c);
    run("flat L1 probe", k_flat<0, 1>, u, false, c);
    run("creduce fp16", k_creduce<1>, u, true, c);
    run("creduce pipe2", k_creduce<1, 2>, u, true, c);
    run("creduce pipe4", k_creduce<1, 4>, u, true, c);
    run("creduce fp32", k_creduce<2>, u, true, c);
    if (ci != 1) {
      run("L1 spin none", k_cluster<1, 0, 3>, u, true, c);
      run("L1 spin mcast", k_cluster<1, 0, 2>, u, true, c);
    }
  }
  {
    int others[3] = {1, 64, 127};
    for (int i = 0; i < 3; ++i) {
      CK(cudaMemset(u.xbuf, 0, 1024)); CK(cudaMemset(u.status, 0, 4));
      k_pingpong<<<NCTA, 32>>>(u, others[i]);
      CK(cudaDeviceSynchronize());
      long long out[2];
      CK(cudaMemcpy(out, u.out, 16, cudaMemcpyDeviceToHost));
      printf("pingpong 0<->%d: one-way store->seen = %.1f cycles\n", others[i], (double)out[0] / out[1]);
    }
  }
  return 0;
}
