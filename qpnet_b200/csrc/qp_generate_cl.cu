// Cluster generator: QPNet.batch_fast_generate (qpnet.py:314-559) for the SI default widths
// (n_resch 512, n_skipch 256, n_quantize 256), up to 32 utterances per launch.
//
// Why this shape (measured with tools/ubench_exchange.cu on B200, see profiles/):
//   * a store -> remote-poll hop through L2 costs ~850 cycles one way; every contraction of the stack
//     needs the whole activation vector, so a sample step is a chain of 35 such hops.  The hop must
//     therefore move as little as possible and be polled by as few threads as possible;
//   * partial-sector writes by many SMs into one line are ~5x slower than one coalesced 256-byte
//     block per producer, so every exchange vector is stored PRODUCER-major: [owner][utt][channels];
//   * distributed shared memory moves only ~10 B/cycle/SM, but its latency (~400 cycles) is half a
//     global hop, so clusters exchange small fp16 partial sums, never activations.
//
// Decomposition: 128 CTAs = 32 clusters x 4 ranks, one CTA per SM, co-resident (cooperative launch;
// a B200 keeps 33 such clusters resident with one 160 KB CTA per SM, but only 15 clusters of 8).
//   * CTA o OWNS residual channels [4o, 4o+4), skip rows [2o, 2o+2) and head rows [2o, 2o+2): it
//     finishes those rows (bias, gate non-linearity, residual carry in fp32, skip accumulation) and
//     publishes them as one tagged 256-byte block.
//   * inside cluster c, rank r contracts over the K-SHARE [128r, 128r+128) of the input vector for all
//     32 rows the cluster owns: it polls only its 8 KB share (two 16-byte loads per thread), runs
//     m16n8k16 tensor-core tiles (utterances = M, rows = N; warp w computes the 8 rows rank w & 3 owns
//     over K-half w >> 2) and sends the 32x8 partial tile to that rank with ONE st.async per lane
//     (fp16, 512 B per warp), completing on its mbarrier.  The owner warp sums the 8 partial tiles.
//   * the past tap x(t-k) (qpnet.py:81-87, 457-502) comes from a per-CTA ring of its own K-share in
//     global memory (the reference's FIFOs, qpnet.py:388-393, 431-437), prefetched with cp.async two
//     phases ahead together with the block's weight tile; k = dil (fixed) or -round(-d[t]*dil) (adaptive, qpnet.py:616-617 / 621-622),
//     k == 0 -> oldest entry (caveat C4).
//   * block 0 needs no exchange at all: its input is the causal layer, a function of the last three
//     symbols, so its gate pre-activation is three table lookups (W.E folded at pack time, fp32).
//   * aux 1x1 (qpnet.py:663-664, 632-633) is a 6-MMA side product of the owner warp, computed while
//     the partial tiles are in flight.
// Exchange words carry a 1-bit epoch tag (LSB of the low bf16 / of the fp32 logit); consumers poll
// the data itself.  Every spin has a watchdog (QP_ETIMEOUT) and every exit goes through a cluster barrier.
#include <algorithm>

#include <cuda_fp16.h>

#include "qp_gen_common.cuh"
#include "qp_pack.cuh"

namespace qp {
namespace cl {

constexpr int NT = 256, NW = 8, CL = 4;
constexpr int UB = 32;                       // utterance slots per launch
constexpr int C = 512, S = 256, Q = 256, AP = 48;
constexpr int NOWN = C / 4;                  // 128 owner CTAs
constexpr int KS = C / CL;                   // 128: K-share of a 512-vector (two warps' worth: K-halves of 64)
constexpr int KH = S / CL;                   // 64: K-share of a 256-vector
constexpr int NR = 8 * CL;                   // 32 rows per cluster tile: 8 per owner rank
constexpr int NPART = 8;                     // partial tiles an owner receives: CL ranks x 2 K-halves
constexpr int PA = KS + 8;                   // activation tile pitch (elements)
constexpr int PWG = 2 * KS + 8, PWR = KS + 8, PWH = KH + 8;   // weight tile pitches
constexpr int PH = AP + 8;                   // aux tile pitch
constexpr int HR = 40;                       // Hraw pitch (floats)
constexpr int WTILE = NR * PWG * 2;          // bytes of the largest weight tile
constexpr int WSLOT = WTILE + UB * PA * 2;   // weight tile + the past-tap rows that travel with a gate tile
constexpr int NSLOT = 3;
constexpr int MAXL = 16;
constexpr int TRACE_EVENTS = 4;
enum { K_GATE = 0, K_RES = 1, K_HEAD1 = 2, K_HEAD2 = 3 };

struct Plan {
  int A, L, nF, nA, U, B, F, M;
  int dil[MAXL], depth[MAXL], ring_size[MAXL];
  int32_t* status;
  const float** tab;
  __nv_bfloat16* Wgate;   // [L][NOWN][NR][2*KS]   rows: 8 per owner rank (s0,t0,s1,t1,s2,t2,s3,t3); K: [current | past]
  __nv_bfloat16* Wres;    // [L][NOWN][NR][KS]     rows per owner rank: r0..r3, k0, k1, 0, 0
  __nv_bfloat16* Whead;   // [2][NOWN][NR][KH]     rows per owner rank: h0, h1, 0 ...
  __nv_bfloat16* Vaux;    // [L][NOWN][8][AP]
  float* bgate;           // [L][NOWN][8]
  float* bres;            // [L][NOWN][8]
  float* bhead;           // [2][NOWN][8]
  float* T0;              // [NOWN][3][Q][8]  block-0 gate tables: cur symbol, previous, the one before
  float* Eo;              // [NOWN][2][Q][4]  causal-layer rows of the owned channels (bias folded into tap 1)
  __nv_bfloat16* ring[MAXL];  // [ring_size][NOWN][UB][KS], l >= 1
  uint32_t* v512;         // [NOWN][UB][2]  x / z exchange, alternating epochs
  uint32_t* v256;         // [NOWN][UB]     relu(skip sum) / relu(head-1)
  uint32_t* vlog;         // [NOWN][UB][2]  fp32 logits
  uint32_t* vsym;         // [UB][32]       fed-back symbol, one line per utterance
  void* tagged_begin; size_t tagged_bytes;
  long long* trace; int trace_step0, trace_nsteps;
};

static int pow2_above(int v) { int q = 1; while (q <= v) q <<= 1; return q; }

bool supported(const QpArch* a, int B) {
  if (a->n_resch != C || a->n_skipch != S || a->n_quantize != Q || a->n_aux > AP) return false;
  if (a->n_fixed + a->n_adaptive > MAXL || a->n_fixed < 1 || a->dil_fixed[0] != 1) return false;
  return B >= 1 && B <= UB;
}

size_t make_plan(const QpArch* a, int B, int F, int M, void* base, size_t cap, Plan* p) {
  p->A = a->n_aux; p->nF = a->n_fixed; p->nA = a->n_adaptive; p->L = p->nF + p->nA; p->U = a->upsampling;
  p->B = B; p->F = F; p->M = M;
  const int L = p->L;
  Arena ar(base, cap);
  p->status = ar.take<int32_t>(64);
  p->tab = ar.take<const float*>(tensor_map(a).count());
  p->Wgate = ar.take<__nv_bfloat16>((size_t)L * NOWN * NR * 2 * KS);
  p->Wres = ar.take<__nv_bfloat16>((size_t)L * NOWN * NR * KS);
  p->Whead = ar.take<__nv_bfloat16>((size_t)2 * NOWN * NR * KH);
  p->Vaux = ar.take<__nv_bfloat16>((size_t)L * NOWN * 8 * AP);
  p->bgate = ar.take<float>((size_t)L * NOWN * 8);
  p->bres = ar.take<float>((size_t)L * NOWN * 8);
  p->bhead = ar.take<float>((size_t)2 * NOWN * 8);
  p->T0 = ar.take<float>((size_t)NOWN * 3 * Q * 8);
  p->Eo = ar.take<float>((size_t)NOWN * 2 * Q * 4);
  for (int l = 0; l < L; ++l) {
    p->dil[l] = l < p->nF ? a->dil_fixed[l] : a->dil_adaptive[l - p->nF];
    p->depth[l] = l < p->nF ? p->dil[l] : p->dil[l] * M;
    p->ring_size[l] = pow2_above(p->depth[l]);
    p->ring[l] = l >= 1 ? ar.take<__nv_bfloat16>((size_t)p->ring_size[l] * NOWN * UB * KS) : nullptr;
  }
  ar.off = align_up(ar.off, 256);
  size_t t0 = ar.off;
  p->v512 = ar.take<uint32_t>((size_t)NOWN * UB * 2);
  p->v256 = ar.take<uint32_t>((size_t)NOWN * UB);
  p->vlog = ar.take<uint32_t>((size_t)NOWN * UB * 2);
  p->vsym = ar.take<uint32_t>((size_t)UB * 32);
  ar.off = align_up(ar.off, 256);
  p->tagged_begin = base ? (char*)base + t0 : nullptr;
  p->tagged_bytes = ar.off - t0;
  p->trace = ar.take<long long>((size_t)8 * (2 * L + 3) * TRACE_EVENTS);
  p->trace_step0 = -100; p->trace_nsteps = 8;
  return align_up(ar.off, 256);
}

// ------------------------------------------------------------------ weight packing
__global__ void pack_kernel(TensorMap tm, Plan p, const float* const* __restrict__ tab) {
  const int L = p.L, A = p.A, nF = p.nF;
  const size_t n_wg = (size_t)L * NOWN * NR * 2 * KS, n_wr = (size_t)L * NOWN * NR * KS, n_wh = (size_t)2 * NOWN * NR * KH;
  const size_t n_va = (size_t)L * NOWN * 8 * AP, n_b = (size_t)L * NOWN * 8, n_bh = (size_t)2 * NOWN * 8;
  const size_t n_eo = (size_t)NOWN * 2 * Q * 4;
  const size_t total = n_wg + n_wr + n_wh + n_va + 2 * n_b + n_bh + n_eo;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t k = idx;
    if (k < n_wg) {
      int kk = (int)(k % (2 * KS)); size_t q = k / (2 * KS);
      int row = (int)(q % NR); q /= NR;
      int s = (int)(q % NOWN), l = (int)(q / NOWN);
      int r = s % CL, o = (s / CL) * CL + row / 8, j = row % 8, g = j & 1, ch = 4 * o + (j >> 1);
      bool past = kk >= KS;
      int col = KS * r + (kk % KS);
      float v;
      if (l < nF) v = tab[tm.dilF_w(g, l)][((size_t)ch * C + col) * 2 + (past ? 0 : 1)];
      else v = tab[past ? tm.dilA_wP(g, l - nF) : tm.dilA_wC(g, l - nF)][(size_t)ch * C + col];
      p.Wgate[k] = __float2bfloat16(v);
      continue;
    }
    k -= n_wg;
    if (k < n_wr) {
      int kk = (int)(k % KS); size_t q = k / KS;
      int row = (int)(q % NR); q /= NR;
      int s = (int)(q % NOWN), l = (int)(q / NOWN);
      int r = s % CL, o = (s / CL) * CL + row / 8, j = row % 8, col = KS * r + kk;
      float v = 0.f;
      if (j < 4) { int ch = 4 * o + j; v = l < nF ? tab[tm.resF_w(l)][(size_t)ch * C + col] : tab[tm.resA_w(l - nF)][(size_t)ch * C + col]; }
      else if (j < 6) { int sr = 2 * o + j - 4; v = l < nF ? tab[tm.skipF_w(l)][(size_t)sr * C + col] : tab[tm.skipA_w(l - nF)][(size_t)sr * C + col]; }
      p.Wres[k] = __float2bfloat16(v);
      continue;
    }
    k -= n_wr;
    if (k < n_wh) {
      int kk = (int)(k % KH); size_t q = k / KH;
      int row = (int)(q % NR); q /= NR;
      int s = (int)(q % NOWN), hd = (int)(q / NOWN);
      int r = s % CL, o = (s / CL) * CL + row / 8, j = row % 8, col = KH * r + kk;
      float v = 0.f;
      if (j < 2) v = tab[hd ? tm.post2_w() : tm.post1_w()][(size_t)(2 * o + j) * S + col];
      p.Whead[k] = __float2bfloat16(v);
      continue;
    }
    k -= n_wh;
    if (k < n_va) {
      int a = (int)(k % AP); size_t q = k / AP;
      int j = (int)(q % 8); q /= 8;
      int o = (int)(q % NOWN), l = (int)(q / NOWN);
      int g = j & 1, ch = 4 * o + (j >> 1);
      float v = 0.f;
      if (a < A) v = l < nF ? tab[tm.auxF_w(g, l)][(size_t)ch * A + a] : tab[tm.auxA_w(g, l - nF)][(size_t)ch * A + a];
      p.Vaux[k] = __float2bfloat16(v);
      continue;
    }
    k -= n_va;
    if (k < n_b) {   // gate biases: every bias that feeds the pre-activation (block 0: + (Wc + Wp) . causal bias)
      int j = (int)(k % 8); size_t q = k / 8;
      int o = (int)(q % NOWN), l = (int)(q / NOWN);
      int g = j & 1, ch = 4 * o + (j >> 1);
      float v;
      if (l < nF) v = tab[tm.dilF_b(g, l)][ch] + tab[tm.auxF_b(g, l)][ch];
      else { int a = l - nF; v = tab[tm.dilA_bC(g, a)][ch] + tab[tm.dilA_bP(g, a)][ch] + tab[tm.auxA_b(g, a)][ch]; }
      if (l == 0) {
        const float* W = tab[tm.dilF_w(g, 0)] + (size_t)ch * C * 2;
        const float* cb = tab[tm.causal_b()];
        float acc = 0.f;
        for (int col = 0; col < C; ++col) acc += (W[2 * col] + W[2 * col + 1]) * cb[col];
        v += acc;
      }
      p.bgate[k] = v;
      continue;
    }
    k -= n_b;
    if (k < n_b) {
      int j = (int)(k % 8); size_t q = k / 8;
      int o = (int)(q % NOWN), l = (int)(q / NOWN);
      float v = 0.f;
      if (j < 4) v = l < nF ? tab[tm.resF_b(l)][4 * o + j] : tab[tm.resA_b(l - nF)][4 * o + j];
      else if (j < 6) v = l < nF ? tab[tm.skipF_b(l)][2 * o + j - 4] : tab[tm.skipA_b(l - nF)][2 * o + j - 4];
      p.bres[k] = v;
      continue;
    }
    k -= n_b;
    if (k < n_bh) {
      int j = (int)(k % 8); size_t q = k / 8;
      int o = (int)(q % NOWN), hd = (int)(q / NOWN);
      p.bhead[k] = j < 2 ? tab[hd ? tm.post2_b() : tm.post1_b()][2 * o + j] : 0.f;
      continue;
    }
    k -= n_bh;
    {
      int j = (int)(k % 4); size_t q = k / 4;
      int sym = (int)(q % Q); q /= Q;
      int tap = (int)(q % 2), o = (int)(q / 2);
      int ch = 4 * o + j;
      p.Eo[k] = tab[tm.causal_w()][((size_t)ch * Q + sym) * 2 + tap] + (tap ? tab[tm.causal_b()][ch] : 0.f);
    }
  }
}

// Block-0 gate tables (fp32): the causal layer output is E0[s(t-2)] + E1[s(t-1)] + b (qpnet.py:447-448,
// 561-564), so Wc.x0(t) + Wp.x0(t-1) = TA[s(t-1)] + TB[s(t-2)] + TC[s(t-3)] + const with
//   TA = Wc.E1,  TB = Wc.E0 + Wp.E1,  TC = Wp.E0.      grid (8 rows, NOWN), block Q threads (one symbol each)
__global__ void table_kernel(TensorMap tm, Plan p, const float* const* __restrict__ tab) {
  const int j = blockIdx.x, o = blockIdx.y, sym = threadIdx.x;
  const int g = j & 1, ch = 4 * o + (j >> 1);
  const float* W = tab[tm.dilF_w(g, 0)] + (size_t)ch * C * 2;   // [col][tap]: tap 0 past, tap 1 current
  const float* E = tab[tm.causal_w()];                          // [col][Q][tap]
  float ta = 0.f, tb = 0.f, tc = 0.f;
  for (int col = 0; col < C; ++col) {
    float wp = W[2 * col], wc = W[2 * col + 1];
    float2 e = *(const float2*)(E + ((size_t)col * Q + sym) * 2);   // (E0, E1)
    ta = fmaf(wc, e.y, ta);
    tb = fmaf(wc, e.x, fmaf(wp, e.y, tb));
    tc = fmaf(wp, e.x, tc);
  }
  float* T = p.T0 + (size_t)o * 3 * Q * 8;
  T[(0 * Q + sym) * 8 + j] = ta;
  T[(1 * Q + sym) * 8 + j] = tb;
  T[(2 * Q + sym) * 8 + j] = tc;
}

// ------------------------------------------------------------------ device helpers
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

struct SmemMap {   // byte offsets into the dynamic shared memory
  int acur, wslot, recv, t0, eo, vaux, haux, hraw, bg, br, bh, bars, abort, total;
};
__host__ __device__ inline SmemMap smem_map(int L) {
  SmemMap m;
  int o = 0;
  m.acur = o; o += 2 * UB * PA * 2;
  o = (o + 127) & ~127;
  m.wslot = o; o += NSLOT * WSLOT;
  m.recv = o; o += 2 * NPART * 32 * 16;
  m.t0 = o; o += 3 * Q * 8 * 4;
  m.eo = o; o += 2 * Q * 4 * 4;
  m.vaux = o; o += L * 8 * PH * 2;
  m.haux = o; o += 2 * UB * PH * 2;
  m.hraw = o; o += UB * HR * 4;
  m.bg = o; o += L * 8 * 4;
  m.br = o; o += L * 8 * 4;
  m.bh = o; o += 2 * 8 * 4;
  o = (o + 15) & ~15;
  m.bars = o; o += 16;
  m.abort = o; o += 16;
  m.total = o;
  return m;
}

template <bool TRACE>
__global__ void __launch_bounds__(NT, 1) cl_gen_kernel(Plan p, GenArgsDev g) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int L = p.L, A = p.A, B = p.B, U = p.U;
  const SmemMap sm = smem_map(L);
  __nv_bfloat16* const sAcur = (__nv_bfloat16*)(smem + sm.acur);
  unsigned char* const sW = smem + sm.wslot;
  uint4* const sRecv = (uint4*)(smem + sm.recv);
  const float* const sT0 = (const float*)(smem + sm.t0);
  const float* const sEo = (const float*)(smem + sm.eo);
  __nv_bfloat16* const sVaux = (__nv_bfloat16*)(smem + sm.vaux);
  __nv_bfloat16* const sHaux = (__nv_bfloat16*)(smem + sm.haux);
  float* const sHraw = (float*)(smem + sm.hraw);
  float* const sBg = (float*)(smem + sm.bg);
  float* const sBr = (float*)(smem + sm.br);
  float* const sBh = (float*)(smem + sm.bh);
  unsigned long long* const sBars = (unsigned long long*)(smem + sm.bars);
  volatile int* const sAbort = (volatile int*)(smem + sm.abort);

  const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned rank = cluster_rank();
  const int q4 = lane >> 2, i4 = lane & 3;      // accumulator fragment coordinates: utterances q4 + 8m, rows 2*i4, 2*i4+1
  const int half = Q / 2;
  const int nphase = 2 * L + 3;
  const long long ldd = (long long)p.F * U;

  // ---- one-time staging ---------------------------------------------------------------
  for (int e = tid; e < (sm.total - sm.acur) / 16; e += NT) ((uint4*)smem)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int e = tid; e < 3 * Q * 8; e += NT) ((float*)(smem + sm.t0))[e] = p.T0[(size_t)s * 3 * Q * 8 + e];
  for (int e = tid; e < 2 * Q * 4; e += NT) ((float*)(smem + sm.eo))[e] = p.Eo[(size_t)s * 2 * Q * 4 + e];
  for (int e = tid; e < L * 8 * AP; e += NT) {
    int l = e / (8 * AP), rem = e - l * 8 * AP, j = rem / AP, a = rem - j * AP;
    sVaux[(l * 8 + j) * PH + a] = p.Vaux[((size_t)l * NOWN + s) * 8 * AP + rem];
  }
  for (int e = tid; e < L * 8; e += NT) {
    sBg[e] = p.bgate[((size_t)(e >> 3) * NOWN + s) * 8 + (e & 7)];
    sBr[e] = p.bres[((size_t)(e >> 3) * NOWN + s) * 8 + (e & 7)];
  }
  if (tid < 16) sBh[tid] = p.bhead[((size_t)(tid >> 3) * NOWN + s) * 8 + (tid & 7)];
  if (tid == 0) {
    mbar_init(smem_u32(&sBars[0]), 1);
    mbar_init(smem_u32(&sBars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // aux rows of the priming region: h_up[:, 0] (replicate pad, qpnet.py:359)
  for (int e = tid; e < UB * A; e += NT) {
    int u = e / A, a = e - u * A;
    sHraw[u * HR + a] = u < B ? g.h[((size_t)u * A + a) * p.F] : 0.f;
  }
  __syncthreads();
  {
    const float w0 = g.up_w[0], bb = g.up_b[0];
    for (int e = tid; e < UB * A; e += NT) {
      int u = e / A, a = e - u * A;
      sHaux[(1 * UB + u) * PH + a] = __float2bfloat16(sHraw[u * HR + a] * w0 + bb);   // slot (t & 1) of t = -1
    }
  }
  __syncthreads();
  cluster_sync();

  auto trace = [&](int t, int phase, int ev) {
    if (TRACE && s == 0 && tid == 0 && t >= p.trace_step0 && t < p.trace_step0 + p.trace_nsteps)
      p.trace[((size_t)(t - p.trace_step0) * nphase + phase) * TRACE_EVENTS + ev] = clock64();
  };

  // ---- weight tile stream: one tile per MMA phase, NSLOT-deep ring, prefetch distance 2 ------
  // per step: tile 2l = res l, tile 2l-1 = gate l (l >= 1), then head-1, head-2 (not in the priming step)
  int pf_t = -1, pf_i = 0, pf_n = 0;   // cursor of the next tile to fetch
  auto issue_next_tile = [&]() {
    if (pf_t < g.max_steps) {
      unsigned char* dst = sW + (pf_n % NSLOT) * WSLOT;
      if (pf_i < 2 * L - 1) {
        if (pf_i & 1) {   // gate (pf_i + 1) / 2: weight tile + the past-tap rows x_l(t - k) of this CTA's K-share
          const int l = (pf_i + 1) >> 1;
          const __nv_bfloat16* src = p.Wgate + ((size_t)l * NOWN + s) * NR * 2 * KS;
          for (int e = tid; e < NR * 32; e += NT) cp_async16(dst + ((e >> 5) * PWG + (e & 31) * 8) * 2, src + (e >> 5) * 2 * KS + (e & 31) * 8);
          if (pf_t >= 0) {
            const int u = tid >> 3;
            int k = p.dil[l];
            if (l >= p.nF) {   // pitch-dependent look-back of this step (qpnet.py:476-483, 613-624)
              k = 0;
              if (u < B)
                k = g.d_is_f64 ? -gen_index_f64(((const double*)g.d)[(long long)u * ldd + pf_t], p.dil[l])
                               : -gen_index_f32(((const float*)g.d)[(long long)u * ldd + pf_t], p.dil[l]);
              if (k <= 0 || k > p.depth[l]) k = p.depth[l];   // k == 0: python index 0 = oldest entry (C4)
            }
            const int slot = (pf_t - k) & (p.ring_size[l] - 1);
            const __nv_bfloat16* rsrc = p.ring[l] + (((size_t)slot * NOWN + s) * UB + u) * KS;
            unsigned char* pdst = dst + WTILE + (u * PA) * 2;
#pragma unroll
            for (int cch = (tid & 7); cch < KS / 8; cch += 8) cp_async16(pdst + cch * 16, rsrc + cch * 8);
          }
        } else {          // res pf_i / 2
          const int l = pf_i >> 1;
          const __nv_bfloat16* src = p.Wres + ((size_t)l * NOWN + s) * NR * KS;
          for (int e = tid; e < NR * 16; e += NT) cp_async16(dst + ((e >> 4) * PWR + (e & 15) * 8) * 2, src + (e >> 4) * KS + (e & 15) * 8);
        }
      } else {
        const int hd = pf_i - (2 * L - 1);
        const __nv_bfloat16* src = p.Whead + ((size_t)hd * NOWN + s) * NR * KH;
        for (int e = tid; e < NR * 8; e += NT) cp_async16(dst + ((e >> 3) * PWH + (e & 7) * 8) * 2, src + (e >> 3) * KH + (e & 7) * 8);
      }
      ++pf_i;
      const int ntiles = pf_t < 0 ? 2 * L - 1 : 2 * L + 1;
      if (pf_i == ntiles) { pf_i = 0; ++pf_t; }
    }
    ++pf_n;
    cp_async_commit();
  };
  issue_next_tile();
  issue_next_tile();

  // ---- owner-warp state ----------------------------------------------------------------
  float xc[4][2], sk[4][2];          // fp32 residual carry (lanes i4 < 2) / skip accumulators (lanes i4 == 2)
#pragma unroll
  for (int m = 0; m < 4; ++m) { xc[m][0] = xc[m][1] = 0.f; sk[m][0] = sk[m][1] = 0.f; }
  int sy_c = half, sy_p1 = half, sy_p2 = half;   // lane u: s(t-1), s(t-2), s(t-3) of utterance u
  int rp = 0;                                    // MMA phase counter: activation / receive / barrier double buffering

  auto spin_check = [&](unsigned& spins, long long& t0) -> bool {
    if ((++spins & 1023u) != 0) return false;
    if (t0 == 0) t0 = clock64();
    if (*((volatile int32_t*)p.status) != 0) return true;
    if (clock64() - t0 > GEN_TIMEOUT_CYCLES) { atomicExch(p.status, QP_ETIMEOUT); return true; }
    return false;
  };

  // aux 1x1 of the owned 8 gate rows for all utterances: 6 MMAs (qpnet.py:663-664 / 632-633)
  auto aux_mma = [&](int l, int t, float (&ax)[2][4]) {
#pragma unroll
    for (int m = 0; m < 2; ++m) ax[m][0] = ax[m][1] = ax[m][2] = ax[m][3] = 0.f;
    const __nv_bfloat16* hp = sHaux + ((t & 1) * UB + (lane & 15)) * PH + (lane >> 4) * 8;
    const __nv_bfloat16* vp = sVaux + (l * 8 + (lane & 7)) * PH + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < AP / 16; ++ks) {
      unsigned b0, b1, a0, a1, a2, a3;
      ldmatrix_x2(b0, b1, vp + ks * 16);
      ldmatrix_x4(a0, a1, a2, a3, hp + ks * 16);
      mma_bf16(ax[0], a0, a1, a2, a3, b0, b1);
      ldmatrix_x4(a0, a1, a2, a3, hp + 16 * PH + ks * 16);
      mma_bf16(ax[1], a0, a1, a2, a3, b0, b1);
    }
  };
  // gate non-linearity + publication of the owned z slice (epoch = parity of the 512-vector write counter)
  auto publish_gate = [&](int l, const float (&pre)[4][2], unsigned par) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float z = fast_sigmoid(pre[m][0] + sBg[l * 8 + 2 * i4]) * fast_tanh(pre[m][1] + sBg[l * 8 + 2 * i4 + 1]);
      float zn = __shfl_xor_sync(0xffffffffu, z, 1);
      if (!(i4 & 1)) st_strong_u32(p.v512 + ((size_t)s * UB + q4 + 8 * m) * 2 + (i4 >> 1), pack_tagged(z, zn, par));
    }
  };

  // =========================================================================== time loop
  for (int t = -1; t < g.max_steps; ++t) {
    const bool prime = t < 0;
    const unsigned w512 = (unsigned)(t + 1) * (unsigned)(2 * L - 1);   // 512-vector writes before this step

    // ================================================================ block 0 gate: symbols -> tables
    trace(t, 0, 0);
    if (warp == 0) {
      int bad = 0;
      if (t == 0) {
        sy_p2 = sy_p1; sy_p1 = sy_c;
        sy_c = lane < B ? (int)(((g.seed[lane] % Q) + Q) % Q) : half;   // qpnet.py:356-358: pad with Q/2, keep the seed last
      } else if (t >= 1) {
        int nw = half;
        if (lane < B) {
          const unsigned want = ((unsigned)(t - 1) & 1u) << 30;
          unsigned spins = 0; long long t0 = 0;
          while (true) {
            unsigned w = ld_strong_u32(p.vsym + lane * 32);
            if (((w ^ want) & 0x40000000u) == 0) { nw = (int)(w & 0xFFFFu) % Q; break; }
            if (spin_check(spins, t0)) { bad = 1; break; }
          }
        }
        sy_p2 = sy_p1; sy_p1 = sy_c; sy_c = nw;
      }
      bad = __any_sync(0xffffffffu, bad);
      trace(t, 0, 1);
      trace(t, 0, 2);
      if (bad) {
        if (lane == 0) *sAbort = 1;
      } else {
        float ax[2][4];
        aux_mma(0, t, ax);
        float pre[4][2];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int u = q4 + 8 * m;
          const int c_ = __shfl_sync(0xffffffffu, sy_c, u), a_ = __shfl_sync(0xffffffffu, sy_p1, u), b_ = __shfl_sync(0xffffffffu, sy_p2, u);
          const float2 ta = *(const float2*)(sT0 + ((0 * Q + c_) * 8 + 2 * i4));
          const float2 tb = *(const float2*)(sT0 + ((1 * Q + a_) * 8 + 2 * i4));
          const float2 tc = *(const float2*)(sT0 + ((2 * Q + b_) * 8 + 2 * i4));
          pre[m][0] = ta.x + tb.x + tc.x + ax[m >> 1][(m & 1) * 2];
          pre[m][1] = ta.y + tb.y + tc.y + ax[m >> 1][(m & 1) * 2 + 1];
          if (i4 < 2) {   // fp32 residual stream of the owned channels restarts from the causal layer
            const float2 e0 = *(const float2*)(sEo + ((0 * Q + a_) * 4 + 2 * i4));
            const float2 e1 = *(const float2*)(sEo + ((1 * Q + c_) * 4 + 2 * i4));
            xc[m][0] = e0.x + e1.x; xc[m][1] = e0.y + e1.y;
          }
          sk[m][0] = sk[m][1] = 0.f;
        }
        publish_gate(0, pre, (w512 + 0u) & 1u);
      }
    }
    trace(t, 0, 3);

    // ================================================================ MMA phases
    const int nmma = prime ? 2 * L - 1 : 2 * L + 1;
    for (int ph = 0; ph < nmma; ++ph) {
      int kind, l;
      if (ph < 2 * L - 1) { kind = (ph & 1) ? K_GATE : K_RES; l = (ph + 1) >> 1; }
      else { kind = ph == 2 * L - 1 ? K_HEAD1 : K_HEAD2; l = 0; }
      const int tph = kind == K_GATE ? 2 * l : kind == K_RES ? 2 * l + 1 : kind == K_HEAD1 ? 2 * L : 2 * L + 1;
      trace(t, tph, 0);
      const int ab = rp & 1;
      __nv_bfloat16* Acur = sAcur + ab * UB * PA;
      const unsigned bar = smem_u32(&sBars[ab]);
      if (tid == 0) mbar_expect_tx(bar, NPART * 512);

      // ---- (1) poll this rank's K-share of the input vector, stage it as the MMA A tile
      int fail = 0;
      if (kind == K_GATE || kind == K_RES) {
        // res l reads z_l (write index 2l), gate l reads x_l (write index 2l - 1)
        const unsigned par = (w512 + (unsigned)(kind == K_RES ? 2 * l : 2 * l - 1)) & 1u;
        // share = owner blocks 32*rank .. 32*rank+31 (16 pieces of 16 bytes each): two pieces per thread
        const uint4* src = (const uint4*)p.v512 + (size_t)(32 * rank) * 16 + tid;
        uint4 v[2];
        unsigned pend = 3;
        unsigned spins = 0; long long t0 = 0;
        while (pend) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (pend & (1u << j)) v[j] = ld_strong_v4(src + 256 * j);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if ((pend & (1u << j)) && fresh4(v[j], par)) pend &= ~(1u << j);
          if (pend && spin_check(spins, t0)) { fail = 1; break; }
        }
        const int rs = kind == K_GATE ? p.ring_size[l] : 1;
        const size_t slot_stride = (size_t)NOWN * UB * KS;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int pidx = tid + 256 * j;
          const int u0 = 2 * (pidx & 15), col = 4 * (pidx >> 4);
          *(uint2*)(Acur + u0 * PA + col) = make_uint2(v[j].x, v[j].y);
          *(uint2*)(Acur + (u0 + 1) * PA + col) = make_uint2(v[j].z, v[j].w);
          if (kind == K_GATE && !fail) {
            // keep x_l(t) for the past taps of later steps; the priming step fills the whole ring with
            // the constant of the pad region (qpnet.py:355-440)
            __nv_bfloat16* r0 = p.ring[l] + ((size_t)s * UB + u0) * KS + col;
            if (prime) {
              for (int sl = 0; sl < rs; ++sl) {
                *(uint2*)(r0 + sl * slot_stride) = make_uint2(v[j].x, v[j].y);
                *(uint2*)(r0 + sl * slot_stride + KS) = make_uint2(v[j].z, v[j].w);
              }
            } else {
              __nv_bfloat16* r1 = r0 + (size_t)(t & (rs - 1)) * slot_stride;
              *(uint2*)r1 = make_uint2(v[j].x, v[j].y);
              *(uint2*)(r1 + KS) = make_uint2(v[j].z, v[j].w);
            }
          }
        }
      } else {
        const unsigned par = kind == K_HEAD1 ? 0u : 1u;
        // share = owner blocks 32*rank .. 32*rank+31 (8 pieces each): one piece per thread
        const uint4* src = (const uint4*)p.v256 + (size_t)(32 * rank) * 8 + tid;
        uint4 v;
        unsigned spins = 0; long long t0 = 0;
        while (true) {
          v = ld_strong_v4(src);
          if (fresh4(v, par)) break;
          if (spin_check(spins, t0)) { fail = 1; break; }
        }
        const int u0 = 4 * (tid & 7), col = 2 * (tid >> 3);
        *(unsigned*)(Acur + u0 * PA + col) = v.x;
        *(unsigned*)(Acur + (u0 + 1) * PA + col) = v.y;
        *(unsigned*)(Acur + (u0 + 2) * PA + col) = v.z;
        *(unsigned*)(Acur + (u0 + 3) * PA + col) = v.w;
      }
      cp_async_wait<1>();   // this phase's weight tile (and the past rows fetched with tile 1) have landed
      if (__syncthreads_or(fail | *sAbort)) goto done;
      trace(t, tph, 1);
      issue_next_tile();

      // ---- (2) tensor-core tile: 32 utterances x the 8 rows cluster rank `warp` owns, over this rank's K-share
      {
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const __nv_bfloat16* Wt = (const __nv_bfloat16*)(sW + (rp % NSLOT) * WSLOT);
        const int pw = kind == K_GATE ? PWG : kind == K_RES ? PWR : PWH;
        const int nt = warp & 3, kh = warp >> 2;   // owner rank whose 8 rows this warp computes, K-half of the share
        const int kbase = kh * ((kind == K_HEAD1 || kind == K_HEAD2) ? KH / 2 : KS / 2);
        const __nv_bfloat16* ap = Acur + (lane & 15) * PA + (lane >> 4) * 8 + kbase;
        const __nv_bfloat16* bp = Wt + (8 * nt + (lane & 7)) * pw + ((lane >> 3) & 1) * 8 + kbase;
        const int ksteps = (kind == K_HEAD1 || kind == K_HEAD2) ? KH / 32 : KS / 32;
        for (int ks = 0; ks < ksteps; ++ks) {
          unsigned b0, b1, a0, a1, a2, a3;
          ldmatrix_x2(b0, b1, bp + ks * 16);
          ldmatrix_x4(a0, a1, a2, a3, ap + ks * 16);
          mma_bf16(acc[0], a0, a1, a2, a3, b0, b1);
          ldmatrix_x4(a0, a1, a2, a3, ap + 16 * PA + ks * 16);
          mma_bf16(acc[1], a0, a1, a2, a3, b0, b1);
        }
        if (kind == K_GATE) {
          // past tap: x_l(t - k) rows that travelled with the weight tile; in the priming region past == present
          const __nv_bfloat16* pp = prime ? ap : (const __nv_bfloat16*)((const unsigned char*)Wt + WTILE) + (lane & 15) * PA + (lane >> 4) * 8 + kbase;
#pragma unroll
          for (int ks = 0; ks < KS / 32; ++ks) {
            unsigned b0, b1, a0, a1, a2, a3;
            ldmatrix_x2(b0, b1, bp + KS + ks * 16);
            ldmatrix_x4(a0, a1, a2, a3, pp + ks * 16);
            mma_bf16(acc[0], a0, a1, a2, a3, b0, b1);
            ldmatrix_x4(a0, a1, a2, a3, pp + 16 * PA + ks * 16);
            mma_bf16(acc[1], a0, a1, a2, a3, b0, b1);
          }
        }
        // partial tile -> owner rank `warp`: utterances (q4, q4+8, q4+16, q4+24) x rows (2*i4, 2*i4+1), fp16
        uint4 pk = make_uint4(pack_h2(acc[0][0], acc[0][1]), pack_h2(acc[0][2], acc[0][3]),
                              pack_h2(acc[1][0], acc[1][1]), pack_h2(acc[1][2], acc[1][3]));
        const unsigned dst = smem_u32(sRecv + (ab * NPART + rank * 2 + kh) * 32 + lane);
        st_async_v4(mapa(dst, nt), pk, mapa(bar, nt));
      }
      trace(t, tph, 2);

      // ---- (3) owner warp: sum the 8 partial tiles, finish the owned rows, publish
      if (warp == 0) {
        float ax[2][4];
        if (kind == K_GATE) aux_mma(l, t, ax);
        int bad = 0;
        {
          unsigned spins = 0; long long t0 = 0;
          const unsigned parity = (unsigned)(rp >> 1) & 1u;
          while (!mbar_try(bar, parity)) {
            if (spin_check(spins, t0)) { bad = 1; break; }
          }
        }
        if (bad) {
          if (lane == 0) *sAbort = 1;
        } else {
          float sum[4][2];
#pragma unroll
          for (int m = 0; m < 4; ++m) sum[m][0] = sum[m][1] = 0.f;
#pragma unroll
          for (int srcr = 0; srcr < NPART; ++srcr) {
            const uint4 qv = sRecv[(ab * NPART + srcr) * 32 + lane];
            float2 f;
            f = unpack_h2(qv.x); sum[0][0] += f.x; sum[0][1] += f.y;
            f = unpack_h2(qv.y); sum[1][0] += f.x; sum[1][1] += f.y;
            f = unpack_h2(qv.z); sum[2][0] += f.x; sum[2][1] += f.y;
            f = unpack_h2(qv.w); sum[3][0] += f.x; sum[3][1] += f.y;
          }
          if (kind == K_GATE) {
#pragma unroll
            for (int m = 0; m < 4; ++m) { sum[m][0] += ax[m >> 1][(m & 1) * 2]; sum[m][1] += ax[m >> 1][(m & 1) * 2 + 1]; }
            publish_gate(l, sum, (w512 + (unsigned)(2 * l)) & 1u);
          } else if (kind == K_RES) {
            const bool last = l == L - 1;
            const unsigned par = (w512 + (unsigned)(2 * l + 1)) & 1u;
            const float b0 = sBr[l * 8 + 2 * i4], b1 = sBr[l * 8 + 2 * i4 + 1];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int u = q4 + 8 * m;
              const float v0 = sum[m][0] + b0, v1 = sum[m][1] + b1;
              if (i4 < 2) {
                // residual projection + current input (qpnet.py:669 / 639); dead after the last block (C7)
                xc[m][0] += v0; xc[m][1] += v1;
                if (!last) st_strong_u32(p.v512 + ((size_t)s * UB + u) * 2 + i4, pack_tagged(xc[m][0], xc[m][1], par));
              } else if (i4 == 2) {
                sk[m][0] += v0; sk[m][1] += v1;
                if (last && !prime)
                  st_strong_u32(p.v256 + (size_t)s * UB + u, pack_tagged(fmaxf(sk[m][0], 0.f), fmaxf(sk[m][1], 0.f), 0u));
              }
            }
          } else if (kind == K_HEAD1) {
            if (i4 == 0) {
#pragma unroll
              for (int m = 0; m < 4; ++m)
                st_strong_u32(p.v256 + (size_t)s * UB + q4 + 8 * m,
                              pack_tagged(fmaxf(sum[m][0] + sBh[0], 0.f), fmaxf(sum[m][1] + sBh[1], 0.f), 1u));
            }
          } else {
            if (i4 == 0) {
              const unsigned par_t = (unsigned)t & 1u;
#pragma unroll
              for (int m = 0; m < 4; ++m)
                st_strong_v2(p.vlog + ((size_t)s * UB + q4 + 8 * m) * 2,
                             (__float_as_uint(sum[m][0] + sBh[8]) & ~1u) | par_t, (__float_as_uint(sum[m][1] + sBh[9]) & ~1u) | par_t);
            }
          }
        }
      } else if (warp >= 2 && kind == K_RES && l == 0) {
        // off the critical path: aux rows of the NEXT step, h_up[:, ta] = h[:, ta / U] * w[ta % U] + b (qpnet.py:143-158, 451)
        const int tn = t + 1;
        if (tn < g.max_steps) {
          const int ta = tn < 0 ? 0 : tn;
          const int f = ta / U, j = ta - f * U;
          const int wt = tid - 64;
          if (j == 0 && tn > 0) {
            for (int e = wt; e < UB * A; e += NT - 64) {
              int u = e / A, a = e - u * A;
              sHraw[u * HR + a] = u < B ? g.h[((size_t)u * A + a) * p.F + f] : 0.f;
            }
            asm volatile("bar.sync 1, 192;\n" ::: "memory");
          }
          const float w = g.up_w[j], bb = g.up_b[0];
          for (int e = wt; e < UB * A; e += NT - 64) {
            int u = e / A, a = e - u * A;
            sHaux[((tn & 1) * UB + u) * PH + a] = __float2bfloat16(sHraw[u * HR + a] * w + bb);
          }
        }
      }
      trace(t, tph, 3);
      ++rp;
    }

    // ================================================================ sampling: one warp per utterance
    if (!prime && warp == 1 && s < B) {
      if (TRACE && s == 0 && lane == 0 && t >= p.trace_step0 && t < p.trace_step0 + p.trace_nsteps)
        p.trace[((size_t)(t - p.trace_step0) * nphase + 2 * L + 2) * TRACE_EVENTS + 0] = clock64();
      const int u = s;
      const unsigned par_t = (unsigned)t & 1u;
      float v[8];
      int bad = 0;
      {
        unsigned pend = 0xF;
        unsigned spins = 0; long long t0 = 0;
        while (pend) {
          uint2 w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (pend & (1u << j)) w[j] = ld_strong_v2(p.vlog + ((size_t)(4 * lane + j) * UB + u) * 2);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((pend & (1u << j)) && (((w[j].x ^ par_t) | (w[j].y ^ par_t)) & 1u) == 0) {
              v[2 * j] = __uint_as_float(w[j].x); v[2 * j + 1] = __uint_as_float(w[j].y); pend &= ~(1u << j);
            }
          if (pend && spin_check(spins, t0)) { bad = 1; break; }
        }
      }
      if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) *sAbort = 1;
      } else {
        float mx = -INFINITY;
        int amax = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > mx) { mx = v[j]; amax = lane * 8 + j; }
        if (g.logits_out && t < g.n_samples[u]) {
          float4* lo = (float4*)(g.logits_out + ((size_t)u * g.max_steps + t) * Q + lane * 8);
          lo[0] = make_float4(v[0], v[1], v[2], v[3]);
          lo[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        float wmx = mx;
        int wam = amax;
        for (int o = 16; o; o >>= 1) {   // warp arg-max, first maximum wins
          float om = __shfl_xor_sync(0xffffffffu, wmx, o);
          int oa = __shfl_xor_sync(0xffffffffu, wam, o);
          if (om > wmx || (om == wmx && oa < wam)) { wmx = om; wam = oa; }
        }
        int sym;
        if (g.mode == QP_MODE_ARGMAX) {
          sym = wam;
        } else {   // softmax + inverse CDF on a uniform (qpnet.py:507-510)
          float local = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[j] = __expf(v[j] - wmx); local += v[j]; }
          float incl = local;
          for (int o = 1; o < 32; o <<= 1) {
            float nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
          }
          const float total = __shfl_sync(0xffffffffu, incl, 31);
          const float uu = g.uniforms ? g.uniforms[(long long)u * g.ld_uniforms + t] : philox_uniform(g.philox_seed, u, t);
          const float target = uu * total;
          float run = incl - local;
          int cnt = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) { run += v[j]; if (run <= target) ++cnt; }
          for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
          sym = min(cnt, Q - 1);
        }
        if (lane == 0) {
          if (t < g.n_samples[u]) g.out[(long long)u * g.ld_out + t] = sym;
          const int fed = g.force ? g.force[(long long)u * g.ld_force + t] : sym;
          st_strong_u32(p.vsym + u * 32, ((unsigned)fed & 0xFFFFu) | (par_t << 30));
        }
      }
      if (TRACE && s == 0 && lane == 0 && t >= p.trace_step0 && t < p.trace_step0 + p.trace_nsteps)
        p.trace[((size_t)(t - p.trace_step0) * nphase + 2 * L + 2) * TRACE_EVENTS + 3] = clock64();
    }
  }
done:
  cp_async_wait<0>();
  __syncthreads();
  cluster_sync();   // no CTA of the cluster leaves while a peer may still write into its shared memory
}

}  // namespace cl
}  // namespace qp

using namespace qp;

namespace qp {

// Launches the cluster generator.  Returns QP_OK, an error, or +1 when the device cannot keep the
// 16 clusters co-resident (the caller then uses the generic kernel).
int cl_generate(const QpArch* arch, const float* const* tensors_host, const QpGenerateArgs* a, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  cl::Plan p;
  size_t need = cl::make_plan(arch, a->B, a->F, a->M, ws, ws_bytes, &p);
  if (need > ws_bytes) return set_error(QP_EWORKSPACE, "generate: workspace %zu < %zu bytes", ws_bytes, need);
  const cl::SmemMap sm = cl::smem_map(p.L);
  QP_REQUIRE(sm.total <= 227 * 1024, "generate: %d bytes of shared memory needed", sm.total);
  const bool tr = getenv("QPNET_GEN_TRACE_STEP") != nullptr;
  if (tr) p.trace_step0 = atoi(getenv("QPNET_GEN_TRACE_STEP"));
  auto kern = tr ? cl::cl_gen_kernel<true> : cl::cl_gen_kernel<false>;
  QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl::NOWN); cfg.blockDim = dim3(cl::NT); cfg.dynamicSmemBytes = (size_t)sm.total; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl::CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  QP_CUDA(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
  if (ncl < cl::NOWN / cl::CL) return 1;
  cfg.numAttrs = 2;

  QP_CUDA(cudaMemsetAsync(p.status, 0, 256, st));
  QP_CUDA(cudaMemsetAsync(p.tagged_begin, 0xFF, p.tagged_bytes, st));   // every word starts with a stale tag
  QP_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(long long) * 8 * (2 * p.L + 3) * cl::TRACE_EVENTS, st));
  if (int e = upload_tensor_table(arch, tensors_host, p.tab, st)) return e;
  TensorMap tm = tensor_map(arch);
  cl::pack_kernel<<<148 * 8, 256, 0, st>>>(tm, p, p.tab);
  QP_LAUNCH_CHECK();
  cl::table_kernel<<<dim3(8, cl::NOWN), cl::Q, 0, st>>>(tm, p, p.tab);
  QP_LAUNCH_CHECK();
  GenArgsDev g;
  g.seed = a->seed; g.h = a->h; g.d = a->d; g.n_samples = a->n_samples;
  g.uniforms = a->uniforms; g.ld_uniforms = a->ld_uniforms; g.philox_seed = a->philox_seed;
  g.force = a->force; g.ld_force = a->ld_force;
  g.out = a->out; g.ld_out = a->ld_out; g.logits_out = a->logits_out;
  g.mode = a->mode; g.max_steps = a->max_steps; g.d_is_f64 = a->d_is_f64;
  g.causal_b = tensors_host[tm.causal_b()]; g.up_w = tensors_host[tm.up_w()]; g.up_b = tensors_host[tm.up_b()];
  QP_CUDA(cudaLaunchKernelEx(&cfg, kern, p, g));
  count_launch();
  return QP_OK;
}

size_t cl_workspace_bytes(const QpArch* arch, int B, int M) {
  cl::Plan p;
  return cl::make_plan(arch, B, 1, M, nullptr, 0, &p);
}

bool cl_supported(const QpArch* arch, int B) { return cl::supported(arch, B); }

int cl_trace_copy(const QpArch* arch, int B, int M, void* ws, size_t ws_bytes, long long* out_host, int n, cudaStream_t st) {
  cl::Plan p;
  cl::make_plan(arch, B, 1, M, ws, ws_bytes, &p);
  int total = 8 * (2 * p.L + 3) * cl::TRACE_EVENTS;
  if (n > total) n = total;
  QP_CUDA(cudaMemcpyAsync(out_host, p.trace, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
  QP_CUDA(cudaStreamSynchronize(st));
  return n;
}

}  // namespace qp
