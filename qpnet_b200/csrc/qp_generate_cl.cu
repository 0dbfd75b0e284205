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
//   * the past tap Wp.x(t-k) (qpnet.py:81-87, 457-502) is never recomputed on the critical path: when
//     x_l(t) is in shared memory the streaming warps also contract it with the PAST-tap weights and keep
//     the fp32 partial tile P'(t) in a per-CTA ring in global memory (the reference's FIFOs,
//     qpnet.py:388-393, 431-437, as partial sums); k steps later it is prefetched with cp.async behind
//     the block's weight tile and seeds the accumulators.  k = dil (fixed) or -round(-d[t]*dil)
//     (adaptive, qpnet.py:616-617 / 621-622), k == 0 -> oldest entry (caveat C4).
//   * block 0 needs no exchange at all: its input is the causal layer, a function of the last three
//     symbols, so its gate pre-activation is three table lookups (W.E folded at pack time, fp32).
//   * aux 1x1 (qpnet.py:663-664, 632-633) is a 6-MMA side product of the owner warp, computed while
//     the partial tiles are in flight.
// Exchange words carry a 1-bit epoch tag (LSB of the low bf16 / of the fp32 logit); consumers poll
// the data itself.  A CTA can publish version v+2 of its block while a slow consumer in ANOTHER cluster
// still reads v+1 (it cannot get further: v+2 needs every owner's v+1, published after that owner
// consumed v), so x / z versions alternate between two buffers and the tag is bit 1 of the version.  Every spin has a watchdog (QP_ETIMEOUT) and every exit goes through a cluster barrier.
#include <algorithm>
#include <type_traits>

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
constexpr int TRACE_EVENTS = 8;   // 0 start, 1 own pieces fresh, 2 barrier passed, 3 MMA done, 4 partials sent, 5 partials arrived, 6 published
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
  uint32_t* v512;         // [2][NOWN][UB][2]  x / z exchange: version v lives in buffer v & 1 with tag (v >> 1) & 1
  uint32_t* v256;         // [2][NOWN][UB]     buffer 0 relu(skip sum), buffer 1 relu(head-1); tag = step & 1
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
  p->v512 = ar.take<uint32_t>((size_t)2 * NOWN * UB * 2);
  p->v256 = ar.take<uint32_t>((size_t)2 * NOWN * UB);
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
  // roles: all 8 warps poll + run the tensor-core tile; warps 0-3 finish the owned rows (warp m: utterances
  // q4 + 8m), warps 4-7 stream weights / past rows and build the aux tile; warp 7 of CTA u < B samples utterance u.
  const bool finisher = warp < 4;
  const int fu = q4 + 8 * warp;                 // utterance this thread finishes (finisher warps)
  const int t128 = tid - 128;                   // index inside the streaming warps

  // ---- one-time staging ---------------------------------------------------------------
  for (int e = tid; e < sm.total / 16; e += NT) ((uint4*)smem)[e] = make_uint4(0, 0, 0, 0);
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

  // ---- weight tile stream (streaming warps only): one tile per MMA phase, NSLOT-deep ring, prefetch
  // distance 2.  Per step: tile 2l = res l, tile 2l-1 = gate l (l >= 1), then head-1, head-2 (not in
  // the priming step).  A gate tile carries the block's past-tap rows behind the weights.
  int pf_t = -1, pf_i = 0, pf_slot = 0;   // cursor of the next tile to fetch
  auto issue_next_tile = [&]() {
    if (pf_t < g.max_steps) {
      unsigned char* dst = sW + pf_slot * WSLOT;
      if (pf_i < 2 * L - 1) {
        if (pf_i & 1) {
          const int l = (pf_i + 1) >> 1;
          const __nv_bfloat16* src = p.Wgate + ((size_t)l * NOWN + s) * NR * 2 * KS;
#pragma unroll
          for (int j = 0; j < NR * 32 / 128; ++j) {
            const int e = t128 + 128 * j;
            cp_async16(dst + ((e >> 5) * PWG + (e & 31) * 8) * 2, src + (e >> 5) * 2 * KS + (e & 31) * 8);
          }
          if (pf_t >= 0) {
            // past-tap partial sums P'(t - k) = Wp . x_l(t - k) over this CTA's K-share, stored k steps ago in
            // MMA fragment order: piece (warp w, utterance group m, lane) = 4 floats of utterance (lane >> 2) + 8m
            const int ln = t128 & 31, w = t128 >> 5;
            const float* ringf = (const float*)p.ring[l];
            const int rmask = p.ring_size[l] - 1;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int u = (ln >> 2) + 8 * m;
              int k = p.dil[l];
              if (l >= p.nF) {   // pitch-dependent look-back of this step (qpnet.py:476-483, 613-624)
                k = 0;
                if (u < B)
                  k = g.d_is_f64 ? -gen_index_f64(((const double*)g.d)[(long long)u * ldd + pf_t], p.dil[l])
                                 : -gen_index_f32(((const float*)g.d)[(long long)u * ldd + pf_t], p.dil[l]);
                if (k <= 0 || k > p.depth[l]) k = p.depth[l];   // k == 0: python index 0 = oldest entry (C4)
              }
              const int slot = (pf_t - k) & rmask;
              cp_async16(dst + WTILE + ((w * 4 + m) * 32 + ln) * 16,
                         ringf + (((size_t)slot * NOWN + s) * 4 + w) * 512 + (m * 32 + ln) * 4);
            }
          }
        } else {
          const int l = pf_i >> 1;
          const __nv_bfloat16* src = p.Wres + ((size_t)l * NOWN + s) * NR * KS;
#pragma unroll
          for (int j = 0; j < NR * 16 / 128; ++j) {
            const int e = t128 + 128 * j;
            cp_async16(dst + ((e >> 4) * PWR + (e & 15) * 8) * 2, src + (e >> 4) * KS + (e & 15) * 8);
          }
        }
      } else {
        const int hd = pf_i - (2 * L - 1);
        const __nv_bfloat16* src = p.Whead + ((size_t)hd * NOWN + s) * NR * KH;
#pragma unroll
        for (int j = 0; j < NR * 8 / 128; ++j) {
          const int e = t128 + 128 * j;
          cp_async16(dst + ((e >> 3) * PWH + (e & 7) * 8) * 2, src + (e >> 3) * KH + (e & 7) * 8);
        }
      }
      ++pf_i;
      const int ntiles = pf_t < 0 ? 2 * L - 1 : 2 * L + 1;
      if (pf_i == ntiles) { pf_i = 0; ++pf_t; }
    }
    pf_slot = pf_slot == NSLOT - 1 ? 0 : pf_slot + 1;
    cp_async_commit();
  };
  if (!finisher) { issue_next_tile(); issue_next_tile(); }

  // ---- finisher state: this thread's utterance fu, rows (2*i4, 2*i4+1) of the owned 8 -------------
  float xc0 = 0.f, xc1 = 0.f;          // fp32 residual carry (lanes i4 < 2)
  float sk0 = 0.f, sk1 = 0.f;          // skip accumulators (lanes i4 == 2)
  int sy_c = half, sy_p1 = half, sy_p2 = half;   // lane u: s(t-1), s(t-2), s(t-3) of utterance u
  int rp = 0;                          // MMA phase counter: activation / receive / barrier double buffering
  int cur_slot = 0;                    // weight slot of the current phase

  auto spin_check = [&](unsigned& spins, long long& t0) -> bool {
    if ((++spins & 1023u) != 0) return false;
    if (t0 == 0) t0 = clock64();
    if (*((volatile int32_t*)p.status) != 0) return true;
    if (clock64() - t0 > GEN_TIMEOUT_CYCLES) { atomicExch(p.status, QP_ETIMEOUT); return true; }
    return false;
  };

  // aux 1x1 of the owned 8 gate rows for this warp's 8 utterances (its half of one m16 tile): 3 MMAs
  // (qpnet.py:663-664 / 632-633).  Returns the (row 2*i4, row 2*i4+1) pair of utterance fu.
  auto aux_pair = [&](int l, int t, float& a0_, float& a1_) {
    float ax[4] = {0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* hp = sHaux + ((t & 1) * UB + 16 * (warp >> 1) + (lane & 15)) * PH + (lane >> 4) * 8;
    const __nv_bfloat16* vp = sVaux + (l * 8 + (lane & 7)) * PH + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < AP / 16; ++ks) {
      unsigned b0, b1, a0, a1, a2, a3;
      ldmatrix_x2(b0, b1, vp + ks * 16);
      ldmatrix_x4(a0, a1, a2, a3, hp + ks * 16);
      mma_bf16(ax, a0, a1, a2, a3, b0, b1);
    }
    a0_ = (warp & 1) ? ax[2] : ax[0];
    a1_ = (warp & 1) ? ax[3] : ax[1];
  };
  // gate non-linearity + publication of (utterance fu, owned channel i4)
  // word of (owner block s, utterance fu) in the buffer of x / z version `ver`
  auto v512_word = [&](unsigned ver) -> uint32_t* { return p.v512 + ((size_t)(ver & 1u) * NOWN * UB + (size_t)s * UB + fu) * 2; };
  auto publish_gate = [&](int l, float pre_s, float pre_t, unsigned ver) {
    const float z = fast_sigmoid(pre_s + sBg[l * 8 + 2 * i4]) * fast_tanh(pre_t + sBg[l * 8 + 2 * i4 + 1]);
    const float zn = __shfl_xor_sync(0xffffffffu, z, 1);
    if (!(i4 & 1)) st_strong_u32(v512_word(ver) + (i4 >> 1), pack_tagged(z, zn, (ver >> 1) & 1u));
  };

  // =========================================================================== time loop
  for (int t = -1; t < g.max_steps; ++t) {
    const bool prime = t < 0;
    const unsigned w512 = (unsigned)(t + 1) * (unsigned)(2 * L - 1);   // 512-vector writes before this step

    // ================================================================ block 0 gate: symbols -> tables
    trace(t, 0, 0);
    if (finisher) {
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
        float a0_, a1_;
        aux_pair(0, t, a0_, a1_);
        const int c_ = __shfl_sync(0xffffffffu, sy_c, fu), a_ = __shfl_sync(0xffffffffu, sy_p1, fu), b_ = __shfl_sync(0xffffffffu, sy_p2, fu);
        const float2 ta = *(const float2*)(sT0 + ((0 * Q + c_) * 8 + 2 * i4));
        const float2 tb = *(const float2*)(sT0 + ((1 * Q + a_) * 8 + 2 * i4));
        const float2 tc = *(const float2*)(sT0 + ((2 * Q + b_) * 8 + 2 * i4));
        if (i4 < 2) {   // fp32 residual stream of the owned channels restarts from the causal layer
          const float2 e0 = *(const float2*)(sEo + ((0 * Q + a_) * 4 + 2 * i4));
          const float2 e1 = *(const float2*)(sEo + ((1 * Q + c_) * 4 + 2 * i4));
          xc0 = e0.x + e1.x; xc1 = e0.y + e1.y;
        }
        sk0 = sk1 = 0.f;
        publish_gate(0, ta.x + tb.x + tc.x + a0_, ta.y + tb.y + tc.y + a1_, w512);
      }
    }
    trace(t, 0, 6);

    // ================================================================ MMA phases
    // One phase = poll the K-share -> tensor-core tiles -> partial tiles to the owners -> finish + publish.
    // KIND is a compile-time constant so every phase type gets straight-line code.
    auto phase = [&](auto kc, const int l) -> bool {
      constexpr int KIND = decltype(kc)::value;
      constexpr bool is512 = KIND == K_GATE || KIND == K_RES;
      const int tph = KIND == K_GATE ? 2 * l : KIND == K_RES ? 2 * l + 1 : KIND == K_HEAD1 ? 2 * L : 2 * L + 1;
      trace(t, tph, 0);
      const int ab = rp & 1;
      __nv_bfloat16* Acur = sAcur + ab * UB * PA;
      const unsigned bar = smem_u32(&sBars[ab]);
      if (tid == 0) mbar_expect_tx(bar, NPART * 512);
      // aux 1x1 of the owned rows needs nothing from this phase's exchange: run it while the input is in flight
      float a0_ = 0.f, a1_ = 0.f;
      if (KIND == K_GATE && finisher) aux_pair(l, t, a0_, a1_);

      // ---- (1) poll this rank's K-share of the input vector, stage it as the MMA A tile
      int fail = 0;
      if (is512) {
        // res l reads z_l (write index 2l), gate l reads x_l (write index 2l - 1)
        const unsigned ver = w512 + (unsigned)(KIND == K_RES ? 2 * l : 2 * l - 1);
        const unsigned par = (ver >> 1) & 1u;
        // share = owner blocks 32*rank .. 32*rank+31 (16 pieces of 16 bytes each): two pieces per thread
        const uint4* src = (const uint4*)p.v512 + (size_t)(ver & 1u) * NOWN * 16 + (size_t)(32 * rank) * 16 + tid;
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
        const int u0 = 2 * (tid & 15), col0 = 4 * (tid >> 4);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          *(uint2*)(Acur + u0 * PA + col0 + 64 * j) = make_uint2(v[j].x, v[j].y);
          *(uint2*)(Acur + (u0 + 1) * PA + col0 + 64 * j) = make_uint2(v[j].z, v[j].w);
        }
      } else {
        const unsigned par = (unsigned)t & 1u;
        // share = owner blocks 32*rank .. 32*rank+31 (8 pieces each): one piece per thread
        const uint4* src = (const uint4*)p.v256 + (size_t)(KIND == K_HEAD1 ? 0 : 1) * NOWN * 8 + (size_t)(32 * rank) * 8 + tid;
        uint4 w;
        unsigned spins = 0; long long t0 = 0;
        while (true) {
          w = ld_strong_v4(src);
          if (fresh4(w, par)) break;
          if (spin_check(spins, t0)) { fail = 1; break; }
        }
        const int uh = 4 * (tid & 7), col = 2 * (tid >> 3);
        *(unsigned*)(Acur + uh * PA + col) = w.x;
        *(unsigned*)(Acur + (uh + 1) * PA + col) = w.y;
        *(unsigned*)(Acur + (uh + 2) * PA + col) = w.z;
        *(unsigned*)(Acur + (uh + 3) * PA + col) = w.w;
      }
      trace(t, tph, 1);
      cp_async_wait<1>();   // streaming warps: this phase's weight tile (and its past partial sums) have landed
      if (__syncthreads_or(fail | *sAbort)) return true;
      trace(t, tph, 2);

      // ---- tensor-core tiles, shared by both roles: warp (w4 = warp & 3) contracts 32 utterances x the 16 rows
      // owner ranks 2*(w4 & 1), +1 finish over K-half (w4 >> 1).  2 x 2 register blocking halves the
      // shared-memory operand traffic of 1 x 2.
      const __nv_bfloat16* Wt = (const __nv_bfloat16*)(sW + cur_slot * WSLOT);
      const int w4 = warp & 3, ntp = w4 & 1, kh = w4 >> 1;
      const int lrow = (lane & 15) * PA + (lane >> 4) * 8;
      // B operand through ldmatrix.x4: lanes 0-7 rows 0-7 k lo, 8-15 rows 0-7 k hi, 16-23 rows 8-15 k lo, 24-31 rows 8-15 k hi
      const int brow = 16 * ntp + (lane & 7) + ((lane >> 4) << 3), bcol = ((lane >> 3) & 1) * 8;
      constexpr int PW = KIND == K_GATE ? PWG : KIND == K_RES ? PWR : PWH;
      constexpr int KHALF = is512 ? KS / 2 : KH / 2;
      const __nv_bfloat16* ap = Acur + lrow + kh * KHALF;
      const __nv_bfloat16* bp = Wt + brow * PW + bcol + kh * KHALF;
      auto kloop = [&](float (&acc)[2][2][4], const __nv_bfloat16* bq) {
#pragma unroll
        for (int ks = 0; ks < KHALF / 16; ++ks) {
          unsigned b0, b1, b2, b3, a0, a1, a2, a3, c0, c1, c2, c3;
          ldmatrix_x4(b0, b1, b2, b3, bq + ks * 16);   // rows 0-7 (k lo, k hi), rows 8-15 (k lo, k hi)
          ldmatrix_x4(a0, a1, a2, a3, ap + ks * 16);
          ldmatrix_x4(c0, c1, c2, c3, ap + 16 * PA + ks * 16);
          mma_bf16(acc[0][0], a0, a1, a2, a3, b0, b1);
          mma_bf16(acc[1][0], a0, a1, a2, a3, b2, b3);
          mma_bf16(acc[0][1], c0, c1, c2, c3, b0, b1);
          mma_bf16(acc[1][1], c0, c1, c2, c3, b2, b3);
        }
      };
      // fragment <-> ring piece: utterance group m = 2*mt + (f >> 1) holds {acc[0][mt][2(f>>1)..+1], acc[1][mt][2(f>>1)..+1]}
      float acc[2][2][4];   // [owner of the pair][m tile][fragment]
#pragma unroll
      for (int a_ = 0; a_ < 2; ++a_)
#pragma unroll
        for (int b_ = 0; b_ < 2; ++b_) acc[a_][b_][0] = acc[a_][b_][1] = acc[a_][b_][2] = acc[a_][b_][3] = 0.f;

      if (finisher) {
        // ---- (2) current tap (+ the past-tap partial sums P'(t - k) that travelled with the weight tile)
        if (KIND == K_GATE) {
          if (!prime) {
            const float4* pin = (const float4*)((const unsigned char*)Wt + WTILE) + w4 * 128 + lane;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const float4 q = pin[m * 32];
              acc[0][m >> 1][(m & 1) * 2] = q.x; acc[0][m >> 1][(m & 1) * 2 + 1] = q.y;
              acc[1][m >> 1][(m & 1) * 2] = q.z; acc[1][m >> 1][(m & 1) * 2 + 1] = q.w;
            }
          } else {
            kloop(acc, bp + KS);   // priming region: the past equals the present (qpnet.py:355-440)
          }
        }
        kloop(acc, bp);
        trace(t, tph, 3);
        // partial tiles -> owner ranks: utterances (q4, q4+8, q4+16, q4+24) x rows (2*i4, 2*i4+1), fp16
#pragma unroll
        for (int a_ = 0; a_ < 2; ++a_) {
          const int nt = 2 * ntp + a_;
          uint4 pk = make_uint4(pack_h2(acc[a_][0][0], acc[a_][0][1]), pack_h2(acc[a_][0][2], acc[a_][0][3]),
                                pack_h2(acc[a_][1][0], acc[a_][1][1]), pack_h2(acc[a_][1][2], acc[a_][1][3]));
          const unsigned dst = smem_u32(sRecv + (ab * NPART + rank * 2 + kh) * 32 + lane);
          st_async_v4(mapa(dst, nt), pk, mapa(bar, nt));
        }
        trace(t, tph, 4);

        // ---- (3) sum the 8 partial tiles of (utterance fu, rows 2*i4, 2*i4+1), finish, publish
        int bad = 0;
        {
          unsigned spins = 0; long long t0 = 0;
          const unsigned parity = (unsigned)(rp >> 1) & 1u;
          while (!mbar_try(bar, parity)) {
            if (spin_check(spins, t0)) { bad = 1; break; }
          }
        }
        trace(t, tph, 5);
        if (bad) {
          if (lane == 0) *sAbort = 1;
        } else {
          float s0 = 0.f, s1 = 0.f;
          const unsigned* rw = (const unsigned*)(sRecv + ab * NPART * 32 + lane) + warp;
#pragma unroll
          for (int srcr = 0; srcr < NPART; ++srcr) {
            const float2 f = unpack_h2(rw[srcr * 32 * 4]);
            s0 += f.x; s1 += f.y;
          }
          if (KIND == K_GATE) {
            publish_gate(l, s0 + a0_, s1 + a1_, w512 + (unsigned)(2 * l));
          } else if (KIND == K_RES) {
            const bool last = l == L - 1;
            const float v0 = s0 + sBr[l * 8 + 2 * i4], v1 = s1 + sBr[l * 8 + 2 * i4 + 1];
            if (i4 < 2) {
              // residual projection + current input (qpnet.py:669 / 639); dead after the last block (C7)
              xc0 += v0; xc1 += v1;
              if (!last) {
                const unsigned ver = w512 + (unsigned)(2 * l + 1);
                st_strong_u32(v512_word(ver) + i4, pack_tagged(xc0, xc1, (ver >> 1) & 1u));
              }
            } else if (i4 == 2) {
              sk0 += v0; sk1 += v1;
              if (last && !prime)
                st_strong_u32(p.v256 + (size_t)s * UB + fu, pack_tagged(fmaxf(sk0, 0.f), fmaxf(sk1, 0.f), (unsigned)t & 1u));
            }
          } else if (KIND == K_HEAD1) {
            if (i4 == 0)
              st_strong_u32(p.v256 + (size_t)(NOWN + s) * UB + fu, pack_tagged(fmaxf(s0 + sBh[0], 0.f), fmaxf(s1 + sBh[1], 0.f), (unsigned)t & 1u));
          } else {
            if (i4 == 0) {
              const unsigned par_t = (unsigned)t & 1u;
              st_strong_v2(p.vlog + ((size_t)s * UB + fu) * 2, (__float_as_uint(s0 + sBh[8]) & ~1u) | par_t,
                           (__float_as_uint(s1 + sBh[9]) & ~1u) | par_t);
            }
          }
        }
        trace(t, tph, 6);
      } else {
        // ---- (3') streaming warps: next weight tile; for a gate, the past-tap partial sums P'(t) = Wp . x_l(t)
        // later steps will need; once per step the aux rows of the NEXT step.
        issue_next_tile();
        if (KIND == K_GATE) {
          kloop(acc, bp + KS);
          float* ringf = (float*)p.ring[l];
          const int rs = p.ring_size[l];
          const size_t slot_f4 = (size_t)NOWN * 512;   // float4 per ring slot: 128 CTAs x 4 warps x 4 groups x 32 lanes
          float4* r0 = (float4*)(ringf + ((size_t)s * 4 + w4) * 512) + lane;
          const int sl0 = prime ? 0 : (t & (rs - 1)), sl1 = prime ? rs : sl0 + 1;
          for (int sl = sl0; sl < sl1; ++sl) {   // the priming step fills the whole ring with the constant of the pad region
            float4* r1 = r0 + sl * slot_f4;
#pragma unroll
            for (int m = 0; m < 4; ++m)
              r1[m * 32] = make_float4(acc[0][m >> 1][(m & 1) * 2], acc[0][m >> 1][(m & 1) * 2 + 1],
                                       acc[1][m >> 1][(m & 1) * 2], acc[1][m >> 1][(m & 1) * 2 + 1]);
          }
        }
        if (KIND == K_RES && l == 0) {
          // h_up[:, ta] = h[:, ta / U] * w[ta % U] + b (qpnet.py:143-158, 451)
          const int tn = t + 1;
          if (tn < g.max_steps) {
            const int ta = tn < 0 ? 0 : tn;
            const int f = ta / U, j = ta - f * U;
            if (j == 0 && tn > 0) {
              for (int e = t128; e < UB * A; e += 128) {
                int u = e / A, a = e - u * A;
                sHraw[u * HR + a] = u < B ? g.h[((size_t)u * A + a) * p.F + f] : 0.f;
              }
              asm volatile("bar.sync 1, 128;\n" ::: "memory");
            }
            const float w = g.up_w[j], bb = g.up_b[0];
            for (int e = t128; e < UB * A; e += 128) {
              int u = e / A, a = e - u * A;
              sHaux[((tn & 1) * UB + u) * PH + a] = __float2bfloat16(sHraw[u * HR + a] * w + bb);
            }
          }
        }
      }
      ++rp;
      cur_slot = cur_slot == NSLOT - 1 ? 0 : cur_slot + 1;
      return false;
    };
    {
      bool stop = false;
      for (int l = 0; l < L && !stop; ++l) {
        stop = phase(std::integral_constant<int, K_RES>(), l);
        if (!stop && l + 1 < L) stop = phase(std::integral_constant<int, K_GATE>(), l + 1);
      }
      if (!stop && !prime) {
        stop = phase(std::integral_constant<int, K_HEAD1>(), 0);
        if (!stop) stop = phase(std::integral_constant<int, K_HEAD2>(), 0);
      }
      if (stop) goto done;
    }

    // ================================================================ sampling: one warp per utterance
    if (!prime && warp == 7 && s < B) {
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
          const float uu = g.uniforms ? g.uniforms[(long long)u * g.ld_uniforms + t] : philox_uniform(g.philox_seed, g.utt_ids ? (unsigned)g.utt_ids[u] : (unsigned)u, t);
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
        p.trace[((size_t)(t - p.trace_step0) * nphase + 2 * L + 2) * TRACE_EVENTS + 6] = clock64();
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
  cfg.numAttrs = getenv("QPNET_GEN_NOCOOP") ? 1 : 2;   // (profilers that cannot replay cooperative cluster launches)

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
  g.force = a->force; g.ld_force = a->ld_force; g.utt_ids = a->utt_ids;
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
