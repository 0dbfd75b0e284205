// fp32 SIMT segmented GEMM with fused epilogues -- the exact-arithmetic path of the
// teacher-forced stack (used for tight-tolerance parity and by backward).
//
//   acc[r, n] = sum_seg sum_k  A_seg[ src_row_seg(r), k ] * W(n, koff_seg + k)
//
// A is assembled on the fly from up to three K-segments of time-major activations
// (past tap rows, current tap rows, aux rows): the 2-tap causal / pitch-adaptive conv is
// a GEMM whose A rows are *gathered*, never materialised (qpnet.py:295-298,657-666).
#pragma once
#include <algorithm>

#include "qp_common.cuh"

namespace qp {

struct Seg {
  const float* base;   // [B][rows][ld]
  int64_t bstride;     // elements between batch elements
  int ld;              // row pitch
  const int* rowmap;   // optional [B][n_rows]: explicit source row (adaptive past tap)
  int row_off;         // else source row = r + row_off ; src < 0 or >= src_rows reads zero
  int src_rows;        // rows available in base per batch element
  int K;               // segment width
  int relu;            // apply relu while loading
};

enum Epi { EPI_PLAIN = 0, EPI_GATE, EPI_RESSKIP, EPI_DGATE, EPI_DX };

struct GemmArgs {
  Seg seg[3];
  int nseg;
  const float* W;  // NT: W[n*ldw + k]   (w_kn = 0)   NN: W[k*ldw + n]  (w_kn = 1)
  int ldw, w_kn;
  const float* bias;  // [N] or null
  int B, n_rows, N;
  int n_begin;  // first output column computed (skip the dead res rows of the last block)
  // ---- epilogue operands (meaning depends on Epi) ----
  float* out;  int64_t out_bstride;  int ldo;     // PLAIN: out[r][n]; GATE: z; RESSKIP: x_next; DGATE: dgate
  const float* mask; int64_t mask_bstride; int ldmask;  // PLAIN: multiply by (mask[r][n] > 0)
  float* gsave; int64_t gsave_bstride;            // GATE: optional (n_rows, 2C) sigmoid/tanh outputs; DGATE: input
  const float* resid; int64_t resid_bstride; int ldresid; int resid_off;  // RESSKIP: + x_cur[r + off][n]; DX: dXnext
  float* skip; int64_t skip_bstride; int skip_row0; int skip_accum; int C;  // RESSKIP
  // DX: scatter dgate * Wg into dX (past rows, atomics), dX (current rows) and dHup
  float* dx; int64_t dx_bstride; const int* dx_rowmap; int dx_past_off; int dx_cur_off; int dx_rows;
  float* dh; int64_t dh_bstride; int dh_off; int A;
};

constexpr int GM = 64, GN = 64, GK = 16;   // (GK = 32 was measured slower: 36 KB of tiles per block halve the residency)
constexpr int LD_IT = GM * GK / 256;   // elements of a 64 x GK tile each of the 256 threads loads
constexpr int GPAD = 8;   // tile pitch GM + 8: the TF32 fragment loads (k = lane % 4, m = lane / 4) hit 32 distinct banks

// ---- TF32 tensor-core inner product (backward of the bf16 training path): the same 64 x 64 x 16 shared tiles,
// mma.sync.m16n8k8 with fp32 accumulation; warp w owns rows 16 (w & 3).., columns 32 (w >> 2)...  The accumulators
// go through a shared 64 x 64 tile back into the (ty, tx) 4 x 4 register layout, so every epilogue is shared with the
// exact fp32 SIMT path.
__device__ __forceinline__ unsigned to_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tile_mma_tf32(const float (*As)[GM + GPAD], const float (*Ws)[GN + GPAD], float (&c)[4][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3, m0 = 16 * (warp & 3), n0 = 32 * (warp >> 2);
#pragma unroll
  for (int ks = 0; ks < GK; ks += 8) {
    const unsigned a0 = to_tf32(As[ks + t][m0 + g]), a1 = to_tf32(As[ks + t][m0 + g + 8]);
    const unsigned a2 = to_tf32(As[ks + t + 4][m0 + g]), a3 = to_tf32(As[ks + t + 4][m0 + g + 8]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const unsigned b0 = to_tf32(Ws[ks + t][n0 + nt * 8 + g]), b1 = to_tf32(Ws[ks + t + 4][n0 + nt * 8 + g]);
      mma_tf32(c[nt], a0, a1, a2, a3, b0, b1);
    }
  }
}
// fragment layout -> (ty, tx) 4 x 4 layout through shared memory (call from all 256 threads)
__device__ __forceinline__ void frag_to_blocked(float (*Cs)[GN + 4], const float (&c)[4][4], float (&acc)[4][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3, m0 = 16 * (warp & 3), n0 = 32 * (warp >> 2);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    Cs[m0 + g][n0 + nt * 8 + 2 * t] = c[nt][0];
    Cs[m0 + g][n0 + nt * 8 + 2 * t + 1] = c[nt][1];
    Cs[m0 + g + 8][n0 + nt * 8 + 2 * t] = c[nt][2];
    Cs[m0 + g + 8][n0 + nt * 8 + 2 * t + 1] = c[nt][3];
  }
  __syncthreads();
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = Cs[ty * 4 + i][tx * 4 + j];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

template <int EPI, bool TC>
__global__ void __launch_bounds__(256, TC ? 4 : 2) gemm_f32_kernel(GemmArgs a) {
  __shared__ __align__(16) float As[GK][GM + GPAD];
  __shared__ __align__(16) float Ws[GK][GN + GPAD];
  __shared__ float Cs[TC ? GM : 1][GN + 4];
  float cfr[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cfr[i][j] = 0.f;
  const int b = blockIdx.z;
  const int r0 = blockIdx.x * GM;
  const int n0 = a.n_begin + blockIdx.y * GN;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  int koff = 0;
  // load slots of this thread, one float4 each per K-slice: A tile row tid / 4, k (tid % 4) * 4 .. +3; W tile: k tid / 16,
  // n (tid % 16) * 4 .. +3 when n is contiguous (w_kn), else n tid / 4, k (tid % 4) * 4 .. +3
  static_assert(GK == 16 && GM == 64 && GN == 64, "one float4 per thread per tile");
  const int ar = tid / 4, ak4 = (tid % 4) * 4;
  const int wk = a.w_kn ? tid / 16 : (tid % 4) * 4, wn = a.w_kn ? (tid % 16) * 4 : tid / 4;
  const bool w_vec = (a.ldw & 3) == 0 && (((size_t)a.W) & 15) == 0;
  for (int s = 0; s < a.nseg; ++s) {
    const Seg sg = a.seg[s];
    // the source row of this thread's tile row is fixed for the whole segment (gathered past tap: one lookup)
    const float* arow = nullptr;
    {
      const int r = r0 + ar;
      if (r < a.n_rows) {
        const int src = sg.rowmap ? sg.rowmap[(int64_t)b * a.n_rows + r] : r + sg.row_off;
        if (src >= 0 && src < sg.src_rows) arow = sg.base + (int64_t)b * sg.bstride + (int64_t)src * sg.ld;
      }
    }
    const bool a_vec = (sg.ld & 3) == 0 && (sg.bstride & 3) == 0 && (((size_t)sg.base) & 15) == 0;
    auto load_a = [&](int k0) -> float4 {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = k0 + ak4;
      if (arow && k < sg.K) {
        if (a_vec && k + 3 < sg.K) v = *(const float4*)(arow + k);
        else {
          v.x = arow[k];
          if (k + 1 < sg.K) v.y = arow[k + 1];
          if (k + 2 < sg.K) v.z = arow[k + 2];
          if (k + 3 < sg.K) v.w = arow[k + 3];
        }
        if (sg.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      }
      return v;
    };
    auto load_w = [&](int k0) -> float4 {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.w_kn) {      // W[k][n]: four consecutive n
        const int k = k0 + wk, n = n0 + wn;
        if (k < sg.K && n < a.N) {
          const float* q = a.W + (int64_t)(koff + k) * a.ldw + n;
          if (w_vec && n + 3 < a.N && (n & 3) == 0) v = *(const float4*)q;
          else { v.x = q[0]; if (n + 1 < a.N) v.y = q[1]; if (n + 2 < a.N) v.z = q[2]; if (n + 3 < a.N) v.w = q[3]; }
        }
      } else {           // W[n][k]: four consecutive k
        const int k = k0 + wk, n = n0 + wn;
        if (k < sg.K && n < a.N) {
          const float* q = a.W + (int64_t)n * a.ldw + koff + k;
          if (w_vec && k + 3 < sg.K && ((koff + k) & 3) == 0) v = *(const float4*)q;
          else { v.x = q[0]; if (k + 1 < sg.K) v.y = q[1]; if (k + 2 < sg.K) v.z = q[2]; if (k + 3 < sg.K) v.w = q[3]; }
        }
      }
      return v;
    };
    // software pipeline inside a segment: the next K-slice is in registers while the current one is contracted
    float4 av = load_a(0), wv = load_w(0);
    for (int k0 = 0; k0 < sg.K; k0 += GK) {
      As[ak4][ar] = av.x; As[ak4 + 1][ar] = av.y; As[ak4 + 2][ar] = av.z; As[ak4 + 3][ar] = av.w;
      if (a.w_kn) *(float4*)&Ws[wk][wn] = wv;
      else { Ws[wk][wn] = wv.x; Ws[wk + 1][wn] = wv.y; Ws[wk + 2][wn] = wv.z; Ws[wk + 3][wn] = wv.w; }
      __syncthreads();
      if (k0 + GK < sg.K) { av = load_a(k0 + GK); wv = load_w(k0 + GK); }
      if (TC) {
        tile_mma_tf32(As, Ws, cfr);
      } else {
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
          float pa[4], pw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pa[i] = As[kk][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) pw[j] = Ws[kk][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pa[i], pw[j], acc[i][j]);
        }
      }
      __syncthreads();
    }
    koff += sg.K;
  }
  if (TC) frag_to_blocked(Cs, cfr, acc);

  // ------------------------------------------------------------------ epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= a.n_rows) continue;
    const int nb = n0 + tx * 4;
    if (EPI == EPI_PLAIN) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = nb + j;
        if (n >= a.N) continue;
        float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
        if (a.mask && !(a.mask[(int64_t)b * a.mask_bstride + (int64_t)r * a.ldmask + n] > 0.f)) v = 0.f;
        a.out[(int64_t)b * a.out_bstride + (int64_t)r * a.ldo + n] = v;
      }
    } else if (EPI == EPI_GATE) {
      // columns are (sigmoid, tanh) pairs: n = 2c + g   (qpnet.py:665-666 / 634-635)
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        int n = nb + j;
        if (n >= a.N) continue;
        float sg_ = sigmoidf_(acc[i][j] + a.bias[n]);
        float th_ = tanhf(acc[i][j + 1] + a.bias[n + 1]);
        a.out[(int64_t)b * a.out_bstride + (int64_t)r * a.ldo + (n >> 1)] = sg_ * th_;
        if (a.gsave) {
          float* g = a.gsave + (int64_t)b * a.gsave_bstride + (int64_t)r * a.N + n;
          g[0] = sg_;
          g[1] = th_;
        }
      }
    } else if (EPI == EPI_RESSKIP) {
      // rows [0,C): residual projection + x_cur (qpnet.py:668-669); rows [C,C+S): skip (667)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = nb + j;
        if (n >= a.N) continue;
        float v = acc[i][j] + a.bias[n];
        if (n < a.C) {
          v += a.resid[(int64_t)b * a.resid_bstride + (int64_t)(r + a.resid_off) * a.ldresid + n];
          a.out[(int64_t)b * a.out_bstride + (int64_t)r * a.ldo + n] = v;
        } else if (r >= a.skip_row0) {
          float* p = a.skip + (int64_t)b * a.skip_bstride + (int64_t)(r - a.skip_row0) * (a.N - a.C) + (n - a.C);
          *p = a.skip_accum ? *p + v : v;
        }
      }
    } else if (EPI == EPI_DGATE) {
      // acc = dz[r][c] (N = C).  dgate[2c] = dz*th*sg*(1-sg);  dgate[2c+1] = dz*sg*(1-th^2)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = nb + j;
        if (c >= a.N) continue;
        const float* g = a.gsave + (int64_t)b * a.gsave_bstride + (int64_t)r * (2 * a.N) + 2 * c;
        float sg_ = g[0], th_ = g[1], dz = acc[i][j];
        float* o = a.out + (int64_t)b * a.out_bstride + (int64_t)r * a.ldo + 2 * c;
        o[0] = dz * th_ * sg_ * (1.f - sg_);
        o[1] = dz * sg_ * (1.f - th_ * th_);
      }
    } else if (EPI == EPI_DX) {
      // acc = (dgate * Wg)[r][k], k over [past C | current C | aux Ap]
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int k = nb + j;
        if (k >= a.N) continue;
        float v = acc[i][j];
        if (k < a.C) {
          int src = a.dx_rowmap ? a.dx_rowmap[(int64_t)b * a.n_rows + r] : r + a.dx_past_off;
          if (src >= 0 && src < a.dx_rows) atomicAdd(a.dx + (int64_t)b * a.dx_bstride + (int64_t)src * a.C + k, v);
        } else if (k < 2 * a.C) {
          int c = k - a.C;
          if (a.resid) v += a.resid[(int64_t)b * a.resid_bstride + (int64_t)r * a.ldresid + c];
          atomicAdd(a.dx + (int64_t)b * a.dx_bstride + (int64_t)(r + a.dx_cur_off) * a.C + c, v);
        } else if (k - 2 * a.C < a.A) {
          float* p = a.dh + (int64_t)b * a.dh_bstride + (int64_t)(r + a.dh_off) * a.A + (k - 2 * a.C);
          *p += v;
        }
      }
    }
  }
}

template <int EPI>
inline int launch_gemm(const GemmArgs& a, cudaStream_t stream, bool tc = false) {
  if (a.n_rows <= 0 || a.B <= 0 || a.N - a.n_begin <= 0) return QP_OK;
  dim3 grid((a.n_rows + GM - 1) / GM, (a.N - a.n_begin + GN - 1) / GN, a.B);
  if (tc) gemm_f32_kernel<EPI, true><<<grid, 256, 0, stream>>>(a);
  else gemm_f32_kernel<EPI, false><<<grid, 256, 0, stream>>>(a);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// weight-gradient GEMM:  out[i][j] (+)= sum_b sum_r  P[b][r][i] * Q_seg[b][src(r)][j - joff]
// P: plain rows with up to two column segments (dXnext | dSkip);  Q: gathered segments.
// ---------------------------------------------------------------------------------------
struct WgradArgs {
  Seg p[2]; int np;     // column segments of P (rows of the output)
  Seg q[3]; int nq;     // column segments of Q (columns of the output)
  int B, n_rows;
  int I, J;             // output dims
  float* out; int ldo;  // out[i*ldo + j], overwritten
  float* colsum;        // optional [I]: sum_r P[r][i]  (bias gradient), overwritten
  int chunk;            // rows per split (set by launch_wgrad): blockIdx.z contracts rows [z*chunk, (z+1)*chunk) and, when
                        // there is more than one split, accumulates into the zeroed output with atomics
};

__device__ __forceinline__ float seg_load(const Seg& sg, int b, int r, int k, int n_rows) {
  int src = sg.rowmap ? sg.rowmap[(int64_t)b * n_rows + r] : r + sg.row_off;
  if (src < 0 || src >= sg.src_rows) return 0.f;
  float v = sg.base[(int64_t)b * sg.bstride + (int64_t)src * sg.ld + k];
  return sg.relu ? fmaxf(v, 0.f) : v;
}

// One load slot of a thread: four consecutive output rows / columns, i.e. four consecutive columns of ONE segment of P
// or Q (segment widths are multiples of 4 except the aux segment, which takes the scalar path), one tile row.
struct WSlot {
  const float* base;   // segment base + first column (nullptr: outside the output, loads zero)
  const int* rowmap;
  int64_t bstride;
  int ld, row_off, src_rows, relu;
  int valid;           // columns of the four that exist in the segment
  int vec;             // 16-byte aligned rows: one float4 load
};
__device__ __forceinline__ WSlot wslot_resolve(const Seg* segs, int nseg, int col, int limit) {
  WSlot w; w.base = nullptr; w.rowmap = nullptr; w.bstride = 0; w.ld = 0; w.row_off = 0; w.src_rows = 0; w.relu = 0;
  w.valid = 0; w.vec = 0;
  if (col >= limit) return w;
  int off = 0;
  for (int s = 0; s < nseg; ++s) {
    if (col - off < segs[s].K) {
      w.base = segs[s].base + (col - off); w.rowmap = segs[s].rowmap; w.bstride = segs[s].bstride; w.ld = segs[s].ld;
      w.row_off = segs[s].row_off; w.src_rows = segs[s].src_rows; w.relu = segs[s].relu;
      w.valid = min(4, segs[s].K - (col - off));
      w.vec = w.valid == 4 && (segs[s].ld & 3) == 0 && (segs[s].bstride & 3) == 0 && ((col - off) & 3) == 0 &&
              (((size_t)segs[s].base) & 15) == 0;
      return w;
    }
    off += segs[s].K;
  }
  return w;
}
__device__ __forceinline__ float4 wslot_load(const WSlot& w, int b, int r, int n_rows, int row_end) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!w.base || r >= row_end) return v;
  const int src = w.rowmap ? w.rowmap[(int64_t)b * n_rows + r] : r + w.row_off;
  if (src < 0 || src >= w.src_rows) return v;
  const float* q = w.base + (int64_t)b * w.bstride + (int64_t)src * w.ld;
  if (w.vec) {
    v = *(const float4*)q;
  } else {
    v.x = q[0];
    if (w.valid > 1) v.y = q[1];
    if (w.valid > 2) v.z = q[2];
    if (w.valid > 3) v.w = q[3];
  }
  if (w.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  return v;
}

template <bool TC>
static __global__ void __launch_bounds__(256, 4) wgrad_f32_kernel(WgradArgs a) {
  static_assert(GK == 16 && GM == 64 && GN == 64, "one float4 per thread per tile: 16 rows x 16 column quads");
  __shared__ __align__(16) float Ps[GK][GM + GPAD];
  __shared__ __align__(16) float Qs[GK][GN + GPAD];
  __shared__ float Cs[TC ? GM : 1][GN + 4];
  float cfr[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cfr[i][j] = 0.f;
  const int i0 = blockIdx.x * GM, j0 = blockIdx.y * GN;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4];
  float csum = 0.f;  // threads with tid < GM accumulate the column sum of P (bias gradient)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // this thread's slot: tile row tid / 16, columns 4 (tid % 16) .. +3 of the P tile and of the Q tile
  const int c4 = 4 * (tid % 16), rr = tid / 16;
  const WSlot ps = wslot_resolve(a.p, a.np, i0 + c4, a.I);
  const WSlot qs = wslot_resolve(a.q, a.nq, j0 + c4, a.J);
  const int row_begin = blockIdx.z * a.chunk, row_end = min(a.n_rows, row_begin + a.chunk);
  const bool split = gridDim.z > 1;
  for (int b = 0; b < a.B; ++b) {
    // software pipeline: the next 16-row tile is in registers while the tensor cores work on the current one
    float4 pv = wslot_load(ps, b, row_begin + rr, a.n_rows, row_end);
    float4 qv = wslot_load(qs, b, row_begin + rr, a.n_rows, row_end);
    for (int r0 = row_begin; r0 < row_end; r0 += GK) {
      *(float4*)&Ps[rr][c4] = pv;
      *(float4*)&Qs[rr][c4] = qv;
      __syncthreads();
      if (r0 + GK < row_end) {
        pv = wslot_load(ps, b, r0 + GK + rr, a.n_rows, row_end);
        qv = wslot_load(qs, b, r0 + GK + rr, a.n_rows, row_end);
      }
      if (TC) {
        tile_mma_tf32(Ps, Qs, cfr);
      } else {
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
          float pw[4], qw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pw[i] = Ps[kk][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) qw[j] = Qs[kk][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pw[i], qw[j], acc[i][j]);
        }
      }
      if (a.colsum && blockIdx.y == 0 && tid < GM) {
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) csum += Ps[kk][tid];
      }
      __syncthreads();
    }
  }
  if (TC) frag_to_blocked(Cs, cfr, acc);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ii = i0 + ty * 4 + i, jj = j0 + tx * 4 + j;
      if (ii < a.I && jj < a.J) {
        if (split) atomicAdd(a.out + (int64_t)ii * a.ldo + jj, acc[i][j]);
        else a.out[(int64_t)ii * a.ldo + jj] = acc[i][j];
      }
    }
  if (a.colsum && blockIdx.y == 0 && tid < GM && i0 + tid < a.I) {
    if (split) atomicAdd(a.colsum + i0 + tid, csum);
    else a.colsum[i0 + tid] = csum;
  }
}

// The contraction runs over the rows of a segment (~20 000 in training) while the output is small, so the rows are
// split over blockIdx.z until the grid holds a few thousand blocks; partial products meet in the zeroed output
// through fp32 atomics (the output must be dense: ldo == J).
inline int launch_wgrad(const WgradArgs& a0, cudaStream_t stream, bool tc = false) {
  WgradArgs a = a0;
  dim3 grid((a.I + GM - 1) / GM, (a.J + GN - 1) / GN);
  int nsplit = 1;
  if (a.ldo == a.J) {
    const int tiles = (int)(grid.x * grid.y);
    nsplit = std::max(1, std::min(std::min(64, 4096 / std::max(tiles, 1)), a.n_rows / 256));
  }
  a.chunk = ((a.n_rows + nsplit - 1) / nsplit + GK - 1) / GK * GK;
  nsplit = (a.n_rows + a.chunk - 1) / std::max(a.chunk, 1);
  if (nsplit < 1) nsplit = 1;
  grid.z = nsplit;
  if (nsplit > 1) {
    cudaError_t e_ = cudaMemsetAsync(a.out, 0, sizeof(float) * (size_t)a.I * a.ldo, stream);
    if (e_ == cudaSuccess && a.colsum) e_ = cudaMemsetAsync(a.colsum, 0, sizeof(float) * (size_t)a.I, stream);
    if (e_ != cudaSuccess) return set_error(QP_ECUDA, "wgrad memset failed: %s", cudaGetErrorString(e_));
  }
  if (tc) wgrad_f32_kernel<true><<<grid, 256, 0, stream>>>(a);
  else wgrad_f32_kernel<false><<<grid, 256, 0, stream>>>(a);
  QP_LAUNCH_CHECK();
  return QP_OK;
}

}  // namespace qp
