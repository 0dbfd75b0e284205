// Interface of the bf16 tcgen05 GEMMs of the teacher-forced stack (qp_tc.cu).
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "qp_common.cuh"

namespace qp {
namespace tc {

// one K-segment of the gathered A operand: bf16 rows [B][src_rows][ld]; K % 64 == 0
struct Seg {
  const __nv_bfloat16* base;
  long long bstride;
  int ld;
  const int* rowmap;   // optional [B][n_rows] explicit source row (adaptive past tap), else r + row_off
  int row_off;
  int src_rows;        // source rows per batch element (reads are clamped into [0, src_rows))
  int K;
};

enum { EPI_GATE = 0, EPI_RESSKIP = 1, EPI_HEAD = 2, EPI_DGATE = 3, EPI_DX = 4 };

struct Args {
  Seg seg[3];
  int nseg;
  const __nv_bfloat16* W;  // [N][ldw] K-major bf16, or (w_mn) [K][ldw] with the N columns contiguous
  int ldw;
  int w_mn;                // backward GEMMs contract against the forward weights un-transposed: B operand MN-major
  const float* bias;       // [N]
  int B, n_rows, N;
  int n_begin;             // first output column computed
  int BN;                  // output columns per CTA (multiple of 32, <= 256)
  // EPI_GATE: N = 2C interleaved (sigmoid, tanh) columns -> z (n_rows, C)
  __nv_bfloat16* z_bf;
  float* z_f32;            // optional fp32 copy (backward)
  float* gsave;            // optional (n_rows, 2C) sigmoid / tanh values (backward)
  __nv_bfloat16* gsave_bf; // ... or the same as bf16 (what the tcgen05 dz GEMM reads back: half the traffic both ways)
  // EPI_RESSKIP: columns [0,C) residual projection + x(current row); [C,C+S) skip accumulation
  int C, S;
  const float* xcur; long long xcur_bstride; int xcur_off;
  float* xnext; __nv_bfloat16* xnext_bf;
  float* skip; long long skip_bstride; int skip_row0; int skip_accum;
  // EPI_HEAD: out (n_rows, ldo) fp32 = acc + bias; optional relu(out) as bf16 with the same shape
  float* out; long long out_bstride; int ldo;
  __nv_bfloat16* out_relu_bf;
  // EPI_DGATE (backward): N = C columns of dz = [dXnext | dskip] * [R ; K]; gin (n_rows, 2C) fp32 sigmoid / tanh values
  // saved by the forward pass -> dgate (n_rows, 2C): (2c) = dz th sg (1 - sg), (2c+1) = dz sg (1 - th^2)
  const float* gin;
  const __nv_bfloat16* gin_bf;   // bf16 sigmoid / tanh values instead of gin
  __nv_bfloat16* dgate_bf;
  float* dgate_f32;        // optional fp32 copy
  // EPI_DX (backward): acc = dgate * Wg, columns [past C | current C | aux 64]: scatter-add into dx (zeroed by the
  // caller) through the past-tap row map / the current-row offset (+ resid = dXnext on the current rows); aux -> dh
  float* dx; long long dx_bstride; const int* dx_rowmap; int dx_past_off; int dx_cur_off; int dx_rows;
  const float* resid; long long resid_bstride;
  float* dh; long long dh_bstride; int dh_off; int A;
  uint32_t mn_lbo, mn_sbo; // MN-major descriptor strides (filled in by the launcher)
  // Optional pre-blocked copy of W (block_pack): 64 x 64 blocks of 8 KB with the 128-byte swizzle applied, ordered
  // [K stage][64-column block], so the weight tile of a stage is ONE contiguous piece that a single thread moves with
  // cp.async.bulk onto the stage mbarrier.  wb_pitch = 64-column blocks per K stage; K stage index = wb_k0 + k / 64.
  const __nv_bfloat16* Wb; int wb_pitch; int wb_k0;
};

int gemm_gate(const Args& a, cudaStream_t st);
int gemm_resskip(const Args& a, cudaStream_t st);
int gemm_head(const Args& a, cudaStream_t st);
int gemm_dgate(const Args& a, cudaStream_t st);
int gemm_dx(const Args& a, cudaStream_t st);

// Weight-gradient contraction on tcgen05: out[i][j] += sum_b sum_r P[b][r][i] * Q[b][src(r)][j], both operands MN-major
// (the contraction runs over the time rows, the operand rows are contiguous in i / j).  P and Q are assembled from
// column segments whose widths are multiples of 64; rows outside a segment's source range contribute zero.  The rows
// are split over blockIdx.z and the partial products meet in `out` through fp32 atomics: the caller zeroes `out`.
struct alignas(64) WgradArgs {
  // TMA descriptors of the non-gathered segments (p[0], p[1], q[0..2]): bf16 [B][rows][cols] tensors, box 64 x 64,
  // SWIZZLE_128B = exactly one MN-major operand block; rows outside the tensor arrive as zeros.  use_tma[s] = 0:
  // the segment is gathered (row map) or absent and its blocks are copied by the producer threads with cp.async.
  CUtensorMap tm[5];
  int use_tma[5];
  Seg p[2]; int np;        // column segments of P (rows i of the output); base == nullptr: an all-zero segment
  Seg q[3]; int nq;        // column segments of Q (columns j of the output)
  int B, n_rows;
  int I, J;                // output rows / columns written (operand columns beyond them are padding)
  float* out; int ldo;
  float* ones_out;         // optional [I]: column `ones_col` of the product (Q carries a column of ones there): bias gradient
  int ones_col;
  int BJ, chunk;           // set by wgrad(): columns per CTA (multiple of 64, <= 256), rows per split
  uint32_t mn_lbo, mn_sbo;
};
int wgrad(const WgradArgs& a, cudaStream_t st);

// dst = bf16(src) (rows x K, K % 4 == 0, K <= 1024; dst may be NULL) and colsum[k] += sum_r src[r][k] (colsum zeroed by the caller, may be NULL)
int f32_to_bf16_colsum(const float* src, long long rows, int K, __nv_bfloat16* dst, float* colsum, cudaStream_t st);

// dst[r][k] = bf16(k < K ? (relu?) src[r][k] : 0), k < Kp
// ones_col >= 0: that (padding) column is set to 1 instead of 0
int f32_to_bf16_pad(const float* src, long long rows, int K, int Kp, __nv_bfloat16* dst, int relu, cudaStream_t st,
                    int ones_col = -1);
// src [mats][R][Cc] bf16 (R, Cc multiples of 64) -> the same elements as 64 x 64 row-major blocks of 8 KB whose 16-byte
// pieces are XOR-ed with (row & 7) (SWIZZLE_128B), blocks ordered [col block][row block] (col_outer: K-major B operand,
// K = columns) or [row block][col block] (MN-major B operand, K = rows)
int block_pack(const __nv_bfloat16* src, int mats, int R, int Cc, int col_outer, __nv_bfloat16* dst, cudaStream_t st);
int pack_wg_bf16(const float* Wg, long long rows, int twoC, int Kg, int Kgp, __nv_bfloat16* dst, cudaStream_t st);

}  // namespace tc
}  // namespace qp
