// Interface of the bf16 tcgen05 GEMMs of the teacher-forced stack (qp_tc.cu).
#pragma once
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

enum { EPI_GATE = 0, EPI_RESSKIP = 1, EPI_HEAD = 2 };

struct Args {
  Seg seg[3];
  int nseg;
  const __nv_bfloat16* W;  // [N][ldw] K-major bf16
  int ldw;
  const float* bias;       // [N]
  int B, n_rows, N;
  int n_begin;             // first output column computed
  int BN;                  // output columns per CTA (multiple of 32, <= 256)
  // EPI_GATE: N = 2C interleaved (sigmoid, tanh) columns -> z (n_rows, C)
  __nv_bfloat16* z_bf;
  float* z_f32;            // optional fp32 copy (backward)
  float* gsave;            // optional (n_rows, 2C) sigmoid / tanh values (backward)
  // EPI_RESSKIP: columns [0,C) residual projection + x(current row); [C,C+S) skip accumulation
  int C, S;
  const float* xcur; long long xcur_bstride; int xcur_off;
  float* xnext; __nv_bfloat16* xnext_bf;
  float* skip; long long skip_bstride; int skip_row0; int skip_accum;
  // EPI_HEAD: out (n_rows, ldo) fp32 = acc + bias; optional relu(out) as bf16 with the same shape
  float* out; long long out_bstride; int ldo;
  __nv_bfloat16* out_relu_bf;
};

int gemm_gate(const Args& a, cudaStream_t st);
int gemm_resskip(const Args& a, cudaStream_t st);
int gemm_head(const Args& a, cudaStream_t st);

// dst[r][k] = bf16(k < K ? (relu?) src[r][k] : 0), k < Kp
int f32_to_bf16_pad(const float* src, long long rows, int K, int Kp, __nv_bfloat16* dst, int relu, cudaStream_t st);
int pack_wg_bf16(const float* Wg, long long rows, int twoC, int Kg, int Kgp, __nv_bfloat16* dst, cudaStream_t st);

}  // namespace tc
}  // namespace qp
