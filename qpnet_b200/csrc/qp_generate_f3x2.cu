// tcgen05 generator for 129..256 utterances per launch: qp_generate_f3.cu with TWO utterance groups of up to 128.
//
// Why a second kernel: a sample step is a chain of L + 3 cross-SM exchanges and ~60 % of a phase is waiting (store -> L2
// -> polled load, the cluster reduction, skew between CTAs).  Here every role works through the items of a step in the
// order (phase, group), so one group's tiles are computed while the other group's exchange is in flight; one weight chunk
// load serves both groups.  256 utterances cost 114 us per sample step against 74 us for 128 in qp_generate_f3.cu
// (profiles/r02x_*): 1.30 x the throughput per GPU.  Below 129 utterances qp_generate_f3.cu is faster (its weight slots
// are double-buffered; here the second group's TMEM tiles and receive buffers take that shared memory).
//
// What differs from qp_generate_f3.cu (read that file's header first; everything not listed is the same design):
//   * per group: TMEM tiles (192 columns each + the head tiles = 448 of 512), cluster receive buffers, exchange buffers,
//     rings, look-back table.  Shared: the three A-tile buffers, the weight slots (single-buffered), every role's warps.
//   * the MMA thread issues Z(j,0) P(j,0) Z(j,1) P(j,1) -- a group's x-products between the groups' z-products -- and every
//     staging role follows the same order.  What was tried and measured no faster: a second issuing thread for the
//     x-products and a non-blocking two-stream scheduler on one thread (both let z-products overtake x-products; at the
//     time a priming pass re-read the ring slot it rewrote one phase later, so a lagging past-tap copy changed the primed
//     state run to run -- the passes now alternate slots 0 and 1, see x_slot), starting the groups up to 8 us apart.
//   * the z-products of an item are ONE sequence of 8 UMMAs with N = 96 ([C_j | U_j | P_{j+1}]); the single-group kernel
//     issues C_j separately first (and is 6 us per step slower with one sequence).  Here, with the bulk polling rounds:
//     256 utterances 121.3 -> 114.1 us per step, 192 utterances 117.1 -> 107.9.
//   * a ring slot is stored as the A tiles its readers stage, [4 K-shares][2 K-blocks][256 rows][128 B] with the 16-byte
//     pieces of a row in SWIZZLE_128B order: the past-tap tile of a block with a fixed look-back is two contiguous 16 KB
//     cp.async.bulk copies issued by one thread (adaptive blocks gather 16-byte pieces with cp.async, one slot per
//     utterance).  2048 cp.async per tile kept the PX warps busy for 2.6k cycles per item.
//   * where the time goes at 256 utterances (tools/f3_trace.py --utts 256, profiles/r02x_trace*): the slowest CTAs are
//     never waiting for data; their MMA thread and PX warps are busy ~8k cycles per (phase, group) item -- z-products
//     1.7k, waiting for the single-buffered Wc / Wp chunks 3.4k (group 0 only), x tile poll 2k, x-products 2k -- and every
//     other CTA waits for their tiles.  Double-buffered A tiles and weight slots would need ~100 KB more shared memory.
//     Finer events (profiles/r02ai_*): a UMMA costs the issuing thread ~75 cycles whatever its N, a tcgen05.commit ~110, so
//     the 24 UMMAs and 7 - 9 commits of an item are ~2.7 k cycles of issue time, the waits for the x and past-tap tiles
//     ~1.3 - 2 k; the weight chunks' latency plays no role (L2 policy sweep and a persisting-L2 window: no change).
//
// tcgen05 generator: QPNet.batch_fast_generate (qpnet.py:314-559) for the SI default widths (n_resch 512, n_skipch 256,
// n_quantize 256), up to 256 utterances per launch, any number of residual blocks up to QP_MAX_LAYERS.
//
// Same algebra as qp_generate_fold2.cu (two-level fold: the gate of block j reads z_{j-1} through G_j = Wc_j R_{j-1},
// z_{j-2} through H_j = Wc_j R_{j-2} and x_{j-2} through Wc_j, so ONE 512-vector is on the critical path of a block and a
// sample step is a chain of L + 3 cross-SM exchanges), re-built around the 5th-generation tensor cores so that the
// utterances fill the M = 128 rows of a UMMA instead of 16-row mma.sync tiles:
//
//   D[128 utterances x 32 rows] (fp32, TMEM) += A[128 utterances x 16] (bf16 activations, shared memory, K-major
//   SWIZZLE_128B) * B[32 weight rows x 16]^T (bf16, shared memory, K-major SWIZZLE_128B)
//
// * 128 CTAs = 32 clusters of 4, one CTA per SM.  Cluster c owns residual channels [16c, 16c+16) of every block (32 gate
//   rows, 16 residual rows, 8 skip rows, 8 rows of each head layer); rank r of a cluster contracts over the K-share
//   [128r, 128r+128) of every 512-vector ([64r, 64r+64) of the 256-vectors of the head).
// * The pre-activation of block j is the sum of two TMEM tiles:
//       P_j  = H_j z_{j-2}(t) + Wc_j x_{j-2}(t) + Wp_j x_j(t-k)     built one phase early, off the critical path
//       C_j  = G_j z_{j-1}(t)                                        the only product that waits for this phase's exchange
//   and every product of a phase that contracts z_{j-1} is ONE UMMA sequence of N = 96 rows, [G_j ; [R ; K]_{j-1} ;
//   H_{j+1}] -> [C_j | U_j | P_{j+1}] (96 TMEM columns, two buffers by phase parity; 8 instructions instead of 24).
//   A phase has one cluster reduction on its critical path: TMEM -> registers (C_j + P_j) -> fp16 partial rows ->
//   st.async to the rank that FINISHES those utterances (rank q finishes utterances [32q, 32q+32) of the cluster's rows)
//   -> sum of the four partial tiles + bias + aux + tables -> z_j = sigmoid * tanh -> published.  U_j = [R ; K]_{j-1}
//   z_{j-1} (residual and skip rows) goes the same way through its own warps and publishes x_j.
// * Past taps (the reference's FIFOs, qpnet.py:388-393, 431-437): x_j(t) is published twice, into the exchange buffer the
//   current step's consumers poll and into a ring xr[j][t mod R_j] of bf16 vectors that serves, k steps later, as the past
//   tap x_j(t-k) (k = dil for fixed blocks, -round(-d[t] * dil) for adaptive ones with the reference's rounding,
//   qpnet.py:616-617 / 621-622; k == 0 -> oldest entry, caveat C4).  1 KB per utterance, block and step instead of the
//   16 KB of un-reduced partial products qp_generate_fold2.cu keeps.
// * The aux 1x1 (qpnet.py:663-664 / 632-633) is evaluated at FRAME rate: h_up[:, t] = h[:, t / U] * w[t % U] + b
//   (qpnet.py:143-158), so V h_up = w[t % U] (V h_f) + b (V 1).  (V h_f) is recomputed in fp32 once per frame by the
//   finishing threads, b (V 1) is folded into the bias.  Block 0 reads the causal layer, a function of the last three
//   symbols: three table lookups and no exchange at all; blocks 1 and 2 reach x_0 through two more tables.
// * Exchange words carry a 1-bit epoch tag (LSB of the low bf16 / of the fp32 logit), so a consumer never needs a fence: it
//   polls the data itself.  The exchange buffers are stored as the A tiles their consumers stage ([group][K-share][K-block]
//   [128 rows][128 B], pieces in SWIZZLE_128B order), so a polling round is ONE 32 KB cp.async.bulk straight into the tile
//   (no registers held, no per-piece instructions), after which the role's threads verify the tags of the staged copy and
//   a barrier with an OR reduction decides whether to repeat the round.  With 2 048 cp.async of 16 bytes a round took ~3 k
//   cycles even when every piece was there; the bulk round made the step 13 % shorter (131 -> 114 us at 256 utterances).
//   (A flag-then-load variant, where a consumer first spins on step counters its producers write after their pieces,
//   moved less data and was slower: one more round trip, profiles/r02g_*.)
// * Warp roles (19 warps, no CTA-wide barrier inside the time loop; everything meets through mbarriers):
//     0-3   ET   T tiles: tcgen05.ld -> partial rows -> finishers of z_j, the two head layers
//     4-7   EU   U tiles: partial rows -> fp32 residual / skip state, x_j and relu(skip sum) published
//     8-11  PZ   poll z_{j-1} (and the head's 256-vectors) -> A tile (bulk copy per round + tag check)
//     12-15 PX   poll x_{j-1} -> A tile; the past taps x_{j+1}(t-k) -> A tile (bulk copies; adaptive blocks: cp.async gather)
//     16    MMA  one thread issues every tcgen05.mma and releases buffers with tcgen05.commit
//     17    LOAD one thread streams the weight chunks of the next phase (cp.async.bulk, two slot groups, L2 evict_last)
//     18    SAMP softmax + inverse-CDF / arg-max of utterance blockIdx.x (qpnet.py:507-512), symbol fed back
//
// * TWO utterance groups of up to 128 share a launch (256 utterances): every role works through the items of a step in the
//   order (phase, group), so while the exchange of one group is in flight (a phase is ~60 % waiting: store -> L2 -> polled
//   load, the cluster reduction, skew between CTAs) the other group computes on the same weight chunk.  The groups have
//   their own TMEM tiles, receive buffers, rings and exchange buffers; they share the A-tile buffers, the weight slots
//   (one chunk load serves both) and every role's warps.
//
// Priming (qpnet.py:355-440): the pad region is constant in time; the constant is found by running the step NP >= L
// times with every past tap reading the previous pass (pass i makes z_i and the ring of block i+1 exact); the last pass
// fills every ring slot.
#include <algorithm>
#include <type_traits>

#include "qp_gen_common.cuh"
#include "qp_pack.cuh"

namespace qp {
namespace f3x2 {

constexpr int CL = 4, NCL = 32, NCTA = CL * NCL;
constexpr int UB = 128;                       // utterance rows of an A tile = UMMA M = utterances of a group
constexpr int NG = 2;                         // utterance groups per launch
constexpr int UT = NG * UB;                   // utterances per launch
constexpr int C = 512, S = 256, Q = 256;
constexpr int KS = C / CL;                    // 128: K-share of a 512-vector
constexpr int KH = S / CL;                    // 64: K-share of a 256-vector
constexpr int MAXL = QP_MAX_LAYERS;
constexpr int WCH_E = 32 * KS, WCH_B = WCH_E * 2;     // weight chunk: 32 rows x 128 K, two 4 KB K-blocks
constexpr int ZPC_E = 96 * KS, ZPC_B = ZPC_E * 2;     // z-product chunk [G ; [R;K] ; H]: 96 rows x 128 K, two 12 KB K-blocks
constexpr int HCH_E = 16 * KH, HCH_B = HCH_E * 2;     // head chunk: 16 rows (8 live) x 64 K
constexpr int ABLK = UB * 128;                         // bytes of one K-block of an A tile (128 rows x 128 B)
constexpr int NWARP = 19, NT = NWARP * 32;
constexpr int MAXA = 8;                                // adaptive blocks (look-back table in shared memory)
// trace events of one CTA (QPNET_GEN_TRACE_STEP, QPNET_GEN_TRACE_CTA), per phase j and group (group g at event + 32 g).
// MMA thread: 0 z tile seen, 24 z UMMAs issued, 1 / 25 / 26 / 27 after each commit, 22 Wc chunk seen, 8 x tile seen, 28 x
// UMMAs issued, 29 x buffer committed, 23 Wp chunk seen, 9 past-tap tile seen, 10 phase issued.  ET thread 0: 2 tiles in TMEM, 3 partial rows sent,
// 4 partial rows of the cluster arrived, 5 z published.  PZ thread 0: 11 z buffer free, 6 z staged.  EU thread 0: 13 U tile
// in TMEM, 14 sent, 15 arrived, 16 x published.  PX thread 0: 17 past-tap buffer free, 18 past-tap copies issued, 19 x buffer
// free, 21 x staged
constexpr int TRACE_EVENTS = 64;   // group g at 32 g

// shared memory map (bytes from a 1024-byte aligned base)
constexpr int SM_Z = 0, SM_X = SM_Z + 2 * ABLK, SM_XP = SM_X + 2 * ABLK;      // A tiles, shared by the groups
constexpr int SM_WZ = SM_XP + 2 * ABLK;                // weight slots (one chunk serves both groups): z-product chunk,
constexpr int SM_WC = SM_WZ + ZPC_B;                   // Wc,
constexpr int SM_WP = SM_WC + WCH_B;                   // Wp
constexpr int SM_WH = SM_WP + WCH_B;                   // the two head tiles (resident)
constexpr int RT_B = 2 * 4 * 32 * 64;                  // recvT [2][4 src][32 utt][64 B]
constexpr int RU_B = 3 * 4 * 32 * 48;                  // recvU [3][4][32][48 B]
constexpr int RH_B = 2 * 4 * 32 * 16;                  // recvH [2][4][32][16 B]
constexpr int RG_B = RT_B + RU_B + RH_B;               // receive buffers of one group
constexpr int SM_R = SM_WH + 2 * HCH_B;
constexpr int SM_K = SM_R + NG * RG_B;                 // look-backs of this step [NG][MAXA][UB] uint16
constexpr int SM_BAR = SM_K + NG * MAXA * UB * 2;      // 64 mbarriers
constexpr int SM_TMEM = SM_BAR + 64 * 8;
constexpr int SM_BH = SM_TMEM + 16;                    // head biases [2][8]
constexpr int SM_END = SM_BH + 64;
static_assert(SM_END + 1024 <= 227 * 1024, "shared memory budget");
// mbarrier indices: shared ones, then per group ([2] = TMEM buffer by phase parity)
constexpr int B_ZPW_FULL = 0, B_ZPW_FREE = 1, B_WCW_FULL = 2, B_WCW_FREE = 3, B_WPW_FULL = 4, B_WPW_FREE = 5,
              B_ZFULL = 6, B_ZFREE = 7, B_XFULL = 8, B_XFREE = 9, B_XPFULL = 10, B_XPFREE = 11, B_ZPOLL = 12, B_XPOLL = 13, B_GROUP0 = 14;
constexpr int G_CFULL = 0, G_UFULL = 2, G_PFULL = 4, G_CFREE = 6, G_UFREE = 8, G_PFREE = 10, G_HFULL = 12,
              G_RT = 14, G_RU = 16, G_RH = 19, G_COUNT = 21;
constexpr int B_COUNT = B_GROUP0 + NG * G_COUNT;
static_assert(B_COUNT <= 64, "mbarrier area / one parity bit per barrier in a 64-bit word");
constexpr int TM_COLS = 512;                           // TMEM: group g at 192 g: buffer b at + 96 b: [C | U | P] 32 each;
constexpr int TC_C = 0, TC_U = 32, TC_P = 64, TC_BUF = 96, TC_GRP = 192, TC_H = NG * TC_GRP;   // heads: TC_H + 32 g + 16 hd

struct Plan {
  int A, L, nF, nA, U, B, F, M, NP;
  int dil[MAXL], depth[MAXL], rlog[MAXL];      // ring of block l: 1 << rlog[l] slots (l >= 1)
  int32_t* status;
  const float** tab;
  __nv_bfloat16* Wzp;     // [L phases 1..L][NCTA][ZPC_E]  pre-swizzled z-product chunks (layout in pack_kernel)
  __nv_bfloat16* Wcp;     // [2: Wc, Wp][L][NCTA][WCH_E]
  __nv_bfloat16* Whead;   // [2][NCTA][HCH_E]
  float* bgate;           // [L][NCL][32]
  float* bres;            // [L][NCL][32]   rows 0-15 residual, 16-23 skip
  float* bhead;           // [2][NCL][8]
  float* T0;              // [NCL][3][Q][32]    block-0 gate tables: newest symbol, previous, the one before
  float* T12;             // [NCL][2][2][Q][32] Wc_1 x_0, Wc_2 x_0: tap 0 = newest symbol (E1 + bias), tap 1 = previous (E0)
  float* Eo;              // [NCL][2][Q][16]    causal-layer rows of the cluster's channels (bias folded into tap 1)
  float* Paux;            // [NCTA][NG][L][32 utt][32 rows]  V h_f of the finishing CTA's utterances, per frame
  uint32_t* xr[MAXL];     // [1 << rlog][CL][2][UT][32]  x_l(t) ring, tagged words, a slot laid out as its readers' A tiles
  uint32_t* vz;           // [L][UT][C / 2]
  uint32_t* vx;           // [L][UT][C / 2]  x_l of the current step (the ring copy serves the past taps)
  uint32_t* v256;         // [2][UT][S / 2]   0: relu(skip sum), 1: relu(head-1)
  uint32_t* vlog;         // [UT][Q]  fp32 logits
  uint32_t* vsym;         // [UT][32] fed-back symbol, one line per utterance
  int16_t* pcm_lut;       // [Q] decode_mu_law(symbol) * 32768 clipped to int16 (qpnet_decode.py:315-318)
  void* tagged_begin; size_t tagged_bytes;
  long long* trace; int trace_step0, trace_nsteps, trace_cta;
  int nowait;             // debug (QPNET_F3_NOWAIT): no poll waits for fresh data -- wrong symbols, but the step time is then the CTAs' local
                          // pipeline alone (profiles/r02ao_*: 107 of 131 us per step at 256 utterances; with the bulk polling rounds 97 of 114)
  long long* gtrace;      // [NCTA][L + 4][NG][4] %globaltimer of one step, every CTA: 0 z published, 1 z staged, 2 partial rows arrived, 3 x published
};

static int log2_above(int v) { int q = 0; while ((1 << q) <= v) ++q; return q; }

bool supported(const QpArch* a, int B) {
  if (a->n_resch != C || a->n_skipch != S || a->n_quantize != Q || a->n_aux > 64 || a->n_aux < 1) return false;
  const int L = a->n_fixed + a->n_adaptive;
  if (L > MAXL || L < 4 || a->n_fixed < 1 || a->dil_fixed[0] != 1 || a->n_adaptive > MAXA) return false;
  return B >= 1 && B <= UT;
}

size_t make_plan(const QpArch* a, int B, int F, int M, void* base, size_t cap, Plan* p) {
  p->A = a->n_aux; p->nF = a->n_fixed; p->nA = a->n_adaptive; p->L = p->nF + p->nA; p->U = a->upsampling;
  p->B = B; p->F = F; p->M = M;
  const int L = p->L;
  p->NP = L + (L & 1);
  Arena ar(base, cap);
  p->status = ar.take<int32_t>(64);
  p->tab = ar.take<const float*>(tensor_map(a).count());
  p->Wzp = ar.take<__nv_bfloat16>((size_t)L * NCTA * ZPC_E);
  p->Wcp = ar.take<__nv_bfloat16>((size_t)2 * L * NCTA * WCH_E);
  p->Whead = ar.take<__nv_bfloat16>((size_t)2 * NCTA * HCH_E);
  p->bgate = ar.take<float>((size_t)L * NCL * 32);
  p->bres = ar.take<float>((size_t)L * NCL * 32);
  p->bhead = ar.take<float>((size_t)2 * NCL * 8);
  p->T0 = ar.take<float>((size_t)NCL * 3 * Q * 32);
  p->T12 = ar.take<float>((size_t)NCL * 4 * Q * 32);
  p->Eo = ar.take<float>((size_t)NCL * 2 * Q * 16);
  p->Paux = ar.take<float>((size_t)NCTA * NG * L * 32 * 32);
  p->pcm_lut = ar.take<int16_t>(Q);
  ar.off = align_up(ar.off, 256);
  size_t t0 = ar.off;
  for (int l = 0; l < MAXL; ++l) { p->dil[l] = 0; p->depth[l] = 0; p->rlog[l] = 0; p->xr[l] = nullptr; }
  for (int l = 0; l < L; ++l) {
    p->dil[l] = l < p->nF ? a->dil_fixed[l] : a->dil_adaptive[l - p->nF];
    p->depth[l] = l < p->nF ? p->dil[l] : p->dil[l] * M;
    p->rlog[l] = log2_above(p->depth[l]);
    p->xr[l] = l >= 1 ? ar.take<uint32_t>(((size_t)1 << p->rlog[l]) * UT * (C / 2)) : nullptr;
  }
  p->vz = ar.take<uint32_t>((size_t)L * UT * (C / 2));
  p->vx = ar.take<uint32_t>((size_t)L * UT * (C / 2));
  p->v256 = ar.take<uint32_t>((size_t)2 * UT * (S / 2));
  p->vlog = ar.take<uint32_t>((size_t)UT * Q);
  p->vsym = ar.take<uint32_t>((size_t)UT * 32);
  ar.off = align_up(ar.off, 256);
  p->tagged_begin = base ? (char*)base + t0 : nullptr;
  p->tagged_bytes = ar.off - t0;
  p->trace = ar.take<long long>((size_t)8 * (L + 4) * TRACE_EVENTS);
  p->gtrace = ar.take<long long>((size_t)NCTA * (L + 4) * 8);
  p->trace_step0 = -1000000; p->trace_nsteps = 8; p->trace_cta = 0;
  return align_up(ar.off, 256);
}

// ------------------------------------------------------------------ weight packing
// element (row, k) of a K-major SWIZZLE_128B operand block whose rows are 128 bytes (64 bf16): 8-row atoms of 1 KB,
// the 16-byte piece index XOR-ed with the row index inside the atom; `kblock_elems` = elements of one 64-wide K-block
__host__ __device__ inline int sw128_elem(int row, int k, int kblock_elems) {
  const int kb = k >> 6, kk = k & 63;
  return kb * kblock_elems + (row >> 3) * 512 + (row & 7) * 64 + ((((kk >> 3) ^ (row & 7)) << 3) | (kk & 7));
}

// gate row `row` of cluster c: channel 16c + row / 2, g = row & 1 (0 sigmoid, 1 tanh)
__device__ __forceinline__ float wc_elem(const TensorMap& tm, const float* const* tab, int nF, int gl, int g, int ch, int col) {
  return gl < nF ? tab[tm.dilF_w(g, gl)][((size_t)ch * C + col) * 2 + 1] : tab[tm.dilA_wC(g, gl - nF)][(size_t)ch * C + col];
}
__device__ __forceinline__ float wp_elem(const TensorMap& tm, const float* const* tab, int nF, int gl, int g, int ch, int col) {
  return gl < nF ? tab[tm.dilF_w(g, gl)][((size_t)ch * C + col) * 2 + 0] : tab[tm.dilA_wP(g, gl - nF)][(size_t)ch * C + col];
}

// Weight chunks of CTA s = 4c + r; column k of a chunk is input channel 128 r + k.
//   Wzp[phase j-1][s] (96 rows): rows 0-31  G_j gate rows (fold_kernel; j <= L-1)
//                                rows 32-47 R_{j-1} row 16c + row - 32 (zero for the dead last projection, C7)
//                                rows 48-55 K_{j-1} row 8c + row - 48, rows 56-63 zero
//                                rows 64-95 H_{j+1} gate rows (fold_kernel; j + 1 <= L-1)
//   Wcp[0][j][s] (32 rows): Wc_j gate rows (current tap; j >= 3, blocks 1 and 2 reach x_0 through tables)
//   Wcp[1][j][s] (32 rows): Wp_j gate rows (past tap; j >= 1)
// plus head tiles, biases and the causal-layer table.  Wzp is zeroed before this kernel; every other element that is ever
// loaded is written here.
__global__ void pack_kernel(TensorMap tm, Plan p, const float* const* __restrict__ tab) {
  const int L = p.L, A = p.A, nF = p.nF;
  const size_t n_ch = (size_t)L * NCTA * WCH_E;              // per kind
  const size_t n_w = 3 * n_ch;                               // [R;K], Wc, Wp
  const size_t n_wh = (size_t)2 * NCTA * HCH_E, n_b = (size_t)L * NCL * 32, n_bh = (size_t)2 * NCL * 8, n_eo = (size_t)NCL * 2 * Q * 16;
  const size_t total = n_w + n_wh + 2 * n_b + n_bh + n_eo;
  const float up_b = tab[tm.up_b()][0];
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t k = idx;
    if (k < n_w) {
      const int which = (int)(k / n_ch);                      // 0 [R;K], 1 Wc, 2 Wp
      size_t q = k % n_ch;
      const int kk = (int)(q % KS); q /= KS;
      const int row = (int)(q % 32); q /= 32;
      const int s = (int)(q % NCTA), l = (int)(q / NCTA);
      const int c = s / CL, r = s % CL, col = KS * r + kk;
      float v = 0.f;
      if (which == 0) {
        if (row < 16) {
          if (l < L - 1) v = l < nF ? tab[tm.resF_w(l)][(size_t)(16 * c + row) * C + col] : tab[tm.resA_w(l - nF)][(size_t)(16 * c + row) * C + col];
        } else if (row < 24) {
          const int sr = 8 * c + row - 16;
          v = l < nF ? tab[tm.skipF_w(l)][(size_t)sr * C + col] : tab[tm.skipA_w(l - nF)][(size_t)sr * C + col];
        }
        p.Wzp[((size_t)l * NCTA + s) * ZPC_E + sw128_elem(32 + row, kk, 96 * 64)] = __float2bfloat16(v);   // phase l + 1
      } else {
        if (l >= 1) v = which == 1 ? wc_elem(tm, tab, nF, l, row & 1, 16 * c + (row >> 1), col)
                                   : wp_elem(tm, tab, nF, l, row & 1, 16 * c + (row >> 1), col);
        p.Wcp[(((size_t)(which - 1) * L + l) * NCTA + s) * WCH_E + sw128_elem(row, kk, 32 * 64)] = __float2bfloat16(v);
      }
      continue;
    }
    k -= n_w;
    if (k < n_wh) {
      const int kk = (int)(k % KH); size_t q = k / KH;
      const int row = (int)(q % 16); q /= 16;
      const int s = (int)(q % NCTA), hd = (int)(q / NCTA);
      const int c = s / CL, r = s % CL;
      float v = 0.f;
      if (row < 8) v = tab[hd ? tm.post2_w() : tm.post1_w()][(size_t)(8 * c + row) * S + KH * r + kk];
      p.Whead[((size_t)hd * NCTA + s) * HCH_E + sw128_elem(row, kk, 16 * 64)] = __float2bfloat16(v);
      continue;
    }
    k -= n_wh;
    if (k < n_b) {   // gate biases: every bias that feeds the pre-activation, + b_up (V 1) of the aux 1x1 (block 0: +
                     // (Wc + Wp) . causal bias; blocks j >= 1: fold_kernel adds Wc_j . (r_{j-1} + r_{j-2}))
      const int row = (int)(k % 32); size_t q = k / 32;
      const int c = (int)(q % NCL), l = (int)(q / NCL);
      const int g = row & 1, ch = 16 * c + (row >> 1);
      float v;
      const float* V;
      if (l < nF) { v = tab[tm.dilF_b(g, l)][ch] + tab[tm.auxF_b(g, l)][ch]; V = tab[tm.auxF_w(g, l)] + (size_t)ch * A; }
      else { int a = l - nF; v = tab[tm.dilA_bC(g, a)][ch] + tab[tm.dilA_bP(g, a)][ch] + tab[tm.auxA_b(g, a)][ch]; V = tab[tm.auxA_w(g, a)] + (size_t)ch * A; }
      float sv = 0.f;
      for (int a = 0; a < A; ++a) sv += V[a];
      v = fmaf(up_b, sv, v);
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
      const int row = (int)(k % 32); size_t q = k / 32;
      const int c = (int)(q % NCL), l = (int)(q / NCL);
      float v = 0.f;
      if (row < 16) v = l < nF ? tab[tm.resF_b(l)][16 * c + row] : tab[tm.resA_b(l - nF)][16 * c + row];
      else if (row < 24) v = l < nF ? tab[tm.skipF_b(l)][8 * c + row - 16] : tab[tm.skipA_b(l - nF)][8 * c + row - 16];
      p.bres[k] = v;
      continue;
    }
    k -= n_b;
    if (k < n_bh) {
      const int j = (int)(k % 8); size_t q = k / 8;
      const int c = (int)(q % NCL), hd = (int)(q / NCL);
      p.bhead[k] = tab[hd ? tm.post2_b() : tm.post1_b()][8 * c + j];
      continue;
    }
    k -= n_bh;
    {
      const int j = (int)(k % 16); size_t q = k / 16;
      const int sym = (int)(q % Q); q /= Q;
      const int tap = (int)(q % 2), c = (int)(q / 2);
      const int ch = 16 * c + j;
      p.Eo[k] = tab[tm.causal_w()][((size_t)ch * Q + sym) * 2 + tap] + (tap ? tab[tm.causal_b()][ch] : 0.f);
    }
  }
}

// Folded products for 8 gate rows of a cluster (fp32 accumulate, rounded to bf16 once).  grid (NCL * 4, L, 2):
//   z = 0: gate block gl = blockIdx.y:  G_gl = Wc_gl R_{gl-1}   (gl >= 1)  -> rows 0-31 of the z-product chunk of phase gl
//   z = 1:                              H_gl = Wc_gl R_{gl-2}   (gl >= 2)  -> rows 64-95 of the chunk of phase gl - 1
// and the bias terms Wc_gl . r_{gl-1} / Wc_gl . r_{gl-2}.  Runs after pack_kernel on the same stream (it adds to bgate).
__global__ void __launch_bounds__(256) fold_kernel(TensorMap tm, Plan p, const float* const* __restrict__ tab) {
  __shared__ float sWc[8][C];
  const int c = blockIdx.x >> 2, rg = blockIdx.x & 3, gl = blockIdx.y, which = blockIdx.z, tid = threadIdx.x;
  const int nF = p.nF;
  const int rl = gl - 1 - which;                    // residual projection folded in
  if (rl < 0) return;                               // G_0, H_0, H_1 do not exist (never loaded)
  for (int e = tid; e < 8 * C; e += 256) {
    const int j8 = e / C, col = e % C, row = 8 * rg + j8;
    sWc[j8][col] = wc_elem(tm, tab, nF, gl, row & 1, 16 * c + (row >> 1), col);
  }
  __syncthreads();
  const float* R = rl < nF ? tab[tm.resF_w(rl)] : tab[tm.resA_w(rl - nF)];   // [out m][in k]
  float acc[8][2];
#pragma unroll
  for (int j8 = 0; j8 < 8; ++j8) acc[j8][0] = acc[j8][1] = 0.f;
  for (int m = 0; m < C; ++m) {
    const float r0 = R[(size_t)m * C + tid], r1 = R[(size_t)m * C + tid + 256];
#pragma unroll
    for (int j8 = 0; j8 < 8; ++j8) {
      const float w = sWc[j8][m];
      acc[j8][0] = fmaf(w, r0, acc[j8][0]);
      acc[j8][1] = fmaf(w, r1, acc[j8][1]);
    }
  }
  // G_gl is rows 0-31 of the chunk of phase gl, H_gl rows 64-95 of the chunk of phase gl - 1
  const int ph = which == 0 ? gl - 1 : gl - 2, row0 = which == 0 ? 0 : 64;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int col = tid + 256 * hh;
    const int s = c * CL + col / KS, kk = col % KS;
    __nv_bfloat16* dst = p.Wzp + ((size_t)ph * NCTA + s) * ZPC_E;
#pragma unroll
    for (int j8 = 0; j8 < 8; ++j8) dst[sw128_elem(row0 + 8 * rg + j8, kk, 96 * 64)] = __float2bfloat16(acc[j8][hh]);
  }
  if (tid < 8) {
    const float* rb = rl < nF ? tab[tm.resF_b(rl)] : tab[tm.resA_b(rl - nF)];
    float a = 0.f;
    for (int m = 0; m < C; ++m) a = fmaf(sWc[tid][m], rb[m], a);
    atomicAdd(&p.bgate[((size_t)gl * NCL + c) * 32 + 8 * rg + tid], a);   // the G and H blocks of one gate run concurrently
  }
}

// Block-0 gate tables (fp32): the causal layer output is x_0 = E0[s(t-2)] + E1[s(t-1)] + b (qpnet.py:447-448, 561-564),
// so Wc.x0(t) + Wp.x0(t-1) = TA[s(t-1)] + TB[s(t-2)] + TC[s(t-3)] + const with TA = Wc.E1, TB = Wc.E0 + Wp.E1, TC = Wp.E0.
// blockIdx.z = 1, 2: the x_0 term of blocks 1 and 2, Wc_b . x_0 = U_b[s(t-1)] + V_b[s(t-2)], U_b = Wc_b.(E1 + bias), V_b = Wc_b.E0.
// grid (32 rows, NCL, 3), block Q threads (one symbol each)
__global__ void table_kernel(TensorMap tm, Plan p, const float* const* __restrict__ tab) {
  const int row = blockIdx.x, c = blockIdx.y, sym = threadIdx.x, b = blockIdx.z;
  const int g = row & 1, ch = 16 * c + (row >> 1);
  const float* E = tab[tm.causal_w()];                          // [col][Q][tap]
  if (b == 0) {
    const float* W = tab[tm.dilF_w(g, 0)] + (size_t)ch * C * 2;   // [col][tap]: tap 0 past, tap 1 current
    float ta = 0.f, tb = 0.f, tc = 0.f;
    for (int col = 0; col < C; ++col) {
      const float wp = W[2 * col], wc = W[2 * col + 1];
      const float2 e = *(const float2*)(E + ((size_t)col * Q + sym) * 2);   // (E0, E1)
      ta = fmaf(wc, e.y, ta);
      tb = fmaf(wc, e.x, fmaf(wp, e.y, tb));
      tc = fmaf(wp, e.x, tc);
    }
    float* T = p.T0 + (size_t)c * 3 * Q * 32;
    T[(0 * Q + sym) * 32 + row] = ta;
    T[(1 * Q + sym) * 32 + row] = tb;
    T[(2 * Q + sym) * 32 + row] = tc;
  } else {
    const float* cb = tab[tm.causal_b()];
    float u = 0.f, v = 0.f;
    for (int col = 0; col < C; ++col) {
      const float wc = wc_elem(tm, tab, p.nF, b, g, ch, col);
      const float2 e = *(const float2*)(E + ((size_t)col * Q + sym) * 2);
      u = fmaf(wc, e.y + cb[col], u);
      v = fmaf(wc, e.x, v);
    }
    float* T = p.T12 + ((size_t)c * 2 + (b - 1)) * 2 * Q * 32;
    T[(0 * Q + sym) * 32 + row] = u;
    T[(1 * Q + sym) * 32 + row] = v;
  }
}

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// named barrier of n threads that also ORs a predicate over them
__device__ __forceinline__ bool bar_or(int id, int n, bool pred) {
  unsigned r;
  asm volatile("{\n .reg .pred p, q;\n setp.ne.u32 p, %1, 0;\n bar.red.or.pred q, %2, %3, p;\n selp.u32 %0, 1, 0, q;\n}\n"
               : "=r"(r) : "r"((unsigned)pred), "r"(id), "r"(n) : "memory");
  return r != 0;
}
__device__ __forceinline__ void st_strong_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// K-major SWIZZLE_128B shared-memory matrix descriptor: start address >> 4, LBO (unused) 1, SBO = 1024 B between 8-row
// atoms, version 1, layout SWIZZLE_128B (the encoding qp_tc.cu runs the teacher-forced GEMMs with)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D fp32, A / B bf16, both K-major
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N) {
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg((const float4*)p); }

// two 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t (&a)[16], uint32_t tb, uint32_t (&b)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
        "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(ta) : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]),
        "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
      : "r"(tb) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// the mbarrier receives one (pre-counted) arrival once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
// wait parities, one bit per mbarrier of the role; "free" barriers start at 1 (the phase before the first one counts as
// complete, so the first wait of a producer passes)
constexpr unsigned long long group_free_bits(int g) {
  return ((3ull << G_CFREE) | (3ull << G_UFREE) | (3ull << G_PFREE)) << (B_GROUP0 + g * G_COUNT);
}
constexpr unsigned long long FREE_BITS = (1ull << B_ZPW_FREE) | (1ull << B_WCW_FREE) | (1ull << B_WPW_FREE) | (1ull << B_ZFREE) |
                                         (1ull << B_XFREE) | (1ull << B_XPFREE) | group_free_bits(0) | group_free_bits(1);

template <bool TRACE>
__global__ void __launch_bounds__(NT, 1) f3x2_gen_kernel(Plan p, GenArgsDev g) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;              // SWIZZLE_128B atoms need 1024-byte alignment
  unsigned char* const sm = smem_raw + (sbase - raw);
  const int L = p.L, B = p.B, U = p.U, NP = p.NP;
  const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = s / CL;
  const int rank = (int)cluster_rank();                       // == s % CL
  const int ng = B > UB ? 2 : 1;                              // utterance groups in this launch
  const int half = Q / 2;
  float* const sBh = (float*)(sm + SM_BH);
  unsigned short* const sK = (unsigned short*)(sm + SM_K);
  auto bar = [&](int i) -> uint32_t { return sbase + SM_BAR + 8 * i; };
  auto gb = [&](int gi, int i) -> int { return B_GROUP0 + gi * G_COUNT + i; };   // barrier i of group gi
  auto gB = [&](int gi) -> int { return min(UB, B - UB * gi); };                  // live utterances of group gi
  auto gnl = [&](int gi) -> int { return (gB(gi) + 31) >> 5; };                   // ranks that finish at least one of them

  // ---- one-time staging ---------------------------------------------------------------
  for (int e = tid; e < SM_END / 16; e += NT) ((uint4*)sm)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int e = tid; e < 2 * HCH_E / 8; e += NT)               // both head tiles stay resident
    ((uint4*)(sm + SM_WH))[e] = ((const uint4*)(p.Whead + ((size_t)(e / (HCH_E / 8)) * NCTA + s) * HCH_E))[e % (HCH_E / 8)];
  if (tid < 16) sBh[tid] = p.bhead[((size_t)(tid >> 3) * NCL + c) * 8 + (tid & 7)];
  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      unsigned cnt = 1;
      if (i == B_ZFULL || i == B_XFULL || i == B_XPFULL) cnt = 128;
      else if (i >= B_GROUP0) {
        const int k = (i - B_GROUP0) % G_COUNT;
        if (k >= G_CFREE && k < G_PFREE + 2) cnt = 4;
      }
      mbar_init(bar(i), cnt);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sbase + SM_TMEM), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  fence_proxy_async();                                        // zero fill / head tiles (generic proxy) before async-proxy reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *(volatile uint32_t*)(sm + SM_TMEM);
  cluster_sync();

  auto trace = [&](int t, int gi, int phase, int ev) {
    if (TRACE && s == p.trace_cta && (tid & 127) == 0 && t >= p.trace_step0 && t < p.trace_step0 + p.trace_nsteps)
      p.trace[((size_t)(t - p.trace_step0) * (L + 4) + phase) * TRACE_EVENTS + 32 * gi + ev] = clock64();
  };
  auto gtrace = [&](int t, int gi, int phase, int ev) {   // the same step on every CTA, on the global clock
    if (TRACE && (tid & 127) == 0 && t == p.trace_step0 + 2) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(ns));
      p.gtrace[((size_t)s * (L + 4) + phase) * 8 + 4 * gi + ev] = (long long)ns;
    }
  };
  // the watchdog: a word or barrier that never arrives is a bug, not a schedule; record it and stop the grid
  // (a polling round is at least an L2 round trip or an mbarrier try_wait time-out, a few hundred cycles: 2^23 rounds
  // are seconds, far beyond any schedule; counting rounds keeps the hot loops free of clock reads and 64-bit state)
  auto spin_check = [&](unsigned& spins) {
    if ((++spins & 0xFFFFu) != 0) return;
    if (*((volatile int32_t*)p.status) != 0 || spins >= (1u << 23)) {
      atomicExch(p.status, QP_ETIMEOUT);
      __threadfence_system();
      __trap();
    }
  };
  auto spin_check_bar = [&](unsigned& spins) {   // mbarrier waits: a try_wait round can be much shorter than an L2 round trip
    if ((++spins & 0xFFFFu) != 0) return;
    if (*((volatile int32_t*)p.status) != 0 || spins >= (1u << 26)) {
      atomicExch(p.status, QP_ETIMEOUT);
      __threadfence_system();
      __trap();
    }
  };
  unsigned long long par = FREE_BITS;                         // this thread's wait parities
  auto waitb = [&](int i) {
    const uint32_t b = bar(i);
    const unsigned parity = (unsigned)(par >> i) & 1u;
    unsigned spins = 0;
    while (!mbar_try(b, parity)) spin_check_bar(spins);
    par ^= 1ull << i;
  };
  // slot of x_l(t) in its ring (priming passes: slot 0); every exchange word of a step carries the step's parity
  // (independent of the batch: an utterance's symbols must not depend on its batch-mates)
  // Priming passes alternate between slots 0 and 1 (every ring has at least two): a pass reads what the PREVIOUS pass wrote
  // (slot (t - 1) & 1) while its own x goes to slot t & 1, so a past-tap copy that is still in flight when block j+1's
  // finishers write can never see the newer value -- the result does not depend on how far a CTA's copies lag.
  auto x_slot = [&](int l, int t) -> int { return t < 0 ? (t & 1) : (t & ((1 << p.rlog[l]) - 1)); };
  // The exchange buffers vz[l], vx[l] are stored as the A tiles their consumers stage (like the ring slots):
  // [group][K-share 4][K-block 2][128 rows][128 B], the 16-byte pieces of a row in SWIZZLE_128B order.  Word offset of piece
  // pw (0..63: 8 channels each) of utterance fu inside one vz[l] / vx[l]:
  auto xoff = [&](int fu, int pw) -> size_t {
    const int gi_ = fu / UB, u = fu % UB;
    return ((size_t)(gi_ * 8 + (pw >> 3)) * UB + u) * 32 + 4 * ((pw ^ u) & 7);
  };
  // Poll a whole 32 KB tile (this rank's K-share of one group: contiguous in the exchange buffer) with ONE bulk copy per
  // round: thread 0 of the role issues it, everyone waits on its mbarrier, checks the tags of its own pieces on the staged
  // copy and a barrier with an OR reduction decides whether the round is repeated.  (2 048 cp.async of 16 bytes took
  // ~3 k cycles per round; the tile has to be on its way anyway, and the copy engine moves it in one piece.)
  auto poll_bulk = [&](const uint32_t* src, uint32_t dst, int pollbar, int named, int i128_, int nrows, unsigned tag) {
    const int pc_ = i128_ & 15, ub_ = i128_ >> 4;       // this thread checks piece position pc_ of rows ub_ + 8 i
    const uint32_t mine = dst + (uint32_t)((pc_ >> 3) * ABLK + ub_ * 128 + ((pc_ & 7) << 4));
    unsigned spins = 0;
    while (true) {
      if (i128_ == 0) {
        // a ragged group: only the 8-row atoms that hold utterances (the first ceil(nrows / 8) KB of each K-block)
        const unsigned kbytes = nrows >= UB ? (unsigned)ABLK : (unsigned)((nrows + 7) >> 3) * 1024u;
        if (kbytes == (unsigned)ABLK) {
          mbar_expect_tx(bar(pollbar), 2 * ABLK);
          bulk_g2s_plain(dst, src, 2 * ABLK, bar(pollbar));
        } else {
          mbar_expect_tx(bar(pollbar), 2 * kbytes);
          bulk_g2s_plain(dst, src, kbytes, bar(pollbar));
          bulk_g2s_plain(dst + ABLK, src + (size_t)UB * 32, kbytes, bar(pollbar));
        }
      }
      waitb(pollbar);
      unsigned bad = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (ub_ + 8 * i < nrows) {
          uint4 x;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(mine + i * 1024) : "memory");
          bad |= (x.x ^ tag) | (x.y ^ tag) | (x.z ^ tag) | (x.w ^ tag);
        }
      if (!bar_or(named, 128, (bad & 1u) != 0) || p.nowait) break;
      spin_check(spins);
    }
  };
  // the head's 256-vectors the same way: [group][K-share 4][128 rows][128 B] (one K-block of 64 channels per rank); word offset
  // of piece pw (0..31) of utterance fu inside v256[hd]
  auto hoff = [&](int fu, int pw) -> size_t {
    const int gi_ = fu / UB, u = fu % UB;
    return ((size_t)(gi_ * 4 + (pw >> 3)) * UB + u) * 32 + 4 * ((pw ^ u) & 7);
  };
  auto poll_bulk_head = [&](const uint32_t* src, uint32_t dst, int pollbar, int named, int i128_, int nrows, unsigned tag) {
    const int pc_ = i128_ & 7, ub_ = i128_ >> 3;        // this thread checks piece position pc_ of rows ub_ + 16 i
    const uint32_t mine = dst + (uint32_t)((ub_ >> 3) * 1024 + (ub_ & 7) * 128 + (pc_ << 4));
    unsigned spins = 0;
    while (true) {
      if (i128_ == 0) {
        const unsigned kbytes = nrows >= UB ? (unsigned)ABLK : (unsigned)((nrows + 7) >> 3) * 1024u;
        mbar_expect_tx(bar(pollbar), kbytes);
        bulk_g2s_plain(dst, src, kbytes, bar(pollbar));
      }
      waitb(pollbar);
      unsigned bad = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (ub_ + 16 * i < nrows) {
          uint4 x;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(mine + i * 2048) : "memory");
          bad |= (x.x ^ tag) | (x.y ^ tag) | (x.z ^ tag) | (x.w ^ tag);
        }
      if (!bar_or(named, 128, (bad & 1u) != 0) || p.nowait) break;
      spin_check(spins);
    }
  };
  // fed-back symbol of utterance u for step t >= 1
  auto poll_symbol = [&](int u, int t) -> int {
    const unsigned want = ((unsigned)(t - 1) & 1u) << 30;
    unsigned spins = 0;
    while (true) {
      const unsigned v = ld_strong_u32(p.vsym + u * 32);
      if (((v ^ want) & 0x40000000u) == 0 || p.nowait) return (int)(v & 0xFFFFu) % Q;
      spin_check(spins);
    }
  };

  if (warp < 4) {
    // ======================================================================================= ET: C + P tiles, z_j, head
    const int w = warp;                                   // TMEM lanes 32w ..: utterance 32w + lane of the group, finished by rank w
    const int lu = tid >> 2, q = tid & 3;                 // finishing role: utterance 32 rank + lu of the group, rows 8q .. 8q+7
    const uint32_t trow = tmem + ((uint32_t)(32 * w) << 16);
    int sy_c[NG], sy_p1[NG], sy_p2[NG];
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) sy_c[gi] = sy_p1[gi] = sy_p2[gi] = half;
    int nT = 0;                                           // tiles finished per group so far: receive buffer nT & 1
    const float* const T0 = p.T0 + (size_t)c * 3 * Q * 32 + 8 * q;
    const float* const T12 = p.T12 + (size_t)c * 4 * Q * 32 + 8 * q;
    const float* const bgp = p.bgate + (size_t)c * 32 + 8 * q;            // + j * NCL * 32

    auto gate4 = [&](int j, int fu, bool live, const float (&pre)[8], unsigned tag) {   // z of channels 4q .. 4q+3: one 16-byte piece per thread pair
      float z[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = fast_sigmoid(pre[2 * i]) * fast_tanh(pre[2 * i + 1]);
      const unsigned w0 = pack_tagged(z[0], z[1], tag), w1 = pack_tagged(z[2], z[3], tag);
      const unsigned o0 = __shfl_down_sync(0xffffffffu, w0, 1), o1 = __shfl_down_sync(0xffffffffu, w1, 1);
      if (!(q & 1) && live)
        st_strong_v4(p.vz + (size_t)j * UT * (C / 2) + xoff(fu, 2 * c + (q >> 1)), make_uint4(w0, w1, o0, o1));
    };

    for (int t = -NP; t < g.max_steps; ++t) {
      const unsigned tagz = (unsigned)(t + NP) & 1u;
      const float wt = __ldg(g.up_w + (t < 0 ? 0 : t % U));
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        if (gi >= ng) break;
        const int fu = UB * gi + 32 * rank + lu;
        const bool fin = rank < gnl(gi), live = fu < B;
        float* const myP = p.Paux + (((size_t)s * NG + gi) * L * 32 + lu) * 32 + 8 * q;   // + j * 1024
        // ---- aux 1x1 at frame rate: P_j = V_j h_f of (utterance fu, rows 8q..8q+7), every block, once per frame
        if (fin && (t == -NP || (t > 0 && t % U == 0))) {
          const int f = t < 0 ? 0 : t / U;
          const TensorMap tm{p.nF, p.nA};
          for (int j = 0; j < L; ++j) {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            if (live) {
              const float* V[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = 8 * q + i, gg = row & 1, ch = 16 * c + (row >> 1);
                V[i] = (j < p.nF ? p.tab[tm.auxF_w(gg, j)] : p.tab[tm.auxA_w(gg, j - p.nF)]) + (size_t)ch * p.A;
              }
              for (int a = 0; a < p.A; ++a) {
                const float hv = __ldg(g.h + ((size_t)fu * p.A + a) * p.F + f);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(__ldg(V[i] + a), hv, acc[i]);
              }
            }
            __stcg((float4*)(myP + (size_t)j * 1024), make_float4(acc[0], acc[1], acc[2], acc[3]));
            __stcg((float4*)(myP + (size_t)j * 1024 + 4), make_float4(acc[4], acc[5], acc[6], acc[7]));
          }
        }
        // ---- symbols of this step (qpnet.py:356-358: pad with Q/2, keep the seed last)
        if (t == 0) {
          sy_p2[gi] = sy_p1[gi]; sy_p1[gi] = sy_c[gi];
          sy_c[gi] = live ? (int)(((g.seed[fu] % Q) + Q) % Q) : half;
        } else if (t >= 1) {
          const int nw = (fin && live) ? poll_symbol(fu, t) : half;
          sy_p2[gi] = sy_p1[gi]; sy_p1[gi] = sy_c[gi]; sy_c[gi] = nw;
        }
        // ---- block 0: three table lookups + aux, no exchange
        trace(t, gi, 0, 2);
        if (fin) {
          float pre[8];
          const float4 pa = ldcg4(myP), pb = ldcg4(myP + 4);
          const float4 a0 = __ldg((const float4*)(T0 + (0 * Q + sy_c[gi]) * 32)), a1 = __ldg((const float4*)(T0 + (0 * Q + sy_c[gi]) * 32 + 4));
          const float4 b0 = __ldg((const float4*)(T0 + (1 * Q + sy_p1[gi]) * 32)), b1 = __ldg((const float4*)(T0 + (1 * Q + sy_p1[gi]) * 32 + 4));
          const float4 c0 = __ldg((const float4*)(T0 + (2 * Q + sy_p2[gi]) * 32)), c1 = __ldg((const float4*)(T0 + (2 * Q + sy_p2[gi]) * 32 + 4));
          const float4 g0 = __ldg((const float4*)bgp), g1 = __ldg((const float4*)(bgp + 4));
          pre[0] = a0.x + b0.x + c0.x + fmaf(wt, pa.x, g0.x); pre[1] = a0.y + b0.y + c0.y + fmaf(wt, pa.y, g0.y);
          pre[2] = a0.z + b0.z + c0.z + fmaf(wt, pa.z, g0.z); pre[3] = a0.w + b0.w + c0.w + fmaf(wt, pa.w, g0.w);
          pre[4] = a1.x + b1.x + c1.x + fmaf(wt, pb.x, g1.x); pre[5] = a1.y + b1.y + c1.y + fmaf(wt, pb.y, g1.y);
          pre[6] = a1.z + b1.z + c1.z + fmaf(wt, pb.z, g1.z); pre[7] = a1.w + b1.w + c1.w + fmaf(wt, pb.w, g1.w);
          gate4(0, fu, live, pre, tagz);
        }
        trace(t, gi, 0, 5);
      }
      // ---- blocks 1 .. L-1, the groups in turn
      for (int j = 1; j < L; ++j) {
        const int b = j & 1, pb_ = (j - 1) & 1, rbuf = nT & 1;
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          if (gi >= ng) break;
          const int nl = gnl(gi);
          const int fu = UB * gi + 32 * rank + lu;
          const bool fin = rank < nl, live = fu < B;
          const uint32_t rbar = bar(gb(gi, G_RT + rbuf));
          const uint32_t tgrp = trow + TC_GRP * gi;
          const uint32_t rt_base = SM_R + gi * RG_B;
          if (fin && tid == 0) mbar_expect_tx(rbar, 4 * 32 * 64);
          // everything the finish needs besides the partial rows: aux, bias, (blocks 1, 2) the x_0 tables
          float pre[8];
          if (fin) {
            const float* myP = p.Paux + (((size_t)s * NG + gi) * L * 32 + lu) * 32 + 8 * q + (size_t)j * 1024;
            const float4 pa = ldcg4(myP), pb = ldcg4(myP + 4);
            const float4 g0 = __ldg((const float4*)(bgp + (size_t)j * NCL * 32)), g1 = __ldg((const float4*)(bgp + (size_t)j * NCL * 32 + 4));
            pre[0] = fmaf(wt, pa.x, g0.x); pre[1] = fmaf(wt, pa.y, g0.y); pre[2] = fmaf(wt, pa.z, g0.z); pre[3] = fmaf(wt, pa.w, g0.w);
            pre[4] = fmaf(wt, pb.x, g1.x); pre[5] = fmaf(wt, pb.y, g1.y); pre[6] = fmaf(wt, pb.z, g1.z); pre[7] = fmaf(wt, pb.w, g1.w);
            if (j <= 2) {
              const float* Ta = T12 + ((size_t)(j - 1) * 2 + 0) * Q * 32 + sy_c[gi] * 32;
              const float* Tb = T12 + ((size_t)(j - 1) * 2 + 1) * Q * 32 + sy_p1[gi] * 32;
              const float4 u0 = __ldg((const float4*)Ta), u1 = __ldg((const float4*)(Ta + 4));
              const float4 v0 = __ldg((const float4*)Tb), v1 = __ldg((const float4*)(Tb + 4));
              pre[0] += u0.x + v0.x; pre[1] += u0.y + v0.y; pre[2] += u0.z + v0.z; pre[3] += u0.w + v0.w;
              pre[4] += u1.x + v1.x; pre[5] += u1.y + v1.y; pre[6] += u1.z + v1.z; pre[7] += u1.w + v1.w;
            }
          }
          waitb(gb(gi, G_PFULL + pb_));                   // P_j: built during the previous phase
          waitb(gb(gi, G_CFULL + b));                     // C_j = G_j z_{j-1}
          tc_fence_after();
          trace(t, gi, j, 2);
          if (w < nl) {
            unsigned hw[16];                              // the 32 partial rows of (utterance 32w + lane) as fp16 pairs
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t cv[16], pv[16];
              tmem_ld16x2(tgrp + TC_BUF * b + TC_C + 16 * hh, cv, tgrp + TC_BUF * pb_ + TC_P + 16 * hh, pv);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                hw[8 * hh + i] = pack_h2(__uint_as_float(cv[2 * i]) + __uint_as_float(pv[2 * i]),
                                         __uint_as_float(cv[2 * i + 1]) + __uint_as_float(pv[2 * i + 1]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar(gb(gi, G_CFREE + b))); mbar_arrive(bar(gb(gi, G_PFREE + pb_))); }
            const uint32_t dst = sbase + rt_base + (uint32_t)(((rbuf * 4 + rank) * 32 + lane) * 64);
            const uint32_t rdst = mapa(dst, (unsigned)w), rb = mapa(rbar, (unsigned)w);
#pragma unroll
            for (int i = 0; i < 4; ++i) st_async_v4(rdst + 16 * i, make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]), rb);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar(gb(gi, G_CFREE + b))); mbar_arrive(bar(gb(gi, G_PFREE + pb_))); }
          }
          trace(t, gi, j, 3);
          if (fin) {
            waitb(gb(gi, G_RT + rbuf));
            trace(t, gi, j, 4);
            gtrace(t, gi, j, 2);
            const unsigned char* rp = sm + rt_base + ((rbuf * 4) * 32 + lu) * 64 + 16 * q;
#pragma unroll
            for (int src = 0; src < 4; ++src) {
              const uint4 h4 = *(const uint4*)(rp + src * 32 * 64);
              const float2 f0 = unpack_h2(h4.x), f1 = unpack_h2(h4.y), f2 = unpack_h2(h4.z), f3 = unpack_h2(h4.w);
              pre[0] += f0.x; pre[1] += f0.y; pre[2] += f1.x; pre[3] += f1.y;
              pre[4] += f2.x; pre[5] += f2.y; pre[6] += f3.x; pre[7] += f3.y;
            }
            gate4(j, fu, live, pre, tagz);
            gtrace(t, gi, j, 0);
          }
          trace(t, gi, j, 5);
        }
        ++nT;
      }
      // ---- head 1, head 2 (qpnet.py:566-571), real steps only
      if (t >= 0) {
        const unsigned par_t = (unsigned)t & 1u;
        for (int hd = 0; hd < 2; ++hd) {
#pragma unroll
          for (int gi = 0; gi < NG; ++gi) {
            if (gi >= ng) break;
            const int nl = gnl(gi);
            const int fu = UB * gi + 32 * rank + lu;
            const bool fin = rank < nl, live = fu < B;
            const uint32_t rbar = bar(gb(gi, G_RH + hd));
            const uint32_t rh_base = SM_R + gi * RG_B + RT_B + RU_B;
            if (fin && tid == 0) mbar_expect_tx(rbar, 4 * 32 * 16);
            waitb(gb(gi, G_HFULL + hd));
            tc_fence_after();
            if (w < nl) {
              uint32_t v[8];
              tmem_ld8(trow + TC_H + 32 * gi + 16 * hd, v);
              const uint32_t dst = sbase + rh_base + (uint32_t)(((hd * 4 + rank) * 32 + lane) * 16);
              st_async_v4(mapa(dst, (unsigned)w),
                          make_uint4(pack_h2(__uint_as_float(v[0]), __uint_as_float(v[1])), pack_h2(__uint_as_float(v[2]), __uint_as_float(v[3])),
                                     pack_h2(__uint_as_float(v[4]), __uint_as_float(v[5])), pack_h2(__uint_as_float(v[6]), __uint_as_float(v[7]))),
                          mapa(rbar, (unsigned)w));
            }
            tc_fence_before();
            if (fin) {
              waitb(gb(gi, G_RH + hd));
              float s0 = sBh[hd * 8 + 2 * q], s1 = sBh[hd * 8 + 2 * q + 1];
              const unsigned char* rp = sm + rh_base + ((hd * 4) * 32 + lu) * 16 + 4 * q;
#pragma unroll
              for (int src = 0; src < 4; ++src) {
                const float2 f = unpack_h2(*(const unsigned*)(rp + src * 32 * 16));
                s0 += f.x; s1 += f.y;
              }
              if (hd == 0) {   // relu(head-1) rows 8c + 2q, +1 -> one 16-byte piece per utterance
                const unsigned w0 = pack_tagged(fmaxf(s0, 0.f), fmaxf(s1, 0.f), par_t);
                const unsigned w1 = __shfl_down_sync(0xffffffffu, w0, 1), w2 = __shfl_down_sync(0xffffffffu, w0, 2),
                               w3 = __shfl_down_sync(0xffffffffu, w0, 3);
                if (q == 0 && live) st_strong_v4(p.v256 + (size_t)UT * (S / 2) + hoff(fu, c), make_uint4(w0, w1, w2, w3));
              } else if (live) {
                st_strong_v2(p.vlog + (size_t)fu * Q + 8 * c + 2 * q, (__float_as_uint(s0) & ~1u) | par_t, (__float_as_uint(s1) & ~1u) | par_t);
              }
            }
          }
        }
      }
    }
  } else if (warp < 8) {
    // ======================================================================================= EU: U tiles, x_j, skip
    const int w = warp - 4, t128 = tid - 128;
    const int lu = t128 >> 2, q = t128 & 3;               // finishing role: residual channels 4q..4q+3, skip rows 2q, 2q+1
    const uint32_t trow = tmem + ((uint32_t)(32 * w) << 16);
    int sy_c[NG], sy_p1[NG];
    float xc[NG][4], sk0[NG], sk1[NG];                    // fp32 residual stream and skip sums of (group, utterance, channels)
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) { sy_c[gi] = sy_p1[gi] = half; sk0[gi] = sk1[gi] = 0.f; xc[gi][0] = xc[gi][1] = xc[gi][2] = xc[gi][3] = 0.f; }
    int nU = 0;
    const float* const Eo = p.Eo + (size_t)c * 2 * Q * 16 + 4 * q;
    const float* const brp = p.bres + (size_t)c * 32;     // + l * NCL * 32
    for (int t = -NP; t < g.max_steps; ++t) {
      const unsigned tagx = (unsigned)(t + NP) & 1u;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        if (gi >= ng) break;
        const int fu = UB * gi + 32 * rank + lu;
        const bool fin = rank < gnl(gi), live = fu < B;
        if (t == 0) {
          sy_p1[gi] = sy_c[gi];
          sy_c[gi] = live ? (int)(((g.seed[fu] % Q) + Q) % Q) : half;
        } else if (t >= 1) {
          const int nw = (fin && live) ? poll_symbol(fu, t) : half;
          sy_p1[gi] = sy_c[gi]; sy_c[gi] = nw;
        }
        // fp32 residual stream restarts from the causal layer x_0 (qpnet.py:447-448), skip sums restart
        sk0[gi] = sk1[gi] = 0.f;
        if (fin) {
          const float4 e0 = __ldg((const float4*)(Eo + (0 * Q + sy_p1[gi]) * 16)), e1 = __ldg((const float4*)(Eo + (1 * Q + sy_c[gi]) * 16));
          xc[gi][0] = e0.x + e1.x; xc[gi][1] = e0.y + e1.y; xc[gi][2] = e0.z + e1.z; xc[gi][3] = e0.w + e1.w;
        }
      }
      for (int j = 1; j <= L; ++j) {
        const int tb = j & 1, rb3 = nU % 3;
        const int l = j - 1;                              // block whose residual / skip projection this is
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          if (gi >= ng) break;
          const int nl = gnl(gi);
          const int fu = UB * gi + 32 * rank + lu;
          const bool fin = rank < nl, live = fu < B;
          const uint32_t rbar = bar(gb(gi, G_RU + rb3));
          const uint32_t ru_base = SM_R + gi * RG_B + RT_B;
          if (fin && t128 == 0) mbar_expect_tx(rbar, 4 * 32 * 48);
          float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, k0 = 0.f, k1 = 0.f;
          if (fin) {
            const float4 br4 = __ldg((const float4*)(brp + (size_t)l * NCL * 32 + 4 * q));
            const float2 bk2 = __ldg((const float2*)(brp + (size_t)l * NCL * 32 + 16 + 2 * q));
            r0 = br4.x; r1 = br4.y; r2 = br4.z; r3 = br4.w; k0 = bk2.x; k1 = bk2.y;
          }
          waitb(gb(gi, G_UFULL + tb));
          tc_fence_after();
          trace(t, gi, j, 13);
          if (w < nl) {
            uint32_t v0[16], v1[8];
            tmem_ld16(trow + TC_GRP * gi + TC_BUF * tb + TC_U, v0);
            tmem_ld8(trow + TC_GRP * gi + TC_BUF * tb + TC_U + 16, v1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(gb(gi, G_UFREE + tb)));
            const uint32_t dst = sbase + ru_base + (uint32_t)(((rb3 * 4 + rank) * 32 + lane) * 48);
            const uint32_t rdst = mapa(dst, (unsigned)w), rb = mapa(rbar, (unsigned)w);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              st_async_v4(rdst + 16 * i,
                          make_uint4(pack_h2(__uint_as_float(v0[8 * i]), __uint_as_float(v0[8 * i + 1])),
                                     pack_h2(__uint_as_float(v0[8 * i + 2]), __uint_as_float(v0[8 * i + 3])),
                                     pack_h2(__uint_as_float(v0[8 * i + 4]), __uint_as_float(v0[8 * i + 5])),
                                     pack_h2(__uint_as_float(v0[8 * i + 6]), __uint_as_float(v0[8 * i + 7]))), rb);
            st_async_v4(rdst + 32,
                        make_uint4(pack_h2(__uint_as_float(v1[0]), __uint_as_float(v1[1])), pack_h2(__uint_as_float(v1[2]), __uint_as_float(v1[3])),
                                   pack_h2(__uint_as_float(v1[4]), __uint_as_float(v1[5])), pack_h2(__uint_as_float(v1[6]), __uint_as_float(v1[7]))), rb);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(gb(gi, G_UFREE + tb)));
          }
          trace(t, gi, j, 14);
          if (fin) {
            waitb(gb(gi, G_RU + rb3));
            trace(t, gi, j, 15);
            const unsigned char* rp = sm + ru_base + ((rb3 * 4) * 32 + lu) * 48;
#pragma unroll
            for (int src = 0; src < 4; ++src) {
              const uint2 h2 = *(const uint2*)(rp + src * 32 * 48 + 8 * q);
              const unsigned hk = *(const unsigned*)(rp + src * 32 * 48 + 32 + 4 * q);
              const float2 f0 = unpack_h2(h2.x), f1 = unpack_h2(h2.y), fk = unpack_h2(hk);
              r0 += f0.x; r1 += f0.y; r2 += f1.x; r3 += f1.y; k0 += fk.x; k1 += fk.y;
            }
            sk0[gi] += k0; sk1[gi] += k1;
            if (j < L) {   // residual projection + current input (qpnet.py:669 / 639): x_j, the input of block j
              xc[gi][0] += r0; xc[gi][1] += r1; xc[gi][2] += r2; xc[gi][3] += r3;
              const unsigned w0 = pack_tagged(xc[gi][0], xc[gi][1], tagx), w1 = pack_tagged(xc[gi][2], xc[gi][3], tagx);
              const unsigned o0 = __shfl_down_sync(0xffffffffu, w0, 1), o1 = __shfl_down_sync(0xffffffffu, w1, 1);
              if (!(q & 1) && live) {
                const uint4 piece = make_uint4(w0, w1, o0, o1);
                st_strong_v4(p.vx + (size_t)j * UT * (C / 2) + xoff(fu, 2 * c + (q >> 1)), piece);   // this step's consumers poll here
                // the past taps of later steps read here.  A ring slot is stored as the A tiles its readers stage,
                // [CL K-shares][2 K-blocks][UT rows][128 B], the 16-byte pieces of a row in SWIZZLE_128B order (piece ^ (row & 7)):
                // the tile of a block with a fixed look-back is two contiguous 16 KB runs
                const int pw = 2 * c + (q >> 1);
                uint32_t* dstp = p.xr[j] + ((size_t)(pw >> 3) * UT + fu) * 32 + 4 * ((pw ^ fu) & 7);
                if (t == -1) {   // the last priming pass fills the whole ring with the constant
                  const int R = 1 << p.rlog[j];
                  for (int sl = 0; sl < R; ++sl) st_strong_v4(dstp + (size_t)sl * UT * (C / 2), piece);
                } else {
                  st_strong_v4(dstp + (size_t)x_slot(j, t) * UT * (C / 2), piece);
                }
              }
              gtrace(t, gi, j, 3);
            } else if (t >= 0) {   // relu(sum of the skip outputs) (qpnet.py:505, 566-567) rows 8c + 2q, +1
              const unsigned par_t = (unsigned)t & 1u;
              const unsigned w0 = pack_tagged(fmaxf(sk0[gi], 0.f), fmaxf(sk1[gi], 0.f), par_t);
              const unsigned w1 = __shfl_down_sync(0xffffffffu, w0, 1), w2 = __shfl_down_sync(0xffffffffu, w0, 2),
                             w3 = __shfl_down_sync(0xffffffffu, w0, 3);
              if (q == 0 && live) st_strong_v4(p.v256 + hoff(fu, c), make_uint4(w0, w1, w2, w3));
            }
          }
          trace(t, gi, j, 16);
        }
        ++nU;
      }
    }
  } else if (warp < 12) {
    // ======================================================================================= PZ: z_{j-1} / head vectors -> A tile
    const int i128 = tid - 256;
    const uint32_t sZ = sbase + SM_Z;
    auto done = [&]() { fence_proxy_async(); mbar_arrive(bar(B_ZFULL)); };
    for (int t = -NP; t < g.max_steps; ++t) {
      const unsigned tagz = (unsigned)(t + NP) & 1u;
      for (int j = 1; j <= L; ++j) {
        for (int gi = 0; gi < ng; ++gi) {
          waitb(B_ZFREE);
          trace(t, gi, j, 11);
          poll_bulk(p.vz + (size_t)(j - 1) * UT * (C / 2) + (size_t)(gi * 8 + rank * 2) * UB * 32, sZ, B_ZPOLL, 2, i128, gB(gi), tagz);
          trace(t, gi, j, 6);
          gtrace(t, gi, j, 1);
          done();
        }
      }
      if (t >= 0) {
        const unsigned par_t = (unsigned)t & 1u;
        for (int hd = 0; hd < 2; ++hd) {
          for (int gi = 0; gi < ng; ++gi) {
            waitb(B_ZFREE);
            poll_bulk_head(p.v256 + (size_t)hd * UT * (S / 2) + (size_t)(gi * 4 + rank) * UB * 32, sZ, B_ZPOLL, 2, i128, gB(gi), par_t);
            done();
          }
        }
      }
    }
  } else if (warp < 16) {
    // ======================================================================================= PX: x_{j-1} and past taps -> A tiles
    const int i128 = tid - 384;
    const int pc = i128 & 15, ub = i128 >> 4;             // piece pc of utterances ub + 8 i, i < 16
    const long long ldd = (long long)p.F * U;
    const uint32_t tile_off = (uint32_t)((pc >> 3) * ABLK + ub * 128 + (((pc & 7) ^ ub) << 4));   // + i * 1024
    // past tap of block jb for step t: x_jb(t - k) of every utterance of the group -> XP tile, asynchronously: the copies
    // complete on the mbarrier the MMA thread waits for, this thread moves on
    auto stage_xp = [&](int jb, int gi, int t) {
      waitb(B_XPFREE);
      trace(t, gi, jb - 1, 17);
      if (t == -NP) {   // the first priming pass has no ring contents yet
        unsigned char* dst = sm + SM_XP + tile_off;
#pragma unroll
        for (int i = 0; i < 16; ++i) *(uint4*)(dst + i * 1024) = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
        mbar_arrive(bar(B_XPFULL));
      } else {
        const uint32_t* ring = p.xr[jb] + ((size_t)rank * 2 * UT + UB * gi) * 32;    // + slot * UT * (C / 2) + kb * UT * 32 + row * 32
        const int rmask = (1 << p.rlog[jb]) - 1;
        if (jb < p.nF) {
          // fixed look-back: every utterance reads the same slot, the tile is two contiguous 16 KB runs of the ring
          if (i128 == 0) {
            const uint32_t* src = ring + (size_t)(t >= 0 ? ((t - p.dil[jb]) & rmask) : ((t - 1) & 1)) * UT * (C / 2);
            mbar_expect_tx(bar(B_XPFULL), 2 * ABLK);
            bulk_g2s_plain(sbase + SM_XP, src, ABLK, bar(B_XPFULL));
            bulk_g2s_plain(sbase + SM_XP + ABLK, src + (size_t)UT * 32, ABLK, bar(B_XPFULL));
          } else {
            mbar_arrive(bar(B_XPFULL));
          }
        } else {
          // pitch-dependent look-backs of this step: piece by piece, one slot per utterance
          const unsigned short* kt = sK + (gi * MAXA + (jb - p.nF)) * UB;
          const int Bg = gB(gi);
          const uint32_t dst = sbase + SM_XP + (uint32_t)((pc >> 3) * ABLK + ub * 128 + ((pc & 7) << 4));
          const uint32_t* src = ring + (size_t)(pc >> 3) * UT * 32 + 4 * (pc & 7);
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int u = ub + 8 * i;
            if (u < Bg) {
              const int slot = t >= 0 ? ((t - (int)kt[u]) & rmask) : ((t - 1) & 1);
              cp_async16_s(dst + i * 1024, src + (size_t)slot * UT * (C / 2) + u * 32);
            }
          }
          cp_async_arrive_noinc(bar(B_XPFULL));
        }
      }
      trace(t, gi, jb - 1, 18);
    };
    // x_jx of the current step (published during the previous phase) -> X tile
    auto stage_x = [&](int jx, int gi, int t) {
      waitb(B_XFREE);
      trace(t, gi, jx + 1, 19);
      poll_bulk(p.vx + (size_t)jx * UT * (C / 2) + (size_t)(gi * 8 + rank * 2) * UB * 32, sbase + SM_X, B_XPOLL, 3, i128, gB(gi), (unsigned)(t + NP) & 1u);
      trace(t, gi, jx + 1, 21);
      fence_proxy_async();
      mbar_arrive(bar(B_XFULL));
    };
    for (int t = -NP; t < g.max_steps; ++t) {
      if (t >= 0 && p.nA > 0) {
        // look-backs of the adaptive blocks for this step: k = -round(-d[t] * dil) with the reference's rounding
        // (qpnet.py:476-483, 613-624); k == 0 (python index 0) and anything beyond the FIFO select the oldest entry (C4)
        asm volatile("bar.sync 1, 128;\n" ::: "memory");         // every PX thread is done with the previous step's table
        for (int gi = 0; gi < ng; ++gi) {
          const int u = UB * gi + i128;
          if (u < B) {
            for (int a = 0; a < p.nA; ++a) {
              const int dl = p.dil[p.nF + a], dep = p.depth[p.nF + a];
              int k = g.d_is_f64 ? -gen_index_f64(__ldg((const double*)g.d + (long long)u * ldd + t), dl)
                                 : -gen_index_f32(__ldg((const float*)g.d + (long long)u * ldd + t), dl);
              if (k <= 0 || k > dep) k = dep;
              sK[(gi * MAXA + a) * UB + i128] = (unsigned short)k;
            }
          }
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
      }
      // Order inside a phase (it must not contradict the MMA thread's: z-products of both groups, Wc x of both, Wp x of
      // both -- the second past-tap tile can only be staged once the first one is consumed, which needs both x tiles)
      for (int gi = 0; gi < ng; ++gi) stage_xp(1, gi, t);
      for (int j = 1; j <= L - 2; ++j) {
        stage_xp(j + 1, 0, t);
        if (j >= 2) for (int gi = 0; gi < ng; ++gi) stage_x(j - 1, gi, t);
        if (ng > 1) stage_xp(j + 1, 1, t);
      }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  } else if (warp == 16) {
    // ======================================================================================= MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t sZ = sbase + SM_Z, sX = sbase + SM_X, sXP = sbase + SM_XP;
      const uint32_t sWZ = sbase + SM_WZ, sWC = sbase + SM_WC, sWP = sbase + SM_WP;
      // D[128 x N] (+)= A[128 x 128] B[N x 128]^T: two 64-wide K-blocks of four K = 16 steps
      auto mma = [&](int N, uint32_t dcol, uint32_t a_base, uint32_t b_base, uint32_t b_kblock, bool fresh) {
        fence_proxy_async();   // generic-proxy stores of the pollers -> tcgen05 operand reads (async proxy)
        tc_fence_after();
        const uint32_t idesc = umma_idesc(128, N);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t ad = umma_desc(a_base + kb * ABLK), bd = umma_desc(b_base + kb * b_kblock);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma(tmem + dcol, ad + 2 * ks, bd + 2 * ks, idesc, (fresh && kb == 0 && ks == 0) ? 0u : 1u);
        }
      };
      // Issue order: fixed, Z(j,0) P(j,0) Z(j,1) P(j,1); every role stages and consumes in the same order.  (Orders in which the
      // x-products may fall behind the next z-products -- a second issuing thread, a non-blocking scheduler -- were not
      // faster, and before the priming passes alternated ring slots (x_slot) they changed what priming converged through.)
      for (int t = -NP; t < g.max_steps; ++t) {
        {   // phase 0: P_1 (buffer 0) <- Wp_1 x_1(t-k), both groups on one chunk
          waitb(B_WPW_FULL);
          for (int gi = 0; gi < ng; ++gi) {
            waitb(gb(gi, G_PFREE + 0)); waitb(B_XPFULL);
            mma(32, TC_GRP * gi + TC_P, sXP, sWP, WCH_B / 2, true);
            umma_commit(bar(B_XPFREE)); umma_commit(bar(gb(gi, G_PFULL + 0)));
          }
          umma_commit(bar(B_WPW_FREE));
        }
        for (int j = 1; j < L; ++j) {
          const int b = j & 1;
          const bool pre = j <= L - 2;                    // this phase also builds P_{j+1}
          waitb(B_ZPW_FULL);
          for (int gi = 0; gi < ng; ++gi) {
            const uint32_t dbuf = TC_GRP * gi + TC_BUF * b;
            waitb(gb(gi, G_CFREE + b)); waitb(gb(gi, G_UFREE + b));
            if (pre) waitb(gb(gi, G_PFREE + b));
            waitb(B_ZFULL);
            trace(t, gi, j, 0);
            // [C_j | U_j | P_{j+1}] <- [G_j ; [R;K]_{j-1} ; H_{j+1}] z_{j-1}: ONE sequence of 8 instructions (N = 96).  The
            // single-group kernel issues C_j on its own first (ready ~300 cycles earlier); here the issuing thread of the
            // slowest CTAs is what every other CTA waits for, and 8 instructions instead of 16 is worth more
            mma(pre ? 96 : 64, dbuf + TC_C, sZ, sWZ, ZPC_B / 2, true);
            trace(t, gi, j, 24);
            umma_commit(bar(gb(gi, G_CFULL + b)));
            trace(t, gi, j, 1);
            umma_commit(bar(B_ZFREE));
            trace(t, gi, j, 25);
            umma_commit(bar(gb(gi, G_UFULL + b)));
            trace(t, gi, j, 26);
            if (gi == ng - 1) umma_commit(bar(B_ZPW_FREE));
            trace(t, gi, j, 27);
            // the x-products of this group go between the groups' z-products: the other group's z tile is staged meanwhile
            if (pre) {
              if (j >= 2) {   // P_{j+1} += Wc_{j+1} x_{j-1}
                if (gi == 0) waitb(B_WCW_FULL);
                trace(t, gi, j, 22);
                waitb(B_XFULL);
                trace(t, gi, j, 8);
                mma(32, dbuf + TC_P, sX, sWC, WCH_B / 2, false);
                trace(t, gi, j, 28);
                umma_commit(bar(B_XFREE));
                trace(t, gi, j, 29);
                if (gi == ng - 1) umma_commit(bar(B_WCW_FREE));
              }
              // P_{j+1} += Wp_{j+1} x_{j+1}(t-k)
              if (gi == 0) waitb(B_WPW_FULL);
              trace(t, gi, j, 23);
              waitb(B_XPFULL);
              trace(t, gi, j, 9);
              mma(32, dbuf + TC_P, sXP, sWP, WCH_B / 2, false);
              umma_commit(bar(B_XPFREE)); umma_commit(bar(gb(gi, G_PFULL + b)));
              if (gi == ng - 1) umma_commit(bar(B_WPW_FREE));
              trace(t, gi, j, 10);
            }
          }
        }
        {   // phase L: U_L <- [R ; K]_{L-1} z_{L-1} (skip rows; the residual rows of the last block are dead, C7)
          const int b = L & 1;
          waitb(B_ZPW_FULL);
          for (int gi = 0; gi < ng; ++gi) {
            waitb(gb(gi, G_UFREE + b)); waitb(B_ZFULL);
            trace(t, gi, L, 0);
            mma(32, TC_GRP * gi + TC_BUF * b + TC_U, sZ, sWZ + 4096, ZPC_B / 2, true);
            umma_commit(bar(B_ZFREE)); umma_commit(bar(gb(gi, G_UFULL + b)));
          }
          umma_commit(bar(B_ZPW_FREE));
        }
        if (t >= 0) {
          for (int hd = 0; hd < 2; ++hd) {   // head layers: 16 rows (8 live) x K-share 64, resident weights
            for (int gi = 0; gi < ng; ++gi) {
              waitb(B_ZFULL);
              fence_proxy_async();
              tc_fence_after();
              const uint32_t idesc16 = umma_idesc(128, 16);
              const uint64_t ad = umma_desc(sZ), bd = umma_desc(sbase + SM_WH + hd * HCH_B);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma(tmem + TC_H + 32 * gi + 16 * hd, ad + 2 * ks, bd + 2 * ks, idesc16, ks ? 1u : 0u);
              umma_commit(bar(gb(gi, G_HFULL + hd))); umma_commit(bar(B_ZFREE));
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ======================================================================================= weight loader (one thread)
    if (lane == 0) {
      const uint64_t pol_keep = l2_policy_evict_last();
      auto load = [&](int free_bar, int full_bar, uint32_t dst, const __nv_bfloat16* src, unsigned bytes) {
        waitb(free_bar);
        mbar_expect_tx(bar(full_bar), bytes);
        bulk_g2s(dst, src, bytes, bar(full_bar), pol_keep);
      };
      const __nv_bfloat16* const Wc = p.Wcp + (size_t)s * WCH_E;                        // + j * NCTA * WCH_E
      const __nv_bfloat16* const Wp = p.Wcp + ((size_t)L * NCTA + s) * WCH_E;
      for (int t = -NP; t < g.max_steps; ++t) {
        load(B_WPW_FREE, B_WPW_FULL, sbase + SM_WP, Wp + (size_t)1 * NCTA * WCH_E, WCH_B);
        for (int j = 1; j <= L; ++j) {
          load(B_ZPW_FREE, B_ZPW_FULL, sbase + SM_WZ, p.Wzp + ((size_t)(j - 1) * NCTA + s) * ZPC_E, ZPC_B);
          if (j >= 2 && j <= L - 2) load(B_WCW_FREE, B_WCW_FULL, sbase + SM_WC, Wc + (size_t)(j + 1) * NCTA * WCH_E, WCH_B);
          if (j <= L - 2) load(B_WPW_FREE, B_WPW_FULL, sbase + SM_WP, Wp + (size_t)(j + 1) * NCTA * WCH_E, WCH_B);
        }
      }
    }
    __syncwarp();
  } else {
    // ======================================================================================= sampler: utterances blockIdx.x + 128 g
    for (int t = 0; t < g.max_steps; ++t) {
      const unsigned par_t = (unsigned)t & 1u;
      for (int gi = 0; gi < ng; ++gi) {
        const int u = UB * gi + s;
        if (u >= B) break;
        float v[8];
        {
          const uint4* src = (const uint4*)(p.vlog + (size_t)u * Q) + 2 * lane;
          uint4 a, b;
          unsigned pend = 3;
          unsigned spins = 0;
          while (pend) {
            if (pend & 1u) { a = ld_strong_v4(src); if (fresh4(a, par_t) || p.nowait) pend &= ~1u; }
            if (pend & 2u) { b = ld_strong_v4(src + 1); if (fresh4(b, par_t) || p.nowait) pend &= ~2u; }
            if (pend) spin_check(spins);
          }
          v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
          v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
        }
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
          const float om = __shfl_xor_sync(0xffffffffu, wmx, o);
          const int oa = __shfl_xor_sync(0xffffffffu, wam, o);
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
            const float nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
          }
          const float total = __shfl_sync(0xffffffffu, incl, 31);
          const float uu = g.uniforms ? g.uniforms[(long long)u * g.ld_uniforms + t]
                                      : philox_uniform(g.philox_seed, g.utt_ids ? (unsigned)g.utt_ids[u] : (unsigned)u, t);
          const float target = uu * total;
          float run = incl - local;
          int cnt = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) { run += v[j]; if (run <= target) ++cnt; }
          for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
          sym = min(cnt, Q - 1);
        }
        if (lane == 0) {
          if (t < g.n_samples[u]) {
            g.out[(long long)u * g.ld_out + t] = sym;
            if (g.out_pcm) g.out_pcm[(long long)u * g.ld_out_pcm + t] = g.pcm_lut[sym];
          }
          const int fed = g.force ? g.force[(long long)u * g.ld_force + t] : sym;
          st_strong_u32(p.vsym + u * 32, ((unsigned)fed & 0xFFFFu) | (par_t << 30));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // no CTA of the cluster leaves while a peer may still write into its shared memory
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TM_COLS) : "memory");
}

}  // namespace f3x2
}  // namespace qp

using namespace qp;

namespace qp {

int pcm_lut_fill(int16_t* lut, int n_quantize, cudaStream_t st);   // qp_util.cu: mu-law decode -> int16 PCM of every symbol

// Launches the tcgen05 generator.  Returns QP_OK, an error, or +1 when the device cannot keep the 32 clusters
// co-resident (the caller then uses another kernel).
int f3x2_generate(const QpArch* arch, const float* const* tensors_host, const QpGenerateArgs* a, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  f3x2::Plan p;
  size_t need = f3x2::make_plan(arch, a->B, a->F, a->M, ws, ws_bytes, &p);
  if (need > ws_bytes) return set_error(QP_EWORKSPACE, "generate: workspace %zu < %zu bytes", ws_bytes, need);
  const int smem = f3x2::SM_END + 1024;
  for (int l = 0; l < p.L; ++l) QP_REQUIRE(p.depth[l] < 65536, "generate: look-back %d of block %d does not fit the 16-bit table", p.depth[l], l);
  QP_REQUIRE(smem <= 227 * 1024, "generate: %d bytes of shared memory needed", smem);
  const bool tr = getenv("QPNET_GEN_TRACE_STEP") != nullptr;
  if (tr) p.trace_step0 = atoi(getenv("QPNET_GEN_TRACE_STEP"));
  if (const char* e = getenv("QPNET_GEN_TRACE_CTA")) p.trace_cta = atoi(e);
  p.nowait = getenv("QPNET_F3_NOWAIT") ? 1 : 0;
  auto kern = tr ? f3x2::f3x2_gen_kernel<true> : f3x2::f3x2_gen_kernel<false>;
  QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(f3x2::NCTA); cfg.blockDim = dim3(f3x2::NT); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = f3x2::CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  QP_CUDA(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
  if (ncl < f3x2::NCL) return 1;
  cfg.numAttrs = getenv("QPNET_GEN_NOCOOP") ? 1 : 2;   // (profilers that cannot replay cooperative cluster launches)

  QP_CUDA(cudaMemsetAsync(p.status, 0, 256, st));
  QP_CUDA(cudaMemsetAsync(p.tagged_begin, 0xFF, p.tagged_bytes, st));   // every word starts with a stale tag
  QP_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(long long) * 8 * (p.L + 4) * f3x2::TRACE_EVENTS, st));
  QP_CUDA(cudaMemsetAsync(p.gtrace, 0, sizeof(long long) * f3x2::NCTA * (p.L + 4) * 8, st));
  if (int e = upload_tensor_table(arch, tensors_host, p.tab, st)) return e;
  TensorMap tm = tensor_map(arch);
  QP_CUDA(cudaMemsetAsync(p.Wzp, 0, sizeof(__nv_bfloat16) * (size_t)p.L * f3x2::NCTA * f3x2::ZPC_E, st));
  f3x2::pack_kernel<<<148 * 8, 256, 0, st>>>(tm, p, p.tab);
  QP_LAUNCH_CHECK();
  f3x2::fold_kernel<<<dim3(f3x2::NCL * 4, p.L, 2), 256, 0, st>>>(tm, p, p.tab);
  QP_LAUNCH_CHECK();
  f3x2::table_kernel<<<dim3(32, f3x2::NCL, 3), f3x2::Q, 0, st>>>(tm, p, p.tab);
  QP_LAUNCH_CHECK();
  GenArgsDev g;
  g.seed = a->seed; g.h = a->h; g.d = a->d; g.n_samples = a->n_samples;
  g.uniforms = a->uniforms; g.ld_uniforms = a->ld_uniforms; g.philox_seed = a->philox_seed;
  g.force = a->force; g.ld_force = a->ld_force; g.utt_ids = a->utt_ids;
  g.out = a->out; g.ld_out = a->ld_out; g.logits_out = a->logits_out;
  g.out_pcm = a->out_pcm; g.ld_out_pcm = a->ld_out_pcm; g.pcm_lut = p.pcm_lut;
  if (a->out_pcm)
    if (int e = pcm_lut_fill(p.pcm_lut, f3x2::Q, st)) return e;
  g.mode = a->mode; g.max_steps = a->max_steps; g.d_is_f64 = a->d_is_f64;
  g.causal_b = tensors_host[tm.causal_b()]; g.up_w = tensors_host[tm.up_w()]; g.up_b = tensors_host[tm.up_b()];
  QP_CUDA(cudaLaunchKernelEx(&cfg, kern, p, g));
  count_launch();
  return QP_OK;
}

size_t f3x2_workspace_bytes(const QpArch* arch, int B, int M) {
  f3x2::Plan p;
  return f3x2::make_plan(arch, B, 1, M, nullptr, 0, &p);
}

bool f3x2_supported(const QpArch* arch, int B) { return f3x2::supported(arch, B); }

int f3x2_trace_copy(const QpArch* arch, int B, int M, void* ws, size_t ws_bytes, long long* out_host, int n, cudaStream_t st) {
  f3x2::Plan p;
  f3x2::make_plan(arch, B, 1, M, ws, ws_bytes, &p);
  // the per-phase clock64 trace of CTA 0, followed by the %globaltimer trace of every CTA (contiguous in the workspace)
  int total = 8 * (p.L + 4) * f3x2::TRACE_EVENTS + f3x2::NCTA * (p.L + 4) * 8;
  if (n > total) n = total;
  QP_CUDA(cudaMemcpyAsync(out_host, p.trace, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
  QP_CUDA(cudaStreamSynchronize(st));
  return n;
}

}  // namespace qp
