// Aggregation kernel for the ReaRev hot shape (both directions, bf16-plane output), |v| variant.
//
// Same math and same CSR / tile decomposition as aggregate.cu (reference: ReasonGNNLayer.reason_layer /
// reason_layer_inv, gnn/modules/kg_reasoning/reasongnn.py:61-116), restructured around what the ncu source view of
// that kernel showed (profiles/README.md): at 333 warp instructions per (row, direction) only a third were the
// FFMA2s that do the work -- 14 % were FMNMX (relu of every gathered table element; 24 % of the stall samples) and
// ~38 % were address / predicate / loop scaffolding.
//
//   * relu(v) = (v + |v|) / 2, and |.| is a free source modifier of FFMA2 on sm_100 (SASS: FFMA2 R, |R|.F32x2, ...).
//     The edge loop accumulates  S = sum c*v  and  Q = sum c*|v|  -- 8 FFMA2 per gathered edge and lane, no FMNMX --
//     and the epilogue uses  sum c*relu(v) = (Q+S)/2,  sum c*relu(-v) = (Q-S)/2  (the 1/2 is folded into the staged
//     relu(+-ins)).  If every v of a row is >= 0 the two chains execute bit-identical operations, so Q-S == 0 exactly
//     (and Q+S == 0 exactly if every v <= 0): exact zeros stay exact zeros.
//     (First attempt, kept in the history: pre-split tables relu(P) | relu(-P).  It doubles the gathered bytes and the
//     gather is L2-bandwidth bound: 250 us instead of 154 us.)
//   * the table is copied once per layer to a 256-column zero-padded layout (gr_pad_table256, 1 KB rows): every
//     lane is in-bounds, so the loop has no clamping or predication, one 64-bit address per gathered edge and the
//     second column chunk is an immediate (+512 B) off it.
//   * ncu on the first version of this file (profiles/): l1tex__data_pipe_lsu_wavefronts at 76 % -- the LSU data pipe
//     (one 128-byte wavefront per clock per SM) was the limiter, with 40 of 115 wavefronts per (row, direction) spent
//     re-reading relu(+-ins) from shared memory, 36 on the gather (two padded 512-byte chunks per edge, odd rows
//     padded to an even edge count) and 16 on the stores.  Hence: relu(+-ins)/2 is staged in shared memory once per
//     tile but held in REGISTERS while a warp stays inside one question; the second column chunk is loaded only by the
//     lanes that own real columns (3 wavefronts instead of 4); edges are taken two at a time with an unpadded
//     single-edge tail.
//   * output: the split-bf16 planes of the e2e GEMM's A operand, segment pitch SEGP (32-byte sectors, see
//     aggregate.cu); columns D..SEGP-1 receive exact zeros (staged ins are zero there).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace gr {

int g_opt_agg_abs_ws = 1;     // gr_set_option("agg_abs_ws", 0|1|2..): persistent warp-specialised kernel when a tile counter is given;
                              // >= 2: TMA-gather variants (agg_abs_tma_kernel)
int g_opt_agg_table_rows = 0;  // experiment plumbing: rows of the padded tables for the gather4 tensor maps
int g_opt_agg_hot_rel = -1;   // gr_set_option("agg_hot_rel", id): relation row kept resident by the TMA-gather kernel

namespace {

constexpr int kRows = 64;         // destination rows per CTA tile
constexpr int kThreads = 256;     // consumer threads
constexpr int kWarps = kThreads / 32;
constexpr int kEdgeCap = 1024;    // staged edges per direction per tile; the rest takes the slow path
constexpr int kPnCols = 256;      // padded table width
constexpr int kPnRowBytes = kPnCols * 4;

struct PnDir {
  const int32_t* rowptr;
  const int32_t* src;
  const int32_t* rel;
  const float* w;
  const float* pn;      // [R1][256] zero-padded relation table
};

struct PnParams {
  PnDir dir[2];
  const float* prior;
  const float* ins;     // [B, I, D]
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  int64_t ld, out_col0, Nt;
  int B, N, I, j0;
  int32_t* tile_counter;   // persistent kernel: dynamic tile scheduler (zeroed before the launch)
  int hot_rel;             // TMA-gather kernel: relation whose table row stays resident in shared memory (-1: none)
  int64_t table_rows;      // rows of each padded relation table (R1): the gather4 tensor maps need the extent
};

__device__ __forceinline__ float4 ldg4(const char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ void fma4(float4& acc, float c, const float4& v) {          // acc += c * v
  const float2 cc = make_float2(c, c);
  const float2 lo = __ffma2_rn(cc, make_float2(v.x, v.y), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(cc, make_float2(v.z, v.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void fma4_abs(float4& acc, float c, const float4& v) {      // acc += c * |v|
  const float2 cc = make_float2(c, c);
  const float2 lo = __ffma2_rn(cc, make_float2(fabsf(v.x), fabsf(v.y)), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(cc, make_float2(fabsf(v.z), fabsf(v.w)), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ float4 addsub4(const float4& q, const float4& s, float sign) {   // q + sign*s, packed
  const float2 ss = make_float2(sign, sign);
  const float2 lo = __ffma2_rn(make_float2(s.x, s.y), ss, make_float2(q.x, q.y));
  const float2 hi = __ffma2_rn(make_float2(s.z, s.w), ss, make_float2(q.z, q.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// y = xp*(Q+S) + xn*(Q-S)  (xp, xn already carry the 1/2) -> (hi, lo) bf16 pairs, 8-byte stores into both planes
__device__ __forceinline__ void emit4(__nv_bfloat16* ph, __nv_bfloat16* pl, bool pred, const float4& xp,
                                      const float4& xn, const float4& U, const float4& V) {
  float2 y01 = __fmul2_rn(make_float2(xp.x, xp.y), make_float2(U.x, U.y));
  float2 y23 = __fmul2_rn(make_float2(xp.z, xp.w), make_float2(U.z, U.w));
  y01 = __ffma2_rn(make_float2(xn.x, xn.y), make_float2(V.x, V.y), y01);
  y23 = __ffma2_rn(make_float2(xn.z, xn.w), make_float2(V.z, V.w), y23);
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(y01.x, y01.y), h23 = __floats2bfloat162_rn(y23.x, y23.y);
  // bf16x2 -> float2 by hand: low half << 16, high half masked (2 ALU ops per pair; the library routine compiles to 4)
  const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
  const float2 f01 = make_float2(__uint_as_float(u01 << 16), __uint_as_float(u01 & 0xffff0000u));
  const float2 f23 = make_float2(__uint_as_float(u23 << 16), __uint_as_float(u23 & 0xffff0000u));
  const float2 m1 = make_float2(-1.f, -1.f);
  const float2 r01 = __ffma2_rn(f01, m1, y01), r23 = __ffma2_rn(f23, m1, y23);   // y - hi, exact, packed
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(r01.x, r01.y);
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(r23.x, r23.y);
  if (pred) {
    *reinterpret_cast<uint2*>(ph) = make_uint2(u01, u23);
    *reinterpret_cast<uint2*>(pl) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  }
}

// relu(+-ins)/2 of one question for this lane's 2 x 4 columns: shared memory -> registers
template <int NI>
struct LaneIns {
  float4 xp[NI][2], xn[NI][2];
  __device__ __forceinline__ void load(const float* xq) {   // xq = &x[q][0][0][lane * 4]
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const float* xj = xq + j * 2 * kPnCols;
      xp[j][0] = *reinterpret_cast<const float4*>(xj);
      xp[j][1] = *reinterpret_cast<const float4*>(xj + 128);
      xn[j][0] = *reinterpret_cast<const float4*>(xj + kPnCols);
      xn[j][1] = *reinterpret_cast<const float4*>(xj + kPnCols + 128);
    }
  }
};

// One (destination row, direction) unit: gather + accumulate the row's in-edges, then emit the NI instruction
// segments.  rc: staged {table byte offset, coefficient} of the tile's edge slice; [beg, end) the row's range in it.
template <int NI, int DT, int SEGP>
__device__ __forceinline__ void row_unit(const int2* __restrict__ rc, int beg, int end, int ebase, const PnDir& dd,
                                         const float* __restrict__ prior, const char* tb, const LaneIns<NI>& x,
                                         __nv_bfloat16* hrow, __nv_bfloat16* lrow, int seg_d, bool ld1, bool wr1) {
  float4 S0 = zero4(), S1 = S0, Q0 = S0, Q1 = S0;
  float4 v01 = zero4(), v11 = zero4();                     // lanes without chunk-1 columns never overwrite these
  const int fast_end = min(end, kEdgeCap);
  int i = beg;
  for (; i + 1 < fast_end; i += 2) {                       // two edges per step: 4 x 16-byte loads in flight per lane
    const int2 m0 = rc[i], m1 = rc[i + 1];
    const char* a0 = tb + (uint32_t)m0.x;
    const char* a1 = tb + (uint32_t)m1.x;
    const float4 v00 = ldg4(a0), v10 = ldg4(a1);
    if (ld1) { v01 = ldg4(a0 + 512); v11 = ldg4(a1 + 512); }
    const float c0 = __int_as_float(m0.y), c1 = __int_as_float(m1.y);
    fma4(S0, c0, v00); fma4_abs(Q0, c0, v00); fma4(S1, c0, v01); fma4_abs(Q1, c0, v01);
    fma4(S0, c1, v10); fma4_abs(Q0, c1, v10); fma4(S1, c1, v11); fma4_abs(Q1, c1, v11);
  }
  if (i < fast_end) {                                      // odd tail: no padded slot
    const int2 m0 = rc[i];
    const char* a0 = tb + (uint32_t)m0.x;
    const float4 v00 = ldg4(a0);
    if (ld1) v01 = ldg4(a0 + 512);
    const float c0 = __int_as_float(m0.y);
    fma4(S0, c0, v00); fma4_abs(Q0, c0, v00); fma4(S1, c0, v01); fma4_abs(Q1, c0, v01);
  }
  for (i = max(beg, kEdgeCap); i < end; ++i) {             // slow path: slice overflowed the staging buffer
    const int64_t e = (int64_t)ebase + i;
    const float w = dd.w ? dd.w[e] : 1.0f;
    const float c = w * (w * prior[dd.src[e]]);
    const char* a = tb + (uint32_t)dd.rel[e] * (uint32_t)kPnRowBytes;
    const float4 v0 = ldg4(a);
    if (ld1) v01 = ldg4(a + 512);
    fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v01); fma4_abs(Q1, c, v01);
  }
  const float4 U0 = addsub4(Q0, S0, 1.f), V0 = addsub4(Q0, S0, -1.f);   // 2 * sum c*relu(v), 2 * sum c*relu(-v)
  const float4 U1 = addsub4(Q1, S1, 1.f), V1 = addsub4(Q1, S1, -1.f);
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    const int seg = seg_d + j * 2 * SEGP;
    emit4(hrow + seg, lrow + seg, true, x.xp[j][0], x.xn[j][0], U0, V0);
    emit4(hrow + seg + 128, lrow + seg + 128, wr1, x.xp[j][1], x.xn[j][1], U1, V1);
  }
}

// relu(+-ins)/2 staging of the tile's two questions (float4 granularity), by `nthr` threads starting at `t`
template <int NI, int DT>
__device__ __forceinline__ void stage_ins(float (*x)[NI][2][kPnCols], const PnParams& p, int b0, int t, int nthr) {
  for (int i = t; i < 2 * NI * (kPnCols / 4); i += nthr) {
    const int c4 = i % (kPnCols / 4), j = (i / (kPnCols / 4)) % NI, q = i / ((kPnCols / 4) * NI);
    const int b = b0 + q;
    float4 v = zero4();
    if (4 * c4 < DT && b < p.B)
      v = __ldg(reinterpret_cast<const float4*>(p.ins + ((int64_t)b * p.I + p.j0 + j) * DT) + c4);
    reinterpret_cast<float4*>(&x[q][j][0][0])[c4] =
        make_float4(0.5f * fmaxf(v.x, 0.f), 0.5f * fmaxf(v.y, 0.f), 0.5f * fmaxf(v.z, 0.f), 0.5f * fmaxf(v.w, 0.f));
    reinterpret_cast<float4*>(&x[q][j][1][0])[c4] = make_float4(0.5f * fmaxf(-v.x, 0.f), 0.5f * fmaxf(-v.y, 0.f),
                                                                 0.5f * fmaxf(-v.z, 0.f), 0.5f * fmaxf(-v.w, 0.f));
  }
}

// ---------------------------------------------------------------------------------------------------------
// One CTA per 64-row tile (used when the caller passes no tile counter)
// ---------------------------------------------------------------------------------------------------------
template <int NI, int DT, int SEGP>
__global__ void __launch_bounds__(kThreads, 2) agg_abs_kernel(const PnParams p) {
  static_assert(DT % 4 == 0 && DT > 128 && DT <= kPnCols, "two column chunks of 128");
  __shared__ int32_t s_rowptr[2][kRows + 1];
  __shared__ int2 s_rc[2][kEdgeCap];                        // {table byte offset rel*1024, float_as_int(c)}
  __shared__ unsigned char s_any[2][kRows];
  __shared__ __align__(16) float s_x[2][NI][2][kPnCols];    // [question of the tile][j][relu(+x)/2 | relu(-x)/2][col]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * kRows;
  const int nrows = (int)min((int64_t)kRows, p.Nt - r0);
  const int N = p.N;
  const int b0 = (int)(r0 / N);
  const int rem0 = (int)(r0 - (int64_t)b0 * N);

  // ---------------- phase 1: row pointers, edge slice -> (table offset, coefficient), instructions ----------
  if (tid <= nrows) {
    s_rowptr[0][tid] = p.dir[0].rowptr[r0 + tid];
    s_rowptr[1][tid] = p.dir[1].rowptr[r0 + tid];
  }
  stage_ins<NI, DT>(s_x, p, b0, tid, kThreads);
  __syncthreads();
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const PnDir& dd = p.dir[d];
    const int64_t eb = s_rowptr[d][0];
    const int ne = (int)min((int64_t)(s_rowptr[d][nrows] - eb), (int64_t)kEdgeCap);
    for (int i = tid; i < ne; i += kThreads) {
      const int s = dd.src[eb + i];
      const int r = dd.rel[eb + i];
      const float w = dd.w ? dd.w[eb + i] : 1.0f;
      const float c = w * (w * p.prior[s]);                  // reasongnn.py:80-84
      s_rc[d][i] = make_int2((int)((uint32_t)r * (uint32_t)kPnRowBytes), __float_as_int(c));
    }
  }
  __syncthreads();
  if (tid < 2 * kRows) {   // rows whose in-edges all carry c == 0 are pure zero stores
    const int d = tid / kRows, lr = tid % kRows;
    unsigned char any = 0;
    if (lr < nrows) {
      const int ebase = s_rowptr[d][0];
      const int beg = s_rowptr[d][lr] - ebase, end = s_rowptr[d][lr + 1] - ebase;
      if (end > kEdgeCap) any = 1;
      for (int i = beg; i < min(end, kEdgeCap) && !any; ++i) any = (s_rc[d][i].y << 1) != 0;
    }
    s_any[d][lr] = any;
  }
  __syncthreads();

  // ---------------- phase 2: one warp per destination row, lane = 4 columns in each of 2 chunks -----------
  const bool ld1 = 128 + lane * 4 < DT;                      // chunk 1: lanes that own real columns
  const bool wr1 = 128 + lane * 4 < SEGP;                    //          lanes that own segment columns (incl. zero pad)
  const char* tb[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) tb[d] = reinterpret_cast<const char*>(p.dir[d].pn) + lane * 16;
  __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
  __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
  const int lr_switch = N - rem0;                            // first tile row of question b0 + 1 (N >= kRows)
  LaneIns<NI> x;
  int cur_q = -1;
  for (int lr = warp; lr < nrows; lr += kWarps) {
    const int q = lr >= lr_switch ? 1 : 0;
    if (q != cur_q) {
      cur_q = q;
      x.load(&s_x[q][0][0][lane * 4]);
    }
    __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
    __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const int ebase = s_rowptr[d][0];
      const int beg = s_rowptr[d][lr] - ebase, end = s_rowptr[d][lr + 1] - ebase;
      if (!s_any[d][lr]) {
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int seg = d * SEGP + j * 2 * SEGP;
          *reinterpret_cast<uint2*>(hrow + seg) = make_uint2(0u, 0u);
          *reinterpret_cast<uint2*>(lrow + seg) = make_uint2(0u, 0u);
          if (wr1) {
            *reinterpret_cast<uint2*>(hrow + seg + 128) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(lrow + seg + 128) = make_uint2(0u, 0u);
          }
        }
        continue;
      }
      row_unit<NI, DT, SEGP>(s_rc[d], beg, end, ebase, p.dir[d], p.prior, tb[d], x, hrow, lrow, d * SEGP, ld1, wr1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent, warp-specialised version: each CTA lives for the whole launch; warp 8 (producer) stages tile t+1 (row
// pointers, {table offset, coefficient} per edge, relu(+-ins)/2) into the other half of a double buffer while warps
// 0-7 (consumers, one row at a time) work on tile t; full/empty mbarriers per buffer; tiles are handed out by an
// atomic counter so the tail balances.  (Measured: the staging round trips it hides were NOT the limiter -- 140 vs
// 142 us -- the LSU data pipe was; kept because it is never slower and frees the consumers from all index work.)
// ---------------------------------------------------------------------------------------------------------
constexpr int kWsThreads = kThreads + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(uint64_t* b) { mbar_arrive(b); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity), "r"(20000u)     // suspend-time hint [ns]: a waiting role does not spin on the
        : "memory");                                       // issue port (measured: 129 us with, 141 us without)
  }
}

template <int NI>
struct alignas(16) WsBuf {
  int2 rc[2][kEdgeCap];
  float x[2][NI][2][kPnCols];
  int32_t rowptr[2][kRows + 4];
  int32_t tile;
};
static_assert(sizeof(WsBuf<2>) % 16 == 0, "double buffer halves must stay 16-byte aligned");


// Producer side of the persistent kernels: stage one 64-row tile (row pointers, relu(+-ins)/2 of its <= 2 questions,
// {table byte offset, coefficient} per edge of both directions) into `bf`.  One warp.
template <int NI, int DT, int CAP, bool FLAG_HOT, class Buf>
__device__ __forceinline__ void produce_tile(Buf& bf, const PnParams& p, int tile, int lane) {
  const int N = p.N;
  const int64_t r0 = (int64_t)tile * kRows;
  const int nrows = (int)min((int64_t)kRows, p.Nt - r0);
  const int b0 = (int)(r0 / N);
  if (lane == 0) bf.tile = tile;
  int eb[2], ne[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int32_t* rp = p.dir[d].rowptr + r0;
    const int e0 = __ldg(rp), e1 = __ldg(rp + nrows);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = lane + 32 * k;
      if (i <= nrows) bf.rowptr[d][i] = __ldg(rp + i);
    }
    eb[d] = e0;
    ne[d] = min(e1 - e0, CAP);
  }
  stage_ins<NI, DT>(bf.x, p, b0, lane, 32);
  // edge slice -> {table byte offset, coefficient}; 4 edges per lane in flight
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const PnDir& dd = p.dir[d];
    for (int i0 = 0; i0 < ne[d]; i0 += 128) {
      int sidx[4], ridx[4];
      float wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + lane + 32 * u;
        const bool ok = i < ne[d];
        sidx[u] = ok ? __ldg(dd.src + eb[d] + i) : 0;
        ridx[u] = ok ? __ldg(dd.rel + eb[d] + i) : 0;
        wv[u] = (ok && dd.w) ? __ldg(dd.w + eb[d] + i) : 1.0f;
      }
      float pr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) pr[u] = __ldg(p.prior + sidx[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + lane + 32 * u;
        if (i < ne[d])
          bf.rc[d][i] = make_int2((int)((uint32_t)ridx[u] * (uint32_t)kPnRowBytes) |
                                      ((FLAG_HOT && ridx[u] == p.hot_rel) ? 1 : 0),
                                  __float_as_int(wv[u] * (wv[u] * pr[u])));
      }
    }
  }
}

template <int NI, int DT, int SEGP>
__global__ void __launch_bounds__(kWsThreads, 2) agg_abs_ws_kernel(const PnParams p, int ntiles) {
  extern __shared__ __align__(16) unsigned char ws_smem[];
  WsBuf<NI>* bufs = reinterpret_cast<WsBuf<NI>*>(ws_smem);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], kThreads); mbar_init(&s_empty[1], kThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kWarps) {
    // =============================== producer warp ===============================
    for (int it = 0;; ++it) {
      WsBuf<NI>& bf = bufs[it & 1];
      if (it >= 2) mbar_wait(&s_empty[it & 1], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[it & 1]);
        break;
      }
      produce_tile<NI, DT, kEdgeCap, false>(bf, p, tile, lane);
      __syncwarp();
      mbar_arrive(&s_full[it & 1]);
    }
    return;
  }

  // =============================== consumer warps ===============================
  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const char* tb[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) tb[d] = reinterpret_cast<const char*>(p.dir[d].pn) + lane * 16;
  LaneIns<NI> x;
  for (int it = 0;; ++it) {
    WsBuf<NI>& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * kRows;
    const int nrows = (int)min((int64_t)kRows, p.Nt - r0);
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    int cur_q = -1;
    for (int lr = warp; lr < nrows; lr += kWarps) {
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
        x.load(&bf.x[q][0][0][lane * 4]);
      }
      __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
      __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int ebase = bf.rowptr[d][0];
        const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
        row_unit<NI, DT, SEGP>(bf.rc[d], beg, end, ebase, p.dir[d], p.prior, tb[d], x, hrow, lrow, d * SEGP, ld1, wr1);
      }
    }
    mbar_arrive(&s_empty[it & 1]);     // every consumer thread arrives (count = kThreads)
  }
}



// ---------------------------------------------------------------------------------------------------------
// Software-pipelined consumers (gr_set_option("agg_abs_ws", 6)): same producer / buffers as agg_abs_ws_kernel, but a
// consumer warp issues the gather loads of the first two edges of its NEXT (row, direction) unit before it runs the
// epilogue of the current one (bf16 hi/lo split + stores, ~75 instructions without a memory dependence), so that
// L2 latency is covered by the epilogue instead of by other warps (there are only 4 per scheduler).  Missing edges
// of a short unit are replaced by (table row 0, coefficient 0): fma(0, finite, acc) == acc, branch-free.
// ---------------------------------------------------------------------------------------------------------
struct Pre2 {
  float4 v00, v01, v10, v11;
  float c0, c1;
};

template <int NI, int DT, int SEGP>
__global__ void __launch_bounds__(kWsThreads, 2) agg_abs_ws2_kernel(const PnParams p, int ntiles) {
  extern __shared__ __align__(16) unsigned char ws_smem[];
  WsBuf<NI>* bufs = reinterpret_cast<WsBuf<NI>*>(ws_smem);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], kThreads); mbar_init(&s_empty[1], kThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kWarps) {
    for (int it = 0;; ++it) {
      WsBuf<NI>& bf = bufs[it & 1];
      if (it >= 2) mbar_wait(&s_empty[it & 1], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[it & 1]);
        break;
      }
      produce_tile<NI, DT, kEdgeCap, false>(bf, p, tile, lane);
      __syncwarp();
      mbar_arrive(&s_full[it & 1]);
    }
    return;
  }

  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const char* const tb0 = reinterpret_cast<const char*>(p.dir[0].pn) + lane * 16;
  const char* const tb1 = reinterpret_cast<const char*>(p.dir[1].pn) + lane * 16;
  LaneIns<NI> x;
  Pre2 pre;
  pre.v01 = zero4(); pre.v11 = zero4();
  for (int it = 0;; ++it) {
    WsBuf<NI>& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * kRows;
    const int nrows = (int)min((int64_t)kRows, p.Nt - r0);
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    const int nunits = warp < nrows ? 2 * ((nrows - warp + kWarps - 1) / kWarps) : 0;
    const int eb0 = bf.rowptr[0][0], eb1 = bf.rowptr[1][0];

    // issue the gather of the first two staged edges of unit u into `pre`
    auto prefetch = [&](int u) {
      const int d = u & 1, lr = warp + kWarps * (u >> 1);
      const int ebase = d ? eb1 : eb0;
      const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
      const int nf = min(end, kEdgeCap) - beg;                 // staged edges of the unit (may be <= 0)
      const int2 m0 = nf >= 1 ? bf.rc[d][beg] : make_int2(0, 0);
      const int2 m1 = nf >= 2 ? bf.rc[d][beg + 1] : make_int2(0, 0);
      const char* tb = d ? tb1 : tb0;
      const char* a0 = tb + (uint32_t)m0.x;
      const char* a1 = tb + (uint32_t)m1.x;
      pre.v00 = ldg4(a0); pre.v10 = ldg4(a1);
      if (ld1) { pre.v01 = ldg4(a0 + 512); pre.v11 = ldg4(a1 + 512); }
      pre.c0 = __int_as_float(m0.y); pre.c1 = __int_as_float(m1.y);
    };

    int cur_q = -1;
    if (nunits > 0) prefetch(0);
    for (int u = 0; u < nunits; ++u) {
      const int d = u & 1, lr = warp + kWarps * (u >> 1);
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
        x.load(&bf.x[q][0][0][lane * 4]);
      }
      const int ebase = d ? eb1 : eb0;
      const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
      const int fast_end = min(end, kEdgeCap);
      const int2* __restrict__ rc = bf.rc[d];
      const char* tb = d ? tb1 : tb0;
      float4 S0 = zero4(), S1 = S0, Q0 = S0, Q1 = S0;
      // the two prefetched edges (coefficient 0 where the unit is shorter)
      fma4(S0, pre.c0, pre.v00); fma4_abs(Q0, pre.c0, pre.v00); fma4(S1, pre.c0, pre.v01); fma4_abs(Q1, pre.c0, pre.v01);
      fma4(S0, pre.c1, pre.v10); fma4_abs(Q0, pre.c1, pre.v10); fma4(S1, pre.c1, pre.v11); fma4_abs(Q1, pre.c1, pre.v11);
      {
        float4 v01 = zero4(), v11 = zero4();
        int i = beg + 2;
        for (; i + 1 < fast_end; i += 2) {
          const int2 m0 = rc[i], m1 = rc[i + 1];
          const char* a0 = tb + (uint32_t)m0.x;
          const char* a1 = tb + (uint32_t)m1.x;
          const float4 v00 = ldg4(a0), v10 = ldg4(a1);
          if (ld1) { v01 = ldg4(a0 + 512); v11 = ldg4(a1 + 512); }
          const float c0 = __int_as_float(m0.y), c1 = __int_as_float(m1.y);
          fma4(S0, c0, v00); fma4_abs(Q0, c0, v00); fma4(S1, c0, v01); fma4_abs(Q1, c0, v01);
          fma4(S0, c1, v10); fma4_abs(Q0, c1, v10); fma4(S1, c1, v11); fma4_abs(Q1, c1, v11);
        }
        if (i < fast_end) {
          const int2 m0 = rc[i];
          const char* a0 = tb + (uint32_t)m0.x;
          const float4 v00 = ldg4(a0);
          if (ld1) v01 = ldg4(a0 + 512);
          const float c0 = __int_as_float(m0.y);
          fma4(S0, c0, v00); fma4_abs(Q0, c0, v00); fma4(S1, c0, v01); fma4_abs(Q1, c0, v01);
        }
        const PnDir& dd = p.dir[d];
        for (i = max(beg, kEdgeCap); i < end; ++i) {           // slow path: slice overflowed the staging buffer
          const int64_t e = (int64_t)ebase + i;
          const float w = dd.w ? dd.w[e] : 1.0f;
          const float c = w * (w * p.prior[dd.src[e]]);
          const char* a = tb + (uint32_t)dd.rel[e] * (uint32_t)kPnRowBytes;
          const float4 v0 = ldg4(a);
          if (ld1) v01 = ldg4(a + 512);
          fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v01); fma4_abs(Q1, c, v01);
        }
      }
      if (u + 1 < nunits) prefetch(u + 1);                     // loads fly while the epilogue below runs
      __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
      __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
      const float4 U0 = addsub4(Q0, S0, 1.f), V0 = addsub4(Q0, S0, -1.f);
      const float4 U1 = addsub4(Q1, S1, 1.f), V1 = addsub4(Q1, S1, -1.f);
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int seg = d * SEGP + j * 2 * SEGP;
        emit4(hrow + seg, lrow + seg, true, x.xp[j][0], x.xn[j][0], U0, V0);
        emit4(hrow + seg + 128, lrow + seg + 128, wr1, x.xp[j][1], x.xn[j][1], U1, V1);
      }
    }
    mbar_arrive(&s_empty[it & 1]);
  }
}


// ---------------------------------------------------------------------------------------------------------
// Half-row consumers (gr_set_option("agg_abs_ws", 7)).
//
// ncu on agg_abs_ws_kernel (profiles/r2_*): 246 warp instructions per (row, direction) unit at 48 % issue-slot
// utilisation; a consumer warp spends ~2300 cycles per unit, most of it in two dependent L2 round trips (two edges
// per trip) -- the kernel is bound by the gather LATENCY per warp, and at 96 registers (32 of them the per-question
// relu(+-ins) values of 8 columns per lane) only 16 consumer warps fit on an SM.  Here a destination row is shared
// by a PAIR of warps, each owning one column half (columns [0,96) / [96,208): the boundary is a whole 32-byte
// sector in both the table row and the bf16 planes) with 4 columns per lane: half the live state per thread ->
// 64 registers -> 28 consumer warps per SM, and every warp keeps 4 edges in flight (one round trip for a row of <= 4
// in-edges).  Bytes of gather in flight per SM: 28 x 4 x ~400 B = 45 KB instead of 16 x 2 x 800 B = 26 KB.
// Same producer, same staged {offset, coefficient} slices, same summation order (slot order within the row) ->
// bit-identical results.  Tiles are 56 rows (7 warp pairs x 8 rows).
// ---------------------------------------------------------------------------------------------------------
constexpr int kHalfSplit = 96;     // first column of the second half (multiple of 16 bf16 / 8 fp32: sector aligned)

template <int NI, int ROWS>
struct alignas(16) HBuf {
  int2 rc[2][kEdgeCap];
  float x[2][NI][2][kPnCols];
  int32_t rowptr[2][ROWS + 4];
  int32_t tile;
  int32_t pad_[3];
};

template <int NI, int DT, int ROWS, int CAP = kEdgeCap, bool REL_INDEX = false, class Buf>
__device__ __forceinline__ void produce_tile_rows(Buf& bf, const PnParams& p, int tile, int lane) {
  // The producer is ONE warp running dependent global loads (row pointers -> src / rel slices -> prior[src]); what
  // bounds it is the number of round trips per tile, not the instruction count.  Both directions and up to 8 edges
  // per lane and direction are therefore fetched per batch (a 64-row cfg2 tile has ~256 edges per direction: one
  // batch): 3 round trips per tile.  (The first version took 128 edges of one direction per batch: ~10 trips, and
  // the consumers of a tile were waiting for it: profiles/README.md.)
  constexpr int U = 8;
  const int N = p.N;
  const int64_t r0 = (int64_t)tile * ROWS;
  const int nrows = (int)min((int64_t)ROWS, p.Nt - r0);
  const int b0 = (int)(r0 / N);
  if (lane == 0) bf.tile = tile;
  int eb[2], ne[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int32_t* rp = p.dir[d].rowptr + r0;
    eb[d] = __ldg(rp);
    ne[d] = __ldg(rp + nrows);
  }
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int32_t* rp = p.dir[d].rowptr + r0;
#pragma unroll
    for (int k = 0; k < (ROWS + 32) / 32; ++k) {
      const int i = lane + 32 * k;
      if (i <= nrows) bf.rowptr[d][i] = __ldg(rp + i);
    }
    ne[d] = min(ne[d] - eb[d], CAP);
  }
  const int nmax = max(ne[0], ne[1]);
  const bool has_w = p.dir[0].w != nullptr || p.dir[1].w != nullptr;
  for (int i0 = 0; i0 < nmax; i0 += 32 * U) {
    int sidx[2][U], ridx[2][U];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const PnDir& dd = p.dir[d];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + lane + 32 * u;
        const bool ok = i < ne[d];
        sidx[d][u] = ok ? __ldg(dd.src + eb[d] + i) : 0;
        ridx[d][u] = ok ? __ldg(dd.rel + eb[d] + i) : 0;
      }
    }
    if (i0 == 0) stage_ins<NI, DT>(bf.x, p, b0, lane, 32);   // overlaps the first batch's index loads
    float pr[2][U];
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
      for (int u = 0; u < U; ++u) pr[d][u] = __ldg(p.prior + sidx[d][u]);
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const PnDir& dd = p.dir[d];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + lane + 32 * u;
        if (i < ne[d]) {
          const float w = (has_w && dd.w) ? __ldg(dd.w + eb[d] + i) : 1.0f;
          bf.rc[d][i] = make_int2(REL_INDEX ? ridx[d][u] : (int)((uint32_t)ridx[d][u] * (uint32_t)kPnRowBytes),
                                  __float_as_int(w * (w * pr[d][u])));
        }
      }
    }
  }
  if (nmax <= 0) stage_ins<NI, DT>(bf.x, p, b0, lane, 32);
}

// y = xp*U + xn*V -> bf16 hi/lo, one 8-byte store per plane
__device__ __forceinline__ void emit4h(__nv_bfloat16* ph, __nv_bfloat16* pl, bool pred, const float4& xp,
                                       const float4& xn, const float4& U, const float4& V) {
  emit4(ph, pl, pred, xp, xn, U, V);
}

template <int NI, int DT, int SEGP, int SLOTS, int RPS>
__global__ void __launch_bounds__((2 * SLOTS + 1) * 32, 2) agg_abs_half_kernel(const PnParams p, int ntiles) {
  constexpr int ROWS = SLOTS * RPS;
  constexpr int CW = 2 * SLOTS;                    // consumer warps
  static_assert(DT % 4 == 0 && DT > kHalfSplit && DT <= kHalfSplit + 128 && SEGP <= kHalfSplit + 128, "two halves");
  using Buf = HBuf<NI, ROWS>;
  extern __shared__ __align__(16) unsigned char ws_smem[];
  Buf* bufs = reinterpret_cast<Buf*>(ws_smem);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], CW * 32); mbar_init(&s_empty[1], CW * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == CW) {
    // =============================== producer warp ===============================
    for (int it = 0;; ++it) {
      Buf& bf = bufs[it & 1];
      if (it >= 2) mbar_wait(&s_empty[it & 1], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[it & 1]);
        break;
      }
      produce_tile_rows<NI, DT, ROWS>(bf, p, tile, lane);
      __syncwarp();
      mbar_arrive(&s_full[it & 1]);
    }
    return;
  }

  // =============================== consumer warps: (slot, column half) ===============================
  const int slot = warp >> 1, half = warp & 1;
  const int col0 = half * kHalfSplit + lane * 4;                      // first of this lane's 4 columns
  const bool ld = col0 < DT && (half == 1 || lane * 4 < kHalfSplit);   // lane owns real table columns
  const bool wr = col0 < SEGP && (half == 1 || lane * 4 < kHalfSplit); // lane owns segment columns (incl. zero pad)
  const char* const tb0 = reinterpret_cast<const char*>(p.dir[0].pn) + col0 * 4;
  const char* const tb1 = reinterpret_cast<const char*>(p.dir[1].pn) + col0 * 4;
  float4 xp[NI], xn[NI];
  for (int it = 0;; ++it) {
    Buf& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * ROWS;
    const int nrows = (int)min((int64_t)ROWS, p.Nt - r0);
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + col0 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + col0 + (int64_t)p.j0 * 2 * SEGP;
    int cur_q = -1;
    for (int lr = slot; lr < nrows; lr += SLOTS) {
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          xp[j] = wr ? *reinterpret_cast<const float4*>(&bf.x[q][j][0][col0 & (kPnCols - 1)]) : zero4();
          xn[j] = wr ? *reinterpret_cast<const float4*>(&bf.x[q][j][1][col0 & (kPnCols - 1)]) : zero4();
        }
      }
      __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
      __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int ebase = bf.rowptr[d][0];
        const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
        const int fast_end = min(end, kEdgeCap);
        const int2* __restrict__ rc = bf.rc[d];
        const char* tb = d ? tb1 : tb0;
        float4 S = zero4(), Q = zero4();
        float4 v0 = zero4(), v1 = zero4(), v2 = zero4(), v3 = zero4();   // lanes without table columns keep zeros
        for (int i = beg; i < fast_end; i += 4) {
          const int n = fast_end - i;
          const int2 m0 = rc[i];
          const int2 m1 = rc[n > 1 ? i + 1 : i];
          if (n > 2) {                                       // 3 or 4 edges: four loads in flight
            const int2 m2 = rc[i + 2];
            const int2 m3 = rc[n > 3 ? i + 3 : i + 2];
            if (ld) {
              v0 = ldg4(tb + (uint32_t)m0.x); v1 = ldg4(tb + (uint32_t)m1.x);
              v2 = ldg4(tb + (uint32_t)m2.x); v3 = ldg4(tb + (uint32_t)m3.x);
            }
            const float c0 = __int_as_float(m0.y), c1 = __int_as_float(m1.y), c2 = __int_as_float(m2.y);
            const float c3 = n > 3 ? __int_as_float(m3.y) : 0.f;
            fma4(S, c0, v0); fma4_abs(Q, c0, v0);
            fma4(S, c1, v1); fma4_abs(Q, c1, v1);
            fma4(S, c2, v2); fma4_abs(Q, c2, v2);
            fma4(S, c3, v3); fma4_abs(Q, c3, v3);
          } else {
            if (ld) { v0 = ldg4(tb + (uint32_t)m0.x); v1 = ldg4(tb + (uint32_t)m1.x); }
            const float c0 = __int_as_float(m0.y);
            const float c1 = n > 1 ? __int_as_float(m1.y) : 0.f;
            fma4(S, c0, v0); fma4_abs(Q, c0, v0);
            fma4(S, c1, v1); fma4_abs(Q, c1, v1);
          }
        }
        {
          const PnDir& dd = p.dir[d];
          for (int i = max(beg, kEdgeCap); i < end; ++i) {   // slow path: slice overflowed the staging buffer
            const int64_t e = (int64_t)ebase + i;
            const float w = dd.w ? dd.w[e] : 1.0f;
            const float c = w * (w * p.prior[dd.src[e]]);
            if (ld) v0 = ldg4(tb + (uint32_t)dd.rel[e] * (uint32_t)kPnRowBytes);
            fma4(S, c, v0); fma4_abs(Q, c, v0);
          }
        }
        const float4 U = addsub4(Q, S, 1.f), V = addsub4(Q, S, -1.f);
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int seg = d * SEGP + j * 2 * SEGP;
          emit4h(hrow + seg, lrow + seg, wr, xp[j], xn[j], U, V);
        }
      }
    }
    mbar_arrive(&s_empty[it & 1]);     // every consumer thread arrives
  }
}


// ---------------------------------------------------------------------------------------------------------
// Shape-generic persistent kernel (gr_set_option("agg_abs_ws", 10..)): the agg_abs_ws_kernel structure with the
// number of consumer warps, the rows per tile and the resident CTAs per SM as template parameters.  The register
// file is split per scheduler (16 K registers each): with W warps per CTA and MINB CTAs per SM the busiest scheduler
// holds ceil(W * MINB / 4) warps, so 9-warp CTAs x 2 leave 96 registers per thread, 10-warp CTAs x 2 still 96 (5 warps
// per scheduler on all four), 12-warp CTAs x 2 leave 80.
// ---------------------------------------------------------------------------------------------------------
template <int NI, int DT, int SEGP, int KW, int ROWS, int MINB>
__global__ void __launch_bounds__((KW + 1) * 32, MINB) agg_abs_wsg_kernel(const PnParams p, int ntiles) {
  using Buf = HBuf<NI, ROWS>;
  extern __shared__ __align__(16) unsigned char ws_smem[];
  Buf* bufs = reinterpret_cast<Buf*>(ws_smem);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], KW * 32); mbar_init(&s_empty[1], KW * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == KW) {
    for (int it = 0;; ++it) {
      Buf& bf = bufs[it & 1];
      if (it >= 2) mbar_wait(&s_empty[it & 1], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[it & 1]);
        break;
      }
      produce_tile_rows<NI, DT, ROWS>(bf, p, tile, lane);
      __syncwarp();
      mbar_arrive(&s_full[it & 1]);
    }
    return;
  }
  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const char* tb[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) tb[d] = reinterpret_cast<const char*>(p.dir[d].pn) + lane * 16;
  LaneIns<NI> x;
  for (int it = 0;; ++it) {
    Buf& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * ROWS;
    const int nrows = (int)min((int64_t)ROWS, p.Nt - r0);
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    int cur_q = -1;
    for (int lr = warp; lr < nrows; lr += KW) {
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
        x.load(&bf.x[q][0][0][lane * 4]);
      }
      __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
      __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int ebase = bf.rowptr[d][0];
        const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
        row_unit<NI, DT, SEGP>(bf.rc[d], beg, end, ebase, p.dir[d], p.prior, tb[d], x, hrow, lrow, d * SEGP, ld1, wr1);
      }
    }
    mbar_arrive(&s_empty[it & 1]);
  }
}

template <int NI, int KW, int ROWS, int MINB>
int launch_wsg(const PnParams& p, cudaStream_t stream) {
  auto kern = agg_abs_wsg_kernel<NI, 200, 208, KW, ROWS, MINB>;
  const size_t smem = 2 * sizeof(HBuf<NI, ROWS>);
  static bool attr_set = false;
  if (!attr_set) {
    GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
  const unsigned tiles = (unsigned)ceil_div(p.Nt, ROWS);
  const unsigned pgrid = std::min<unsigned>(tiles, (unsigned)MINB * (unsigned)sm_count());
  kern<<<pgrid, (KW + 1) * 32, smem, stream>>>(p, (int)tiles);
  GR_CHECK_LAUNCH();
  return GR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// TMA-gather version of the persistent kernel (gr_set_option("agg_abs_ws", 2..4)).
//
// ncu on agg_abs_ws_kernel (profiles/r1b_agg_kernel.txt): DRAM 38 % and L2 41 % of peak, issue slots 48 % busy,
// 58 % of the stall samples on long_scoreboard -- the per-lane LDGs of the gathered table rows: with 16 consumer
// warps per SM and two edges in flight per warp there are ~25 KB of gather outstanding per SM, not enough to cover
// the L2 latency.  Here the gather does not occupy warps at all: every consumer warp owns a private ring of NS
// 800-byte slots in shared memory and issues one 1-D bulk copy (cp.async.bulk, SASS UBLKCP) per gathered edge --
// lane i copies the table row of edge i of a (row, direction) unit -- for units up to kBars ahead of the one it is
// accumulating; the unit's mbarrier (expect_tx = n * 800 bytes) flips when all its rows have landed, and the
// accumulation loop reads the rows with conflict-free LDS.128.  Bytes in flight per SM = ring bytes (~150 KB)
// instead of 25 KB.  The prefetch cursor runs across tile boundaries (it peeks at the producer's next tile with a
// non-blocking mbarrier.test_wait), so there is no gather bubble at the start of a tile.
//
// Optional resident row (p.hot_rel >= 0): every real node carries a self-loop fact with the same relation id
// (gnn/dataset_load.py:499-506), i.e. ~25 % of all gathered rows are ONE table row per direction.  The producer
// flags those edges (bit 0 of the staged table offset); they are served from a copy of that row held in shared
// memory for the whole launch and never enter the ring.  Pure optimisation: any hot_rel gives the same result.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTmaCap = 512;       // staged edges per direction per tile (mean 256 at cfg2); the rest -> slow path
constexpr int kRingMaxUnit = 8;    // ring slots one (row, direction) unit may take; further staged edges use LDG
constexpr int kBars = 4;           // units in flight per consumer warp

template <int NI>
struct alignas(16) TmaBuf {
  int2 rc[2][kTmaCap];
  float x[2][NI][2][kPnCols];
  int32_t rowptr[2][kRows + 4];
  int32_t tile;
  int32_t pad_[3];
};
static_assert(sizeof(TmaBuf<2>) % 16 == 0, "double buffer halves must stay 16-byte aligned");

__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NI, int DT, int SEGP, int KW, int NS, int MINB, bool HOT>
__global__ void __launch_bounds__((KW + 1) * 32, MINB) agg_abs_tma_kernel(const PnParams p, int ntiles) {
  static_assert(DT % 4 == 0 && DT > 128 && DT <= kPnCols && NS >= kRingMaxUnit, "two column chunks of 128");
  using Buf = TmaBuf<NI>;
  constexpr int kSlot = DT * 4;                    // bytes per gathered table row
  extern __shared__ __align__(16) unsigned char ws_smem[];
  Buf* bufs = reinterpret_cast<Buf*>(ws_smem);
  unsigned char* ring_all = ws_smem + 2 * sizeof(Buf);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  __shared__ __align__(8) uint64_t s_bar[KW][kBars];
  __shared__ __align__(16) float s_hot[2][kPnCols];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], KW * 32); mbar_init(&s_empty[1], KW * 32);
    for (int w = 0; w < KW; ++w)
      for (int b = 0; b < kBars; ++b) mbar_init(&s_bar[w][b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (HOT) {
    for (int i = tid; i < 2 * kPnCols; i += (KW + 1) * 32) {
      const int d = i / kPnCols, c = i % kPnCols;
      s_hot[d][c] = p.hot_rel >= 0 ? __ldg(p.dir[d].pn + (int64_t)p.hot_rel * kPnCols + c) : 0.f;
    }
  }
  __syncthreads();

  if (warp == KW) {
    // =============================== producer warp ===============================
    for (int it = 0;; ++it) {
      Buf& bf = bufs[it & 1];
      if (it >= 2) mbar_wait(&s_empty[it & 1], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[it & 1]);
        break;
      }
      produce_tile<NI, DT, kTmaCap, HOT>(bf, p, tile, lane);
      __syncwarp();
      mbar_arrive(&s_full[it & 1]);
    }
    return;
  }

  // =============================== consumer warps ===============================
  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const char* const pnb0 = reinterpret_cast<const char*>(p.dir[0].pn);
  const char* const pnb1 = reinterpret_cast<const char*>(p.dir[1].pn);
  unsigned char* const ring = ring_all + (size_t)warp * NS * kSlot;
  const uint32_t ring_s = smem_u32(ring);
  uint64_t* const bars = s_bar[warp];

  // number of (row, direction) units this warp owns in the tile staged in `bf` (-1: end marker)
  auto tile_units = [&](const Buf& bf) -> int {
    const int t = bf.tile;
    if (t < 0) return -1;
    const int nr = (int)min((int64_t)kRows, p.Nt - (int64_t)t * kRows);
    return warp < nr ? 2 * ((nr - warp + KW - 1) / KW) : 0;
  };
  // ring edges of unit u of the tile in `bf`: the first <= kRingMaxUnit non-resident edges among its first 32 staged
  // edges.  Returns n; `mine` = this lane's edge is one of them, `rank` its slot rank, `off` its table offset.
  auto unit_ring = [&](const Buf& bf, int u, bool& mine, int& rank, int& off, int& dd) -> int {
    const int d = u & 1, lr = warp + KW * (u >> 1);
    const int ebase = bf.rowptr[d][0];
    const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
    const int cand = max(0, min(min(end, kTmaCap) - beg, 32));
    dd = d;
    if (!HOT) {
      const int n = min(cand, kRingMaxUnit);
      mine = lane < n;
      rank = lane;
      off = mine ? bf.rc[d][beg + lane].x : 0;
      return n;
    }
    off = lane < cand ? bf.rc[d][beg + lane].x : 1;
    const unsigned m = __ballot_sync(0xffffffffu, (off & 1) == 0);
    rank = __popc(m & ((1u << lane) - 1u));
    mine = (off & 1) == 0 && rank < kRingMaxUnit;
    return min(__popc(m), kRingMaxUnit);
  };

  int p_it = 0, p_u = 0, p_nu = -2;     // prefetch cursor: tile iteration, unit, units in that tile (-2: not entered)
  uint32_t seq_p = 0, seq_c = 0;        // units (with n > 0) issued / consumed
  int head_p = 0, head_c = 0, used = 0; // ring slot cursors, slots in flight
  LaneIns<NI> x;

  for (int it = 0;; ++it) {
    Buf& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int nunits = tile_units(bf);
    if (p_it < it || p_nu == -2) { p_it = it; p_u = 0; p_nu = nunits; }
    const int64_t r0 = (int64_t)tile * kRows;
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    int cur_q = -1;
    for (int u = 0; u < nunits; ++u) {
      // ---------------- prefetch: issue the bulk copies of units ahead of u ----------------
      for (;;) {
        if (p_u >= p_nu) {                               // cursor at the end of its tile: try to enter the next one
          if (p_nu < 0 || p_it > it) break;              // end marker seen / already one tile ahead
          if (!mbar_test(&s_full[(p_it + 1) & 1], ((p_it + 1) >> 1) & 1)) break;
          ++p_it; p_u = 0;
          p_nu = tile_units(bufs[p_it & 1]);
          continue;
        }
        if (seq_p - seq_c >= (uint32_t)kBars) break;
        bool mine; int rank, off, d;
        const int n = unit_ring(bufs[p_it & 1], p_u, mine, rank, off, d);
        if (used + n > NS) break;
        if (n > 0) {
          uint64_t* bar = &bars[seq_p % kBars];
          if (lane == 0) mbar_expect_tx(bar, (uint32_t)(n * kSlot));
          __syncwarp();
          if (mine) {
            int slot = head_p + rank;
            slot = slot >= NS ? slot - NS : slot;
            bulk_g2s(ring_s + (uint32_t)(slot * kSlot), (d ? pnb1 : pnb0) + (uint32_t)off, (uint32_t)kSlot, bar);
          }
          head_p += n; head_p = head_p >= NS ? head_p - NS : head_p;
          used += n;
          ++seq_p;
        }
        ++p_u;
      }
      // ---------------- consume unit u ----------------
      const int d = u & 1, lr = warp + KW * (u >> 1);
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
        x.load(&bf.x[q][0][0][lane * 4]);
      }
      __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
      __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
      const int ebase = bf.rowptr[d][0];
      const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
      const int fast_end = min(end, kTmaCap);
      int n;
      {
        bool mine; int rank, off, dd;
        n = unit_ring(bf, u, mine, rank, off, dd);
      }
      if (n > 0) {
        mbar_wait(&bars[seq_c % kBars], (seq_c / kBars) & 1);
        ++seq_c;
      }
      const int2* __restrict__ rc = bf.rc[d];
      float4 S0 = zero4(), S1 = S0, Q0 = S0, Q1 = S0;
      int slot = head_c, taken = 0;
      for (int i = beg; i < fast_end; ++i) {
        const int2 m = rc[i];
        const float c = __int_as_float(m.y);
        float4 v0, v1;
        const bool hot = HOT && (m.x & 1);
        if (hot || taken < n) {
          const unsigned char* base = hot ? reinterpret_cast<const unsigned char*>(&s_hot[d][0])
                                          : ring + slot * kSlot;
          if (!hot) { ++taken; slot = slot + 1 == NS ? 0 : slot + 1; }
          v0 = *reinterpret_cast<const float4*>(base + lane * 16);
          v1 = ld1 ? *reinterpret_cast<const float4*>(base + 512 + lane * 16) : zero4();
        } else {                                         // staged but not in the ring (long rows): direct gather
          const char* a0 = (d ? pnb1 : pnb0) + lane * 16 + (uint32_t)(m.x & ~1);
          v0 = ldg4(a0);
          v1 = ld1 ? ldg4(a0 + 512) : zero4();
        }
        fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
      }
      for (int i = max(beg, kTmaCap); i < end; ++i) {    // slow path: slice overflowed the staging buffer
        const PnDir& dd = p.dir[d];
        const int64_t e = (int64_t)ebase + i;
        const float w = dd.w ? dd.w[e] : 1.0f;
        const float c = w * (w * p.prior[dd.src[e]]);
        const char* a = (d ? pnb1 : pnb0) + lane * 16 + (uint32_t)dd.rel[e] * (uint32_t)kPnRowBytes;
        const float4 v0 = ldg4(a);
        const float4 v1 = ld1 ? ldg4(a + 512) : zero4();
        fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
      }
      head_c += n; head_c = head_c >= NS ? head_c - NS : head_c;
      used -= n;
      const float4 U0 = addsub4(Q0, S0, 1.f), V0 = addsub4(Q0, S0, -1.f);
      const float4 U1 = addsub4(Q1, S1, 1.f), V1 = addsub4(Q1, S1, -1.f);
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int seg = d * SEGP + j * 2 * SEGP;
        emit4(hrow + seg, lrow + seg, true, x.xp[j][0], x.xn[j][0], U0, V0);
        emit4(hrow + seg + 128, lrow + seg + 128, wr1, x.xp[j][1], x.xn[j][1], U1, V1);
      }
    }
    mbar_arrive(&s_empty[it & 1]);     // every consumer thread arrives (count = KW * 32)
  }
}


// ---------------------------------------------------------------------------------------------------------
// Ring kernel (gr_set_option("agg_abs_ws", 20)): three warp roles per CTA.
//   stager  (1 warp)  : as in agg_abs_ws_kernel -- stages tile t+1 (row pointers, {table offset, coefficient} per edge,
//                       relu(+-ins)/2) while tile t is processed.
//   issuer  (1 warp)  : walks the staged tile in the consumers' order, one "block" (KW rows x both directions) at a
//                       time, and issues one 1-D bulk copy (cp.async.bulk -> UBLKCP) per gathered edge: table row ->
//                       next slot of a CTA-wide ring in shared memory; the block's mbarrier carries the byte count.
//                       It also writes, per (row, direction) unit, {first ring slot, #ring edges} and the edge
//                       coefficients in ring order.  It runs ahead of the consumers by as many blocks as fit in the
//                       ring (reclaimed block by block through "empty" mbarriers).
//   consumers (KW)    : one destination row per block each; wait for the block's bytes, accumulate S / Q from the ring
//                       with LDS.128 (no global loads, no per-edge address arithmetic on 64-bit pointers), epilogue as
//                       before.
// Why: ncu on the LDG kernels (profiles/README.md) shows nothing saturated (L1TEX 57 %, issue 48 %, L2 41 %, DRAM 38 %)
// and 58 % of the stall samples on the first use of gathered data; adding consumer warps does not help (18 warps:
// -3 %; 22: +10 %; 28 half-row warps: +40 %) -- the LSU/L1 miss path cannot keep more gather requests in flight.  The
// bulk copies bypass LSU and L1 and their number in flight is bounded by the ring size only.  A first version that
// let every consumer warp issue its own copies (kept below as agg_abs_tma_kernel) removed the long-scoreboard stalls but
// doubled the instruction count (bookkeeping + 8 instructions per copy for the lane -> uniform-register waterfall).
// ---------------------------------------------------------------------------------------------------------
constexpr int kRingUnitCap = 6;    // ring slots one (row, direction) unit may take; further staged edges use LDG
constexpr int kChunks = 4;         // blocks in flight (mbarrier pairs)

template <int NI, int KW>
struct alignas(16) RingBuf {
  int2 rc[2][kTmaCap];
  float x[2][NI][2][kPnCols];
  int32_t rowptr[2][KW * 8 + 4];
  uint32_t ud[2][KW * 8];          // per unit: first ring slot | #ring edges << 16   (written by the issuer)
  int32_t tile;
  int32_t pad_[3];
};

template <int NI, int DT, int SEGP, int KW, int RS>
__global__ void __launch_bounds__((KW + 2) * 32, 2) agg_abs_ring_kernel(const PnParams p, int ntiles) {
  constexpr int ROWS = KW * 8;
  constexpr int kSlot = DT * 4;
  static_assert(DT % 4 == 0 && DT > 128 && DT <= kPnCols && RS >= 2 * KW * kRingUnitCap, "ring must hold a block");
  using Buf = RingBuf<NI, KW>;
  extern __shared__ __align__(16) unsigned char ws_smem[];
  Buf* bufs = reinterpret_cast<Buf*>(ws_smem);
  unsigned char* const ring = ws_smem + 2 * sizeof(Buf);
  float* const ringc = reinterpret_cast<float*>(ring + (size_t)RS * kSlot);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];      // tile buffers (stager <-> issuer + consumers)
  __shared__ __align__(8) uint64_t c_full[kChunks], c_empty[kChunks];   // ring blocks (issuer <-> consumers)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], (KW + 1) * 32); mbar_init(&s_empty[1], (KW + 1) * 32);
    for (int c = 0; c < kChunks; ++c) { mbar_init(&c_full[c], 1); mbar_init(&c_empty[c], KW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == KW) {
    // =============================== stager ===============================
    for (int it = 0;; ++it) {
      Buf& bf = bufs[it & 1];
      if (it >= 2) mbar_wait(&s_empty[it & 1], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[it & 1]);
        break;
      }
      produce_tile_rows<NI, DT, ROWS, kTmaCap>(bf, p, tile, lane);
      __syncwarp();
      mbar_arrive(&s_full[it & 1]);
    }
    return;
  }

  if (warp == KW + 1) {
    // =============================== issuer ===============================
    const char* const pnb0 = reinterpret_cast<const char*>(p.dir[0].pn);
    const char* const pnb1 = reinterpret_cast<const char*>(p.dir[1].pn);
    const uint32_t ring_s = smem_u32(ring);
    uint32_t chunk_seq = 0, oldest = 0;
    int head = 0, inflight = 0;
    int size_hist = 0;                       // sizes of the <= kChunks blocks in flight, 8 bits each (lane-uniform)
    for (int it = 0;; ++it) {
      Buf& bf = bufs[it & 1];
      mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      const int tile = bf.tile;
      if (tile < 0) break;
      const int nrows = (int)min((int64_t)ROWS, p.Nt - (int64_t)tile * ROWS);
      const int nblocks = (nrows + KW - 1) / KW;
      const int eb0 = bf.rowptr[0][0], eb1 = bf.rowptr[1][0];
      for (int b = 0; b < nblocks; ++b) {
        // lane l < 2*KW describes unit (row b*KW + (l >> 1), direction l & 1)
        const int ud_d = lane & 1, ud_row = b * KW + (lane >> 1);
        int ubeg = 0, un = 0;
        if (lane < 2 * KW && ud_row < nrows) {
          const int ebase = ud_d ? eb1 : eb0;
          ubeg = bf.rowptr[ud_d][ud_row] - ebase;
          const int uend = bf.rowptr[ud_d][ud_row + 1] - ebase;
          un = max(0, min(min(uend, kTmaCap) - ubeg, kRingUnitCap));
        }
        int incl = un;                        // inclusive prefix sum over the 2*KW unit lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const int T = __shfl_sync(0xffffffffu, incl, 31);
        const int ustart = incl - un;
        // reclaim ring space / barrier slots of finished blocks
        while (inflight + T > RS || chunk_seq - oldest >= (uint32_t)kChunks) {
          mbar_wait(&c_empty[oldest % kChunks], (oldest / kChunks) & 1);
          inflight -= (size_hist >> (8 * (oldest % kChunks))) & 0xff;
          ++oldest;
        }
        const int ch = chunk_seq % kChunks;
        size_hist = (size_hist & ~(0xff << (8 * ch))) | (T << (8 * ch));
        uint64_t* const bar = &c_full[ch];
        if (lane < 2 * KW && ud_row < nrows) {
          int s0 = head + ustart;
          s0 = s0 >= RS ? s0 - RS : s0;
          bf.ud[ud_d][ud_row] = (uint32_t)s0 | ((uint32_t)un << 16);
        }
        // the copies: unit by unit, lane i = edge i of the unit
#pragma unroll 1
        for (int u = 0; u < 2 * KW; ++u) {
          const int n = __shfl_sync(0xffffffffu, un, u);
          if (n == 0) continue;
          const int beg = __shfl_sync(0xffffffffu, ubeg, u);
          const int st = __shfl_sync(0xffffffffu, ustart, u);
          const int d = u & 1;
          if (lane < n) {
            const int2 m = bf.rc[d][beg + lane];
            int slot = head + st + lane;
            slot = slot >= RS ? slot - RS : slot;
            ringc[slot] = __int_as_float(m.y);
            bulk_g2s(ring_s + (uint32_t)(slot * kSlot), (d ? pnb1 : pnb0) + (uint32_t)m.x, (uint32_t)kSlot, bar);
          }
        }
        // publish: the arrive (release) orders this warp's ud / ringc stores before the consumers' reads; the bytes of
        // copies that already landed were counted negative and are balanced by the expect_tx
        __syncwarp();
        if (lane == 0) {
          if (T > 0) mbar_expect_tx(bar, (uint32_t)(T * kSlot));
          else mbar_arrive1(bar);
        }
        head += T; head = head >= RS ? head - RS : head;
        inflight += T;
        ++chunk_seq;
      }
      // this warp is done reading the tile buffer
      mbar_arrive(&s_empty[it & 1]);
    }
    return;
  }

  // =============================== consumer warps ===============================
  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const char* const tb0 = reinterpret_cast<const char*>(p.dir[0].pn) + lane * 16;
  const char* const tb1 = reinterpret_cast<const char*>(p.dir[1].pn) + lane * 16;
  const unsigned char* const ring_lane = ring + lane * 16;
  LaneIns<NI> x;
  uint32_t chunk_seq = 0;
  for (int it = 0;; ++it) {
    Buf& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * ROWS;
    const int nrows = (int)min((int64_t)ROWS, p.Nt - r0);
    const int nblocks = (nrows + KW - 1) / KW;
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    int cur_q = -1;
    for (int b = 0; b < nblocks; ++b, ++chunk_seq) {
      const int lr = b * KW + warp;
      const int ch = chunk_seq % kChunks;
      if (lr < nrows) {
        const int q = lr >= lr_switch ? 1 : 0;
        if (q != cur_q) {
          cur_q = q;
          x.load(&bf.x[q][0][0][lane * 4]);
        }
        mbar_wait(&c_full[ch], (chunk_seq / kChunks) & 1);
        __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
        __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const uint32_t desc = bf.ud[d][lr];
          int slot = (int)(desc & 0xffffu);
          const int n = (int)(desc >> 16);
          float4 S0 = zero4(), S1 = S0, Q0 = S0, Q1 = S0;
          float4 v1 = zero4();                               // lanes without chunk-1 columns never overwrite it
          for (int i = 0; i < n; ++i) {
            const unsigned char* a = ring_lane + slot * kSlot;
            const float c = ringc[slot];
            const float4 v0 = *reinterpret_cast<const float4*>(a);
            if (ld1) v1 = *reinterpret_cast<const float4*>(a + 512);
            slot = slot + 1 == RS ? 0 : slot + 1;
            fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
          }
          {                                                  // rows longer than the ring share / the staging buffer
            const int ebase = bf.rowptr[d][0];
            const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
            if (end - beg > n) {
              const char* tb = d ? tb1 : tb0;
              const int fast_end = min(end, kTmaCap);
              for (int i = beg + n; i < fast_end; ++i) {
                const int2 m = bf.rc[d][i];
                const char* a0 = tb + (uint32_t)m.x;
                const float4 v0 = ldg4(a0);
                if (ld1) v1 = ldg4(a0 + 512);
                const float c = __int_as_float(m.y);
                fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
              }
              const PnDir& dd = p.dir[d];
              for (int i = max(beg + n, kTmaCap); i < end; ++i) {
                const int64_t e = (int64_t)ebase + i;
                const float w = dd.w ? dd.w[e] : 1.0f;
                const float c = w * (w * p.prior[dd.src[e]]);
                const char* a0 = tb + (uint32_t)dd.rel[e] * (uint32_t)kPnRowBytes;
                const float4 v0 = ldg4(a0);
                if (ld1) v1 = ldg4(a0 + 512);
                fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
              }
            }
          }
          const float4 U0 = addsub4(Q0, S0, 1.f), V0 = addsub4(Q0, S0, -1.f);
          const float4 U1 = addsub4(Q1, S1, 1.f), V1 = addsub4(Q1, S1, -1.f);
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const int seg = d * SEGP + j * 2 * SEGP;
            emit4(hrow + seg, lrow + seg, true, x.xp[j][0], x.xn[j][0], U0, V0);
            emit4(hrow + seg + 128, lrow + seg + 128, wr1, x.xp[j][1], x.xn[j][1], U1, V1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive1(&c_empty[ch]);               // this warp's ring slots of the block are free
    }
    mbar_arrive(&s_empty[it & 1]);
  }
}

template <int NI, int KW, int RS>
int launch_ring(const PnParams& p, cudaStream_t stream) {
  auto kern = agg_abs_ring_kernel<NI, 200, 208, KW, RS>;
  const size_t smem = 2 * sizeof(RingBuf<NI, KW>) + (size_t)RS * 200 * 4 + (size_t)RS * 4;
  static bool attr_set = false;
  if (!attr_set) {
    GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
  const unsigned tiles = (unsigned)ceil_div(p.Nt, KW * 8);
  const unsigned pgrid = std::min<unsigned>(tiles, 2u * (unsigned)sm_count());
  kern<<<pgrid, (KW + 2) * 32, smem, stream>>>(p, (int)tiles);
  GR_CHECK_LAUNCH();
  return GR_OK;
}


// ---------------------------------------------------------------------------------------------------------
// TMA-gather kernel, second version (gr_set_option("agg_abs_ws", 30)): per-warp double-buffered stages.
//
// scripts/micro/gather_bw.cu (profiles/r2_gather_bw.txt): random 800-byte rows of the L2-resident table are gathered
// at 10 TB/s by 16 warps/SM of LDG.128 (independent of the loads in flight per warp), 16 TB/s by 32 warps/SM -- and at
// 15.7 TB/s by 16 warps/SM that issue one cp.async.bulk per row into shared memory.  The aggregation kernel cannot
// have 32 warps (96 registers), so the gather moves to the bulk-copy path and the kernel becomes issue-bound: the
// whole design below is about instructions per (row, direction) unit.
//   * every consumer warp owns two stages of NSS slots (800 B each) and one mbarrier per stage; while it accumulates
//     unit g from stage g & 1 the copies of unit g + 1 are in flight into the other stage (fixed distance 1: no ring
//     bookkeeping).  Lane i issues the copy of edge i (ptxas serialises the lanes through ELECT / R2UR: UBLKCP takes
//     uniform registers, ~8 instructions per copy).
//   * the accumulation loop is unrolled over the NSS slots: every shared-memory address is the lane's stage base plus
//     an immediate, coefficients are LDS.32 broadcasts; rows with more staged edges than NSS finish through LDG.
// ---------------------------------------------------------------------------------------------------------
constexpr int kStageSlots = 6;

__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds1(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

template <int NI, int DT, int SEGP, int KW, int ROWS>
__global__ void __launch_bounds__((KW + 1) * 32, 2) agg_abs_tma2_kernel(const PnParams p, int ntiles) {
  static_assert(DT % 4 == 0 && DT > 128 && DT <= kPnCols, "two column chunks of 128");
  constexpr int NSS = kStageSlots;
  constexpr int kSlot = DT * 4;
  struct alignas(16) Buf {
    int2 rc[2][kTmaCap];
    float x[2][NI][2][kPnCols];
    int32_t rowptr[2][ROWS + 4];
    int32_t tile;
    int32_t pad_[3];
  };
  extern __shared__ __align__(16) unsigned char ws_smem[];
  Buf* bufs = reinterpret_cast<Buf*>(ws_smem);
  unsigned char* const ring_all = ws_smem + 2 * sizeof(Buf);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  __shared__ __align__(8) uint64_t s_bar[KW][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], KW * 32); mbar_init(&s_empty[1], KW * 32);
    for (int w = 0; w < KW; ++w) { mbar_init(&s_bar[w][0], 1); mbar_init(&s_bar[w][1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == KW) {
    for (int it = 0;; ++it) {
      Buf& bf = bufs[it & 1];
      if (it >= 2) mbar_wait(&s_empty[it & 1], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[it & 1]);
        break;
      }
      produce_tile_rows<NI, DT, ROWS, kTmaCap>(bf, p, tile, lane);
      __syncwarp();
      mbar_arrive(&s_full[it & 1]);
    }
    return;
  }

  // =============================== consumer warps ===============================
  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const char* const pnb0 = reinterpret_cast<const char*>(p.dir[0].pn);
  const char* const pnb1 = reinterpret_cast<const char*>(p.dir[1].pn);
  const uint32_t stage_s = smem_u32(ring_all) + (uint32_t)warp * 2u * NSS * kSlot;   // this warp's two stages
  const uint32_t stage_lane = stage_s + lane * 16;
  uint64_t* const bars = s_bar[warp];
  LaneIns<NI> x;
  uint32_t g = 0;                                             // units issued by this warp so far (stage = g & 1)

  for (int it = 0;; ++it) {
    Buf& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * ROWS;
    const int nrows = (int)min((int64_t)ROWS, p.Nt - r0);
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    const int nunits = warp < nrows ? 2 * ((nrows - warp + KW - 1) / KW) : 0;
    const int eb0 = bf.rowptr[0][0], eb1 = bf.rowptr[1][0];
    const uint32_t rc_s = smem_u32(&bf.rc[0][0]);

    // issue the bulk copies of unit u (first <= NSS staged edges) into stage gi & 1; returns beg | n << 16
    auto issue = [&](int u, uint32_t gi, int& beg, int& end) -> int {
      const int d = u & 1, lr = warp + KW * (u >> 1);
      const int ebase = d ? eb1 : eb0;
      beg = bf.rowptr[d][lr] - ebase;
      end = bf.rowptr[d][lr + 1] - ebase;
      const int n = max(0, min(min(end, kTmaCap) - beg, NSS));
      uint64_t* const bar = &bars[gi & 1];
      if (lane == 0) {
        if (n > 0) mbar_expect_tx(bar, (uint32_t)(n * kSlot));
        else mbar_arrive1(bar);
      }
      if (lane < n) {
        const int off = bf.rc[d][beg + lane].x;
        bulk_g2s(stage_s + (uint32_t)(((gi & 1) * NSS + lane) * kSlot), (d ? pnb1 : pnb0) + (uint32_t)off,
                 (uint32_t)kSlot, bar);
      }
      return n;
    };

    int cur_q = -1;
    int beg = 0, end = 0, n = 0, nbeg = 0, nend = 0, nn = 0;
    if (nunits > 0) n = issue(0, g, beg, end);
    for (int u = 0; u < nunits; ++u, ++g) {
      if (u + 1 < nunits) nn = issue(u + 1, g + 1, nbeg, nend);
      const int d = u & 1, lr = warp + KW * (u >> 1);
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
        x.load(&bf.x[q][0][0][lane * 4]);
      }
      mbar_wait(&bars[g & 1], (g >> 1) & 1);
      const uint32_t sl = stage_lane + (g & 1) * (NSS * kSlot);
      const uint32_t cy = rc_s + (uint32_t)((d * kTmaCap + beg) * 8 + 4);        // &rc[d][beg].y
      float4 S0 = zero4(), S1 = S0, Q0 = S0, Q1 = S0;
      float4 v1 = zero4();                                    // lanes without chunk-1 columns never overwrite it
#pragma unroll
      for (int i = 0; i < NSS; ++i) {
        if (i < n) {
          const float c = lds1(cy + i * 8);
          const float4 v0 = lds4(sl + i * kSlot);
          if (ld1) v1 = lds4(sl + i * kSlot + 512);
          fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
        }
      }
      if (end - beg > n) {                                    // long rows: staged edges beyond the stage, then the rest
        const char* tb = (d ? pnb1 : pnb0) + lane * 16;
        const int fast_end = min(end, kTmaCap);
        for (int i = beg + n; i < fast_end; ++i) {
          const int2 m = bf.rc[d][i];
          const char* a0 = tb + (uint32_t)m.x;
          const float4 v0 = ldg4(a0);
          if (ld1) v1 = ldg4(a0 + 512);
          const float c = __int_as_float(m.y);
          fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
        }
        const PnDir& dd = p.dir[d];
        const int ebase = d ? eb1 : eb0;
        for (int i = max(beg + n, kTmaCap); i < end; ++i) {
          const int64_t e = (int64_t)ebase + i;
          const float w = dd.w ? dd.w[e] : 1.0f;
          const float c = w * (w * p.prior[dd.src[e]]);
          const char* a0 = tb + (uint32_t)dd.rel[e] * (uint32_t)kPnRowBytes;
          const float4 v0 = ldg4(a0);
          if (ld1) v1 = ldg4(a0 + 512);
          fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
        }
      }
      __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
      __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
      const float4 U0 = addsub4(Q0, S0, 1.f), V0 = addsub4(Q0, S0, -1.f);
      const float4 U1 = addsub4(Q1, S1, 1.f), V1 = addsub4(Q1, S1, -1.f);
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int seg = d * SEGP + j * 2 * SEGP;
        emit4(hrow + seg, lrow + seg, true, x.xp[j][0], x.xn[j][0], U0, V0);
        emit4(hrow + seg + 128, lrow + seg + 128, wr1, x.xp[j][1], x.xn[j][1], U1, V1);
      }
      beg = nbeg; end = nend; n = nn;
    }
    mbar_arrive(&s_empty[it & 1]);
  }
}

template <int NI, int KW, int ROWS>
int launch_tma2(const PnParams& p, cudaStream_t stream) {
  auto kern = agg_abs_tma2_kernel<NI, 200, 208, KW, ROWS>;
  const size_t buf = (sizeof(int2) * 2 * kTmaCap + sizeof(float) * 2 * NI * 2 * kPnCols + 4 * 2 * (ROWS + 4) + 16 + 15) / 16 * 16;
  const size_t smem = 2 * buf + (size_t)KW * 2 * kStageSlots * 200 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
  const unsigned tiles = (unsigned)ceil_div(p.Nt, ROWS);
  const unsigned pgrid = std::min<unsigned>(tiles, 2u * (unsigned)sm_count());
  kern<<<pgrid, (KW + 1) * 32, smem, stream>>>(p, (int)tiles);
  GR_CHECK_LAUNCH();
  return GR_OK;
}


// ---------------------------------------------------------------------------------------------------------
// TMA-gather kernel, third version (gr_set_option("agg_abs_ws", 32)): tile::gather4.
//
// scripts/micro/gather4_test.cu (profiles/r2_gather4_test.txt): `cp.async.bulk.tensor.2d ... tile::gather4` with a
// {D columns, 1 row} box fetches FOUR table rows, given by four row coordinates, into 4 x 800 contiguous bytes; a row
// coordinate beyond the tensor is zero-filled without a memory read and still counts its bytes on the mbarrier.  16
// warps/SM issuing two of them per 8 edges gather at 16.5 TB/s -- with a quarter of the copy instructions of the
// per-row bulk copies (each costs ~8-12 issue slots: UBLKCP / UTMALDG take uniform registers, ptxas walks the lanes).
// Structure: one CTA per SM, 14 consumer warps + the staging warp; every consumer owns two stages of 8 slots (two
// gather4 groups) and prefetches exactly one (row, direction) unit ahead, also across the tile boundary (non-blocking
// peek at the next staged tile).  The accumulation loop is a real loop over the unit's edges (ptxas predicates an
// unrolled `if (i < n)` chain: all 6 bodies issued for 4 edges on average -- measured, 340 instructions per unit).
// ---------------------------------------------------------------------------------------------------------
constexpr int kG4Slots = 8;                 // slots per stage = two gather4 groups
constexpr int kOobRow = 0x3fffffff;         // row coordinate outside any table: zero fill, no memory traffic

typedef CUresult (*AggEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_table_tmap(CUtensorMap* m, const float* pn, int64_t rows, int cols) {
  static AggEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<AggEncodeTiledFn>(q);
  }
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)kPnCols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kPnRowBytes};
  cuuint32_t box[2] = {(cuuint32_t)cols, 1u};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(pn), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int r0, int r1, int r2,
                                            int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(0), "r"(r0), "r"(r1),
      "r"(r2), "r"(r3)
      : "memory");
}

template <int NI, int ROWS>
struct alignas(16) G4Buf {
  int2 rc[2][kTmaCap];               // {relation row index, coefficient}
  float x[2][NI][2][kPnCols];
  int32_t rowptr[2][ROWS + 4];
  int32_t tile;
  int32_t nrows;
  int32_t pad_[2];
};

template <int NI, int DT, int SEGP, int KW, int RPW>
__global__ void __launch_bounds__((KW + 1) * 32, 1)
agg_abs_tma3_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                    const PnParams p, int ntiles) {
  static_assert(DT % 4 == 0 && DT > 128 && DT <= kPnCols, "two column chunks of 128");
  constexpr int ROWS = KW * RPW;
  constexpr int NSS = kG4Slots;
  constexpr int kSlot = DT * 4;
  using Buf = G4Buf<NI, ROWS>;
  extern __shared__ __align__(128) unsigned char ws_smem_raw[];
  unsigned char* const ring_all = ws_smem_raw + ((128u - (smem_u32(ws_smem_raw) & 127u)) & 127u);
  Buf* bufs = reinterpret_cast<Buf*>(ring_all + (size_t)KW * 2 * NSS * kSlot);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  __shared__ __align__(8) uint64_t s_bar[KW][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], KW * 32); mbar_init(&s_empty[1], KW * 32);
    for (int w = 0; w < KW; ++w) { mbar_init(&s_bar[w][0], 1); mbar_init(&s_bar[w][1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == KW) {
    // =============================== staging warp ===============================
    for (int it = 0;; ++it) {
      Buf& bf = bufs[it & 1];
      if (it >= 2) mbar_wait(&s_empty[it & 1], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[it & 1]);
        break;
      }
      if (lane == 0) bf.nrows = (int)min((int64_t)ROWS, p.Nt - (int64_t)tile * ROWS);
      produce_tile_rows<NI, DT, ROWS, kTmaCap, true>(bf, p, tile, lane);
      __syncwarp();
      mbar_arrive(&s_full[it & 1]);
    }
    return;
  }

  // =============================== consumer warps ===============================
  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const char* const pnb0 = reinterpret_cast<const char*>(p.dir[0].pn) + lane * 16;
  const char* const pnb1 = reinterpret_cast<const char*>(p.dir[1].pn) + lane * 16;
  const uint32_t stage_s = smem_u32(ring_all) + (uint32_t)warp * 2u * NSS * kSlot;
  const uint32_t stage_lane = stage_s + lane * 16;
  uint64_t* const bars = s_bar[warp];
  LaneIns<NI> x;
  uint32_t g = 0;                     // units issued so far minus the one in flight: unit g lives in stage g & 1
  int beg = 0, end = 0, n = 0;        // the unit whose copies are in flight / about to be consumed
  bool pre = false;                   // unit 0 of the current tile was issued while finishing the previous tile

  // issue the gather of unit u of the tile staged in b into stage gi & 1
  auto issue = [&](const Buf& b, int u, uint32_t gi, int& ubeg, int& uend) -> int {
    const int d = u & 1, lr = warp + KW * (u >> 1);
    const int ebase = b.rowptr[d][0];
    ubeg = b.rowptr[d][lr] - ebase;
    uend = b.rowptr[d][lr + 1] - ebase;
    const int un = max(0, min(min(uend, kTmaCap) - ubeg, NSS));
    const int ng = (un + 3) >> 2;
    uint64_t* const bar = &bars[gi & 1];
    if (lane == 0) {
      if (ng > 0) mbar_expect_tx(bar, (uint32_t)(ng * 4 * kSlot));
      else mbar_arrive1(bar);
    }
    if (lane < ng) {
      const int2* e = &b.rc[d][ubeg + 4 * lane];
      const int left = un - 4 * lane;                         // >= 1
      const int r0 = e[0].x;
      const int r1 = left > 1 ? e[1].x : kOobRow;
      const int r2 = left > 2 ? e[2].x : kOobRow;
      const int r3 = left > 3 ? e[3].x : kOobRow;
      tma_gather4(stage_s + (uint32_t)(((gi & 1) * NSS + 4 * lane) * kSlot), d ? &map1 : &map0, bar, r0, r1, r2, r3);
    }
    return un;
  };

  for (int it = 0;; ++it) {
    Buf& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * ROWS;
    const int nrows = bf.nrows;
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    const int nunits = warp < nrows ? 2 * ((nrows - warp + KW - 1) / KW) : 0;
    const uint32_t rc_s = smem_u32(&bf.rc[0][0]);
    int cur_q = -1;
    if (nunits > 0 && !pre) n = issue(bf, 0, g, beg, end);
    pre = false;
    for (int u = 0; u < nunits; ++u, ++g) {
      int nbeg = 0, nend = 0, nn = 0;
      if (u + 1 < nunits) {
        nn = issue(bf, u + 1, g + 1, nbeg, nend);
      } else if (mbar_test(&s_full[(it + 1) & 1], ((it + 1) >> 1) & 1)) {   // last unit: peek at the next staged tile
        const Buf& nb = bufs[(it + 1) & 1];
        if (nb.tile >= 0 && warp < nb.nrows) {
          nn = issue(nb, 0, g + 1, nbeg, nend);
          pre = true;
        }
      }
      const int d = u & 1, lr = warp + KW * (u >> 1);
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
        x.load(&bf.x[q][0][0][lane * 4]);
      }
      mbar_wait(&bars[g & 1], (g >> 1) & 1);
      uint32_t sl = stage_lane + (g & 1) * (NSS * kSlot);
      uint32_t cy = rc_s + (uint32_t)((d * kTmaCap + beg) * 8 + 4);        // &rc[d][beg].y
      float4 S0 = zero4(), S1 = S0, Q0 = S0, Q1 = S0;
      float4 v1 = zero4();                                    // lanes without chunk-1 columns never overwrite it
#pragma unroll 1
      for (int i = 0; i < n; ++i, sl += kSlot, cy += 8) {
        const float c = lds1(cy);
        const float4 v0 = lds4(sl);
        if (ld1) v1 = lds4(sl + 512);
        fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
      }
      if (end - beg > n) {                                    // long rows: staged edges beyond the stage, then the rest
        const char* tb = d ? pnb1 : pnb0;
        const int fast_end = min(end, kTmaCap);
        for (int i = beg + n; i < fast_end; ++i) {
          const int2 m = bf.rc[d][i];
          const char* a0 = tb + (size_t)(uint32_t)m.x * kPnRowBytes;
          const float4 v0 = ldg4(a0);
          if (ld1) v1 = ldg4(a0 + 512);
          const float c = __int_as_float(m.y);
          fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
        }
        const PnDir& dd = p.dir[d];
        const int ebase = bf.rowptr[d][0];
        for (int i = max(beg + n, kTmaCap); i < end; ++i) {
          const int64_t e = (int64_t)ebase + i;
          const float w = dd.w ? dd.w[e] : 1.0f;
          const float c = w * (w * p.prior[dd.src[e]]);
          const char* a0 = tb + (size_t)(uint32_t)dd.rel[e] * kPnRowBytes;
          const float4 v0 = ldg4(a0);
          if (ld1) v1 = ldg4(a0 + 512);
          fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
        }
      }
      __nv_bfloat16* const hrow = hi_lane + (int64_t)lr * p.ld;
      __nv_bfloat16* const lrow = lo_lane + (int64_t)lr * p.ld;
      const float4 U0 = addsub4(Q0, S0, 1.f), V0 = addsub4(Q0, S0, -1.f);
      const float4 U1 = addsub4(Q1, S1, 1.f), V1 = addsub4(Q1, S1, -1.f);
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int seg = d * SEGP + j * 2 * SEGP;
        emit4(hrow + seg, lrow + seg, true, x.xp[j][0], x.xn[j][0], U0, V0);
        emit4(hrow + seg + 128, lrow + seg + 128, wr1, x.xp[j][1], x.xn[j][1], U1, V1);
      }
      beg = nbeg; end = nend; n = nn;
    }
    mbar_arrive(&s_empty[it & 1]);
  }
}

template <int NI, int KW, int RPW>
int launch_tma3(const PnParams& p, cudaStream_t stream) {
  auto kern = agg_abs_tma3_kernel<NI, 200, 208, KW, RPW>;
  const size_t smem = 128 + (size_t)KW * 2 * kG4Slots * 200 * 4 + 2 * sizeof(G4Buf<NI, KW * RPW>);
  static bool attr_set = false;
  if (!attr_set) {
    GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  CUtensorMap m0, m1;
  if (!make_table_tmap(&m0, p.dir[0].pn, p.table_rows, 200) || !make_table_tmap(&m1, p.dir[1].pn, p.table_rows, 200)) {
    set_error("gr_aggregate_dual_abs: cuTensorMapEncodeTiled failed for the padded relation table");
    return GR_ERR_CUDA;
  }
  GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
  const unsigned tiles = (unsigned)ceil_div(p.Nt, KW * RPW);
  const unsigned pgrid = std::min<unsigned>(tiles, (unsigned)sm_count());
  kern<<<pgrid, (KW + 1) * 32, smem, stream>>>(m0, m1, p, (int)tiles);
  GR_CHECK_LAUNCH();
  return GR_OK;
}


// ---------------------------------------------------------------------------------------------------------
// gather4 kernel, refined (gr_set_option("agg_abs_ws", 33)).  What the profile of agg_abs_tma3_kernel said
// (profiles/README.md): 308 instructions per unit -- 75 to issue the prefetch (row pointers, clamping, address
// arithmetic, the coordinate loads), 30 around the mbarrier wait, 20 CS2R of accumulator zeroing, 17 per edge, 94 in
// the epilogue -- and 15 % of the stall samples on the tile hand-over: one staging warp cannot feed 14 consumers.
// Here:
//   * two staging warps (even / odd tiles; a shared-memory turn counter keeps their tile grabs in hand-over order);
//   * the staging warp leaves, per (row, direction) unit, a descriptor {first staged edge, #stage edges, long-row flag}
//     and the two ready-made gather4 coordinate quads (out-of-table coordinates where the unit has fewer edges), so a
//     consumer's prefetch is one LDS + one LDS.128 + the copy;
//   * the direction loop is unrolled inside the row loop: descriptor / quad / tensor-map addresses are immediates;
//   * accumulators start from the first edge's products (no zeroing), two edges per loop trip.
// ---------------------------------------------------------------------------------------------------------
template <int NI, int ROWS>
struct alignas(16) G5Buf {
  int4 quad[2][ROWS][2];             // gather4 row coordinates of the unit's stage edges 0-3 / 4-7
  int2 rc[2][kTmaCap];               // {relation row index, coefficient}
  float x[2][NI][2][kPnCols];
  int32_t rowptr[2][ROWS + 4];
  uint32_t ud[2][ROWS];              // first staged edge | #stage edges << 16 | (unit has more edges than that) << 31
  int32_t tile;
  int32_t nrows;
  int32_t pad_[2];
};

template <int NI, int DT, int SEGP, int KW, int RPW>
__global__ void __launch_bounds__((KW + 2) * 32, 1)
agg_abs_g4_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                  const PnParams p, int ntiles) {
  static_assert(DT % 4 == 0 && DT > 128 && DT <= kPnCols, "two column chunks of 128");
  constexpr int ROWS = KW * RPW;
  constexpr int NSS = kG4Slots;
  constexpr int kSlot = DT * 4;
  using Buf = G5Buf<NI, ROWS>;
  extern __shared__ __align__(128) unsigned char ws_smem_raw[];
  // gather4 destinations must be 128-byte aligned; the dynamic segment only follows the static one at 16 bytes
  unsigned char* const ring_all = ws_smem_raw + ((128u - (smem_u32(ws_smem_raw) & 127u)) & 127u);
  Buf* bufs = reinterpret_cast<Buf*>(ring_all + (size_t)KW * 2 * NSS * kSlot);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  __shared__ __align__(8) uint64_t s_bar[KW][2];
  __shared__ volatile int s_turn;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], KW * 32); mbar_init(&s_empty[1], KW * 32);
    for (int w = 0; w < KW; ++w) { mbar_init(&s_bar[w][0], 1); mbar_init(&s_bar[w][1], 1); }
    s_turn = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= KW) {
    // =============================== staging warps: warp KW -> even iterations, KW + 1 -> odd ===============
    const int par = warp - KW;
    Buf& bf = bufs[par];
    for (int it = par;; it += 2) {
      if (it >= 2) mbar_wait(&s_empty[par], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) {
        while (s_turn != it) __nanosleep(20);                 // tiles are grabbed in hand-over order
        tile = atomicAdd(p.tile_counter, 1);
        __threadfence_block();
        s_turn = it + 1;
      }
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[par]);
        break;
      }
      const int nrows = (int)min((int64_t)ROWS, p.Nt - (int64_t)tile * ROWS);
      if (lane == 0) bf.nrows = nrows;
      produce_tile_rows<NI, DT, ROWS, kTmaCap, true>(bf, p, tile, lane);
      __syncwarp();
      for (int un = lane; un < 2 * nrows; un += 32) {          // unit descriptors + gather4 coordinate quads
        const int d = un >= nrows ? 1 : 0, lr = un - d * nrows;
        const int ebase = bf.rowptr[d][0];
        const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
        const int n = max(0, min(min(end, kTmaCap) - beg, NSS));
        int r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = k < n ? bf.rc[d][beg + k].x : kOobRow;
        bf.quad[d][lr][0] = make_int4(r[0], r[1], r[2], r[3]);
        bf.quad[d][lr][1] = make_int4(r[4], r[5], r[6], r[7]);
        bf.ud[d][lr] = (uint32_t)min(beg, kTmaCap) | ((uint32_t)n << 16) | (end - beg > n ? 0x80000000u : 0u);
      }
      __syncwarp();
      mbar_arrive(&s_full[par]);
    }
    return;
  }

  // =============================== consumer warps ===============================
  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const uint32_t stage_s = smem_u32(ring_all) + (uint32_t)warp * 2u * NSS * kSlot;   // stage 0: direction 0, stage 1: 1
  const uint32_t stage_lane = stage_s + lane * 16;
  uint64_t* const bars = s_bar[warp];
  LaneIns<NI> x;
  uint32_t ph = 0;                    // rows consumed so far by this warp: both stage barriers are at phase parity ph & 1
  uint32_t desc0 = 0;                 // descriptor of the direction-0 unit in flight
  bool pre = false;

  // prefetch unit (lr, D) of the tile staged in b into stage D; returns its descriptor
  auto issue = [&](const Buf& b, int lr, auto dir) -> uint32_t {
    constexpr int D_ = decltype(dir)::value;
    const uint32_t desc = b.ud[D_][lr];
    const int ng = (int)(((desc >> 16) & 0xffu) + 3u) >> 2;
    uint64_t* const bar = &bars[D_];
    if (lane == 0) {
      if (ng > 0) mbar_expect_tx(bar, (uint32_t)(ng * 4 * kSlot));
      else mbar_arrive1(bar);
    }
    if (lane < ng) {
      const int4 q = b.quad[D_][lr][lane];
      tma_gather4(stage_s + (uint32_t)((D_ * NSS + 4 * lane) * kSlot), D_ ? &map1 : &map0, bar, q.x, q.y, q.z, q.w);
    }
    return desc;
  };

  // accumulate + emit one unit from its stage
  auto consume = [&](Buf& bf, int lr, auto dir, uint32_t desc, __nv_bfloat16* hrow, __nv_bfloat16* lrow) {
    constexpr int D_ = decltype(dir)::value;
    const int beg = (int)(desc & 0xffffu), n = (int)((desc >> 16) & 0xffu);
    uint32_t sl = stage_lane + D_ * (NSS * kSlot);
    uint32_t cy = smem_u32(&bf.rc[D_][0]) + (uint32_t)(beg * 8 + 4);
    float4 S0, S1, Q0, Q1;
    float4 v1 = zero4();                                      // lanes without chunk-1 columns never overwrite it
    if (n > 0) {
      const float c = lds1(cy);
      const float4 v0 = lds4(sl);
      if (ld1) v1 = lds4(sl + 512);
      S0 = zero4(); S1 = S0; Q0 = S0; Q1 = S0;
      fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
      int i = 1;
#pragma unroll 1
      for (; i + 1 < n; i += 2) {
        sl += 2 * kSlot; cy += 16;
        const float ca = lds1(cy - 8), cb = lds1(cy);
        const float4 a0 = lds4(sl - kSlot), b0 = lds4(sl);
        float4 a1 = v1, b1 = v1;
        if (ld1) { a1 = lds4(sl - kSlot + 512); b1 = lds4(sl + 512); }
        fma4(S0, ca, a0); fma4_abs(Q0, ca, a0); fma4(S1, ca, a1); fma4_abs(Q1, ca, a1);
        fma4(S0, cb, b0); fma4_abs(Q0, cb, b0); fma4(S1, cb, b1); fma4_abs(Q1, cb, b1);
      }
      if (i < n) {
        sl += kSlot; cy += 8;
        const float cc = lds1(cy);
        const float4 c0 = lds4(sl);
        if (ld1) v1 = lds4(sl + 512);
        fma4(S0, cc, c0); fma4_abs(Q0, cc, c0); fma4(S1, cc, v1); fma4_abs(Q1, cc, v1);
      }
    } else {
      S0 = zero4(); S1 = S0; Q0 = S0; Q1 = S0;
    }
    if (desc >> 31) {                                         // long rows: staged edges beyond the stage, then the rest
      const char* tb = reinterpret_cast<const char*>(p.dir[D_].pn) + lane * 16;
      const int ebase = bf.rowptr[D_][0];
      const int rb = bf.rowptr[D_][lr] - ebase, end = bf.rowptr[D_][lr + 1] - ebase;
      const int fast_end = min(end, kTmaCap);
      for (int i = rb + n; i < fast_end; ++i) {
        const int2 m = bf.rc[D_][i];
        const char* a0 = tb + (size_t)(uint32_t)m.x * kPnRowBytes;
        const float4 v0 = ldg4(a0);
        if (ld1) v1 = ldg4(a0 + 512);
        const float c = __int_as_float(m.y);
        fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
      }
      const PnDir& dd = p.dir[D_];
      for (int i = max(rb + n, kTmaCap); i < end; ++i) {
        const int64_t e = (int64_t)ebase + i;
        const float w = dd.w ? dd.w[e] : 1.0f;
        const float c = w * (w * p.prior[dd.src[e]]);
        const char* a0 = tb + (size_t)(uint32_t)dd.rel[e] * kPnRowBytes;
        const float4 v0 = ldg4(a0);
        if (ld1) v1 = ldg4(a0 + 512);
        fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
      }
    }
    const float4 U0 = addsub4(Q0, S0, 1.f), V0 = addsub4(Q0, S0, -1.f);
    const float4 U1 = addsub4(Q1, S1, 1.f), V1 = addsub4(Q1, S1, -1.f);
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int seg = D_ * SEGP + j * 2 * SEGP;
      emit4(hrow + seg, lrow + seg, true, x.xp[j][0], x.xn[j][0], U0, V0);
      emit4(hrow + seg + 128, lrow + seg + 128, wr1, x.xp[j][1], x.xn[j][1], U1, V1);
    }
  };
  using Dir0 = std::integral_constant<int, 0>;
  using Dir1 = std::integral_constant<int, 1>;

  for (int it = 0;; ++it) {
    Buf& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * ROWS;
    const int nrows = bf.nrows;
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    const int64_t ld = p.ld;
    __nv_bfloat16* hrow = p.out_hi + (r0 + warp) * ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* lrow = p.out_lo + (r0 + warp) * ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    int cur_q = -1;
    if (warp < nrows && !pre) desc0 = issue(bf, warp, Dir0{});
    pre = false;
    for (int lr = warp; lr < nrows; lr += KW, ++ph, hrow += KW * ld, lrow += KW * ld) {
      const uint32_t desc1 = issue(bf, lr, Dir1{});           // direction 1 of this row flies while direction 0 is consumed
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
        x.load(&bf.x[q][0][0][lane * 4]);
      }
      mbar_wait(&bars[0], ph & 1);
      consume(bf, lr, Dir0{}, desc0, hrow, lrow);
      if (lr + KW < nrows) {                                  // direction 0 of the next row
        desc0 = issue(bf, lr + KW, Dir0{});
      } else if (mbar_test(&s_full[(it + 1) & 1], ((it + 1) >> 1) & 1)) {     // ... or of the next staged tile
        const Buf& nb = bufs[(it + 1) & 1];
        if (nb.tile >= 0 && warp < nb.nrows) {
          desc0 = issue(nb, warp, Dir0{});
          pre = true;
        }
      }
      mbar_wait(&bars[1], ph & 1);
      consume(bf, lr, Dir1{}, desc1, hrow, lrow);
    }
    mbar_arrive(&s_empty[it & 1]);
  }
}

template <int NI, int KW, int RPW>
int launch_g4(const PnParams& p, cudaStream_t stream) {
  auto kern = agg_abs_g4_kernel<NI, 200, 208, KW, RPW>;
  const size_t smem = 128 + (size_t)KW * 2 * kG4Slots * 200 * 4 + 2 * sizeof(G5Buf<NI, KW * RPW>);
  static bool attr_set = false;
  if (!attr_set) {
    GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  CUtensorMap m0, m1;
  if (!make_table_tmap(&m0, p.dir[0].pn, p.table_rows, 200) || !make_table_tmap(&m1, p.dir[1].pn, p.table_rows, 200)) {
    set_error("gr_aggregate_dual_abs: cuTensorMapEncodeTiled failed for the padded relation table");
    return GR_ERR_CUDA;
  }
  GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
  const unsigned tiles = (unsigned)ceil_div(p.Nt, KW * RPW);
  const unsigned pgrid = std::min<unsigned>(tiles, (unsigned)sm_count());
  kern<<<pgrid, (KW + 2) * 32, smem, stream>>>(m0, m1, p, (int)tiles);
  GR_CHECK_LAUNCH();
  return GR_OK;
}


// ---------------------------------------------------------------------------------------------------------
// gather4 kernel with a deeper prefetch (gr_set_option("agg_abs_ws", 34)).
// Three gather kernels with very different instruction counts (tma2 340, tma3 308, g4 303 per unit) all ran 157.5 us:
// 1730 units per SM / 14 warps x 1.27 us.  With the copies of unit k + 1 issued when unit k starts, a unit cannot
// take less than the latency of a TMA gather (issue -> bytes landed -> mbarrier flip -> waiter resumes), ~1.3 us here,
// three times the LDG round trip.  So the per-warp ring is cut into FOUR groups of four slots (one gather4 each): a
// unit takes one or two groups (its first <= 8 staged edges), its mbarrier is the one of its first group, and a flat
// prefetch cursor keeps issuing ahead (across the tile boundary) while groups are free: 2-3 units in flight.
// ---------------------------------------------------------------------------------------------------------
template <int NI, int DT, int SEGP, int KW, int RPW>
__global__ void __launch_bounds__((KW + 2) * 32, 1)
agg_abs_g5_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                  const PnParams p, int ntiles) {
  static_assert(DT % 4 == 0 && DT > 128 && DT <= kPnCols, "two column chunks of 128");
  constexpr int ROWS = KW * RPW;
  constexpr int kSlot = DT * 4;
  constexpr int kGroup = 4 * kSlot;            // bytes of one gather4 group
  using Buf = G5Buf<NI, ROWS>;
  extern __shared__ __align__(128) unsigned char ws_smem_raw[];
  unsigned char* const ring_all = ws_smem_raw + ((128u - (smem_u32(ws_smem_raw) & 127u)) & 127u);
  Buf* bufs = reinterpret_cast<Buf*>(ring_all + (size_t)KW * 4 * kGroup);
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
  __shared__ __align__(8) uint64_t s_bar[KW][4];
  __shared__ volatile int s_turn;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  if (tid == 0) {
    mbar_init(&s_full[0], 32); mbar_init(&s_full[1], 32);
    mbar_init(&s_empty[0], KW * 32); mbar_init(&s_empty[1], KW * 32);
    for (int w = 0; w < KW; ++w)
      for (int b = 0; b < 4; ++b) mbar_init(&s_bar[w][b], 1);
    s_turn = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= KW) {
    // =============================== staging warps: warp KW -> even iterations, KW + 1 -> odd ===============
    const int par = warp - KW;
    Buf& bf = bufs[par];
    for (int it = par;; it += 2) {
      if (it >= 2) mbar_wait(&s_empty[par], ((it >> 1) - 1) & 1);
      int tile = 0;
      if (lane == 0) {
        while (s_turn != it) __nanosleep(20);                 // tiles are grabbed in hand-over order
        tile = atomicAdd(p.tile_counter, 1);
        __threadfence_block();
        s_turn = it + 1;
      }
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= ntiles) {
        if (lane == 0) bf.tile = -1;
        __syncwarp();
        mbar_arrive(&s_full[par]);
        break;
      }
      const int nrows = (int)min((int64_t)ROWS, p.Nt - (int64_t)tile * ROWS);
      if (lane == 0) bf.nrows = nrows;
      produce_tile_rows<NI, DT, ROWS, kTmaCap, true>(bf, p, tile, lane);
      __syncwarp();
      for (int un = lane; un < 2 * nrows; un += 32) {          // unit descriptors + gather4 coordinate quads
        const int d = un >= nrows ? 1 : 0, lr = un - d * nrows;
        const int ebase = bf.rowptr[d][0];
        const int beg = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
        const int n = max(0, min(min(end, kTmaCap) - beg, 8));
        int r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = k < n ? bf.rc[d][beg + k].x : kOobRow;
        bf.quad[d][lr][0] = make_int4(r[0], r[1], r[2], r[3]);
        bf.quad[d][lr][1] = make_int4(r[4], r[5], r[6], r[7]);
        bf.ud[d][lr] = (uint32_t)min(beg, kTmaCap) | ((uint32_t)n << 16) | (end - beg > n ? 0x80000000u : 0u);
      }
      __syncwarp();
      mbar_arrive(&s_full[par]);
    }
    return;
  }

  // =============================== consumer warps ===============================
  const bool ld1 = 128 + lane * 4 < DT;
  const bool wr1 = 128 + lane * 4 < SEGP;
  const uint32_t ring_s = smem_u32(ring_all) + (uint32_t)warp * 4u * kGroup;   // this warp's four groups
  const uint32_t ring_lane = ring_s + lane * 16;
  uint64_t* const bars = s_bar[warp];
  LaneIns<NI> x;
  // ring state (all warp-uniform)
  int gh = 0, gc = 0, gfree = 4;      // next group to fill / group of the unit being consumed / free groups
  uint32_t parbits = 0;               // phase parity of the four group barriers
  // prefetch cursor: next unit to issue = (tile iteration pit, row plr, direction pd); pnrows = rows of that tile
  int pit = -1, plr = 0, pd = 0, pnrows = 0;
  bool pend = false;                  // the cursor has seen the end marker

  for (int it = 0;; ++it) {
    Buf& bf = bufs[it & 1];
    mbar_wait(&s_full[it & 1], (it >> 1) & 1);
    const int tile = bf.tile;
    if (tile < 0) break;
    const int64_t r0 = (int64_t)tile * ROWS;
    const int nrows = bf.nrows;
    const int b0 = (int)(r0 / N);
    const int lr_switch = N - (int)(r0 - (int64_t)b0 * N);
    const int64_t ld = p.ld;
    __nv_bfloat16* hrow = p.out_hi + (r0 + warp) * ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    __nv_bfloat16* lrow = p.out_lo + (r0 + warp) * ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
    if (pit < it) { pit = it; plr = warp; pd = 0; pnrows = nrows; }
    int cur_q = -1;
    for (int lr = warp; lr < nrows; lr += KW, hrow += KW * ld, lrow += KW * ld) {
      const int q = lr >= lr_switch ? 1 : 0;
      if (q != cur_q) {
        cur_q = q;
        x.load(&bf.x[q][0][0][lane * 4]);
      }
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        // ---------------- keep the ring full ----------------
        for (;;) {
          if (plr >= pnrows) {                                // cursor at the end of its tile
            if (pend || pit > it) break;                      // never two tiles ahead (that buffer is this one)
            if (!mbar_test(&s_full[(it + 1) & 1], ((it + 1) >> 1) & 1)) break;
            const Buf& nb = bufs[(it + 1) & 1];
            pit = it + 1; plr = warp; pd = 0;
            if (nb.tile < 0) { pend = true; pnrows = 0; break; }
            pnrows = nb.nrows;
            continue;
          }
          const Buf& pb = bufs[pit & 1];
          const uint32_t pdesc = pb.ud[pd][plr];
          const int ng = (int)(((pdesc >> 16) & 0xffu) + 3u) >> 2;
          if (ng > gfree) break;
          if (ng > 0) {
            uint64_t* const bar = &bars[gh];
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)(ng * kGroup));
            if (lane < ng) {
              const int4 qd = pb.quad[pd][plr][lane];
              tma_gather4(ring_s + (uint32_t)(((gh + lane) & 3) * kGroup), pd ? &map1 : &map0, bar, qd.x, qd.y, qd.z, qd.w);
            }
            gh = (gh + ng) & 3;
            gfree -= ng;
          }
          pd ^= 1;
          if (pd == 0) plr += KW;
        }
        // ---------------- consume unit (lr, d) ----------------
        const uint32_t desc = bf.ud[d][lr];
        const int beg = (int)(desc & 0xffffu), n = (int)((desc >> 16) & 0xffu);
        float4 S0 = zero4(), S1 = S0, Q0 = S0, Q1 = S0;
        float4 v1 = zero4();                                  // lanes without chunk-1 columns never overwrite it
        if (n > 0) {
          mbar_wait(&bars[gc], (parbits >> gc) & 1u);
          parbits ^= 1u << gc;
          uint32_t cy = smem_u32(&bf.rc[d][0]) + (uint32_t)(beg * 8 + 4);
          uint32_t sl = ring_lane + (uint32_t)gc * kGroup;
          const int n0 = min(n, 4);
#pragma unroll 1
          for (int i = 0; i < n0; ++i, sl += kSlot, cy += 8) {
            const float c = lds1(cy);
            const float4 v0 = lds4(sl);
            if (ld1) v1 = lds4(sl + 512);
            fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
          }
          if (n > 4) {
            sl = ring_lane + (uint32_t)((gc + 1) & 3) * kGroup;
#pragma unroll 1
            for (int i = 4; i < n; ++i, sl += kSlot, cy += 8) {
              const float c = lds1(cy);
              const float4 v0 = lds4(sl);
              if (ld1) v1 = lds4(sl + 512);
              fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
            }
          }
          const int ngc = (n + 3) >> 2;
          gc = (gc + ngc) & 3;
          gfree += ngc;
        }
        if (desc >> 31) {                                     // long rows: staged edges beyond the ring, then the rest
          const char* tb = reinterpret_cast<const char*>(p.dir[d].pn) + lane * 16;
          const int ebase = bf.rowptr[d][0];
          const int rb = bf.rowptr[d][lr] - ebase, end = bf.rowptr[d][lr + 1] - ebase;
          const int fast_end = min(end, kTmaCap);
          for (int i = rb + n; i < fast_end; ++i) {
            const int2 m = bf.rc[d][i];
            const char* a0 = tb + (size_t)(uint32_t)m.x * kPnRowBytes;
            const float4 v0 = ldg4(a0);
            if (ld1) v1 = ldg4(a0 + 512);
            const float c = __int_as_float(m.y);
            fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
          }
          const PnDir& dd = p.dir[d];
          for (int i = max(rb + n, kTmaCap); i < end; ++i) {
            const int64_t e = (int64_t)ebase + i;
            const float w = dd.w ? dd.w[e] : 1.0f;
            const float c = w * (w * p.prior[dd.src[e]]);
            const char* a0 = tb + (size_t)(uint32_t)dd.rel[e] * kPnRowBytes;
            const float4 v0 = ldg4(a0);
            if (ld1) v1 = ldg4(a0 + 512);
            fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
          }
        }
        const float4 U0 = addsub4(Q0, S0, 1.f), V0 = addsub4(Q0, S0, -1.f);
        const float4 U1 = addsub4(Q1, S1, 1.f), V1 = addsub4(Q1, S1, -1.f);
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int seg = d * SEGP + j * 2 * SEGP;
          emit4(hrow + seg, lrow + seg, true, x.xp[j][0], x.xn[j][0], U0, V0);
          emit4(hrow + seg + 128, lrow + seg + 128, wr1, x.xp[j][1], x.xn[j][1], U1, V1);
        }
      }
    }
    mbar_arrive(&s_empty[it & 1]);
  }
}

template <int NI, int KW, int RPW>
int launch_g5(const PnParams& p, cudaStream_t stream) {
  auto kern = agg_abs_g5_kernel<NI, 200, 208, KW, RPW>;
  const size_t smem = 128 + (size_t)KW * 16 * 200 * 4 + 2 * sizeof(G5Buf<NI, KW * RPW>);
  static bool attr_set = false;
  if (!attr_set) {
    GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  CUtensorMap m0, m1;
  if (!make_table_tmap(&m0, p.dir[0].pn, p.table_rows, 200) || !make_table_tmap(&m1, p.dir[1].pn, p.table_rows, 200)) {
    set_error("gr_aggregate_dual_abs: cuTensorMapEncodeTiled failed for the padded relation table");
    return GR_ERR_CUDA;
  }
  GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
  const unsigned tiles = (unsigned)ceil_div(p.Nt, KW * RPW);
  const unsigned pgrid = std::min<unsigned>(tiles, (unsigned)sm_count());
  kern<<<pgrid, (KW + 2) * 32, smem, stream>>>(m0, m1, p, (int)tiles);
  GR_CHECK_LAUNCH();
  return GR_OK;
}

// table [rows, D] fp32 (row stride ldt) -> zero-padded [rows][256]
__global__ void pad_table_kernel(const float* __restrict__ table, int64_t ldt, int64_t rows, int D,
                                 float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row, 4-column group)
  if (i >= rows * (kPnCols / 4)) return;
  const int64_t r = i / (kPnCols / 4);
  const int g = (int)(i % (kPnCols / 4));
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = 4 * g + k < D ? table[r * ldt + 4 * g + k] : 0.f;
  reinterpret_cast<float4*>(out + r * kPnCols)[g] = make_float4(v[0], v[1], v[2], v[3]);
}

template <int NI, int KW, int NS, int MINB, bool HOT>
int launch_tma(const PnParams& p, unsigned grid, cudaStream_t stream) {
  auto kern = agg_abs_tma_kernel<NI, 200, 208, KW, NS, MINB, HOT>;
  const size_t smem = 2 * sizeof(TmaBuf<NI>) + (size_t)KW * NS * 200 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
  const unsigned pgrid = std::min<unsigned>(grid, (unsigned)MINB * (unsigned)sm_count());
  kern<<<pgrid, (KW + 1) * 32, smem, stream>>>(p, (int)grid);
  GR_CHECK_LAUNCH();
  return GR_OK;
}

template <int NI>
int launch_pn(const PnParams& p, cudaStream_t stream) {
  const unsigned grid = (unsigned)ceil_div(p.Nt, kRows);
  if (p.tile_counter && g_opt_agg_abs_ws >= 2 && g_opt_agg_abs_ws <= 5 && NI == 2) {
    switch (g_opt_agg_abs_ws) {
      case 2: return launch_tma<2, 8, 12, 2, false>(p, grid, stream);
      case 3: return launch_tma<2, 8, 12, 2, true>(p, grid, stream);
      case 4: return launch_tma<2, 16, 12, 1, true>(p, grid, stream);
      default: return launch_tma<2, 12, 16, 1, true>(p, grid, stream);
    }
  }
  if constexpr (NI == 2) {
    if (p.tile_counter && g_opt_agg_abs_ws == 20) return launch_ring<2, 8, 96>(p, stream);
    if (p.tile_counter && g_opt_agg_abs_ws == 30) return launch_tma2<2, 8, 64>(p, stream);
    if (p.tile_counter && g_opt_agg_abs_ws == 32) return launch_tma3<2, 14, 4>(p, stream);
    if (p.tile_counter && g_opt_agg_abs_ws == 33) return launch_g4<2, 14, 4>(p, stream);
    if (p.tile_counter && g_opt_agg_abs_ws == 34) return launch_g5<2, 14, 4>(p, stream);
    if (p.tile_counter && g_opt_agg_abs_ws == 35) return launch_g5<2, 14, 8>(p, stream);
    if (p.tile_counter && g_opt_agg_abs_ws == 31) return launch_tma2<2, 8, 128>(p, stream);
    if (p.tile_counter && g_opt_agg_abs_ws >= 10) {
      switch (g_opt_agg_abs_ws) {
        case 10: return launch_wsg<2, 8, 64, 2>(p, stream);
        case 11: return launch_wsg<2, 9, 72, 2>(p, stream);
        case 12: return launch_wsg<2, 11, 66, 2>(p, stream);
        case 13: return launch_wsg<2, 19, 76, 1>(p, stream);
        case 14: return launch_wsg<2, 7, 56, 3>(p, stream);
        default: return launch_wsg<2, 9, 63, 2>(p, stream);
      }
    }
  }
  if (p.tile_counter && g_opt_agg_abs_ws == 7 && NI <= 2) {
    constexpr int SLOTS = 7, RPS = 8;
    using Buf = HBuf<NI, SLOTS * RPS>;
    auto kern = agg_abs_half_kernel<NI, 200, 208, SLOTS, RPS>;
    const size_t smem = 2 * sizeof(Buf);
    static bool attr_set = false;
    if (!attr_set) {
      GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
    const unsigned tiles = (unsigned)ceil_div(p.Nt, SLOTS * RPS);
    const unsigned pgrid = std::min<unsigned>(tiles, 2u * (unsigned)sm_count());
    kern<<<pgrid, (2 * SLOTS + 1) * 32, smem, stream>>>(p, (int)tiles);
    GR_CHECK_LAUNCH();
    return GR_OK;
  }
  if (p.tile_counter && g_opt_agg_abs_ws == 6) {
    const size_t smem = 2 * sizeof(WsBuf<NI>);
    static bool attr_set = false;
    if (!attr_set) {
      GR_CHECK_CUDA(cudaFuncSetAttribute(agg_abs_ws2_kernel<NI, 200, 208>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
    const unsigned pgrid = std::min<unsigned>(grid, 2u * (unsigned)sm_count());
    agg_abs_ws2_kernel<NI, 200, 208><<<pgrid, kWsThreads, smem, stream>>>(p, (int)grid);
    GR_CHECK_LAUNCH();
    return GR_OK;
  }
  if (p.tile_counter && g_opt_agg_abs_ws) {
    const size_t smem = 2 * sizeof(WsBuf<NI>);
    static bool attr_set = false;
    if (!attr_set) {
      GR_CHECK_CUDA(cudaFuncSetAttribute(agg_abs_ws_kernel<NI, 200, 208>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    GR_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream));
    const unsigned pgrid = std::min<unsigned>(grid, 2u * (unsigned)sm_count());
    agg_abs_ws_kernel<NI, 200, 208><<<pgrid, kWsThreads, smem, stream>>>(p, (int)grid);
    GR_CHECK_LAUNCH();
    return GR_OK;
  }
  agg_abs_kernel<NI, 200, 208><<<grid, kThreads, 0, stream>>>(p);
  GR_CHECK_LAUNCH();
  return GR_OK;
}

}  // namespace
}  // namespace gr

extern "C" int gr_pad_table256(const float* table, int64_t ldt, int64_t rows, int D, float* pn, void* stream_) {
  using namespace gr;
  GR_CHECK_ARG(table && pn, "null pointer");
  GR_CHECK_ARG(rows > 0 && D > 0 && D <= kPnCols && ldt >= D, "bad shape (D <= 256)");
  GR_CHECK_ARG((reinterpret_cast<uintptr_t>(pn) & 15) == 0, "pn must be 16-byte aligned");
  const int64_t work = rows * (kPnCols / 4);
  pad_table_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      table, ldt, rows, D, pn);
  GR_CHECK_LAUNCH();
  return GR_OK;
}

extern "C" int gr_aggregate_dual_abs_supported(int N, int D, int64_t seg_pitch, int64_t R1) {
  return (D == 200 && seg_pitch == 208 && N >= gr::kRows && R1 > 0 && R1 < (1 << 21)) ? 1 : 0;
}

extern "C" int gr_aggregate_dual_abs(const int32_t* rowptr_t, const int32_t* src_t, const int32_t* rel_t,
                                    const float* w_t, const int32_t* rowptr_h, const int32_t* src_h,
                                    const int32_t* rel_h, const float* w_h, const float* prior,
                                    const float* pn_fwd, const float* pn_inv, const float* ins, void* out_hi,
                                    void* out_lo, int64_t ld_planes, int64_t out_col0, int64_t seg_pitch, int B,
                                    int N, int D, int I, int64_t F, int32_t* tile_counter, void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(rowptr_t && rowptr_h && prior && pn_fwd && pn_inv && ins && out_hi && out_lo, "null pointer");
  GR_CHECK_ARG(F == 0 || (src_t && rel_t && src_h && rel_h), "null edge arrays");
  GR_CHECK_ARG(B > 0 && N >= kRows && I > 0, "B, I must be positive and N >= 64");
  GR_CHECK_ARG(D == 200 && seg_pitch == 208, "this build specialises D = 200, seg_pitch = 208 (use gr_aggregate_dual)");
  GR_CHECK_ARG(ld_planes % 4 == 0 && out_col0 % 4 == 0 && ld_planes >= out_col0 + 2 * (int64_t)I * seg_pitch,
               "plane row pitch / column offset must be multiples of 4 and cover all segments");
  GR_CHECK_ARG((reinterpret_cast<uintptr_t>(out_hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(out_lo) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(pn_fwd) & 15) == 0 && (reinterpret_cast<uintptr_t>(pn_inv) & 15) == 0,
               "misaligned planes / padded tables");
  PnParams p{};
  p.dir[0] = PnDir{rowptr_t, src_t, rel_t, w_t, pn_fwd};
  p.dir[1] = PnDir{rowptr_h, src_h, rel_h, w_h, pn_inv};
  p.prior = prior; p.ins = ins;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(out_hi); p.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
  p.ld = ld_planes; p.out_col0 = out_col0; p.Nt = (int64_t)B * N;
  p.B = B; p.N = N; p.I = I; p.tile_counter = tile_counter;
  p.hot_rel = g_opt_agg_hot_rel;
  p.table_rows = g_opt_agg_table_rows;
  for (int j0 = 0; j0 < I; j0 += 4) {
    p.j0 = j0;
    const int ni = I - j0 < 4 ? I - j0 : 4;
    int rc = ni == 1 ? launch_pn<1>(p, stream) : ni == 2 ? launch_pn<2>(p, stream)
             : ni == 3 ? launch_pn<3>(p, stream) : launch_pn<4>(p, stream);
    if (rc != GR_OK) return rc;
  }
  return GR_OK;
}
