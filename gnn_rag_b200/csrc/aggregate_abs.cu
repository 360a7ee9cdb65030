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
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace gr {

int g_opt_agg_abs_ws = 1;     // gr_set_option("agg_abs_ws", 0|1): persistent warp-specialised kernel when a tile counter is given

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
  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
  const float2 m1 = make_float2(-1.f, -1.f);
  const float2 r01 = __ffma2_rn(f01, m1, y01), r23 = __ffma2_rn(f23, m1, y23);   // y - hi, exact, packed
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(r01.x, r01.y);
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(r23.x, r23.y);
  if (pred) {
    *reinterpret_cast<uint2*>(ph) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
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
  const int fast_end = min(end, kEdgeCap);
  int i = beg;
  for (; i + 1 < fast_end; i += 2) {                       // two edges per step: 4 x 16-byte loads in flight per lane
    const int2 m0 = rc[i], m1 = rc[i + 1];
    const char* a0 = tb + (uint32_t)m0.x;
    const char* a1 = tb + (uint32_t)m1.x;
    const float4 v00 = ldg4(a0), v10 = ldg4(a1);
    const float4 v01 = ld1 ? ldg4(a0 + 512) : zero4(), v11 = ld1 ? ldg4(a1 + 512) : zero4();
    const float c0 = __int_as_float(m0.y), c1 = __int_as_float(m1.y);
    fma4(S0, c0, v00); fma4_abs(Q0, c0, v00); fma4(S1, c0, v01); fma4_abs(Q1, c0, v01);
    fma4(S0, c1, v10); fma4_abs(Q0, c1, v10); fma4(S1, c1, v11); fma4_abs(Q1, c1, v11);
  }
  if (i < fast_end) {                                      // odd tail: no padded slot
    const int2 m0 = rc[i];
    const char* a0 = tb + (uint32_t)m0.x;
    const float4 v00 = ldg4(a0);
    const float4 v01 = ld1 ? ldg4(a0 + 512) : zero4();
    const float c0 = __int_as_float(m0.y);
    fma4(S0, c0, v00); fma4_abs(Q0, c0, v00); fma4(S1, c0, v01); fma4_abs(Q1, c0, v01);
  }
  for (i = max(beg, kEdgeCap); i < end; ++i) {             // slow path: slice overflowed the staging buffer
    const int64_t e = (int64_t)ebase + i;
    const float w = dd.w ? dd.w[e] : 1.0f;
    const float c = w * (w * prior[dd.src[e]]);
    const char* a = tb + (uint32_t)dd.rel[e] * (uint32_t)kPnRowBytes;
    const float4 v0 = ldg4(a);
    const float4 v1 = ld1 ? ldg4(a + 512) : zero4();
    fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
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
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
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
        ne[d] = min(e1 - e0, kEdgeCap);
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
              bf.rc[d][i] = make_int2((int)((uint32_t)ridx[u] * (uint32_t)kPnRowBytes),
                                      __float_as_int(wv[u] * (wv[u] * pr[u])));
          }
        }
      }
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

template <int NI>
int launch_pn(const PnParams& p, cudaStream_t stream) {
  const unsigned grid = (unsigned)ceil_div(p.Nt, kRows);
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
  for (int j0 = 0; j0 < I; j0 += 4) {
    p.j0 = j0;
    const int ni = I - j0 < 4 ? I - j0 : 4;
    int rc = ni == 1 ? launch_pn<1>(p, stream) : ni == 2 ? launch_pn<2>(p, stream)
             : ni == 3 ? launch_pn<3>(p, stream) : launch_pn<4>(p, stream);
    if (rc != GR_OK) return rc;
  }
  return GR_OK;
}
