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
//   * relu(+-ins)/2 of the tile's (at most two) questions live in shared memory instead of 32 registers per
//     thread, which is what lets 3 CTAs (24 warps) fit per SM.
//   * output: the split-bf16 planes of the e2e GEMM's A operand, segment pitch SEGP (32-byte sectors, see
//     aggregate.cu); columns D..SEGP-1 receive exact zeros (staged ins are zero there).
#include <cuda_bf16.h>

#include "common.cuh"

namespace gr {

int g_opt_agg_abs_eb = 2;     // gr_set_option("agg_abs_eb", 2|4): gathered edges per branch-free block
int g_opt_agg_abs_minb = 3;   // gr_set_option("agg_abs_minb", 2|3): CTAs per SM the kernel is compiled for

namespace {

constexpr int kRows = 64;         // destination rows per CTA tile
constexpr int kThreads = 256;
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
};

__device__ __forceinline__ float4 ldg4(const char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void fma4(float4& acc, float c, const float4& v) {
  const float2 cc = make_float2(c, c);
  const float2 lo = __ffma2_rn(cc, make_float2(v.x, v.y), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(cc, make_float2(v.z, v.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ void fma4_abs(float4& acc, float c, const float4& v) {   // acc += c * |v|
  const float2 cc = make_float2(c, c);
  const float2 lo = __ffma2_rn(cc, make_float2(fabsf(v.x), fabsf(v.y)), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(cc, make_float2(fabsf(v.z), fabsf(v.w)), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
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
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(y01.x - f01.x, y01.y - f01.y);
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(y23.x - f23.x, y23.y - f23.y);
  if (pred) {
    *reinterpret_cast<uint2*>(ph) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    *reinterpret_cast<uint2*>(pl) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  }
}

template <int NI, int DT, int SEGP, int MINB, int EB>   // EB: gathered edges per branch-free block (2 or 4)
__global__ void __launch_bounds__(kThreads, MINB) agg_abs_kernel(const PnParams p) {
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
  for (int i = tid; i < 2 * NI * kPnCols; i += kThreads) {
    const int col = i % kPnCols, j = (i / kPnCols) % NI, q = i / (kPnCols * NI);
    const int b = b0 + q;
    const float x = (col < DT && b < p.B) ? p.ins[((int64_t)b * p.I + p.j0 + j) * DT + col] : 0.f;
    s_x[q][j][0][col] = 0.5f * fmaxf(x, 0.f);
    s_x[q][j][1][col] = 0.5f * fmaxf(-x, 0.f);
  }
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
  const bool wr1 = 128 + lane * 4 < SEGP;                    // chunk 1 columns that exist in the segment
  const char* tb[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) tb[d] = reinterpret_cast<const char*>(p.dir[d].pn) + lane * 16;
  __nv_bfloat16* const hi_lane = p.out_hi + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
  __nv_bfloat16* const lo_lane = p.out_lo + r0 * p.ld + p.out_col0 + lane * 4 + (int64_t)p.j0 * 2 * SEGP;
  const int lr_switch = N - rem0;                            // first tile row of question b0 + 1 (N >= kRows)

  for (int lr = warp; lr < nrows; lr += kWarps) {
    const float* xq = &s_x[lr >= lr_switch ? 1 : 0][0][0][lane * 4];
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
      float4 S0 = make_float4(0.f, 0.f, 0.f, 0.f), S1 = S0, Q0 = S0, Q1 = S0;
      const int fast_end = min(end, kEdgeCap);
      // EB edges per block; slots past the row end re-read the last edge with c forced to 0
      for (int i = beg; i < fast_end; i += EB) {
        int2 m[EB];
        float4 v0[EB], v1[EB];
#pragma unroll
        for (int u = 0; u < EB; ++u) {
          m[u] = s_rc[d][min(i + u, fast_end - 1)];
          if (u > 0 && i + u >= fast_end) m[u].y = 0;
        }
#pragma unroll
        for (int u = 0; u < EB; ++u) {
          const char* a = tb[d] + (uint32_t)m[u].x;
          v0[u] = ldg4(a);
          v1[u] = ldg4(a + 512);
        }
#pragma unroll
        for (int u = 0; u < EB; ++u) {
          const float c = __int_as_float(m[u].y);
          fma4(S0, c, v0[u]); fma4_abs(Q0, c, v0[u]); fma4(S1, c, v1[u]); fma4_abs(Q1, c, v1[u]);
        }
      }
      for (int i = max(beg, kEdgeCap); i < end; ++i) {       // slow path: slice overflowed the staging buffer
        const int64_t e = (int64_t)ebase + i;
        const PnDir& dd = p.dir[d];
        const float w = dd.w ? dd.w[e] : 1.0f;
        const float c = w * (w * p.prior[dd.src[e]]);
        const char* a = tb[d] + (uint32_t)dd.rel[e] * (uint32_t)kPnRowBytes;
        const float4 v0 = ldg4(a), v1 = ldg4(a + 512);
        fma4(S0, c, v0); fma4_abs(Q0, c, v0); fma4(S1, c, v1); fma4_abs(Q1, c, v1);
      }
      const float4 A0 = make_float4(Q0.x + S0.x, Q0.y + S0.y, Q0.z + S0.z, Q0.w + S0.w);   // 2 * sum c*relu(v)
      const float4 B0 = make_float4(Q0.x - S0.x, Q0.y - S0.y, Q0.z - S0.z, Q0.w - S0.w);   // 2 * sum c*relu(-v)
      const float4 A1 = make_float4(Q1.x + S1.x, Q1.y + S1.y, Q1.z + S1.z, Q1.w + S1.w);
      const float4 B1 = make_float4(Q1.x - S1.x, Q1.y - S1.y, Q1.z - S1.z, Q1.w - S1.w);
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int seg = d * SEGP + j * 2 * SEGP;
        const float* xj = xq + j * 2 * kPnCols;
        const float4 xp0 = *reinterpret_cast<const float4*>(xj), xn0 = *reinterpret_cast<const float4*>(xj + kPnCols);
        emit4(hrow + seg, lrow + seg, true, xp0, xn0, A0, B0);
        const float4 xp1 = *reinterpret_cast<const float4*>(xj + 128),
                     xn1 = *reinterpret_cast<const float4*>(xj + kPnCols + 128);
        emit4(hrow + seg + 128, lrow + seg + 128, wr1, xp1, xn1, A1, B1);
      }
    }
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
  const bool m3 = g_opt_agg_abs_minb >= 3, e4 = g_opt_agg_abs_eb >= 4;
  if (m3 && !e4) agg_abs_kernel<NI, 200, 208, 3, 2><<<grid, kThreads, 0, stream>>>(p);
  else if (m3) agg_abs_kernel<NI, 200, 208, 3, 4><<<grid, kThreads, 0, stream>>>(p);
  else if (!e4) agg_abs_kernel<NI, 200, 208, 2, 2><<<grid, kThreads, 0, stream>>>(p);
  else agg_abs_kernel<NI, 200, 208, 2, 4><<<grid, kThreads, 0, stream>>>(p);
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
                                    int N, int D, int I, int64_t F, void* stream_) {
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
  p.B = B; p.N = N; p.I = I;
  for (int j0 = 0; j0 < I; j0 += 4) {
    p.j0 = j0;
    const int ni = I - j0 < 4 ? I - j0 : 4;
    int rc = ni == 1 ? launch_pn<1>(p, stream) : ni == 2 ? launch_pn<2>(p, stream)
             : ni == 3 ? launch_pn<3>(p, stream) : launch_pn<4>(p, stream);
    if (rc != GR_OK) return rc;
  }
  return GR_OK;
}
