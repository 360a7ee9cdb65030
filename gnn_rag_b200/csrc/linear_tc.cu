// Tensor-core linear layer for sm_100a:  C = act(A W^T + bias)  with fp32 in / fp32 out and fp32-class
// accuracy, via the 3-product split-bf16 scheme on tcgen05:
//     x = hi + lo (hi = bf16(x), lo = bf16(x - hi));   A W^T ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T
// (dropped term A_lo W_lo^T <= 2^-18 relative).  All three products accumulate into one fp32 TMEM
// accumulator.  This is the `e2e_linear` of ReasonGNNLayer.forward / NSMBaseLayer.forward
// (reference gnn/modules/kg_reasoning/reasongnn.py:163, nsm_gnn.py:63): a genuine dense contraction,
// M = B*N node rows, K = (2*num_ins+1)*D, N = D.
//
// Kernel: one CTA per 128-row tile of A, full N (<= 256) per CTA.  Warp roles (canonical Blackwell
// GEMM): warp 0 = TMA producer (cp.async.bulk.tensor, 128B-swizzled K-major tiles of the four bf16
// planes), warp 1 = MMA issuer (single thread, tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x Npad x 16,
// 3 MMAs per K-step), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld 32x32b, bias + relu,
// stores).  smem ring of kStages {A_hi, A_lo, W_hi, W_lo} stages with full/empty mbarriers;
// tcgen05.commit releases a stage and finally signals the epilogue.
//
// Roofline: tensor-pipe work 3 * 2*M*N*K flop; HBM traffic ~ 2 planes * M*K*2 B = M*K*4 B (same as fp32 A).
#include "tcgen05.cuh"

namespace gr {

int g_opt_linear_tc = 0;   // gr_set_option("linear_tc", 0|1): route e2e linears through this kernel
int g_tc_cluster = 2;      // gr_set_option("tc_cluster", 1|2): CTAs per cluster sharing W tiles by TMA multicast
int g_tc_bk = 32;          // gr_set_option("tc_bk", 32|64): k-block width (64B / 128B swizzle)
int g_tc_tma_store = 1;    // gr_set_option("tc_tma_store", 0|1): staged TMA-store epilogue vs direct per-row stores

namespace {

using namespace tc;
// k-block width BK (bf16 elements) is a template parameter: 64 (128-byte swizzle rows, 2-3 smem stages)
// or 32 (64-byte swizzle rows, twice as many, finer stages -> more TMA requests in flight)
constexpr int kThreads = 384;   // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue (2 column halves)
constexpr int kEpiWarps = 8;

// ---------------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) bf16 planes
// ---------------------------------------------------------------------------------------------------
// W [N, nseg*seg] dense -> hi/lo planes [N, >= nseg*pitch] with segment s at columns [s*pitch, s*pitch+seg),
// zeros in the padding columns (matches the padded activation-plane layout)
__global__ void split_bf16_seg_kernel(const float* __restrict__ W, int64_t ldw, int64_t N, int64_t Kpad,
                                      int seg, int pitch, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo, int64_t ldo) {
  const int64_t total = N * Kpad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t n = i / Kpad, k = i - n * Kpad;
    const int sidx = (int)(k / pitch), c = (int)(k - (int64_t)sidx * pitch);
    float v = c < seg ? __ldg(W + n * ldw + (int64_t)sidx * seg + c) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[n * ldo + k] = h;
    lo[n * ldo + k] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

__global__ void split_bf16_kernel(const float* __restrict__ A, int64_t lda, int64_t M, int64_t K,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  int64_t ldo, int vec_ok, int vec_st) {
  const int64_t kq = (K + 3) / 4;
  const int64_t total = M * kq;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t m = i / kq, k = (i - m * kq) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const float* src = A + m * lda + k;
    if (vec_ok && k + 3 < K) {
      float4 t = __ldg(reinterpret_cast<const float4*>(src));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k + q < K) v[q] = __ldg(src + q);
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      h[q] = __float2bfloat16_rn(v[q]);
      l[q] = __float2bfloat16_rn(v[q] - __bfloat162float(h[q]));
    }
    __nv_bfloat16* ph = hi + m * ldo + k;
    __nv_bfloat16* pl = lo + m * ldo + k;
    if (vec_st && k + 3 < ldo) {   // 8-byte aligned plane bases and ldo % 4 == 0
      *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<uint2*>(h);
      *reinterpret_cast<uint2*>(pl) = *reinterpret_cast<uint2*>(l);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k + q < ldo) { ph[q] = h[q]; pl[q] = l[q]; }
    }
  }
}

struct TcParams {
  const float* bias;
  float* C;                 // fp32 output [M, N] (may be null)
  int64_t ldc;
  __nv_bfloat16* c_hi;      // optional bf16 hi/lo planes of the output (next layer's A operand)
  __nv_bfloat16* c_lo;
  int64_t ldc16;
  const float* w_score;     // optional: score_func dot product (reasongnn.py:165), as two partial sums:
  float* dots;              //   dots[half*M + m] = sum over this half's columns of out[m,n] * w_score[n]
  int M, N, K, n_pad, stages, num_tiles;
  uint32_t flags;
  int tma_store;            // 1: epilogue stages 128x16 chunks in smem and writes them with TMA stores
};

constexpr int kStageOutBytes = BM * 16 * 4 + 2 * BM * 16 * 2;   // per column-half: fp32 8 KB + hi 4 KB + lo 4 KB

constexpr int kAccStride = 256;   // TMEM columns per accumulator buffer (two buffers -> 512 columns)

// ---------------------------------------------------------------------------------------------------
// the GEMM kernel: persistent over 128-row tiles; TMEM accumulators double-buffered so the epilogue of
// tile i overlaps the TMA/MMA mainloop of tile i+1.
// ---------------------------------------------------------------------------------------------------
template <int CS, int BK>   // CS: CTAs per cluster sharing W tiles by multicast; BK: k-block width
__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c_hi,
                 const __grid_constant__ CUtensorMap map_c_lo, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages] x {A_hi 16K, A_lo 16K, W_hi n_pad*128, W_lo n_pad*128}, then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = BM * BK * 2;            // 16384
  const int w_bytes = p.n_pad * BK * 2;       // n_pad * 128
  // GR_LINEAR_BF16_SINGLE: bf16 activation storage -- one product A_hi W_hi, stages hold {A_hi, W_hi} only
  const bool single = (p.flags & GR_LINEAR_BF16_SINGLE) != 0;
  const int w_off = single ? a_bytes : 2 * a_bytes;          // W_hi tile inside a stage
  const int stage_bytes = single ? a_bytes + w_bytes : 2 * a_bytes + 2 * w_bytes;
  // epilogue staging (TMA-store source, must be 128-byte aligned): right after the 1024-aligned stages
  uint8_t* s_out = smem + (size_t)p.stages * stage_bytes;    // [2 halves] x {fp32 128x16, hi 128x16, lo 128x16}
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + 2 * kStageOutBytes);
  uint64_t* full_bar = bars;                       // [stages]
  uint64_t* empty_bar = bars + p.stages;           // [stages]
  uint64_t* tmem_full_bar = bars + 2 * p.stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [256] zero padded
  float* s_ws = s_bias + 256;                                // [256] zero padded
  for (int i = threadIdx.x; i < 256; i += kThreads) {
    s_bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
    s_ws[i] = (p.w_score && i < p.N) ? p.w_score[i] : 0.f;
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + BK - 1) / BK;
  const int crank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int ncluster = gridDim.x / CS, cid = blockIdx.x / CS;
  const int ngroups = (p.num_tiles + CS - 1) / CS;       // tile groups: CS consecutive 128-row tiles
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CS);                      // every CTA of the cluster must release the slot
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kEpiWarps);    // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(2u * kAccStride)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CS > 1) cluster_sync_all();                        // all barriers of the cluster are initialised
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t phase = 0;
      int s = 0;
      const int w_rows = p.n_pad / CS;                   // W rows this CTA fetches (and multicasts)
      const int w_slice = w_rows * BK * 2;               // bytes
      for (int g = cid; g < ngroups; g += ncluster) {
        const int m0 = (g * CS + crank) * BM;            // may lie beyond M for the last group: zero-filled
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], phase ^ 1);
          uint8_t* st = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
          tma_load_2d(st, &map_a_hi, &full_bar[s], kb * BK, m0);
          if (!single) tma_load_2d(st + a_bytes, &map_a_lo, &full_bar[s], kb * BK, m0);
          if (CS == 1) {
            tma_load_2d(st + w_off, &map_w_hi, &full_bar[s], kb * BK, 0);
            if (!single) tma_load_2d(st + w_off + w_bytes, &map_w_lo, &full_bar[s], kb * BK, 0);
          } else {
            tma_load_2d_mc(st + w_off + crank * w_slice, &map_w_hi, &full_bar[s], kb * BK,
                           crank * w_rows, kMask);
            if (!single)
              tma_load_2d_mc(st + w_off + w_bytes + crank * w_slice, &map_w_lo, &full_bar[s], kb * BK,
                             crank * w_rows, kMask);
          }
          if (++s == p.stages) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n_pad
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      uint32_t phase = 0;
      int s = 0, it = 0;
      for (int g = cid; g < ngroups; g += ncluster, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccStride);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t da_hi = make_smem_desc<BK>(sa), da_lo = make_smem_desc<BK>(sa + a_bytes);
          const uint64_t dw_hi = make_smem_desc<BK>(sa + w_off);
          const uint64_t dw_lo = make_smem_desc<BK>(sa + w_off + w_bytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);   // +32 B per K step inside the swizzle row
            umma_bf16(tmem_d, da_hi + adv, dw_hi + adv, idesc, (kb | k) ? 1u : 0u);
            if (!single) {
              umma_bf16(tmem_d, da_hi + adv, dw_lo + adv, idesc, 1u);
              umma_bf16(tmem_d, da_lo + adv, dw_hi + adv, idesc, 1u);
            }
          }
          // free this smem stage (in every CTA of the cluster: their producers multicast into it)
          if (CS == 1) umma_commit(&empty_bar[s]); else umma_commit_mc(&empty_bar[s], kMask);
          if (++s == p.stages) { s = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);           // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    // warp e = warp-4: TMEM lane quarter q = warp & 3 (hardware rule), column half = e >> 2
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int row_in_tile = q * 32 + lane;
    const bool relu = p.flags & GR_LINEAR_RELU;
    const int nchunks = p.n_pad / 16;
    const int ch_beg = half == 0 ? 0 : (nchunks + 1) / 2, ch_end = half == 0 ? (nchunks + 1) / 2 : nchunks;
    const bool vec_ok = p.C && (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    const bool vec16_ok = p.c_hi && (p.ldc16 % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.c_hi) & 7) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.c_lo) & 7) == 0);
    int it = 0;
    for (int g = cid; g < ngroups; g += ncluster, ++it) {
      const int tile = g * CS + crank;
      const int acc = it & 1;
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row = (int64_t)tile * BM + row_in_tile;
      const bool row_ok = row < p.M;
      float* crow = p.C ? p.C + row * p.ldc : nullptr;
      __nv_bfloat16* hrow = p.c_hi ? p.c_hi + row * p.ldc16 : nullptr;
      __nv_bfloat16* lrow = p.c_hi ? p.c_lo + row * p.ldc16 : nullptr;
      float dot = 0.f;
      const uint32_t taddr = tmem_base + (uint32_t)(acc * kAccStride) + ((uint32_t)(q * 32) << 16);
      if (p.tma_store) {
        // ---- staged epilogue: 128x16 chunk -> smem (dense rows) -> TMA store (clips rows >= M, cols >= N)
        uint8_t* stg = s_out + (size_t)half * kStageOutBytes;
        float* s_c = reinterpret_cast<float*>(stg) + row_in_tile * 16;
        uint32_t* s_h = reinterpret_cast<uint32_t*>(stg + BM * 16 * 4) + row_in_tile * 8;
        uint32_t* s_l = reinterpret_cast<uint32_t*>(stg + BM * 16 * 4 + BM * 16 * 2) + row_in_tile * 8;
        const bool issuer = (warp == 4 + 4 * half) && lane == 0;
        for (int ch = ch_beg; ch < ch_end; ++ch) {
          const int c0 = ch * 16;
          uint32_t r[16];
          tmem_ld16(taddr + (uint32_t)c0, r);
          if (ch == ch_end - 1) {        // last TMEM read of this tile: hand the accumulator back early
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j);
            const float4 w4 = *reinterpret_cast<const float4*>(s_ws + c0 + j);
            float x0 = __uint_as_float(r[j]) + b4.x, x1 = __uint_as_float(r[j + 1]) + b4.y;
            float x2 = __uint_as_float(r[j + 2]) + b4.z, x3 = __uint_as_float(r[j + 3]) + b4.w;
            if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
            dot = fmaf(x0, w4.x, dot); dot = fmaf(x1, w4.y, dot);
            dot = fmaf(x2, w4.z, dot); dot = fmaf(x3, w4.w, dot);
            v[j] = x0; v[j + 1] = x1; v[j + 2] = x2; v[j + 3] = x3;
          }
          uint32_t h[8], l[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j], v[j + 1]);
            const float2 hf = __bfloat1622float2(h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[j] - hf.x, v[j + 1] - hf.y);
            h[j / 2] = *reinterpret_cast<const uint32_t*>(&h2);
            l[j / 2] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          // the previous chunk's TMA stores must have finished READING the staging buffer
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          named_bar_sync(1 + half, 128);
          if (p.C) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(s_c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (p.c_hi) {
            *reinterpret_cast<uint4*>(s_h) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(s_h + 4) = make_uint4(h[4], h[5], h[6], h[7]);
            *reinterpret_cast<uint4*>(s_l) = make_uint4(l[0], l[1], l[2], l[3]);
            *reinterpret_cast<uint4*>(s_l + 4) = make_uint4(l[4], l[5], l[6], l[7]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          named_bar_sync(1 + half, 128);
          if (issuer) {
            const int m0 = tile * BM;
            if (p.C) tma_store_2d(&map_c, stg, c0, m0);
            if (p.c_hi) {
              tma_store_2d(&map_c_hi, stg + BM * 16 * 4, c0, m0);
              tma_store_2d(&map_c_lo, stg + BM * 16 * 4 + BM * 16 * 2, c0, m0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (p.dots && row_ok) p.dots[(int64_t)half * p.M + row] = dot;
        continue;
      }
      for (int ch = ch_beg; ch < ch_end; ++ch) {
        const int c0 = ch * 16;
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c0, r);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          // columns >= N: accumulator 0 (zero-filled W rows), bias 0, score weight 0 -> contribute 0
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j);
          const float4 w4 = *reinterpret_cast<const float4*>(s_ws + c0 + j);
          float x0 = __uint_as_float(r[j]) + b4.x, x1 = __uint_as_float(r[j + 1]) + b4.y;
          float x2 = __uint_as_float(r[j + 2]) + b4.z, x3 = __uint_as_float(r[j + 3]) + b4.w;
          if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
          dot = fmaf(x0, w4.x, dot); dot = fmaf(x1, w4.y, dot);
          dot = fmaf(x2, w4.z, dot); dot = fmaf(x3, w4.w, dot);
          v[j] = x0; v[j + 1] = x1; v[j + 2] = x2; v[j + 3] = x3;
        }
        if (row_ok) {
          const bool full = c0 + 16 <= p.N;
          if (crow) {
            if (vec_ok && full) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(crow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.N) crow[c0 + j] = v[j];
            }
          }
          if (hrow) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j], v[j + 1]);
              const float2 hf = __bfloat1622float2(h2);
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[j] - hf.x, v[j + 1] - hf.y);
              h[j / 2] = *reinterpret_cast<const uint32_t*>(&h2);
              l[j / 2] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            if (vec16_ok && full) {
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                *reinterpret_cast<uint2*>(hrow + c0 + 2 * j) = make_uint2(h[j], h[j + 1]);
                *reinterpret_cast<uint2*>(lrow + c0 + 2 * j) = make_uint2(l[j], l[j + 1]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.N) {
                  const uint32_t hw = h[j / 2], lw = l[j / 2];
                  reinterpret_cast<unsigned short*>(hrow)[c0 + j] = (unsigned short)((j & 1) ? (hw >> 16) : (hw & 0xFFFF));
                  reinterpret_cast<unsigned short*>(lrow)[c0 + j] = (unsigned short)((j & 1) ? (lw >> 16) : (lw & 0xFFFF));
                }
            }
          }
        }
      }
      if (p.dots && row_ok) p.dots[(int64_t)half * p.M + row] = dot;
      // release the accumulator to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (p.tma_store && lane == 0 && (warp == 4 || warp == 8))
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // outstanding TMA stores complete
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // nobody exits while a peer may still multicast into / arrive on this CTA
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(2u * kAccStride)
                 : "memory");
  }
}

struct TcPlan {
  int64_t kp;          // plane row stride (elements), multiple of 8
  int n_pad, stages, bk;
  size_t a_plane_bytes, w_plane_bytes, total_bytes, w_only_bytes, smem_bytes;
  bool ok;
};

TcPlan plan_tc(int64_t M, int64_t N, int64_t K, bool single = false) {
  TcPlan t{};
  const int BK = g_tc_bk == 64 ? 64 : 32;
  t.ok = (N >= 8 && N <= 256 && K >= 8 && M >= 1);
  t.kp = (K + 7) / 8 * 8;
  t.n_pad = (int)((N + 15) / 16 * 16);
  const size_t stage = (single ? 1 : 2) * ((size_t)BM * BK * 2 + (size_t)t.n_pad * BK * 2);
  // 227 KB usable smem minus alignment slack, bias/score arrays, barriers and the epilogue staging buffers
  int stages = (int)((227 * 1024 - 1024 - 2048 - 256 - 2 * kStageOutBytes) / stage);
  t.stages = stages > 8 ? 8 : stages;
  if (t.stages < 2) t.ok = false;
  t.bk = BK;
  t.smem_bytes = (size_t)t.stages * stage + 1024 /*align slack*/ + (2 * t.stages + 4) * 8 + 16 + 2 * 256 * 4 + 2 * kStageOutBytes;
  t.a_plane_bytes = align_up((size_t)M * t.kp * 2, 256);
  t.w_plane_bytes = align_up((size_t)N * t.kp * 2, 256);
  t.total_bytes = 2 * t.a_plane_bytes + 2 * t.w_plane_bytes;
  t.w_only_bytes = 2 * t.w_plane_bytes;
  return t;
}

int split_launch(const float* A, int64_t lda, int64_t M, int64_t K, __nv_bfloat16* hi, __nv_bfloat16* lo,
                 int64_t ldo, cudaStream_t stream) {
  int va = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  int64_t work = M * ((K + 3) / 4);
  int grid = (int)std::min<int64_t>(ceil_div(work, 256), 32LL * sm_count());
  int vs = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(hi) & 7) == 0) &&
           ((reinterpret_cast<uintptr_t>(lo) & 7) == 0);
  split_bf16_kernel<<<grid, 256, 0, stream>>>(A, lda, M, K, hi, lo, ldo, va, vs);
  GR_CHECK_LAUNCH();
  return GR_OK;
}

template <int CS, int BK>
int launch_tc_cs(const CUtensorMap& m_a_hi, const CUtensorMap& m_a_lo, const CUtensorMap& m_w_hi,
                 const CUtensorMap& m_w_lo, const CUtensorMap& m_c, const CUtensorMap& m_c_hi,
                 const CUtensorMap& m_c_lo, const TcPlan& t, const TcParams& p, cudaStream_t stream) {
  static bool attr_done[64] = {};
  if (first_use_on_device(attr_done)) {
    GR_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel<CS, BK>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int ngroups = (p.num_tiles + CS - 1) / CS;
  const int nclusters = std::max(1, std::min(ngroups, sm_count() / CS));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(nclusters * CS));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = t.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, linear_tc_kernel<CS, BK>, m_a_hi, m_a_lo, m_w_hi, m_w_lo, m_c, m_c_hi,
                                   m_c_lo, p));
  return GR_OK;
}

int launch_tc(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, int64_t lda16,
              const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, int64_t ldw16, const TcPlan& t,
              TcParams p, cudaStream_t stream) {
  p.n_pad = t.n_pad; p.stages = t.stages;
  p.num_tiles = (int)ceil_div(p.M, BM);
  // cluster multicast of W needs 8-row-aligned W slices and at least two tiles
  int cs = (g_tc_cluster >= 2 && (t.n_pad / 2) % 8 == 0 && p.num_tiles >= 2) ? 2 : 1;
  CUtensorMap m_a_hi, m_a_lo, m_w_hi, m_w_lo;
  if (!make_tmap(&m_a_hi, a_hi, p.M, p.K, lda16, BM, t.bk) || !make_tmap(&m_a_lo, a_lo, p.M, p.K, lda16, BM, t.bk) ||
      !make_tmap(&m_w_hi, w_hi, p.N, p.K, ldw16, t.n_pad / cs, t.bk) ||
      !make_tmap(&m_w_lo, w_lo, p.N, p.K, ldw16, t.n_pad / cs, t.bk)) {
    set_error("gr_linear_tc: cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned, row strides "
              "multiples of 8 elements)");
    return GR_ERR_CUDA;
  }
  // output tensor maps for the staged TMA-store epilogue (need 16-byte aligned bases and row pitches)
  CUtensorMap m_c, m_c_hi, m_c_lo;
  memset(&m_c, 0, sizeof(m_c)); memset(&m_c_hi, 0, sizeof(m_c_hi)); memset(&m_c_lo, 0, sizeof(m_c_lo));
  bool ok = g_tc_tma_store != 0;
  if (ok && p.C) ok = make_out_tmap(&m_c, p.C, p.M, p.N, p.ldc, 4);
  // the planes also receive the (exactly zero) columns N .. round16(N): whole 32-byte sectors per row
  const int64_t n16 = std::min<int64_t>((p.N + 15) / 16 * 16, p.ldc16);
  if (ok && p.c_hi) ok = make_out_tmap(&m_c_hi, p.c_hi, p.M, n16, p.ldc16, 2) &&
                         make_out_tmap(&m_c_lo, p.c_lo, p.M, n16, p.ldc16, 2);
  p.tma_store = ok ? 1 : 0;
  if (t.bk == 64) {
    if (cs == 2) return launch_tc_cs<2, 64>(m_a_hi, m_a_lo, m_w_hi, m_w_lo, m_c, m_c_hi, m_c_lo, t, p, stream);
    return launch_tc_cs<1, 64>(m_a_hi, m_a_lo, m_w_hi, m_w_lo, m_c, m_c_hi, m_c_lo, t, p, stream);
  }
  if (cs == 2) return launch_tc_cs<2, 32>(m_a_hi, m_a_lo, m_w_hi, m_w_lo, m_c, m_c_hi, m_c_lo, t, p, stream);
  return launch_tc_cs<1, 32>(m_a_hi, m_a_lo, m_w_hi, m_w_lo, m_c, m_c_hi, m_c_lo, t, p, stream);
}

}  // namespace

bool linear_tc_supported(int64_t M, int64_t N, int64_t K) {
  return plan_tc(M, N, K).ok && get_encode_fn() != nullptr;
}

}  // namespace gr

extern "C" size_t gr_linear_tc_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return gr::plan_tc(M, N, K).total_bytes;
}

extern "C" size_t gr_linear_tc_planes_workspace_bytes(int64_t N, int64_t K) {
  if (N <= 0 || K <= 0) return 0;
  return gr::plan_tc(1, N, K).w_only_bytes;
}

extern "C" int gr_split_bf16(const float* A, int64_t lda, int64_t M, int64_t K, void* hi, void* lo,
                             int64_t ld_out, void* stream_) {
  using namespace gr;
  GR_CHECK_ARG(A && hi && lo, "null pointer");
  GR_CHECK_ARG(M > 0 && K > 0 && lda >= K && ld_out >= K && ld_out % 8 == 0, "bad shape / ld_out % 8 != 0");
  return split_launch(A, lda, M, K, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo),
                      ld_out, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int gr_linear_tc(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                            float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, uint32_t flags,
                            void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(A && W && C && workspace, "null pointer");
  GR_CHECK_ARG(M > 0 && N > 0 && K > 0, "M, N, K must be positive");
  GR_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, "leading dimension smaller than row length");
  GR_CHECK_ARG(M < (int64_t)0x7fffffff - BM, "M exceeds int32 range");
  TcPlan t = plan_tc(M, N, K);
  if (!t.ok) {
    set_error("gr_linear_tc: unsupported shape M=%lld N=%lld K=%lld (need 8 <= N <= 256, K >= 8)",
              (long long)M, (long long)N, (long long)K);
    return GR_ERR_UNSUPPORTED;
  }
  if (workspace_bytes < t.total_bytes) {
    set_error("gr_linear_tc: workspace too small (%zu < %zu)", workspace_bytes, t.total_bytes);
    return GR_ERR_WORKSPACE;
  }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("gr_linear_tc: workspace must be 256-byte aligned");
    return GR_ERR_INVALID_ARG;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* a_lo = reinterpret_cast<__nv_bfloat16*>(ws + t.a_plane_bytes);
  __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(ws + 2 * t.a_plane_bytes);
  __nv_bfloat16* w_lo = reinterpret_cast<__nv_bfloat16*>(ws + 2 * t.a_plane_bytes + t.w_plane_bytes);
  int rc = split_launch(A, lda, M, K, a_hi, a_lo, t.kp, stream);
  if (rc != GR_OK) return rc;
  rc = split_launch(W, ldw, N, K, w_hi, w_lo, t.kp, stream);
  if (rc != GR_OK) return rc;
  TcParams p{};
  p.bias = bias; p.C = C; p.ldc = ldc;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.flags = flags;
  return launch_tc(a_hi, a_lo, t.kp, w_hi, w_lo, t.kp, t, p, stream);
}

extern "C" int gr_linear_tc_planes(const void* A_hi, const void* A_lo, int64_t lda16, const float* W,
                                   int64_t ldw, const float* bias, float* C, int64_t ldc, void* C_hi,
                                   void* C_lo, int64_t ldc16, const float* w_score, float* dots, int64_t M,
                                   int64_t N, int64_t K, int64_t k_seg, int64_t k_seg_pitch, uint32_t flags,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(A_hi && (A_lo || (flags & GR_LINEAR_BF16_SINGLE)) && W && workspace, "null pointer");
  GR_CHECK_ARG(C || C_hi, "no output requested");
  GR_CHECK_ARG(M > 0 && N > 0 && K > 0, "M, N, K must be positive");
  const bool segmented = k_seg > 0 && k_seg_pitch > k_seg;
  GR_CHECK_ARG(lda16 >= K && lda16 % 8 == 0, "lda16 must be >= K and a multiple of 8");
  GR_CHECK_ARG(!segmented || K % k_seg_pitch == 0, "K must be a multiple of k_seg_pitch");
  GR_CHECK_ARG(ldw >= (segmented ? K / k_seg_pitch * k_seg : K), "ldw smaller than the weight row length");
  GR_CHECK_ARG(!C || ldc >= N, "ldc smaller than N");
  GR_CHECK_ARG(!C_hi || (C_lo && ldc16 >= N), "C_lo missing or ldc16 smaller than N");
  GR_CHECK_ARG(!dots || w_score, "dots requested without w_score");
  GR_CHECK_ARG(M < (int64_t)0x7fffffff - BM, "M exceeds int32 range");
  TcPlan t = plan_tc(M, N, K, (flags & GR_LINEAR_BF16_SINGLE) != 0);
  if (!t.ok) {
    set_error("gr_linear_tc_planes: unsupported shape M=%lld N=%lld K=%lld", (long long)M, (long long)N,
              (long long)K);
    return GR_ERR_UNSUPPORTED;
  }
  if (workspace_bytes < t.w_only_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("gr_linear_tc_planes: workspace too small or not 256-byte aligned");
    return GR_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* w_lo = reinterpret_cast<__nv_bfloat16*>(ws + t.w_plane_bytes);
  int rc = GR_OK;
  if (flags & GR_LINEAR_W_PRESPLIT) {
    // the caller kept the workspace of an earlier call with the same W / N / K / k_seg / k_seg_pitch
  } else if (segmented) {
    int64_t work = N * K;
    int grid = (int)std::min<int64_t>(ceil_div(work, 256), 32LL * sm_count());
    split_bf16_seg_kernel<<<grid, 256, 0, stream>>>(W, ldw, N, K, (int)k_seg, (int)k_seg_pitch, w_hi, w_lo, t.kp);
    GR_CHECK_LAUNCH();
  } else {
    rc = split_launch(W, ldw, N, K, w_hi, w_lo, t.kp, stream);
  }
  if (rc != GR_OK) return rc;
  TcParams p{};
  p.bias = bias; p.C = C; p.ldc = ldc;
  p.c_hi = reinterpret_cast<__nv_bfloat16*>(C_hi); p.c_lo = reinterpret_cast<__nv_bfloat16*>(C_lo);
  p.ldc16 = ldc16; p.w_score = w_score; p.dots = dots;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.flags = flags;
  return launch_tc(reinterpret_cast<const __nv_bfloat16*>(A_hi),
                   reinterpret_cast<const __nv_bfloat16*>(A_lo ? A_lo : A_hi), lda16, w_hi, w_lo, t.kp, t, p, stream);
}
