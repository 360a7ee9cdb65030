// Backward of the relation-typed aggregation (the kernel Trainer_KBQA.train_epoch needs,
// gnn/train_model.py:209-233).  Forward (gr_aggregate; ReasonGNNLayer.reason_layer / reason_layer_inv,
// gnn/modules/kg_reasoning/reasongnn.py:61-116; NSMLayer.reason_layer, nsm_gnn.py:87-112):
//
//     out[n, j, :] = sum_{e -> n} c_e * relu(P[r_e, :] * x_j[b(n), :]),     c_e = w_e^2 * p[s_e]
//
// Given G = dL/dout the three gradients are reductions over the same edge list with three different keys:
//
//     dP[r, :]    = sum_{e: r_e = r}   c_e * sum_j  G[n_e, j, :] * x_j[b, :] * m_{e,j}       (key: relation)
//     dx_j[b, :]  = sum_{e in b}       c_e *        G[n_e, j, :] * P[r_e, :] * m_{e,j}       (key: question)
//     dp[s]       = sum_{e: s_e = s} w_e^2 * sum_j <G[n_e, j, :], relu(P[r_e, :] * x_j[b, :])>   (key: source node)
//
// with m_{e,j} = [P[r_e] * x_j[b] > 0] elementwise.  One warp per destination row walks the row's in-edges in the
// destination CSR the forward uses (so G[n] and x_j[b] are loaded once per row and stay in registers); dx_j is
// accumulated in registers across all rows a warp handles inside one question and flushed with one atomicAdd per
// column on a question change; dP and dp go out through fp32 atomics (relation rows / source nodes are random).  The
// gradient buffers are ACCUMULATED into (caller zeroes them): the two directions and the T x K layer calls of one
// backward pass add up in place.  Atomic accumulation order is not deterministic -- as in the reference, whose
// torch.sparse.mm backward on CUDA is an atomic scatter as well.
#include <algorithm>

#include "common.cuh"

namespace gr {
namespace {

constexpr int kBwdThreads = 256;
constexpr int kCPL = 8;                 // columns per lane: D <= 256

struct BwdParams {
  const int32_t* rowptr;
  const int32_t* src;
  const int32_t* rel;
  const float* w;
  const float* prior;
  const float* table;     // [R1, D]
  const float* ins;       // [B, I, D]
  const float* gout;      // [Nt, ld]: G[n, j, d] at n * ld + col0 + j * seg + d
  int64_t ld, col0, seg;
  float* gtable;          // [R1, D]   +=
  float* gins;            // [B, I, D] +=
  float* gprior;          // [Nt]      +=
  int64_t Nt;
  int N, D, I;
};

template <int NI>
__global__ void __launch_bounds__(kBwdThreads) agg_bwd_kernel(const BwdParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // contiguous row ranges per warp: a warp stays inside one question for ~all of its rows
  const int64_t per = (p.Nt + nwarps - 1) / nwarps;
  const int64_t row_beg = warp * per, row_end = min(p.Nt, row_beg + per);
  const int D = p.D;
  float x[NI][kCPL], dx[NI][kCPL];
  int cur_b = -1;
  auto flush = [&]() {
    if (cur_b < 0) return;
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int k = 0; k < kCPL; ++k) {
        const int c = lane + 32 * k;
        if (c < D && dx[j][k] != 0.f) atomicAdd(p.gins + ((int64_t)cur_b * p.I + j) * D + c, dx[j][k]);
      }
  };
  for (int64_t n = row_beg; n < row_end; ++n) {
    const int b = (int)(n / p.N);
    if (b != cur_b) {
      flush();
      cur_b = b;
#pragma unroll
      for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int k = 0; k < kCPL; ++k) {
          const int c = lane + 32 * k;
          x[j][k] = c < D ? __ldg(p.ins + ((int64_t)b * p.I + j) * D + c) : 0.f;
          dx[j][k] = 0.f;
        }
    }
    const int beg = p.rowptr[n], end = p.rowptr[n + 1];
    if (beg == end) continue;
    float g[NI][kCPL];
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int k = 0; k < kCPL; ++k) {
        const int c = lane + 32 * k;
        g[j][k] = c < D ? __ldg(p.gout + n * p.ld + p.col0 + (int64_t)j * p.seg + c) : 0.f;
      }
    for (int e = beg; e < end; ++e) {
      const int s = p.src[e], r = p.rel[e];
      const float w = p.w ? p.w[e] : 1.f;
      const float w2 = w * w;
      const float c_e = w2 * p.prior[s];
      const float* prow = p.table + (int64_t)r * D;
      float* gprow = p.gtable + (int64_t)r * D;
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < kCPL; ++k) {
        const int c = lane + 32 * k;
        if (c < D) {
          const float pv = __ldg(prow + c);
          float dp_c = 0.f;
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const float pre = pv * x[j][k];
            if (pre > 0.f) {
              dp_c = fmaf(g[j][k], x[j][k], dp_c);
              dx[j][k] = fmaf(c_e * g[j][k], pv, dx[j][k]);
              dot = fmaf(g[j][k], pre, dot);
            }
          }
          if (c_e != 0.f && dp_c != 0.f) atomicAdd(gprow + c, c_e * dp_c);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (lane == 0 && dot != 0.f) atomicAdd(p.gprior + s, w2 * dot);
    }
  }
  flush();
}

}  // namespace
}  // namespace gr

extern "C" int gr_aggregate_backward(const int32_t* rowptr, const int32_t* src, const int32_t* rel, const float* w,
                                     const float* prior, const float* table, const float* ins, const float* grad_out,
                                     int64_t grad_row_stride, int64_t grad_col0, int64_t seg_stride, float* grad_table,
                                     float* grad_ins, float* grad_prior, int B, int N, int D, int I, int64_t F,
                                     void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(rowptr && prior && table && ins && grad_out && grad_table && grad_ins && grad_prior, "null pointer");
  GR_CHECK_ARG(F == 0 || (src && rel), "null edge arrays");
  GR_CHECK_ARG(B > 0 && N > 0 && D > 0 && D <= 32 * kCPL && I > 0 && I <= 4, "need 0 < D <= 256 and 0 < I <= 4");
  GR_CHECK_ARG(seg_stride >= D && grad_row_stride >= grad_col0 + (int64_t)(I - 1) * seg_stride + D,
               "grad_out row stride / segment stride smaller than the rows it must hold");
  if (F == 0) return GR_OK;
  BwdParams p{};
  p.rowptr = rowptr; p.src = src; p.rel = rel; p.w = w; p.prior = prior; p.table = table; p.ins = ins;
  p.gout = grad_out; p.ld = grad_row_stride; p.col0 = grad_col0; p.seg = seg_stride;
  p.gtable = grad_table; p.gins = grad_ins; p.gprior = grad_prior;
  p.Nt = (int64_t)B * N; p.N = N; p.D = D; p.I = I;
  const int grid = (int)std::min<int64_t>(ceil_div(p.Nt * 32, kBwdThreads), 16LL * sm_count());
  switch (I) {
    case 1: agg_bwd_kernel<1><<<grid, kBwdThreads, 0, stream>>>(p); break;
    case 2: agg_bwd_kernel<2><<<grid, kBwdThreads, 0, stream>>>(p); break;
    case 3: agg_bwd_kernel<3><<<grid, kBwdThreads, 0, stream>>>(p); break;
    default: agg_bwd_kernel<4><<<grid, kBwdThreads, 0, stream>>>(p); break;
  }
  GR_CHECK_LAUNCH();
  return GR_OK;
}
