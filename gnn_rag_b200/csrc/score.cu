// Answer scoring + per-question masked softmax, and the seed-weighted row pick of QueryReform.
//
// gr_score_softmax: reference ReasonGNNLayer.forward tail (gnn/modules/kg_reasoning/reasongnn.py:165-169)
// and NSMBaseLayer.forward (nsm_gnn.py:67-74):  score = score_func(h) + (1-mask)*VERY_NEG_NUMBER,
// dist = Softmax(dim=1)(score).
// gr_seed_retrieve: torch.bmm(seed_info.unsqueeze(1), ent_emb) in QueryReform.forward
// (gnn/modules/query_update.py:40).
#include "common.cuh"

namespace gr {
namespace {

constexpr float kVeryNeg = -100000000000.0f;   // VERY_NEG_NUMBER, reasongnn.py:9

// one warp per node row: logits[n] = dot(h[n,:], w) + b + (1-mask[n]) * VERY_NEG
__global__ void score_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ w,
                             const float* __restrict__ bptr, const float* __restrict__ mask,
                             float* __restrict__ logits, int64_t Nt, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float bias = bptr ? bptr[0] : 0.f;
  const bool vec = (D % 4 == 0) && (ldh % 4 == 0) && (reinterpret_cast<size_t>(h) % 16 == 0) &&
                   (reinterpret_cast<size_t>(w) % 16 == 0);
  for (int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < Nt; n += warps) {
    const float* row = h + n * ldh;
    float s = 0.f;
    if (vec) {
      for (int d = lane * 4; d < D; d += 128) {
        float4 a = __ldg(reinterpret_cast<const float4*>(row + d));
        float4 b = __ldg(reinterpret_cast<const float4*>(w + d));
        s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
      }
    } else {
      for (int d = lane; d < D; d += 32) s = fmaf(__ldg(row + d), __ldg(w + d), s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      float m = mask[n];
      logits[n] = (s + bias) + (1.0f - m) * kVeryNeg;   // reasongnn.py:168
    }
  }
}

__device__ __forceinline__ float block_reduce(float v, float* sm, bool is_max) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float y = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, y) : v + y;
  }
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float x = lane < nw ? sm[lane] : (is_max ? -INFINITY : 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float y = __shfl_xor_sync(0xffffffffu, x, o);
      x = is_max ? fmaxf(x, y) : x + y;
    }
    if (lane == 0) sm[0] = x;
  }
  __syncthreads();
  float r = sm[0];
  __syncthreads();
  return r;
}

// logits[n] = (dots[n] + b) + (1-mask[n]) * VERY_NEG   (in place allowed)
__global__ void logits_from_dots_kernel(const float* __restrict__ dots, const float* __restrict__ dots2,
                                        const float* __restrict__ bptr, const float* __restrict__ mask,
                                        float* __restrict__ logits, int64_t Nt) {
  const float bias = bptr ? bptr[0] : 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < Nt; n += stride) {
    float d = dots[n];
    if (dots2) d += dots2[n];
    logits[n] = (d + bias) + (1.0f - mask[n]) * kVeryNeg;
  }
}

// one CTA per question: dist[b,:] = softmax(logits[b,:])
__global__ void softmax_kernel(const float* __restrict__ logits, float* __restrict__ dist, int N) {
  __shared__ float sm[32];
  const float* x = logits + (int64_t)blockIdx.x * N;
  float* y = dist + (int64_t)blockIdx.x * N;
  float mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, x[n]);
  mx = block_reduce(mx, sm, true);
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(x[n] - mx);
  s = block_reduce(s, sm, false);
  for (int n = threadIdx.x; n < N; n += blockDim.x) y[n] = expf(x[n] - mx) / s;
}

// one CTA per question: dist[b,:] = softmax((dots + dots2 + bias) + (1-mask)*VERY_NEG); the logits are staged in
// dist itself (each thread re-reads only what it wrote)
__global__ void masked_softmax_kernel(const float* __restrict__ dots, const float* __restrict__ dots2,
                                      const float* __restrict__ bptr, const float* __restrict__ mask,
                                      float* __restrict__ dist, int N) {
  __shared__ float sm[32];
  const int64_t off = (int64_t)blockIdx.x * N;
  const float bias = bptr ? bptr[0] : 0.f;
  float* y = dist + off;
  float mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float d = dots[off + n];
    if (dots2) d += dots2[off + n];
    const float l = (d + bias) + (1.0f - mask[off + n]) * kVeryNeg;
    y[n] = l;
    mx = fmaxf(mx, l);
  }
  mx = block_reduce(mx, sm, true);
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(y[n] - mx);
  s = block_reduce(s, sm, false);
  for (int n = threadIdx.x; n < N; n += blockDim.x) y[n] = expf(y[n] - mx) / s;
}

// one CTA per question; seeds are visited in local-index order (deterministic)
__global__ void seed_retrieve_kernel(const float* __restrict__ seed, const float* __restrict__ h,
                                     int64_t ldh, float* __restrict__ out, int N, int D) {
  __shared__ int s_list[1024];
  __shared__ float s_val[1024];
  __shared__ int s_woff[33];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const float* sd = seed + (int64_t)b * N;
  // accumulators: thread t owns columns t, t+blockDim, ...
  constexpr int kMaxCols = 4;
  float acc[kMaxCols] = {0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < N; base += blockDim.x) {
    int n = base + tid;
    float v = n < N ? sd[n] : 0.f;
    bool nzf = v != 0.f;
    unsigned bal = __ballot_sync(0xffffffffu, nzf);
    if (lane == 0) s_woff[wid + 1] = __popc(bal);
    __syncthreads();
    if (tid == 0) {
      s_woff[0] = 0;
      for (int i = 0; i < nw; ++i) s_woff[i + 1] += s_woff[i];
    }
    __syncthreads();
    if (nzf) {
      int pos = s_woff[wid] + __popc(bal & ((1u << lane) - 1));
      s_list[pos] = n;
      s_val[pos] = v;
    }
    __syncthreads();
    int cnt = s_woff[nw];
    for (int i = 0; i < cnt; ++i) {
      const float* row = h + ((int64_t)b * N + s_list[i]) * ldh;
      float sv = s_val[i];
#pragma unroll
      for (int c = 0; c < kMaxCols; ++c) {
        int d = tid + c * blockDim.x;
        if (d < D) acc[c] = fmaf(sv, row[d], acc[c]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c) {
    int d = tid + c * blockDim.x;
    if (d < D) out[(int64_t)b * D + d] = acc[c];
  }
}

}  // namespace
}  // namespace gr

extern "C" int gr_score_softmax(const float* h, int64_t ldh, const float* w_score,
                                const float* b_score, const float* mask, float* dist,
                                float* logits_out, int B, int N, int D, void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(h && w_score && mask && dist, "null pointer");
  GR_CHECK_ARG(B > 0 && N > 0 && D > 0 && ldh >= D, "bad shape");
  int64_t Nt = (int64_t)B * N;
  // logits are staged in logits_out if given, else in dist itself (softmax is then in place)
  float* logits = logits_out ? logits_out : dist;
  int grid = (int)std::min<int64_t>(ceil_div(Nt, 8), 16LL * sm_count());
  score_kernel<<<grid, 256, 0, stream>>>(h, ldh, w_score, b_score, mask, logits, Nt, D);
  GR_CHECK_LAUNCH();
  int threads = N >= 1024 ? 1024 : (N >= 256 ? 256 : 64);
  softmax_kernel<<<B, threads, 0, stream>>>(logits, dist, N);
  GR_CHECK_LAUNCH();
  return GR_OK;
}

extern "C" int gr_seed_retrieve(const float* seed_info, const float* h, int64_t ldh, float* out, int B,
                                int N, int D, void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(seed_info && h && out, "null pointer");
  GR_CHECK_ARG(B > 0 && N > 0 && D > 0 && ldh >= D, "bad shape");
  GR_CHECK_ARG(D <= 4 * 256, "D > 1024 unsupported");
  seed_retrieve_kernel<<<B, 256, 0, stream>>>(seed_info, h, ldh, out, N, D);
  GR_CHECK_LAUNCH();
  return GR_OK;
}

extern "C" int gr_masked_softmax(const float* dots, const float* dots2, const float* b_score,
                                 const float* mask, float* dist, int B, int N, void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(dots && mask && dist, "null pointer");
  GR_CHECK_ARG(B > 0 && N > 0, "bad shape");
  int threads = N >= 1024 ? 1024 : (N >= 256 ? 256 : 64);
  masked_softmax_kernel<<<B, threads, 0, stream>>>(dots, dots2, b_score, mask, dist, N);
  GR_CHECK_LAUNCH();
  return GR_OK;
}
