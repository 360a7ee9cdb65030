// Candidate ranking: the "retrieved answer-node set" of the evaluator.
//
// Reference: Evaluator.evaluate (gnn/evaluate.py:156,188-209) builds candidate2prob by dropping seeds
// (s == 1 after the LongTensor cast at :173), pads (c == len(id2entity)) and p < (1-eps)/N; f1_and_hits
// (:25-50) sorts with python's STABLE sorted(..., reverse=True) (equal probabilities keep local-index
// order) and keeps the prefix up to and including the item where the running float64 sum exceeds eps.
// Here: one CTA per question, order-preserving compaction, a 64-bit key sort
// (key = (~bits(p)) << 32 | local_index: ascending key == descending p, ascending index on ties), and the
// same sequential float64 running sum.  Integer/bit work end to end: bit-exact w.r.t. the reference
// given the same probabilities.
#include <math.h>

#include "common.cuh"

namespace gr {
namespace {

constexpr int kRankThreads = 512;
constexpr int kSmemKeys = 4096;

__device__ void bitonic_sort_u64(unsigned long long* a, int n) {
  // all-ascending bitonic network (virtual +inf padding): sorts arbitrary n
  for (int k = 2; (k >> 1) < n; k <<= 1) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int l = i ^ (k - 1);
      if (l > i && l < n) {
        unsigned long long x = a[i], y = a[l];
        if (x > y) { a[i] = y; a[l] = x; }
      }
    }
    __syncthreads();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int l = i ^ j;
        if (l > i && l < n) {
          unsigned long long x = a[i], y = a[l];
          if (x > y) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kRankThreads)
rank_kernel(const float* __restrict__ dist, const int64_t* __restrict__ local_entity,
            const float* __restrict__ query_entities, int64_t pad_id, double eps, double ignore_prob,
            int32_t* __restrict__ cand_idx, int32_t* __restrict__ cand_count,
            int32_t* __restrict__ cand_total, int N, unsigned long long* __restrict__ ws, int exact_ok) {
  __shared__ unsigned long long s_keys[kSmemKeys];
  __shared__ int s_woff[kRankThreads / 32 + 1];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int nw = kRankThreads / 32;
  const float* p = dist + (int64_t)b * N;
  const int64_t* le = local_entity + (int64_t)b * N;
  const float* qe = query_entities + (int64_t)b * N;
  unsigned long long* gkeys = ws + (int64_t)b * N;
  __shared__ int s_bad;      // a kept term > 1.0: not a probability -> no exactness argument, sequential sum
  if (tid == 0) { s_base = 0; s_bad = 0; }
  __syncthreads();
  // 1. order-preserving compaction of surviving candidates into gkeys
  for (int base = 0; base < N; base += kRankThreads) {
    int n = base + tid;
    bool keep = false;
    float pv = 0.f;
    if (n < N) {
      pv = p[n];
      bool is_seed = ((long long)qe[n]) == 1LL;          // evaluate.py:173,194
      bool is_pad = le[n] == pad_id;                      // :201
      bool small = (double)pv < ignore_prob;              // :203 (python float compare)
      keep = !is_seed && !is_pad && !small;
      if (keep && !(pv <= 1.0f)) s_bad = 1;
    }
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_woff[wid + 1] = __popc(bal);
    __syncthreads();
    if (tid == 0) {
      s_woff[0] = s_base;
      for (int i = 0; i < nw; ++i) s_woff[i + 1] += s_woff[i];
      s_base = s_woff[nw];
    }
    __syncthreads();
    if (keep) {
      int pos = s_woff[wid] + __popc(bal & ((1u << lane) - 1));
      unsigned hi = 0xFFFFFFFFu - __float_as_uint(pv);
      gkeys[pos] = ((unsigned long long)hi << 32) | (unsigned)n;
    }
    __syncthreads();
  }
  const int total = s_base;
  __syncthreads();
  // 2. sort
  unsigned long long* keys = gkeys;
  if (total <= kSmemKeys) {
    for (int i = tid; i < total; i += kRankThreads) s_keys[i] = gkeys[i];
    __syncthreads();
    keys = s_keys;
  } else {
    __threadfence_block();
  }
  bitonic_sort_u64(keys, total);
  // 3. eps-mass prefix = first i with (sum_{k<=i} p_k in float64) > eps (f1_and_hits, evaluate.py:41-50).
  //    The reference adds sequentially in python floats (fp64).  Every surviving p_k is an fp32 value
  //    >= ignore_prob, so with exact_ok (host: 24 + ceil(log2(1/ignore_prob)) + 1 <= 53) every partial sum of
  //    any subset is exactly representable in fp64: the fp64 sum is ORDER-INDEPENDENT and a parallel scan is
  //    bit-identical to the sequential loop.  Otherwise fall back to the sequential loop.
  if (exact_ok && !s_bad) {
    __shared__ double s_wsum[kRankThreads / 32];
    __shared__ double s_carry;
    __shared__ int s_first;
    if (tid == 0) { s_carry = 0.0; s_first = total; }
    __syncthreads();
    for (int base = 0; base < total && s_first == total; base += kRankThreads) {
      const int i = base + tid;
      double v = 0.0;
      if (i < total) v = (double)__uint_as_float(0xFFFFFFFFu - (unsigned)(keys[i] >> 32));
      double x = v;                                   // inclusive warp scan
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) s_wsum[wid] = x;
      __syncthreads();
      double off = s_carry;
      for (int w = 0; w < wid; ++w) off += s_wsum[w];
      const double cum = off + x;
      if (i < total && cum > eps) atomicMin(&s_first, i);
      __syncthreads();
      if (tid == kRankThreads - 1) s_carry = cum;     // carry = inclusive sum of the whole chunk
      __syncthreads();
    }
    if (tid == 0) {
      cand_count[b] = s_first < total ? s_first + 1 : total;
      cand_total[b] = total;
    }
  } else if (tid == 0) {
    double tp = 0.0;
    int cnt = 0;
    for (int i = 0; i < total; ++i) {
      unsigned hi = (unsigned)(keys[i] >> 32);
      float pv = __uint_as_float(0xFFFFFFFFu - hi);
      tp += (double)pv;
      cnt = i + 1;
      if (tp > eps) break;
    }
    cand_count[b] = cnt;
    cand_total[b] = total;
  }
  // 4. ordered local indices
  for (int i = tid; i < N; i += kRankThreads)   // slots past `total` are zero-filled (defined output)
    cand_idx[(int64_t)b * N + i] = i < total ? (int32_t)(keys[i] & 0xFFFFFFFFull) : 0;
}

}  // namespace
}  // namespace gr

extern "C" size_t gr_rank_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (size_t)B * (size_t)N * sizeof(unsigned long long);
}

extern "C" int gr_rank_candidates(const float* dist, const int64_t* local_entity,
                                  const float* query_entities, int64_t pad_id, double eps,
                                  int32_t* cand_idx, int32_t* cand_count, int32_t* cand_total, int B,
                                  int N, void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(dist && local_entity && query_entities && cand_idx && cand_count && cand_total,
               "null pointer");
  GR_CHECK_ARG(B > 0 && N > 0, "bad shape");
  if (workspace_bytes < gr_rank_workspace_bytes(B, N) || !workspace) {
    set_error("gr_rank_candidates: workspace too small");
    return GR_ERR_WORKSPACE;
  }
  double ignore_prob = (1 - eps) / N;   // evaluate.py:156
  // order-independence of the fp64 running sum (see rank_kernel step 3)
  int exact_ok = 0;
  if (ignore_prob > 0.0 && ignore_prob < 1.0) {
    int e_min;
    frexp(ignore_prob, &e_min);                       // ignore_prob = m * 2^e_min, m in [0.5, 1)
    // all terms are multiples of 2^(e_min - 24); terms are <= 1 (checked on device) and the scan only trusts
    // partial sums up to the first crossing of eps (< eps + 1 < 2): 24 + (1 - e_min) + 1 bits suffice
    int bits = 24 + (1 - e_min) + 1;
    exact_ok = bits <= 53 && eps < 1.0;
  }
  rank_kernel<<<B, kRankThreads, 0, stream>>>(dist, local_entity, query_entities, pad_id, eps,
                                              ignore_prob, cand_idx, cand_count, cand_total, N,
                                              reinterpret_cast<unsigned long long*>(workspace), exact_ok);
  GR_CHECK_LAUNCH();
  return GR_OK;
}
