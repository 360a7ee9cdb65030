// Shortest-path node sets between seed entities and retrieved candidates.
//
// Reference: build_graph (llm/src/utils/graph_utils.py:10-21) makes an UNDIRECTED nx.Graph from the
// question's triples; get_truth_paths (:49-75) enumerates nx.all_shortest_paths(seed, answer) for every
// (seed, candidate) pair.  The set of nodes on those paths is {v : d(s,v) + d(v,t) = d(s,t)}.
// One CTA per question: level-synchronous BFS from every source and every target over the union of the
// tail-CSR and head-CSR (which together are the undirected adjacency), then a marking pass.
// Pure integer work: bit-exact.
#include "common.cuh"

namespace gr {
namespace {

constexpr int kThreads = 512;

__device__ void bfs_from(int root, int32_t* __restrict__ dist, int N, int64_t row0,
                         const int32_t* __restrict__ rp_t, const int32_t* __restrict__ src_t,
                         const int32_t* __restrict__ rp_h, const int32_t* __restrict__ src_h,
                         int* s_changed) {
  for (int v = threadIdx.x; v < N; v += blockDim.x) dist[v] = (v == root) ? 0 : -1;
  __syncthreads();
  for (int level = 0; level < N; ++level) {
    if (threadIdx.x == 0) *s_changed = 0;
    __syncthreads();
    for (int v = threadIdx.x; v < N; v += blockDim.x) {
      if (dist[v] != level) continue;
      const int64_t g = row0 + v;
      for (int e = rp_t[g]; e < rp_t[g + 1]; ++e) {
        int u = (int)(src_t[e] - row0);
        if (u >= 0 && u < N && dist[u] < 0) { dist[u] = level + 1; *s_changed = 1; }
      }
      for (int e = rp_h[g]; e < rp_h[g + 1]; ++e) {
        int u = (int)(src_h[e] - row0);
        if (u >= 0 && u < N && dist[u] < 0) { dist[u] = level + 1; *s_changed = 1; }
      }
    }
    __syncthreads();
    int ch = *s_changed;
    __syncthreads();
    if (!ch) break;
  }
}

__global__ void __launch_bounds__(kThreads)
paths_kernel(const int32_t* __restrict__ rp_t, const int32_t* __restrict__ src_t,
             const int32_t* __restrict__ rp_h, const int32_t* __restrict__ src_h,
             const int32_t* __restrict__ source_idx, const int32_t* __restrict__ source_cnt,
             int max_sources, const int32_t* __restrict__ target_idx,
             const int32_t* __restrict__ target_cnt, int max_targets, uint8_t* __restrict__ on_path,
             int32_t* __restrict__ pair_dist, int N, int32_t* __restrict__ ws) {
  __shared__ int s_changed;
  const int b = blockIdx.x;
  const int64_t row0 = (int64_t)b * N;
  const int ns = min(source_cnt[b], max_sources), nt = min(target_cnt[b], max_targets);
  int32_t* base = ws + (int64_t)b * (max_sources + max_targets) * N;
  for (int i = 0; i < ns; ++i)
    bfs_from(source_idx[(int64_t)b * max_sources + i], base + (int64_t)i * N, N, row0, rp_t, src_t,
             rp_h, src_h, &s_changed);
  for (int j = 0; j < nt; ++j)
    bfs_from(target_idx[(int64_t)b * max_targets + j], base + (int64_t)(max_sources + j) * N, N, row0,
             rp_t, src_t, rp_h, src_h, &s_changed);
  __syncthreads();
  for (int v = threadIdx.x; v < N; v += blockDim.x) on_path[row0 + v] = 0;
  for (int i = threadIdx.x; i < max_sources * max_targets; i += blockDim.x)
    pair_dist[(int64_t)b * max_sources * max_targets + i] = -1;
  __syncthreads();
  for (int i = 0; i < ns; ++i) {
    const int32_t* ds = base + (int64_t)i * N;
    for (int j = 0; j < nt; ++j) {
      const int32_t* dt = base + (int64_t)(max_sources + j) * N;
      const int t = target_idx[(int64_t)b * max_targets + j];
      const int dst = ds[t];
      if (threadIdx.x == 0) pair_dist[((int64_t)b * max_sources + i) * max_targets + j] = dst;
      if (dst < 0) continue;
      for (int v = threadIdx.x; v < N; v += blockDim.x)
        if (ds[v] >= 0 && dt[v] >= 0 && ds[v] + dt[v] == dst) on_path[row0 + v] = 1;
    }
  }
}

}  // namespace
}  // namespace gr

extern "C" size_t gr_paths_workspace_bytes(int B, int N, int max_sources, int max_targets) {
  if (B <= 0 || N <= 0 || max_sources < 0 || max_targets < 0) return 0;
  return (size_t)B * (size_t)(max_sources + max_targets) * (size_t)N * sizeof(int32_t) + 16;
}

extern "C" int gr_shortest_path_nodes(const int32_t* rowptr_t, const int32_t* src_t,
                                      const int32_t* rowptr_h, const int32_t* src_h,
                                      const int32_t* source_idx, const int32_t* source_cnt,
                                      int max_sources, const int32_t* target_idx,
                                      const int32_t* target_cnt, int max_targets, uint8_t* on_path,
                                      int32_t* pair_dist, int B, int N, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
  using namespace gr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GR_CHECK_ARG(rowptr_t && rowptr_h && source_idx && source_cnt && target_idx && target_cnt &&
                   on_path && pair_dist,
               "null pointer");
  GR_CHECK_ARG(B > 0 && N > 0 && max_sources > 0 && max_targets > 0, "bad shape");
  if (!workspace || workspace_bytes < gr_paths_workspace_bytes(B, N, max_sources, max_targets)) {
    set_error("gr_shortest_path_nodes: workspace too small");
    return GR_ERR_WORKSPACE;
  }
  paths_kernel<<<B, kThreads, 0, stream>>>(rowptr_t, src_t, rowptr_h, src_h, source_idx, source_cnt,
                                           max_sources, target_idx, target_cnt, max_targets, on_path,
                                           pair_dist, N, reinterpret_cast<int32_t*>(workspace));
  GR_CHECK_LAUNCH();
  return GR_OK;
}
