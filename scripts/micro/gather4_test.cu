// Micro-test: semantics and throughput of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_test gather4_test.cu && ./gather4_test
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void g4_check(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, int ncols, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(4 * ncols * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(s32(smem)), "l"(&map), "r"(s32(&bar)), "r"(0), "r"(r0), "r"(r1),
        "r"(r2), "r"(r3)
        : "memory");
    uint32_t ok = 0;
    int spins = 0;
    while (!ok && spins < 1000000) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
      ++spins;
    }
    out[4 * 256] = ok ? 1.f : -1.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * ncols; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

// throughput: every warp gathers NG groups of 4 rows per stage, 2 stages
template <int NG>
__global__ void g4_bw(const __grid_constant__ CUtensorMap map, const int* __restrict__ rel, int nedges, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwb = blockDim.x >> 5;
  __shared__ __align__(8) unsigned long long bars[32][2];
  unsigned char* ring = smem + (size_t)warp * 2 * NG * 3200;
  if (lane == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[warp][b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int gw = blockIdx.x * nwb + warp, nw = gridDim.x * nwb;
  float4 acc0 = make_float4(0, 0, 0, 0), acc1 = acc0;
  const bool ld1 = 128 + lane * 4 < 200;
  constexpr int NE = NG * 4;
  auto issue = [&](int e, int stage) {
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[warp][stage])), "r"(NE * 800) : "memory");
    __syncwarp();
    if (lane < NG) {
      const int4 r = *reinterpret_cast<const int4*>(rel + e + 4 * lane);
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
          " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(s32(ring + (stage * NG + lane) * 3200)), "l"(&map),
          "r"(s32(&bars[warp][stage])), "r"(0), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w)
          : "memory");
    }
  };
  int e = gw * NE, it = 0;
  if (e + NE <= nedges) issue(e, 0);
  for (; e + NE <= nedges; e += nw * NE, ++it) {
    const int stage = it & 1;
    if (e + nw * NE + NE <= nedges) issue(e + nw * NE, stage ^ 1);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(s32(&bars[warp][stage])), "r"((it >> 1) & 1) : "memory");
#pragma unroll
    for (int u = 0; u < NE; ++u) {
      const unsigned char* a = ring + stage * NG * 3200 + u * 800 + lane * 16;
      const float4 v0 = *reinterpret_cast<const float4*>(a);
      const float4 v1 = ld1 ? *reinterpret_cast<const float4*>(a + 512) : make_float4(0, 0, 0, 0);
      acc0.x += v0.x; acc0.y += v0.y; acc0.z += v0.z; acc0.w += v0.w;
      acc1.x += v1.x; acc1.y += v1.y; acc1.z += v1.z; acc1.w += v1.w;
    }
  }
  if (acc0.x + acc0.y + acc0.z + acc0.w + acc1.x + acc1.y + acc1.z + acc1.w == 12345.678f) out[0] = 1.f;
}

int main() {
  const int R = 12214, COLS = 256, E = 1 << 20;
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  float* tab; float* out; int* rel;
  cudaMalloc(&tab, (size_t)R * COLS * 4); cudaMalloc(&out, (4 * 256 + 8) * 4); cudaMalloc(&rel, E * 4);
  std::vector<float> h((size_t)R * COLS);
  for (int r = 0; r < R; ++r) for (int c = 0; c < COLS; ++c) h[(size_t)r * COLS + c] = r * 1000.f + c;
  cudaMemcpy(tab, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  std::vector<int> hr(E); srand(1); for (int i = 0; i < E; ++i) hr[i] = rand() % R;
  cudaMemcpy(rel, hr.data(), E * 4, cudaMemcpyHostToDevice);
  for (int box_rows : {1}) {
    for (int ncols : {200, 256}) {
      CUtensorMap m;
      cuuint64_t dims[2] = {(cuuint64_t)COLS, (cuuint64_t)R};
      cuuint64_t strides[1] = {(cuuint64_t)COLS * 4};
      cuuint32_t box[2] = {(cuuint32_t)ncols, (cuuint32_t)box_rows};
      cuuint32_t estr[2] = {1, 1};
      CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("box {%d cols, %d rows}: encode rc=%d\n", ncols, box_rows, (int)rc);
      if (rc != CUDA_SUCCESS) continue;
      cudaMemset(out, 0, (4 * 256 + 8) * 4);
      g4_check<<<1, 128, 4 * 256 * 4 + 128>>>(m, 5, 17, 3, R + 7, ncols, out);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> o(4 * 256 + 8);
      cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
      printf("   run: %s  barrier=%g  row0: %g %g .. %g | row1: %g %g | row2: %g %g | row3(oob): %g %g\n", cudaGetErrorString(e),
             o[4 * 256], o[0], o[1], o[ncols - 1], o[ncols], o[ncols + 1], o[2 * ncols], o[2 * ncols + 1], o[3 * ncols], o[3 * ncols + 1]);
      if (e != cudaSuccess) { printf("   (sticky error: stop)\n"); return 0; }
    }
  }
  // throughput with the {200, 1} map (if it worked) -- try box rows 1 first, then 4
  for (int box_rows : {1}) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)COLS, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)COLS * 4};
    cuuint32_t box[2] = {200u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) continue;
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    auto time = [&](auto launch, const char* name) {
      for (int i = 0; i < 3; ++i) launch();
      float best = 1e9;
      for (int i = 0; i < 10; ++i) {
        cudaEventRecord(s); launch(); cudaEventRecord(e); cudaEventSynchronize(e);
        float ms; cudaEventElapsedTime(&ms, s, e); best = ms < best ? ms : best;
      }
      printf("box_rows=%d %-40s %8.1f us  %6.2f TB/s  (%s)\n", box_rows, name, best * 1e3, (double)E * 800 / (best * 1e-3) / 1e12,
             cudaGetErrorString(cudaGetLastError()));
    };
    {
      const size_t smem = (size_t)8 * 2 * 2 * 3200 + 128;
      cudaFuncSetAttribute(g4_bw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      time([&] { g4_bw<2><<<296, 256, smem>>>(m, rel, E, out); }, "gather4 NG=2 (8 rows/stage), 8 warps x 2 CTA");
    }
    {
      const size_t smem = (size_t)8 * 2 * 1 * 3200 + 128;
      cudaFuncSetAttribute(g4_bw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      time([&] { g4_bw<1><<<296, 256, smem>>>(m, rel, E, out); }, "gather4 NG=1 (4 rows/stage), 8 warps x 2 CTA");
    }
    if (cudaDeviceSynchronize() != cudaSuccess) break;
  }
  return 0;
}
