// Microbenchmark: what HBM bandwidth does the stage-1 kernel's ACCESS PATTERN reach when nothing but the loads runs?
// The reference's pool is [page][K|V][kv-head][D] fp16: one kv-head's K row is 256 contiguous bytes, the next
// token's is 4096 bytes further.  Every CTA streams one kv-head's K and V tiles (128 tokens) with the kernel's own
// TMA boxes {64 elements, 1 head, 32 pages} into a ring of shared-memory stages and drops them.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a kv_stream.cu -o kv_stream && ./kv_stream
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

constexpr int kStages = 3, kTileBytes = 2 * 128 * 256;  // K + V of 128 tokens of one head

// mode 0: CTA c streams head c % 8 of token chunk c / 8 (all heads of a token range at the same time, as when one tree
// keeps every CTA on one chain); mode 1: every CTA streams a token range of its own (heads at unrelated times)
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap mk, const __grid_constant__ CUtensorMap mv,
                                                         int tiles_per_cta, int mode, int n_tokens) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[kStages];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int head = blockIdx.x % 8;
  const int chunk = mode == 0 ? blockIdx.x / 8 : blockIdx.x;
  const int n_chunks = mode == 0 ? (gridDim.x + 7) / 8 : gridDim.x;
  const int tok0 = (int)((long long)chunk * (n_tokens - tiles_per_cta * 128) / (n_chunks > 1 ? n_chunks - 1 : 1)) / 128 * 128;
  auto issue = [&](int t) {
    const int st = t % kStages;
    const uint32_t bar = smem_u32(&full[st]);
    mbar_expect(bar, kTileBytes);
    const uint32_t dst = base + st * kTileBytes;
    for (int kv = 0; kv < 2; ++kv)
      for (int r = 0; r < 4; ++r)
        for (int pn = 0; pn < 2; ++pn)
          tma3(dst + kv * 32768 + pn * 16384 + r * 32 * 128, kv == 0 ? &mk : &mv, bar, pn * 64, head, tok0 + t * 128 + r * 32);
  };
  for (int t = 0; t < kStages && t < tiles_per_cta; ++t) issue(t);
  for (int t = 0; t < tiles_per_cta; ++t) {
    mbar_wait(smem_u32(&full[t % kStages]), (t / kStages) & 1);
    if (t + kStages < tiles_per_cta) issue(t + kStages);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int n_tokens = 400000, HKV = 8, D = 128;
  __half* pool;
  cudaMalloc(&pool, (size_t)n_tokens * 2 * HKV * D * 2);
  cudaMemset(pool, 0, (size_t)n_tokens * 2 * HKV * D * 2);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)f;
  CUtensorMap mk, mv;
  const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)HKV, (cuuint64_t)n_tokens};
  const cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)2 * HKV * D * 2};
  const cuuint32_t box[3] = {64, 1, 32}, es[3] = {1, 1, 1};
  enc(&mk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, pool, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&mv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, pool + HKV * D, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = kStages * kTileBytes + 1024;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode)
    for (int tiles : {6, 48, 300}) {
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<<<148, 128, smem>>>(mk, mv, tiles, mode, n_tokens);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2)
          printf("mode %d (%s), %3d tiles per CTA: %.1f us, %.0f GB/s (%s)\n", mode, mode == 0 ? "8 heads of a token range together" : "heads at unrelated times",
                 tiles, ms * 1e3, 148.0 * tiles * kTileBytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
      }
    }
  return 0;
}
