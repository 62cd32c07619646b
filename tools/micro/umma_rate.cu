// Microbenchmark: issue rate of tcgen05.mma kind::f16 M=128 (cta_group::1) for the shapes the stage-1
// kernel uses.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int n, bool b_mn) {
  return (1u << 4) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// mode 0: SS N=128 K-major B; 1: SS N=256; 2: TS N=128 MN-major B; 3: SS N=128, fresh accumulator every 8 MMAs
template <int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint64_t a = desc_sw128(base, 16, 1024), b = desc_sw128(base + 32768, 16, 1024);
    const uint64_t bv = desc_sw128(base + 32768, 16384, 1024);
    constexpr uint32_t id = MODE == 1 ? idesc(256, false) : MODE == 2 ? idesc(128, true) : MODE == 4 ? idesc(64, false) : MODE == 5 ? idesc(32, false) : idesc(128, false);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t off = (uint64_t)(((ks >> 2) * 16384 + (ks & 3) * 32) >> 4);
        const uint32_t acc = (MODE == 3) ? (uint32_t)(ks > 0) : (uint32_t)(it > 0 || ks > 0);
        if (MODE == 2)
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(tmem + 256),
                       "r"(tmem + ks * 8), "l"(bv + (uint64_t)((ks * 2048) >> 4)), "r"(id), "r"(acc) : "memory");
        else
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem),
                       "l"(a + off), "l"(b + off), "r"(id), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    t1 = clock64();  // issue time
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int MODE>
void run(const char* name, int grid, int iters) {
  long long* d;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<MODE><<<grid, 128, 200 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s grid %3d: %s  issue %.1f cyc/MMA, complete %.1f cyc/MMA (%d MMAs)\n", name, grid, cudaGetErrorString(e),
         (double)h[0] / (iters * 8), (double)h[1] / (iters * 8), iters * 8);
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<0>("SS M128 N128 K16 (accumulate)", grid, 64);
    run<3>("SS M128 N128 K16 (new acc / 8)", grid, 64);
    run<1>("SS M128 N256 K16", grid, 64);
    run<2>("TS M128 N128 K16 (B MN-major)", grid, 64);
    run<4>("SS M128 N64 K16", grid, 64);
    run<5>("SS M128 N32 K16", grid, 64);
    run<4>("SS M128 N64 K16, 8 MMAs only", grid, 1);
    run<0>("SS M128 N128 K16, 8 MMAs only", grid, 1);
    run<2>("TS M128 N128 K16, 8 MMAs only", grid, 1);
  }
  return 0;
}
