// Stage 2: tree-topology log-sum-exp combine of the stage-1 partials.
//
// Replaces DeFT_splitBynode_Triton_stage2 (DeFT/deft/layers/attention/tree_attention.py:297-416 with
// kernels :420-445 and :485-546).  The reference scatters with atomics (zero-initialised atomic max,
// fp32 atomic add of the weights, fp16 atomic add into the output, then o.div_(L)); here one warp
// owns one (query, head), walks the query's partial rows through a CSR in ascending row order and
// merges in fp32 with the true maximum, so the result is deterministic, needs no zeroed output and is
// rounded to fp16 exactly once.
#include "combine.cuh"

namespace deft {
namespace {

constexpr int kThreads = 256;

template <int D>
__global__ void __launch_bounds__(kThreads) stage2_kernel(const AttnParams p) {
  constexpr int DL = D >= 32 ? D / 32 : 1;   // head_dim 16: lanes 16-31 own no output dim
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (w >= (int64_t)p.nq * p.H) return;
  const int q = (int)(w / p.H), h = (int)(w % p.H);
  const int beg = p.csr_off[q], end = p.csr_off[q + 1];

  float m = -INFINITY;
  for (int i = beg + lane; i < end; i += 32) m = fmaxf(m, p.plse[(int64_t)p.csr_rows[i] * p.H + h]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));

  float L = 0.f, acc[DL];
#pragma unroll
  for (int i = 0; i < DL; ++i) acc[i] = 0.f;
  if (m > -INFINITY) {
    for (int i = beg; i < end; ++i) {
      const int64_t row = p.csr_rows[i];
      const float wgt = __expf(p.plse[row * p.H + h] - m);
      L += wgt;
      const float* src = p.po + (row * p.H + h) * D + lane * DL;
      if constexpr (DL == 4) {
        const float4 x = *reinterpret_cast<const float4*>(src);
        acc[0] = fmaf(wgt, x.x, acc[0]); acc[1] = fmaf(wgt, x.y, acc[1]);
        acc[2] = fmaf(wgt, x.z, acc[2]); acc[3] = fmaf(wgt, x.w, acc[3]);
      } else if constexpr (DL == 2) {
        const float2 x = *reinterpret_cast<const float2*>(src);
        acc[0] = fmaf(wgt, x.x, acc[0]); acc[1] = fmaf(wgt, x.y, acc[1]);
      } else {
        if (lane < D) acc[0] = fmaf(wgt, src[0], acc[0]);
      }
    }
  }
  const float inv = L > 0.f ? 1.f / L : 0.f;
  __half* dst = p.o + (int64_t)q * p.o_row_stride + (int64_t)h * p.o_head_stride + lane * DL;
  if constexpr (DL == 4) {
    __half2 a = __floats2half2_rn(acc[0] * inv, acc[1] * inv);
    __half2 b = __floats2half2_rn(acc[2] * inv, acc[3] * inv);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst) = pk;
  } else if constexpr (DL == 2) {
    *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(acc[0] * inv, acc[1] * inv);
  } else {
    if (lane < D) dst[0] = __float2half_rn(acc[0] * inv);
  }
}

// Standalone merge of the tile partials of the tcgen05 stage 1 (one warp per item, combine.cuh).  Used
// when the fused tail of the stage-1 kernel is disabled or cannot run (see attn_umma.cu).
// kBatch partials of a query are in flight at a time.  16 suits a single tree (a query of the BASELINE trees has 6-12
// partials, the kernel is one wave of latency-bound warps); a forest has thousands of queries with two or three
// partials each, and what binds it is how many warps fit an SM: 4 in flight need a third of the registers.
template <int D, int G, int kBatch>
__global__ void __launch_bounds__(kThreads) stage2_tiles_kernel(const AttnParams p) {
  griddep_launch_dependents();  // the next kernel of the stream may start its own prologue
  const int64_t w = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (w >= (int64_t)p.nq * p.HKV * CombineShape<D, G>::NCG) return;
  if (p.new_k != nullptr) {
    // Fused KV append (KVCacheUpdater.update, tree_cache.py:67-76): stage 1 has read this step's K / V rows straight
    // from the activations; here they go to their pages for the steps to come.  Warp (query, kv-head, chunk group)
    // copies its chunks of the row -- nothing of this call reads these pages, so no need to wait for anything.
    using S = CombineShape<D, G>;
    const int lane = threadIdx.x & 31;
    const int cg = (int)(w % S::NCG), kvh = (int)((w / S::NCG) % p.HKV), q = (int)(w / ((int64_t)S::NCG * p.HKV));
    const int c = cg * S::CPW + lane;
    if (lane < S::CPW && c < S::CH) {
      const int64_t src = (int64_t)q * p.new_row_stride + (int64_t)kvh * p.new_head_stride + c * 8;
      const int64_t dst = (int64_t)p.cache_loc[q] * p.kv_tok_stride + (int64_t)kvh * p.kv_head_stride + c * 8;
      *reinterpret_cast<uint4*>(const_cast<__half*>(p.k) + dst) = *reinterpret_cast<const uint4*>(p.new_k + src);
      *reinterpret_cast<uint4*>(const_cast<__half*>(p.v) + dst) = *reinterpret_cast<const uint4*>(p.new_v + src);
    }
  }
  combine_tiles_item<D, G, true, kBatch>(p, w, threadIdx.x & 31);
}

template <int D, int G>
int launch_tiles_t(const AttnParams& p, cudaStream_t stream) {
  const int64_t warps = (int64_t)p.nq * p.HKV * CombineShape<D, G>::NCG;
  const int64_t blocks = (warps + kThreads / 32 - 1) / (kThreads / 32);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(kThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // launch early, wait inside (griddep_wait)
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  if (warps >= 148 * 64 * 2) DEFT_CUDA(cudaLaunchKernelEx(&cfg, stage2_tiles_kernel<D, G, 4>, p));   // several waves of warps
  else DEFT_CUDA(cudaLaunchKernelEx(&cfg, stage2_tiles_kernel<D, G, 16>, p));
  return DEFT_OK;
}

template <int D>
int launch_t(const AttnParams& p, cudaStream_t stream) {
  const int64_t warps = (int64_t)p.nq * p.H;
  const int64_t blocks = (warps + kThreads / 32 - 1) / (kThreads / 32);
  stage2_kernel<D><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
  DEFT_CUDA(cudaGetLastError());
  return DEFT_OK;
}

}  // namespace

int launch_stage2(const AttnParams& p, cudaStream_t stream) {
  if (p.nq <= 0) return DEFT_OK;
  switch (p.D) {
    case 16: return launch_t<16>(p, stream);
    case 32: return launch_t<32>(p, stream);
    case 64: return launch_t<64>(p, stream);
    case 128: return launch_t<128>(p, stream);
  }
  set_error("unsupported head_dim %d", p.D);
  return DEFT_E_ARG;
}

int launch_stage2_tiles(const AttnParams& p, cudaStream_t stream) {
  if (p.nq <= 0) return DEFT_OK;
  const int G = p.H / p.HKV;
#define DEFT_CASE(DD, GG) \
  if (p.D == DD && G == GG) return launch_tiles_t<DD, GG>(p, stream);
  DEFT_CASE(128, 4) DEFT_CASE(128, 2) DEFT_CASE(128, 1) DEFT_CASE(64, 4) DEFT_CASE(64, 2) DEFT_CASE(64, 1)
#undef DEFT_CASE
  set_error("tile combine does not cover head_dim %d / GQA group %d", p.D, G);
  return DEFT_E_ARG;
}

}  // namespace deft
