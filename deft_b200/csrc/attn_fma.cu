// Stage 1, warp-FMA path: partial softmax of one (item, kv-head) per CTA.
//
// Computes what the reference's stage-1 kernels compute (DeFT/deft/layers/attention/
// tree_attention.py:860-976 Flatten kernel2, :170-293 Node kernel): for every query row of every
// group of the item, over the item's KV tokens,  S = q.k^T/sqrt(D)  masked by the per-token
// bitmask, partial_o = softmax(S) V and partial_lse = m + log l.  Unlike the reference, one CTA
// serves all G = H/HKV query heads that share the kv-head, so a KV tile is staged in shared memory
// once for G x (up to 32) rows, and a long item is walked in 64-token tiles with online softmax.
//
// This is the general path (any row count, any item length) and the one used for sparse tiles;
// dense tiles go to the tcgen05 path (attn_umma.cu).
#include "common.cuh"

namespace deft {
namespace {

constexpr int kTile = 64;     // tokens staged per step
constexpr int kThreads = 256; // 8 warps
constexpr int kWarps = kThreads / 32;
constexpr int kPad = 8;       // halves of row padding: 16-byte shift per token row, conflict-free LDS.128

__device__ __forceinline__ uint4 ldg_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void unpack8(const uint4& raw, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int D, int G>
struct Smem {
  __half k[kTile][D + kPad];
  __half v[kTile][D + kPad];
  __half q[kMaxGroupQ * G][D];
  float p[kWarps][kTile][G];
  uint32_t mask[kTile];
  int64_t page[kTile];
};

template <int D, int G>
__global__ void __launch_bounds__(kThreads) stage1_fma_kernel(const AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<D, G>& sm = *reinterpret_cast<Smem<D, G>*>(smem_raw);

  const int n_items = p.n_items_dev ? *p.n_items_dev : p.n_items;
  const int item_id = blockIdx.x;
  if (item_id >= n_items) return;
  const int hkv = blockIdx.y;
  const deft_item_t item = p.items[item_id];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int DL = D >= 32 ? D / 32 : 1;   // output dims owned by a lane (head_dim 16: the upper lanes own none)
  constexpr int CH = D / 8;        // 16-byte chunks per row

  for (int gi = 0; gi < item.n_grp; ++gi) {
    const deft_group_t grp = p.groups[item.grp_off + gi];
    const int nrow = grp.q_cnt;
    __syncthreads();  // previous group's smem is no longer read
    // ---- stage Q rows of the group: [row][g][D] halves
    for (int c = tid; c < nrow * G * CH; c += kThreads) {
      const int r = c / (G * CH), rem = c % (G * CH), g = rem / CH, ch = rem % CH;
      const int64_t qid = p.q_list[grp.q_off + r];
      const __half* src = p.q + qid * p.q_row_stride + (int64_t)(hkv * G + g) * p.q_head_stride + ch * 8;
      *reinterpret_cast<uint4*>(&sm.q[r * G + g][ch * 8]) = ldg_nc_16(src);
    }
    // per-warp running state for the (up to 4) queries this warp owns: rows warp, warp+8, ...
    float m_run[4][G], l_run[4][G], acc[4][G][DL];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        m_run[a][g] = -INFINITY;
        l_run[a][g] = 0.f;
#pragma unroll
        for (int i = 0; i < DL; ++i) acc[a][g][i] = 0.f;
      }

    for (int t0 = 0; t0 < item.kv_len; t0 += kTile) {
      const int tlen = min(kTile, item.kv_len - t0);
      __syncthreads();  // previous tile fully consumed
      if (tid < kTile) {
        int64_t pg = 0;
        uint32_t mk = 0;
        if (tid < tlen) {
          const int64_t e = item.kv_off + t0 + tid;
          pg = p.kv_idx_bytes == 8 ? reinterpret_cast<const int64_t*>(p.kv_idx)[e]
                                   : (int64_t) reinterpret_cast<const int32_t*>(p.kv_idx)[e];
          mk = grp.mask_off >= 0 ? (uint32_t)p.masks[grp.mask_off + t0 + tid] : 0xffffffffu;
        }
        sm.page[tid] = pg;
        sm.mask[tid] = mk;
      }
      __syncthreads();
      // ---- stage K and V tiles (16-byte chunks, coalesced along D)
      for (int c = tid; c < kTile * CH; c += kThreads) {
        const int t = c / CH, ch = c % CH;
        uint4 kk = make_uint4(0, 0, 0, 0), vv = kk;
        if (t < tlen) {
          const int64_t base = sm.page[t] * p.kv_tok_stride + (int64_t)hkv * p.kv_head_stride + ch * 8;
          kk = ldg_nc_16(p.k + base);
          vv = ldg_nc_16(p.v + base);
        }
        *reinterpret_cast<uint4*>(&sm.k[t][ch * 8]) = kk;
        *reinterpret_cast<uint4*>(&sm.v[t][ch * 8]) = vv;
      }
      __syncthreads();

#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int r = warp + a * kWarps;
        if (r >= nrow) break;  // warp-uniform
        // ---- phase A: scores of tokens lane, lane+32 for the G heads of row r
        float s[G][2];
#pragma unroll
        for (int g = 0; g < G; ++g) s[g][0] = s[g][1] = 0.f;
#pragma unroll 4
        for (int ch = 0; ch < CH; ++ch) {
          float k0[8], k1[8];
          unpack8(*reinterpret_cast<const uint4*>(&sm.k[lane][ch * 8]), k0);
          unpack8(*reinterpret_cast<const uint4*>(&sm.k[lane + 32][ch * 8]), k1);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            float qf[8];
            unpack8(*reinterpret_cast<const uint4*>(&sm.q[r * G + g][ch * 8]), qf);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              s[g][0] = fmaf(qf[i], k0[i], s[g][0]);
              s[g][1] = fmaf(qf[i], k1[i], s[g][1]);
            }
          }
        }
        const bool ok0 = lane < tlen && ((sm.mask[lane] >> r) & 1u);
        const bool ok1 = lane + 32 < tlen && ((sm.mask[lane + 32] >> r) & 1u);
        float alpha[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float s0 = ok0 ? s[g][0] * p.scale : -INFINITY;
          const float s1 = ok1 ? s[g][1] * p.scale : -INFINITY;
          float tmax = fmaxf(s0, s1);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
          const float m_new = fmaxf(m_run[a][g], tmax);
          float p0 = 0.f, p1 = 0.f;
          alpha[g] = 1.f;
          if (m_new > -INFINITY) {  // warp-uniform
            p0 = ok0 ? __expf(s0 - m_new) : 0.f;
            p1 = ok1 ? __expf(s1 - m_new) : 0.f;
            alpha[g] = __expf(m_run[a][g] - m_new);  // exp(-inf) = 0 on the first live tile
          }
          float psum = p0 + p1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
          l_run[a][g] = l_run[a][g] * alpha[g] + psum;
          m_run[a][g] = m_new;
          sm.p[warp][lane][g] = p0;
          sm.p[warp][lane + 32][g] = p1;
        }
        __syncwarp();
        // ---- phase B: acc = acc*alpha + P V ; lane owns DL consecutive output dims
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int i = 0; i < DL; ++i) acc[a][g][i] *= alpha[g];
#pragma unroll 4
        for (int t = 0; t < tlen; ++t) {
          float pv[G];
#pragma unroll
          for (int g = 0; g < G; ++g) pv[g] = sm.p[warp][t][g];
          float vf[DL];
          if constexpr (DL == 4) {
            const uint2 raw = *reinterpret_cast<const uint2*>(&sm.v[t][lane * 4]);
            const __half2* h = reinterpret_cast<const __half2*>(&raw);
            float2 a0 = __half22float2(h[0]), a1 = __half22float2(h[1]);
            vf[0] = a0.x; vf[1] = a0.y; vf[2] = a1.x; vf[3] = a1.y;
          } else if constexpr (DL == 2) {
            float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&sm.v[t][lane * 2]));
            vf[0] = a0.x; vf[1] = a0.y;
          } else {
            vf[0] = lane < D ? __half2float(sm.v[t][lane]) : 0.f;
          }
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < DL; ++i) acc[a][g][i] = fmaf(pv[g], vf[i], acc[a][g][i]);
        }
        __syncwarp();  // sm.p[warp] is rewritten by the next row
      }
    }

    // ---- write the group's partials: po[row][h][d] = acc / l, plse[row][h] = m + log l
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = warp + a * kWarps;
      if (r >= nrow) break;
      const int64_t prow = (int64_t)grp.part_base + r;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int h = hkv * G + g;
        const float l = l_run[a][g];
        const float inv = l > 0.f ? 1.f / l : 0.f;
        float* dst = p.po + (prow * p.H + h) * D + lane * DL;
        if (lane * DL < D) {
#pragma unroll
          for (int i = 0; i < DL; ++i) dst[i] = acc[a][g][i] * inv;
        }
        if (lane == 0) p.plse[prow * p.H + h] = l > 0.f ? m_run[a][g] + __logf(l) : -INFINITY;
      }
    }
  }
}

template <int D, int G>
int launch_t(const AttnParams& p, cudaStream_t stream) {
  static PerDeviceOnce once;  // per device; benign race: the attribute call is idempotent
  const size_t smem = sizeof(Smem<D, G>);
  const int dev = current_device_index();
  if (once.slot[dev] == 0) {
    DEFT_CUDA(cudaFuncSetAttribute(stage1_fma_kernel<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    once.slot[dev] = 1;
  }
  dim3 grid(p.n_items, p.HKV);
  stage1_fma_kernel<D, G><<<grid, kThreads, smem, stream>>>(p);
  DEFT_CUDA(cudaGetLastError());
  return DEFT_OK;
}

template <int D>
int launch_d(const AttnParams& p, cudaStream_t stream) {
  switch (p.H / p.HKV) {
    case 1: return launch_t<D, 1>(p, stream);
    case 2: return launch_t<D, 2>(p, stream);
    case 4: return launch_t<D, 4>(p, stream);
    case 8: return launch_t<D, 8>(p, stream);
  }
  set_error("unsupported GQA group size H/HKV = %d (supported: 1, 2, 4, 8)", p.H / p.HKV);
  return DEFT_E_ARG;
}

}  // namespace

int launch_stage1_fma(const AttnParams& p, cudaStream_t stream) {
  if (p.n_items <= 0) return DEFT_OK;
  switch (p.D) {
    case 16: return launch_d<16>(p, stream);
    case 32: return launch_d<32>(p, stream);
    case 64: return launch_d<64>(p, stream);
    case 128: return launch_d<128>(p, stream);
  }
  set_error("unsupported head_dim %d (supported: 16, 32, 64, 128)", p.D);
  return DEFT_E_ARG;
}

}  // namespace deft
