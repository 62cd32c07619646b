// Warp-level merge of the tile partials of one (query, kv-head, chunk group): shared by the standalone
// stage-2 kernel (combine.cu) and the fused tail of the tcgen05 stage-1 kernel (attn_umma.cu).
//
// Tile partials: po16 [slot tile][D/8 chunks][32*G rows][8] fp16 and plse16 [slot tile][32*G] fp32,
// slot tile = (partial row / 32) * HKV + kv_head, row inside the tile = (partial row % 32) * G + g.
// lane = (g, chunk): the G heads of a query read 16*G contiguous bytes per chunk and write whole
// 128-byte lines of the output.  The partials of a query are merged in ascending partial-row order
// (CSR), single pass with a running maximum, all in fp32, one fp16 rounding at the end -- what the
// reference does with atomics (tree_attention.py:297-416), made deterministic.
#pragma once
#include "common.cuh"

namespace deft {

template <int D, int G>
struct CombineShape {
  static constexpr int CH = D / 8;                   // 16-byte chunks per head row
  static constexpr int CPW = 32 / G;                 // chunks one warp covers (lane = g + G * chunk)
  static constexpr int NCG = (CH + CPW - 1) / CPW;   // warps per (query, kv-head)
  static constexpr int R = kMaxGroupQ * G;           // rows per slot tile
};

// item = (q * HKV + kv_head) * NCG + chunk group.  Loads bypass L1 (__ldcg): in the fused kernel the
// partials were written by other SMs during the same launch.
// Programmatic dependent launch: everything above `griddep_wait` (index arithmetic, CSR loads -- plan data
// that no kernel of this call writes) overlaps the tail of the stage-1 kernel.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// The merge of one item in two steps, so that a caller can put a wait between them: prefetch() touches plan
// data only (what no kernel of the call writes: index arithmetic, the CSR bounds, the first 32 partial-row ids),
// run() reads the partials.
template <int D, int G>
struct TileMerge {
  using S = CombineShape<D, G>;
  int q, kvh, g, c, beg, end, first_row;
  bool active;
  __device__ __forceinline__ void prefetch(const AttnParams& p, int64_t item, int lane) {
    const int cg = (int)(item % S::NCG);
    kvh = (int)((item / S::NCG) % p.HKV);
    q = (int)(item / ((int64_t)S::NCG * p.HKV));
    g = lane % G;
    c = lane / G + cg * S::CPW;
    active = c < S::CH;
    beg = p.u_csr_off[q];
    end = p.u_csr_off[q + 1];
    first_row = beg + lane < end ? p.u_csr_rows[beg + lane] : -1;
  }
  template <int kBatch>
  __device__ __forceinline__ void run_batched(const AttnParams& p, int lane) const;
  __device__ __forceinline__ void run(const AttnParams& p, int lane) const { run_batched<16>(p, lane); }
};

template <int D, int G, bool kWaitDep = false, int kBatch = 16>
__device__ __forceinline__ void combine_tiles_item(const AttnParams& p, int64_t item, int lane) {
  TileMerge<D, G> tm;
  tm.prefetch(p, item, lane);
  if constexpr (kWaitDep) {
    tm.first_row = __shfl_sync(0xffffffffu, tm.first_row, lane);  // the row ids have landed before the wait returns
    griddep_wait();                                                // stage 1 has completed: its partials are visible
  }
  tm.template run_batched<kBatch>(p, lane);
}

template <int D, int G>
template <int kBatch>
__device__ __forceinline__ void TileMerge<D, G>::run_batched(const AttnParams& p, int lane) const {
  const uint4* po = reinterpret_cast<const uint4*>(p.po16);
  float m = -INFINITY, L = 0.f, acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  // Batches of 16 partials (a query of the BASELINE trees has 6-12): the row ids come in with one coalesced load
  // per 32, then the 16 log-sum-exps and the 16 data chunks are all in flight before the first one is consumed.
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int my_row = i0 == beg ? first_row : (i0 + lane < end ? p.u_csr_rows[i0 + lane] : -1);
    const int n_here = min(32, end - i0);
    for (int k0 = 0; k0 < n_here; k0 += kBatch) {
      float lse[kBatch];
      uint4 raw[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int row = __shfl_sync(0xffffffffu, my_row, (k0 + k) & 31);
        lse[k] = -INFINITY;
        raw[k] = make_uint4(0u, 0u, 0u, 0u);
        if (k0 + k < n_here) {
          const int64_t tile = (int64_t)(row >> 5) * p.HKV + kvh;
          const int rr = (row & 31) * G + g;
          lse[k] = __ldcg(p.plse16 + tile * S::R + rr);
          if (active) raw[k] = __ldcg(po + (tile * S::CH + c) * S::R + rr);
        }
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        if (lse[k] > -INFINITY) {
          const float m_new = fmaxf(m, lse[k]);
          const float a = __expf(m - m_new), wgt = __expf(lse[k] - m_new);  // a = 0 on the first live partial
          m = m_new;
          L = L * a + wgt;
          const __half2* h = reinterpret_cast<const __half2*>(&raw[k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x = __half22float2(h[j]);
            acc[2 * j] = fmaf(wgt, x.x, acc[2 * j] * a);
            acc[2 * j + 1] = fmaf(wgt, x.y, acc[2 * j + 1] * a);
          }
        }
      }
    }
  }
  if (!active) return;
  const float inv = L > 0.f ? 1.f / L : 0.f;
  uint4 pk;
  __half2 h0 = __floats2half2_rn(acc[0] * inv, acc[1] * inv), h1 = __floats2half2_rn(acc[2] * inv, acc[3] * inv);
  __half2 h2 = __floats2half2_rn(acc[4] * inv, acc[5] * inv), h3 = __floats2half2_rn(acc[6] * inv, acc[7] * inv);
  pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
  pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p.o + (int64_t)q * p.o_row_stride + (int64_t)(kvh * G + g) * p.o_head_stride + c * 8) = pk;
}

}  // namespace deft
