// Device-side derivation of the work plan from the reference's index tables.
//
// When the caller hands in tables built by the reference's own Python builder
// (TreeMetadata.from_tree_cache, DeFT/deft/tree_decoding/tree_cache.py:618-881) no host-built plan
// exists, so one small single-CTA kernel derives it on the stream, without a host round trip:
//   * Flatten: consecutive table blocks that repeat the same 128 KV tokens (the reference duplicates
//     a KV block once per 32-query sub-block, tree_cache.py:680-708) are fused into ONE item with
//     several groups, so the KV tile is staged once for all of its queries;
//   * Node: entries longer than `split` tokens are cut into several items (the reference walks a
//     4096-token root entry in one serial loop, tree_attention.py:230-282);
//   * the query -> partial-row CSR that makes stage 2 deterministic (rows ascending per query).
// deft_b200_build_tables() produces the same arrays on the host for tables it builds itself.
#include "common.cuh"

namespace deft {
namespace {

constexpr int kThreads = 1024;

// exclusive block scan of one int per thread; returns the block total through `total`
__device__ int block_excl_scan(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // warp_sums reuse
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int s = warp_sums[lane];
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    warp_sums[lane] = si - s;          // exclusive warp offsets
    if (lane == 31) warp_sums[32] = si;  // total
  }
  __syncthreads();
  total = warp_sums[32];
  return warp_sums[warp] + incl - v;
}

// CSR of partial rows per query from the group list; rows ascending inside a query.
__device__ void build_csr(const deft_group_t* groups, int n_groups, const int64_t* q_list, int nq,
                          const PlanBuffers& pb, int* warp_sums) {
  const int tid = threadIdx.x;
  for (int q = tid; q <= nq; q += kThreads) pb.csr_off[q] = 0;
  __syncthreads();
  for (int g = tid; g < n_groups; g += kThreads) {
    const deft_group_t grp = groups[g];
    for (int r = 0; r < grp.q_cnt; ++r) {
      const int q = (int)q_list[grp.q_off + r];
      if (q >= 0 && q < nq) atomicAdd(&pb.csr_off[q], 1);
    }
  }
  __syncthreads();
  int carry = 0;
  for (int base = 0; base <= nq; base += kThreads) {
    const int q = base + tid;
    const int c = q < nq ? pb.csr_off[q] : 0;
    int total;
    const int ex = block_excl_scan(c, warp_sums, total);
    if (q <= nq) {
      pb.csr_off[q] = carry + ex;
      if (q < nq) pb.cursor[q] = carry + ex;
    }
    carry += total;
    __syncthreads();
  }
  if (tid == 0) pb.counters[1] = carry;
  __syncthreads();
  for (int g = tid; g < n_groups; g += kThreads) {
    const deft_group_t grp = groups[g];
    for (int r = 0; r < grp.q_cnt; ++r) {
      const int q = (int)q_list[grp.q_off + r];
      if (q >= 0 && q < nq) pb.csr_rows[atomicAdd(&pb.cursor[q], 1)] = grp.part_base + r;
    }
  }
  __syncthreads();
  // rank sort of every query's list by one warp (rows are distinct)
  const int lane = tid & 31, warp = tid >> 5;
  for (int q = warp; q < nq; q += kThreads / 32) {
    const int beg = pb.csr_off[q], n = pb.csr_off[q + 1] - beg;
    if (n <= 1) continue;
    if (n <= 32 * 8) {
      int val[8], rank[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = lane + 32 * k;
        val[k] = i < n ? pb.csr_rows[beg + i] : 0x7fffffff;
        rank[k] = 0;
      }
      for (int j = 0; j < n; ++j) {
        const int x = pb.csr_rows[beg + j];
#pragma unroll
        for (int k = 0; k < 8; ++k) rank[k] += x < val[k];
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (lane + 32 * k < n) pb.csr_rows[beg + rank[k]] = val[k];
    } else if (lane == 0) {  // very long lists: insertion sort by one lane
      for (int i = 1; i < n; ++i) {
        const int x = pb.csr_rows[beg + i];
        int j = i - 1;
        while (j >= 0 && pb.csr_rows[beg + j] > x) {
          pb.csr_rows[beg + j + 1] = pb.csr_rows[beg + j];
          --j;
        }
        pb.csr_rows[beg + j + 1] = x;
      }
    }
  }
}

// Units of the tcgen05 path: one per (item, pair of groups); the item's KV range becomes a chain of
// 128-token tiles, the pair's groups become the two slots.
__device__ void build_units(const deft_item_t* items, const deft_group_t* groups, int n_items, const PlanBuffers& pb,
                            int* warp_sums) {
  const int tid = threadIdx.x;
  int carry = 0;
  for (int base = 0; base < n_items; base += kThreads) {
    const int i = base + tid;
    deft_item_t it{};
    int np = 0;
    if (i < n_items) {
      it = items[i];
      np = (it.n_grp + 1) / 2;
    }
    int total;
    const int u0 = carry + block_excl_scan(np, warp_sums, total);
    for (int k = 0; k < np; ++k) {
      deft_unit_t u{};
      u.kv_off = it.kv_off;
      u.kv_tile_stride = 128;
      u.mask_tile_stride = 128;
      u.n_tiles = max(1, (it.kv_len + 127) / 128);
      u.last_len = max(0, it.kv_len - (u.n_tiles - 1) * 128);
      u.page0 = -1;
      u.q_id0[0] = u.q_id0[1] = -1;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int gi = 2 * k + s;
        if (gi < it.n_grp) {
          const deft_group_t g = groups[it.grp_off + gi];
          u.mask_off[s] = g.mask_off;
          u.q_off[s] = g.q_off;
          u.q_cnt[s] = g.q_cnt;
          u.part_base[s] = g.part_base;
        } else {
          u.mask_off[s] = -1;
        }
      }
      pb.units[u0 + k] = u;
    }
    carry += total;
    __syncthreads();
  }
  if (tid == 0) pb.counters[2] = carry;
}

__global__ void __launch_bounds__(kThreads) plan_flatten_kernel(
    const int64_t* __restrict__ block_q_cnts, const int64_t* __restrict__ block_q_offset,
    const int64_t* __restrict__ block_lens, const int64_t* __restrict__ block_kv,
    const int64_t* __restrict__ block_q, int n_blocks, int block_len, int nq, PlanBuffers pb) {
  __shared__ int warp_sums[33];
  const int tid = threadIdx.x;
  int carry = 0;
  for (int base = 0; base < n_blocks; base += kThreads) {
    const int b = base + tid;
    int head = 0;
    if (b < n_blocks)
      head = b == 0 || block_kv[(int64_t)b * block_len] != block_kv[(int64_t)(b - 1) * block_len];
    int total;
    const int idx = carry + block_excl_scan(head, warp_sums, total);
    if (b < n_blocks) {
      if (head) {
        deft_item_t it;
        it.kv_off = (int64_t)b * block_len;
        it.kv_len = (int)block_lens[b];
        it.grp_off = b;
        it.n_grp = 0;  // patched below
        it.cost = 0;
        pb.items[idx] = it;
      }
      deft_group_t g;
      g.mask_off = (int64_t)b * block_len;
      g.q_off = (int)block_q_offset[b];
      g.q_cnt = min((int)block_q_cnts[b], kMaxGroupQ);
      g.part_base = b * kMaxGroupQ;  // 32 partial rows per group: the tile-partial layout needs the alignment
      g.pad = 0;
      pb.groups[b] = g;
    }
    carry += total;
    __syncthreads();
  }
  const int n_items = carry;
  if (tid == 0) pb.counters[0] = n_items;
  __syncthreads();
  for (int i = tid; i < n_items; i += kThreads) {
    const int nxt = i + 1 < n_items ? pb.items[i + 1].grp_off : n_blocks;
    const int ng = nxt - pb.items[i].grp_off;
    pb.items[i].n_grp = ng;
    pb.items[i].cost = pb.items[i].kv_len * ng;
  }
  __syncthreads();
  build_units(pb.items, pb.groups, n_items, pb, warp_sums);
  __syncthreads();
  build_csr(pb.groups, n_blocks, block_q, nq, pb, warp_sums);
}

__global__ void __launch_bounds__(kThreads) plan_node_kernel(
    const int64_t* __restrict__ kv_offset, const int64_t* __restrict__ kv_len,
    const int64_t* __restrict__ q_offset, const int64_t* __restrict__ q_len,
    const int64_t* __restrict__ node_q, int n_entries, int split, int nq, PlanBuffers pb) {
  __shared__ int warp_sums[33];
  const int tid = threadIdx.x;
  int carry_items = 0;
  for (int base = 0; base < n_entries; base += kThreads) {
    const int e = base + tid;
    int nch = 0, qn = 0, kn = 0;
    if (e < n_entries) {
      kn = (int)kv_len[e];
      qn = min((int)q_len[e], kMaxGroupQ);
      nch = split > 0 ? max(1, (kn + split - 1) / split) : 1;
    }
    int tot_i;
    const int ib = carry_items + block_excl_scan(nch, warp_sums, tot_i);
    for (int j = 0; j < nch; ++j) {
      const int step = split > 0 ? split : kn;
      deft_item_t it;
      it.kv_off = kv_offset[e] + (int64_t)j * step;
      it.kv_len = max(0, min(step, kn - j * step));
      it.grp_off = ib + j;
      it.n_grp = 1;
      it.cost = it.kv_len;
      pb.items[ib + j] = it;
      deft_group_t g;
      g.mask_off = -1;
      g.q_off = (int)q_offset[e];
      g.q_cnt = qn;
      g.part_base = (ib + j) * kMaxGroupQ;
      g.pad = 0;
      pb.groups[ib + j] = g;
    }
    carry_items += tot_i;
    __syncthreads();
  }
  if (tid == 0) pb.counters[0] = carry_items;
  __syncthreads();
  build_units(pb.items, pb.groups, carry_items, pb, warp_sums);
  __syncthreads();
  build_csr(pb.groups, carry_items, node_q, nq, pb, warp_sums);
}

}  // namespace

int launch_plan_flatten(const int64_t* block_q_cnts, const int64_t* block_q_offset,
                        const int64_t* block_lens, const int64_t* block_kv, const int64_t* block_q,
                        int64_t n_blocks, int32_t block_len, int32_t nq, const PlanBuffers& pb,
                        cudaStream_t stream) {
  plan_flatten_kernel<<<1, kThreads, 0, stream>>>(block_q_cnts, block_q_offset, block_lens, block_kv,
                                                  block_q, (int)n_blocks, block_len, nq, pb);
  DEFT_CUDA(cudaGetLastError());
  return DEFT_OK;
}

int launch_plan_node(const int64_t* kv_offset, const int64_t* kv_len, const int64_t* q_offset,
                     const int64_t* q_len, const int64_t* node_q, int64_t n_entries, int32_t split,
                     int32_t nq, const PlanBuffers& pb, cudaStream_t stream) {
  plan_node_kernel<<<1, kThreads, 0, stream>>>(kv_offset, kv_len, q_offset, q_len, node_q,
                                               (int)n_entries, split, nq, pb);
  DEFT_CUDA(cudaGetLastError());
  return DEFT_OK;
}

}  // namespace deft
