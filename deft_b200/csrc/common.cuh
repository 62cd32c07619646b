// Shared declarations of the libdeft_b200 translation units.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "deft_b200.h"

namespace deft {

void set_error(const char* fmt, ...);

#define DEFT_CHECK_ARG(cond, ...)   \
  do {                              \
    if (!(cond)) {                  \
      deft::set_error(__VA_ARGS__); \
      return DEFT_E_ARG;            \
    }                               \
  } while (0)

#define DEFT_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      deft::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                      __LINE__);                                                     \
      return DEFT_E_CUDA;                                                            \
    }                                                                                \
  } while (0)

// Once-only launch configuration is PER DEVICE: cudaFuncSetAttribute applies to the device that is current, and a
// process may drive several GPUs.  slot[d] == 0: device d has not been configured yet (value = its SM count after).
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  int slot[kMaxDevices];  // zero-initialised (static storage); a racing second configuration is idempotent
};
inline int current_device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

static_assert(sizeof(deft_unit_t) == 80, "deft_unit_t is part of the ABI: 80 bytes");
static_assert(sizeof(deft_job_t) == 96, "deft_job_t is part of the ABI: 96 bytes");
constexpr int kMaxGroupQ = 32;     // queries per group (reference max_q_len / BLOCK_M, tree_cache.py:623)
constexpr int kNodeSplit = 256;    // tokens per item when long Node entries are split on the device

// Everything stage 1 / stage 2 need to know about one attention call.
struct AttnParams {
  // TMA tensor maps (tcgen05 path), valid when the flags below are set:
  //   tmap_k / tmap_v: [pool][HKV][D] fp16, box {64, 1, 32 pages}, 128B swizzle
  //   tmap_q:          [nq][H][D]     fp16, box {64, G, 32 queries}, 128B swizzle
  //   tmap_kg / tmap_vg: the same pool as a 2-D matrix of head rows [pool * row_ratio][D], box {64, 1}: the
  //                    tile::gather4 form (four arbitrary rows per instruction) for pages that are not consecutive;
  //                    row of (page, kv-head) = page * kv_row_ratio + kv-head, rows >= kv_rows read as zeros
  CUtensorMap tmap_k, tmap_v, tmap_q, tmap_kg, tmap_vg;
  CUtensorMap tmap_k16, tmap_v16, tmap_k8, tmap_v8;  // tmap_k / tmap_v with boxes of 16 and 8 pages (aligned runs inside a block)
  // fused KV append: the same six box maps and two gather maps over the step's new K / V rows ([nq][HKV][D] views of the
  // fused qkv output; "page" = query id), valid when new_k != nullptr
  CUtensorMap tmap_nk, tmap_nv, tmap_nk16, tmap_nv16, tmap_nk8, tmap_nv8, tmap_nkg, tmap_nvg;
  const __half* new_k;
  const __half* new_v;
  int64_t new_row_stride, new_head_stride;
  const int32_t* cache_loc;   // [nq] page of every query's new row (stage 2 writes the rows there)
  int32_t new_row_ratio, new_rows;
  int32_t tma_kv, tma_q, tma_gather;
  int32_t kv_row_ratio, kv_rows;
  const __half* q;
  const __half* k;
  const __half* v;
  __half* o;
  int64_t q_row_stride, q_head_stride;
  int64_t kv_tok_stride, kv_head_stride;
  int64_t o_row_stride, o_head_stride;
  int32_t nq, H, HKV, D;
  float scale;
  // index tables
  const void* kv_idx;     // int64 or int32 page ids
  int32_t kv_idx_bytes;   // 8 or 4
  const int64_t* q_list;  // query id per (group, row)
  const int64_t* masks;   // per-token bitmask (bit r = row r of the group attends), may be null
  // plan
  const deft_item_t* items;
  const deft_group_t* groups;
  const int32_t* csr_off;
  const int32_t* csr_rows;
  const int32_t* n_items_dev;  // device-resident item count (device-built plans), or null
  int32_t n_items;             // launch bound on the item count
  // partial softmax buffers: po [rows][H][D] fp32, plse [rows][H] fp32
  float* po;
  float* plse;
  float* dbg;  // debug dump of the tensor-core path (raw S and O of the first unit), normally null
  int* trace;  // per-CTA timeline of the tensor-core path ([cta][128] SM cycles), normally null
  // ---- unit plan (tcgen05 path): tables may be the reference's (int64) or the builder's (32-bit)
  const deft_unit_t* units;
  const int32_t* n_units_dev;  // device-resident unit count (device-built plans), or null
  int32_t n_units;             // launch bound on the unit count
  const void* u_kv;    int32_t u_kv_bytes;    // page id per token slot
  const int32_t* u_blk;                       // load descriptor per chunk of 8 token slots (native tables), or null
  const void* u_mask;  int32_t u_mask_bytes;  // per-token row bitmask, may be null
  const void* u_q;     int32_t u_q_bytes;     // query id per (slot, row)
  const int32_t* u_csr_off;
  const int32_t* u_csr_rows;
  const int32_t* job_off;  // per-CTA job lists (host-balanced), or null: CTA c runs jobs c, c+grid, ...
  const deft_job_t* jobs;  // job records (first job of CTA c at [c], see deft_job_t)
  int32_t n_ctas;          // CTAs the job lists cover (grid size), 0 when job_off is null
  // tile partials: po16 [slot tile][D/8][32*G rows][8] fp16, plse16 [slot tile][32*G] fp32,
  // slot tile = (part_base / 32) * HKV + kv_head
  __half* po16;
  float* plse16;
  int32_t pdl;         // launch the tcgen05-path kernels with programmatic stream serialization
  int32_t plan_fresh;  // the unit plan was derived on the device by the preceding kernel of this call
  int32_t clustered;   // launched as clusters of two CTAs (the plan's job lists pair slot-jobs that share K/V tiles)
  int32_t experiment;  // bit flags switching kernel variants for A/B measurements (deft_b200_set_experiment)
};

__device__ __forceinline__ int64_t load_index(const void* p, int bytes, int64_t i) {
  return bytes == 8 ? reinterpret_cast<const int64_t*>(p)[i] : (int64_t) reinterpret_cast<const int32_t*>(p)[i];
}

// stage 1 (warp-FMA path) and stage 2, attn_fma.cu / combine.cu
int launch_stage1_fma(const AttnParams& p, cudaStream_t stream);
int launch_stage2(const AttnParams& p, cudaStream_t stream);
// stage 1 (tcgen05 path) over the unit plan, attn_umma.cu; stage 2 over its tile partials, combine.cu
bool stage1_umma_supported(const AttnParams& p);
int launch_stage1_umma(const AttnParams& p, cudaStream_t stream);
int launch_stage2_tiles(const AttnParams& p, cudaStream_t stream);

// device-side plan derivation from the reference tables, plan.cu
struct PlanBuffers {
  deft_item_t* items;
  deft_group_t* groups;
  int32_t* csr_off;
  int32_t* csr_rows;
  int32_t* cursor;    // nq ints of scratch
  int32_t* counters;  // [0] = n_items, [1] = n_part_rows, [2] = n_units
  deft_unit_t* units; // one per (item, pair of groups)
};
int launch_plan_flatten(const int64_t* block_q_cnts, const int64_t* block_q_offset,
                        const int64_t* block_lens, const int64_t* block_kv, const int64_t* block_q,
                        int64_t n_blocks, int32_t block_len, int32_t nq, const PlanBuffers& pb,
                        cudaStream_t stream);
int launch_plan_node(const int64_t* kv_offset, const int64_t* kv_len, const int64_t* q_offset,
                     const int64_t* q_len, const int64_t* node_q, int64_t n_entries, int32_t split,
                     int32_t nq, const PlanBuffers& pb, cudaStream_t stream);

}  // namespace deft
