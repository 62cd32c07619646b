// C ABI of libdeft_b200.so (see include/deft_b200.h for the contract of every entry point).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <unordered_map>

#include "common.cuh"

namespace deft {

static thread_local char g_error[512] = "";
static thread_local int g_stages = DEFT_STAGE_PLAN | DEFT_STAGE_1 | DEFT_STAGE_2;
static thread_local int g_stage1_impl = DEFT_STAGE1_AUTO;
static thread_local float* g_debug = nullptr;
static thread_local int* g_trace = nullptr;
static thread_local bool g_no_tma = false;
static thread_local bool g_no_pdl = false;
static thread_local bool g_no_gather4 = false;
static thread_local int g_experiment = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- TMA tensor maps.  cuTensorMapEncodeTiled comes from the driver through the runtime (no -lcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// fp16 tensor [d2][d1][d0] with element strides s1, s2; box {64, b1, b2}; 128-byte swizzle.
struct MapKey {
  const void* ptr; int64_t d0, d1, d2, s1, s2; int32_t b1, b2;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 && b1 == o.b1 && b2 == o.b2;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {(uint64_t)k.d0, (uint64_t)k.d1, (uint64_t)k.d2, (uint64_t)k.s1, (uint64_t)k.s2, (uint64_t)k.b1, (uint64_t)k.b2})
      h = (h ^ v) * 0x100000001B3ull + (h >> 29);
    return (size_t)h;
  }
};
// Per-thread cache: a decode step calls with the same layer pools over and over (5 maps per call: 160 keys for 32
// layers, 400 for 80), in cyclic order -- a hash map, emptied if a caller ever walks through thousands of buffers.
bool make_map(CUtensorMap* out, const MapKey& key) {
  constexpr size_t kMaxEntries = 4096;
  static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn fn = encode_tiled_fn();
  const bool two_d = key.d2 == 0 && key.b2 == 0;  // [d1][d0], box {64, b1}
  if (!fn || key.d0 < 64 || key.d1 <= 0 || (!two_d && key.d2 <= 0)) return false;
  alignas(64) CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)key.d0, (cuuint64_t)key.d1, (cuuint64_t)key.d2};
  const cuuint64_t strides[2] = {(cuuint64_t)key.s1 * 2, (cuuint64_t)key.s2 * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)key.b1, (cuuint32_t)key.b2};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, two_d ? 2 : 3, const_cast<void*>(key.ptr), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (cache.size() >= kMaxEntries) cache.clear();
  cache.emplace(key, map);
  *out = map;
  return true;
}

void setup_tma(AttnParams& p, int64_t kv_pool_tokens) {
  p.tma_kv = p.tma_q = p.tma_gather = 0;
  if (g_no_tma) return;
  const int G = p.H / p.HKV;
  if (kv_pool_tokens > 0 && p.kv_head_stride >= p.D && p.kv_tok_stride >= p.D) {
    p.tma_kv = make_map(&p.tmap_k, MapKey{p.k, p.D, p.HKV, kv_pool_tokens, p.kv_head_stride, p.kv_tok_stride, 1, 32}) &&
               make_map(&p.tmap_v, MapKey{p.v, p.D, p.HKV, kv_pool_tokens, p.kv_head_stride, p.kv_tok_stride, 1, 32}) &&
               make_map(&p.tmap_k16, MapKey{p.k, p.D, p.HKV, kv_pool_tokens, p.kv_head_stride, p.kv_tok_stride, 1, 16}) &&
               make_map(&p.tmap_v16, MapKey{p.v, p.D, p.HKV, kv_pool_tokens, p.kv_head_stride, p.kv_tok_stride, 1, 16}) &&
               make_map(&p.tmap_k8, MapKey{p.k, p.D, p.HKV, kv_pool_tokens, p.kv_head_stride, p.kv_tok_stride, 1, 8}) &&
               make_map(&p.tmap_v8, MapKey{p.v, p.D, p.HKV, kv_pool_tokens, p.kv_head_stride, p.kv_tok_stride, 1, 8});
    // gather4 view: head rows at a constant pitch (the token stride must be a whole number of head strides)
    const int64_t ratio = p.kv_tok_stride / p.kv_head_stride;
    const int64_t rows = (kv_pool_tokens - 1) * ratio + p.HKV;
    if (!g_no_gather4 && p.tma_kv && ratio * p.kv_head_stride == p.kv_tok_stride && ratio >= p.HKV && rows < (1ll << 31)) {
      p.kv_row_ratio = (int32_t)ratio;
      p.kv_rows = (int32_t)rows;
      p.tma_gather = make_map(&p.tmap_kg, MapKey{p.k, p.D, rows, 0, p.kv_head_stride, 0, 1, 0}) &&
                     make_map(&p.tmap_vg, MapKey{p.v, p.D, rows, 0, p.kv_head_stride, 0, 1, 0});
    }
  }
  if (p.q_head_stride >= p.D && p.q_row_stride >= p.D)
    p.tma_q = make_map(&p.tmap_q, MapKey{p.q, p.D, p.H, p.nq, p.q_head_stride, p.q_row_stride, G, 32});
}

// Fused append: tensor maps over the step's new K / V rows, the mirror image of the pool's
int setup_append(AttnParams& p, const deft_append_t* a) {
  DEFT_CHECK_ARG(a->new_k && a->new_v && a->cache_loc, "append: null pointer");
  DEFT_CHECK_ARG(((uintptr_t)a->new_k | (uintptr_t)a->new_v) % 16 == 0 && (a->new_row_stride | a->new_head_stride) % 8 == 0,
                 "append: new_k / new_v must be 16-byte aligned with strides in multiples of 8 elements");
  DEFT_CHECK_ARG(p.tma_kv && p.tma_gather, "append: needs the TMA paths of the tensor-core kernel");
  const int64_t ratio = a->new_head_stride > 0 ? a->new_row_stride / a->new_head_stride : 0;
  DEFT_CHECK_ARG(a->new_head_stride >= p.D && ratio >= p.HKV && ratio * a->new_head_stride == a->new_row_stride,
                 "append: the row stride of new_k / new_v must be a whole number of head strides");
  const int64_t rows = (int64_t)(p.nq - 1) * ratio + p.HKV;
  p.new_k = static_cast<const __half*>(a->new_k);
  p.new_v = static_cast<const __half*>(a->new_v);
  p.new_row_stride = a->new_row_stride; p.new_head_stride = a->new_head_stride;
  p.cache_loc = a->cache_loc;
  p.new_row_ratio = (int32_t)ratio; p.new_rows = (int32_t)rows;
  bool ok = true;
  const int boxes[3] = {32, 16, 8};
  CUtensorMap* km[3] = {&p.tmap_nk, &p.tmap_nk16, &p.tmap_nk8};
  CUtensorMap* vm[3] = {&p.tmap_nv, &p.tmap_nv16, &p.tmap_nv8};
  for (int i = 0; i < 3; ++i) {
    ok = ok && make_map(km[i], MapKey{a->new_k, p.D, p.HKV, p.nq, a->new_head_stride, a->new_row_stride, 1, boxes[i]});
    ok = ok && make_map(vm[i], MapKey{a->new_v, p.D, p.HKV, p.nq, a->new_head_stride, a->new_row_stride, 1, boxes[i]});
  }
  ok = ok && make_map(&p.tmap_nkg, MapKey{a->new_k, p.D, rows, 0, a->new_head_stride, 0, 1, 0});
  ok = ok && make_map(&p.tmap_nvg, MapKey{a->new_v, p.D, rows, 0, a->new_head_stride, 0, 1, 0});
  DEFT_CHECK_ARG(ok, "append: could not encode the tensor maps of new_k / new_v");
  return DEFT_OK;
}

// Which stage-1 kernel a call of this thread runs (the stage-2 kernel and the partial layout follow it).
bool use_umma(int32_t H, int32_t HKV, int32_t D) {
  if (g_stage1_impl == DEFT_STAGE1_FMA) return false;
  AttnParams probe{};
  probe.H = H; probe.HKV = HKV; probe.D = D;
  return g_stage1_impl == DEFT_STAGE1_UMMA || stage1_umma_supported(probe);
}

// Carves the caller's workspace.  Layout (all 256-byte aligned):
//   tcgen05 path: po16 [slots*HKV][D/8][32G][8] f16 | plse16 [slots*HKV][32G] f32
//   warp-FMA path: po [rows][H][D] f32 | plse [rows][H] f32
//   device-derived plan: items | groups | units | csr_off | csr_rows | cursor | counters
struct Workspace {
  float* po;
  float* plse;
  __half* po16;
  float* plse16;
  PlanBuffers pb;
  size_t bytes;
};

struct Sizes {
  int64_t fma_rows;     // partial rows of the warp-FMA layout
  int64_t slots;        // unit slots (32 partial rows each) of the tile layout
  int64_t groups;       // device plan: bound on items / groups / units
  int64_t csr_entries;  // device plan: bound on sum of q_cnt over groups
};

Workspace carve(void* base, const Sizes& z, bool umma, bool with_plan, int32_t nq, int32_t H, int32_t HKV, int32_t D) {
  Workspace w{};
  size_t off = 0;
  auto take = [&](size_t n) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(n);
    return p;
  };
  if (umma) {
    const size_t tile_rows = (size_t)kMaxGroupQ * (H / HKV);
    w.po16 = static_cast<__half*>(take((size_t)z.slots * HKV * tile_rows * D * sizeof(__half)));
    w.plse16 = static_cast<float*>(take((size_t)z.slots * HKV * tile_rows * sizeof(float)));
  } else {
    w.po = static_cast<float*>(take((size_t)z.fma_rows * H * D * sizeof(float)));
    w.plse = static_cast<float*>(take((size_t)z.fma_rows * H * sizeof(float)));
  }
  if (with_plan) {
    w.pb.items = static_cast<deft_item_t*>(take((size_t)z.groups * sizeof(deft_item_t)));
    w.pb.groups = static_cast<deft_group_t*>(take((size_t)z.groups * sizeof(deft_group_t)));
    w.pb.units = static_cast<deft_unit_t*>(take((size_t)z.groups * sizeof(deft_unit_t)));
    w.pb.csr_off = static_cast<int32_t*>(take((size_t)(nq + 1) * sizeof(int32_t)));
    w.pb.csr_rows = static_cast<int32_t*>(take((size_t)z.csr_entries * sizeof(int32_t)));
    w.pb.cursor = static_cast<int32_t*>(take((size_t)(nq + 1) * sizeof(int32_t)));
    w.pb.counters = static_cast<int32_t*>(take(16));
  }
  w.bytes = off;
  return w;
}

inline int64_t node_items_bound(int64_t n_entries, int64_t total_kv_bound) {
  return n_entries + (total_kv_bound > 0 ? total_kv_bound / kNodeSplit + 1 : 0);
}
inline int64_t node_csr_bound(int64_t n_partials, int64_t total_kv_bound) {
  return n_partials + (total_kv_bound > 0 ? kMaxGroupQ * (total_kv_bound / kNodeSplit + 1) : 0);
}

Sizes flatten_sizes(const deft_plan_t* plan, int64_t n_partials, int64_t n_blocks) {
  if (plan) return Sizes{plan->n_part_rows, plan->n_unit_slots, 0, 0};
  return Sizes{n_blocks * kMaxGroupQ, n_blocks, n_blocks, n_partials};
}
Sizes node_sizes(const deft_plan_t* plan, int64_t n_partials, int64_t n_entries, int64_t total_kv_bound) {
  if (plan) return Sizes{plan->n_part_rows, plan->n_unit_slots, 0, 0};
  const int64_t items = node_items_bound(n_entries, total_kv_bound);
  return Sizes{items * kMaxGroupQ, items, items, node_csr_bound(n_partials, total_kv_bound)};
}

int check_common(const void* q, const void* k, const void* v, const void* o, int32_t nq, int32_t H,
                 int32_t HKV, int32_t D, int64_t q_row_stride, int64_t q_head_stride,
                 int64_t kv_tok_stride, int64_t kv_head_stride, int64_t o_row_stride,
                 int64_t o_head_stride) {
  DEFT_CHECK_ARG(q && k && v && o, "null tensor pointer");
  DEFT_CHECK_ARG(nq > 0 && H > 0 && HKV > 0 && H % HKV == 0, "bad head geometry nq=%d H=%d HKV=%d", nq, H, HKV);
  // the reference asserts head_dim in {16, 32, 64, 128} (tree_attention.py:100,305,582)
  DEFT_CHECK_ARG(D == 16 || D == 32 || D == 64 || D == 128, "head_dim %d not supported (16, 32, 64, 128)", D);
  DEFT_CHECK_ARG(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) % 16 == 0,
                 "q/k/v/o must be 16-byte aligned");
  DEFT_CHECK_ARG((q_row_stride | q_head_stride | kv_tok_stride | kv_head_stride | o_row_stride | o_head_stride) % 8 == 0,
                 "strides must be multiples of 8 elements (16 bytes)");
  return DEFT_OK;
}

AttnParams base_params(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                       const void* v, int64_t kv_tok_stride, int64_t kv_head_stride, void* o,
                       int64_t o_row_stride, int64_t o_head_stride, int32_t nq, int32_t H, int32_t HKV,
                       int32_t D) {
  AttnParams p{};
  p.q = static_cast<const __half*>(q);
  p.k = static_cast<const __half*>(k);
  p.v = static_cast<const __half*>(v);
  p.o = static_cast<__half*>(o);
  p.q_row_stride = q_row_stride; p.q_head_stride = q_head_stride;
  p.kv_tok_stride = kv_tok_stride; p.kv_head_stride = kv_head_stride;
  p.o_row_stride = o_row_stride; p.o_head_stride = o_head_stride;
  p.nq = nq; p.H = H; p.HKV = HKV; p.D = D;
  p.scale = 1.0f / sqrtf((float)D);  // recomputed from head_dim like the reference (tree_attention.py:104)
  return p;
}

// Host-built plan: the warp-FMA path reads the item/group layer (reference tables), the tcgen05 path
// the unit layer (native 32-bit tables).
int use_plan(AttnParams& p, const deft_plan_t* plan, bool umma) {
  if (umma) {
    DEFT_CHECK_ARG(plan->units && plan->u_csr_off && plan->u_csr_rows && plan->u_kv && plan->u_q,
                   "plan has no unit layer (rebuild the tables with this library version)");
    p.units = plan->units; p.n_units = plan->n_units; p.n_units_dev = nullptr;
    p.u_kv = plan->u_kv; p.u_kv_bytes = 4;
    p.u_blk = plan->u_blk;
    p.u_mask = plan->u_mask; p.u_mask_bytes = 4;
    p.u_q = plan->u_q; p.u_q_bytes = 4;
    p.u_csr_off = plan->u_csr_off; p.u_csr_rows = plan->u_csr_rows;
    if (plan->u_job_off && plan->u_jobs && plan->hkv == p.HKV && plan->n_ctas > 0) {
      p.job_off = plan->u_job_off; p.jobs = plan->u_jobs; p.n_ctas = plan->n_ctas;
      p.clustered = plan->paired;  // pair-aligned lists: clusters of two CTAs (the launch drops it when the grid is odd)
    }
  } else {
    p.items = plan->items; p.groups = plan->groups;
    p.csr_off = plan->csr_off; p.csr_rows = plan->csr_rows;
    p.n_items = plan->n_items; p.n_items_dev = nullptr;
  }
  return DEFT_OK;
}
// Device-derived plan: both layers index the reference tables the call was given.
void use_plan(AttnParams& p, const PlanBuffers& pb, int64_t bound) {
  p.items = pb.items; p.groups = pb.groups;
  p.csr_off = pb.csr_off; p.csr_rows = pb.csr_rows;
  p.n_items = (int32_t)bound; p.n_items_dev = pb.counters;
  p.units = pb.units; p.n_units = (int32_t)bound; p.n_units_dev = pb.counters + 2;
  p.u_kv = p.kv_idx; p.u_kv_bytes = p.kv_idx_bytes;
  p.u_blk = nullptr;
  p.u_mask = p.masks; p.u_mask_bytes = 8;
  p.u_q = p.q_list; p.u_q_bytes = 8;
  p.u_csr_off = pb.csr_off; p.u_csr_rows = pb.csr_rows;
  p.plan_fresh = 1;
}

int run_stages(const AttnParams& p, bool umma, cudaStream_t stream) {
  if (g_stages & DEFT_STAGE_1) {
    int rc;
    if (umma) {
      AttnParams pd = p;
      pd.dbg = g_debug;
      pd.trace = g_trace;
      pd.experiment = g_experiment;
      rc = launch_stage1_umma(pd, stream);
    } else {
      rc = launch_stage1_fma(p, stream);
    }
    if (rc) return rc;
  }
  if (g_stages & DEFT_STAGE_2) return umma ? launch_stage2_tiles(p, stream) : launch_stage2(p, stream);
  return DEFT_OK;
}

__global__ void kv_append_kernel(__half* k, __half* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                                 const __half* nk, const __half* nv, int64_t new_row_stride,
                                 int64_t new_head_stride, const int32_t* loc, int n, int HKV, int CH) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
  const int64_t total = (int64_t)n * HKV * CH;
  // launched programmatically (the launch and the block scheduling overlap the previous kernel's tail): nothing is
  // read or written before the previous kernel's memory is visible.  A no-op for a plain launch.
  // ... and the kernel after this one (stage 1) may start ITS prologue right away (barrier init, TMEM allocation, job
  // record, tensor-map prefetch: ~1.5 us): it waits for this grid's completion before it touches the pool
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (i >= total) return;
  const int ch = (int)(i % CH), h = (int)((i / CH) % HKV), r = (int)(i / ((int64_t)CH * HKV));
  const int64_t src = (int64_t)r * new_row_stride + (int64_t)h * new_head_stride + ch * 8;
  const int64_t dst = (int64_t)loc[r] * kv_tok_stride + (int64_t)h * kv_head_stride + ch * 8;
  *reinterpret_cast<uint4*>(k + dst) = *reinterpret_cast<const uint4*>(nk + src);
  *reinterpret_cast<uint4*>(v + dst) = *reinterpret_cast<const uint4*>(nv + src);
}

}  // namespace
}  // namespace deft

using namespace deft;

extern "C" {

int deft_b200_abi_version(void) { return DEFT_B200_ABI_VERSION; }
const char* deft_b200_last_error(void) { return g_error; }
void deft_b200_set_stages(int32_t mask) { g_stages = mask; }
void deft_b200_set_stage1_impl(int32_t impl) { g_stage1_impl = impl; }
void deft_b200_set_debug_buffer(void* dev) { g_debug = static_cast<float*>(dev); }
void deft_b200_set_trace_buffer(void* dev) { g_trace = static_cast<int*>(dev); }
void deft_b200_set_tma(int32_t enabled) { g_no_tma = enabled == 0; }
void deft_b200_set_pdl(int32_t enabled) { g_no_pdl = enabled == 0; }
void deft_b200_set_gather4(int32_t enabled) { g_no_gather4 = enabled == 0; }
void deft_b200_set_experiment(int32_t flags) { g_experiment = flags; }

size_t deft_b200_flatten_workspace_bytes(int32_t nq, int32_t H, int32_t HKV, int32_t D,
                                         int64_t n_partials, int64_t n_blocks, const deft_plan_t* plan) {
  if (HKV <= 0 || H % HKV) return 0;
  return carve(nullptr, flatten_sizes(plan, n_partials, n_blocks), use_umma(H, HKV, D), plan == nullptr, nq, H, HKV, D).bytes;
}

size_t deft_b200_node_workspace_bytes(int32_t nq, int32_t H, int32_t HKV, int32_t D, int64_t n_partials,
                                      int64_t n_entries, int64_t total_kv_bound, const deft_plan_t* plan) {
  if (HKV <= 0 || H % HKV) return 0;
  return carve(nullptr, node_sizes(plan, n_partials, n_entries, total_kv_bound), use_umma(H, HKV, D),
               plan == nullptr, nq, H, HKV, D).bytes;
}

int deft_b200_flatten_fwd(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                          const void* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                          int64_t kv_pool_tokens, void* o,
                          int64_t o_row_stride, int64_t o_head_stride, int32_t nq, int32_t H,
                          int32_t HKV, int32_t D, int32_t block_len, const int64_t* block_q,
                          int64_t n_partials, const int64_t* block_q_cnts,
                          const int64_t* block_q_offset, const int64_t* block_lens, int64_t n_blocks,
                          const int64_t* block_bitmasks, const int64_t* block_kv,
                          const deft_plan_t* plan, void* workspace, size_t workspace_bytes,
                          void* stream_) {
  return deft_b200_flatten_fwd_append(q, q_row_stride, q_head_stride, k, v, kv_tok_stride, kv_head_stride, kv_pool_tokens, o,
                                      o_row_stride, o_head_stride, nq, H, HKV, D, block_len, block_q, n_partials, block_q_cnts,
                                      block_q_offset, block_lens, n_blocks, block_bitmasks, block_kv, plan, nullptr, workspace,
                                      workspace_bytes, stream_);
}

int deft_b200_flatten_fwd_append(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                                 const void* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                                 int64_t kv_pool_tokens, void* o,
                                 int64_t o_row_stride, int64_t o_head_stride, int32_t nq, int32_t H,
                                 int32_t HKV, int32_t D, int32_t block_len, const int64_t* block_q,
                                 int64_t n_partials, const int64_t* block_q_cnts,
                                 const int64_t* block_q_offset, const int64_t* block_lens, int64_t n_blocks,
                                 const int64_t* block_bitmasks, const int64_t* block_kv,
                                 const deft_plan_t* plan, const deft_append_t* append, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_common(q, k, v, o, nq, H, HKV, D, q_row_stride, q_head_stride, kv_tok_stride,
                        kv_head_stride, o_row_stride, o_head_stride);
  if (rc) return rc;
  DEFT_CHECK_ARG(block_len == 128, "block_len must be 128 (got %d): the reference Flatten kernel hard-wires it", block_len);
  const bool umma = use_umma(H, HKV, D);
  // (a native-only build, deft_b200_layout_set_native_only: the unit plan is all the tensor-core path reads)
  const bool plan_only = umma && plan && plan->units && plan->u_blk && n_blocks == 0 && n_partials == 0;
  DEFT_CHECK_ARG(plan_only || (block_q && block_q_cnts && block_q_offset && block_lens && block_bitmasks && block_kv),
                 "null table pointer");
  DEFT_CHECK_ARG(plan_only || (n_blocks > 0 && n_partials > 0), "empty tables");
  DEFT_CHECK_ARG(workspace, "null workspace");
  Workspace w = carve(workspace, flatten_sizes(plan, n_partials, n_blocks), umma, plan == nullptr, nq, H, HKV, D);
  if (w.bytes > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return DEFT_E_WORKSPACE;
  }
  AttnParams p = base_params(q, q_row_stride, q_head_stride, k, v, kv_tok_stride, kv_head_stride, o,
                             o_row_stride, o_head_stride, nq, H, HKV, D);
  p.kv_idx = block_kv; p.kv_idx_bytes = 8;
  p.q_list = block_q; p.masks = block_bitmasks;
  p.po = w.po; p.plse = w.plse; p.po16 = w.po16; p.plse16 = w.plse16;
  if (umma) setup_tma(p, kv_pool_tokens);
  p.pdl = g_no_pdl ? 0 : 1;
  if (append) {
    DEFT_CHECK_ARG(umma && plan && plan->u_blk && plan->fresh, "append: needs the tensor-core path and a host-built plan (tables built with fresh_page)");
    rc = setup_append(p, append);
    if (rc) return rc;
  }
  DEFT_CHECK_ARG(!(plan && plan->fresh && !append), "this plan reads the step's tokens from the activations: call the *_append form");
  if (plan) {
    rc = use_plan(p, plan, umma);
    if (rc) return rc;
  } else {
    if (g_stages & DEFT_STAGE_PLAN) {
      rc = launch_plan_flatten(block_q_cnts, block_q_offset, block_lens, block_kv, block_q, n_blocks,
                               block_len, nq, w.pb, stream);
      if (rc) return rc;
    }
    use_plan(p, w.pb, n_blocks);
  }
  return run_stages(p, umma, stream);
}

int deft_b200_node_fwd(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                       const void* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                       int64_t kv_pool_tokens, void* o,
                       int64_t o_row_stride, int64_t o_head_stride, int32_t nq, int32_t H,
                       int32_t HKV, int32_t D, const void* kv_indices, int32_t kv_index_bytes,
                       const int64_t* kv_offset, const int64_t* kv_len, const int64_t* node_q,
                       int64_t n_partials, const int64_t* q_offset, const int64_t* q_len,
                       int64_t n_entries, int64_t total_kv_bound, const deft_plan_t* plan,
                       void* workspace, size_t workspace_bytes, void* stream_) {
  return deft_b200_node_fwd_append(q, q_row_stride, q_head_stride, k, v, kv_tok_stride, kv_head_stride, kv_pool_tokens, o,
                                   o_row_stride, o_head_stride, nq, H, HKV, D, kv_indices, kv_index_bytes, kv_offset, kv_len, node_q,
                                   n_partials, q_offset, q_len, n_entries, total_kv_bound, plan, nullptr, workspace,
                                   workspace_bytes, stream_);
}

int deft_b200_node_fwd_append(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                              const void* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                              int64_t kv_pool_tokens, void* o,
                              int64_t o_row_stride, int64_t o_head_stride, int32_t nq, int32_t H,
                              int32_t HKV, int32_t D, const void* kv_indices, int32_t kv_index_bytes,
                              const int64_t* kv_offset, const int64_t* kv_len, const int64_t* node_q,
                              int64_t n_partials, const int64_t* q_offset, const int64_t* q_len,
                              int64_t n_entries, int64_t total_kv_bound, const deft_plan_t* plan,
                              const deft_append_t* append, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_common(q, k, v, o, nq, H, HKV, D, q_row_stride, q_head_stride, kv_tok_stride,
                        kv_head_stride, o_row_stride, o_head_stride);
  if (rc) return rc;
  DEFT_CHECK_ARG(kv_index_bytes == 8 || kv_index_bytes == 4, "kv_index_bytes must be 8 or 4");
  const bool umma = use_umma(H, HKV, D);
  const bool plan_only = umma && plan && plan->units && plan->u_blk && n_entries == 0 && n_partials == 0;
  DEFT_CHECK_ARG(plan_only || (kv_indices && kv_offset && kv_len && node_q && q_offset && q_len), "null table pointer");
  DEFT_CHECK_ARG(plan_only || (n_entries > 0 && n_partials > 0), "empty tables");
  DEFT_CHECK_ARG(workspace, "null workspace");
  const int64_t items = node_items_bound(n_entries, total_kv_bound);
  Workspace w = carve(workspace, node_sizes(plan, n_partials, n_entries, total_kv_bound), umma, plan == nullptr,
                      nq, H, HKV, D);
  if (w.bytes > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return DEFT_E_WORKSPACE;
  }
  AttnParams p = base_params(q, q_row_stride, q_head_stride, k, v, kv_tok_stride, kv_head_stride, o,
                             o_row_stride, o_head_stride, nq, H, HKV, D);
  p.kv_idx = kv_indices; p.kv_idx_bytes = kv_index_bytes;
  p.q_list = node_q; p.masks = nullptr;
  p.po = w.po; p.plse = w.plse; p.po16 = w.po16; p.plse16 = w.plse16;
  if (umma) setup_tma(p, kv_pool_tokens);
  p.pdl = g_no_pdl ? 0 : 1;
  if (append) {
    DEFT_CHECK_ARG(umma && plan && plan->u_blk && plan->fresh, "append: needs the tensor-core path and a host-built plan (tables built with fresh_page)");
    rc = setup_append(p, append);
    if (rc) return rc;
  }
  DEFT_CHECK_ARG(!(plan && plan->fresh && !append), "this plan reads the step's tokens from the activations: call the *_append form");
  if (plan) {
    rc = use_plan(p, plan, umma);
    if (rc) return rc;
  } else {
    if (g_stages & DEFT_STAGE_PLAN) {
      // long entries are cut into items of `split` tokens: kNodeSplit (what the workspace was sized for) when
      // the call is small, longer -- fewer jobs and partials -- when there is work for several waves of CTAs
      int32_t split = 0;
      if (total_kv_bound > 0) {
        const int64_t want = total_kv_bound * HKV / (2 * 148) / 128 * 128;
        split = (int32_t)(want < kNodeSplit ? kNodeSplit : want > 2048 ? 2048 : want);
      }
      rc = launch_plan_node(kv_offset, kv_len, q_offset, q_len, node_q, n_entries, split, nq, w.pb, stream);
      if (rc) return rc;
    }
    use_plan(p, w.pb, items);
  }
  return run_stages(p, umma, stream);
}

int deft_b200_kv_append(void* k, void* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                        const void* new_k, const void* new_v, int64_t new_row_stride,
                        int64_t new_head_stride, const int32_t* cache_loc, int32_t n, int32_t HKV,
                        int32_t D, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEFT_CHECK_ARG(k && v && new_k && new_v && cache_loc, "null pointer");
  DEFT_CHECK_ARG(n >= 0 && HKV > 0 && D > 0 && D % 8 == 0, "bad shape n=%d HKV=%d D=%d", n, HKV, D);
  DEFT_CHECK_ARG(((uintptr_t)k | (uintptr_t)v | (uintptr_t)new_k | (uintptr_t)new_v) % 16 == 0,
                 "buffers must be 16-byte aligned");
  DEFT_CHECK_ARG((kv_tok_stride | kv_head_stride | new_row_stride | new_head_stride) % 8 == 0,
                 "strides must be multiples of 8 elements");
  if (n == 0) return DEFT_OK;
  const int CH = D / 8;
  const int64_t total = (int64_t)n * HKV * CH;
  const int threads = 256;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((total + threads - 1) / threads));
  cfg.blockDim = dim3(threads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_no_pdl ? 0 : 1;
  DEFT_CUDA(cudaLaunchKernelEx(&cfg, kv_append_kernel, static_cast<__half*>(k), static_cast<__half*>(v), kv_tok_stride,
                               kv_head_stride, static_cast<const __half*>(new_k), static_cast<const __half*>(new_v),
                               new_row_stride, new_head_stride, cache_loc, n, HKV, CH));
  return DEFT_OK;
}

}  // extern "C"
