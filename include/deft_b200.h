/*
 * deft_b200.h -- C ABI of libdeft_b200.so: B200 (sm_100a) tree-attention decode for DeFT.
 *
 * This is the drop-in boundary for the ONE path this repository accelerates: DeFT's decode-step
 * tree attention (DeFT-Flatten / DeFT-Node / DeFT-Node-Chunk / Tree-Index) over its token-granular
 * paged KV pool, plus the KV-guided-grouping metadata builder that drives it.  Every entry point
 * names the reference interface it replaces (paths under LINs-lab/DeFT, DeFT/deft/...).
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are marked [dev], host pointers [host]
 *   - all tensors are caller-owned and must outlive the (asynchronous) call
 *   - launches go to the cudaStream_t passed as `void* stream` (0 = legacy default stream)
 *   - return 0 on success, a negative DEFT_E_* code otherwise; deft_b200_last_error() gives text
 *   - launches happen on the CURRENT device, which must be the device `stream` and the tensors belong to (the
 *     Python shim switches to the tensors' device around a call when it is not the current one)
 *   - thread-compatible: no global mutable state besides a thread-local error string, per-thread tensor-map
 *     caches and once-only PER-DEVICE cudaFuncSetAttribute calls (a process may drive several GPUs)
 *   - activations / KV are IEEE fp16 (the reference hard-codes torch.float16, model_runner.py:271,336)
 */
#ifndef DEFT_B200_H_
#define DEFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEFT_B200_ABI_VERSION 5

enum {
  DEFT_OK = 0,
  DEFT_E_ARG = -1,       /* bad argument (shape, alignment, null pointer) */
  DEFT_E_WORKSPACE = -2, /* workspace too small */
  DEFT_E_CUDA = -3,      /* a CUDA runtime call failed */
  DEFT_E_TREE = -4       /* malformed tree handed to the metadata builder */
};

int deft_b200_abi_version(void);
const char* deft_b200_last_error(void);

/* Measurement hook (bench.py): which stages the *_fwd calls of THIS thread launch.
 * bit 0 = device plan derivation, bit 1 = stage 1 (partial softmax), bit 2 = stage 2 (combine).
 * Default 7.  Lets a kernel be timed alone, back to back, on the launching stream. */
#define DEFT_STAGE_PLAN 1
#define DEFT_STAGE_1 2
#define DEFT_STAGE_2 4
void deft_b200_set_stages(int32_t mask);

/* Stage-1 kernel selection of THIS thread.  AUTO: the tcgen05 tensor-core kernel when the geometry
 * fits it (head_dim 64/128, H/HKV in {1,2,4}), else the warp-FMA kernel. */
#define DEFT_STAGE1_AUTO 0
#define DEFT_STAGE1_FMA 1
#define DEFT_STAGE1_UMMA 2
void deft_b200_set_stage1_impl(int32_t impl);
/* Test hook: with DEFT_STAGE1_UMMA forced, the first work unit dumps its raw S [128][128] and
 * O [128][D] accumulators (fp32) to this device buffer.  NULL (default) disables. */
void deft_b200_set_debug_buffer(void* dev);
/* Profiling hook: the tcgen05 stage 1 records a per-CTA timeline ([n_ctas][128] int32, SM cycles since
 * the CTA started; event ids in csrc/attn_umma.cu) into this device buffer.  NULL (default) disables. */
void deft_b200_set_trace_buffer(void* dev);
/* Test hook: 0 = never use TMA (every K/V/Q row is gathered with cp.async), 1 (default) = TMA for runs
 * of 128 consecutive pages and for slots of consecutive query ids. */
void deft_b200_set_tma(int32_t enabled);
/* Test hook: 0 = plain launches, 1 (default) = the tcgen05-path kernels are launched with programmatic
 * stream serialization: each starts while its predecessor drains, does its dependency-free prologue
 * (barriers, TMEM, plan tables) and only then waits for the predecessor (griddepcontrol.wait). */
void deft_b200_set_pdl(int32_t enabled);
/* Test hook: 0 = pages that are not consecutive are gathered with cp.async (16-byte copies), 1 (default) =
 * with TMA tile::gather4 (four K or V rows per instruction) when the pool's token stride is a whole
 * number of head strides. */
void deft_b200_set_gather4(int32_t enabled);
/* Profiling hook: bit flags selecting kernel variants for A/B measurements on one box (0 = the default path). */
void deft_b200_set_experiment(int32_t flags);

/* ------------------------------------------------------------------------------------------
 * Work plan (device side).  Two layers:
 *
 * (1) items / groups over the REFERENCE tables (used by the warp-FMA stage 1).  One *item* = one KV
 *     token range of the table, attended by 1..n *groups* of <= 32 queries; a group row r stands for
 *     the G = H/HKV GQA heads of query q_list[q_off + r]; `part_base + r` is the partial row.
 * (2) units (used by the tcgen05 stage 1).  One *unit* = a chain of n_tiles KV tiles (128 tokens
 *     each, the last one possibly shorter) attended by one or two *slots* of <= 32 queries; the
 *     kernel walks the chain with an online softmax and emits ONE partial per (unit, slot, kv-head).
 *     `part_base[s]` is a multiple of 32; partial row = part_base[s] + r.  Units index either the
 *     reference tables (device-derived plans) or the builder's compact native tables.
 *
 * The plan is either built on the host by deft_b200_build_tables() (and uploaded by the caller in
 * one copy) or derived on the device from the reference tables by the *_fwd calls.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t kv_off;   /* first element of the KV index table this item reads */
  int32_t kv_len;   /* tokens in the item (any length; tiled inside the kernel) */
  int32_t grp_off;  /* first group */
  int32_t n_grp;    /* number of groups sharing the KV range */
  int32_t cost;     /* scheduling weight (tokens x rows); informational */
} deft_item_t;

typedef struct {
  int64_t mask_off;  /* first element of the per-token bitmask table, -1 = no mask (all attend) */
  int32_t q_off;     /* first element of the query list */
  int32_t q_cnt;     /* 1..32 */
  int32_t part_base; /* first partial row */
  int32_t pad;
} deft_group_t;

typedef struct {
  int64_t kv_off;           /* element of the KV index table holding token 0 of tile 0 */
  int64_t mask_off[2];      /* element of the mask table for token 0 of tile 0, per slot; -1 = every
                               live row attends every token of every tile (no table read) */
  int32_t kv_tile_stride;   /* elements between consecutive tiles in the KV index table */
  int32_t mask_tile_stride; /* elements between consecutive tiles in the mask table */
  int32_t n_tiles;          /* >= 1 */
  int32_t last_len;         /* tokens in the last tile, 1..128 (all other tiles hold 128) */
  int32_t q_off[2];         /* first element of the query list, per slot */
  int32_t q_cnt[2];         /* 1..32; q_cnt[1] == 0: single-slot unit */
  int32_t part_base[2];     /* multiple of 32 */
  /* Shortcuts the builder fills in when it can (-1 otherwise): with them the kernel issues its TMA loads
   * straight from the unit record, without the dependent reads of the index tables. */
  int32_t page0;            /* >= 0: every tile is full and the unit's tokens sit on consecutive pages
                               page0, page0 + 1, ... (the prompt) */
  int32_t q_id0[2];         /* >= 0: the slot's queries have consecutive ids q_id0[s], q_id0[s] + 1, ... */
  int32_t dense_tiles;      /* the unit's first dense_tiles tiles are full and attended by every row of its live
                               slots (a prompt ahead of the first subtree tile): no mask is read for them */
} deft_unit_t;              /* 80 bytes */

/* One entry of the per-CTA job lists of the unit plan.  A job is ONE slot of a unit on one kv-head:
 * job = ((unit * HKV + kv_head) << 1) | slot.  The record carries a copy of its unit so that a CTA
 * starts from one load.  Records [0, n_ctas) are the FIRST job of every CTA (job < 0: none) with, in
 * n_jobs / next, how many jobs the CTA has and where its further records start (consecutive). */
typedef struct {
  int32_t job;
  int32_t n_jobs;  /* first record of a CTA only: jobs of this CTA (>= 1 when job >= 0) */
  int32_t next;    /* first record of a CTA only: index of its second record */
  int32_t shared;  /* 1: CTA (c ^ 1) holds, at the same position of its list, the other slot of the same (unit,
                      kv_head): the pair loads every K/V tile once (each CTA half of it, TMA multicast) */
  deft_unit_t unit;
} deft_job_t;               /* 96 bytes */

typedef struct {
  /* (1) item/group plan over the reference tables */
  const deft_item_t* items;   /* [dev] */
  const deft_group_t* groups; /* [dev] */
  const int32_t* csr_off;     /* [dev] nq+1: partial rows of query q are csr_rows[csr_off[q]:csr_off[q+1]] */
  const int32_t* csr_rows;    /* [dev] ascending within a query (deterministic merge order) */
  int32_t n_items;
  int32_t n_groups;
  int32_t n_part_rows;
  int32_t n_units;
  /* (2) unit plan over the native tables (all [dev]; NULL/0 when absent) */
  const deft_unit_t* units;
  const int32_t* u_csr_off;   /* nq+1 */
  const int32_t* u_csr_rows;
  const int32_t* u_kv;        /* page id per token slot, tiles of 128; -1 = dummy token (a zero row nobody attends);
                                 bit 30 set: a token of THIS decode step, the low bits are its query id (its K/V row is
                                 read from the step's activations, see deft_append_t) */
  const int32_t* u_blk;       /* per chunk of 8 token slots (16 per tile): how the chunk is loaded, (kind << 28) | first page.
                                 kind 3 / 2 / 1: the chunk lies in an aligned run of 32 / 16 / 8 consecutive pages (one TMA
                                 box per panel for the whole run), 0: gathered four rows at a time; bit 27: the chunk holds
                                 tokens of this step (rows of the activations, "page" = first query id).  May be NULL. */
  const uint32_t* u_mask;     /* 128 words per (tile, slot): bit r = row r of the slot attends */
  const int32_t* u_q;         /* query id per (slot, row) */
  const int32_t* u_job_off;   /* n_ctas+1: CTA c has u_job_off[c+1] - u_job_off[c] jobs (u_jobs carries the lists) */
  const deft_job_t* u_jobs;   /* job records, balanced over CTAs by the builder (see deft_job_t) */
  int32_t n_unit_slots;       /* partial tiles per kv-head */
  int32_t n_ctas;             /* CTAs the job lists were balanced for */
  int32_t hkv;                /* kv-head count the job lists were built for */
  int32_t paired;             /* 1: pair-aligned job lists (deft_job_t.shared): the kernel is launched as clusters of 2 */
  int32_t fresh;              /* 1: the tables were built with fresh_page: the plan can only run through the *_append forms */
  int32_t pad;
} deft_plan_t;

/* ------------------------------------------------------------------------------------------
 * DeFT-Flatten operator.
 * Replaces tree_attention_subtree_fwd (layers/attention/tree_attention.py:552-667: stage-1 kernel2
 * :860-976 + DeFT_splitBynode_Triton_stage2 :297-416).  Argument meaning is the reference's:
 *   q  [nq, H, D] fp16, strides in elements (row stride 6144 for the fused-qkv view)
 *   k/v[pool, HKV, D] fp16 views of kv_data[layer][:,0] / [:,1] (memory_pool.py:68-72);
 *      kv_pool_tokens = pool size (pages): bounds the TMA tensor maps used for runs of consecutive
 *      pages; 0 disables TMA (every page is then gathered with cp.async)
 *   o  [nq, H, D] fp16, fully overwritten (the reference requires it pre-zeroed; we do not)
 *   block_q [n_partials], block_q_cnts/offset/lens [n_blocks], block_bitmasks/block_kv
 *   [n_blocks*block_len] -- int64 device tables of TreeMetadata (tree_cache.py:591-616)
 * block_len must be 128 (the reference kernel hard-wires BLOCK_N=128, tree_attention.py:655-657).
 * `plan` may be NULL: the plan is then derived on the device inside `workspace`.
 * The workspace size depends on the plan (pass the same `plan`, or NULL, to *_workspace_bytes).
 * ------------------------------------------------------------------------------------------ */
size_t deft_b200_flatten_workspace_bytes(int32_t nq, int32_t H, int32_t HKV, int32_t D,
                                         int64_t n_partials, int64_t n_blocks,
                                         const deft_plan_t* plan);

int deft_b200_flatten_fwd(const void* q, int64_t q_row_stride, int64_t q_head_stride,
                          const void* k, const void* v, int64_t kv_tok_stride,
                          int64_t kv_head_stride, int64_t kv_pool_tokens, void* o, int64_t o_row_stride,
                          int64_t o_head_stride, int32_t nq, int32_t H, int32_t HKV, int32_t D,
                          int32_t block_len, const int64_t* block_q, int64_t n_partials,
                          const int64_t* block_q_cnts, const int64_t* block_q_offset,
                          const int64_t* block_lens, int64_t n_blocks,
                          const int64_t* block_bitmasks, const int64_t* block_kv,
                          const deft_plan_t* plan, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * DeFT-Node / Node-Chunk / Tree-Index operator.
 * Replaces tree_attention_fwd (layers/attention/tree_attention.py:14-68: stage-1 :82-293 + stage 2).
 *   kv_indices: node_kv, int64 -- or the int32 node->page table of tree-index mode
 *               (tree_decoding/tree_index_pool.py:11-50); kv_index_bytes is 8 or 4
 *   kv_offset/kv_len/q_offset/q_len [n_entries] int64; node_q [n_partials] int64
 *   max_kv_len: upper bound of kv_len[] (sizes the split of long entries; e.g. the pool size)
 * ------------------------------------------------------------------------------------------ */
size_t deft_b200_node_workspace_bytes(int32_t nq, int32_t H, int32_t HKV, int32_t D,
                                      int64_t n_partials, int64_t n_entries, int64_t total_kv_bound,
                                      const deft_plan_t* plan);

int deft_b200_node_fwd(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                       const void* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                       int64_t kv_pool_tokens, void* o,
                       int64_t o_row_stride, int64_t o_head_stride, int32_t nq, int32_t H,
                       int32_t HKV, int32_t D, const void* kv_indices, int32_t kv_index_bytes,
                       const int64_t* kv_offset, const int64_t* kv_len, const int64_t* node_q,
                       int64_t n_partials, const int64_t* q_offset, const int64_t* q_len,
                       int64_t n_entries, int64_t total_kv_bound, const deft_plan_t* plan,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused KV append.  Inside a decode step the reference first scatters the step's K/V rows into the pool
 * (KVCacheUpdater.update, tree_decoding/tree_cache.py:67-76, called from deft_attention.py:121) and then attends over the
 * pool.  The *_append forms below do both in the attention's own two launches: wherever the plan marks a token as this
 * step's (tables built with `fresh_page`, see deft_b200_build_tables) stage 1 reads its K/V row straight from the
 * activations, and stage 2 writes the rows to their pages cache_loc[i] for the steps to come.  Pool contents after
 * the call are those of the reference's two index_puts.  Needs a host-built plan and the tensor-core path.
 *   new_k / new_v [nq, HKV, D] fp16 views of the fused qkv output (strides in elements), cache_loc [nq] int32 [dev]
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* new_k;
  const void* new_v;
  int64_t new_row_stride;
  int64_t new_head_stride;
  const int32_t* cache_loc;
} deft_append_t;

int deft_b200_flatten_fwd_append(const void* q, int64_t q_row_stride, int64_t q_head_stride,
                                 const void* k, const void* v, int64_t kv_tok_stride,
                                 int64_t kv_head_stride, int64_t kv_pool_tokens, void* o, int64_t o_row_stride,
                                 int64_t o_head_stride, int32_t nq, int32_t H, int32_t HKV, int32_t D,
                                 int32_t block_len, const int64_t* block_q, int64_t n_partials,
                                 const int64_t* block_q_cnts, const int64_t* block_q_offset,
                                 const int64_t* block_lens, int64_t n_blocks,
                                 const int64_t* block_bitmasks, const int64_t* block_kv,
                                 const deft_plan_t* plan, const deft_append_t* append, void* workspace,
                                 size_t workspace_bytes, void* stream);

int deft_b200_node_fwd_append(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                              const void* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                              int64_t kv_pool_tokens, void* o,
                              int64_t o_row_stride, int64_t o_head_stride, int32_t nq, int32_t H,
                              int32_t HKV, int32_t D, const void* kv_indices, int32_t kv_index_bytes,
                              const int64_t* kv_offset, const int64_t* kv_len, const int64_t* node_q,
                              int64_t n_partials, const int64_t* q_offset, const int64_t* q_len,
                              int64_t n_entries, int64_t total_kv_bound, const deft_plan_t* plan,
                              const deft_append_t* append, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * KV append.  Replaces KVCacheUpdater.update (tree_decoding/tree_cache.py:67-76):
 *   key_buffer[cache_loc] = cache_k; value_buffer[cache_loc] = cache_v   in ONE launch.
 *   new_k/new_v [n, HKV, D] fp16 (row stride in elements), cache_loc [n] int32.
 * ------------------------------------------------------------------------------------------ */
int deft_b200_kv_append(void* k, void* v, int64_t kv_tok_stride, int64_t kv_head_stride,
                        const void* new_k, const void* new_v, int64_t new_row_stride,
                        int64_t new_head_stride, const int32_t* cache_loc, int32_t n, int32_t HKV,
                        int32_t D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Metadata builder (host only, no CUDA, thread-safe).
 * Replaces TreeMetadata.from_tree_cache (tree_decoding/tree_cache.py:618-881) and
 * from_tree_cache_node (:883-1018): KV-guided grouping + flattened-tree KV split.  Tables are
 * bit-identical to the reference's; the work plan above is produced in the same pass.
 *
 * The tree is handed over as flat arrays, nodes in DFS pre-order with children in creation
 * (dict insertion) order -- the order tree_cache.py:725-791 visits them:
 *   parent[n]        index of the parent in this order, -1 for the root (node 0).  Further -1 entries start
 *                    further trees: a forest of independent trees over one pool is built in one pass and
 *                    attended in one launch (query ids and page ids are global to the forest)
 *   kv_off[n+1], kv  per-node page lists (node.kv_indices, any order; sorted inside, :736)
 *   q_off[n+1], qs   per-node attending queries = rank by leaf id of node.refs (:650-652, :737)
 *   tix_row[n]       tree-index mode only: node.node_indices_id, else NULL
 * hkv / n_ctas: kv-head count and CTA count (SMs) the native unit plan is balanced for: the
 * builder cuts long KV chains into pieces and assigns (unit, kv-head) jobs to CTAs (longest first).
 * fresh_page [query_num] or NULL: the page this decode step appends for every query (TreeCache.alloc's cache_loc,
 *   -1: none).  The native plan then reads those tokens from the step's activations (deft_b200_*_fwd_append).
 * The result is one packed host buffer (upload with a single copy) + a directory.
 * ------------------------------------------------------------------------------------------ */
enum {
  DEFT_T_NODE_Q = 0, DEFT_T_NODE_KV, DEFT_T_NODE_Q_LEN, DEFT_T_NODE_KV_LEN, DEFT_T_NODE_Q_OFFSET,
  DEFT_T_NODE_KV_OFFSET, DEFT_T_BLOCK_Q, DEFT_T_BLOCK_Q_CNTS, DEFT_T_BLOCK_Q_OFFSET,
  DEFT_T_BLOCK_BITMASKS, DEFT_T_BLOCK_KV, DEFT_T_BLOCK_LENS,          /* int64 reference tables */
  DEFT_T_FLAT_ITEMS, DEFT_T_FLAT_GROUPS, DEFT_T_FLAT_CSR_OFF, DEFT_T_FLAT_CSR_ROWS, /* Flatten plan */
  DEFT_T_NODE_ITEMS, DEFT_T_NODE_GROUPS, DEFT_T_NODE_CSR_OFF, DEFT_T_NODE_CSR_ROWS, /* Node plan */
  DEFT_T_U_UNITS, DEFT_T_U_CSR_OFF, DEFT_T_U_CSR_ROWS, DEFT_T_U_KV, DEFT_T_U_MASK, DEFT_T_U_Q,
  DEFT_T_U_JOB_OFF, DEFT_T_U_JOBS, DEFT_T_U_BLK,                      /* native unit plan */
  DEFT_T_COUNT
};

typedef struct deft_tables deft_tables_t;

/* Capacity-padded packing.  A decode step appends one page per leaf (tree_generate.py:109), so every table grows a
 * little from step to step; packed tightly, their offsets move and whatever holds device addresses of the tables (a
 * captured CUDA graph of the step) dies.  A layout handle remembers a capacity per table: a table that fits keeps its
 * offset, one that outgrows its region takes 50-100 % more than it needs and moves deft_b200_layout_version() on.  NULL
 * packs tightly.  One handle per decode loop, used by one thread at a time. */
typedef struct deft_layout deft_layout_t;
deft_layout_t* deft_b200_layout_new(void);
void deft_b200_layout_free(deft_layout_t* layout);
int64_t deft_b200_layout_version(const deft_layout_t* layout);
/* on != 0 (set before the first build with the handle): builds leave the twelve reference tables and the item / group
 * plans EMPTY and make the native unit plan only -- all that deft_b200_*_fwd(_append) read on the tensor-core path
 * when they are given the plan (their table pointers may then be NULL, their counts 0).  For decode loops that call
 * the operators themselves (DecodeStepGraph); what the reference's own code reads from TreeMetadata is not there. */
void deft_b200_layout_set_native_only(deft_layout_t* layout, int on);

deft_tables_t* deft_b200_build_tables(int32_t n_nodes, const int32_t* parent, const int64_t* kv_off,
                                      const int64_t* kv, const int64_t* q_off, const int64_t* qs,
                                      const int64_t* tix_row, int64_t tix_max_ctx,
                                      int32_t query_num, int32_t block_len, int32_t max_q_len,
                                      int32_t max_block_len, int32_t node_split, int32_t hkv,
                                      int32_t n_ctas, deft_layout_t* layout, const int32_t* fresh_page);

/* Native tree mirror: SURVEY.md 8(f).1 "keep topology / page lists in C++".  What TreeCache holds as Python objects
 * (tree_cache.py:94-129: TreeNode.kv_indices, .refs, .children) is mirrored here once (deft_b200_tree_set, the flat arrays
 * of deft_b200_build_tables for ONE tree), after which a decode step only reports the page it appended to each leaf
 * (TreeCache.alloc, tree_cache.py:261-283 -> deft_b200_tree_append: node = DFS index, -1 entries are skipped) and the
 * tables are built straight from the mirrors (deft_b200_build_tables_trees: the trees one after the other, queries of
 * tree t after those of tree t-1; fresh_page indexed by that global query id).  Any structural change (branch, cut,
 * merge) is followed by another deft_b200_tree_set.  One thread at a time per handle. */
typedef struct deft_tree deft_tree_t;
deft_tree_t* deft_b200_tree_new(void);
void deft_b200_tree_free(deft_tree_t* tree);
int deft_b200_tree_set(deft_tree_t* tree, int32_t n_nodes, const int32_t* parent, const int64_t* kv_off, const int64_t* kv,
                       const int64_t* q_off, const int64_t* qs, const int64_t* tix_row /* or NULL */, int32_t query_num);
int deft_b200_tree_append(deft_tree_t* tree, int32_t n, const int32_t* node, const int64_t* page);
int64_t deft_b200_tree_pages(const deft_tree_t* tree);   /* pages held (consistency check of the caller) */
deft_tables_t* deft_b200_build_tables_trees(deft_tree_t* const* trees, int32_t n_trees, int64_t tix_max_ctx,
                                            int32_t block_len, int32_t max_q_len, int32_t max_block_len,
                                            int32_t node_split, int32_t hkv, int32_t n_ctas, deft_layout_t* layout,
                                            const int32_t* fresh_page);
const void* deft_b200_tables_data(const deft_tables_t* t);  /* packed host buffer */
size_t deft_b200_tables_bytes(const deft_tables_t* t);
/* dir[2*i] = byte offset of array i in the packed buffer, dir[2*i+1] = element count */
int deft_b200_tables_directory(const deft_tables_t* t, int64_t* dir /* [2*DEFT_T_COUNT] */);
/* scalars: {query_num, node_num, total_kv_len, block_len, flat_part_rows, node_part_rows,
 *           n_unit_slots, n_ctas, paired (1: the job lists are pair-aligned, see deft_job_t.shared),
 *           unit-slot capacity (= n_unit_slots without a layout handle: what the workspace is carved for),
 *           fresh (1: built with fresh_page)} */
int deft_b200_tables_scalars(const deft_tables_t* t, int64_t* out /* [11] */);
void deft_b200_tables_free(deft_tables_t* t);

#ifdef __cplusplus
}
#endif
#endif /* DEFT_B200_H_ */
