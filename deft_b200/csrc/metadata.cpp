// Host metadata builder: KV-guided grouping + flattened-tree KV split, and the native work plan.
//
// Replaces TreeMetadata.from_tree_cache (DeFT/deft/tree_decoding/tree_cache.py:618-881) and
// from_tree_cache_node (:883-1018).  The reference walks the tree in Python and uploads eleven
// tensors one by one (2-7 ms per decode step in its own logs); this builder makes one pass in C++
// and lays every table -- the twelve int64 reference tables, bit-identical, plus the item / group /
// CSR plan of include/deft_b200.h -- into ONE packed buffer that the caller uploads with one copy.
// No CUDA here: pure host code, re-entrant.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <queue>
#include <utility>
#include <vector>

#include "deft_b200.h"

namespace deft {
void set_error(const char* fmt, ...);
}

namespace {

using i64 = int64_t;
using i32 = int32_t;

struct Csr {
  std::vector<i32> off, rows;
};

// q -> ascending partial rows, from the group list
Csr make_csr(const std::vector<deft_group_t>& groups, const std::vector<i64>& q_list, i32 nq) {
  Csr c;
  c.off.assign((size_t)nq + 1, 0);
  size_t total = 0;
  for (const auto& g : groups)
    for (i32 r = 0; r < g.q_cnt; ++r) {
      const i64 q = q_list[(size_t)g.q_off + r];
      if (q >= 0 && q < nq) { ++c.off[(size_t)q + 1]; ++total; }
    }
  for (i32 q = 0; q < nq; ++q) c.off[(size_t)q + 1] += c.off[(size_t)q];
  c.rows.assign(total, 0);
  std::vector<i32> cur(c.off.begin(), c.off.end() - 1);
  for (const auto& g : groups)  // part_base grows with the group index, so rows come out ascending
    for (i32 r = 0; r < g.q_cnt; ++r) {
      const i64 q = q_list[(size_t)g.q_off + r];
      if (q >= 0 && q < nq) c.rows[(size_t)cur[(size_t)q]++] = g.part_base + r;
    }
  for (i32 q = 0; q < nq; ++q) std::sort(c.rows.begin() + c.off[(size_t)q], c.rows.begin() + c.off[(size_t)q + 1]);
  return c;
}

std::vector<i64> offsets_of(const std::vector<i64>& lens) {  // cat([0], cumsum(len)[:-1]), tree_cache.py:822-833
  std::vector<i64> o(lens.size(), 0);
  for (size_t i = 1; i < lens.size(); ++i) o[i] = o[i - 1] + lens[i - 1];
  return o;
}

}  // namespace

struct deft_tables {
  std::vector<unsigned char> packed;
  i64 dir[2 * DEFT_T_COUNT];
  i64 scalars[11];
};

// Capacities of the packed buffer's regions, kept from one build to the next (deft_b200_layout_new): a table that
// fits its region keeps its offset, so the device addresses a captured CUDA graph was given stay valid while the
// tree grows.  A table that outgrows its region takes ~25 % more than it needs and moves the version on.
struct deft_layout {
  i64 cap_bytes[DEFT_T_COUNT] = {0};
  i64 slot_cap = 0;      // unit slots (partial tiles per kv-head) the workspace is carved for
  i64 version = 0;
  int native_only = 0;   // leave the reference tables and the item / group plans empty (deft_b200_layout_set_native_only)
};

// Native mirror of one decoding tree (deft_b200_tree_new): topology and per-node page lists kept on this side of the
// ABI, so that a decode step only hands over the pages it appended (TreeCache.alloc) instead of the whole tree.
// What the native tiler keeps per tree from one build to the next (valid while the tree only grows by appends): the
// token order of its tiles.  A build then lays out only the tokens appended since -- its cost no longer grows with
// the tree -- and cuts the tiles from the kept order.
struct deft_tiler_cache {
  bool valid = false;
  i32 query_base = 0;                      // rank of the tree's first query in the build the order was made for
  std::vector<i32> run_page, run_node;     // tokens of the RUN tiles (128 each), nodes tree-local
  struct Group {                           // tokens attended by one set of slots, in tile order (-1: dummy)
    i64 sig;
    i32 max_page;
    std::vector<i32> page, node;
  };
  std::vector<Group> groups;               // ascending sig
  std::vector<std::pair<i32, i64>> pending;   // (node, page) appended since the kept order was last brought up to date
};

struct deft_tree {
  std::vector<i32> parent;                 // DFS pre-order, -1 for the root
  std::vector<std::vector<i64>> pages;     // per node, in the order they were handed out
  std::vector<i64> q_off, qs, tix;
  i32 query_num = 0;
  i64 n_pages = 0;
  deft_tiler_cache tiler;
};

namespace {
template <class T>
struct Span {   // a stretch of one of the pooled arrays below
  T* p = nullptr;
  size_t n = 0;
  T* begin() const { return p; }
  T* end() const { return p + n; }
  T* data() const { return p; }
  size_t size() const { return n; }
  T& operator[](size_t i) const { return p[i]; }
};
// native tile (one per 128 token slots, KV NOT duplicated per 32-query sub-block); the arrays live in the Scratch pools
struct Tile {
  i32 n_live;
  size_t s0;                 // first entry of the tile in the pools
  Span<i32> slots;           // query slots with at least one attending row, ascending
  Span<uint32_t> masks;      // [touched slot][128]: bit r = rank 32*slot + r attends the token
  Span<uint32_t> rows_or;    // [touched slot]: OR of the slot's words over the live tokens
  Span<uint8_t> dense;       // [touched slot]: 128 live tokens, every row of the slot attends every one
};
struct RestTok { uint64_t key; i32 page, node; };   // a token outside the RUN tiles; key = (group of slots, page)
constexpr i32 kFreshToken = 1 << 30;   // u_kv entry of a token of THIS step: the bit + its query id (deft_plan_t.u_kv)
// Storage kept from one call to the next (per thread): the tables of a decode step are a few hundred KB, and fresh
// allocations of that size are handed back to the kernel at every free and page-faulted in again at the next build
// (measured: a third of the build time).  Everything is cleared, nothing is read, at the start of a build.
struct Scratch {
  std::vector<i64> node_q, node_kv, node_q_len, node_kv_len, node_kv_offset_ti;
  std::vector<i64> block_q, block_q_cnts, block_kv, block_masks, block_lens;
  std::vector<i32> u_kv, u_blk;             // native tiles: page id per token slot, load descriptor per chunk of 8 slots
  std::vector<i32> tok_page, tok_node;      // the trees' tokens in DFS order (page id, node)
  std::vector<i32> nw_off, nw_slot;         // per node: the slots its queries sit in ...
  std::vector<uint32_t> nw_word;            // ... and their rows in each
  std::vector<i64> node_sig;
  std::vector<uint32_t> u_mask;
  std::vector<i64> kvs;
  std::vector<Tile> tiles;                  // native tiles and their pooled arrays (one entry per (tile, touched slot))
  std::vector<i32> tp_slots;
  std::vector<uint32_t> tp_masks, tp_rows_or;
  std::vector<uint8_t> tp_dense;
  std::vector<RestTok> rest, rest_tmp;      // the native tiler's subtree tokens (radix-sorted by group and page)
  std::vector<i32> out_page, out_node;
  deft_tables* spare = nullptr;   // a freed handle whose packed buffer is reused by the next build
  // the last plan search: the piece length depends only on the chains' shape, which most decode steps keep
  std::vector<uint64_t> plan_key;
  double plan_len = 0.0;
  ~Scratch() { delete spare; }
};
thread_local Scratch g_scratch;
// the mirrors the flat arrays of the running build were made from, tree by tree (deft_b200_build_tables_trees), or null
thread_local deft_tree_t* const* g_tree_hints = nullptr;
thread_local i32 g_n_tree_hints = 0;
}  // namespace

extern "C" {

deft_tables_t* deft_b200_build_tables(int32_t n_nodes, const int32_t* parent, const int64_t* kv_off,
                                      const int64_t* kv, const int64_t* q_off, const int64_t* qs,
                                      const int64_t* tix_row, int64_t tix_max_ctx,
                                      int32_t query_num, int32_t block_len, int32_t max_q_len,
                                      int32_t max_block_len, int32_t node_split, int32_t hkv,
                                      int32_t n_ctas, deft_layout_t* layout, const int32_t* fresh_page) {
  if (n_nodes <= 0 || !parent || !kv_off || !kv || !q_off || !qs) {
    deft::set_error("build_tables: null or empty tree");
    return nullptr;
  }
  if (block_len <= 0 || max_q_len <= 0 || max_q_len > 32 || query_num <= 0) {
    deft::set_error("build_tables: need block_len > 0, 0 < max_q_len <= 32, query_num > 0");
    return nullptr;
  }
  if (parent[0] != -1) {
    deft::set_error("build_tables: node 0 must be the root (parent -1)");
    return nullptr;
  }
  // Further nodes with parent -1 start further trees: a FOREST of independent decoding trees over one page
  // pool (batched serving; the reference handles one tree per call).  The tables are what the reference's
  // DFS would emit walking the trees one after the other.
  for (i32 i = 1; i < n_nodes; ++i)
    if (parent[i] < -1 || parent[i] >= i) {
      deft::set_error("build_tables: nodes must be in DFS pre-order (parent[%d] = %d)", i, parent[i]);
      return nullptr;
    }

  const char* prof_env = std::getenv("DEFT_BUILD_PROFILE");
  const bool prof = prof_env && prof_env[0] == '1';
  auto t_prof = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!prof) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "build_tables %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(now - t_prof).count());
    t_prof = now;
  };
  // A decode loop on the tensor-core kernels reads nothing but the native unit plan: its layout may ask for just that
  // (the int64 tables of the reference and the item / group plans of the warp-FMA path are then left empty)
  const bool native_only = layout && layout->native_only && block_len == 128 && hkv > 0;
  Scratch& S = g_scratch;
  auto& node_q = S.node_q; auto& node_kv = S.node_kv; auto& node_q_len = S.node_q_len; auto& node_kv_len = S.node_kv_len;
  auto& node_kv_offset_ti = S.node_kv_offset_ti;
  auto& block_q = S.block_q; auto& block_q_cnts = S.block_q_cnts; auto& block_kv = S.block_kv;
  auto& block_masks = S.block_masks; auto& block_lens = S.block_lens;
  node_q.clear(); node_kv.clear(); node_q_len.clear(); node_kv_len.clear(); node_kv_offset_ti.clear();
  block_q.clear(); block_q_cnts.clear(); block_kv.clear(); block_masks.clear(); block_lens.clear();
  std::vector<deft_item_t> f_items;
  std::vector<deft_group_t> f_groups;
  i64 total_kv_len = 0;

  lap("reference tables");
  // ---- native plan, part 1: fixed query slots.  Queries are ranked in DFS LEAF order (the order the
  // pre-order walk meets the childless nodes), so that every node's query set is a contiguous rank range;
  // slot k holds ranks [32k, 32k+32).  A tile touches few slots and consecutive tiles touch the same ones.
  std::vector<i32> rank_of((size_t)query_num, -1);
  {
    std::vector<char> has_child((size_t)n_nodes, 0);
    for (i32 n = 1; n < n_nodes; ++n)
      if (parent[n] >= 0) has_child[(size_t)parent[n]] = 1;
    i32 next = 0;
    for (i32 n = 0; n < n_nodes; ++n)
      if (!has_child[(size_t)n])
        for (i64 i = q_off[n]; i < q_off[n + 1]; ++i) {
          const i64 qv = qs[i];
          if (qv >= 0 && qv < query_num && rank_of[(size_t)qv] < 0) rank_of[(size_t)qv] = next++;
        }
    for (i32 qv = 0; qv < query_num; ++qv)
      if (rank_of[(size_t)qv] < 0) rank_of[(size_t)qv] = next++;
  }
  auto& tiles = S.tiles;
  auto& tp_slots = S.tp_slots; auto& tp_masks = S.tp_masks; auto& tp_rows_or = S.tp_rows_or; auto& tp_dense = S.tp_dense;
  tiles.clear(); tp_slots.clear(); tp_masks.clear(); tp_rows_or.clear(); tp_dense.clear();
  auto& u_kv = S.u_kv;
  auto& tok_page = S.tok_page;
  auto& tok_node = S.tok_node;
  u_kv.clear();
  tok_page.clear();
  tok_node.clear();

  // open block (tree_cache.py:654-658)
  std::vector<i64> seg_tokens;
  std::vector<i64> seg_lens;
  std::vector<std::vector<i64>> seg_qs;  // sorted query ids per segment
  std::vector<i64> uni;                  // union, kept sorted + unique at close time

  auto close_block = [&]() {  // pack_new_block, tree_cache.py:661-723
    const i64 n_live = (i64)seg_tokens.size();
    const size_t n_seg = seg_qs.size();
    const i64 n_pad = n_live < block_len ? block_len - n_live : 0;  // the pad segment has an empty query set
    if (n_seg > 1) std::sort(uni.begin(), uni.end());  // one segment: its sorted query list is the union already
    uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
    deft_item_t item{};
    item.kv_off = (i64)block_lens.size() * block_len;
    item.kv_len = (i32)n_live;
    item.grp_off = (i32)f_groups.size();
    for (size_t s0 = 0; s0 < uni.size(); s0 += (size_t)max_q_len) {
      const size_t s1 = std::min(uni.size(), s0 + (size_t)max_q_len);
      deft_group_t g{};
      g.mask_off = (i64)block_lens.size() * block_len;
      g.q_off = (i32)block_q.size();
      g.q_cnt = (i32)(s1 - s0);
      g.part_base = g.q_off;
      f_groups.push_back(g);
      block_q.insert(block_q.end(), uni.begin() + (long)s0, uni.begin() + (long)s1);
      block_q_cnts.push_back((i64)(s1 - s0));
      block_kv.insert(block_kv.end(), seg_tokens.begin(), seg_tokens.end());
      block_kv.insert(block_kv.end(), (size_t)n_pad, (i64)-1);
      block_lens.push_back(n_live);
      for (size_t s = 0; s < n_seg; ++s) {
        i64 bits = 0;
        if (n_seg == 1) {
          bits = (i64)(((uint64_t)1 << (s1 - s0)) - 1);   // the sub-list [s0, s1) IS the segment's query set (<= 32 of them)
        } else {
          for (i64 qv : seg_qs[s]) {
            // position of qv inside the sub-list [s0, s1)
            auto it = std::lower_bound(uni.begin() + (long)s0, uni.begin() + (long)s1, qv);
            if (it != uni.begin() + (long)s1 && *it == qv) bits |= (i64)1 << (it - (uni.begin() + (long)s0));
          }
        }
        block_masks.insert(block_masks.end(), (size_t)seg_lens[s], bits);
      }
      block_masks.insert(block_masks.end(), (size_t)n_pad, (i64)0);
    }
    item.n_grp = (i32)f_groups.size() - item.grp_off;
    item.cost = item.kv_len * item.n_grp;
    if (item.n_grp > 0) f_items.push_back(item);
    seg_tokens.clear(); seg_lens.clear(); seg_qs.clear(); uni.clear();
  };

  auto& kvs = S.kvs;
  std::vector<i64> q;
  std::vector<std::pair<i32, i32>> fresh_sorted;   // (page, query) of this step's tokens (fused KV append)
  if (fresh_page) {
    for (i32 qv = 0; qv < query_num; ++qv)
      if (fresh_page[qv] >= 0) fresh_sorted.emplace_back(fresh_page[qv], qv);
    std::sort(fresh_sorted.begin(), fresh_sorted.end());
  }
  const char* env_r = std::getenv("DEFT_PLAN_REGROUP");
  const bool regroup = fresh_page != nullptr || !(env_r && env_r[0] == '0');   // (fresh tokens need chunks of their own)
  // A native-only build of trees that all keep their tile order needs no token stream at all: the tokens of a tree
  // are then listed only if its kept order turns out not to serve (gen_tokens below).
  bool lazy_tokens = native_only && regroup && g_tree_hints != nullptr;
  if (lazy_tokens) {
    i32 qb = 0;
    for (i32 ti = 0; ti < g_n_tree_hints && lazy_tokens; ++ti) {
      lazy_tokens = g_tree_hints[ti]->tiler.valid && g_tree_hints[ti]->tiler.query_base == qb;
      qb += g_tree_hints[ti]->query_num;
    }
  }
  auto push_tokens = [&]() {   // of the node whose sorted pages are in `kvs`; this step's tokens name their query instead of a page
    for (i64 kvp : kvs) {
      i32 v = (i32)kvp;
      if (fresh_page && !fresh_sorted.empty() && kvp >= fresh_sorted.front().first && kvp <= fresh_sorted.back().first) {
        auto it = std::lower_bound(fresh_sorted.begin(), fresh_sorted.end(), std::make_pair((i32)kvp, (i32)-1));
        if (it != fresh_sorted.end() && it->first == (i32)kvp) v = kFreshToken | it->second;
      }
      tok_page.push_back(v);
    }
  };
  auto gen_tokens = [&](i32 n0, i32 n1) {
    tok_page.clear();
    tok_node.clear();
    for (i32 n = n0; n < n1; ++n) {
      kvs.assign(kv + kv_off[n], kv + kv_off[n + 1]);
      if (!std::is_sorted(kvs.begin(), kvs.end())) std::sort(kvs.begin(), kvs.end());
      push_tokens();
      tok_node.insert(tok_node.end(), kvs.size(), n);
    }
  };
  for (i32 n = 0; n < n_nodes; ++n) {  // pre-order visit == the reference's recursive dfs (:725-791)
    const i64 k0 = kv_off[n], k1 = kv_off[n + 1], q0 = q_off[n], q1 = q_off[n + 1];
    if (k1 < k0 || q1 <= q0) {
      deft::set_error("build_tables: node %d has no attending query or a negative page count", n);
      return nullptr;
    }
    if (lazy_tokens) {
      for (i64 i = q0; i < q1; ++i)
        if (qs[i] < 0 || qs[i] >= query_num) {
          deft::set_error("build_tables: node %d lists query %lld outside [0, %d)", n, (long long)qs[i], query_num);
          return nullptr;
        }
      total_kv_len += k1 - k0;
      continue;
    }
    kvs.assign(kv + k0, kv + k1);
    const i64 n_kv = (i64)kvs.size();
    q.assign(qs + q0, qs + q1);
    std::sort(q.begin(), q.end());
    for (i64 qv : q)
      if (qv < 0 || qv >= query_num) {
        deft::set_error("build_tables: node %d lists query %lld outside [0, %d)", n, (long long)qv, query_num);
        return nullptr;
      }
    total_kv_len += n_kv;
    if (max_block_len == -1 && n_kv == 0) {
      // the reference raises here (range() step 0, tree_cache.py:746-748); report instead of crashing
      deft::set_error("build_tables: node %d has no KV pages (call alloc() before building tables)", n);
      return nullptr;
    }
    const i64 step = max_block_len == -1 ? n_kv : max_block_len;
    const bool kv_sorted = std::is_sorted(kvs.begin(), kvs.end());  // pages of a node are handed out ascending
    if (!tix_row && !kv_sorted) std::sort(kvs.begin(), kvs.end());  // :736 (the tree-index table keeps page order)
    for (size_t s0 = 0; s0 < q.size() && !native_only; s0 += (size_t)max_q_len) {
      const size_t s1 = std::min(q.size(), s0 + (size_t)max_q_len);
      for (i64 c0 = 0; c0 < n_kv; c0 += step) {
        const i64 c1 = std::min(n_kv, c0 + step);
        node_q.insert(node_q.end(), q.begin() + (long)s0, q.begin() + (long)s1);
        node_q_len.push_back((i64)(s1 - s0));
        node_kv_len.push_back(c1 - c0);
        if (tix_row) node_kv_offset_ti.push_back(tix_row[n] * tix_max_ctx + c0);  // tree_index_pool.py:47-49
        else node_kv.insert(node_kv.end(), kvs.begin() + c0, kvs.begin() + c1);
      }
    }
    if (tix_row && !kv_sorted) std::sort(kvs.begin(), kvs.end());
    push_tokens();               // the native tiler's token stream (DFS order)
    tok_node.insert(tok_node.end(), kvs.size(), n);
    // flatten packing (:763-788)
    i64 room = block_len - (i64)seg_tokens.size();
    i64 done = native_only ? n_kv : 0;
    while (done < n_kv) {
      if (n_kv - done < room) {
        seg_tokens.insert(seg_tokens.end(), kvs.begin() + done, kvs.end());
        seg_lens.push_back(n_kv - done);
        seg_qs.push_back(q);
        uni.insert(uni.end(), q.begin(), q.end());
        break;
      }
      seg_tokens.insert(seg_tokens.end(), kvs.begin() + done, kvs.begin() + done + room);
      seg_lens.push_back(room);
      seg_qs.push_back(q);
      uni.insert(uni.end(), q.begin(), q.end());
      close_block();
      done += room;
      room = block_len;
    }
  }
  if (!seg_lens.empty()) close_block();  // :797-798

  std::vector<i64> node_q_offset = offsets_of(node_q_len);
  std::vector<i64> node_kv_offset = tix_row ? node_kv_offset_ti : offsets_of(node_kv_len);
  std::vector<i64> block_q_offset = offsets_of(block_q_cnts);

  lap("dfs + reference tables");
  // ---- native plan, part 1b: the native tiles.
  // The reference cuts the DFS-ordered token stream of the whole table into blocks of 128 (block_kv above, bit for bit).
  // The native tiles are cut from the same tokens, tree by tree, with two freedoms the kernel's arithmetic allows (a
  // softmax does not care about the order of its tokens, and the per-token masks travel with them):
  //  * every tree starts a new tile, and a stretch of >= 128 consecutive pages attended by one set of slots (a prompt)
  //    becomes whole RUN tiles from its first token on -- four TMA boxes per operand, no mask, and chains that do not
  //    straddle trees (a forest's trees are independent jobs);
  //  * what is left (the subtree) is grouped by the set of slots that attend it, so that a slot walks whole tiles of
  //    its own, and inside a group laid out as aligned chunks of 32 / 16 / 8 / 4 / 2 / 1 consecutive pages, longest
  //    first.  The kernel loads a tile as blocks of 32 rows: 32 consecutive pages are one TMA box per panel, aligned
  //    runs of 16 or 8 smaller boxes, the rest goes four rows at a time (TMA gather4) -- and the SM's copy engine is
  //    bound by the NUMBER of such instructions: 2.2 us per K + V tile of gathered rows against 0.6 us for boxes, more
  //    than the softmax and the tensor pipe need for the tile.  The allocator hands one decode step's pages to the
  //    leaves in ascending order (tree_cache.py:261-283), so the tokens of one step over neighbouring leaves DO sit on
  //    consecutive pages; the DFS order just strings them node by node.  A group is padded to whole blocks with dummy
  //    tokens (page -1: a zero row nobody attends).  DEFT_PLAN_REGROUP=0: the rest keeps its DFS order.
  {
    // per node: the words of its attending ranks, slot by slot (CSR over the nodes), and the slots it touches as a key
    auto& nw_off = S.nw_off; auto& nw_slot = S.nw_slot; auto& nw_word = S.nw_word; auto& node_sig = S.node_sig;
    nw_off.assign((size_t)n_nodes + 1, 0);
    nw_slot.clear();
    nw_word.clear();
    node_sig.assign((size_t)n_nodes, -1);
    for (i32 n = 0; n < n_nodes; ++n) {
      const size_t w0 = nw_slot.size();
      for (i64 i = q_off[n]; i < q_off[n + 1]; ++i) {
        const i32 rk = rank_of[(size_t)qs[i]];
        size_t k = w0;
        while (k < nw_slot.size() && nw_slot[k] != rk / 32) ++k;
        if (k == nw_slot.size()) { nw_slot.push_back(rk / 32); nw_word.push_back(0u); }
        nw_word[k] |= 1u << (rk % 32);
      }
      for (size_t i = w0 + 1; i < nw_slot.size(); ++i)          // (a node touches one or two slots, rarely more: insertion sort)
        for (size_t j = i; j > w0 && nw_slot[j] < nw_slot[j - 1]; --j) { std::swap(nw_slot[j], nw_slot[j - 1]); std::swap(nw_word[j], nw_word[j - 1]); }
      nw_off[(size_t)n + 1] = (i32)nw_slot.size();
      if (nw_slot.size() > w0) node_sig[(size_t)n] = ((i64)nw_slot[w0] << 32) | (i64)nw_slot.back();
    }
    auto sig_of = [&](i32 n) -> i64 { return n < 0 ? -1 : node_sig[(size_t)n]; };
    // one native tile from up to 128 (page, node) tokens; node < 0: a dummy token
    auto emit_tile = [&](const i32* pages, const i32* nodes, i32 n_live) {
      tiles.emplace_back();
      Tile& tl = tiles.back();
      tl.n_live = n_live;
      const size_t s0 = tl.s0 = tp_slots.size();
      const size_t k0 = u_kv.size();
      u_kv.resize(k0 + 128, 0);
      std::copy(pages, pages + n_live, u_kv.begin() + (long)k0);
      i32 last_node = -2;
      for (i32 i = 0; i < n_live; ++i) {
        if (nodes[i] == last_node || nodes[i] < 0) continue;
        last_node = nodes[i];
        for (i32 k = nw_off[(size_t)last_node]; k < nw_off[(size_t)last_node + 1]; ++k)
          if (std::find(tp_slots.begin() + (long)s0, tp_slots.end(), nw_slot[(size_t)k]) == tp_slots.end()) tp_slots.push_back(nw_slot[(size_t)k]);
      }
      std::sort(tp_slots.begin() + (long)s0, tp_slots.end());
      const size_t ns = tp_slots.size() - s0;
      tp_masks.resize((s0 + ns) * 128, 0u);
      tp_rows_or.resize(s0 + ns, 0u);
      tp_dense.resize(s0 + ns, n_live == 128 ? 1 : 0);
      const i32* slots = tp_slots.data() + s0;
      uint32_t* masks = tp_masks.data() + s0 * 128;
      uint32_t* rows_or = tp_rows_or.data() + s0;
      uint32_t full[8], all_and[8];    // per touched slot (a tile touches a handful): the slot's live rows, AND over the tokens
      std::vector<uint32_t> full_v, and_v;
      uint32_t* fullp = full; uint32_t* andp = all_and;
      if (ns > 8) { full_v.resize(ns); and_v.resize(ns); fullp = full_v.data(); andp = and_v.data(); }
      for (size_t si = 0; si < ns; ++si) {
        const i32 cnt = std::min(32, query_num - 32 * slots[si]);
        fullp[si] = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
        andp[si] = 0xffffffffu;
      }
      for (i32 i = 0; i < n_live;) {    // tokens of one node in a row share their words
        i32 j = i + 1;
        while (j < n_live && nodes[j] == nodes[i]) ++j;
        size_t si = 0;
        if (nodes[i] >= 0)
          for (i32 k = nw_off[(size_t)nodes[i]]; k < nw_off[(size_t)nodes[i] + 1]; ++k) {
            while (slots[si] != nw_slot[(size_t)k]) { andp[si] = 0; ++si; }   // slots the node does not touch
            const uint32_t word = nw_word[(size_t)k];
            std::fill(masks + si * 128 + (size_t)i, masks + si * 128 + (size_t)j, word);
            rows_or[si] |= word;
            andp[si] &= word;
            ++si;
          }
        for (; si < ns; ++si) andp[si] = 0;
        i = j;
      }
      for (size_t si = 0; si < ns; ++si)
        if ((andp[si] & fullp[si]) != fullp[si]) tp_dense[s0 + si] = 0;
    };
    struct Chunk { i32 first, len; };       // `len` (a power of two) consecutive pages: tokens [first, first + len) of a sorted list
    auto& rest = S.rest; auto& rest_tmp = S.rest_tmp; auto& out_page = S.out_page; auto& out_node = S.out_node;
    std::vector<Chunk> chunks;
    std::vector<i64> sigs;                  // the distinct slot sets of one tree's subtree tokens (a handful)
    std::vector<i32> sig_rank;
    std::vector<i32> gseq_page, gseq_node;  // one group's tokens in tile order
    // power-of-two chunks of the runs of consecutive pages of tokens [g0, g1) of `rest` (sorted by page)
    auto chunk_runs = [&](size_t g0, size_t g1) {
      chunks.clear();
      for (size_t i = g0; i < g1;) {
        size_t j = i + 1;
        while (j < g1 && rest[j].page == rest[j - 1].page + 1) ++j;
        size_t at = i;
        for (i32 len = 32; len >= 1; len >>= 1)
          while (j - at >= (size_t)len) {
            chunks.push_back({(i32)at, len});
            at += (size_t)len;
          }
        i = j;
      }
    };
    // `rest` by ascending key: a byte-wise radix sort (LSD, bytes that are the same for all keys skipped: two or three
    // passes for a pool of < 16 M pages)
    auto sort_rest = [&]() {
      const size_t n_rest = rest.size();
      uint32_t hist[5][256] = {};
      uint64_t any = 0;
      for (const RestTok& tk : rest) {
        any |= tk.key;
        for (int d = 0; d < 5; ++d) ++hist[d][(tk.key >> (8 * d)) & 0xff];
      }
      if ((any >> 40) != 0) {             // (more slot sets than a byte counts: never seen; the comparison sort is the fallback)
        std::sort(rest.begin(), rest.end(), [](const RestTok& x, const RestTok& y) { return x.key < y.key; });
        return;
      }
      rest_tmp.resize(n_rest);
      RestTok* src = rest.data();
      RestTok* dst = rest_tmp.data();
      for (int d = 0; d < 5 && n_rest > 1; ++d) {
        uint32_t* h = hist[d];
        if (h[(src[0].key >> (8 * d)) & 0xff] == n_rest) continue;   // every key has this byte
        uint32_t sum = 0;
        for (int v = 0; v < 256; ++v) { const uint32_t c = h[v]; h[v] = sum; sum += c; }
        for (size_t x = 0; x < n_rest; ++x) dst[h[(src[x].key >> (8 * d)) & 0xff]++] = src[x];
        std::swap(src, dst);
      }
      if (src != rest.data()) rest.swap(rest_tmp);
    };
    // the groups of `rest` (sorted) one after the other into out_page / out_node: inside a group the chunks longest
    // first (in page order among equals), between groups dummies up to a whole block of 32.  `keep`: where the
    // groups without the fresh flag are also kept for the builds to come (nodes tree-local).
    auto layout_rest = [&](deft_tiler_cache* keep, i32 node_base) {
      const size_t n_rest = rest.size();
      for (size_t g0 = 0; g0 < n_rest;) {
        size_t g1 = g0;
        while (g1 < n_rest && (rest[g1].key >> 32) == (rest[g0].key >> 32)) ++g1;
        chunk_runs(g0, g1);
        gseq_page.clear();
        gseq_node.clear();
        for (i32 len = 32; len >= 1; len >>= 1)
          for (const Chunk& c : chunks) {
            if (c.len != len) continue;
            for (i32 k = 0; k < len; ++k) {
              gseq_page.push_back(rest[(size_t)(c.first + k)].page);
              gseq_node.push_back(rest[(size_t)(c.first + k)].node);
            }
          }
        out_page.insert(out_page.end(), gseq_page.begin(), gseq_page.end());
        out_node.insert(out_node.end(), gseq_node.begin(), gseq_node.end());
        const i64 sig = sigs[(size_t)sig_rank[(size_t)(rest[g0].key >> 32)]];
        if (keep && !((sig >> 62) & 1)) {
          keep->groups.emplace_back();
          deft_tiler_cache::Group& g = keep->groups.back();
          g.sig = sig;
          g.max_page = rest[g1 - 1].page;
          g.page = gseq_page;
          g.node.resize(gseq_node.size());
          for (size_t k = 0; k < gseq_node.size(); ++k) g.node[k] = gseq_node[k] - node_base;
        }
        if (g1 < n_rest)
          while (out_page.size() % 32) {                       // whole blocks per group: the next group's runs stay aligned
            out_page.push_back(-1);
            out_node.push_back(-1);
          }
        g0 = g1;
      }
    };
    // rest tokens carry the index of their slot set in `sigs`; the sort wants the RANK of the set (ascending sig)
    auto rank_sigs = [&]() {
      sig_rank.assign(sigs.size(), 0);       // sig_rank[rank] = index in sigs (after the loop below: inverted)
      std::vector<i32> rk(sigs.size(), 0);
      for (size_t x = 0; x < sigs.size(); ++x)
        for (size_t y = 0; y < sigs.size(); ++y) rk[x] += sigs[y] < sigs[x] ? 1 : 0;
      for (RestTok& tk : rest) tk.key = ((uint64_t)rk[(size_t)(tk.key >> 32)] << 32) | (tk.key & 0xffffffffu);
      for (size_t x = 0; x < sigs.size(); ++x) sig_rank[(size_t)rk[x]] = (i32)x;
    };
    auto fresh_query = [&](i64 page) -> i32 {     // the query whose page of this step `page` is, or -1
      if (!fresh_page || fresh_sorted.empty() || page < fresh_sorted.front().first || page > fresh_sorted.back().first) return -1;
      auto it = std::lower_bound(fresh_sorted.begin(), fresh_sorted.end(), std::make_pair((i32)page, (i32)-1));
      return it != fresh_sorted.end() && it->first == (i32)page ? it->second : -1;
    };
    // A tree whose tile order was kept (deft_tiler_cache): the tokens appended since go to the END of their groups, the
    // ones of this very step (fresh_page) into groups of their own as ever, and the tiles are cut from the kept order.
    // false: the kept order cannot serve this build (nothing has been touched).
    struct Incoming { i32 group, page, node; };
    std::vector<Incoming> incoming;
    std::vector<std::pair<i32, i64>> still_pending;
    auto serve_kept = [&](deft_tree_t* ht, i32 node_base, i32 query_base) -> bool {
      deft_tiler_cache& c = ht->tiler;
      if (!c.valid || c.query_base != query_base) return false;
      size_t named = 0, found = 0;
      if (fresh_page)
        for (i32 qv = query_base; qv < query_base + ht->query_num; ++qv) named += fresh_page[qv] >= 0 ? 1 : 0;
      incoming.clear();
      still_pending.clear();
      rest.clear();
      sigs.clear();
      for (const auto& pn : c.pending) {
        const i64 sig = sig_of(node_base + pn.first);
        const i32 fq = fresh_query(pn.second);
        if (fq >= 0) {                      // of this step: read from the activations, kept pending for the next build
          ++found;
          still_pending.push_back(pn);
          const i64 fsig = sig | ((i64)1 << 62);
          size_t gi = 0;
          while (gi < sigs.size() && sigs[gi] != fsig) ++gi;
          if (gi == sigs.size()) sigs.push_back(fsig);
          rest.push_back({((uint64_t)gi << 32) | (uint32_t)fq, kFreshToken | fq, node_base + pn.first});
          continue;
        }
        size_t g = 0;
        while (g < c.groups.size() && c.groups[g].sig != sig) ++g;
        if (g < c.groups.size() && pn.second <= (i64)c.groups[g].max_page) return false;   // not an append in page order
        if (pn.second >= ((i64)1 << 27)) return false;
        incoming.push_back({g < c.groups.size() ? (i32)g : -1, (i32)pn.second, pn.first});
      }
      if (found != named) return false;     // fresh_page names tokens the kept order already holds as pool pages
      for (Incoming& in : incoming)         // slot sets seen for the first time get (empty) groups, kept in ascending order
        if (in.group < 0) {
          const i64 sig = sig_of(node_base + in.node);
          size_t g = 0;
          while (g < c.groups.size() && c.groups[g].sig < sig) ++g;
          if (g == c.groups.size() || c.groups[g].sig != sig) {
            deft_tiler_cache::Group ng;
            ng.sig = sig;
            ng.max_page = -1;
            c.groups.insert(c.groups.begin() + (long)g, std::move(ng));
            for (Incoming& o : incoming)
              if (o.group >= (i32)g) ++o.group;
          }
          in.group = (i32)g;
        }
      // per group: the new tokens by page, runs -> power-of-two chunks, longest first, at the end of the group.  A chunk
      // of 8 or more is aligned with dummies when the whole batch is a multiple of it (the alignment then lasts: one
      // page per leaf and step), so that it loads as one TMA box per panel.
      std::sort(incoming.begin(), incoming.end(), [](const Incoming& x, const Incoming& y) { return x.group != y.group ? x.group < y.group : x.page < y.page; });
      std::vector<RestTok> batch;
      for (size_t i0 = 0; i0 < incoming.size();) {
        size_t i1 = i0;
        while (i1 < incoming.size() && incoming[i1].group == incoming[i0].group) ++i1;
        deft_tiler_cache::Group& g = c.groups[(size_t)incoming[i0].group];
        batch.clear();
        for (size_t i = i0; i < i1; ++i) batch.push_back({0, incoming[i].page, incoming[i].node});
        batch.swap(rest);                   // (chunk_runs reads `rest`)
        chunk_runs(0, rest.size());
        const size_t batch_len = rest.size();
        for (i32 len = 32; len >= 1; len >>= 1)
          for (const Chunk& ch : chunks) {
            if (ch.len != len) continue;
            if (len >= 8 && batch_len % (size_t)len == 0)
              while (g.page.size() % (size_t)len) { g.page.push_back(-1); g.node.push_back(-1); }
            for (i32 k = 0; k < len; ++k) {
              g.page.push_back(rest[(size_t)(ch.first + k)].page);
              g.node.push_back(rest[(size_t)(ch.first + k)].node);
            }
          }
        g.max_page = rest.back().page;
        batch.swap(rest);
        i0 = i1;
      }
      c.pending.swap(still_pending);
      // the tiles: RUN tiles, the kept groups, this step's groups
      out_page.clear();
      out_node.clear();
      for (size_t k = 0; k < c.run_page.size(); k += 128) {
        out_node.resize(128);
        for (size_t x = 0; x < 128; ++x) out_node[x] = c.run_node[k + x] + node_base;
        emit_tile(&c.run_page[k], out_node.data(), 128);
      }
      out_node.clear();
      rank_sigs();
      sort_rest();
      bool first = true;
      for (const deft_tiler_cache::Group& g : c.groups) {
        if (g.page.empty()) continue;
        if (!first)
          while (out_page.size() % 32) { out_page.push_back(-1); out_node.push_back(-1); }
        first = false;
        out_page.insert(out_page.end(), g.page.begin(), g.page.end());
        for (i32 nd : g.node) out_node.push_back(nd < 0 ? -1 : nd + node_base);
      }
      if (!rest.empty()) {
        if (!first)
          while (out_page.size() % 32) { out_page.push_back(-1); out_node.push_back(-1); }
        layout_rest(nullptr, node_base);
      }
      for (size_t k = 0; k < out_page.size(); k += 128)
        emit_tile(&out_page[k], &out_node[k], (i32)std::min<size_t>(128, out_page.size() - k));
      return true;
    };
    // the tiles of the tree whose tokens are [a, b) of the stream, laid out from scratch (and the order kept, if it has a mirror)
    auto layout_tree = [&](size_t a, size_t b, deft_tree_t* ht, i32 node_base, i32 query_base) {
      deft_tiler_cache* keep = nullptr;
      if (ht && regroup) {                   // this build's order is kept for the builds to come
        keep = &ht->tiler;
        *keep = deft_tiler_cache();
        keep->query_base = query_base;
      }
      rest.clear();
      out_page.clear();
      out_node.clear();
      sigs.clear();
      for (size_t i = a; i < b;) {
        // a stretch of consecutive pages attended by one set of slots
        // (this step's tokens are not in the pool yet: they never join a run of pool pages, and get groups of their own)
        const bool fresh = (tok_page[i] & kFreshToken) != 0;
        const i64 sig = sig_of(tok_node[i]) | (fresh ? (i64)1 << 62 : 0);
        size_t j = i + 1;
        while (j < b && tok_page[j] == tok_page[j - 1] + 1 && (tok_node[j] == tok_node[j - 1] || sig_of(tok_node[j]) == (sig & ~((i64)1 << 62)))) ++j;
        const size_t whole = fresh ? 0 : (j - i) / 128 * 128;
        for (size_t k = i; k < i + whole; k += 128) emit_tile(&tok_page[k], &tok_node[k], 128);   // RUN tiles
        if (keep) {
          for (size_t k = i; k < i + whole; ++k) {
            keep->run_page.push_back(tok_page[k]);
            keep->run_node.push_back(tok_node[k] - node_base);
          }
          if (fresh)                         // this step's tokens join the kept order at the next build
            for (size_t k = i; k < j; ++k)
              keep->pending.emplace_back(tok_node[k] - node_base, (i64)fresh_page[tok_page[k] & ~kFreshToken]);
        }
        if (i + whole < j) {
          size_t gi = 0;
          while (gi < sigs.size() && sigs[gi] != sig) ++gi;
          if (gi == sigs.size()) sigs.push_back(sig);
          // (a fresh token's "page" is its query id with a flag: the flag is in the group already)
          for (size_t k = i + whole; k < j; ++k)
            rest.push_back({((uint64_t)gi << 32) | (uint32_t)(tok_page[k] & ~kFreshToken), tok_page[k], tok_node[k]});
        }
        i = j;
      }
      if (regroup) {
        rank_sigs();
        sort_rest();
        layout_rest(keep, node_base);
        if (keep) {
          keep->valid = true;
          for (const RestTok& tk : rest)     // (a page that does not fit a load descriptor fails the build further down)
            if (!(tk.page & kFreshToken) && tk.page >= (1 << 27)) keep->valid = false;
        }
      } else {
        for (const RestTok& tk : rest) { out_page.push_back(tk.page); out_node.push_back(tk.node); }
      }
      for (size_t k = 0; k < out_page.size(); k += 128)
        emit_tile(&out_page[k], &out_node[k], (i32)std::min<size_t>(128, out_page.size() - k));
    };
    i32 tree_i = 0, query_base = 0, nodes_before = 0;
    if (lazy_tokens) {
      for (; tree_i < g_n_tree_hints; ++tree_i) {
        deft_tree_t* ht = g_tree_hints[tree_i];
        const i32 n_nodes_t = (i32)ht->parent.size();
        if (!serve_kept(ht, nodes_before, query_base)) {
          gen_tokens(nodes_before, nodes_before + n_nodes_t);
          layout_tree(0, tok_page.size(), ht, nodes_before, query_base);
        }
        nodes_before += n_nodes_t;
        query_base += ht->query_num;
      }
    }
    const size_t n_tok = lazy_tokens ? 0 : tok_page.size();
    for (size_t a = 0; a < n_tok; ++tree_i) {
      // the tokens [a, b) of one tree (its nodes are consecutive in the pre-order; parent -1 starts the next tree)
      deft_tree_t* ht = g_tree_hints && tree_i < g_n_tree_hints ? g_tree_hints[tree_i] : nullptr;
      size_t b = a + 1;
      if (ht) b = a + (size_t)ht->n_pages;
      else
        while (b < n_tok && !(tok_node[b] != tok_node[b - 1] && parent[tok_node[b]] == -1)) ++b;
      const i32 node_base = ht ? nodes_before : tok_node[a];
      if (ht) nodes_before += (i32)ht->parent.size();
      if (!(ht && regroup && serve_kept(ht, node_base, query_base))) layout_tree(a, b, ht, node_base, query_base);
      if (ht) query_base += ht->query_num;
      a = b;
    }
    for (size_t t = 0; t < tiles.size(); ++t) {   // the pools no longer move: hand every tile its stretches
      Tile& tl = tiles[t];
      const size_t ns = (t + 1 < tiles.size() ? tiles[t + 1].s0 : tp_slots.size()) - tl.s0;
      tl.slots = {tp_slots.data() + tl.s0, ns};
      tl.masks = {tp_masks.data() + tl.s0 * 128, ns * 128};
      tl.rows_or = {tp_rows_or.data() + tl.s0, ns};
      tl.dense = {tp_dense.data() + tl.s0, ns};
    }
  }

  lap("native tiler");
  // per chunk of 8 token slots: how the kernel's producers load it (deft_plan_t.u_blk)
  auto& u_blk = S.u_blk;
  u_blk.assign(tiles.size() * 16, 0);
  for (size_t t = 0; t < tiles.size(); ++t) {
    auto runs = [&](size_t from, size_t len) {   // `len` live tokens on consecutive pages
      if (from + len > (size_t)tiles[t].n_live || u_kv[t * 128 + from] < 0) return false;
      for (size_t kk = t * 128 + from + 1; kk < t * 128 + from + len; ++kk)
        if (u_kv[kk] != u_kv[kk - 1] + 1) return false;
      return true;
    };
    for (size_t c = 0; c < 16; ++c) {
      const i32 kind = runs(c / 4 * 32, 32) ? 3 : runs(c / 2 * 16, 16) ? 2 : runs(c * 8, 8) ? 1 : 0;
      const i32 page = u_kv[t * 128 + c * 8];
      const i32 fresh = page >= 0 && (page & kFreshToken) ? 1 : 0;
      if (page >= 0 && (page & ~kFreshToken) >= (1 << 27)) {
        deft::set_error("build_tables: page id %d does not fit the 27 bits of a load descriptor", page);
        return nullptr;
      }
      u_blk[t * 16 + c] = (kind << 28) | (fresh << 27) | (kind ? (page & ~kFreshToken) : 0);
    }
  }

  lap("slots + tiles + flat plan");
  // ---- Node plan: long entries are cut into node_split-token items
  std::vector<deft_item_t> n_items;
  std::vector<deft_group_t> n_groups;
  i32 n_rows = 0;
  for (size_t e = 0; e < node_q_len.size(); ++e) {
    const i32 kn = (i32)node_kv_len[e], qn = (i32)node_q_len[e];
    const i32 step = node_split > 0 ? node_split : std::max(kn, 1);
    const i32 nch = std::max(1, (kn + step - 1) / step);
    for (i32 j = 0; j < nch; ++j) {
      deft_item_t it{};
      it.kv_off = node_kv_offset[e] + (i64)j * step;
      it.kv_len = std::max(0, std::min(step, kn - j * step));
      it.grp_off = (i32)n_groups.size();
      it.n_grp = 1;
      it.cost = it.kv_len;
      n_items.push_back(it);
      deft_group_t g{};
      g.mask_off = -1;
      g.q_off = (i32)node_q_offset[e];
      g.q_cnt = qn;
      g.part_base = n_rows;
      n_groups.push_back(g);
      n_rows += qn;
    }
  }
  Csr f_csr = make_csr(f_groups, block_q, query_num);
  Csr n_csr = make_csr(n_groups, node_q, query_num);

  lap("node plan");
  // ---- native plan, part 2: units.  For every PAIR of slots, the tiles touching it form chains of
  // consecutive tiles (the prompt plus the pair's own subtree are contiguous in DFS order); every chain
  // is cut into pieces, one unit per piece.  All units of a pair share their Q tiles.  The piece length
  // and the (unit, kv-head) -> CTA assignment come from a longest-first balance over n_ctas CTAs with
  // the cost model below (tile steps; calibrated on B200).
  std::vector<deft_unit_t> units;
  std::vector<i32> u_q, u_job_off;
  std::vector<deft_job_t> u_jobs;
  bool plan_paired = false;  // the job lists pair slot-jobs on CTAs (2c, 2c + 1): launch as clusters of two
  auto& u_mask = S.u_mask;
  u_mask.clear();
  Csr u_csr;
  u_csr.off.assign((size_t)query_num + 1, 0);
  i32 n_unit_slots = 0;
  if (block_len == 128 && !tiles.empty()) {
    const i32 n_slots = (query_num + 31) / 32, n_pairs = (n_slots + 1) / 2;
    u_q.assign((size_t)n_slots * 32, 0);
    for (i32 qv = 0; qv < query_num; ++qv) u_q[(size_t)rank_of[(size_t)qv]] = qv;
    auto slot_cnt = [&](i32 sl) { return std::min(32, query_num - 32 * sl); };
    auto tile_slot = [&](const Tile& tl, i32 sl) -> const uint32_t* {  // mask words of (tile, slot) or null
      auto it = std::lower_bound(tl.slots.begin(), tl.slots.end(), sl);
      return it != tl.slots.end() && *it == sl ? tl.masks.data() + (size_t)(it - tl.slots.begin()) * 128 : nullptr;
    };

    struct Chain { i32 pair; size_t t0, n_tiles; };
    std::vector<Chain> chains;
    // A chain also ends where the set of live slots changes (both -> one, one -> the other): a slot then never
    // walks tiles none of its rows attends.  (DEFT_PLAN_SPLIT_LIVE=0: chains by contiguity only.)
    const char* env_s = std::getenv("DEFT_PLAN_SPLIT_LIVE");
    const bool split_live = !(env_s && env_s[0] == '0');
    {
      // one pass over the tiles, every tile visiting only the pairs it touches (a forest has as many pairs as trees
      // and every tile belongs to one of them): a pair's run ends at a gap or, with split_live, where its set of
      // live slots changes
      struct Run { size_t run0 = 0, run_n = 0, last = 0; int sig = 0; };
      std::vector<Run> runs((size_t)n_pairs);
      auto close_run = [&](i32 pr) {
        Run& r = runs[(size_t)pr];
        if (r.run_n) chains.push_back({pr, r.run0, r.run_n});
        r.run_n = 0;
      };
      for (size_t t = 0; t < tiles.size(); ++t) {
        const Span<i32>& sl = tiles[t].slots;  // ascending
        for (size_t i = 0; i < sl.size();) {
          const i32 pr = sl[i] / 2;
          int sig = 0;
          for (; i < sl.size() && sl[i] / 2 == pr; ++i) sig |= (sl[i] & 1) ? 2 : 1;
          Run& r = runs[(size_t)pr];
          if (r.run_n && (r.last + 1 != t || (split_live && sig != r.sig))) close_run(pr);
          if (r.run_n == 0) { r.run0 = t; r.sig = sig; }
          ++r.run_n;
          r.last = t;
        }
      }
      for (i32 pr = 0; pr < n_pairs; ++pr) close_run(pr);
      std::sort(chains.begin(), chains.end(), [](const Chain& a, const Chain& b) { return a.pair != b.pair ? a.pair < b.pair : a.t0 < b.t0; });
    }

    const i32 heads = hkv > 0 ? hkv : 1;
    const i32 ctas = n_ctas > 0 ? n_ctas : 148;
    // one job = one SLOT of a unit on one kv-head (the kernel works one M = 128 accumulator per CTA).  Costs in
    // tile steps (calibrated on B200): a tile whose 128 pages are consecutive arrives as TMA boxes, a tile of
    // scattered pages is gathered row by row and is bound by that (~1.5x); the constant is a job's start-up
    // + epilogue.
    const char* env_g = std::getenv("DEFT_PLAN_GATHER_COST");
    const double gather_cost = env_g ? std::atof(env_g) : 2.2;
    const char* env_m = std::getenv("DEFT_PLAN_MASK_COST");
    const double mask_cost = env_m ? std::atof(env_m) : 0.5;
    std::vector<double> tile_cost(tiles.size());
    for (size_t t = 0; t < tiles.size(); ++t) {
      // what a tile costs the copy engine goes with its TMA instructions per panel: one box per aligned run of 32 / 16 / 8
      // consecutive pages, a gather4 per four rows of anything else (the gather cost above: a tile of nothing else)
      double scattered = 0.0;
      for (size_t c = 0; c < 16; ++c) {
        const i32 kind = u_blk[t * 16 + c] >> 28;
        const double instr = kind == 3 ? 0.25 : kind == 2 ? 0.5 : kind == 1 ? 1.0 : 2.0;   // of this chunk
        scattered += (instr - 0.25) / (2.0 - 0.25) / 4.0;
      }
      // ... and a tile some row of which does not attend every token costs the softmax warps a pass over the mask
      bool masked = false;
      for (uint8_t d : tiles[t].dense) masked = masked || d == 0;
      tile_cost[t] = 1.0 + (masked ? mask_cost : 0.0) + (gather_cost - 1.0) * scattered / 4.0;
    }
    // (DEFT_PLAN_JOB_CONST / DEFT_PLAN_GATHER_COST: calibration overrides for A/B runs on one box)
    const char* env_c = std::getenv("DEFT_PLAN_JOB_CONST");
    const double kJobConst = env_c ? std::atof(env_c) : 2.0;
    // pieces of a chain for a given maximum piece cost: (first tile, count), near-equal costs
    std::vector<std::pair<size_t, size_t>> pcs;
    std::vector<double> pcs_cost;
    auto pieces_of = [&](const Chain& c, double max_cost) {
      pcs.clear();
      pcs_cost.clear();
      double total = 0.0;
      for (size_t t = 0; t < c.n_tiles; ++t) total += tile_cost[c.t0 + t];
      const size_t np = std::max<size_t>(1, (size_t)std::ceil(total / max_cost - 1e-9));
      size_t t = 0;
      double done = 0.0;
      for (size_t i = 0; i < np && t < c.n_tiles; ++i) {
        const double goal = total * (double)(i + 1) / (double)np;
        const size_t t_begin = t;
        double acc = 0.0;
        while (t < c.n_tiles && (t == t_begin || i + 1 == np || done + acc + 0.5 * tile_cost[c.t0 + t] <= goal)) {
          acc += tile_cost[c.t0 + t];
          ++t;
        }
        done += acc;
        pcs.emplace_back(t_begin, t - t_begin);
        pcs_cost.push_back(acc);
      }
    };
    double longest = 1.0;
    for (const Chain& c : chains) {
      double total = 0.0;
      for (size_t t = 0; t < c.n_tiles; ++t) total += tile_cost[c.t0 + t];
      longest = std::max(longest, total);
    }
    // candidate piece costs: from the cost one CTA would carry if the work split evenly (shorter pieces only
    // add jobs and partials) up to the longest chain, ~15 % apart
    double all_cost = 0.0;
    for (const Chain& c : chains) {
      double total = 0.0;
      for (size_t t = 0; t < c.n_tiles; ++t) total += tile_cost[c.t0 + t];
      all_cost += total * heads * (slot_cnt(2 * c.pair + 1) > 0 ? 2 : 1);
    }
    std::vector<double> cand;
    for (double l = std::max(1.0, std::min(0.5 * all_cost / ctas, longest / 8.0)); l < longest; l = std::max(l + 0.5, l * 1.15))
      cand.push_back(l);
    cand.push_back(longest);
    // slots of a pair with at least one attending row in tiles [ta, tb): each is one job per kv-head
    // (per-(tile, slot) liveness once, so that every candidate of the search below is O(tiles))
    std::vector<uint8_t> tile_live(tiles.size() * (size_t)n_slots, 0);
    for (size_t t = 0; t < tiles.size(); ++t)
      for (size_t si = 0; si < tiles[t].slots.size(); ++si)
        tile_live[t * (size_t)n_slots + (size_t)tiles[t].slots[si]] = tiles[t].rows_or[si] != 0 ? 1 : 0;
    auto live_slots_in = [&](i32 pr, size_t ta, size_t tb) {
      int n_live = 0;
      for (int sl = 0; sl < 2; ++sl) {
        if (2 * pr + sl >= n_slots) continue;
        bool live = false;
        for (size_t t = ta; t < tb && !live; ++t) live = tile_live[t * (size_t)n_slots + (size_t)(2 * pr + sl)] != 0;
        n_live += live ? 1 : 0;
      }
      return n_live;
    };
    auto makespan = [&](double max_cost) {
      std::vector<double> costs;
      for (const Chain& c : chains) {
        pieces_of(c, max_cost);
        for (size_t pi = 0; pi < pcs.size(); ++pi) {
          const int n_live = live_slots_in(c.pair, c.t0 + pcs[pi].first, c.t0 + pcs[pi].first + pcs[pi].second);
          for (i32 h = 0; h < heads * n_live; ++h) costs.push_back(kJobConst + pcs_cost[pi]);
        }
      }
      if (costs.size() <= (size_t)ctas)  // a CTA each: the longest job is the makespan
        return costs.empty() ? 0.0 : *std::max_element(costs.begin(), costs.end());
      std::sort(costs.begin(), costs.end(), [](double a, double b) { return a > b; });
      std::priority_queue<double, std::vector<double>, std::greater<double>> bins;
      for (i32 c = 0; c < ctas; ++c) bins.push(0.0);
      double worst = 0.0;
      for (double c : costs) {
        const double load = bins.top() + c;
        bins.pop();
        bins.push(load);
        worst = std::max(worst, load);
      }
      return worst;
    };
    // The search depends only on the shape of the work (chains, tile costs, liveness, grid): consecutive decode steps
    // mostly keep it, so the last answer is kept per thread under a hash of exactly those inputs.  (A colliding
    // hash could only cost balance, never correctness: any piece length gives a valid plan.)
    std::vector<uint64_t> key;
    {
      uint64_t hsh = 1469598103934665603ull;
      auto mix = [&](uint64_t v) { hsh = (hsh ^ v) * 1099511628211ull; };
      for (const Chain& c : chains) { mix((uint64_t)c.pair); mix(c.t0); mix(c.n_tiles); }
      for (double tc : tile_cost) { uint64_t b; std::memcpy(&b, &tc, 8); mix(b); }
      for (size_t i = 0; i + 8 <= tile_live.size(); i += 8) { uint64_t b; std::memcpy(&b, tile_live.data() + i, 8); mix(b); }
      for (size_t i = tile_live.size() & ~(size_t)7; i < tile_live.size(); ++i) mix(tile_live[i]);
      uint64_t bj, bg;
      std::memcpy(&bj, &kJobConst, 8);
      std::memcpy(&bg, &gather_cost, 8);
      bg ^= (uint64_t)(mask_cost * 1024.0);
      key = {hsh, (uint64_t)chains.size(), (uint64_t)tiles.size(), (uint64_t)n_slots, (uint64_t)query_num,
             (uint64_t)heads, (uint64_t)ctas, bj, bg};
    }
    double best_len = cand.back();
    const char* env_k = std::getenv("DEFT_PLAN_CACHE");
    if (!(env_k && env_k[0] == '0') && key == S.plan_key) {
      best_len = S.plan_len;
    } else {
      double best = -1.0;
      // the pieces of a chain depend on the candidate only through their NUMBER: neighbouring candidates that cut
      // every chain into the same number of pieces have the same makespan
      std::vector<double> chain_total(chains.size(), 0.0);
      for (size_t ci = 0; ci < chains.size(); ++ci)
        for (size_t t = 0; t < chains[ci].n_tiles; ++t) chain_total[ci] += tile_cost[chains[ci].t0 + t];
      std::vector<size_t> nps(chains.size()), prev_nps;
      double prev_m = 0.0;
      for (double l : cand) {  // ascending: ties go to the longer piece (fewer partials)
        for (size_t ci = 0; ci < chains.size(); ++ci)
          nps[ci] = std::max<size_t>(1, (size_t)std::ceil(chain_total[ci] / l - 1e-9));
        const double m = nps == prev_nps ? prev_m : makespan(l);
        prev_nps = nps;
        prev_m = m;
        if (best < 0.0 || m <= best + 1e-9) { best = m; best_len = l; }
      }
      S.plan_key = key;
      S.plan_len = best_len;
    }

    std::vector<double> ucost;
    std::vector<std::vector<i32>> rows_of((size_t)query_num);  // CSR: partial rows of every query
    for (const Chain& c : chains) {
      pieces_of(c, best_len);
      const std::vector<std::pair<size_t, size_t>> chain_pcs = pcs;
      const std::vector<double> chain_cost = pcs_cost;
      for (size_t pi = 0; pi < chain_pcs.size(); ++pi) {
        const auto& pc = chain_pcs[pi];
        const size_t ta = c.t0 + pc.first, tb = ta + pc.second;  // tiles [ta, tb)
        // which slots of the pair have attending rows in this piece, and which rows
        i32 live_slots[2];
        uint32_t live_rows[2] = {0u, 0u};
        int n_live_slots = 0;
        bool dense[2] = {true, true};
        for (int sl = 0; sl < 2; ++sl) {
          const i32 slot = 2 * c.pair + sl;
          uint32_t rows = 0;
          bool all_dense = true;
          for (size_t t = ta; t < tb; ++t) {
            const Tile& tl = tiles[t];
            auto it = std::lower_bound(tl.slots.begin(), tl.slots.end(), slot);
            if (it != tl.slots.end() && *it == slot) {
              const size_t si = (size_t)(it - tl.slots.begin());
              rows |= tl.rows_or[si];
              all_dense = all_dense && tl.dense[si] != 0;
            } else {
              all_dense = false;
            }
          }
          if (rows) { live_slots[n_live_slots] = slot; live_rows[n_live_slots] = rows; dense[n_live_slots] = all_dense; ++n_live_slots; }
        }
        if (n_live_slots == 0) continue;
        deft_unit_t u{};
        u.kv_off = (i64)ta * 128;
        u.kv_tile_stride = 128;
        u.mask_tile_stride = n_live_slots * 128;
        u.n_tiles = (i32)pc.second;
        u.last_len = tiles[tb - 1].n_live;
        const i64 mask_base = (i64)u_mask.size();
        const bool unit_dense = dense[0] && (n_live_slots < 2 || dense[1]);  // nothing would read its masks
        if (!unit_dense)
          for (size_t t = ta; t < tb; ++t)
            for (int k = 0; k < n_live_slots; ++k) {
              const uint32_t* w = tile_slot(tiles[t], live_slots[k]);
              const size_t at = u_mask.size();
              u_mask.resize(at + 128, 0u);   // tokens past the tile's live ones stay zero (the masks' own tail is zero too)
              if (w) std::copy(w, w + tiles[t].n_live, u_mask.begin() + (long)at);
            }
        u.mask_off[1] = -1;
        u.q_id0[0] = u.q_id0[1] = -1;
        u.dense_tiles = 0;     // leading tiles every live slot attends densely (a prompt ahead of the subtree)
        for (size_t t = ta; t < tb; ++t) {
          bool all = true;
          for (int k = 0; k < n_live_slots && all; ++k) {
            auto it = std::lower_bound(tiles[t].slots.begin(), tiles[t].slots.end(), live_slots[k]);
            all = it != tiles[t].slots.end() && *it == live_slots[k] && tiles[t].dense[(size_t)(it - tiles[t].slots.begin())] != 0;
          }
          if (!all) break;
          ++u.dense_tiles;
        }
        {  // shortcut: all tiles full and on consecutive pages
          const size_t k0 = ta * 128, k1 = tb * 128;
          bool runp = tiles[tb - 1].n_live == 128 && u_kv[k0] >= 0 && !(u_kv[k0] & kFreshToken);
          for (size_t kk = k0 + 1; kk < k1 && runp; ++kk) runp = u_kv[kk] == u_kv[k0] + (i32)(kk - k0);
          u.page0 = runp ? u_kv[k0] : -1;
        }
        for (int k = 0; k < n_live_slots; ++k) {
          u.mask_off[k] = dense[k] ? -1 : mask_base + (i64)k * 128;
          u.q_off[k] = 32 * live_slots[k];
          u.q_cnt[k] = slot_cnt(live_slots[k]);
          u.part_base[k] = 32 * n_unit_slots++;
          bool runq = true;  // shortcut: consecutive query ids
          for (i32 rr = 1; rr < u.q_cnt[k] && runq; ++rr) runq = u_q[(size_t)u.q_off[k] + (size_t)rr] == u_q[(size_t)u.q_off[k]] + rr;
          u.q_id0[k] = runq ? u_q[(size_t)u.q_off[k]] : -1;
          for (i32 rr = 0; rr < u.q_cnt[k]; ++rr)
            if ((live_rows[k] >> rr) & 1u) rows_of[(size_t)u_q[(size_t)u.q_off[k] + (size_t)rr]].push_back(u.part_base[k] + rr);
        }
        units.push_back(u);
        ucost.push_back(kJobConst + chain_cost[pi]);
      }
    }
    for (i32 qv = 0; qv < query_num; ++qv) {
      std::sort(rows_of[(size_t)qv].begin(), rows_of[(size_t)qv].end());
      u_csr.off[(size_t)qv + 1] = u_csr.off[(size_t)qv] + (i32)rows_of[(size_t)qv].size();
      u_csr.rows.insert(u_csr.rows.end(), rows_of[(size_t)qv].begin(), rows_of[(size_t)qv].end());
    }
    // (unit, kv-head, slot) jobs -> CTAs, longest first onto the least loaded CTA; job = ((unit * hkv + head) << 1) | slot
    if (hkv > 0) {
      std::vector<i32> order(units.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (i32)i;
      std::stable_sort(order.begin(), order.end(), [&](i32 a, i32 b) { return ucost[(size_t)a] > ucost[(size_t)b]; });
      using Bin = std::pair<double, i32>;
      std::priority_queue<Bin, std::vector<Bin>, std::greater<Bin>> bins;
      for (i32 c = 0; c < ctas; ++c) bins.push({0.0, c});
      std::vector<std::vector<i32>> per((size_t)ctas);
      std::vector<size_t> n_shared((size_t)ctas, 0);  // leading jobs of a CTA's list shared with its pair
      size_t n_jobs_total = 0;
      for (const deft_unit_t& u : units) n_jobs_total += (size_t)hkv * (size_t)((u.q_cnt[0] > 0) + (u.q_cnt[1] > 0));
      // Throughput regime (many jobs per CTA, e.g. a forest of trees): the two slot-jobs of a unit read the same
      // K/V tiles, so they go to the two CTAs of a PAIR at the same position of their lists -- they then run
      // side by side and the second read hits L2 instead of HBM.  Latency regime (about one job per CTA): every
      // job on its own, longest first onto the least loaded CTA (the CTAs all start together anyway).
      const char* env_p = std::getenv("DEFT_PLAN_PAIR");
      const bool pair_mode = ctas >= 2 && ctas % 2 == 0 && (env_p ? env_p[0] == '1' : n_jobs_total >= (size_t)4 * (size_t)ctas);
      plan_paired = pair_mode;
      if (pair_mode) {
        const i32 n_pairs_b = ctas / 2;
        std::priority_queue<Bin, std::vector<Bin>, std::greater<Bin>> pbins;
        for (i32 b = 0; b < n_pairs_b; ++b) pbins.push({0.0, b});
        std::vector<double> load((size_t)ctas, 0.0);
        for (i32 ui : order) {
          if (units[(size_t)ui].q_cnt[0] <= 0 || units[(size_t)ui].q_cnt[1] <= 0) continue;
          for (i32 h = 0; h < hkv; ++h) {
            Bin b = pbins.top();
            pbins.pop();
            for (i32 k = 0; k < 2; ++k) {
              per[(size_t)(2 * b.second + k)].push_back(((ui * hkv + h) << 1) | k);
              load[(size_t)(2 * b.second + k)] += ucost[(size_t)ui];
              ++n_shared[(size_t)(2 * b.second + k)];
            }
            b.first += ucost[(size_t)ui];
            pbins.push(b);
          }
        }
        while (!bins.empty()) bins.pop();
        for (i32 c = 0; c < ctas; ++c) bins.push({load[(size_t)c], c});
        for (i32 ui : order) {
          const bool a = units[(size_t)ui].q_cnt[0] > 0, b2 = units[(size_t)ui].q_cnt[1] > 0;
          if (a && b2) continue;
          const i32 k = a ? 0 : 1;
          for (i32 h = 0; h < hkv; ++h) {
            Bin b = bins.top();
            bins.pop();
            per[(size_t)b.second].push_back(((ui * hkv + h) << 1) | k);
            b.first += ucost[(size_t)ui];
            bins.push(b);
          }
        }
      } else {
      for (i32 ui : order)
        for (i32 h = 0; h < hkv; ++h)
          for (i32 k = 0; k < 2; ++k) {
            if (units[(size_t)ui].q_cnt[k] <= 0) continue;
            Bin b = bins.top();
            bins.pop();
            per[(size_t)b.second].push_back(((ui * hkv + h) << 1) | k);
            b.first += ucost[(size_t)ui];
            bins.push(b);
          }
      }
      // records [0, ctas): every CTA's first job; the others follow, consecutive per CTA
      auto record = [&](i32 job) {
        deft_job_t r{};
        r.job = job;
        if (job >= 0) r.unit = units[(size_t)((job >> 1) / hkv)];
        return r;
      };
      u_jobs.resize((size_t)ctas);
      u_job_off.push_back(0);
      i32 total = 0;
      for (i32 c = 0; c < ctas; ++c) {
        const std::vector<i32>& mine = per[(size_t)c];
        deft_job_t first = record(mine.empty() ? -1 : mine[0]);
        first.n_jobs = (i32)mine.size();
        first.next = (i32)u_jobs.size();
        first.shared = n_shared[(size_t)c] > 0 ? 1 : 0;
        u_jobs[(size_t)c] = first;
        for (size_t i = 1; i < mine.size(); ++i) {
          deft_job_t r = record(mine[i]);
          r.shared = i < n_shared[(size_t)c] ? 1 : 0;
          u_jobs.push_back(r);
        }
        total += (i32)mine.size();
        u_job_off.push_back(total);
      }
    }
  }

  lap("units + jobs");
  // ---- pack
  deft_tables_t* t = S.spare ? S.spare : new (std::nothrow) deft_tables_t();
  S.spare = nullptr;
  if (!t) {
    deft::set_error("build_tables: out of memory");
    return nullptr;
  }
  struct Src { const void* p; size_t n; size_t elem; };
  const Src src[DEFT_T_COUNT] = {
      {node_q.data(), node_q.size(), 8}, {node_kv.data(), node_kv.size(), 8},
      {node_q_len.data(), node_q_len.size(), 8}, {node_kv_len.data(), node_kv_len.size(), 8},
      {node_q_offset.data(), node_q_offset.size(), 8}, {node_kv_offset.data(), node_kv_offset.size(), 8},
      {block_q.data(), block_q.size(), 8}, {block_q_cnts.data(), block_q_cnts.size(), 8},
      {block_q_offset.data(), block_q_offset.size(), 8}, {block_masks.data(), block_masks.size(), 8},
      {block_kv.data(), block_kv.size(), 8}, {block_lens.data(), block_lens.size(), 8},
      {f_items.data(), f_items.size(), sizeof(deft_item_t)}, {f_groups.data(), f_groups.size(), sizeof(deft_group_t)},
      {f_csr.off.data(), f_csr.off.size(), 4}, {f_csr.rows.data(), f_csr.rows.size(), 4},
      {n_items.data(), n_items.size(), sizeof(deft_item_t)}, {n_groups.data(), n_groups.size(), sizeof(deft_group_t)},
      {n_csr.off.data(), n_csr.off.size(), 4}, {n_csr.rows.data(), n_csr.rows.size(), 4},
      {units.data(), units.size(), sizeof(deft_unit_t)}, {u_csr.off.data(), u_csr.off.size(), 4},
      {u_csr.rows.data(), u_csr.rows.size(), 4}, {u_kv.data(), u_kv.size(), 4}, {u_mask.data(), u_mask.size(), 4},
      {u_q.data(), u_q.size(), 4}, {u_job_off.data(), u_job_off.size(), 4}, {u_jobs.data(), u_jobs.size(), sizeof(deft_job_t)},
      {u_blk.data(), u_blk.size(), 4},
  };
  if (layout) {   // capacity-padded regions: offsets only move when a table outgrows its region
    bool grow = n_unit_slots > layout->slot_cap;
    for (int i = 0; i < DEFT_T_COUNT; ++i) grow = grow || (i64)((src[i].n * src[i].elem + 255) / 256 * 256) > layout->cap_bytes[i];
    if (grow) {     // ... and then every region takes its headroom afresh: the tables grow together, so do the layouts.
      // What moving costs is a re-capture of the decode step's CUDA graphs (3-5 ms, measured: more than three steps),
      // what headroom costs is upload bytes: +50 % for the reference's big int64 tables, +100 % (and 4 KB: the job
      // and unit tables move in jumps when the plan search changes the piece length) for the small native ones.
      for (int i = 0; i < DEFT_T_COUNT; ++i) {
        const size_t region = (src[i].n * src[i].elem + 255) / 256 * 256;
        const size_t want = i < 12 ? region + region / 2 + 1024 : 2 * region + 4096;
        layout->cap_bytes[i] = std::max<i64>(layout->cap_bytes[i], (i64)((want + 255) / 256 * 256));
      }
      layout->slot_cap = std::max<i64>(layout->slot_cap, n_unit_slots + n_unit_slots / 2 + 8);
      ++layout->version;
    }
  }
  size_t off = 0;
  for (int i = 0; i < DEFT_T_COUNT; ++i) {
    t->dir[2 * i] = (i64)off;
    t->dir[2 * i + 1] = (i64)src[i].n;
    off += layout ? (size_t)layout->cap_bytes[i] : (src[i].n * src[i].elem + 255) / 256 * 256;
  }
  t->packed.resize(std::max<size_t>(off, 256));
  for (int i = 0; i < DEFT_T_COUNT; ++i) {  // every byte is written: the tables, and zeros up to the next 256-byte boundary
    unsigned char* dst = t->packed.data() + t->dir[2 * i];
    const size_t nb = src[i].n * src[i].elem;
    const size_t end = (i + 1 < DEFT_T_COUNT ? (size_t)t->dir[2 * i + 2] : t->packed.size()) - (size_t)t->dir[2 * i];
    if (nb) std::memcpy(dst, src[i].p, nb);
    std::memset(dst + nb, 0, end - nb);
  }
  t->scalars[0] = query_num;
  t->scalars[1] = (i64)node_q_len.size();
  t->scalars[2] = total_kv_len;
  t->scalars[3] = block_len;
  t->scalars[4] = (i64)block_q.size();
  t->scalars[5] = n_rows;
  t->scalars[6] = n_unit_slots;
  t->scalars[7] = u_job_off.empty() ? 0 : (i64)u_job_off.size() - 1;
  t->scalars[8] = plan_paired ? 1 : 0;
  t->scalars[9] = layout ? layout->slot_cap : n_unit_slots;
  t->scalars[10] = fresh_page ? 1 : 0;
  return t;
}

const void* deft_b200_tables_data(const deft_tables_t* t) { return t ? t->packed.data() : nullptr; }
size_t deft_b200_tables_bytes(const deft_tables_t* t) { return t ? t->packed.size() : 0; }

int deft_b200_tables_directory(const deft_tables_t* t, int64_t* dir) {
  if (!t || !dir) return DEFT_E_ARG;
  std::memcpy(dir, t->dir, sizeof(t->dir));
  return DEFT_OK;
}

int deft_b200_tables_scalars(const deft_tables_t* t, int64_t* out) {
  if (!t || !out) return DEFT_E_ARG;
  std::memcpy(out, t->scalars, sizeof(t->scalars));
  return DEFT_OK;
}

deft_tree_t* deft_b200_tree_new(void) { return new (std::nothrow) deft_tree_t(); }
void deft_b200_tree_free(deft_tree_t* t) { delete t; }
int64_t deft_b200_tree_pages(const deft_tree_t* t) { return t ? t->n_pages : -1; }

int deft_b200_tree_set(deft_tree_t* t, int32_t n_nodes, const int32_t* parent, const int64_t* kv_off, const int64_t* kv,
                       const int64_t* q_off, const int64_t* qs, const int64_t* tix_row, int32_t query_num) {
  if (!t || n_nodes <= 0 || !parent || !kv_off || !kv || !q_off || !qs || query_num <= 0) {
    deft::set_error("tree_set: null or empty tree");
    return DEFT_E_ARG;
  }
  for (i32 n = 0; n < n_nodes; ++n)
    if (kv_off[n + 1] < kv_off[n] || q_off[n + 1] < q_off[n]) {
      deft::set_error("tree_set: offsets of node %d decrease", n);
      return DEFT_E_ARG;
    }
  t->parent.assign(parent, parent + n_nodes);
  t->pages.resize((size_t)n_nodes);
  for (i32 n = 0; n < n_nodes; ++n) t->pages[(size_t)n].assign(kv + kv_off[n], kv + kv_off[n + 1]);
  t->q_off.assign(q_off, q_off + n_nodes + 1);
  t->qs.assign(qs, qs + q_off[n_nodes]);
  if (tix_row) t->tix.assign(tix_row, tix_row + n_nodes);
  else t->tix.clear();
  t->query_num = query_num;
  t->n_pages = kv_off[n_nodes] - kv_off[0];
  t->tiler = deft_tiler_cache();          // (another topology: the kept tile order is void)
  return DEFT_OK;
}

int deft_b200_tree_append(deft_tree_t* t, int32_t n, const int32_t* node, const int64_t* page) {
  if (!t || n < 0 || (n > 0 && (!node || !page))) {
    deft::set_error("tree_append: null argument");
    return DEFT_E_ARG;
  }
  for (i32 i = 0; i < n; ++i)
    if (node[i] >= (i32)t->pages.size()) {
      deft::set_error("tree_append: node %d of a tree of %zu nodes", node[i], t->pages.size());
      return DEFT_E_ARG;
    }
  for (i32 i = 0; i < n; ++i) {
    if (node[i] < 0) continue;            // a leaf the walk left out (paused)
    t->pages[(size_t)node[i]].push_back(page[i]);
    ++t->n_pages;
    if (t->tiler.valid) t->tiler.pending.emplace_back(node[i], page[i]);
  }
  return DEFT_OK;
}

deft_tables_t* deft_b200_build_tables_trees(deft_tree_t* const* trees, int32_t n_trees, int64_t tix_max_ctx,
                                            int32_t block_len, int32_t max_q_len, int32_t max_block_len,
                                            int32_t node_split, int32_t hkv, int32_t n_ctas, deft_layout_t* layout,
                                            const int32_t* fresh_page) {
  if (!trees || n_trees <= 0) {
    deft::set_error("build_tables_trees: no trees");
    return nullptr;
  }
  // the flat arrays of deft_b200_build_tables: the trees one after the other (nodes, pages and queries offset)
  thread_local std::vector<i32> parent;
  thread_local std::vector<i64> kv_off, kv, q_off, qs, tix;
  parent.clear(); kv.clear(); qs.clear(); tix.clear();
  kv_off.assign(1, 0);
  q_off.assign(1, 0);
  i64 query_base = 0;
  bool all_tix = tix_max_ctx > 0;
  for (i32 ti = 0; ti < n_trees; ++ti) {
    const deft_tree_t* t = trees[ti];
    if (!t || t->parent.empty()) {
      deft::set_error("build_tables_trees: tree %d is empty (deft_b200_tree_set first)", ti);
      return nullptr;
    }
    const i32 node_base = (i32)parent.size();
    const i64 qs_base = (i64)qs.size();
    for (size_t n = 0; n < t->parent.size(); ++n) {
      parent.push_back(t->parent[n] < 0 ? -1 : t->parent[n] + node_base);
      kv.insert(kv.end(), t->pages[n].begin(), t->pages[n].end());
      kv_off.push_back((i64)kv.size());
      q_off.push_back(qs_base + t->q_off[n + 1] - t->q_off[0]);
    }
    for (i64 qv : t->qs) qs.push_back(qv + query_base);
    if (t->tix.size() == t->parent.size()) tix.insert(tix.end(), t->tix.begin(), t->tix.end());
    else all_tix = false;
    query_base += t->query_num;
  }
  if (tix_max_ctx > 0 && !all_tix) {
    deft::set_error("build_tables_trees: tree-index mode needs the index rows of every tree");
    return nullptr;
  }
  g_tree_hints = trees;
  g_n_tree_hints = n_trees;
  deft_tables_t* out = deft_b200_build_tables((i32)parent.size(), parent.data(), kv_off.data(), kv.data(), q_off.data(),
                                              qs.data(), tix_max_ctx > 0 ? tix.data() : nullptr, tix_max_ctx, (i32)query_base,
                                              block_len, max_q_len, max_block_len, node_split, hkv, n_ctas, layout, fresh_page);
  g_tree_hints = nullptr;
  g_n_tree_hints = 0;
  return out;
}

deft_layout_t* deft_b200_layout_new(void) { return new (std::nothrow) deft_layout_t(); }
void deft_b200_layout_set_native_only(deft_layout_t* l, int on) {
  if (l) l->native_only = on ? 1 : 0;
}
void deft_b200_layout_free(deft_layout_t* l) { delete l; }
int64_t deft_b200_layout_version(const deft_layout_t* l) { return l ? l->version : -1; }

void deft_b200_tables_free(deft_tables_t* t) {
  if (t && !g_scratch.spare) g_scratch.spare = t;   // its buffer serves the next build of this thread
  else delete t;
}

}  // extern "C"
