// Stage 1, tensor-core path (sm_100a): tcgen05.mma with TMEM accumulators, warp-specialised persistent CTAs
// walking the unit plan.  Third generation of the kernel ("task-parity softmax groups").
//
// One job = (unit, kv-head, slot): ONE slot of <= 32 queries x G = H/HKV heads = one M = 128 UMMA accumulator,
// along the unit's chain of KV tiles (128 tokens x D, K and V, gathered from the token-granular paged pool into
// 128B-swizzled shared memory by TMA boxes / TMA gather4 / cp.async).  Per KV tile t of the chain:
//
//   S(t)[128 x 128] = Q[128 x D] . K(t)^T     tcgen05.mma kind::f16, SS (Q, K in swizzled smem) -> TMEM
//   P(t)            = exp2(S*c - m_ref)       one softmax THREAD per row: the whole 128-column row comes out of TMEM
//                                             once, the row maximum is thread-local (no exchange), P goes back to
//                                             TMEM as packed fp16 -- the A operand of the next MMA
//   O[128 x D]     += P(t) . V(t)             tcgen05.mma TS (A = P in TMEM, B = V MN-major swizzled smem)
//
// The softmax is what binds a tile step (16 K exponentials per tile on the SM's 16-lane MUFU = 0.53 us, the same as
// the two MMAs on the tensor pipe), and a single chain of softmax threads leaves the MUFU idle while it waits for
// TMEM loads, stores and barriers.  So the CTA's work is a sequence of TASKS -- the tiles of job 0, the epilogue of
// job 0, the tiles of job 1, ... -- dealt alternately to TWO softmax groups (warps 0-3 and 4-7, 128 threads each, both
// on all four sub-partitions): while one group is in the latency-bound part of its task the other one is in its
// exponentials.  Each group has its own S buffer and its own P buffer in TMEM:
//
//   TMEM (512 columns):  O [0, 128)   S_A [128, 256)   S_B [256, 384)   P_A [384, 448)   P_B [448, 512)
//
// S_g is free again as soon as the group has the row in registers (the S issuer then computes the group's NEXT tile
// into it while the exponentials run), P_g when P V of the tile has completed.  Both groups accumulate into the ONE
// O: they share the row's reference maximum m_ref (shared memory, one float per row).  Order is kept by a chain of
// "checked" barriers: task k reads / raises m_ref only after task k-1 has done so.  A tile whose row maximum exceeds
// m_ref by more than 2^15 raises it: the thread waits for P V of the previous tile, rescales its row of O in TMEM
// and publishes the new m_ref before it lets task k+1 go on (test_reference_maximum_is_raised_mid_chain).  The row
// sums are kept per thread and merged in the epilogue (log-sum-exp per group, through shared memory).  The epilogue
// of a job is a task like any other: one group turns O into the job's partial (O / l as fp16 + log-sum-exp) while
// the other group already works the first tile of the next job.
//
// Warp roles (512 threads, 1 CTA per SM, register budgets moved with setmaxnreg: softmax 184, the rest 72):
//   warps 0-3 / 4-7   softmax group A / B (thread = row = TMEM lane)
//   warp  8           S issuer (one elected thread), TMEM alloc / dealloc
//   warp  9           P V issuer
//   warp  10          Q tile: one TMA box per panel when the slot's query ids are consecutive, else cp.async gathers
//   warp  11          row masks: per-token words -> per-query token bitmasks (warp-shuffle bit transpose) + "dense
//                     tile" flag; skipped for jobs whose tiles are all dense (the prompt)
//   warps 12-13/14-15 K / V producers, 64 rows each, on rings of their own.  32 consecutive pages = ONE TMA box per
//                     panel, scattered pages = TMA tile::gather4, or cp.async when TMA is switched off
// Cluster pairs (throughput regime, pair-aligned job lists): the two CTAs of a cluster work the two slots of the same
// (unit, kv-head); each K/V tile is loaded ONCE, the CTA of rank r issuing rows [64r, 64r + 64) with
// .multicast::cluster into both CTAs' shared memory.
//
// Reference semantics: DeFT/deft/layers/attention/tree_attention.py:860-976 (Flatten stage 1) and :170-293 (Node
// stage 1).  Same-box A/B against the previous generation (two threads per row agreeing on the maximum through a named
// barrier, S triple-buffered with P written in place): DESIGN.md section 6.
#include "umma_ptx.cuh"

namespace deft {
namespace {
using namespace umma;

constexpr int kThreads = 512;  // 16 warps: 4 register-budget groups of 4 (setmaxnreg works per warpgroup)
constexpr int kMmaWarp = 8, kPvWarp = 9, kQWarp = 10, kMaskWarp = 11, kKvWarp0 = 12;  // 12-13: K producers; 14-15: V
constexpr int kSoftmaxRegs = 184, kProducerRegs = 72;  // 256 * 184 + 256 * 72 = 64 K registers (launch: 128 each)
constexpr int kKStages = 3, kVStages = 2, kMaskStages = 4;
constexpr float kRaise = 15.f;  // a row maximum this far (log2) above m_ref raises it: P <= 2^15 stays in fp16's range

// Optional per-CTA timeline (test/profiling hook, deft_b200_set_trace_buffer): trace[cta][event] = SM cycles since the
// CTA started.  Per-tile events of the first job's first six tiles take 8 slots each from kTrTile0.
constexpr int kTraceSlots = 128;
enum : int {
  kTrStart = 0, kTrQIds = 1, kTrQIssued = 2, kTrMask0 = 4, kTrKUnit = 5, kTrMmaQFull = 6, kTrEpiBegin = 7, kTrEpiEnd = 8,
  kTrEnd = 9, kTrEpiODone = 10, kTrKRole = 12, kTrJob0 = 64,  // + 2 * job: epilogue end | tiles and kind of the job
  kTrTile0 = 16,  // + 8 * tile: K issued, S issued, S in registers, checked, P half 0 handed, P half 1 handed, V issued,
                  //             P V issued
};
#define DEFT_TRACE(ev)                                                                          \
  do {                                                                                          \
    if (p.trace != nullptr && (ev) < kTraceSlots) p.trace[(int64_t)blockIdx.x * kTraceSlots + (ev)] = (int)(clock64() - t_start); \
  } while (0)

// barrier indices
enum : int {
  K_FULL = 0, K_EMPTY = K_FULL + kKStages, V_FULL = K_EMPTY + kKStages, V_EMPTY = V_FULL + kVStages,
  Q_FULL = V_EMPTY + kVStages, Q_EMPTY = Q_FULL + 1,
  M_FULL = Q_EMPTY + 1,                    // [stage]
  M_EMPTY = M_FULL + kMaskStages,
  S_FULL = M_EMPTY + kMaskStages,          // [group]: S of the group's tile is in TMEM
  S_FREE = S_FULL + 2,                     // [group]: the group has the row in registers
  P_FULL = S_FREE + 2,                     // [group][half]: P of one 64-token half is in TMEM
  P_FREE = P_FULL + 4,                     // [group]: P V of the group's tile has completed
  CHK = P_FREE + 2,                        // [group]: the group's task has read / raised the shared reference maximum
  O_DONE = CHK + 2,                        // one phase per job: the last P V has landed, O is complete
  O_EMPTY = O_DONE + 1,                    // one phase per job: the epilogue has read O
  kNumBars = O_EMPTY + 1
};

template <int D>
struct Layout {
  static constexpr int kOperandBytes = kRows * D * 2;                      // Q, K or V tile
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kOperandBytes;                            // [stage]
  static constexpr int kV = kK + kKStages * kOperandBytes;                 // [stage]
  static constexpr int kMask = kV + kVStages * kOperandBytes;              // [stage][32 queries][4] u32
  static constexpr int kFlag = kMask + kMaskStages * kTileN * 4;           // [stage] u32
  static constexpr int kMref = kFlag + kMaskStages * 4;                    // [128] f32: the rows' reference maximum (log2 domain)
  static constexpr int kLse = kMref + kRows * 4;                           // [job parity][group][128] f32: m + log2(l) per group
  static constexpr int kBars = kLse + 2 * 2 * kRows * 4;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kBytes = kTmemSlot + 16;
  static constexpr int kAlloc = kBytes + 1024;  // slack for the manual 1024-byte alignment
};

// The jobs of one CTA.  job = ((unit * HKV + kv-head) << 1) | slot of the unit's pair.  Either the host-balanced
// record lists (deft_job_t: the CTA's first record sits at [blockIdx.x] and carries its unit, so a CTA starts from
// ONE load), or jobs c, c + grid, ... over the unit table.
struct Jobs {
  const deft_job_t* recs;  // null: strided over the unit table
  int n, next;
  __device__ __forceinline__ Jobs(const AttnParams& p) {
    if (p.job_off != nullptr) {
      recs = p.jobs;
      n = next = 0;
      if ((int)blockIdx.x < p.n_ctas) {
        prefetch_l1(recs + blockIdx.x);  // (the record may straddle two lines: both are on their way)
        const int4 hdr = *reinterpret_cast<const int4*>(recs + blockIdx.x);
        n = hdr.x >= 0 ? hdr.y : 0;
        next = hdr.z;
      }
    } else {
      recs = nullptr;
      const int n_units = p.n_units_dev ? *p.n_units_dev : p.n_units;
      const int total = n_units * p.HKV * 2;
      n = (int)blockIdx.x < total ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
      next = 0;
    }
  }
  // i-th job of this CTA -> (unit, kv-head, slot); a pair's second slot may be empty (then the job is nobody's).
  // `shared`: the other CTA of my cluster pair works the other slot of the same (unit, kv-head) at the same
  // position of its list, so every K/V tile is loaded once for both (each CTA issues half of it, multicast).
  __device__ __forceinline__ bool get(const AttnParams& p, int i, deft_unit_t& u, int& hkv, int& k, bool& shared) const {
    int job;
    shared = false;
    if (recs != nullptr) {
      const deft_job_t* r = i == 0 ? recs + blockIdx.x : recs + next + (i - 1);
      job = r->job;
      shared = r->shared != 0 && p.tma_kv != 0 && p.tma_gather != 0 && p.clustered != 0;
      u = r->unit;
      if (i + 1 < n) prefetch_l1(recs + next + i);  // the next job's record: no load latency between two jobs
    } else {
      job = (int)blockIdx.x + i * (int)gridDim.x;
      u = p.units[(job >> 1) / p.HKV];
      if (i + 1 < n) prefetch_l1(p.units + ((job + (int)gridDim.x) >> 1) / p.HKV);
    }
    k = job & 1;
    hkv = (job >> 1) % p.HKV;
    return (k == 0 ? u.q_cnt[0] : u.q_cnt[1]) > 0;
  }
};

// the factor that takes a sum kept against the reference `from` to the reference `to` >= from (no NaN for -inf)
__device__ __forceinline__ float rescale(float from, float to) { return from == to ? 1.f : fast_exp2(from - to); }

template <int D, int G, bool kDbg>
__global__ void __launch_bounds__(kThreads, 1) stage1_umma_kernel(const __grid_constant__ AttnParams p) {
  using L = Layout<D>;
  constexpr int CH = D / 8;           // 16-byte chunks per row
  constexpr int R = kMaxGroupQ * G;   // live rows of a full slot
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kColS = 128, kColP = 384;  // S_g at kColS + 128 g, P_g at kColP + 64 g
  constexpr uint32_t kIdescQK = instr_desc(kTileN, false);
  constexpr uint32_t kIdescPV = instr_desc(D, true);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + L::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_start = clock64();
  griddep_launch_dependents();       // stage 2 may launch early: its CTAs wait for this grid to finish
  if (tid == 32) {  // the first TMA of a map otherwise pays the fetch of its 128-byte descriptor
    if (p.tma_kv) { prefetch_tensormap(&p.tmap_k); prefetch_tensormap(&p.tmap_v); }
    if (p.tma_gather) { prefetch_tensormap(&p.tmap_kg); prefetch_tensormap(&p.tmap_vg); }
    if (p.tma_q) prefetch_tensormap(&p.tmap_q);
  }
  if (p.plan_fresh) griddep_wait();  // the plan itself comes from the preceding (plan) kernel
  const Jobs jobs(p);  // (its load is in flight under the barrier set-up and the TMEM allocation below)
  if (tid == 0) {
    for (int s = 0; s < kKStages; ++s) { mbar_init(bar(K_FULL + s), 64); mbar_init(bar(K_EMPTY + s), 2); }  // EMPTY: my issuer + the pair's (or mine twice)
    for (int s = 0; s < kVStages; ++s) { mbar_init(bar(V_FULL + s), 64); mbar_init(bar(V_EMPTY + s), 2); }
    mbar_init(bar(Q_FULL), 32); mbar_init(bar(Q_EMPTY), 1);
    for (int m = 0; m < kMaskStages; ++m) { mbar_init(bar(M_FULL + m), 32); mbar_init(bar(M_EMPTY + m), 128); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(bar(S_FULL + g), 1); mbar_init(bar(S_FREE + g), 128); mbar_init(bar(P_FREE + g), 1);
      mbar_init(bar(CHK + g), 128);
      for (int h = 0; h < 2; ++h) mbar_init(bar(P_FULL + 2 * g + h), 128);
    }
    mbar_init(bar(O_DONE), 1); mbar_init(bar(O_EMPTY), 128);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(base + L::kTmemSlot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (p.clustered) cluster_sync();  // the pair's barriers exist before anything of mine is multicast to them
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gbase + L::kTmemSlot);
  const uint32_t crank = p.clustered ? cluster_ctarank() : 0u;

  // Programmatic dependent launch: everything up to here (barrier init, TMEM allocation, job list) overlapped
  // the tail of the preceding kernel; q, the KV pool and the partial workspace may still be in its hands.  The K / V
  // producers go further before they wait: their first job's record, load descriptors and page ids are plan data
  // (nothing of this stream's recent past writes them), and reading them is two dependent trips to memory.
  if (warp < kKvWarp0 && warp != kMaskWarp) griddep_wait();   // (the mask warp reads plan data only: it never waits)
  if (tid == 0) DEFT_TRACE(kTrStart);
  if (warp >= 8) {
  reg_dealloc<kProducerRegs>();  // warps 8-15: two whole warpgroups give registers away
  if (warp >= kKvWarp0) {
    // ============================== K / V producers ==============================
    // warps 12-13: K rows [64w, 64w + 64) of every tile; warps 14-15: V likewise.  K and V run on rings of their own (K
    // is released as soon as S is done, a tile earlier than V).  Three ways to load a tile:
    //  (a) the unit's tokens sit on consecutive pages (page0 >= 0: a prompt): lane (block of 32 rows, panel) issues
    //      one TMA box, nothing is looked up;
    //  (b) native tables (p.u_blk): the builder has laid the scattered tokens out as aligned runs of 32 / 16 / 8
    //      consecutive pages wherever the pages allow (metadata.cpp, part 1b) and says so per chunk of 8 rows: lane
    //      (chunk, role) issues a box for a run, or gather4s for four arbitrary rows.  The copy engine is bound by the
    //      NUMBER of these instructions: a tile of gather4s costs it 2.2 us, a tile of boxes 0.6 us.  The descriptors
    //      and page ids are asked for two tiles ahead of their use (a table read costs about a tile step);
    //  (c) anything else (the reference's int64 tables, TMA or gather4 switched off): per block of 32 rows a box when
    //      its pages are consecutive, else gather4, else cp.async.
    // Rows past the tile's length and dummy tokens (page < 0) arrive as zeros: they name a row outside the gather map.
    constexpr int NP = D / 64;
    const int kv = (warp - kKvWarp0) >> 1, w = (warp - kKvWarp0) & 1;
    const int stages = kv == 0 ? kKStages : kVStages;
    const int FULL = kv == 0 ? K_FULL : V_FULL, EMPTY = kv == 0 ? K_EMPTY : V_EMPTY;
    const CUtensorMap* m32 = kv == 0 ? &p.tmap_k : &p.tmap_v;
    const CUtensorMap* mg = kv == 0 ? &p.tmap_kg : &p.tmap_vg;
    const bool native = p.u_blk != nullptr && p.u_kv_bytes == 4 && p.tma_kv != 0 && p.tma_gather != 0;
    uint32_t cnt = 0;  // tiles produced
    bool waited = false;
    auto dep_wait = [&]() {   // the preceding kernel's memory, once, before the first byte of q / K / V moves
      if (!waited) griddep_wait();
      waited = true;
    };
    if (warp == kKvWarp0 && lane == 0) DEFT_TRACE(kTrKRole);
    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      if (warp == kKvWarp0 && lane == 0 && ji == 0 && u.n_tiles > 0) DEFT_TRACE(kTrKUnit);
      const bool known_run = u.page0 >= 0 && p.tma_kv != 0;
      // shared job: the pair loads every tile ONCE -- the CTA of rank r issues rows [64r, 64r + 64) for both
      const bool mine = !shared || (uint32_t)w == crank;
      if (known_run) {
        // ---- (a)
        dep_wait();
        for (int t = 0; t < u.n_tiles; ++t, ++cnt) {
          const int st = cnt % stages;
          mbar_wait<64>(bar(EMPTY + st), ((cnt / stages) & 1) ^ 1);
          const uint32_t full = bar(FULL + st);
          if (lane == 0) mbar_expect_tx(full, 64 * D * 2);
          if (mine && lane < 2 * NP) {
            const int b = lane / NP, pn = lane % NP;
            const uint32_t dst = base + (kv == 0 ? L::kK : L::kV) + st * L::kOperandBytes + pn * kPanelBytes + (w * 64 + b * 32) * 128;
            const int page = u.page0 + t * kTileN + w * 64 + b * 32;
            if (shared) tma_load_3d_mc(dst, m32, full, pn * 64, hkv, page, 0x3);
            else tma_load_3d(dst, m32, full, pn * 64, hkv, page);
          }
          mbar_arrive(full);
          if (w == 0 && lane == 0 && ji == 0 && t < 6) DEFT_TRACE(kTrTile0 + 8 * t + (kv == 0 ? 0 : 6));
        }
      } else if (native) {
        // ---- (b) lane = (chunk c of my 8, role li): boxes are issued by li < NP (panel li); gather4s by li < 2 NP
        // (rows 4 g4 .. 4 g4 + 3 of the chunk, panel pn)
        const int c = lane >> 2, li = lane & 3;
        const int g4 = li / NP, pn = li % NP;
        const int32_t* blk = p.u_blk + (u.kv_off >> 7) * 16 + w * 8 + c;
        const int32_t* pgs = reinterpret_cast<const int32_t*>(p.u_kv) + u.kv_off + w * 64 + c * 8 + (g4 & 1) * 4;
        auto desc_of = [&](int t) { return t < u.n_tiles ? blk[t * 16] : 0; };
        auto pages_of = [&](int t) { return t < u.n_tiles ? *reinterpret_cast<const int4*>(pgs + t * kTileN) : make_int4(-1, -1, -1, -1); };
        int d_next = desc_of(0), d_after = desc_of(1);
        int4 g_next = pages_of(0), g_after = pages_of(1);
        if (!waited) {   // (the loads above are on their way while the preceding kernel drains; they have landed by the wait's end)
          asm volatile("" ::"r"(d_next), "r"(g_next.x));
          dep_wait();
        }
        for (int t = 0; t < u.n_tiles; ++t, ++cnt) {
          const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
          const int st = cnt % stages;
          const int desc = d_next;
          const int4 pg = g_next;
          d_next = d_after;
          g_next = g_after;
          d_after = desc_of(t + 2);
          g_after = pages_of(t + 2);
          mbar_wait<64>(bar(EMPTY + st), ((cnt / stages) & 1) ^ 1);
          const uint32_t full = bar(FULL + st);
          if (lane == 0) mbar_expect_tx(full, 64 * D * 2);
          if (mine) {
            const uint32_t dst = base + (kv == 0 ? L::kK : L::kV) + st * L::kOperandBytes + (w * 64 + c * 8) * 128;
            // bit 27: the chunk holds tokens of THIS step -- rows of the step's new K / V (fused append), "page" = query id
            const int kind = (desc >> 28) & 3, page = desc & 0x07ffffff;
            const bool fresh = (desc >> 27) & 1;
            if (kind == 3) {
              if ((c & 3) == 0 && li < NP) {
                const CUtensorMap* m = fresh ? (kv == 0 ? &p.tmap_nk : &p.tmap_nv) : m32;
                if (shared) tma_load_3d_mc(dst + li * kPanelBytes, m, full, li * 64, hkv, page, 0x3);
                else tma_load_3d(dst + li * kPanelBytes, m, full, li * 64, hkv, page);
              }
            } else if (kind == 2) {
              if ((c & 1) == 0 && li < NP) {
                const CUtensorMap* m = fresh ? (kv == 0 ? &p.tmap_nk16 : &p.tmap_nv16) : (kv == 0 ? &p.tmap_k16 : &p.tmap_v16);
                if (shared) tma_load_3d_mc(dst + li * kPanelBytes, m, full, li * 64, hkv, page, 0x3);
                else tma_load_3d(dst + li * kPanelBytes, m, full, li * 64, hkv, page);
              }
            } else if (kind == 1) {
              if (li < NP) {
                const CUtensorMap* m = fresh ? (kv == 0 ? &p.tmap_nk8 : &p.tmap_nv8) : (kv == 0 ? &p.tmap_k8 : &p.tmap_v8);
                if (shared) tma_load_3d_mc(dst + li * kPanelBytes, m, full, li * 64, hkv, page, 0x3);
                else tma_load_3d(dst + li * kPanelBytes, m, full, li * 64, hkv, page);
              }
            } else if (li < 2 * NP) {
              // (the builder keeps this step's tokens in chunks of their own: four rows are all fresh -- or dummies -- or none)
              const int row = w * 64 + c * 8 + g4 * 4;   // my four rows of the tile
              constexpr int kFresh = 1 << 30;
              auto is_fresh = [](int v) { return v >= 0 && (v & kFresh) != 0; };   // (a dummy is -1: every bit set)
              const bool fr = is_fresh(pg.x) || is_fresh(pg.y) || is_fresh(pg.z) || is_fresh(pg.w);
              const int ratio = fr ? p.new_row_ratio : p.kv_row_ratio, oob = fr ? p.new_rows : p.kv_rows;
              auto row_of = [&](int v, int i) { return row + i < tlen && v >= 0 ? (v & ~kFresh) * ratio + hkv : oob; };
              const int r0 = row_of(pg.x, 0), r1 = row_of(pg.y, 1), r2 = row_of(pg.z, 2), r3 = row_of(pg.w, 3);
              const CUtensorMap* m = fr ? (kv == 0 ? &p.tmap_nkg : &p.tmap_nvg) : mg;
              if (shared) tma_gather4_mc(dst + pn * kPanelBytes + g4 * 4 * 128, m, full, pn * 64, r0, r1, r2, r3, 0x3);
              else tma_gather4(dst + pn * kPanelBytes + g4 * 4 * 128, m, full, pn * 64, r0, r1, r2, r3);
            }
          }
          mbar_arrive(full);
          if (w == 0 && lane == 0 && ji == 0 && t < 6) DEFT_TRACE(kTrTile0 + 8 * t + (kv == 0 ? 0 : 6));
        }
      } else {
        // ---- (c)
        auto page_of = [&](int t, int row) -> int {  // page of `row` of tile t (-1 past the end)
          const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
          return t < u.n_tiles && row < tlen ? (int)load_index(p.u_kv, p.u_kv_bytes, u.kv_off + (int64_t)t * u.kv_tile_stride + row) : -1;
        };
        int pg_next0 = page_of(0, w * 64 + lane), pg_next1 = page_of(0, w * 64 + 32 + lane);
        dep_wait();
        for (int t = 0; t < u.n_tiles; ++t, ++cnt) {
          const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
          const int st = cnt % stages;
          const int pg[2] = {pg_next0, pg_next1};
          pg_next0 = page_of(t + 1, w * 64 + lane);  // the next tile's page ids are in flight while this tile is issued
          pg_next1 = page_of(t + 1, w * 64 + 32 + lane);
          mbar_wait<64>(bar(EMPTY + st), ((cnt / stages) & 1) ^ 1);
          const uint32_t dst_base = base + (kv == 0 ? L::kK : L::kV) + st * L::kOperandBytes;
          const uint32_t full = bar(FULL + st);
          bool run[2];
          uint32_t tx = 0;  // bytes the TMA engine completes on MY barrier for my 64 rows (whichever CTA issues them)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int page0 = __shfl_sync(0xffffffffu, pg[b], 0);
            run[b] = __all_sync(0xffffffffu, p.tma_kv != 0 && tlen == kTileN && pg[b] >= 0 && pg[b] == page0 + lane);
            if (run[b] || p.tma_gather != 0) tx += 32 * D * 2;
          }
          if (lane == 0 && tx != 0) mbar_expect_tx(full, tx);
          bool any_cp_async = false;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int row0 = w * 64 + b * 32;  // this block's 32 rows
            const bool ok = row0 + lane < tlen && pg[b] >= 0;
            const int page0 = __shfl_sync(0xffffffffu, pg[b], 0);
            if (run[b]) {
              if (mine && lane < NP) {
                if (shared) tma_load_3d_mc(dst_base + lane * kPanelBytes + row0 * 128, m32, full, lane * 64, hkv, page0, 0x3);
                else tma_load_3d(dst_base + lane * kPanelBytes + row0 * 128, m32, full, lane * 64, hkv, page0);
              }
            } else if (p.tma_gather != 0) {
              // lane (g, panel) moves the four rows 4g .. 4g+3 of the block with one gather4 per panel
              const int g = lane / NP, pn = lane % NP;
              const int my_row = ok ? pg[b] * p.kv_row_ratio + hkv : p.kv_rows;
              const int r0 = __shfl_sync(0xffffffffu, my_row, (4 * g) & 31), r1 = __shfl_sync(0xffffffffu, my_row, (4 * g + 1) & 31);
              const int r2 = __shfl_sync(0xffffffffu, my_row, (4 * g + 2) & 31), r3 = __shfl_sync(0xffffffffu, my_row, (4 * g + 3) & 31);
              if (mine && lane < 8 * NP) {
                const uint32_t dst = dst_base + pn * kPanelBytes + (row0 + 4 * g) * 128;
                if (shared) tma_gather4_mc(dst, mg, full, pn * 64, r0, r1, r2, r3, 0x3);
                else tma_gather4(dst, mg, full, pn * 64, r0, r1, r2, r3);
              }
            } else {
              const __half* src_base = (kv == 0 ? p.k : p.v) + (int64_t)hkv * p.kv_head_stride;
              constexpr int TOK_PER_INSTR = 32 / CH;  // tokens covered by one warp-wide copy
#pragma unroll 4
              for (int i = 0; i < 32 / TOK_PER_INSTR; ++i) {
                const int nl = i * TOK_PER_INSTR + lane / CH;  // row inside the block
                const int ch = lane % CH;
                const int64_t page = __shfl_sync(0xffffffffu, pg[b], nl);
                const bool okr = row0 + nl < tlen && page >= 0;
                cp_async_16(dst_base + tile_off(row0 + nl, ch), src_base + (okr ? page : 0) * p.kv_tok_stride + ch * 8, okr ? 16u : 0u);
              }
              any_cp_async = true;
            }
          }
          if (any_cp_async) cp_async_arrive(full);  // arrives once my copies have landed
          else mbar_arrive(full);
          if (w == 0 && lane == 0 && ji == 0 && t < 6) DEFT_TRACE(kTrTile0 + 8 * t + (kv == 0 ? 0 : 6));
        }
      }
    }
  } else if (warp == kQWarp) {
    // ============================== Q tile of the job's slot ==============================
    uint32_t q_cnt = 0;  // jobs
    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      // row r = (query r / G, head r % G); rows past q_cnt*G are zero
      const int n_q = k == 0 ? u.q_cnt[0] : u.q_cnt[1];
      const int q_id0 = k == 0 ? u.q_id0[0] : u.q_id0[1];
      const int q_off = k == 0 ? u.q_off[0] : u.q_off[1];
      const bool known_run = q_id0 >= 0 && p.tma_q != 0;  // the builder's shortcut: no query-table read
      const int64_t my_q = known_run ? (int64_t)q_id0 + lane : (lane < n_q ? load_index(p.u_q, p.u_q_bytes, q_off + lane) : 0);
      mbar_wait<64>(bar(Q_EMPTY), (q_cnt & 1) ^ 1);
      const uint32_t qs = base + L::kQ;
      if (lane == 0 && ji == 0 && my_q >= 0) DEFT_TRACE(kTrQIds);
      // consecutive query ids: the slot's G heads x 32 queries are ONE box of q's tensor map per panel
      // (rows past q_cnt then hold the next queries or zeros: finite, never stored)
      const int64_t q0 = __shfl_sync(0xffffffffu, my_q, 0);
      const bool run = p.tma_q != 0 && (lane >= n_q || my_q == q0 + lane);
      if (known_run || __all_sync(0xffffffffu, run)) {
        if (lane == 0) {
          mbar_arrive_expect_tx(bar(Q_FULL), R * D * 2);
#pragma unroll
          for (int pn = 0; pn < D / 64; ++pn)
            tma_load_3d(qs + pn * kPanelBytes, &p.tmap_q, bar(Q_FULL), pn * 64, hkv * G, (int)q0);
        } else {
          mbar_arrive(bar(Q_FULL));
        }
      } else {
#pragma unroll 4
        for (int i = 0; i < kRows * CH / 32; ++i) {
          const int c = lane + i * 32;
          const int r = c / CH, ch = c % CH;
          const int qi = r / G, g = r % G;
          const int64_t qid = __shfl_sync(0xffffffffu, my_q, qi & 31);
          const bool ok = qi < n_q;
          const __half* src = p.q + qid * p.q_row_stride + (int64_t)(hkv * G + g) * p.q_head_stride + ch * 8;
          cp_async_16(qs + tile_off(r, ch), ok ? src : p.q, ok ? 16u : 0u);
        }
        cp_async_arrive(bar(Q_FULL));
      }
      if (lane == 0 && ji == 0) DEFT_TRACE(kTrQIssued);
      ++q_cnt;
    }
  } else if (warp == kMaskWarp) {
    // ============================== mask words + dense flag per tile ==============================
    uint32_t m_cnt = 0;  // tiles
    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      const int n_q = k == 0 ? u.q_cnt[0] : u.q_cnt[1];
      const int64_t mask_off = k == 0 ? u.mask_off[0] : u.mask_off[1];
      if (mask_off < 0 && u.last_len == kTileN) continue;  // every tile dense: the softmax warps do not ask
      const uint32_t fullw = n_q >= 32 ? 0xffffffffu : ((1u << n_q) - 1u);
      for (int t = u.dense_tiles; t < u.n_tiles; ++t, ++m_cnt) {   // (the unit's leading dense tiles: nobody asks)
        const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
        const int st = m_cnt % kMaskStages;
        // per-token words: bit r = row r of the slot attends token lane + 32j (loaded ahead of the wait)
        uint32_t m[kTileN / 32];
        bool dense = tlen == kTileN;
#pragma unroll
        for (int j = 0; j < kTileN / 32; ++j) {
          const int n = lane + 32 * j;
          m[j] = 0;
          if (n < tlen)
            m[j] = mask_off >= 0 ? (uint32_t)load_index(p.u_mask, p.u_mask_bytes, mask_off + (int64_t)t * u.mask_tile_stride + n)
                                 : 0xffffffffu;
          dense = dense && ((m[j] & fullw) == fullw);
        }
        dense = __all_sync(0xffffffffu, dense);
        mbar_wait<64>(bar(M_EMPTY + st), ((m_cnt / kMaskStages) & 1) ^ 1);
        uint32_t* ms = reinterpret_cast<uint32_t*>(gbase + L::kMask) + st * kTileN;
        if (!dense) {
          // transpose to row masks: lane = query, word j bit n = the query attends token 32j + n
#pragma unroll 1
          for (int j = 0; j < kTileN / 32; ++j) {
            const uint32_t w = j == 0 ? m[0] : j == 1 ? m[1] : j == 2 ? m[2] : m[3];
            ms[lane * 4 + j] = warp_transpose32(w, lane);
          }
        }
        if (lane == 0) reinterpret_cast<uint32_t*>(gbase + L::kFlag)[st] = dense ? 1u : 0u;
        mbar_arrive(bar(M_FULL + st));
        if (lane == 0 && ji == 0 && t == 0) DEFT_TRACE(kTrMask0);
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== S issuer ==============================
    // The whole warp runs the (uniform) control flow and the waits; lane 0 alone executes the tcgen05.mma /
    // tcgen05.commit instructions.  S(t) = Q K(t)^T goes into the S buffer of the group that will work tile t, as soon
    // as K(t) has landed and the group has taken its previous tile's row out of the buffer -- i.e. while that group is
    // still in the exponentials of its previous tile.
    const bool leader = lane == 0;
    const uint64_t q_desc = smem_desc_sw128(base + L::kQ, 16, 1024);
    uint32_t k_cnt = 0, task = 0, j_cnt = 0, tc0 = 0, tc1 = 0;  // K tiles consumed, tasks, jobs, tile tasks per group
    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      const int n = u.n_tiles;
      const bool tr0 = ji == 0 && leader;
      mbar_wait(bar(Q_FULL), j_cnt & 1);
      if (tr0) DEFT_TRACE(kTrMmaQFull);
      for (int t = 0; t < n; ++t, ++task) {
        const int g = task & 1, st = (k_cnt + t) % kKStages;
        const uint32_t c = g == 0 ? tc0++ : tc1++;
        if (c >= 1) mbar_wait(bar(S_FREE + g), (c - 1) & 1);
        mbar_wait(bar(K_FULL + st), ((k_cnt + t) / kKStages) & 1);
        tc_fence_after();
        const uint32_t s_tmem = tmem + kColS + g * 128;
        const uint64_t k_desc = smem_desc_sw128(base + L::kK + st * L::kOperandBytes, 16, 1024);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint64_t koff = (uint64_t)(((ks >> 2) * kPanelBytes + (ks & 3) * 32) >> 4);
            umma_ss(s_tmem, q_desc + koff, k_desc + koff, kIdescQK, ks > 0);
          }
          umma_commit(bar(S_FULL + g));
          if (shared) {
            umma_commit_mc(bar(K_EMPTY + st), 0x3);  // K(t) is free here; the pair's producers hear it too
          } else {
            umma_commit(bar(K_EMPTY + st));
            umma_commit(bar(K_EMPTY + st));
          }
          if (t == n - 1) umma_commit(bar(Q_EMPTY));
        }
        __syncwarp();
        if (tr0 && t < 6) DEFT_TRACE(kTrTile0 + 8 * t + 1);
      }
      ++task;  // the job's epilogue task
      k_cnt += n;
      ++j_cnt;
    }
  } else if (warp == kPvWarp) {
    // ============================== P V issuer ==============================
    // O (+)= P(t) V(t) in two 64-token halves, each as soon as the tile's group has written that half of P.
    const bool leader = lane == 0;
    const uint32_t o_tmem = tmem;
    uint32_t v_cnt = 0, task = 0, j_cnt = 0, tc0 = 0, tc1 = 0;
    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      const int n = u.n_tiles;
      const bool tr0 = ji == 0 && leader;
      for (int t = 0; t < n; ++t, ++task) {
        const int g = task & 1, st = (v_cnt + t) % kVStages;
        const uint32_t c = g == 0 ? tc0++ : tc1++;
        mbar_wait(bar(V_FULL + st), ((v_cnt + t) / kVStages) & 1);
        if (t == 0 && j_cnt > 0) mbar_wait(bar(O_EMPTY), (j_cnt - 1) & 1);  // the previous job's epilogue has read O
        const uint64_t v_desc = smem_desc_sw128(base + L::kV + st * L::kOperandBytes, kPanelBytes, 1024);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(bar(P_FULL + 2 * g + half), c & 1);
          tc_fence_after();
          const uint32_t p_tmem = tmem + kColP + g * 64 + half * 32;  // 64 tokens = 32 columns of packed fp16 pairs
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < kHalfN / 16; ++ks)
              umma_ts(o_tmem, p_tmem + ks * 8, v_desc + (uint64_t)(((half * kHalfN + ks * 16) * 128) >> 4), kIdescPV,
                      t > 0 || half > 0 || ks > 0);
          }
        }
        if (leader) {
          if (shared) {
            umma_commit_mc(bar(V_EMPTY + st), 0x3);
          } else {
            umma_commit(bar(V_EMPTY + st));
            umma_commit(bar(V_EMPTY + st));
          }
          umma_commit(bar(P_FREE + g));
          if (t == n - 1) umma_commit(bar(O_DONE));
        }
        __syncwarp();
        if (tr0 && t < 6) DEFT_TRACE(kTrTile0 + 8 * t + 7);
      }
      ++task;
      v_cnt += n;
      ++j_cnt;
    }
  }
  } else {
    reg_alloc<kSoftmaxRegs>();   // warps 0-7
    // ============================== softmax groups + epilogue ==============================
    const int g = warp >> 2;     // my group: tasks with (task & 1) == g are mine
    const int r = tid & 127;     // my row == my TMEM lane
    const int qi = r / G;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + kColS + g * 128, t_p = t_lane + kColP + g * 64;
    const float c = p.scale * 1.4426950408889634f;  // scores are handled in the log2 domain
    volatile float* m_ref = reinterpret_cast<volatile float*>(gbase + L::kMref);
    volatile float* lse_s = reinterpret_cast<volatile float*>(gbase + L::kLse);  // [job parity][group][row]
    // tasks, jobs, MY tile tasks, THEIR tile tasks, masked tiles (of both groups)
    uint32_t task = 0, j_cnt = 0, tcnt = 0, ocnt = 0, m_cnt = 0;
    bool first_job = blockIdx.x == 0;

    // task k waits until task k - 1 (the other group's) has read / raised the shared reference maximum
    auto wait_checked = [&](uint32_t k) {
      if (k > 0) mbar_wait<32>(bar(CHK + (g ^ 1)), ((k - 1) >> 1) & 1);
    };

    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      const int n = u.n_tiles;
      const int n_q = k == 0 ? u.q_cnt[0] : u.q_cnt[1];
      const int part_base = k == 0 ? u.part_base[0] : u.part_base[1];
      const bool dbg = kDbg && p.dbg != nullptr && first_job;  // (the dumps live in an instantiation of their own)
      first_job = false;
      const int64_t mask_off = k == 0 ? u.mask_off[0] : u.mask_off[1];
      const bool job_dense = mask_off < 0 && u.last_len == kTileN;  // no tile of this job needs a mask
      const uint32_t jp = j_cnt & 1;
      const uint32_t task0 = task;                                    // task of tile 0
      const int my_last = ((task0 + n - 1) & 1) == (uint32_t)g ? n - 1 : n - 2;  // my last tile of the job (-1: none)
      float m_loc = -INFINITY, l_run = 0.f;   // the reference maximum my sum is kept against, my sum over MY tiles

      for (int t = 0; t < n; ++t, ++task) {
        if ((task & 1) != (uint32_t)g) {      // the other group's tile
          if (!job_dense && t >= u.dense_tiles) ++m_cnt;
          ++ocnt;
          continue;
        }
        const uint32_t cn = tcnt++;
        const bool tr = ji == 0 && (tid & 127) == 0 && t < 6;
        float sv[kTileN];   // my S row: out of TMEM once, kept in registers
        mbar_wait<32>(bar(S_FULL + g), cn & 1);
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < kTileN / 32; ++cb) tmem_ld32_nowait(t_s + cb * 32, sv + cb * 32);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(bar(S_FREE + g));   // the S issuer may compute my next tile into the buffer
        if (tr) DEFT_TRACE(kTrTile0 + 8 * t + 2);
        if (dbg && t == 0)
          for (int j = 0; j < kTileN; ++j) p.dbg[r * kTileN + j] = sv[j];

        // ---- mask: tokens my query does not attend score -inf (the unit's leading dense tiles -- a prompt ahead of the
        // subtree -- carry none)
        if (!job_dense && t >= u.dense_tiles) {
          const int mst = m_cnt % kMaskStages;
          mbar_wait<32>(bar(M_FULL + mst), (m_cnt / kMaskStages) & 1);
          ++m_cnt;
          const uint32_t* ms = reinterpret_cast<const uint32_t*>(gbase + L::kMask) + mst * kTileN;
          const bool dense = reinterpret_cast<const volatile uint32_t*>(gbase + L::kFlag)[mst] != 0;
          if (!dense) {
            const uint4 rm = qi < 32 ? *reinterpret_cast<const uint4*>(ms + qi * 4) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int j = 0; j < kTileN; ++j) {
              const uint32_t w = (j >> 5) == 0 ? rm.x : (j >> 5) == 1 ? rm.y : (j >> 5) == 2 ? rm.z : rm.w;
              if (!((w >> (j & 31)) & 1u)) sv[j] = -INFINITY;
            }
          }
          mbar_arrive(bar(M_EMPTY + mst));
        }
        // ---- my row's maximum (thread-local)
        float mt;
        {
          float m0 = sv[0], m1 = sv[1], m2 = sv[2], m3 = sv[3];
#pragma unroll
          for (int j = 4; j < kTileN; j += 4) {
            m0 = fmaxf(m0, sv[j]); m1 = fmaxf(m1, sv[j + 1]); m2 = fmaxf(m2, sv[j + 2]); m3 = fmaxf(m3, sv[j + 3]);
          }
          mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;  // c > 0; -inf when the row attends nothing here
        }
        // ---- the shared reference maximum, in task order
        wait_checked(task);
        const float m_cur = t == 0 ? -INFINITY : m_ref[r];
        if (m_cur != m_loc) {          // another tile of the row raised it since my last tile
          l_run *= rescale(m_loc, m_cur);
          m_loc = m_cur;
        }
        if (t == 0) {
          m_ref[r] = mt;
          m_loc = mt;
        } else {
          const bool raise = mt > m_cur + kRaise;
          // tcgen05.ld / st are warp-wide: if any row of my warp raises, the whole warp rewrites its 32 rows of O
          // (the rows that do not raise with a factor of one)
          if (__any_sync(0xffffffffu, raise)) {
            // P V of every earlier tile of the job has landed in O once the OTHER group's P buffer is free again: tile
            // t - 1 is theirs (their tile task ocnt - 1), and the tensor pipe completes in order.  Their next tile's
            // P V cannot complete before mine, so the barrier is at most this one phase ahead: a sound parity wait.
            mbar_wait<64>(bar(P_FREE + (g ^ 1)), (ocnt - 1) & 1);
            tc_fence_after();
            const float alpha = !raise ? 1.f : m_cur == -INFINITY ? 0.f : fast_exp2(m_cur - mt);
            float ov[32];
#pragma unroll 1
            for (int cb = 0; cb < D / 32; ++cb) {
              tmem_ld32(t_lane + cb * 32, ov);
#pragma unroll
              for (int j = 0; j < 32; ++j) ov[j] *= alpha;
              tmem_st32(t_lane + cb * 32, ov);
            }
            tmem_wait_st();
            tc_fence_before();
            if (raise) {
              l_run *= alpha;
              m_ref[r] = mt;
              m_loc = mt;
            }
          }
        }
        mbar_arrive(bar(CHK + g));
        if (tr) DEFT_TRACE(kTrTile0 + 8 * t + 3);

        // ---- P = exp2(S c - m_ref), 64 tokens at a time; each half goes to the tensor pipe as soon as it is in TMEM
        const float m_use = m_loc == -INFINITY ? 0.f : m_loc;
        if (cn >= 1) mbar_wait<32>(bar(P_FREE + g), (cn - 1) & 1);   // P V of my previous tile has read my P buffer
        float hs = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t pk[kHalfN / 2];
          float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
#pragma unroll
          for (int j = 0; j < kHalfN; j += 4) {
            const int jj = h * kHalfN + j;
            const float e0 = fast_exp2(fmaf(sv[jj], c, -m_use)), e1 = fast_exp2(fmaf(sv[jj + 1], c, -m_use));
            const float e2 = fast_exp2(fmaf(sv[jj + 2], c, -m_use)), e3 = fast_exp2(fmaf(sv[jj + 3], c, -m_use));
            ps0 += e0; ps1 += e1; ps2 += e2; ps3 += e3;
            pk[j / 2] = pack_half2(e0, e1);
            pk[j / 2 + 1] = pack_half2(e2, e3);
          }
          hs += (ps0 + ps1) + (ps2 + ps3);
          tmem_st32(t_p + h * 32, reinterpret_cast<const float*>(pk));
          if (h == 1 && t == my_last) {
            // my last tile of the job: my share of the row sum goes to the epilogue as a log-sum-exp (log2 domain)
            const float l_fin = l_run + hs;
            lse_s[(jp * 2 + g) * kRows + r] = l_fin > 0.f ? m_loc + log2f(l_fin) : -INFINITY;
          }
          tmem_wait_st();
          tc_fence_before();  // my TMEM stores (P, a rescaled O) are ordered before the MMA issued after the barrier
          mbar_arrive(bar(P_FULL + 2 * g + h));
          if (tr) DEFT_TRACE(kTrTile0 + 8 * t + 4 + h);
        }
        l_run += hs;
      }

      // ---- the job's epilogue task: partial = O / l as fp16, log-sum-exp in the natural-log domain
      if ((task & 1) == (uint32_t)g) {
        wait_checked(task);
        const float m_fin = m_ref[r];        // every tile of the job has passed its check
        mbar_arrive(bar(CHK + g));           // the next job's first tile may overwrite it
        mbar_wait<32>(bar(O_DONE), j_cnt & 1);
        tc_fence_after();
        if (ji == 0 && r == 0) DEFT_TRACE(kTrEpiODone);
        float l_row = 0.f;
        {
          const bool both = n >= 2;
          const int g0 = task0 & 1;          // the group of tile 0
          const float la = (both || g0 == 0) ? lse_s[(jp * 2 + 0) * kRows + r] : -INFINITY;
          const float lb = (both || g0 == 1) ? lse_s[(jp * 2 + 1) * kRows + r] : -INFINITY;
          if (la > -INFINITY) l_row += fast_exp2(la - m_fin);
          if (lb > -INFINITY) l_row += fast_exp2(lb - m_fin);
        }
        if (ji == 0 && r == 0) DEFT_TRACE(kTrEpiBegin);
        const bool live = qi < n_q;
        const float inv = l_row > 0.f ? 1.f / l_row : 0.f;
        const int64_t tile = (int64_t)(part_base >> 5) * p.HKV + hkv;
        uint4* dst = reinterpret_cast<uint4*>(p.po16) + tile * (CH * R) + r;  // [chunk][row] of 16 bytes
#pragma unroll
        for (int hb = 0; hb < D / 64; ++hb) {
          float ov[64];
          tmem_ld32_nowait(t_lane + hb * 64, ov);
          tmem_ld32_nowait(t_lane + hb * 64 + 32, ov + 32);
          tmem_wait_ld();
          if (hb == D / 64 - 1) {
            tc_fence_before();  // my reads of O are ordered before the next job's first P V (accumulate = 0)
            mbar_arrive(bar(O_EMPTY));
          }
          if (dbg)
            for (int j = 0; j < 64; ++j) p.dbg[kRows * kTileN + r * D + hb * 64 + j] = ov[j];
          if (live) {
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              uint4 o4;
              o4.x = pack_half2(ov[c8 * 8 + 0] * inv, ov[c8 * 8 + 1] * inv);
              o4.y = pack_half2(ov[c8 * 8 + 2] * inv, ov[c8 * 8 + 3] * inv);
              o4.z = pack_half2(ov[c8 * 8 + 4] * inv, ov[c8 * 8 + 5] * inv);
              o4.w = pack_half2(ov[c8 * 8 + 6] * inv, ov[c8 * 8 + 7] * inv);
              dst[(hb * 8 + c8) * R] = o4;
            }
          }
        }
        if (live) p.plse16[tile * R + r] = l_row > 0.f ? (m_fin + log2f(l_row)) * 0.6931471805599453f : -INFINITY;
        if (ji == 0 && r == 0) DEFT_TRACE(kTrEpiEnd);
        if (r == 0 && j_cnt < 32 && p.trace != nullptr) {   // per-job record: epilogue end, tiles and kind of the job
          DEFT_TRACE(kTrJob0 + 2 * (int)j_cnt);
          p.trace[(int64_t)blockIdx.x * kTraceSlots + kTrJob0 + 2 * (int)j_cnt + 1] = n | ((u.page0 >= 0) << 16) | ((int)shared << 17) | ((int)job_dense << 18);
        }
      }
      ++task;
      ++j_cnt;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.clustered) cluster_sync();  // the pair no longer multicasts into my shared memory or arrives on my barriers
  if (tid == 0) DEFT_TRACE(kTrEnd);
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

template <int D, int G>
int launch_t(const AttnParams& p, cudaStream_t stream) {
  static PerDeviceOnce once;  // per device: the attribute belongs to the current device's copy of the function
  using L = Layout<D>;
  const int dev = current_device_index();
  int num_sms = once.slot[dev];
  if (num_sms == 0) {
    DEFT_CUDA(cudaFuncSetAttribute(stage1_umma_kernel<D, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    DEFT_CUDA(cudaFuncSetAttribute(stage1_umma_kernel<D, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    DEFT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    once.slot[dev] = num_sms;
  }
  int grid;
  if (p.job_off != nullptr) {
    grid = p.n_ctas;
  } else {
    const int64_t n_jobs = (int64_t)p.n_units * p.HKV * 2;
    grid = (int)(n_jobs < num_sms ? n_jobs : num_sms);
  }
  if (grid <= 0) return DEFT_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L::kAlloc;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (p.pdl) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // launch early, wait inside (griddep_wait)
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  AttnParams pl = p;
  pl.clustered = p.job_off != nullptr && grid % 2 == 0 && p.clustered;  // CTA pairs (2c, 2c + 1): see deft_job_t.shared
  if (pl.clustered) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = 2;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  if (pl.dbg != nullptr) DEFT_CUDA(cudaLaunchKernelEx(&cfg, stage1_umma_kernel<D, G, true>, pl));
  else DEFT_CUDA(cudaLaunchKernelEx(&cfg, stage1_umma_kernel<D, G, false>, pl));
  return DEFT_OK;
}

}  // namespace

bool stage1_umma_supported(const AttnParams& p) {
  const int G = p.H / p.HKV;
  return (p.D == 128 || p.D == 64) && (G == 1 || G == 2 || G == 4);
}

int launch_stage1_umma(const AttnParams& p, cudaStream_t stream) {
  if (p.n_units <= 0) return DEFT_OK;
  const int G = p.H / p.HKV;
#define DEFT_CASE(DD, GG) \
  if (p.D == DD && G == GG) return launch_t<DD, GG>(p, stream);
  DEFT_CASE(128, 4) DEFT_CASE(128, 2) DEFT_CASE(128, 1) DEFT_CASE(64, 4) DEFT_CASE(64, 2) DEFT_CASE(64, 1)
#undef DEFT_CASE
  set_error("tcgen05 stage 1 does not cover head_dim %d / GQA group %d", p.D, G);
  return DEFT_E_ARG;
}

}  // namespace deft
