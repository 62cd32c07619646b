// Stage 1, tensor-core path (sm_100a): tcgen05.mma with TMEM accumulators, warp-specialised
// persistent CTAs walking the unit plan.
//
// One job = (unit, kv-head).  A unit is a chain of KV tiles (128 tokens x D, K and V) attended by one
// or two *slots* of <= 32 queries; with G = H/HKV query heads per kv-head a slot is one M = 128 UMMA
// tile (row r = query r/G, head r%G).  Each KV tile is gathered from the token-granular paged pool
// ONCE into 128B-swizzled shared memory and serves both slots and all G heads:
//
//   S_s[128 x 128]  = Q_s[128 x D] . K^T        tcgen05.mma kind::f16, SS: A/B K-major SW128 smem -> TMEM
//   P_s             = exp2(S_s*c - m_ref), masked by the per-token row bitmask; written back over S_s
//                     in TMEM as packed fp16 (the A operand of the next MMA never touches smem)
//   O_s[128 x D]   += P_s[128 x 128] . V        tcgen05.mma TS: A = P_s in TMEM, B = V (MN-major SW128 smem)
//
// The chain is walked with an online softmax whose reference maximum m_ref is only raised when a tile
// exceeds it by more than 2^8 (P stays <= 256 in fp16, O and l stay consistent), so the accumulator in
// TMEM is almost never rescaled.  ONE partial (O/l as fp16, log-sum-exp as fp32) leaves the SM per
// (job, slot); stage 2 (combine.cu) merges the partials of every query.
//
// Warp roles (512 threads, 1 CTA per SM, all 512 TMEM columns):
//   warps 0-3   softmax + epilogue of slot 0 (thread = row = TMEM lane; 200 registers via setmaxnreg:
//               the whole 128-column S row is read from TMEM once and kept in registers)
//   warps 4-7   softmax + epilogue of slot 1; the two slots ping-pong on the tensor pipe:
//               S_0(t) S_1(t) PV_0(t) | S_0(t+1) PV_1(t) S_1(t+1) PV_0(t+1) | ...
//   warp  8/9   MMA issuer of slot 0 / slot 1 (one elected thread each; warp 8 also owns the TMEM allocation).
//               One issuer per slot: a single thread's issue stream (~110 cycles per MMA with its waits and
//               commits) was the serial bottleneck of a two-slot step
//   warp  10    Q tiles of both slots: one TMA box per 64-wide panel when a slot's query ids are
//               consecutive, else cp.async 16-byte gathers
//   warp  11    per-(tile, slot) row masks (token bitmask per query, transposed from the per-token
//               words of the table) + "dense tile" flag
//   warps 12-15 K / V producers: each warp owns 32 token rows of every tile -- one TMA box per panel
//               when its 32 pages are consecutive (prompt), else cp.async 16-byte gathers (the in-flight
//               depth of cp.async is per warp, hence four warps)
// All hand-offs are mbarriers (cp.async arrive-on, tcgen05.commit, plain arrive); no __syncthreads in
// the steady state.
//
// Reference semantics: DeFT/deft/layers/attention/tree_attention.py:860-976 (Flatten stage 1) and
// :170-293 (Node stage 1).
#include "umma_ptx.cuh"

namespace deft {
namespace {

constexpr int kThreads = 640;  // 20 warps: 5 register-budget groups of 4 (setmaxnreg works per warpgroup)
constexpr int kMmaWarp = 8, kQWarp = 9, kMaskWarp = 10, kPvWarp = 11, kKvWarp0 = 12;  // 12-15: K producers; 16-19: V
// a gathering warp is bound by its copies in flight (~8 x 512 bytes against the memory latency), so scattered
// pages want many producer warps: four per operand, 32 rows each
constexpr int kSoftmaxRegs = 152, kProducerRegs = 56;  // 256 * 152 + 384 * 56 <= 64 K registers (launch: 96 each)
constexpr int kKStages = 3, kVStages = 2, kMaskStages = 2;
constexpr int kSBufs = 3;  // S tiles in TMEM: O [0, 128) + 3 x 128 columns = all 512

using namespace umma;

// Optional per-CTA timeline (test/profiling hook, deft_b200_set_trace_buffer): trace[cta][event] =
// SM cycles since the CTA started.  Events: see kTrace* below; per-tile events take 8 slots per tile.
constexpr int kTraceSlots = 128;
enum : int {
  kTrStart = 0, kTrQIds = 1, kTrQ0Issued = 2, kTrQ1Issued = 3, kTrMask0 = 4, kTrKUnit = 5, kTrMmaQFull = 6, kTrEpiBegin = 7,
  kTrEpiEnd = 8, kTrEnd = 9,
  kTrTile0 = 16,  // + 8 * tile: K issued, K_FULL seen by MMA, S_FULL seen by softmax 0, pass 1 done, P_FULL arrive,
                  //             P_FULL seen by MMA, V issued, (spare)
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
  S_FULL = M_EMPTY + kMaskStages,          // [S buffer]: S of one tile is in TMEM
  P_FULL = S_FULL + kSBufs,                // [S buffer][half]: P of one 64-token half has been written over S
  S_FREE = P_FULL + 2 * kSBufs,            // [S buffer]: P V of the buffer's tile has completed
  PV_DONE = S_FREE + kSBufs,               // one phase per tile: P V of the tile has landed in O
  O_DONE = PV_DONE + 1,                    // one phase per job: the last P V has landed, O is complete
  O_EMPTY = O_DONE + 1,                    // one phase per job: the epilogue has read O
  kNumBars = O_EMPTY + 1
};

template <int D>
struct Layout {
  static constexpr int kOperandBytes = kRows * D * 2;                      // Q, K or V tile
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kOperandBytes;                            // [stage]
  static constexpr int kV = kK + kKStages * kOperandBytes;                 // [stage]
  static constexpr int kMask = kV + kVStages * kOperandBytes;              // [stage][128] u32
  static constexpr int kFlag = kMask + kMaskStages * kTileN * 4;           // [stage] u32
  static constexpr int kPvCnt = kFlag + 16;                                // u32: tiles whose P V the issuer has seen complete
  static constexpr int kXchg = kPvCnt + 16;                                 // [tile parity][round parity][half][128] f32
  static constexpr int kBars = kXchg + 2 * 2 * 2 * kRows * 4;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kBytes = kTmemSlot + 16;
  static constexpr int kAlloc = kBytes + 1024;  // slack for the manual 1024-byte alignment
};

// The jobs of one CTA.  job = ((unit * HKV + kv-head) << 1) | slot of the unit's pair: one CTA works ONE
// slot (<= 32 queries x G heads = one M = 128 accumulator) over the unit's chain of KV tiles; the two
// slots of a pair are separate jobs, on different SMs when the balance allows (their K/V tile reads meet
// in L2).  Either the host-balanced record lists (deft_job_t: the CTA's first record sits at [blockIdx.x]
// and carries its unit, so a CTA starts from ONE load), or jobs c, c + grid, ... over the unit table.
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

// pair of warps w, w + 4 (the two threads of a row sit in them): named barrier 1 + (w & 3)
__device__ __forceinline__ void pair_sync(int warp) {
  asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
}

// ... and the same barrier OR-reducing a predicate over the 64 threads
__device__ __forceinline__ bool pair_sync_or(int warp, bool pred) {
  uint32_t out;
  asm volatile(
      "{ .reg .pred p, q; setp.ne.b32 q, %2, 0; barrier.cta.red.or.pred p, %1, 64, q; selp.u32 %0, 1, 0, p; }"
      : "=r"(out)
      : "r"(1 + (warp & 3)), "r"((uint32_t)pred)
      : "memory");
  return out != 0;
}

template <int D, int G, bool kDbg>
__global__ void __launch_bounds__(kThreads, 1) stage1_umma_v2_kernel(const __grid_constant__ AttnParams p) {
  using L = Layout<D>;
  constexpr int CH = D / 8;           // 16-byte chunks per row
  constexpr int R = kMaxGroupQ * G;   // live rows of a full slot
  constexpr uint32_t kTmemCols = 512; // O [0, D)   S_0 [128, 256)   S_1 [256, 384)   S_2 [384, 512)
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
    for (int s = 0; s < kKStages; ++s) { mbar_init(bar(K_FULL + s), 128); mbar_init(bar(K_EMPTY + s), 2); }  // EMPTY: my issuer + the pair's (or mine twice)
    for (int s = 0; s < kVStages; ++s) { mbar_init(bar(V_FULL + s), 128); mbar_init(bar(V_EMPTY + s), 2); }
    mbar_init(bar(Q_FULL), 32); mbar_init(bar(Q_EMPTY), 1);
    for (int m = 0; m < kMaskStages; ++m) { mbar_init(bar(M_FULL + m), 32); mbar_init(bar(M_EMPTY + m), 256); }
    for (int b = 0; b < kSBufs; ++b) {
      mbar_init(bar(S_FULL + b), 1); mbar_init(bar(S_FREE + b), 1);
      for (int h = 0; h < 2; ++h) mbar_init(bar(P_FULL + 2 * b + h), 128);
    }
    mbar_init(bar(PV_DONE), 1); mbar_init(bar(O_DONE), 1); mbar_init(bar(O_EMPTY), 256);
    *reinterpret_cast<volatile uint32_t*>(gbase + L::kPvCnt) = 0u;
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
  // the tail of the preceding kernel; q, the KV pool and the partial workspace may still be in its hands.
  griddep_wait();
  if (tid == 0) DEFT_TRACE(kTrStart);
  if (warp >= 8) {
  reg_dealloc<kProducerRegs>();  // warps 8-19: three whole warpgroups give registers away
  if (warp >= kKvWarp0) {
    // ============================== K / V producers ==============================
    // warps 12-15: K rows [32w, 32w + 32) of every tile; warps 16-19: V likewise.  K and V run on rings of their
    // own (K is released as soon as S is done, a tile earlier than V).
    const int kv = (warp - kKvWarp0) >> 2, w = (warp - kKvWarp0) & 3;
    const int stages = kv == 0 ? kKStages : kVStages;
    const int FULL = kv == 0 ? K_FULL : V_FULL, EMPTY = kv == 0 ? K_EMPTY : V_EMPTY;
    const int row0 = w * 32;  // my 32 rows
    uint32_t cnt = 0;  // tiles produced
    if (warp == kKvWarp0 && lane == 0) DEFT_TRACE(12);
    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      if (warp == kKvWarp0 && lane == 0 && ji == 0 && u.n_tiles > 0) DEFT_TRACE(kTrKUnit);
      const bool known_run = u.page0 >= 0 && p.tma_kv != 0;  // the builder's shortcut: no index-table read at all
      auto page_of = [&](int t) -> int {  // page of my row of tile t (0 past the end)
        const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
        if (known_run) return u.page0 + t * kTileN + row0 + lane;
        return t < u.n_tiles && row0 + lane < tlen ? (int)load_index(p.u_kv, p.u_kv_bytes, u.kv_off + (int64_t)t * u.kv_tile_stride + row0 + lane) : 0;
      };
      int pg_next = page_of(0);
      for (int t = 0; t < u.n_tiles; ++t, ++cnt) {
        const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
        const int st = cnt % stages;
        const uint32_t ph = ((cnt / stages) & 1) ^ 1;
        const int pg = pg_next;
        pg_next = page_of(t + 1);  // the next tile's page ids are in flight while this tile is issued
        const bool trp = kv == 0 && w == 0 && lane == 0 && ji == 0 && t < 6;
        if (trp) DEFT_TRACE(64 + 8 * t + 0);
        mbar_wait<64>(bar(EMPTY + st), ph);
        if (trp) DEFT_TRACE(64 + 8 * t + 1);
        const uint32_t dst_base = base + (kv == 0 ? L::kK : L::kV) + st * L::kOperandBytes;
        const uint32_t full = bar(FULL + st);
        // 32 consecutive pages of a full tile are ONE box of the pool's tensor map per 64-wide panel
        const int page0 = __shfl_sync(0xffffffffu, pg, 0);
        const bool run = known_run || __all_sync(0xffffffffu, p.tma_kv != 0 && tlen == kTileN && pg == page0 + lane);
        // shared job: the pair loads every tile ONCE -- the CTA of rank r issues rows [64r, 64r + 64) for both
        const bool mine = !shared || (uint32_t)(w >> 1) == crank;
        if (run) {
          if (lane == 0) {
            mbar_arrive_expect_tx(full, 32 * D * 2);
            if (mine) {
#pragma unroll
              for (int pn = 0; pn < D / 64; ++pn) {
                if (shared)
                  tma_load_3d_mc(dst_base + pn * kPanelBytes + row0 * 128, kv == 0 ? &p.tmap_k : &p.tmap_v, full, pn * 64, hkv, page0, 0x3);
                else
                  tma_load_3d(dst_base + pn * kPanelBytes + row0 * 128, kv == 0 ? &p.tmap_k : &p.tmap_v, full, pn * 64, hkv, page0);
              }
            }
          } else {
            mbar_arrive(full);
          }
        } else if (p.tma_gather != 0) {
          // scattered pages: lane (g, panel) moves the four rows 4g .. 4g+3 of my 32 with one gather4 per panel;
          // rows past the tile's length name a row outside the map and arrive as zeros
          constexpr int NP = D / 64;
          const int g = lane / NP, pn = lane % NP;
          const int my_row = row0 + lane < tlen ? pg * p.kv_row_ratio + hkv : p.kv_rows;
          const int r0 = __shfl_sync(0xffffffffu, my_row, (4 * g) & 31), r1 = __shfl_sync(0xffffffffu, my_row, (4 * g + 1) & 31);
          const int r2 = __shfl_sync(0xffffffffu, my_row, (4 * g + 2) & 31), r3 = __shfl_sync(0xffffffffu, my_row, (4 * g + 3) & 31);
          if (lane == 0) mbar_arrive_expect_tx(full, 32 * D * 2);
          else mbar_arrive(full);
          if (mine && lane < 8 * NP) {
            if (shared)
              tma_gather4_mc(dst_base + pn * kPanelBytes + (row0 + 4 * g) * 128, kv == 0 ? &p.tmap_kg : &p.tmap_vg, full, pn * 64,
                             r0, r1, r2, r3, 0x3);
            else
              tma_gather4(dst_base + pn * kPanelBytes + (row0 + 4 * g) * 128, kv == 0 ? &p.tmap_kg : &p.tmap_vg, full, pn * 64,
                          r0, r1, r2, r3);
          }
        } else {
          const __half* src_base = (kv == 0 ? p.k : p.v) + (int64_t)hkv * p.kv_head_stride;
          constexpr int TOK_PER_INSTR = 32 / CH;  // tokens covered by one warp-wide copy
#pragma unroll 4
          for (int i = 0; i < 32 / TOK_PER_INSTR; ++i) {
            const int nl = i * TOK_PER_INSTR + lane / CH;  // row inside my 32
            const int ch = lane % CH;
            const int64_t page = __shfl_sync(0xffffffffu, pg, nl);
            const bool ok = row0 + nl < tlen;
            cp_async_16(dst_base + tile_off(row0 + nl, ch), src_base + page * p.kv_tok_stride + ch * 8, ok ? 16u : 0u);
          }
          cp_async_arrive(full);
        }
        if (w == 0 && lane == 0 && ji == 0) DEFT_TRACE(kTrTile0 + 8 * t + (kv == 0 ? 0 : 6));
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
      if (lane == 0 && ji == 0) DEFT_TRACE(kTrQ0Issued);
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
      for (int t = 0; t < u.n_tiles; ++t, ++m_cnt) {
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
        if (lane == 0 && ji == 0 && t >= 1 && t <= 3) DEFT_TRACE(120 + 2 * (t - 1));
        mbar_wait<64>(bar(M_EMPTY + st), ((m_cnt / kMaskStages) & 1) ^ 1);
        uint32_t* ms = reinterpret_cast<uint32_t*>(gbase + L::kMask) + st * kTileN;
        if (!dense) {
          // transpose to row masks: lane = query, word j bit n = the query attends token 32j + n
          // (one copy of the transpose in the binary: this warp's loop shares the SM's 32 KB instruction cache with the
          // softmax, producer and issuer loops)
#pragma unroll 1
          for (int j = 0; j < kTileN / 32; ++j) {
            const uint32_t w = j == 0 ? m[0] : j == 1 ? m[1] : j == 2 ? m[2] : m[3];
            ms[lane * 4 + j] = warp_transpose32(w, lane);
          }
        }
        if (lane == 0) reinterpret_cast<uint32_t*>(gbase + L::kFlag)[st] = dense ? 1u : 0u;
        mbar_arrive(bar(M_FULL + st));
        if (lane == 0 && ji == 0 && t == 0) DEFT_TRACE(kTrMask0);
        if (lane == 0 && ji == 0 && t >= 1 && t <= 3) DEFT_TRACE(121 + 2 * (t - 1));
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== S issuer ==============================
    // The whole warp runs the (uniform) control flow and the waits; lane 0 alone executes the tcgen05.mma /
    // tcgen05.commit instructions.  S is triple-buffered in TMEM: S(t) = Q K(t)^T is issued as soon as K(t) has
    // landed and P V of tile t - 3 (the buffer's previous tenant) has completed, i.e. up to two tiles ahead of
    // the softmax warps, which therefore never wait for the tensor pipe in the steady state.
    const bool leader = lane == 0;
    const uint64_t q_desc = smem_desc_sw128(base + L::kQ, 16, 1024);
    uint32_t k_cnt = 0, g0 = 0, j_cnt = 0;  // K tiles consumed (ring position), tiles of earlier jobs, jobs
    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      const int n = u.n_tiles;
      const bool tr0 = ji == 0 && leader;
      mbar_wait(bar(Q_FULL), j_cnt & 1);
      if (tr0) DEFT_TRACE(kTrMmaQFull);
      for (int t = 0; t < n; ++t) {
        const uint32_t gt = g0 + t, c = k_cnt + t;
        const int sb = gt % kSBufs, st = c % kKStages;
        if (gt >= (uint32_t)kSBufs) mbar_wait(bar(S_FREE + sb), (gt / kSBufs - 1) & 1);
        mbar_wait(bar(K_FULL + st), (c / kKStages) & 1);
        tc_fence_after();
        if (tr0) DEFT_TRACE(kTrTile0 + 8 * t + 1);
        const uint32_t s_tmem = tmem + 128 + sb * 128;
        const uint64_t k_desc = smem_desc_sw128(base + L::kK + st * L::kOperandBytes, 16, 1024);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint64_t koff = (uint64_t)(((ks >> 2) * kPanelBytes + (ks & 3) * 32) >> 4);
            umma_ss(s_tmem, q_desc + koff, k_desc + koff, kIdescQK, ks > 0);
          }
          umma_commit(bar(S_FULL + sb));
          if (shared) {
            umma_commit_mc(bar(K_EMPTY + st), 0x3);  // K(t) is free here; the pair's producers hear it too
          } else {
            umma_commit(bar(K_EMPTY + st));
            umma_commit(bar(K_EMPTY + st));
          }
          if (t == n - 1) umma_commit(bar(Q_EMPTY));
        }
        __syncwarp();
      }
      k_cnt += n;
      g0 += n;
      ++j_cnt;
    }
  } else if (warp == kPvWarp) {
    // ============================== P V issuer ==============================
    // O (+)= P(t) V(t) in two 64-token halves, each as soon as the softmax warps have written that half of P
    // over S(t).  After a tile's MMAs and commits this warp sees the tile's PV_DONE phase through and publishes
    // the count of completed tiles (the softmax warps' rare rescale path reads it: a parity wait is sound only
    // one phase ahead of what is known complete, and only this warp sees every phase).
    const bool leader = lane == 0;
    const uint32_t o_tmem = tmem;
    uint32_t v_cnt = 0, g0 = 0, j_cnt = 0;
    volatile uint32_t* pv_cnt = reinterpret_cast<volatile uint32_t*>(gbase + L::kPvCnt);
    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      const int n = u.n_tiles;
      const bool tr0 = ji == 0 && leader;
      for (int t = 0; t < n; ++t) {
        const uint32_t gt = g0 + t, c = v_cnt + t;
        const int buf = gt % kSBufs, st = c % kVStages;
        mbar_wait(bar(V_FULL + st), (c / kVStages) & 1);
        if (t == 0) mbar_wait(bar(O_EMPTY), (j_cnt & 1) ^ 1);
        const uint64_t v_desc = smem_desc_sw128(base + L::kV + st * L::kOperandBytes, kPanelBytes, 1024);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(bar(P_FULL + 2 * buf + half), (gt / kSBufs) & 1);
          tc_fence_after();
          if (tr0 && half == 0) DEFT_TRACE(kTrTile0 + 8 * t + 5);
          const uint32_t p_tmem = tmem + 128 + buf * 128 + half * kHalfN;  // P_a: columns [0, 32), P_b: [64, 96) of S
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < kHalfN / 16; ++ks)
              umma_ts(o_tmem, p_tmem + ks * 8, v_desc + (uint64_t)(((half * kHalfN + ks * 16) * 128) >> 4), kIdescPV,
                      t > 0 || half > 0 || ks > 0);
          }
        }
        if (leader) {
          umma_commit(bar(PV_DONE));
          if (shared) {
            umma_commit_mc(bar(V_EMPTY + st), 0x3);
          } else {
            umma_commit(bar(V_EMPTY + st));
            umma_commit(bar(V_EMPTY + st));
          }
          umma_commit(bar(S_FREE + buf));
          if (t == n - 1) umma_commit(bar(O_DONE));
        }
        __syncwarp();
        mbar_wait(bar(PV_DONE), gt & 1);
        if (leader) *pv_cnt = gt + 1;
      }
      v_cnt += n;
      g0 += n;
      ++j_cnt;
    }
  }
  } else {
    reg_alloc<kSoftmaxRegs>();   // warps 0-7
    // ============================== softmax + epilogue ==============================
    // Two threads per row: warp w < 4 takes columns [0, 64) of every S tile, warp w + 4 columns [64, 128) of
    // the same 32 rows (the same TMEM lanes).  They agree on the row's reference maximum through shared
    // memory and a named barrier of the two warps, once per tile.
    const int h = warp >> 2;
    const int r = tid & 127;  // my row == my TMEM lane
    const int qi = r / G;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_o = t_lane + h * (D / 2);  // my half of the O row
    const float c = p.scale * 1.4426950408889634f;  // scores are handled in the log2 domain
    float* xchg = reinterpret_cast<float*>(gbase + L::kXchg);
    uint32_t g0 = 0, j_cnt = 0, m_cnt = 0;  // tiles of earlier jobs, jobs, masked tiles
    bool first_job = blockIdx.x == 0;

    for (int ji = 0; ji < jobs.n; ++ji) {
      deft_unit_t u; int hkv, k; bool shared;
      if (!jobs.get(p, ji, u, hkv, k, shared)) continue;
      const int n_q = k == 0 ? u.q_cnt[0] : u.q_cnt[1];
      const int part_base = k == 0 ? u.part_base[0] : u.part_base[1];
      const bool dbg = kDbg && p.dbg != nullptr && first_job;  // (the dumps live in an instantiation of their own)
      first_job = false;
      float m_ref = -INFINITY, l_run = 0.f;

      const int64_t mask_off = k == 0 ? u.mask_off[0] : u.mask_off[1];
      const bool job_dense = mask_off < 0 && u.last_len == kTileN;  // no tile of this job needs a mask
      float sv[kHalfN];   // my half of the current S row (64 columns): out of TMEM once, kept in registers
      bool have_next = false;

      for (int t = 0; t < u.n_tiles; ++t) {
        const uint32_t gt = g0 + t;
        const int buf = gt % kSBufs;
        const bool tr = ji == 0 && (tid & 127) == 0 && t < 5;
        const int tr0 = kTrTile0 + (h == 0 ? 0 : 48) + 8 * t;  // the second half's events sit 48 slots higher
        const uint32_t t_s = t_lane + 128 + buf * 128 + h * kHalfN;
        if (have_next) {  // my half of this tile's S has been on its way since the previous tile's P went out
          tmem_wait_ld();
        } else {
          mbar_wait<32>(bar(S_FULL + buf), (gt / kSBufs) & 1);
          tc_fence_after();
#pragma unroll
          for (int cb = 0; cb < kHalfN / 32; ++cb) tmem_ld32_nowait(t_s + cb * 32, sv + cb * 32);
          tmem_wait_ld();
        }
        if (tr) DEFT_TRACE(tr0 + 2);
        if (tr && t == 3 && h == 0 && have_next) DEFT_TRACE(117);   // tile 3's S came out of TMEM ahead of time
        if (dbg && t == 0)
          for (int j = 0; j < kHalfN; ++j) p.dbg[r * kTileN + h * kHalfN + j] = sv[j];
        uint32_t rw[2] = {0xffffffffu, 0xffffffffu};  // my query's token bitmask over my 64 columns (kept: a redo re-masks)
        bool masked = false;                           // CTA-uniform: this tile carries a mask
        auto apply_mask = [&]() {
          if ((rw[0] & rw[1]) != 0xffffffffu) {
#pragma unroll
            for (int j = 0; j < kHalfN; ++j)
              if (!((rw[j >> 5] >> (j & 31)) & 1u)) sv[j] = -INFINITY;
          }
        };
        if (!job_dense) {
          const int mst = m_cnt % kMaskStages;
          mbar_wait<32>(bar(M_FULL + mst), (m_cnt / kMaskStages) & 1);
          if (tr && t == 2 && h == 0) DEFT_TRACE(118);
          ++m_cnt;
          const uint32_t* ms = reinterpret_cast<const uint32_t*>(gbase + L::kMask) + mst * kTileN;
          const bool dense = reinterpret_cast<const volatile uint32_t*>(gbase + L::kFlag)[mst] != 0;
          if (!dense) {  // masked-out tokens score -inf: my query's token bitmask comes from the mask warp
            const uint2 rm = qi < 32 ? *reinterpret_cast<const uint2*>(ms + qi * 4 + h * 2) : make_uint2(0u, 0u);
            rw[0] = rm.x; rw[1] = rm.y;
            if (t == 0) apply_mask();   // the first tile's exact maximum needs S masked ...
            else masked = true;         // ... later tiles mask inside the exp loop (no pass of its own over S)
          }
          if (tr && t == 2 && h == 0) DEFT_TRACE(119);
          mbar_arrive(bar(M_EMPTY + mst));
        }
        auto half_max = [&]() {
          float m0 = sv[0], m1 = sv[1], m2 = sv[2], m3 = sv[3];
#pragma unroll
          for (int j = 4; j < kHalfN; j += 4) {
            m0 = fmaxf(m0, sv[j]); m1 = fmaxf(m1, sv[j + 1]); m2 = fmaxf(m2, sv[j + 2]); m3 = fmaxf(m3, sv[j + 3]);
          }
          return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;  // c > 0; -inf when the row attends nothing here
        };
        // ---- reference maximum m_ref (shared by the row's two threads).  The first tile of a job sets it to
        // the exact row maximum.  Every later tile is exponentiated against the m_ref it finds (no dependent
        // max -> exp chain); only if a half row's exponentials sum past 2^15 (a P might leave fp16) does its
        // thread ask for a raise, and then both threads rescale their half of O and redo the tile.
        int round = 0;
        float* xq = xchg + (gt & 1) * (4 * kRows);  // [round parity][half][row]
        if (t == 0) {
          xq[h * kRows + r] = half_max();
          pair_sync(warp);
          m_ref = fmaxf(xq[r], xq[kRows + r]);
          round = 1;
        }
        if (tr) DEFT_TRACE(tr0 + 3);
        uint32_t pk[kHalfN / 2];
        float hsum;
        bool redo;
        int n_redo = 0;
        do {
          const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
          float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
          if (!masked) {
#pragma unroll
            for (int j = 0; j < kHalfN; j += 4) {
              const float e0 = fast_exp2(fmaf(sv[j], c, -m_use)), e1 = fast_exp2(fmaf(sv[j + 1], c, -m_use));
              const float e2 = fast_exp2(fmaf(sv[j + 2], c, -m_use)), e3 = fast_exp2(fmaf(sv[j + 3], c, -m_use));
              ps0 += e0; ps1 += e1; ps2 += e2; ps3 += e3;
              pk[j / 2] = pack_half2(e0, e1);
              pk[j / 2 + 1] = pack_half2(e2, e3);
            }
          } else {
            // masked tile: P = 0 where my query does not attend the token; the exponential is predicated on the
            // mask bit, so a column no row of the warp attends costs no MUFU cycles at all
#pragma unroll
            for (int j = 0; j < kHalfN; j += 4) {
              const uint32_t w = rw[j >> 5];
              const float e0 = exp2_if(fmaf(sv[j], c, -m_use), w & (1u << (j & 31)));
              const float e1 = exp2_if(fmaf(sv[j + 1], c, -m_use), w & (1u << ((j + 1) & 31)));
              const float e2 = exp2_if(fmaf(sv[j + 2], c, -m_use), w & (1u << ((j + 2) & 31)));
              const float e3 = exp2_if(fmaf(sv[j + 3], c, -m_use), w & (1u << ((j + 3) & 31)));
              ps0 += e0; ps1 += e1; ps2 += e2; ps3 += e3;
              pk[j / 2] = pack_half2(e0, e1);
              pk[j / 2 + 1] = pack_half2(e2, e3);
            }
          }
          hsum = (ps0 + ps1) + (ps2 + ps3);
          if (tr && t == 2 && h == 0) DEFT_TRACE(112);
          // sv is dead from here (unless the tile is redone: S(t) is still in TMEM then, P has not been written over
          // it).  S of the next tile is normally there already (the S issuer runs ahead): my half of it starts its
          // way out of TMEM now, under the agreement barrier, the store of P and the hand-off.
          have_next = false;
          if (t + 1 < u.n_tiles) {
            const int nb = (gt + 1) % kSBufs;
            if (mbar_test_wait(bar(S_FULL + nb), ((gt + 1) / kSBufs) & 1)) {
              tc_fence_after();
#pragma unroll
              for (int cb = 0; cb < kHalfN / 32; ++cb) tmem_ld32_nowait(t_lane + 128 + nb * 128 + h * kHalfN + cb * 32, sv + cb * 32);
              have_next = true;
            }
          }
          // every P >= 0, so a half-row sum below 2^15 proves that no P left fp16's range; a row that had seen
          // nothing yet (m_ref = -inf) asks at its first live token.  (!(x < y) also catches NaN.)
          const bool over = !(hsum < 32768.f) || (m_ref == -INFINITY && hsum > 0.f);
          // the row pairs' two warps learn whether anybody asked (the common answer is no)
          float rq = -INFINITY;
          redo = pair_sync_or(warp, over);  // one barrier with an OR reduction; shared memory only on a request
          if (tr && t == 2 && h == 0) DEFT_TRACE(113);
          if (redo) {
            // back to this tile's S: whatever was on its way for the next tile lands first, then S(t) again
            if (have_next) tmem_wait_ld();
            have_next = false;
#pragma unroll
            for (int cb = 0; cb < kHalfN / 32; ++cb) tmem_ld32_nowait(t_s + cb * 32, sv + cb * 32);
            tmem_wait_ld();
            apply_mask();
            masked = false;   // S carries -inf now: the plain loop
            float* xr = xq + (round & 1) * (2 * kRows);
            xr[h * kRows + r] = over ? half_max() : -INFINITY;
            pair_sync(warp);
            rq = fmaxf(xr[r], xr[kRows + r]);  // the row's request, seen alike by its two threads
            ++round;
          }
          if (redo) {
            float alpha = 1.f;
            if (rq > -INFINITY) {
              alpha = fast_exp2(m_ref - rq);  // 0 when m_ref = -inf
              m_ref = rq;
              l_run *= alpha;
            }
            if (t > 0) {
              // P V of the previous tile has landed in O (this tile's waits on my P): the issuer warp follows
              // the PV_DONE phases and publishes how many tiles are complete
              {
                volatile uint32_t* pv_cnt = reinterpret_cast<volatile uint32_t*>(gbase + L::kPvCnt);
                uint32_t spins = 0;
                while (*pv_cnt < gt) {
                  __nanosleep(64);
                  if (++spins > (1u << 20)) __trap();
                }
              }
              tc_fence_after();
              float* ov = reinterpret_cast<float*>(pk);  // P is recomputed: its registers carry O meanwhile
#pragma unroll 1
              for (int cb = 0; cb < D / 64; ++cb) {
                tmem_ld32(t_o + cb * 32, ov);
#pragma unroll
                for (int j = 0; j < 32; ++j) ov[j] *= alpha;
                tmem_st32(t_o + cb * 32, ov);
              }
              tmem_wait_st();
            }
          }
          if (redo && ++n_redo > 4) __trap();  // a raise settles in one more pass: anything else is a bug, not a hang
        } while (redo);
        l_run += hsum;
        tmem_st32(t_s, reinterpret_cast<const float*>(pk));  // P_a over columns [0, 32) of S, P_b over [64, 96)
        if (tr && t == 2 && h == 0) DEFT_TRACE(114);
        if (tr && t == 2 && h == 0) DEFT_TRACE(115);
        tmem_wait_st();
        if (tr && t == 2 && h == 0) DEFT_TRACE(116);
        tc_fence_before();  // my TMEM stores (P, rescaled O) are ordered before the MMA issued after the barrier
        mbar_arrive(bar(P_FULL + 2 * buf + h));
        if (tr) DEFT_TRACE(tr0 + 4);
      }
      g0 += u.n_tiles;

      // ---- epilogue: partial = O / l as fp16, log-sum-exp in the natural-log domain
      float* xq = xchg + (g0 & 1) * (4 * kRows) + 2 * kRows;  // a slot no tile of this parity is using right now
      xq[h * kRows + r] = l_run;
      mbar_wait(bar(O_DONE), j_cnt & 1);
      ++j_cnt;
      tc_fence_after();
      pair_sync(warp);
      const float l_row = xq[r] + xq[kRows + r];
      if (ji == 0 && tid == 0) DEFT_TRACE(kTrEpiBegin);
      const bool live = qi < n_q;
      const float inv = l_row > 0.f ? 1.f / l_row : 0.f;
      const int64_t tile = (int64_t)(part_base >> 5) * p.HKV + hkv;
      uint4* dst = reinterpret_cast<uint4*>(p.po16) + tile * (CH * R) + r;  // [chunk][row] of 16 bytes
      {
        float ov[D / 2];  // my half of the O row in one round trip to TMEM
#pragma unroll
        for (int cb = 0; cb < D / 64; ++cb) tmem_ld32_nowait(t_o + cb * 32, ov + cb * 32);
        tmem_wait_ld();
        tc_fence_before();  // my reads of O are ordered before the next job's first P V (accumulate = 0)
        mbar_arrive(bar(O_EMPTY));
        if (dbg)
          for (int j = 0; j < D / 2; ++j) p.dbg[kRows * kTileN + r * D + h * (D / 2) + j] = ov[j];
        if (live) {
#pragma unroll
          for (int c8 = 0; c8 < D / 16; ++c8) {
            uint4 o4;
            o4.x = pack_half2(ov[c8 * 8 + 0] * inv, ov[c8 * 8 + 1] * inv);
            o4.y = pack_half2(ov[c8 * 8 + 2] * inv, ov[c8 * 8 + 3] * inv);
            o4.z = pack_half2(ov[c8 * 8 + 4] * inv, ov[c8 * 8 + 5] * inv);
            o4.w = pack_half2(ov[c8 * 8 + 6] * inv, ov[c8 * 8 + 7] * inv);
            dst[(h * (D / 16) + c8) * R] = o4;
          }
        }
      }
      if (live && h == 0) p.plse16[tile * R + r] = l_row > 0.f ? (m_ref + log2f(l_row)) * 0.6931471805599453f : -INFINITY;
      if (ji == 0 && tid == 0) DEFT_TRACE(kTrEpiEnd);
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
    DEFT_CUDA(cudaFuncSetAttribute(stage1_umma_v2_kernel<D, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    DEFT_CUDA(cudaFuncSetAttribute(stage1_umma_v2_kernel<D, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
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
  if (pl.dbg != nullptr) DEFT_CUDA(cudaLaunchKernelEx(&cfg, stage1_umma_v2_kernel<D, G, true>, pl));
  else DEFT_CUDA(cudaLaunchKernelEx(&cfg, stage1_umma_v2_kernel<D, G, false>, pl));
  return DEFT_OK;
}

}  // namespace

int launch_stage1_umma_v2(const AttnParams& p, cudaStream_t stream) {
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
