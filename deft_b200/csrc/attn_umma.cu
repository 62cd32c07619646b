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
#include "combine.cuh"

namespace deft {
namespace {

constexpr int kTileN = 128;  // tokens per KV tile (= the reference's BLOCK_LEN)
constexpr int kHalfN = 64;   // ... worked by the tensor pipe and the softmax warps in two halves
constexpr int kRows = 128;   // UMMA M
constexpr int kThreads = 512;  // 16 warps: 4 register-budget groups of 4 (setmaxnreg works per warpgroup)
constexpr int kMmaWarp0 = 8, kMmaWarp1 = 9, kQWarp = 10, kMaskWarp = 11, kKvWarp0 = 12;  // 12-15: K/V producers
constexpr int kSoftmaxRegs = 192, kProducerRegs = 64;  // 256 * 192 + 256 * 64 = 64 K registers
constexpr int kKvStages = 2, kMaskStages = 2;
constexpr float kRescaleLog2 = 14.f;  // raise m_ref only when a half tile tops it by more than 2^14 (P stays < 2^15 in fp16)

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a launch failure (trap) instead of a hung GPU.
// kSleepNs > 0: the waiting warp backs off between polls -- the producer / issuer warps share their SM
// sub-partition's issue slots with a softmax warp and must not spin in them.
template <int kSleepNs = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (kSleepNs > 0) __nanosleep(kSleepNs);
    if (++spins > (1u << 26)) __trap();
  }
}
// arrives on `bar` when all cp.async of this thread issued so far have landed (counts as one arrival)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// 16-byte global->shared copy; src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA: one box of a 3-D tensor map -> shared memory (swizzled by the map), completes `bytes` on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// TMA tile::gather4: four arbitrary rows of a 2-D tensor map (box {64, 1}) -> four consecutive 128-byte rows of
// shared memory (swizzled by the map), completes 4 x 128 bytes on `bar`; rows outside the tensor read as zeros
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
// NB: no fence.proxy.async on the consumer side.  Data staged by cp.async or TMA is handed over through
// an mbarrier the MMA thread waits on; a proxy fence there also waits for every async-proxy copy still
// in flight to this CTA (the NEXT tiles' loads), which serialised the tensor pipe behind the loads.
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrives on `bar` when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fast_exp2(float x) {  // MUFU.EX2; exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 32 x 32 bit-matrix transpose across a warp: lane l holds row l; on return lane l holds column l
// (bit b of the result = bit l of lane b's input).  Five butterfly stages of one shuffle each.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    const int j = 16 >> st;
    const uint32_t m = st == 0 ? 0x0000FFFFu : st == 1 ? 0x00FF00FFu : st == 2 ? 0x0F0F0F0Fu : st == 3 ? 0x33333333u : 0x55555555u;
    const uint32_t other = __shfl_xor_sync(0xffffffffu, x, j);
    x = (lane & j) ? ((x & (m << j)) | ((other >> j) & m)) : ((x & m) | ((other & m) << j));
  }
  return x;
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor)
// ------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, 128-byte swizzle; offsets in bytes
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // LayoutType::SWIZZLE_128B
  return d;
}
// instruction descriptor: fp16 x fp16 -> fp32, M = 128
__host__ __device__ constexpr uint32_t instr_desc(int n, bool b_mn_major) {
  return (1u << 4)                         // c_format = F32
         | (0u << 7) | (0u << 10)          // a_format = b_format = F16
         | (0u << 15)                      // A K-major
         | ((b_mn_major ? 1u : 0u) << 16)  // B major
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}

// One operand tile in shared memory: [D/64 or 2 panels][128 rows][128 bytes], 16-byte chunks XOR-swizzled
// by (row & 7) -- the canonical SWIZZLE_128B layout.  For K-major operands (Q, K) a row is an M/N
// index and a panel is 64 elements of the contraction dim; for the MN-major operand (V) a row is a
// token (contraction index) and a panel is 64 elements of D.
constexpr int kPanelBytes = kRows * 128;
__device__ __forceinline__ uint32_t tile_off(int row, int chunk16) {
  return (uint32_t)((chunk16 >> 3) * kPanelBytes + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
}

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
  K_FULL = 0, K_EMPTY = K_FULL + kKvStages, V_FULL = K_EMPTY + kKvStages, V_EMPTY = V_FULL + kKvStages,
  Q_FULL = V_EMPTY + kKvStages, Q_EMPTY = Q_FULL + 2,
  M_FULL = Q_EMPTY + 2,                    // [slot][stage]
  M_EMPTY = M_FULL + 2 * kMaskStages,
  S_FULL = M_EMPTY + 2 * kMaskStages,      // [slot]: S of one tile is in TMEM
  P_FULL = S_FULL + 2,                     // [slot][half]: P of one 64-token half has been written over S
  O_FULL = P_FULL + 4,                     // [slot]: one phase per tile (P V of the whole tile has landed in O)
  O_DONE = O_FULL + 2,                     // [slot]: one phase per job (the last P V has landed: O is complete)
  O_EMPTY = O_DONE + 2,                    // [slot]
  PVA_DONE = O_EMPTY + 2,                  // [slot]: one phase per tile (P V of the tile's first half has landed in O)
  ORDER = PVA_DONE + 2,                    // [slot]: the slot's turn on the exp (MUFU) section
  kNumBars = ORDER + 2
};

template <int D>
struct Layout {
  static constexpr int kOperandBytes = kRows * D * 2;                      // Q, K or V tile
  static constexpr int kK = 0;                                             // [stage]
  static constexpr int kV = kK + kKvStages * kOperandBytes;                // [stage]
  static constexpr int kQ = kV + kKvStages * kOperandBytes;                // [slot]
  static constexpr int kMask = kQ + 2 * kOperandBytes;                     // [slot][stage][128] u32
  static constexpr int kFlag = kMask + 2 * kMaskStages * kTileN * 4;       // [slot][stage] u32
  static constexpr int kBars = kFlag + 2 * kMaskStages * 4;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kBytes = kTmemSlot + 16;
  static constexpr int kAlloc = kBytes + 1024;  // slack for the manual 1024-byte alignment
};

// The (unit, kv-head) jobs of one CTA: an explicit host-balanced list, or jobs c, c+grid, ...
struct Jobs {
  const int32_t* list;
  int begin, end, stride;
  __device__ __forceinline__ Jobs(const AttnParams& p) {
    if (p.job_off != nullptr) {
      list = p.jobs;
      begin = (int)blockIdx.x < p.n_ctas ? p.job_off[blockIdx.x] : 0;
      end = (int)blockIdx.x < p.n_ctas ? p.job_off[blockIdx.x + 1] : 0;
      stride = 1;
    } else {
      list = nullptr;
      const int n_units = p.n_units_dev ? *p.n_units_dev : p.n_units;
      begin = blockIdx.x;
      end = n_units * p.HKV;
      stride = gridDim.x;
    }
  }
  __device__ __forceinline__ int get(int i) const { return list ? list[i] : i; }
};

template <int D, int G>
__global__ void __launch_bounds__(kThreads, 1) stage1_umma_kernel(const __grid_constant__ AttnParams p) {
  using L = Layout<D>;
  constexpr int CH = D / 8;           // 16-byte chunks per row
  constexpr int R = kMaxGroupQ * G;   // live rows of a full slot
  constexpr uint32_t kTmemCols = 512; // S_0 [0,128) S_1 [128,256) O_0 [256,256+D) O_1 [384,384+D)
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
  if (tid == 0) {
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(bar(K_FULL + s), 128); mbar_init(bar(K_EMPTY + s), 2);  // one commit per slot issuer
      mbar_init(bar(V_FULL + s), 128); mbar_init(bar(V_EMPTY + s), 2);  // (slot 0's commits twice in a one-slot job)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(Q_FULL + s), 32); mbar_init(bar(Q_EMPTY + s), 1);
      for (int m = 0; m < kMaskStages; ++m) {
        mbar_init(bar(M_FULL + s * kMaskStages + m), 32);
        mbar_init(bar(M_EMPTY + s * kMaskStages + m), 128);
      }
      mbar_init(bar(S_FULL + s), 1);
      for (int h = 0; h < 2; ++h) mbar_init(bar(P_FULL + 2 * s + h), 128);
      mbar_init(bar(O_FULL + s), 1); mbar_init(bar(O_DONE + s), 1); mbar_init(bar(O_EMPTY + s), 128);
      mbar_init(bar(PVA_DONE + s), 1); mbar_init(bar(ORDER + s), 128);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp0) tmem_alloc(base + L::kTmemSlot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gbase + L::kTmemSlot);

  const Jobs jobs(p);
  // Programmatic dependent launch: everything up to here (barrier init, TMEM allocation, job list) overlapped
  // the tail of the preceding kernel; q, the KV pool and the partial workspace may still be in its hands.
  griddep_wait();
  if (tid == 0) DEFT_TRACE(kTrStart);
  if (warp >= 8) {
  reg_dealloc<kProducerRegs>();  // warps 8-11 and 12-15: two whole warpgroups give registers away
  if (warp >= kKvWarp0) {
    // ============================== K / V producers: warp w owns token rows [32w, 32w+32) ==============================
    const int w = warp - kKvWarp0;
    uint32_t cnt = 0;  // tiles produced
    for (int ji = jobs.begin; ji < jobs.end; ji += jobs.stride) {
      const int job = jobs.get(ji);
      const int hkv = job % p.HKV;
      const deft_unit_t u = p.units[job / p.HKV];
      if (w == 0 && lane == 0 && ji == jobs.begin) DEFT_TRACE(kTrKUnit);
      const int n_mine = w * 32 + lane;  // my token row
      const bool known_run = u.page0 >= 0 && p.tma_kv != 0;  // the builder's shortcut: no index-table read at all
      auto page_of = [&](int t) -> int64_t {
        const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
        if (known_run) return (int64_t)u.page0 + t * kTileN + n_mine;
        return t < u.n_tiles && n_mine < tlen ? load_index(p.u_kv, p.u_kv_bytes, u.kv_off + (int64_t)t * u.kv_tile_stride + n_mine) : 0;
      };
      int64_t pg_next = page_of(0);
      for (int t = 0; t < u.n_tiles; ++t, ++cnt) {
        const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
        const int st = cnt % kKvStages;
        const uint32_t ph = ((cnt / kKvStages) & 1) ^ 1;
        const int64_t pg = pg_next;
        pg_next = page_of(t + 1);  // the next tile's page id is in flight while this tile is issued
        // 32 consecutive pages of a full tile are ONE box of the pool's tensor map per 64-wide panel
        const int64_t page0 = __shfl_sync(0xffffffffu, pg, 0);
        const bool run = known_run || __all_sync(0xffffffffu, p.tma_kv != 0 && tlen == kTileN && pg == page0 + lane);
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
          const int full = (kv == 0 ? K_FULL : V_FULL) + st;
          mbar_wait<64>(bar((kv == 0 ? K_EMPTY : V_EMPTY) + st), ph);
          const uint32_t dst_base = base + (kv == 0 ? L::kK : L::kV) + st * L::kOperandBytes;
          if (run) {
            if (lane == 0) {
              mbar_arrive_expect_tx(bar(full), 32 * D * 2);
#pragma unroll
              for (int pn = 0; pn < D / 64; ++pn)
                tma_load_3d(dst_base + pn * kPanelBytes + w * 32 * 128, kv == 0 ? &p.tmap_k : &p.tmap_v, bar(full), pn * 64,
                            hkv, (int)page0);
            } else {
              mbar_arrive(bar(full));
            }
          } else if (p.tma_gather != 0) {
            // scattered pages: lane (g, panel) moves the four rows 4g .. 4g+3 of my 32 with one gather4 per panel;
            // rows past the tile's length name a row outside the map and arrive as zeros
            constexpr int NP = D / 64;
            const int g = lane / NP, pn = lane % NP;
            const int my_row = n_mine < tlen ? (int)pg * p.kv_row_ratio + hkv : p.kv_rows;
            const int r0 = __shfl_sync(0xffffffffu, my_row, (4 * g) & 31), r1 = __shfl_sync(0xffffffffu, my_row, (4 * g + 1) & 31);
            const int r2 = __shfl_sync(0xffffffffu, my_row, (4 * g + 2) & 31), r3 = __shfl_sync(0xffffffffu, my_row, (4 * g + 3) & 31);
            if (lane == 0) mbar_arrive_expect_tx(bar(full), 32 * D * 2);
            else mbar_arrive(bar(full));
            if (lane < 8 * NP)
              tma_gather4(dst_base + pn * kPanelBytes + (w * 32 + 4 * g) * 128, kv == 0 ? &p.tmap_kg : &p.tmap_vg, bar(full), pn * 64,
                          r0, r1, r2, r3);
          } else {
            const __half* src_base = (kv == 0 ? p.k : p.v) + (int64_t)hkv * p.kv_head_stride;
            constexpr int TOK_PER_INSTR = 32 / CH;  // tokens covered by one warp-wide copy
#pragma unroll 4
            for (int i = 0; i < 32 / TOK_PER_INSTR; ++i) {
              const int nl = i * TOK_PER_INSTR + lane / CH;  // row inside my 32
              const int ch = lane % CH;
              const int64_t page = __shfl_sync(0xffffffffu, pg, nl);
              const bool ok = w * 32 + nl < tlen;
              cp_async_16(dst_base + tile_off(w * 32 + nl, ch), src_base + page * p.kv_tok_stride + ch * 8, ok ? 16u : 0u);
            }
            cp_async_arrive(bar(full));
          }
          if (w == 0 && lane == 0 && ji == jobs.begin) DEFT_TRACE(kTrTile0 + 8 * t + (kv == 0 ? 0 : 6));
        }
      }
    }
  } else if (warp == kQWarp) {
    // ============================== Q tiles of both slots ==============================
    uint32_t q_cnts[2] = {0, 0};  // jobs per slot
    for (int ji = jobs.begin; ji < jobs.end; ji += jobs.stride) {
      const int job = jobs.get(ji);
      const int hkv = job % p.HKV;
      const deft_unit_t u = p.units[job / p.HKV];
#pragma unroll
     for (int s = 0; s < 2; ++s) {
      if (u.q_cnt[s] == 0) continue;
      uint32_t& q_cnt = q_cnts[s];
      // row r = (query r / G, head r % G); rows past q_cnt*G are zero
      const bool known_run = u.q_id0[s] >= 0 && p.tma_q != 0;  // the builder's shortcut: no query-table read
      const int64_t my_q = known_run ? (int64_t)u.q_id0[s] + lane
                                     : (lane < u.q_cnt[s] ? load_index(p.u_q, p.u_q_bytes, u.q_off[s] + lane) : 0);
      mbar_wait<64>(bar(Q_EMPTY + s), (q_cnt & 1) ^ 1);
      const uint32_t qs = base + L::kQ + s * L::kOperandBytes;
      if (lane == 0 && ji == jobs.begin && s == 0 && my_q >= 0) DEFT_TRACE(kTrQIds);
      // consecutive query ids: the slot's G heads x 32 queries are ONE box of q's tensor map per panel
      // (rows past q_cnt then hold the next queries or zeros: finite, never stored)
      const int64_t q0 = __shfl_sync(0xffffffffu, my_q, 0);
      const bool run = p.tma_q != 0 && (lane >= u.q_cnt[s] || my_q == q0 + lane);
      if (known_run || __all_sync(0xffffffffu, run)) {
        if (lane == 0) {
          mbar_arrive_expect_tx(bar(Q_FULL + s), R * D * 2);
#pragma unroll
          for (int pn = 0; pn < D / 64; ++pn)
            tma_load_3d(qs + pn * kPanelBytes, &p.tmap_q, bar(Q_FULL + s), pn * 64, hkv * G, (int)q0);
        } else {
          mbar_arrive(bar(Q_FULL + s));
        }
      } else {
#pragma unroll 4
        for (int i = 0; i < kRows * CH / 32; ++i) {
          const int c = lane + i * 32;
          const int r = c / CH, ch = c % CH;
          const int qi = r / G, g = r % G;
          const int64_t qid = __shfl_sync(0xffffffffu, my_q, qi & 31);
          const bool ok = qi < u.q_cnt[s];
          const __half* src = p.q + qid * p.q_row_stride + (int64_t)(hkv * G + g) * p.q_head_stride + ch * 8;
          cp_async_16(qs + tile_off(r, ch), ok ? src : p.q, ok ? 16u : 0u);
        }
        cp_async_arrive(bar(Q_FULL + s));
      }
      if (lane == 0 && ji == jobs.begin) DEFT_TRACE(kTrQ0Issued + s);
      ++q_cnt;
     }
    }
  } else if (warp == kMaskWarp) {
    // ============================== mask words + dense flag per (tile, slot) ==============================
    uint32_t m_cnt[2] = {0, 0};  // tiles per slot
    for (int ji = jobs.begin; ji < jobs.end; ji += jobs.stride) {
      const int job = jobs.get(ji);
      const deft_unit_t u = p.units[job / p.HKV];
      const int n_slots = u.q_cnt[1] > 0 ? 2 : 1;
      for (int t = 0; t < u.n_tiles; ++t) {
        const int tlen = t == u.n_tiles - 1 ? u.last_len : kTileN;
        for (int s = 0; s < n_slots; ++s) {
          const int st = m_cnt[s] % kMaskStages;
          mbar_wait<64>(bar(M_EMPTY + s * kMaskStages + st), ((m_cnt[s] / kMaskStages) & 1) ^ 1);
          uint32_t* ms = reinterpret_cast<uint32_t*>(gbase + L::kMask) + (s * kMaskStages + st) * kTileN;
          const uint32_t full = u.q_cnt[s] >= 32 ? 0xffffffffu : ((1u << u.q_cnt[s]) - 1u);
          // per-token words: bit r = row r of the slot attends token lane + 32j
          uint32_t m[kTileN / 32];
          bool dense = tlen == kTileN;
#pragma unroll
          for (int j = 0; j < kTileN / 32; ++j) {
            const int n = lane + 32 * j;
            m[j] = 0;
            if (n < tlen)
              m[j] = u.mask_off[s] >= 0
                         ? (uint32_t)load_index(p.u_mask, p.u_mask_bytes, u.mask_off[s] + (int64_t)t * u.mask_tile_stride + n)
                         : 0xffffffffu;
            dense = dense && ((m[j] & full) == full);
          }
          dense = __all_sync(0xffffffffu, dense);
          if (!dense) {
            // transpose to row masks: lane = query, word j bit n = the query attends token 32j + n
            *reinterpret_cast<uint4*>(ms + lane * 4) = make_uint4(warp_transpose32(m[0], lane), warp_transpose32(m[1], lane),
                                                                  warp_transpose32(m[2], lane), warp_transpose32(m[3], lane));
          }
          if (lane == 0) reinterpret_cast<uint32_t*>(gbase + L::kFlag)[s * kMaskStages + st] = dense ? 1u : 0u;
          mbar_arrive(bar(M_FULL + s * kMaskStages + st));
          if (lane == 0 && ji == jobs.begin && t == 0 && s == 0) DEFT_TRACE(kTrMask0);
          ++m_cnt[s];
        }
      }
    }
  } else if (warp == kMmaWarp0 || warp == kMmaWarp1) {
    // ============================== MMA issuer of one slot ==============================
    // The whole warp runs the (uniform) control flow and the waits; lane 0 alone executes the tcgen05.mma /
    // tcgen05.commit instructions.  Per tile:  S_s(t) = Q_s K(t)^T (N = 128: SS MMAs are bound by the 128 B/clk
    // of shared memory, so narrower ones cost more per token), then O_s (+)= P_s(t) V(t) in two 64-token halves,
    // each as soon as the softmax warps have written that half of P over S (the second half's exponentials
    // overlap the first half's MMAs), then straight on to S_s(t+1) in the same in-order stream.
    const int s = warp == kMmaWarp0 ? 0 : 1;
    const bool leader = lane == 0;
    const uint32_t s_tmem = tmem + s * 128, o_tmem = tmem + 256 + s * 128;
    const uint64_t q_desc = smem_desc_sw128(base + L::kQ + s * L::kOperandBytes, 16, 1024);
    uint32_t kv_cnt = 0;  // KV tiles consumed (ring position; every job advances both slots' issuers alike)
    uint32_t s_cnt = 0;   // tiles of this slot (S_FULL / P_FULL / O_FULL phases)
    uint32_t j_cnt = 0;   // jobs of this slot (Q_FULL / O_EMPTY phases)
    for (int ji = jobs.begin; ji < jobs.end; ji += jobs.stride) {
      const int job = jobs.get(ji);
      const deft_unit_t u = p.units[job / p.HKV];
      const int n = u.n_tiles;
      const bool has_b = u.q_cnt[1] > 0;
      if (s == 1 && !has_b) {  // one-slot job: slot 0's issuer signs off the KV stages for both
        kv_cnt += n;
        continue;
      }
      const bool tr0 = ji == jobs.begin && leader && s == 0;
      // S_s(t): descriptors advance by compile-time constants (the start-address field holds bytes >> 4 and
      // cannot carry out of its 14 bits inside 227 KB of shared memory).
      auto issue_s = [&](int t) {
        const uint32_t c = kv_cnt + t;
        const int st = c % kKvStages;
        mbar_wait<20>(bar(K_FULL + st), (c / kKvStages) & 1);
        if (tr0) DEFT_TRACE(kTrTile0 + 8 * t + 1);
        if (t == 0) mbar_wait<20>(bar(Q_FULL + s), j_cnt & 1);
        if (tr0 && t == 0) DEFT_TRACE(kTrMmaQFull);
        tc_fence_after();
        const uint64_t k_desc = smem_desc_sw128(base + L::kK + st * L::kOperandBytes, 16, 1024);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint64_t koff = (uint64_t)(((ks >> 2) * kPanelBytes + (ks & 3) * 32) >> 4);
            umma_ss(s_tmem, q_desc + koff, k_desc + koff, kIdescQK, ks > 0);
          }
          umma_commit(bar(S_FULL + s));
          umma_commit(bar(K_EMPTY + st));  // K(t) and, after the last tile, Q_s are free
          if (!has_b) umma_commit(bar(K_EMPTY + st));
          if (t == n - 1) umma_commit(bar(Q_EMPTY + s));
        }
      };
      issue_s(0);
      __syncwarp();
      for (int t = 0; t < n; ++t) {
        const uint32_t c = kv_cnt + t;
        const int st = c % kKvStages;
        mbar_wait<20>(bar(V_FULL + st), (c / kKvStages) & 1);
        const uint64_t v_desc = smem_desc_sw128(base + L::kV + st * L::kOperandBytes, kPanelBytes, 1024);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // ---- O_s (+)= P_s(t)[:, 64h .. 64h+64) V(t)[64h .. 64h+64)
          mbar_wait<20>(bar(P_FULL + 2 * s + h), (s_cnt + t) & 1);
          if (tr0 && h == 0) DEFT_TRACE(kTrTile0 + 8 * t + 5);
          if (t == 0 && h == 0) mbar_wait<20>(bar(O_EMPTY + s), (j_cnt & 1) ^ 1);
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < kHalfN / 16; ++ks)
              umma_ts(o_tmem, s_tmem + h * (kHalfN / 2) + ks * 8, v_desc + (uint64_t)(((h * kHalfN + ks * 16) * 128) >> 4), kIdescPV,
                      t > 0 || h > 0 || ks > 0);
            if (h == 0) umma_commit(bar(PVA_DONE + s));
            if (tr0 && t == 1) DEFT_TRACE(14 + h);
          }
        }
        if (t + 1 < n) issue_s(t + 1);
        if (leader) {  // the commits (each covers every MMA issued before it) come after S(t+1), off the critical path
          umma_commit(bar(O_FULL + s));
          umma_commit(bar(V_EMPTY + st));
          if (!has_b) umma_commit(bar(V_EMPTY + st));
          if (t == n - 1) umma_commit(bar(O_DONE + s));
        }
        __syncwarp();
      }
      kv_cnt += n;
      s_cnt += n;
      ++j_cnt;
    }
  }
  } else {
    reg_alloc<kSoftmaxRegs>();   // warps 0-3 and 4-7
    // ============================== softmax + epilogue (slot = warp / 4) ==============================
    const int s = warp >> 2;
    const int r = tid & 127;  // my row == my TMEM lane
    const int qi = r / G;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + s * 128, t_o = t_lane + 256 + s * 128;
    const float c = p.scale * 1.4426950408889634f;  // scores are handled in the log2 domain
    uint32_t s_cnt = 0, m_cnt = 0, j_cnt = 0;
    uint32_t ord_cnt = 0;  // tiles of two-slot jobs: the slots take turns on the exp section
    bool first_job = blockIdx.x == 0;

    for (int ji = jobs.begin; ji < jobs.end; ji += jobs.stride) {
      const int job = jobs.get(ji);
      const int hkv = job % p.HKV;
      const deft_unit_t u = p.units[job / p.HKV];
      if (s == 1 && u.q_cnt[1] == 0) continue;
      const bool dbg = p.dbg != nullptr && first_job && s == 0;
      first_job = false;
      float m_ref = -INFINITY, l_run = 0.f;

      for (int t = 0; t < u.n_tiles; ++t, ++s_cnt, ++m_cnt) {
        const int mst = m_cnt % kMaskStages;
        mbar_wait<32>(bar(M_FULL + s * kMaskStages + mst), (m_cnt / kMaskStages) & 1);
        const uint32_t* ms = reinterpret_cast<const uint32_t*>(gbase + L::kMask) + (s * kMaskStages + mst) * kTileN;
        const bool dense = reinterpret_cast<const volatile uint32_t*>(gbase + L::kFlag)[s * kMaskStages + mst] != 0;
        const bool tr = ji == jobs.begin && (tid & 127) == 0 && t < 5;
        const int tr0 = kTrTile0 + (s == 0 ? 0 : 48) + 8 * t;  // slot 1 events sit 48 slots higher
        const bool trf = tr && s == 0 && t == 1;
        mbar_wait<32>(bar(S_FULL + s), s_cnt & 1);
        tc_fence_after();
        if (tr) DEFT_TRACE(tr0 + 2);

        // In a two-slot job the slots alternate on the exp section (slot 0 first), which keeps them half a
        // period apart: one exponentiates while the tensor pipe serves the other.
        const bool ordered = u.q_cnt[1] > 0;

        // ---- P = exp2(S*c - m_ref) -> packed fp16, written over S in TMEM (columns [0, 64) of the slot), worked
        // and handed to the tensor pipe in two 64-token halves: the first half's P V runs under the second
        // half's exponentials.
        // Reference maximum m_ref: the first half tile of a job sets it to its exact row maximum.  Every later
        // half is exponentiated against the m_ref it finds (no dependent max -> exp chain): its own maximum is
        // tracked inside the MUFU-bound loop, and only if it tops m_ref by more than 2^14 (P would leave fp16)
        // is m_ref raised, the accumulator rescaled and the half redone.
        float sv[2][kHalfN];  // my row of S, one 64-column half at a time: out of TMEM once, kept in registers
#pragma unroll
        for (int cb = 0; cb < kHalfN / 32; ++cb) tmem_ld32_nowait(t_s + cb * 32, sv[0] + cb * 32);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tmem_wait_ld();
          if (trf && h == 0) DEFT_TRACE(10);
          if (dbg && t == 0)
            for (int j = 0; j < kHalfN; ++j) p.dbg[r * kTileN + h * kHalfN + j] = sv[h][j];
          if (!dense) {  // masked-out tokens score -inf: my query's token bitmask comes from the mask warp
            const uint2 rm = qi < 32 ? *reinterpret_cast<const uint2*>(ms + qi * 4 + h * 2) : make_uint2(0u, 0u);
            const uint32_t rw[2] = {rm.x, rm.y};
#pragma unroll
            for (int j = 0; j < kHalfN; ++j)
              if (!((rw[j >> 5] >> (j & 31)) & 1u)) sv[h][j] = -INFINITY;
          }
          if (h == 1) mbar_arrive(bar(M_EMPTY + s * kMaskStages + mst));
          auto half_max = [&]() {
            float m0 = sv[h][0], m1 = sv[h][1], m2 = sv[h][2], m3 = sv[h][3];
#pragma unroll
            for (int j = 4; j < kHalfN; j += 4) {
              m0 = fmaxf(m0, sv[h][j]); m1 = fmaxf(m1, sv[h][j + 1]); m2 = fmaxf(m2, sv[h][j + 2]); m3 = fmaxf(m3, sv[h][j + 3]);
            }
            return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;  // c > 0; -inf when the row attends nothing here
          };
          if (t == 0 && h == 0) m_ref = half_max();
          if (h == 0) {
            if (tr) DEFT_TRACE(tr0 + 3);
            if (ordered) mbar_wait<32>(bar(ORDER + s), s == 0 ? ((ord_cnt & 1) ^ 1) : (ord_cnt & 1));
            if (trf) DEFT_TRACE(11);
          }
          uint32_t pk[kHalfN / 2];
          float hsum;
          bool redo;
          do {
            const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
            float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
#pragma unroll
            for (int j = 0; j < kHalfN; j += 4) {
              const float e0 = fast_exp2(fmaf(sv[h][j], c, -m_use)), e1 = fast_exp2(fmaf(sv[h][j + 1], c, -m_use));
              const float e2 = fast_exp2(fmaf(sv[h][j + 2], c, -m_use)), e3 = fast_exp2(fmaf(sv[h][j + 3], c, -m_use));
              ps0 += e0; ps1 += e1; ps2 += e2; ps3 += e3;
              pk[j / 2] = pack_half2(e0, e1);
              pk[j / 2 + 1] = pack_half2(e2, e3);
            }
            hsum = (ps0 + ps1) + (ps2 + ps3);
            // every P >= 0, so a half-row sum below 2^15 proves that no P left fp16's range; a row that had seen
            // nothing yet (m_ref = -inf) takes the slow path at its first live token.  (!(x < y) also catches NaN.)
            const bool over = !(hsum < 32768.f) || (m_ref == -INFINITY && hsum > 0.f);
            redo = __any_sync(0xffffffffu, over);
            if (redo) {
              float alpha = 1.f;
              if (over) {
                const float hmax = half_max();
                alpha = fast_exp2(m_ref - hmax);  // 0 when m_ref = -inf
                m_ref = hmax;
                l_run *= alpha;
              }
              if (t > 0 || h > 0) {
                // every P V issued so far has landed in O: the previous tile's, or this tile's first half
                if (h == 0) mbar_wait(bar(O_FULL + s), (s_cnt - 1) & 1);
                else mbar_wait(bar(PVA_DONE + s), s_cnt & 1);
                tc_fence_after();
                float* ov = reinterpret_cast<float*>(pk);  // P is recomputed: its registers carry O meanwhile
#pragma unroll 1
                for (int cb = 0; cb < D / 32; ++cb) {
                  tmem_ld32(t_o + cb * 32, ov);
#pragma unroll
                  for (int j = 0; j < 32; ++j) ov[j] *= alpha;
                  tmem_st32(t_o + cb * 32, ov);
                }
                tmem_wait_st();
              }
            }
          } while (redo);
          l_run += hsum;
          if (h == 0) {  // the second half of S is on its way out of TMEM while the first half of P goes in
#pragma unroll
            for (int cb = 0; cb < kHalfN / 32; ++cb) tmem_ld32_nowait(t_s + kHalfN + cb * 32, sv[1] + cb * 32);
          } else if (ordered) {
            mbar_arrive(bar(ORDER + (s ^ 1)));  // the MUFU-bound part of my turn is over
            ++ord_cnt;
          }
          if (trf && h == 0) DEFT_TRACE(12);
          tmem_st32(t_s + h * (kHalfN / 2), reinterpret_cast<const float*>(pk));
          tmem_wait_st();
          if (trf && h == 0) DEFT_TRACE(13);
          tc_fence_before();  // my TMEM stores (P, rescaled O) are ordered before the MMA issued after the barrier
          mbar_arrive(bar(P_FULL + 2 * s + h));
          if (tr) DEFT_TRACE(tr0 + (h == 0 ? 4 : 7));
        }
      }

      // ---- epilogue: partial = O / l as fp16, log-sum-exp in the natural-log domain
      // (not O_FULL: a thread that is two of its phases behind would read the parity of an older phase)
      mbar_wait(bar(O_DONE + s), j_cnt & 1);
      ++j_cnt;
      tc_fence_after();
      if (ji == jobs.begin && tid == 0) DEFT_TRACE(kTrEpiBegin);
      const bool live = qi < u.q_cnt[s];
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const int64_t tile = (int64_t)(u.part_base[s] >> 5) * p.HKV + hkv;
      uint4* dst = reinterpret_cast<uint4*>(p.po16) + tile * (CH * R) + r;  // [chunk][row] of 16 bytes
      {
        float ov[D];  // the whole O row in one round trip to TMEM, then O(1) arrives as soon as it is in registers
#pragma unroll
        for (int cb = 0; cb < D / 32; ++cb) tmem_ld32_nowait(t_o + cb * 32, ov + cb * 32);
        tmem_wait_ld();
        tc_fence_before();  // my reads of O are ordered before the next job's first P V (accumulate = 0)
        mbar_arrive(bar(O_EMPTY + s));
        if (dbg)
          for (int j = 0; j < D; ++j) p.dbg[kRows * kTileN + r * D + j] = ov[j];
        if (live) {
#pragma unroll
          for (int c8 = 0; c8 < D / 8; ++c8) {
            uint4 pk;
            pk.x = pack_half2(ov[c8 * 8 + 0] * inv, ov[c8 * 8 + 1] * inv);
            pk.y = pack_half2(ov[c8 * 8 + 2] * inv, ov[c8 * 8 + 3] * inv);
            pk.z = pack_half2(ov[c8 * 8 + 4] * inv, ov[c8 * 8 + 5] * inv);
            pk.w = pack_half2(ov[c8 * 8 + 6] * inv, ov[c8 * 8 + 7] * inv);
            dst[c8 * R] = pk;
          }
        }
      }
      if (live) p.plse16[tile * R + r] = l_run > 0.f ? (m_ref + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
      if (ji == jobs.begin && tid == 0) DEFT_TRACE(kTrEpiEnd);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) DEFT_TRACE(kTrEnd);
  if (warp == kMmaWarp0) tmem_dealloc(tmem, kTmemCols);
}

template <int D, int G>
int launch_t(const AttnParams& p, cudaStream_t stream) {
  static bool configured = false;
  static int num_sms = 0;
  using L = Layout<D>;
  if (!configured) {
    DEFT_CUDA(cudaFuncSetAttribute(stage1_umma_kernel<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    int dev = 0;
    DEFT_CUDA(cudaGetDevice(&dev));
    DEFT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  int grid;
  if (p.job_off != nullptr) {
    grid = p.n_ctas;
  } else {
    const int64_t n_jobs = (int64_t)p.n_units * p.HKV;
    grid = (int)(n_jobs < num_sms ? n_jobs : num_sms);
  }
  if (grid <= 0) return DEFT_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L::kAlloc;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // launch early, wait inside (griddep_wait)
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  DEFT_CUDA(cudaLaunchKernelEx(&cfg, stage1_umma_kernel<D, G>, p));
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
