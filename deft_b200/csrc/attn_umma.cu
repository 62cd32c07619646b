// Stage 1, tensor-core path (sm_100a): tcgen05.mma with TMEM accumulators, persistent CTAs.
//
// One work unit = (item, kv-head).  A unit's KV tile (128 tokens x D, K and V) is gathered from the
// token-granular paged pool ONCE into 128B-swizzled shared memory and serves every query group of
// the item and all G = H/HKV query heads that share the kv-head: a group of <= 32 queries x G heads
// is one M = 128 UMMA tile (row r = query r/G, head r%G).
//
//   S[128 x 128]  = Q[128 x D] . K^T          tcgen05.mma kind::f16, A/B K-major SW128 smem, D in TMEM
//   P             = exp2(S*c - m), masked by the per-token bitmask     (4 softmax warps, row = TMEM lane)
//   O[128 x D]   += P[128 x 128] . V          A = P (K-major smem), B = V (MN-major SW128 smem)
//
// Warp roles (192 threads): warps 0-3 softmax/epilogue (thread 0 also issues the MMAs), warps 4-5
// producers (cp.async 16-byte gathers of paged KV rows, Q rows, masks -> mbarrier rings).
// Reference semantics: DeFT/deft/layers/attention/tree_attention.py:860-976 (Flatten stage 1) and
// :170-293 (Node stage 1); items longer than 128 tokens are walked tile by tile with online softmax.
#include "common.cuh"

namespace deft {
namespace {

constexpr int kTileN = 128;           // tokens per KV tile (= the reference's BLOCK_LEN)
constexpr int kRows = 128;            // UMMA M
constexpr int kComputeThreads = 128;  // warps 0-3
constexpr int kProducerThreads = 64;  // warps 4-5
constexpr int kThreads = kComputeThreads + kProducerThreads;
constexpr int kKvStages = 2, kQStages = 2, kMaskStages = 2;

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
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
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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
// by (row & 7) -- the canonical SWIZZLE_128B layout.  For K-major operands (Q, K, P) a row is an M/N
// index and a panel is 64 elements of the contraction dim; for the MN-major operand (V) a row is a
// token (contraction index) and a panel is 64 elements of D.
constexpr int kPanelBytes = kRows * 128;
__device__ __forceinline__ uint32_t tile_off(int row, int chunk16) {
  return (uint32_t)((chunk16 >> 3) * kPanelBytes + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
}

template <int D>
struct Layout {
  static constexpr int kOperandBytes = kRows * D * 2;  // Q, K or V tile
  static constexpr int kPBytes = kRows * kTileN * 2;
  static constexpr int kKv = 0;                                               // [stage][K|V]
  static constexpr int kQ = kKv + kKvStages * 2 * kOperandBytes;              // [stage]
  static constexpr int kP = kQ + kQStages * kOperandBytes;
  static constexpr int kMask = kP + kPBytes;                                  // [stage][128] u32
  static constexpr int kBars = kMask + kMaskStages * kTileN * 4;
  static constexpr int kNumBars = 2 * kKvStages + 2 * kQStages + 2 * kMaskStages + 2;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kBytes = kTmemSlot + 16;
  static constexpr int kAlloc = kBytes + 1024;  // slack for the manual 1024-byte alignment
};

struct Ring {  // position in an mbarrier ring
  int stage = 0;
  uint32_t phase = 0;
  template <int N>
  __device__ __forceinline__ void next() {
    if (++stage == N) {
      stage = 0;
      phase ^= 1;
    }
  }
};

template <int D, int G>
__global__ void __launch_bounds__(kThreads, 1) stage1_umma_kernel(const AttnParams p) {
  using L = Layout<D>;
  constexpr int CH = D / 8;            // 16-byte chunks per row
  constexpr int kTmemCols = 256;       // S: columns [0,128), O: columns [128,128+D)
  constexpr uint32_t kIdescQK = instr_desc(kTileN, false);
  constexpr uint32_t kIdescPV = instr_desc(D, true);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + L::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  // barrier indices
  constexpr int KV_FULL = 0, KV_EMPTY = KV_FULL + kKvStages, Q_FULL = KV_EMPTY + kKvStages,
                Q_EMPTY = Q_FULL + kQStages, M_FULL = Q_EMPTY + kQStages, M_EMPTY = M_FULL + kMaskStages,
                S_FULL = M_EMPTY + kMaskStages, O_FULL = S_FULL + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kKvStages; ++s) { mbar_init(bar(KV_FULL + s), kProducerThreads); mbar_init(bar(KV_EMPTY + s), 1); }
    for (int s = 0; s < kQStages; ++s) { mbar_init(bar(Q_FULL + s), kProducerThreads); mbar_init(bar(Q_EMPTY + s), 1); }
    for (int s = 0; s < kMaskStages; ++s) { mbar_init(bar(M_FULL + s), kProducerThreads); mbar_init(bar(M_EMPTY + s), kComputeThreads); }
    mbar_init(bar(S_FULL), 1);
    mbar_init(bar(O_FULL), 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(base + L::kTmemSlot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gbase + L::kTmemSlot);

  const int n_items = p.n_items_dev ? *p.n_items_dev : p.n_items;
  const int n_units = n_items * p.HKV;

  if (warp >= 4) {
    // ============================== producers ==============================
    const int pt = tid - kComputeThreads;  // 0..63
    const int pw = warp - 4;               // 0: K, 1: V
    Ring kv, qr, mr;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int item_id = u / p.HKV, hkv = u % p.HKV;
      const deft_item_t item = p.items[item_id];
      const int ntile = max(1, (item.kv_len + kTileN - 1) / kTileN);
      for (int gi = 0; gi < item.n_grp; ++gi) {
        const deft_group_t grp = p.groups[item.grp_off + gi];
        // ---- Q tile of the group: row r = (query r / G, head r % G); rows past q_cnt*G are zero
        {
          mbar_wait(bar(Q_EMPTY + qr.stage), qr.phase ^ 1);
          const int64_t my_q = lane < grp.q_cnt ? p.q_list[grp.q_off + lane] : 0;
          const uint32_t qs = base + L::kQ + qr.stage * L::kOperandBytes;
#pragma unroll 4
          for (int i = 0; i < kRows * CH / kProducerThreads; ++i) {
            const int c = pt + i * kProducerThreads;
            const int r = c / CH, ch = c % CH;
            const int qi = r / G, g = r % G;
            const int64_t qid = __shfl_sync(0xffffffffu, my_q, qi & 31);
            const bool ok = qi < grp.q_cnt;
            const __half* src = p.q + qid * p.q_row_stride + (int64_t)(hkv * G + g) * p.q_head_stride + ch * 8;
            cp_async_16(qs + tile_off(r, ch), ok ? src : p.q, ok ? 16u : 0u);
          }
          cp_async_arrive(bar(Q_FULL + qr.stage));
          qr.next<kQStages>();
        }
        for (int t = 0; t < ntile; ++t) {
          const int t0 = t * kTileN;
          const int tlen = max(0, min(kTileN, item.kv_len - t0));
          // ---- KV tile: loaded once per item when the item is a single tile, else once per (group, tile)
          if (ntile > 1 || gi == 0) {
            mbar_wait(bar(KV_EMPTY + kv.stage), kv.phase ^ 1);
            int64_t pg[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = lane + 32 * j;
              pg[j] = 0;
              if (n < tlen) {
                const int64_t e = item.kv_off + t0 + n;
                pg[j] = p.kv_idx_bytes == 8 ? reinterpret_cast<const int64_t*>(p.kv_idx)[e]
                                            : (int64_t) reinterpret_cast<const int32_t*>(p.kv_idx)[e];
              }
            }
            const __half* src_base = (pw == 0 ? p.k : p.v) + (int64_t)hkv * p.kv_head_stride;
            const uint32_t dst_base = base + L::kKv + (kv.stage * 2 + pw) * L::kOperandBytes;
            constexpr int TOK_PER_INSTR = 32 / CH;  // tokens covered by one warp-wide copy
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll 4
              for (int i = 0; i < 32 / TOK_PER_INSTR; ++i) {
                const int n = j * 32 + i * TOK_PER_INSTR + lane / CH;
                const int ch = lane % CH;
                const int64_t page = __shfl_sync(0xffffffffu, pg[j], n & 31);
                const bool ok = n < tlen;
                cp_async_16(dst_base + tile_off(n, ch), src_base + page * p.kv_tok_stride + ch * 8, ok ? 16u : 0u);
              }
            }
            cp_async_arrive(bar(KV_FULL + kv.stage));
            kv.next<kKvStages>();
          }
          // ---- mask words of (group, tile): bit q = query q of the group attends; 0 past the tile end
          {
            mbar_wait(bar(M_EMPTY + mr.stage), mr.phase ^ 1);
            uint32_t* ms = reinterpret_cast<uint32_t*>(gbase + L::kMask) + mr.stage * kTileN;
#pragma unroll
            for (int j = 0; j < kTileN / kProducerThreads; ++j) {
              const int n = pt + j * kProducerThreads;
              uint32_t m = 0;
              if (n < tlen) m = grp.mask_off >= 0 ? (uint32_t)p.masks[grp.mask_off + t0 + n] : 0xffffffffu;
              ms[n] = m;
            }
            mbar_arrive(bar(M_FULL + mr.stage));
            mr.next<kMaskStages>();
          }
        }
      }
    }
  } else {
    // ============================== softmax / MMA issue / epilogue ==============================
    const int r = tid;             // my row == my TMEM lane
    const int qi = r / G, g = r % G;
    const uint32_t qbit = qi < 32 ? (1u << qi) : 0u;   // my query's bit in the per-token masks
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t t_s = t_lane, t_o = t_lane + 128;
    const float c = p.scale * 1.4426950408889634f;  // scores are handled in the log2 domain
    Ring kv, qr, mr;
    uint32_t s_phase = 0, o_phase = 0;
    const uint32_t p_smem = base + L::kP;

    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int item_id = u / p.HKV, hkv = u % p.HKV;
      const deft_item_t item = p.items[item_id];
      const int ntile = max(1, (item.kv_len + kTileN - 1) / kTileN);
      for (int gi = 0; gi < item.n_grp; ++gi) {
        const deft_group_t grp = p.groups[item.grp_off + gi];
        float m_run = -INFINITY, l_run = 0.f;
        for (int t = 0; t < ntile; ++t) {
          const bool new_kv = ntile > 1 || gi == 0;
          const bool last_kv_use = ntile > 1 || gi == item.n_grp - 1;
          // after a single-tile item's first group the KV ring has already advanced: look one back
          int kv_stage = kv.stage;
          if (!new_kv) kv_stage = kv.stage == 0 ? kKvStages - 1 : kv.stage - 1;
          const uint32_t k_smem = base + L::kKv + (kv_stage * 2 + 0) * L::kOperandBytes;
          const uint32_t v_smem = base + L::kKv + (kv_stage * 2 + 1) * L::kOperandBytes;
          const uint32_t q_smem = base + L::kQ + qr.stage * L::kOperandBytes;

          // ---- S = Q K^T
          if (tid == 0) {
            if (t == 0) mbar_wait(bar(Q_FULL + qr.stage), qr.phase);
            if (new_kv) mbar_wait(bar(KV_FULL + kv.stage), kv.phase);
            fence_proxy_async();
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint32_t koff = (ks >> 2) * kPanelBytes + (ks & 3) * 32;
              umma_f16(tmem, smem_desc_sw128(q_smem + koff, 16, 1024), smem_desc_sw128(k_smem + koff, 16, 1024),
                       kIdescQK, ks > 0);
            }
            umma_commit(bar(S_FULL));
            if (t == ntile - 1) umma_commit(bar(Q_EMPTY + qr.stage));  // Q is only read by these MMAs
          }
          if (new_kv) kv.next<kKvStages>();

          // ---- softmax of my row
          mbar_wait(bar(M_FULL + mr.stage), mr.phase);
          const uint32_t* ms = reinterpret_cast<const uint32_t*>(gbase + L::kMask) + mr.stage * kTileN;
          mbar_wait(bar(S_FULL), s_phase);
          s_phase ^= 1;
          tc_fence_after();
          float v[32];
          float m_tile = -INFINITY;
          const bool dbg = p.dbg != nullptr && u == 0 && gi == 0 && t == 0;
#pragma unroll 1
          for (int cb = 0; cb < kTileN / 32; ++cb) {
            tmem_ld32(t_s + cb * 32, v);
            if (dbg)
              for (int j = 0; j < 32; ++j) p.dbg[r * kTileN + cb * 32 + j] = v[j];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const uint4 mk = *reinterpret_cast<const uint4*>(ms + cb * 32 + j);
              if (mk.x & qbit) m_tile = fmaxf(m_tile, v[j]);
              if (mk.y & qbit) m_tile = fmaxf(m_tile, v[j + 1]);
              if (mk.z & qbit) m_tile = fmaxf(m_tile, v[j + 2]);
              if (mk.w & qbit) m_tile = fmaxf(m_tile, v[j + 3]);
            }
          }
          m_tile *= c;  // c > 0
          const float m_new = fmaxf(m_run, m_tile);
          const float m_use = m_new == -INFINITY ? 0.f : m_new;
          const float alpha = exp2f(m_run - m_use);  // 0 on the first live tile
          if (t > 0) {
            // previous P V has completed (o_full was waited below); rescale the running output
            if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
              for (int cb = 0; cb < D / 32; ++cb) {
                tmem_ld32(t_o + cb * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= alpha;
                tmem_st32(t_o + cb * 32, v);
              }
            }
          }
          float psum = 0.f;
#pragma unroll 1
          for (int cb = 0; cb < kTileN / 32; ++cb) {
            tmem_ld32(t_s + cb * 32, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const uint4 mk = *reinterpret_cast<const uint4*>(ms + cb * 32 + j);
              v[j] = (mk.x & qbit) ? exp2f(fmaf(v[j], c, -m_use)) : 0.f;
              v[j + 1] = (mk.y & qbit) ? exp2f(fmaf(v[j + 1], c, -m_use)) : 0.f;
              v[j + 2] = (mk.z & qbit) ? exp2f(fmaf(v[j + 2], c, -m_use)) : 0.f;
              v[j + 3] = (mk.w & qbit) ? exp2f(fmaf(v[j + 3], c, -m_use)) : 0.f;
              psum += (v[j] + v[j + 1]) + (v[j + 2] + v[j + 3]);
            }
            // P row -> K-major SW128 tile: 4 chunks of 8 halves
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint4 pk;
              __half2 h0 = __floats2half2_rn(v[q4 * 8 + 0], v[q4 * 8 + 1]);
              __half2 h1 = __floats2half2_rn(v[q4 * 8 + 2], v[q4 * 8 + 3]);
              __half2 h2 = __floats2half2_rn(v[q4 * 8 + 4], v[q4 * 8 + 5]);
              __half2 h3 = __floats2half2_rn(v[q4 * 8 + 6], v[q4 * 8 + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&h0);
              pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2);
              pk.w = *reinterpret_cast<uint32_t*>(&h3);
              const uint32_t dst = p_smem + tile_off(r, cb * 4 + q4);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
            }
          }
          l_run = l_run * alpha + psum;
          m_run = m_new;
          mbar_arrive(bar(M_EMPTY + mr.stage));
          mr.next<kMaskStages>();
          fence_proxy_async();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
          tc_fence_before();    // my TMEM loads/stores are ordered before the MMAs issued after the barrier
          compute_bar();

          // ---- O (+)= P V
          if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < kTileN / 16; ++ks) {
              const uint32_t poff = (ks >> 2) * kPanelBytes + (ks & 3) * 32;
              umma_f16(tmem + 128, smem_desc_sw128(p_smem + poff, 16, 1024),
                       smem_desc_sw128(v_smem + ks * 2048, kPanelBytes, 1024), kIdescPV, t > 0 || ks > 0);
            }
            umma_commit(bar(O_FULL));
            if (last_kv_use) umma_commit(bar(KV_EMPTY + kv_stage));
          }
          mbar_wait(bar(O_FULL), o_phase);
          o_phase ^= 1;
          tc_fence_after();
        }
        qr.next<kQStages>();

        // ---- epilogue: po[row][h][:] = O / l, plse[row][h] = ln-domain log-sum-exp
        const bool live = qi < grp.q_cnt;
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        float* dst = p.po + (((int64_t)grp.part_base + qi) * p.H + hkv * G + g) * D;
        float v[32];
#pragma unroll 1
        for (int cb = 0; cb < D / 32; ++cb) {
          tmem_ld32(t_o + cb * 32, v);
          if (p.dbg != nullptr && u == 0 && gi == 0)
            for (int j = 0; j < 32; ++j) p.dbg[kRows * kTileN + r * D + cb * 32 + j] = v[j];
          if (live) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + cb * 32 + j) = make_float4(v[j] * inv, v[j + 1] * inv, v[j + 2] * inv, v[j + 3] * inv);
          }
        }
        if (live)
          p.plse[((int64_t)grp.part_base + qi) * p.H + hkv * G + g] =
              l_run > 0.f ? (m_run + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
        tc_fence_before();  // O is overwritten by the next group's first P V (issued after the next compute_bar)
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
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
  const int units = p.n_items * p.HKV;
  const int grid = units < num_sms ? units : num_sms;
  stage1_umma_kernel<D, G><<<grid, kThreads, L::kAlloc, stream>>>(p);
  DEFT_CUDA(cudaGetLastError());
  return DEFT_OK;
}

}  // namespace

bool stage1_umma_supported(const AttnParams& p) {
  const int G = p.H / p.HKV;
  return (p.D == 128 || p.D == 64) && (G == 1 || G == 2 || G == 4);
}

int launch_stage1_umma(const AttnParams& p, cudaStream_t stream) {
  if (p.n_items <= 0) return DEFT_OK;
  const int G = p.H / p.HKV;
#define DEFT_CASE(DD, GG) \
  if (p.D == DD && G == GG) return launch_t<DD, GG>(p, stream);
  DEFT_CASE(128, 4) DEFT_CASE(128, 2) DEFT_CASE(128, 1) DEFT_CASE(64, 4) DEFT_CASE(64, 2) DEFT_CASE(64, 1)
#undef DEFT_CASE
  set_error("tcgen05 stage 1 does not cover head_dim %d / GQA group %d", p.D, G);
  return DEFT_E_ARG;
}

}  // namespace deft
