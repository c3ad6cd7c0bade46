// 2-CTA (cta_group::2) variant of the persistent pointwise GEMM: a CLUSTER of two CTAs (two SMs of one TPC) computes a
// 256(Cout) x 256(frames) tile per step with ONE tcgen05.mma.cta_group::2 stream issued by the leader CTA.
//
// Why: with cta_group::1 every M128 x N256 x K16 MMA re-reads 4 KB of A and 8 KB of B from its SM's shared memory per
// 128 tensor cycles; measured on B200 (pwgemm2.cu with the TMA loads disabled) that caps the tensor pipe at ~58-64 %
// busy.  In pair mode each SM holds its own 128 rows of A and only HALF of the B tile (the hardware reads the other
// half from the peer's shared memory), so per-SM operand fill drops to 32 KB per k-chunk and the accumulator
// (128 lanes x 256 columns per CTA) can be double buffered: the epilogue of tile i overlaps the MMAs of tile i+1.
//
//   both CTAs   warp 0: TMA producer for its own A rows / B half, signalling the LEADER's full barrier
//               warps 4-11: epilogue for its own 128 output channels (TMEM -> bf16 -> swizzled staging -> TMA store)
//   leader CTA  warp 1: MMA issuer; tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs
#include "ts_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"
#include <type_traits>

namespace ts {
namespace pw3 {

constexpr int BM = 256, BK = 64, UMMA_K = 16;  // tile of the CTA PAIR: 256 (Cout) x BN (frames), BN = 256 or 128 (template)
constexpr int A_BYTES = 128 * BK * 2;          // this CTA's 128 weight rows: 16 KB
constexpr int STAGES = 6;                      // streaming mode, BN = 256: 6 stages of (A rows + B half) = 32 KB
constexpr int SMALL_STAGES = 3;                // BN = 128 ("small footprint", see launch_pw_gemm_pair): 3 stages of 24 KB
constexpr int UNITS = 12;                      // 16 KB units next to the epilogue staging: resident A chunks + B-only stages
constexpr int MAX_STAGES = 12;
constexpr int ACC = 2;
constexpr int STG_BYTES = 32 * 128;
// every epilogue warp owns 32 output channels x BN / (EW / 4) frames of a tile (EW = epilogue warps, 4 per column block)
__host__ __device__ constexpr int threads(int ew) { return 128 + 32 * ew; }
__host__ __device__ constexpr int b_bytes(int bn) { return BK * (bn / 2) * 2; }   // this CTA's half of the activation tile
__host__ __device__ constexpr int smem_bytes(int bn, int ew, int nst) {
  return nst * (A_BYTES + b_bytes(bn)) + ew * STG_BYTES + 256 + 1024;
}
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared::cluster address -> leader CTA

struct Params {
  CUtensorMap a0, b0, a1, b1, out;
  int kc0, kc1;
  int Cout, T, B;
  int m_tiles, n_tiles, num_tiles;
  const float* shift;
  const int32_t* lens;
  float* stats;        // optional [B, Cout, 2 * n_tiles, 2]: (sum, sum of squares) of the STORED bf16 values per 128-frame block
  int out_pitch;
  int relu;
  unsigned long long* pool;
  const float* se_scale;
  const __nv_bfloat16* y1;
  int y1_pitch;
  int f16;            // operands and 16-bit output rows are IEEE fp16 instead of bf16
  int res_kc;         // > 0: WEIGHT-STATIONARY mode -- this CTA's 128 weight rows x all res_kc k-chunks stay in shared
                      // memory for the whole kernel (every pair owns ONE m-tile) and the ring streams activations only
  int nst;            // ring stages (6 x 32 KB streaming, 12 - res_kc x 16 KB weight-stationary)
  int rev;            // walk the (utterance, frame-tile) space from the far end (see next_walk_reversed())
  int dbg;            // timing experiment: 32 = the epilogue only drains TMEM (no math, no stores)
  unsigned long long* trace;   // ts_trace slot of this launch (nullptr: off)
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(ptx::smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(leader_bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(leader_bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mma2_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma2_commit_mcast(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// <256, 8>  the streaming kernel;
// <128, 4>  the small-footprint variant: 256 threads at <= 128 registers (half the register file), 256 TMEM columns.
// (<128, 8> -- finer tiles with all resources for launches with only 1-3 tiles per pair -- was measured and removed:
//  QuartzNet at 16 / 32 / 64 utterances 1.79 / 2.33 / 3.61 -> 1.93 / 2.72 / 3.72 ms.  A 256 x 128 tile re-reads the 256
//  weight rows for half the columns, and at these sizes the kernel is bound by operand bytes per SM, not by tile quanta.)
template <int BN, int EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(threads(EPI_WARPS), EPI_WARPS == 8 ? 1 : 2)
pw_gemm_pair_kernel(const __grid_constant__ Params p) {
  constexpr int CW = BN / (EPI_WARPS / 4);   // frames per epilogue warp: 128 or 64
  constexpr int B_BYTES = b_bytes(BN);
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = ACC * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // [resident A chunks | ring] then the staging tiles (BN = 256: 6 x 32 KB == UNITS x 16 KB in either mode)
  uint8_t* stg_base = smem + (BN == 256 ? STAGES * STAGE_BYTES : p.nst * STAGE_BYTES);   // 1024-aligned either way
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + EPI_WARPS * STG_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + ACC;
  uint64_t* a_full = tmem_empty + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);
  const bool resident = p.res_kc > 0;
  uint8_t* ring = smem + p.res_kc * A_BYTES;
  const int ring_stage = resident ? B_BYTES : STAGE_BYTES;
  const int nst = p.nst;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.kc0 + p.kc1;
  const uint32_t rank = cluster_ctarank();     // 0 = leader
  const int tw = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 2 ? 1 : -1);   // traced CTAs: leaders of the first / last pair
  if (threadIdx.x == 0) {
    trace_head(p.trace, tw, 1);
    trace_stamp(p.trace, tw, 1);
    trace_cta(p.trace, p.dbg, blockIdx.x, 0);
  }
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  // i-th tile of this pair -> (m-tile, frame tile, utterance).  Streaming: tiles are dealt round-robin.  Weight-stationary:
  // pair p owns m-tile p % m_tiles and strides over the (frame tile, utterance) space with the pairs that share it.
  const int nb_tiles = p.num_tiles / p.m_tiles;
  const int pairs_per_mt = npairs / p.m_tiles;
  auto tile_at = [&](int i, int& mt, int& nt, int& b) -> bool {
    int rest;
    if (resident) {
      if (pair >= pairs_per_mt * p.m_tiles) return false;
      rest = pair / p.m_tiles + i * pairs_per_mt;
      if (rest >= nb_tiles) return false;
      mt = pair % p.m_tiles;
    } else {
      const int tile = pair + i * npairs;
      if (tile >= p.num_tiles) return false;
      mt = tile % p.m_tiles;
      rest = tile / p.m_tiles;
    }
    if (p.rev) rest = nb_tiles - 1 - rest;
    nt = rest % p.n_tiles;
    b = rest / p.n_tiles;
    return true;
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.a0);
    ptx::prefetch_tensormap(&p.b0);
    ptx::prefetch_tensormap(&p.out);
    if (p.kc1 > 0) {
      ptx::prefetch_tensormap(&p.a1);
      ptx::prefetch_tensormap(&p.b1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 2);      // only the leader's copy is used: one arrival per CTA + both CTAs' bytes
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(a_full, 2);
    for (int a = 0; a < ACC; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2 * EPI_WARPS * 32);   // leader's copy: epilogue threads of both CTAs
    }
    ptx::fence_barrier_init();
  }
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / before the paired TMEM allocation
  if (warp == 2) {
    tmem2_alloc(tmem_slot, TMEM_COLS);
    tmem2_relinquish();
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(p.trace, tw, 2);
  // PDL: everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail; from here on we
  // read its output.  Let the next kernel start its own prologue as soon as all our CTAs got this far.
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) trace_stamp(p.trace, tw, 3);

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: runs ahead over this CTA's whole tile list =====
    uint32_t cnt = 0;
    long long w12 = 0;
    int mt, nt, b;
    if (resident && tile_at(0, mt, nt, b)) {   // the weights of this pair's m-tile: loaded ONCE, all k-chunks
      const int m0 = mt * BM + 128 * (int)rank;
      if (rank == 0)
        ptx::mbar_arrive_expect_tx(a_full, 2 * p.res_kc * A_BYTES);
      else
        mbar_arrive_remote(a_full, 0);
      for (int kc = 0; kc < num_k; ++kc) {
        const bool seg1 = kc >= p.kc0;
        tma2_load_2d(smem + kc * A_BYTES, seg1 ? &p.a1 : &p.a0, a_full, (seg1 ? kc - p.kc0 : kc) * BK, m0);
      }
    }
    for (int i = 0; tile_at(i, mt, nt, b); ++i) {
      const int m0 = mt * BM + 128 * (int)rank;          // this CTA's 128 output channels
      const int t0 = nt * BN + (BN / 2) * (int)rank;     // this CTA's half of the frame tile
      for (int kc = 0; kc < num_k; ++kc, ++cnt) {
        const int s = cnt % nst;
        TS_TIMED_WAIT((p.trace != nullptr && tw == 0), w12, ptx::mbar_wait(&empty_bar[s], ((cnt / nst) & 1) ^ 1));
        uint8_t* sa = ring + s * ring_stage;
        uint8_t* sb = resident ? sa : sa + A_BYTES;
        // dbg 16 / 8 (timing experiments, results wrong): skip the weight (A) / activation (B) loads of every stage
        const bool skip_a = (p.dbg & 16) != 0, skip_b = (p.dbg & 8) != 0;
        if (rank == 0)
          ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * (ring_stage - (skip_a && !resident ? A_BYTES : 0) - (skip_b ? B_BYTES : 0)));
        else
          mbar_arrive_remote(&full_bar[s], 0);
        const bool seg1 = kc >= p.kc0;
        const CUtensorMap* ma = seg1 ? &p.a1 : &p.a0;
        const CUtensorMap* mb = seg1 ? &p.b1 : &p.b0;
        const int k0 = (seg1 ? kc - p.kc0 : kc) * BK;
        if (!resident && !skip_a) tma2_load_2d(sa, ma, &full_bar[s], k0, m0);       // [128 rows x 64 k]
#pragma unroll
        for (int j = 0; j < BN / 128; ++j)     // 64-frame boxes of this CTA's half of the frame tile
          if (!skip_b) tma2_load_3d(sb + j * (BK * 128), mb, &full_bar[s], t0 + 64 * j, k0, b);
      }
    }
    trace_put(p.trace, tw, 12, w12);
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ===== MMA issuer (leader CTA only) =====
    const uint32_t idesc = ptx::umma_idesc_16(256, BN, 0, 1, p.f16);
    uint32_t cnt = 0;
    int mt, nt, b;
    long long w10 = 0, w11 = 0;
    const long long t_loop0 = trace_clock();     // whole life of the issue loop = denominator of the stall shares
    if (resident && tile_at(0, mt, nt, b)) {
      ptx::mbar_wait(a_full, 0);
      ptx::tc_fence_after();
    }
    for (int it = 0; tile_at(it, mt, nt, b); ++it) {
      const int a = it % ACC;
      TS_TIMED_WAIT((p.trace != nullptr && tw == 0), w11, ptx::mbar_wait(&tmem_empty[a], ((it / ACC) & 1) ^ 1));   // epilogue has drained this accumulator set
      ptx::tc_fence_after();
      for (int kc = 0; kc < num_k; ++kc, ++cnt) {
        const int s = cnt % nst;
        TS_TIMED_WAIT((p.trace != nullptr && tw == 0), w10, ptx::mbar_wait(&full_bar[s], (cnt / nst) & 1));
        ptx::tc_fence_after();
        if (cnt == 0) trace_stamp(p.trace, tw, 4);
        const uint32_t st = ptx::smem_u32(ring + s * ring_stage);
        const uint32_t sa = resident ? ptx::smem_u32(smem + kc * A_BYTES) : st;
        const uint32_t sb = resident ? st : st + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t db = ptx::umma_desc(sb + k * 2048, BK * 128, 1024);
          const uint64_t da = ptx::umma_desc(sa + k * 32, 0, 1024);
          mma2_bf16_ss(tmem_base + a * BN, da, db, idesc, (kc > 0 || k > 0) ? 1u : 0u);
        }
        mma2_commit_mcast(&empty_bar[s]);
      }
      mma2_commit_mcast(&tmem_full[a]);
    }
    trace_stamp(p.trace, tw, 5);
    trace_put(p.trace, tw, 10, w10);
    trace_put(p.trace, tw, 11, w11);
    trace_put(p.trace, tw, 15, trace_clock() - t_loop0);
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int e = warp - 4;
    const int q = warp & 3;          // TMEM lane quarter accessible to this warp
    const int h = e >> 2;            // column block [CW h, CW (h + 1)) of the tile handled by this warp
    uint8_t* stg = stg_base + e * STG_BYTES;
    const uint32_t stg_row = ptx::smem_u32(stg) + lane * 128;
    long long w13 = 0, w14 = 0;
    int mt, nt, b;
    for (uint32_t it = 0; tile_at((int)it, mt, nt, b); ++it) {
      const int t0 = nt * BN + h * CW;
      const int mrow0 = mt * BM + (int)rank * 128 + q * 32;
      const int m = mrow0 + lane;
      const bool m_ok = m < p.Cout;
      const float shift = (m_ok && p.shift) ? p.shift[m] : 0.f;
      const int len = p.lens ? min(p.lens[b], p.T) : p.T;
      const float gate = (m_ok && p.se_scale) ? p.se_scale[(size_t)b * p.Cout + m] : 0.f;
      float pooled = 0.f;
      float st_s = 0.f, st_ss = 0.f;
      const int a = it % ACC;
      TS_TIMED_WAIT((p.trace != nullptr && tw == 0 && threadIdx.x == 128), w13, ptx::mbar_wait(&tmem_full[a], (it / ACC) & 1));
      ptx::tc_fence_after();
      if (it == 0 && threadIdx.x == 128) trace_stamp(p.trace, tw, 6);
#pragma unroll 1
      for (int cc = 0; cc < CW / 64; ++cc) {
        uint32_t v[64];
        __syncwarp();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + h * CW + cc * 64);
        ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        ptx::tmem_ld_wait();
        if (cc == CW / 64 - 1) {  // this thread's part of the accumulator is read: tell the leader's MMA warp
          ptx::tc_fence_before();
          if (rank == 0)
            ptx::mbar_arrive(&tmem_empty[a]);
          else
            mbar_arrive_remote(&tmem_empty[a], 0);
        }
        const int tb = t0 + cc * 64;
        if (mrow0 < p.Cout && tb < p.out_pitch && !(p.dbg & 32)) {   // warp-uniform: something of this 32 x 64 block is stored
          float r[64];
#pragma unroll
          for (int j = 0; j < 64; ++j) r[j] = __uint_as_float(v[j]) + shift;
          if (p.pool) {
#pragma unroll
            for (int j = 0; j < 64; ++j) pooled += (tb + j < p.T) ? r[j] : 0.f;
          }
          if (p.y1 && m_ok) {
            const uint4* yp = reinterpret_cast<const uint4*>(p.y1 + ((size_t)b * p.Cout + m) * p.y1_pitch + tb);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const uint4 u = yp[g];
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int hh = 0; hh < 4; ++hh) {
                const float2 f = unpack16x2(w[hh], p.f16 != 0);
                r[g * 8 + 2 * hh] += gate * f.x;
                r[g * 8 + 2 * hh + 1] += gate * f.y;
              }
            }
          }
          if (tb + 64 > len) {   // block crosses the utterance end (warp-uniform): zero the tail
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (tb + j >= len) r[j] = 0.f;
          }
          if (lane == 0) TS_TIMED_WAIT((p.trace != nullptr && tw == 0 && threadIdx.x == 128), w14, bulk_wait_read0());   // previous TMA store has finished reading the staging tile
          __syncwarp();
          // (uniform branches on row format / activation so that each path carries exactly one conversion per pair; the
          // ReLU rides in the conversion instruction: F2FP.RELU)
          // Tried and rejected (round 2): storing each thread's 128-byte output line straight from registers with four
          // 256-bit stores (STG.256) instead of staging + TMA store.  32 lanes x 32 different lines per instruction is the
          // worst case for the LSU write path: 512 -> 1024 at 256 x 751 frames 172 -> 222 us, 512 x 512 96 -> 108 us.
          auto pack_and_stage = [&](auto f16_tag, auto relu_tag) {
            constexpr bool kF16 = decltype(f16_tag)::value, kRelu = decltype(relu_tag)::value;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              uint32_t pk[4];
#pragma unroll
              for (int hh = 0; hh < 4; ++hh) {
                const int j = g * 8 + 2 * hh;
                pk[hh] = kF16 ? (kRelu ? pack_f16x2_relu(r[j], r[j + 1]) : pack_f16x2(r[j], r[j + 1]))
                              : (kRelu ? pack_bf16x2_relu(r[j], r[j + 1]) : pack_bf16x2(r[j], r[j + 1]));
                if (!kF16 && !kRelu && p.stats) {
                  const float2 f = unpack_bf16x2(pk[hh]);
                  st_s += f.x + f.y;
                  st_ss = fmaf(f.x, f.x, fmaf(f.y, f.y, st_ss));
                }
              }
              // SWIZZLE_128B: 16-byte chunk g of row `lane` lives at chunk (g ^ (row & 7))
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row + ((g ^ (lane & 7)) << 4)),
                           "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                           : "memory");
            }
          };
          if (p.f16) {
            if (p.relu) pack_and_stage(std::true_type{}, std::true_type{});
            else pack_and_stage(std::true_type{}, std::false_type{});
          } else {
            if (p.relu) pack_and_stage(std::false_type{}, std::true_type{});
            else pack_and_stage(std::false_type{}, std::false_type{});
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&p.out, stg, tb, mrow0, b);
            bulk_commit();
          }
        }
      }
      if (p.pool && m_ok) se_pool_add(p.pool + (size_t)b * p.Cout + m, pooled);
      if (BN == 256 && EPI_WARPS == 8 && p.stats && m_ok)   // one slot per (row, 128-frame block): written exactly once, no atomics (deterministic)
        *reinterpret_cast<float2*>(p.stats + ((((size_t)b * p.Cout + m) * (2 * p.n_tiles)) + nt * 2 + h) * 2) =
            make_float2(st_s, st_ss);
    }
    if (threadIdx.x == 128) {
      trace_stamp(p.trace, tw, 7);
      trace_put(p.trace, tw, 13, w13);
      trace_put(p.trace, tw, 14, w14);
    }
    if (lane == 0) bulk_wait0();   // all stores of this warp are complete before the CTA exits
    if (threadIdx.x == 128) trace_stamp(p.trace, tw, 8);
  }
  ptx::tc_fence_before();
  cluster_sync_all();   // the peer's shared memory / TMEM stay alive until both CTAs are done
  if (warp == 2) {
    ptx::tc_fence_after();
    tmem2_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) {
    trace_stamp(p.trace, tw, 9);
    trace_cta(p.trace, p.dbg, blockIdx.x, 1);
  }
}

}  // namespace pw3

int option_pw_resident();
int option_pw_ws();
int launch_pw_gemm_ws(const void* w0, const void* x0, int cin0, int x0_pitch, const void* w1, const void* x1, int cin1,
                      int x1_pitch, int B, int Cout, int T, const float* shift, const int32_t* lens, void* out,
                      int out_pitch, int relu, unsigned long long* pool, const float* se_scale, const void* y1, int y1_pitch,
                      cudaStream_t st, int f16, int wconst);
int small_footprint(long long frames);

// bf16-row outputs with Cout > 128 on CTA pairs; TS_ERR_UNSUPPORTED otherwise (caller falls back)
int launch_pw_gemm_pair(const void* w0, const void* x0, int cin0, int x0_pitch, const void* w1, const void* x1, int cin1,
                        int x1_pitch, int B, int Cout, int T, const float* shift, const int32_t* lens, void* out,
                        int out_pitch, int relu, unsigned long long* pool, const float* se_scale, const void* y1, int y1_pitch,
                        cudaStream_t st, float* stats, int f16, int wconst) {
  if (Cout <= 128 || out_pitch % 64 != 0) return TS_ERR_UNSUPPORTED;
  // K <= 512: the weight-stationary kernel with the weights in tensor memory (pwgemm4.cu)
  if (stats == nullptr && option_pw_ws() != 0 && small_footprint((long long)B * out_pitch) == 0) {
    const int rc = launch_pw_gemm_ws(w0, x0, cin0, x0_pitch, w1, x1, cin1, x1_pitch, B, Cout, T, shift, lens, out, out_pitch,
                                     relu, pool, se_scale, y1, y1_pitch, st, f16, wconst);
    if (rc != TS_ERR_UNSUPPORTED) return rc;
  }
  pw3::Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = tma::make_2d_bf16(&p.a0, w0, cin0, Cout, (uint64_t)cin0 * 2, pw3::BK, 128)) != TS_OK) return rc;
  if ((rc = tma::make_3d_bf16(&p.b0, x0, T, cin0, B, (uint64_t)x0_pitch * 2, (uint64_t)cin0 * x0_pitch * 2, 64, pw3::BK,
                              1)) != TS_OK)
    return rc;
  p.kc0 = ceil_div(cin0, pw3::BK);
  if (cin1 > 0) {
    if ((rc = tma::make_2d_bf16(&p.a1, w1, cin1, Cout, (uint64_t)cin1 * 2, pw3::BK, 128)) != TS_OK) return rc;
    if ((rc = tma::make_3d_bf16(&p.b1, x1, T, cin1, B, (uint64_t)x1_pitch * 2, (uint64_t)cin1 * x1_pitch * 2, 64,
                                pw3::BK, 1)) != TS_OK)
      return rc;
    p.kc1 = ceil_div(cin1, pw3::BK);
  }
  if ((rc = tma::make_3d_bf16(&p.out, out, out_pitch, Cout, B, (uint64_t)out_pitch * 2, (uint64_t)Cout * out_pitch * 2,
                              64, 32, 1)) != TS_OK)
    return rc;
  // Small-footprint mode (option small, off by default): 256 x 128 tiles, 3 stages of 24 KB, 4 epilogue warps and 256 TMEM
  // columns -- 89 KB of shared memory, half of the registers and of the tensor memory, so that the CTAs of this launch
  // are RESIDENT next to the (equally slimmed) Toeplitz CTAs of the previous launch and run their prologue under its tail.
  // Measured on B200 (tools/ab_small.sh): the overlap is real (tools/trace_chain.py: entry 2.5 us BEFORE the predecessor
  // ends instead of 2.5 us after) but with half the shared memory each kernel has half the bytes in flight: QuartzNet at
  // 32 utterances 2.30 -> 3.27 ms, at 256 11.2 -> 19.2 ms.  Kept as an A/B knob.
  const bool small = stats == nullptr && small_footprint((long long)B * out_pitch) != 0;
  const int BN = small ? 128 : 256;
  p.Cout = Cout; p.T = T; p.B = B;
  p.m_tiles = ceil_div(Cout, pw3::BM);
  p.n_tiles = ceil_div(out_pitch, BN);
  p.num_tiles = p.m_tiles * p.n_tiles * B;
  p.shift = shift; p.lens = lens; p.out_pitch = out_pitch; p.relu = relu;
  p.stats = stats;
  p.f16 = f16;
  p.rev = next_walk_reversed();
  p.dbg = option_dbg();
  // Weight-stationary mode (option pw_resident, OFF by default): all k-chunks of this CTA's weight rows stay in shared
  // memory (K <= 512) and the ring streams activations only.  It removes the weight re-reads (~40 % of the L2 -> SM bytes
  // of a 512 x 512 layer) but leaves room for only 4 activation stages of 16 KB: measured on B200 (ncu, 512 x 512 layer at
  // 256 x 751 frames) 90.2 -> 99.4 us, tensor pipe 64 -> 56 % busy, L2 throughput 43 -> 30 % of peak -- the kernel is bound
  // by bytes in flight, not by L2 bandwidth, so the 6 x 32 KB ring stays the default.
  const int kc_total = p.kc0 + p.kc1;
  p.res_kc = (!small && option_pw_resident() && kc_total <= pw3::UNITS - 4) ? kc_total : 0;
  p.nst = small ? pw3::SMALL_STAGES : (p.res_kc ? pw3::UNITS - p.res_kc : pw3::STAGES);
  p.pool = pool; p.se_scale = se_scale; p.y1 = reinterpret_cast<const __nv_bfloat16*>(y1); p.y1_pitch = y1_pitch;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    TS_CUDA(cudaGetDevice(&dev));
    TS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    TS_CUDA(cudaFuncSetAttribute(pw3::pw_gemm_pair_kernel<256, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 pw3::smem_bytes(256, 8, pw3::STAGES)));
    TS_CUDA(cudaFuncSetAttribute(pw3::pw_gemm_pair_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 pw3::smem_bytes(128, 4, pw3::SMALL_STAGES)));
  }
  int pairs = num_sms / 2;
  if (p.num_tiles < pairs) pairs = p.num_tiles;
  p.trace = trace_next_slot(1, 2 * pairs);
  if (small)
    TS_CUDA(launch_pdl(pw3::pw_gemm_pair_kernel<128, 4>, dim3(2 * pairs), dim3(pw3::threads(4)),
                       pw3::smem_bytes(128, 4, pw3::SMALL_STAGES), st, option_pdl() != 0, p));
  else
    TS_CUDA(launch_pdl(pw3::pw_gemm_pair_kernel<256, 8>, dim3(2 * pairs), dim3(pw3::threads(8)),
                       pw3::smem_bytes(256, 8, pw3::STAGES), st, option_pdl() != 0, p));
  TS_LAUNCH_CHECK("pw_gemm_pair_kernel");
  return TS_OK;
}

}  // namespace ts
