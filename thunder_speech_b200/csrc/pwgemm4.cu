// Weight-stationary pair GEMM with the WEIGHTS IN TENSOR MEMORY (tcgen05.mma "ts" form: A from TMEM, B from shared memory).
//
// Why (profiles/r02_gemm_ablation.txt): the streaming pair kernel (pwgemm3.cu) is bound by the SM's 128 B/cycle
// shared-memory port -- per 64-wide k-chunk it carries 32 KB of TMA writes, 32 KB of SS-operand reads and the epilogue's
// staging traffic, 156 B/cycle asked of 128 at K = 512.  The weights of a 1x1 conv are the same for every frame tile, so
// here every CTA pair OWNS one 256-row m-tile for the whole kernel: its CTAs copy their 128 weight rows x K (<= 512) once
// into 256 columns of their tensor memory (global -> registers -> tcgen05.st; row-major bf16 weights ARE the TMEM operand
// image: lane = row, 32-bit column c = elements 2c, 2c + 1), before griddepcontrol.wait, i.e. under the previous kernel's
// tail.  After that only activations move: 8 KB per k-chunk per CTA instead of 32 KB through TMA, and 2 KB instead of 8 KB
// of operand reads per MMA -- the port sees ~96 B/cycle at full tensor rate, and the L2 -> SM bytes halve.
//
//   tile of the pair   256 (Cout) x 128 (frames): M256 x N128 x K16 MMAs, two accumulator stages of 128 columns
//   tensor memory      [0, 256) weights (K/2 columns used), [256, 512) accumulators
//   shared memory      ring of NST x 8 KB activation stages (this CTA's 64 frames x 64 channels, MN-major SW128) + 8 x 4 KB staging
//   both CTAs   warp 0      TMA producer (activations only), signalling the leader's full barrier
//               warps 4-11  weight copy into TMEM, then epilogue: 32 channels x 64 frames per warp and tile
//   leader CTA  warp 1      MMA issuer; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs
//
// Same epilogue semantics as pwgemm3.cu (shift, SE pool / gate, ReLU, tail mask, both row formats); bit-identical results
// are NOT expected only where the accumulation order differs -- it does not: k runs 0..K in both kernels.
#include "ts_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"
#include <type_traits>

namespace ts {
namespace pw4 {

constexpr int BM = 256, BN = 128, BK = 64, UMMA_K = 16;
constexpr int B_BYTES = BK * (BN / 2) * 2;     // this CTA's half of the activation tile: 8 KB
constexpr int NST = 20;                        // activation stages: 160 KB in flight per CTA
constexpr int ACC = 2;
constexpr int EPI_WARPS = 8;
constexpr int STG_BYTES = 32 * 128;
constexpr int THREADS = 128 + 32 * EPI_WARPS;
constexpr int A_COLS = 256;                    // TMEM columns reserved for the weights (K <= 512)
constexpr int TMEM_COLS = 512;
constexpr int MAX_KC = A_COLS * 2 / BK;        // 8 k-chunks
constexpr int SMEM_BYTES = NST * B_BYTES + EPI_WARPS * STG_BYTES + 512 + 1024;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;

struct Params {
  CUtensorMap b0, b1, out;
  const uint16_t* w0;   // [Cout, cin0] 16-bit, row-major
  const uint16_t* w1;   // [Cout, cin1] or nullptr (second K segment: the residual 1x1 conv)
  int cin0, cin1;
  int kc0, kc1;
  int Cout, T, B;
  int m_tiles, n_tiles, nb_tiles, pairs_per_mt;
  const float* shift;
  const int32_t* lens;
  int out_pitch;
  int relu;
  unsigned long long* pool;
  const float* se_scale;
  const __nv_bfloat16* y1;
  int y1_pitch;
  int f16;
  int rev;
  int wconst;   // weights are launch-invariant constants: copy them before griddepcontrol.wait
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
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(leader_bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the weights come from tensor memory (same lanes as D in each CTA of the pair)
__device__ __forceinline__ void mma2_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
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
// 32 lanes x 32 consecutive 32-bit columns <- 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
pw_gemm_ws_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* stg_base = smem + NST * B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + EPI_WARPS * STG_BYTES);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* tmem_full = empty_bar + NST;
  uint64_t* tmem_empty = tmem_full + ACC;
  uint64_t* a_ready = tmem_empty + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.kc0 + p.kc1;
  const uint32_t rank = cluster_ctarank();     // 0 = leader
  const int pair = blockIdx.x >> 1;
  // weight-stationary tiling: pair q owns m-tile q % m_tiles and strides over the (frame tile, utterance) space with the
  // pairs that share it
  const int mt = pair % p.m_tiles;
  const bool active = pair < p.pairs_per_mt * p.m_tiles;
  auto tile_at = [&](int i, int& nt, int& b) -> bool {
    if (!active) return false;
    int rest = pair / p.m_tiles + i * p.pairs_per_mt;
    if (rest >= p.nb_tiles) return false;
    if (p.rev) rest = p.nb_tiles - 1 - rest;
    nt = rest % p.n_tiles;
    b = rest / p.n_tiles;
    return true;
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.b0);
    ptx::prefetch_tensormap(&p.out);
    if (p.kc1 > 0) ptx::prefetch_tensormap(&p.b1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) {
      ptx::mbar_init(&full_bar[s], 2);      // only the leader's copy is used: one arrival per CTA + both CTAs' bytes
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(a_ready, 2);             // leader's copy: one arrival per CTA once its weights sit in tensor memory
    for (int a = 0; a < ACC; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2 * EPI_WARPS * 32);   // leader's copy: epilogue threads of both CTAs
    }
    ptx::fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem2_alloc(tmem_slot, TMEM_COLS);
    tmem2_relinquish();
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base + A_COLS;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: activations only =====
    pdl_wait();
    uint32_t cnt = 0;
    int nt, b;
    for (int i = 0; tile_at(i, nt, b); ++i) {
      const int t0 = nt * BN + (BN / 2) * (int)rank;     // this CTA's half of the frame tile
      for (int kc = 0; kc < num_k; ++kc, ++cnt) {
        const int s = cnt % NST;
        ptx::mbar_wait(&empty_bar[s], ((cnt / NST) & 1) ^ 1);
        if (rank == 0)
          ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * B_BYTES);
        else
          mbar_arrive_remote(&full_bar[s], 0);
        const bool seg1 = kc >= p.kc0;
        tma2_load_3d(ring + s * B_BYTES, seg1 ? &p.b1 : &p.b0, &full_bar[s], t0, (seg1 ? kc - p.kc0 : kc) * BK, b);
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ===== MMA issuer (leader CTA only) =====
    const uint32_t idesc = ptx::umma_idesc_16(256, BN, 0, 1, p.f16);
    uint32_t cnt = 0;
    int nt, b;
    ptx::mbar_wait(a_ready, 0);               // both CTAs' weights are in tensor memory
    ptx::tc_fence_after();
    for (int it = 0; tile_at(it, nt, b); ++it) {
      const int a = it % ACC;
      ptx::mbar_wait(&tmem_empty[a], ((it / ACC) & 1) ^ 1);
      ptx::tc_fence_after();
      for (int kc = 0; kc < num_k; ++kc, ++cnt) {
        const int s = cnt % NST;
        ptx::mbar_wait(&full_bar[s], (cnt / NST) & 1);
        ptx::tc_fence_after();
        const uint32_t sb = ptx::smem_u32(ring + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t db = ptx::umma_desc(sb + k * 2048, BK * 128, 1024);
          mma2_bf16_ts(tmem_acc + a * BN, tmem_base + kc * (BK / 2) + k * (UMMA_K / 2), db, idesc, (kc > 0 || k > 0) ? 1u : 0u);
        }
        mma2_commit_mcast(&empty_bar[s]);
      }
      mma2_commit_mcast(&tmem_full[a]);
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    const int q = warp & 3;          // TMEM lane quarter accessible to this warp
    const int h = e >> 2;            // which half of the work this warp group takes (k-chunks of the copy / columns of a tile)
    // ===== weights -> tensor memory (constants: no dependency on the previous kernel, runs under its tail) =====
    {
      if (!p.wconst) pdl_wait();     // weights produced on this stream (training): they are only final once the predecessor is
      const int m = mt * BM + (int)rank * 128 + q * 32 + lane;      // this thread's weight row = its TMEM lane
      const bool m_ok = active && m < p.Cout;
      for (int kc = h; kc < num_k; kc += 2) {                        // the two warp groups interleave the k-chunks
        const bool seg1 = kc >= p.kc0;
        const int cin = seg1 ? p.cin1 : p.cin0;
        const int k0 = (seg1 ? kc - p.kc0 : kc) * BK;
        const uint16_t* wr = (seg1 ? p.w1 : p.w0) + (size_t)m * cin + k0;
        uint32_t v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {                                // 8 elements (16 bytes) per load; cin % 8 == 0
          uint4 u = make_uint4(0, 0, 0, 0);
          if (m_ok && k0 + 8 * g < cin) u = __ldg(reinterpret_cast<const uint4*>(wr + 8 * g));
          v[4 * g + 0] = u.x; v[4 * g + 1] = u.y; v[4 * g + 2] = u.z; v[4 * g + 3] = u.w;
        }
        tmem_st_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kc * (BK / 2)), v);
      }
      tmem_st_wait();
      ptx::tc_fence_before();
      named_bar_sync(1, EPI_WARPS * 32);
      if (threadIdx.x == 128) {
        if (rank == 0)
          ptx::mbar_arrive(a_ready);
        else
          mbar_arrive_remote(a_ready, 0);
      }
    }
    // ===== epilogue: 32 channels x 64 frames per warp and tile =====
    pdl_wait();                      // no global write (and no y1 read) of this grid may overtake the previous grid
    uint8_t* stg = stg_base + e * STG_BYTES;
    const uint32_t stg_row = ptx::smem_u32(stg) + lane * 128;
    int nt, b;
    for (uint32_t it = 0; tile_at((int)it, nt, b); ++it) {
      const int tb = nt * BN + h * 64;
      const int mrow0 = mt * BM + (int)rank * 128 + q * 32;
      const int m = mrow0 + lane;
      const bool m_ok = m < p.Cout;
      const float shift = (m_ok && p.shift) ? p.shift[m] : 0.f;
      const int len = p.lens ? min(p.lens[b], p.T) : p.T;
      const float gate = (m_ok && p.se_scale) ? p.se_scale[(size_t)b * p.Cout + m] : 0.f;
      float pooled = 0.f;
      const int a = it % ACC;
      ptx::mbar_wait(&tmem_full[a], (it / ACC) & 1);
      ptx::tc_fence_after();
      uint32_t v[64];
      __syncwarp();
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + h * 64);
      ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      if (rank == 0)
        ptx::mbar_arrive(&tmem_empty[a]);
      else
        mbar_arrive_remote(&tmem_empty[a], 0);
      if (mrow0 < p.Cout && tb < p.out_pitch) {   // warp-uniform: something of this 32 x 64 block is stored
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
        if (lane == 0) bulk_wait_read0();   // previous TMA store has finished reading the staging tile
        __syncwarp();
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
            }
            // SWIZZLE_128B: 16-byte chunk g of row `lane` lives at chunk (g ^ (row & 7))
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row + ((g ^ (lane & 7)) << 4)), "r"(pk[0]),
                         "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
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
      if (p.pool && m_ok) se_pool_add(p.pool + (size_t)b * p.Cout + m, pooled);
    }
    if (lane == 0) bulk_wait0();   // all stores of this warp are complete before the CTA exits
  }
  ptx::tc_fence_before();
  cluster_sync_all();   // the peer's shared memory / TMEM stay alive until both CTAs are done
  if (warp == 2) {
    ptx::tc_fence_after();
    tmem2_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace pw4

// K = cin0 + cin1 <= 512 (in 64-wide chunks), 16-bit row outputs, Cout > 128; TS_ERR_UNSUPPORTED otherwise (the streaming
// pair kernel takes over)
int launch_pw_gemm_ws(const void* w0, const void* x0, int cin0, int x0_pitch, const void* w1, const void* x1, int cin1,
                      int x1_pitch, int B, int Cout, int T, const float* shift, const int32_t* lens, void* out,
                      int out_pitch, int relu, unsigned long long* pool, const float* se_scale, const void* y1, int y1_pitch,
                      cudaStream_t st, int f16, int wconst) {
  if (Cout <= 128 || out_pitch % 64 != 0 || cin0 % 8 != 0 || cin1 % 8 != 0) return TS_ERR_UNSUPPORTED;
  pw4::Params p;
  memset(&p, 0, sizeof(p));
  p.kc0 = ceil_div(cin0, pw4::BK);
  p.kc1 = cin1 > 0 ? ceil_div(cin1, pw4::BK) : 0;
  if (p.kc0 + p.kc1 > pw4::MAX_KC) return TS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(w0) & 15) != 0 || (w1 && (reinterpret_cast<uintptr_t>(w1) & 15) != 0)) return TS_ERR_UNSUPPORTED;
  int rc;
  if ((rc = tma::make_3d_bf16(&p.b0, x0, T, cin0, B, (uint64_t)x0_pitch * 2, (uint64_t)cin0 * x0_pitch * 2, 64, pw4::BK,
                              1)) != TS_OK)
    return rc;
  if (cin1 > 0) {
    if ((rc = tma::make_3d_bf16(&p.b1, x1, T, cin1, B, (uint64_t)x1_pitch * 2, (uint64_t)cin1 * x1_pitch * 2, 64,
                                pw4::BK, 1)) != TS_OK)
      return rc;
  }
  if ((rc = tma::make_3d_bf16(&p.out, out, out_pitch, Cout, B, (uint64_t)out_pitch * 2, (uint64_t)Cout * out_pitch * 2,
                              64, 32, 1)) != TS_OK)
    return rc;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    TS_CUDA(cudaGetDevice(&dev));
    TS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    TS_CUDA(cudaFuncSetAttribute(pw4::pw_gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pw4::SMEM_BYTES));
  }
  p.w0 = reinterpret_cast<const uint16_t*>(w0);
  p.w1 = reinterpret_cast<const uint16_t*>(w1);
  p.cin0 = cin0; p.cin1 = cin1;
  p.Cout = Cout; p.T = T; p.B = B;
  p.m_tiles = ceil_div(Cout, pw4::BM);
  p.n_tiles = ceil_div(out_pitch, pw4::BN);
  p.nb_tiles = p.n_tiles * B;
  int pairs = num_sms / 2;
  if (p.m_tiles > pairs) return TS_ERR_UNSUPPORTED;
  p.pairs_per_mt = pairs / p.m_tiles;
  if (p.pairs_per_mt > p.nb_tiles) p.pairs_per_mt = p.nb_tiles;
  pairs = p.pairs_per_mt * p.m_tiles;
  p.shift = shift; p.lens = lens; p.out_pitch = out_pitch; p.relu = relu;
  p.pool = pool; p.se_scale = se_scale; p.y1 = reinterpret_cast<const __nv_bfloat16*>(y1); p.y1_pitch = y1_pitch;
  p.f16 = f16;
  p.wconst = wconst;
  p.rev = next_walk_reversed();
  TS_CUDA(launch_pdl(pw4::pw_gemm_ws_kernel, dim3(2 * pairs), dim3(pw4::THREADS), pw4::SMEM_BYTES, st, option_pdl() != 0, p));
  TS_LAUNCH_CHECK("pw_gemm_ws_kernel");
  return TS_OK;
}

}  // namespace ts
