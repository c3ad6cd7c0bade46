// Toeplitz-MMA depthwise conv, PERSISTENT edition (math: dwmma.cu; TMA tile format: dwmma2.cu).
//
// dwmma2.cu runs one CTA per (channel, tile group): every CTA pays TMEM allocation, barrier set-up, the Toeplitz build of
// its channel (~3 tile-times) and a pipeline ramp for as few as 4 tiles (B = 32 per GPU: 57 % of the CTA's life is useful),
// and C x groups CTAs on 296 slots quantise into waves (measured: 75 % of HBM at B = 256, 40 % at B = 32).  Here the grid
// is 2 CTAs per SM and each CTA walks a CONTIGUOUS, balanced range of the channel-major work list (channel, utterance
// tile): prologue and ramp are paid once per CTA, ranges differ by at most one tile, and the Toeplitz blocks of the NEXT
// channel are built by two dedicated warps into a second buffer while the tensor core still streams the current channel.
//
//   warp 0      TMA producer: one 4-D box per tile (NB utterances x R windows x 64 frames of one channel)
//   warp 1      MMA issuer: M128 x N64 x K16 per non-zero k-slice of the NQ Toeplitz blocks, 4 TMEM accumulator stages;
//               switches Toeplitz buffer at channel boundaries (b_ready / b_free barriers)
//   warps 2-3   Toeplitz builders (taps -> zero-padded dilated tap line in shared memory -> SW128 bf16/fp16 blocks)
//   warps 4-7   epilogue: TMEM -> tail mask -> 16-bit -> swizzled staging -> ONE 4-D TMA store per tile
#include "ts_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"
#include <type_traits>

namespace ts {
namespace dwt3 {

constexpr int L = 64;
constexpr int MROWS = 128;
constexpr int A_STAGE = 17 * 1024;     // 132 rows x 128 B rounded up to a multiple of 1024
constexpr int NSTAGE = 2;               // input stages per CTA at 2 CTAs / SM
constexpr int SMALL_NSTAGE = 3;         // ... at 1 CTA / SM ("small footprint" mode, see launch_dw_persist)
constexpr int MAX_NSTAGE = 3;
constexpr int BQ = 64 * 128;           // one Toeplitz block: 64 rows (r) x 64 k (j) 16-bit = 8 KB
constexpr int MAX_NQ = 5;
constexpr int OUT_STAGE = MROWS * 128; // 16 KB
constexpr int ACC_STAGES = 4;
constexpr int TMEM_COLS = ACC_STAGES * L;  // 256
constexpr int THREADS = 256;
constexpr int BUILDERS = 64;
constexpr int WP_FLOATS = 64 * MAX_NQ + 80;   // zero-padded dilated tap line
__host__ __device__ constexpr int smem_bytes(int nq, int nbuf, int nstage) {
  return nstage * A_STAGE + nbuf * nq * BQ + OUT_STAGE + WP_FLOATS * 4 + 256 + 1024;
}

struct Params {
  CUtensorMap in, out;   // (64 frames, W windows, C, B), box (64, R, 1, NB)
  const float* w;
  const int32_t* lens;
  int B, C, T, K, P, D;
  int W, R, NB;
  int HL, NQ, nbuf;      // left halo windows; Toeplitz blocks; Toeplitz buffers (2, or 1 when two sets do not fit)
  int nstage;            // input stages (NSTAGE, or SMALL_NSTAGE with one CTA per SM)
  int tiles_per_chan;
  long long total;       // C * tiles_per_chan work items, channel-major
  int f16;
  int rev;
  unsigned long long* trace;   // ts_trace slot of this launch (nullptr: off)
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// K-major SWIZZLE_128B descriptor (8-row groups 1024 B apart); a tile whose start is shifted by q*128 B is addressed
// correctly with base offset 0 (measured on B200, see dwmma2.cu)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int NQ_T>
__global__ void __launch_bounds__(THREADS, 2)
dw_persist_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  const int NSTAGE = p.nstage;          // shadows the namespace constant: stage count of THIS launch
  uint8_t* sB = sA + NSTAGE * A_STAGE;
  uint8_t* sO = sB + p.nbuf * p.NQ * BQ;
  float* wp = reinterpret_cast<float*>(sO + OUT_STAGE);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wp + WP_FLOATS);
  uint64_t* empty_bar = full_bar + MAX_NSTAGE;
  uint64_t* acc_full = empty_bar + MAX_NSTAGE;
  uint64_t* acc_empty = acc_full + ACC_STAGES;
  uint64_t* b_ready = acc_empty + ACC_STAGES;
  uint64_t* b_free = b_ready + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_free + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tw = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 1 ? 1 : -1);
  if (tid == 0) {
    trace_head(p.trace, tw, 2);
    trace_stamp(p.trace, tw, 1);
  }
  // contiguous balanced share of the channel-major work list
  const long long g0 = p.total * blockIdx.x / gridDim.x, g1 = p.total * (blockIdx.x + 1) / gridDim.x;
  const int ntiles = (int)(g1 - g0);
  // n-th work item of this CTA -> (channel, utterance tile); reversed walks start at the far end of the tensor
  auto item = [&](int n, int& c, int& tl) {
    const long long g = p.rev ? p.total - 1 - (g0 + n) : g0 + n;
    c = (int)(g / p.tiles_per_chan);
    tl = (int)(g - (long long)c * p.tiles_per_chan);
  };
  const uint32_t box_bytes = (uint32_t)(128 * p.R * p.NB);
  // the FIRST channel's taps are on the critical path of the whole CTA (its Toeplitz blocks gate the first MMA): issue their
  // global loads before anything else so that their latency hides under the prologue (one tap per builder thread, K <= 192)
  float tap0 = 0.f;
  if (tid >= 64 && tid - 64 < p.K && ntiles > 0) {
    int c0, t0;
    item(0, c0, t0);
    tap0 = __ldg(p.w + (size_t)c0 * p.K + (tid - 64));
  }

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.in);
    ptx::prefetch_tensormap(&p.out);
    for (int s = 0; s < NSTAGE; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], 128);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&b_ready[i], BUILDERS);
      ptx::mbar_init(&b_free[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  // stage rows the TMA box never writes (>= NB*R) feed only accumulator rows that are never stored; zero them once
  for (int i = tid; i < NSTAGE * (A_STAGE / 16); i += THREADS) {
    const int s = i / (A_STAGE / 16), off = (i % (A_STAGE / 16)) * 16;
    if (off >= (int)box_bytes) *reinterpret_cast<uint4*>(sA + s * A_STAGE + off) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) trace_stamp(p.trace, tw, 2);

  pdl_launch_dependents();   // the next kernel may begin its prologue once every CTA of this grid is here
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      pdl_wait();              // first read of the previous kernel's output
      trace_stamp(p.trace, tw, 3);
      for (int n = 0; n < ntiles; ++n) {
        int c, tl;
        item(n, c, tl);
        const int s = n % NSTAGE;
        ptx::mbar_wait(&empty_bar[s], ((n / NSTAGE) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&full_bar[s], box_bytes);
        tma_load_4d(sA + s * A_STAGE, &p.in, &full_bar[s], 0, -p.HL, c, tl * p.NB);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_16(MROWS, L, 0, 0, p.f16);
      const int jlo = 64 * p.HL - p.P, jhi = 64 * p.HL + 63 + p.P;  // non-zero band of the stacked Toeplitz rows
      int cur_c = -1, j = -1;
      uint32_t sb = 0;
      for (int n = 0; n < ntiles; ++n) {
        int c, tl;
        item(n, c, tl);
        if (c != cur_c) {   // channel boundary: release the old Toeplitz set once its MMAs retire, wait for the new one
          if (cur_c >= 0) ptx::mma_commit(&b_free[j % p.nbuf]);
          cur_c = c;
          ++j;
          ptx::mbar_wait(&b_ready[j % p.nbuf], (j / p.nbuf) & 1);
          sb = ptx::smem_u32(sB + (j % p.nbuf) * p.NQ * BQ);
        }
        const int s = n % NSTAGE, a = n % ACC_STAGES;
        ptx::mbar_wait(&acc_empty[a], ((n / ACC_STAGES) & 1) ^ 1);
        ptx::mbar_wait(&full_bar[s], (n / NSTAGE) & 1);
        ptx::tc_fence_after();
        if (n == 0) trace_stamp(p.trace, tw, 4);
        const uint32_t sa = ptx::smem_u32(sA + s * A_STAGE);
        uint32_t acc = 0;
#pragma unroll
        for (int q = 0; q < NQ_T; ++q) {
          if (q >= p.NQ) break;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int j0 = 64 * q + 16 * ks;
            if (j0 + 15 < jlo || j0 > jhi) continue;
            ptx::mma_bf16_ss(tmem_base + a * L, desc_sw128(sa + q * 128 + ks * 32), desc_sw128(sb + q * BQ + ks * 32), idesc,
                             acc);
            acc = 1;
          }
        }
        ptx::mma_commit(&empty_bar[s]);
        ptx::mma_commit(&acc_full[a]);
      }
      trace_stamp(p.trace, tw, 5);
    }
  } else {
    // ===== Toeplitz builders: Tq[r][jj] = wline[64 q + jj - 64 HL - r + P], SW128 rows of 128 B =====
    // wline = the dilated tap line (tap k at position k * D) zero-padded on both sides, so the build needs neither bounds
    // checks nor divisions: wp[OFF + d] = wline[d], OFF = 64 HL + 63 - P >= 0, d in [-OFF, 64 (NQ - HL) + P + 6].
    // The FIRST channel's blocks are on the critical path of the whole CTA, so all six non-issuing warps build them
    // (the epilogue warps have nothing to do yet); afterwards warps 2-3 alone build one channel ahead of the tensor core.
    const int OFF = 64 * p.HL + 63 - p.P;
    const int span = 64 * p.NQ + 70;
    auto build = [&](int c, int buf, int bt, int nthr, int bar_id, bool first) {
      named_bar_sync(bar_id, nthr);   // every builder is done reading the previous tap line
      for (int i = bt; i < span; i += nthr) wp[i] = 0.f;
      named_bar_sync(bar_id, nthr);
      if (first && p.K <= THREADS - 64) {   // preloaded at kernel entry (bt == tid - 64 for the first build)
        if (bt < p.K) wp[OFF + bt * p.D] = tap0;
      } else {
        for (int k = bt; k < p.K; k += nthr) wp[OFF + k * p.D] = __ldg(p.w + (size_t)c * p.K + k);
      }
      named_bar_sync(bar_id, nthr);
      uint8_t* dst = sB + buf * p.NQ * BQ;
      for (int ci = bt; ci < p.NQ * 64 * 8; ci += nthr) {
        const int q = ci >> 9, r = (ci >> 3) & 63, g = ci & 7;
        const float* src = wp + (64 * q + 8 * g + 63 - r);   // = wp[OFF + d0], d0 = 64 q + 8 g - 64 HL - r + P
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) pk[h] = pack16x2(src[2 * h], src[2 * h + 1], p.f16 != 0);
        *reinterpret_cast<uint4*>(dst + q * BQ + r * 128 + ((g ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
    };
    int cur_c, tl0;
    item(0, cur_c, tl0);
    build(cur_c, 0, tid - 64, THREADS - 64, 3, true);
    named_bar_sync(3, THREADS - 64);   // all 192 threads' writes are fenced before the 64 arrivals complete the phase
    if (warp < 4) {
      ptx::mbar_arrive(&b_ready[0]);
      int j = 0;
      for (int n = 1; n < ntiles; ++n) {
        int c, tl;
        item(n, c, tl);
        if (c == cur_c) continue;
        cur_c = c;
        ++j;
        ptx::mbar_wait(&b_free[j % p.nbuf], ((j / p.nbuf) & 1) ^ 1);   // the MMAs that read this buffer have retired
        build(c, j % p.nbuf, tid - 64, BUILDERS, 2, false);
        ptx::mbar_arrive(&b_ready[j % p.nbuf]);
      }
    }
  }
  if (warp >= 4) {
    // ===== epilogue: TMEM -> 16-bit -> swizzled staging -> one TMA store per tile =====
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int bl = row / p.R, i = row - bl * p.R;
    const int t = 64 * i;
    const uint32_t srow = ptx::smem_u32(sO) + row * 128;
    pdl_wait();                // no global write of this grid may overtake the previous grid's reads
    for (int n = 0; n < ntiles; ++n) {
      int c, tl;
      item(n, c, tl);
      const int a = n % ACC_STAGES;
      const int b0 = tl * p.NB;
      const int b = b0 + bl;
      int lout = p.T;
      if (p.lens && bl < p.NB && b < p.B) lout = min(lout, max(__ldg(p.lens + b), 0));
      ptx::mbar_wait(&acc_full[a], (n / ACC_STAGES) & 1);
      ptx::tc_fence_after();
      if (n == 0 && tid == 128) trace_stamp(p.trace, tw, 6);
      uint32_t v[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * L);
      ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc_empty[a]);
      // the previous tile's TMA store must have finished READING the staging buffer
      if (tid == 128) bulk_wait_read0();
      named_bar_sync(1, 128);
      const bool interior = t + 64 <= lout;
      auto pack_and_stage = [&](auto f16_tag) {   // uniform branch on the row format: one conversion per pair on each path
        constexpr bool kF16 = decltype(f16_tag)::value;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t pk[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int e = g * 8 + 2 * h;
            float lo = __uint_as_float(v[e]), hi = __uint_as_float(v[e + 1]);
            if (!interior) {
              if (t + e >= lout) lo = 0.f;
              if (t + e + 1 >= lout) hi = 0.f;
            }
            pk[h] = kF16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((g ^ (row & 7)) << 4)), "r"(pk[0]),
                       "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                       : "memory");
        }
      };
      if (p.f16)
        pack_and_stage(std::true_type{});
      else
        pack_and_stage(std::false_type{});
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (tid == 128) {
        tma_store_4d(&p.out, sO, 0, 0, c, b0);
        bulk_commit();
      }
    }
    if (tid == 128) {
      trace_stamp(p.trace, tw, 7);
      bulk_wait0();
      trace_stamp(p.trace, tw, 8);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (tid == 0) trace_stamp(p.trace, tw, 9);
}

}  // namespace dwt3

int option_dw_share_halo();
int small_footprint(long long frames);
int option_dwp_nbuf();
int option_dwp_nstage();

int launch_dw_persist(const __nv_bfloat16* x, int B, int C, int T, int pitch_in, const float* w, int K, int D, int P,
                      const int32_t* lens, __nv_bfloat16* y, int pitch_out, cudaStream_t st, int f16) {
  // stride 1, length preserving ("same") padding: 2 P == D (K - 1)
  if (!(2 * P == D * (K - 1) && D >= 1 && pitch_in == pitch_out && pitch_in % 64 == 0 && C <= 65535))
    return TS_ERR_UNSUPPORTED;
  dwt3::Params p;
  memset(&p, 0, sizeof(p));
  p.w = w; p.lens = lens; p.f16 = f16;
  p.B = B; p.C = C; p.T = T; p.K = K; p.P = P; p.D = D;
  p.W = pitch_in / 64;
  p.HL = ceil_div(P, 64);
  p.NQ = p.HL + 1 + (63 + P) / 64;
  if (p.NQ > dwt3::MAX_NQ || P == 0 || (K - 1) * D + 64 * p.HL + 63 - P >= dwt3::WP_FLOATS) return TS_ERR_UNSUPPORTED;
  // rows per utterance in the stacked tile: W data windows + the larger halo (all-zero halo rows are shared between
  // neighbouring utterances, see dwmma2.cu)
  const int HR = p.NQ - 1 - p.HL;
  p.R = option_dw_share_halo() ? p.W + (p.HL > HR ? p.HL : HR) : p.W + p.NQ - 1;
  if (p.R > dwt3::MROWS) return TS_ERR_UNSUPPORTED;
  p.NB = dwt3::MROWS / p.R;
  if (p.NB > 256 || p.R > 256) return TS_ERR_UNSUPPORTED;
  p.tiles_per_chan = ceil_div(B, p.NB);
  p.total = (long long)C * p.tiles_per_chan;
  p.nbuf = (p.NQ <= 3) ? 2 : 1;     // two Toeplitz sets of 5 blocks would leave room for one CTA per SM only
  if (option_dwp_nbuf() > 0 && p.NQ <= 3) p.nbuf = option_dwp_nbuf() >= 2 ? 2 : 1;
  int rc;
  cuuint64_t dims[4] = {64, (cuuint64_t)p.W, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[3] = {128, (cuuint64_t)pitch_in * 2, (cuuint64_t)C * pitch_in * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)p.R, 1, (cuuint32_t)p.NB};
  if ((rc = tma::encode(&p.in, x, 4, dims, strides, box)) != TS_OK) return rc;
  if ((rc = tma::encode(&p.out, y, 4, dims, strides, box)) != TS_OK) return rc;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    TS_CUDA(cudaGetDevice(&dev));
    TS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    TS_CUDA(cudaFuncSetAttribute(dwt3::dw_persist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dwt3::smem_bytes(3, 2, dwt3::MAX_NSTAGE)));
    TS_CUDA(cudaFuncSetAttribute(dwt3::dw_persist_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dwt3::smem_bytes(dwt3::MAX_NQ, 1, dwt3::MAX_NSTAGE)));
  }
  p.rev = next_walk_reversed();
  // Small-footprint mode (few tiles per CTA): ONE CTA per SM with three input stages (117 KB, 256 TMEM columns), the
  // counterpart of the slim pair GEMM (pwgemm3.cu): a Toeplitz CTA and a GEMM CTA fit on an SM together, so every launch
  // of the dw -> pw -> dw chain is resident -- prologue done, first Toeplitz blocks built -- while its predecessor still
  // computes, and starts streaming the moment the dependency resolves.
  const bool small = p.NQ <= 3 && small_footprint((long long)B * pitch_in) != 0;
  p.nstage = small ? dwt3::SMALL_NSTAGE : dwt3::NSTAGE;
  if (option_dwp_nstage() > 0 && !small) {   // experiment: as many stages as two CTAs per SM allow
    p.nstage = option_dwp_nstage() > dwt3::MAX_NSTAGE ? dwt3::MAX_NSTAGE : option_dwp_nstage();
    while (p.nstage > 2 && 2 * (dwt3::smem_bytes(p.NQ, p.nbuf, p.nstage) + 1024) > 232448) --p.nstage;
  }
  long long grid = (small ? 1ll : 2ll) * num_sms;
  if (grid > p.total) grid = p.total;
  const int smem = dwt3::smem_bytes(p.NQ, p.nbuf, p.nstage);
  p.trace = trace_next_slot(2, (unsigned)grid);
  if (p.NQ <= 3)
    TS_CUDA(launch_pdl(dwt3::dw_persist_kernel<3>, dim3((unsigned)grid), dim3(dwt3::THREADS), smem, st, option_pdl() != 0, p));
  else
    TS_CUDA(launch_pdl(dwt3::dw_persist_kernel<5>, dim3((unsigned)grid), dim3(dwt3::THREADS), smem, st, option_pdl() != 0, p));
  TS_LAUNCH_CHECK("dw_persist_kernel");
  return TS_OK;
}

}  // namespace ts
