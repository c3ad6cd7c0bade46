// Toeplitz-MMA depthwise conv, TMA edition (see dwmma.cu for the math).
//
// Differences to dwmma.cu: the input tile ([NB utterances] x [W+2 windows] x 64 frames of ONE channel) arrives with a
// single 4-D TMA box per tile -- out-of-range windows (the left/right zero padding of the conv) and utterances
// beyond the batch are zero-filled by the TMA unit -- into SWIZZLE_128B rows of 128 bytes (one 64-frame window per
// row), which is also the K-major operand layout of tcgen05.mma; the three row-shifted A tiles are the same staged
// copy addressed +128*q bytes.  The output tile leaves through a swizzled staging buffer and ONE 4-D TMA store per
// tile (halo windows and tail utterances are clipped by the tensor map), so every global transaction is a full
// 128-byte row.  The caller guarantees that the input rows are already zero beyond each utterance's length (true for
// every producer in this library); only the OUTPUT tail mask needs `lens`.
#include "ts_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"
#include <type_traits>

namespace ts {
namespace dwt2 {

constexpr int L = 64;
constexpr int MROWS = 128;
constexpr int A_STAGE = 17 * 1024;     // 130 rows x 128 B rounded up to a multiple of 1024
constexpr int NSTAGE = 4;              // default input stages (option dw_nstage: 2..4), clamped to what two CTAs per SM allow:
                                       // 4 x 17 KB + 24 KB Toeplitz + 16 KB staging = 109 KB per CTA.  Alternating A/B on one box,
                                       // 3 vs 4 stages: QuartzNet 256 x 15 s 11.24 -> 11.16 ms (three pairs of runs, all the same
                                       // sign), Citrinet-1024 128 x 20 s 25.02 -> 24.86 ms
constexpr int MAX_NSTAGE = 4;
constexpr int BQ = 64 * 128;           // one Toeplitz block: 64 rows (r) x 64 k (j) bf16 = 8 KB
constexpr int MAX_NQ = 5;              // Toeplitz blocks: left halo + centre + right halo windows (3 for K <= 129, dilation 1)
constexpr int OUT_STAGE = MROWS * 128; // 16 KB
constexpr int ACC_STAGES = 4;
constexpr int TMEM_COLS = ACC_STAGES * L;  // 256
constexpr int THREADS = 256;
__host__ __device__ constexpr int smem_bytes(int nq, int nstage) { return nstage * A_STAGE + nq * BQ + OUT_STAGE + 256 + 1024; }

struct Params {
  CUtensorMap in, out;   // (64 frames, W windows, C, B), box (64, R, 1, NB)
  const float* w;
  const int32_t* lens;
  int B, C, T, K, P, D;
  int W, R, NB;
  int HL, NQ;            // left halo windows = ceil(P / 64); Toeplitz blocks = HL + 1 + (63 + P) / 64
  int nstage;            // input stages of this launch
  int tiles_per_chan, tiles_per_cta;
  int f16;               // rows (and the Toeplitz blocks) are IEEE fp16 instead of bf16
  int rev;               // walk the utterance tiles from the far end (see next_walk_reversed())
  unsigned long long* trace;   // ts_trace slot of this launch (nullptr: off)
  int dbg;               // timing experiments (results are wrong): 1 = one MMA per tile, 4 = no MMA, 2 = no staging / store
  int base_offset_mode;  // experiment switch, default 0: MEASURED on B200 -- a SW128 tile whose start is shifted by
                         // q*128 B is addressed correctly with base-offset 0 (the XOR uses absolute address bits);
                         // writing q into bits 49..51 gives wrong results
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

// K-major SWIZZLE_128B descriptor (8-row groups 1024 B apart) with an explicit base-offset field (bits 49..51)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int NQ_T>   // compile-time bound of the Toeplitz block loop (3: K <= 129 undilated; 5: the general case)
__global__ void __launch_bounds__(THREADS, 2)
dw_tma_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  const int NSTAGE = p.nstage;           // shadows the namespace default: stage count of THIS launch
  uint8_t* sB = sA + NSTAGE * A_STAGE;
  uint8_t* sO = sB + p.NQ * BQ;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sO + OUT_STAGE);
  uint64_t* empty_bar = full_bar + MAX_NSTAGE;
  uint64_t* acc_full = empty_bar + MAX_NSTAGE;
  uint64_t* acc_empty = acc_full + ACC_STAGES;
  uint64_t* b_ready = acc_empty + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x;
  const int tw = (blockIdx.x == 0 && blockIdx.y == 0) ? 0 : -1;
  if (tid == 0) {
    trace_head(p.trace, tw, 3);
    trace_stamp(p.trace, tw, 1);
    trace_cta(p.trace, p.dbg, blockIdx.y * gridDim.x + blockIdx.x, 0);
  }
  const int tile0 = blockIdx.y * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.tiles_per_chan - tile0);
  // n-th tile of this CTA; reversed walks start at the last utterances (CTAs with a low blockIdx.y are scheduled first)
  auto tile_at = [&](int n) { return p.rev ? p.tiles_per_chan - 1 - (tile0 + n) : tile0 + n; };
  const uint32_t box_bytes = (uint32_t)(128 * p.R * p.NB);
  // this channel's taps gate the Toeplitz build and with it the first MMA: their global loads are issued before anything
  // else so that the latency hides under the prologue (one tap per builder thread; K <= 192)
  float tap0 = 0.f;
  if (tid >= 64 && tid - 64 < p.K) tap0 = __ldg(p.w + (size_t)c * p.K + (tid - 64));

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
    ptx::mbar_init(b_ready, THREADS - 64);
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
    // ===== TMA producer (starts streaming immediately; the Toeplitz build overlaps) =====
    if (lane == 0) {
      pdl_wait();              // first read of the previous kernel's output
      trace_stamp(p.trace, tw, 3);
      long long w12 = 0;
      for (int n = 0; n < ntiles; ++n) {
        const int s = n % NSTAGE;
        TS_TIMED_WAIT((p.trace != nullptr && tw == 0), w12, ptx::mbar_wait(&empty_bar[s], ((n / NSTAGE) & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(&full_bar[s], box_bytes);
        tma_load_4d(sA + s * A_STAGE, &p.in, &full_bar[s], 0, -p.HL, c, tile_at(n) * p.NB);
      }
      trace_put(p.trace, tw, 12, w12);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_16(MROWS, L, 0, 0, p.f16);
      const uint32_t sb = ptx::smem_u32(sB);
      const int jlo = 64 * p.HL - p.P, jhi = 64 * p.HL + 63 + p.P;  // non-zero band of the stacked Toeplitz rows
      long long w10 = 0, w11 = 0;
      const long long t_loop0 = trace_clock();
      ptx::mbar_wait(b_ready, 0);
      for (int n = 0; n < ntiles; ++n) {
        const int s = n % NSTAGE, a = n % ACC_STAGES;
        TS_TIMED_WAIT((p.trace != nullptr && tw == 0), w11, ptx::mbar_wait(&acc_empty[a], ((n / ACC_STAGES) & 1) ^ 1));
        TS_TIMED_WAIT((p.trace != nullptr && tw == 0), w10, ptx::mbar_wait(&full_bar[s], (n / NSTAGE) & 1));
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
            const uint64_t da = desc_sw128(sa + q * 128 + ks * 32, p.base_offset_mode ? q : 0);
            const uint64_t db = desc_sw128(sb + q * BQ + ks * 32, 0);
            ptx::mma_bf16_ss(tmem_base + a * L, da, db, idesc, acc);
            acc = 1;
          }
        }
        ptx::mma_commit(&empty_bar[s]);
        ptx::mma_commit(&acc_full[a]);
      }
      trace_put(p.trace, tw, 10, w10);
      trace_put(p.trace, tw, 11, w11);
      trace_put(p.trace, tw, 15, trace_clock() - t_loop0);
    }
  }
  if (warp >= 2) {
    // ===== Toeplitz blocks of this channel: Tq[r][j] = w[(64 q + j - 64 HL - r + P) / D], SW128 rows of 128 B =====
    // the K taps of this channel are staged once in shared memory (borrowing the output staging buffer, which the epilogue
    // first touches only after b_ready): the build below then needs no global loads
    float* wc = reinterpret_cast<float*>(sO);
    if (p.K <= THREADS - 64) {
      if (tid - 64 < p.K) wc[tid - 64] = tap0;      // fetched at kernel entry
    } else {
      for (int k = tid - 64; k < p.K; k += THREADS - 64) wc[k] = __ldg(p.w + (size_t)c * p.K + k);
    }
    named_bar_sync(2, THREADS - 64);
    const int kd = p.K * p.D;
    for (int ci = tid - 64; ci < p.NQ * 64 * 8; ci += THREADS - 64) {
      const int q = ci >> 9, r = (ci >> 3) & 63, g = ci & 7;
      const int d0 = 64 * q + 8 * g - 64 * p.HL - r + p.P;
      uint32_t pk[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int da = d0 + 2 * h, db = da + 1;
        float fa, fb;
        if (p.D == 1) {  // uniform branch: no integer division on the common path
          fa = (da >= 0 && da < kd) ? wc[da] : 0.f;
          fb = (db >= 0 && db < kd) ? wc[db] : 0.f;
        } else {
          fa = (da >= 0 && da < kd && da % p.D == 0) ? wc[da / p.D] : 0.f;
          fb = (db >= 0 && db < kd && db % p.D == 0) ? wc[db / p.D] : 0.f;
        }
        pk[h] = pack16x2(fa, fb, p.f16 != 0);
      }
      *reinterpret_cast<uint4*>(sB + q * BQ + r * 128 + ((g ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    fence_proxy_async();
    ptx::mbar_arrive(b_ready);
  }
  if (warp >= 4) {
    // ===== epilogue: TMEM -> bf16 -> swizzled staging -> one TMA store per tile =====
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int bl = row / p.R, i = row - bl * p.R;
    const int t = 64 * i;
    const uint32_t srow = ptx::smem_u32(sO) + row * 128;
    pdl_wait();                // no global write of this grid may overtake the previous grid's reads
    long long w13 = 0, w14 = 0;
    for (int n = 0; n < ntiles; ++n) {
      const int a = n % ACC_STAGES;
      const int b0 = tile_at(n) * p.NB;
      const int b = b0 + bl;
      int lout = p.T;
      if (p.lens && bl < p.NB && b < p.B) lout = min(lout, max(__ldg(p.lens + b), 0));
      TS_TIMED_WAIT((p.trace != nullptr && tw == 0 && tid == 128), w13, ptx::mbar_wait(&acc_full[a], (n / ACC_STAGES) & 1));
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
      if (tid == 128) TS_TIMED_WAIT((p.trace != nullptr && tw == 0), w14, bulk_wait_read0());
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
      trace_put(p.trace, tw, 13, w13);
      trace_put(p.trace, tw, 14, w14);
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
  if (tid == 0) {
    trace_stamp(p.trace, tw, 9);
    trace_cta(p.trace, p.dbg, blockIdx.y * gridDim.x + blockIdx.x, 1);
  }
}

}  // namespace dwt2

int option_dw_base_offset();
int option_dw_share_halo();
int option_dw_pro();
int option_dw_nstage();

int launch_dw_tma(const __nv_bfloat16* x, int B, int C, int T, int pitch_in, const float* w, int K, int D, int P,
                  const int32_t* lens, __nv_bfloat16* y, int pitch_out, cudaStream_t st, int f16) {
  // stride 1, length preserving ("same") padding: 2 P == D (K - 1)
  if (!(2 * P == D * (K - 1) && D >= 1 && pitch_in == pitch_out && pitch_in % 64 == 0 && C <= 65535))
    return TS_ERR_UNSUPPORTED;
  dwt2::Params p;
  memset(&p, 0, sizeof(p));
  p.rev = next_walk_reversed();
  p.w = w; p.lens = lens; p.f16 = f16;
  p.B = B; p.C = C; p.T = T; p.K = K; p.P = P; p.D = D;
  p.W = pitch_in / 64;
  p.HL = ceil_div(P, 64);
  p.NQ = p.HL + 1 + (63 + P) / 64;
  if (p.NQ > dwt2::MAX_NQ || P == 0) return TS_ERR_UNSUPPORTED;
  // Rows per utterance in the stacked tile.  Both halos of an utterance are all-zero windows (conv padding on the left,
  // the zero pad beyond the pitch on the right), so neighbouring utterances SHARE them: HL zero rows in front of every
  // utterance double as the right halo of the previous one (the rows behind the last utterance are zeroed once at kernel
  // start).  W + max(HL, HR) rows instead of W + HL + HR: 25 instead of 21 utterances per tile at T = 251.
  const int HR = p.NQ - 1 - p.HL;
  p.R = option_dw_share_halo() ? p.W + (p.HL > HR ? p.HL : HR) : p.W + p.NQ - 1;
  if (p.R > dwt2::MROWS) return TS_ERR_UNSUPPORTED;
  p.NB = dwt2::MROWS / p.R;
  if (p.NB > 256 || p.R > 256) return TS_ERR_UNSUPPORTED;
  p.tiles_per_chan = ceil_div(B, p.NB);
  {
    double best = -1.0;
    int best_tpc = p.tiles_per_chan;
    for (int tpc = 1; tpc <= p.tiles_per_chan; ++tpc) {
      const int g = ceil_div(p.tiles_per_chan, tpc);
      const double waves = (double)C * g / 296.0;
      const double pro = option_dw_pro() / 10.0;   // per-CTA prologue (TMEM alloc, Toeplitz build, first load) in tile-times
      const double eff = waves / (double)((long long)(waves + 0.999999)) * (double)p.tiles_per_chan / (double)(g * tpc) *
                         (double)tpc / ((double)tpc + pro);
      if (eff > best + 1e-9) {
        best = eff;
        best_tpc = tpc;
      }
    }
    p.tiles_per_cta = best_tpc;
  }
  const int groups = ceil_div(p.tiles_per_chan, p.tiles_per_cta);
  p.base_offset_mode = option_dw_base_offset();
  p.dbg = option_dbg();
  int rc;
  cuuint64_t dims[4] = {64, (cuuint64_t)p.W, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[3] = {128, (cuuint64_t)pitch_in * 2, (cuuint64_t)C * pitch_in * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)p.R, 1, (cuuint32_t)p.NB};
  if ((rc = tma::encode(&p.in, x, 4, dims, strides, box)) != TS_OK) return rc;
  if ((rc = tma::encode(&p.out, y, 4, dims, strides, box)) != TS_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    TS_CUDA(cudaFuncSetAttribute(dwt2::dw_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dwt2::smem_bytes(3, dwt2::MAX_NSTAGE)));
    TS_CUDA(cudaFuncSetAttribute(dwt2::dw_tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 dwt2::smem_bytes(dwt2::MAX_NQ, dwt2::MAX_NSTAGE)));
    attr_set = true;
  }
  dim3 grid(C, groups);
  p.nstage = option_dw_nstage() > 0 ? option_dw_nstage() : dwt2::NSTAGE;
  if (p.nstage > dwt2::MAX_NSTAGE) p.nstage = dwt2::MAX_NSTAGE;
  while (p.nstage > 2 && 2 * (dwt2::smem_bytes(p.NQ, p.nstage) + 1024) > 232448) --p.nstage;   // two CTAs per SM
  p.trace = trace_next_slot(3, (unsigned)(C * groups));
  if (p.NQ <= 3)
    TS_CUDA(launch_pdl(dwt2::dw_tma_kernel<3>, grid, dim3(dwt2::THREADS), dwt2::smem_bytes(p.NQ, p.nstage), st, option_pdl() != 0, p));
  else
    TS_CUDA(launch_pdl(dwt2::dw_tma_kernel<5>, grid, dim3(dwt2::THREADS), dwt2::smem_bytes(p.NQ, p.nstage), st, option_pdl() != 0, p));
  TS_LAUNCH_CHECK("dw_tma_kernel");
  return TS_OK;
}

}  // namespace ts
