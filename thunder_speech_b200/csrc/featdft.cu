// Log-mel front-end, tensor-core edition: the STFT as a DFT-matrix contraction on tcgen05 (north_star kernel 1, variant B;
// variant A, the shared-memory radix-8 FFT, is features.cu -- `ts_logmel` picks one, see the measurements in DESIGN.md).
//
// Math (replaces PreEmphasisFilter / PowerSpectrum / MelScale, src/thunder/quartznet/transform.py:136-144,186-208,243-255;
// the DFT-matrix STFT itself is the reference's own second implementation, src/thunder/blocks.py:38-91).  With n_fft = 512
// and the window supported on [96, 416) the frame is symmetric about n = 255.5, so with y = window * pre-emphasised audio,
//     y+[m] = y[256 + m],  y-[m] = y[255 - m],  ye = y+ + y-,  yo = y+ - y-,   m = 0..159,
//     |X[k]|^2 = (sum_m ye[m] cos(th))^2 + (sum_m yo[m] sin(th))^2,   th = 2 pi k (m + 1/2) / 512,
// i.e. two real GEMMs  [frames x 160] x [160 x 256 bins]  (half the flops of the direct DFT matrix).  Bin 256 rides in the
// unused column 0 of the sine GEMM (sin(0) = 0 there; cos(pi (m + 1/2)) = 0 and sin(pi (m + 1/2)) = (-1)^m at k = 256).
// In frame coordinates y+[m] and y-[m] are the samples hop*f + m and hop*f - 1 - m of the signal: the A operand is the
// signal re-read in rows of `hop`, no per-frame gather.
//
// Precision: fp16 tensor-core operands cannot hold the 1e-4 feature parity (11-bit mantissa), so both operands are split
// hi + lo (two fp16 each, ~22 bits) and every k-step issues  hi*hi + lo*hi + hi*lo  (the lo*lo term is 2^-22 relative).
// The audio side is scaled by 2^8 before the split so that quiet signals stay out of fp16's subnormals; power is unscaled
// by 2^-16 in the epilogue.  (|window * audio| must stay below 255: float audio in [-1, 1] like torchaudio.load.)
//
// One CLUSTER of two CTAs per 256 frames (tcgen05 cta_group::2, M = 256): each CTA owns 128 frames (its TMEM lanes) and
// keeps HALF of the split basis resident in shared memory for the whole kernel (4 matrices x 160 x 128 bins fp16 = 160 KB;
// the pair's MMA reads both halves), the audio-side slices stream through a 3-stage ring built by SIMT warps.
//
//   warp 0            MMA issuer (leader CTA): per k-step 6 x M256 N256 K16 -> TMEM cols [0,256) cosine, [256,512) sine
//   warps 4-7         epilogue, thread = frame: TMEM -> power -> sliding two-filter mel accumulation over the ordered bins
//                     -> log -> coalesced store; per-(32 frames, filter) sum / sum-of-squares partials for the normaliser
//   warps 8-15        builders: coalesced signal loads (16 frames x 16 m per warp step) -> pre-emphasis, window, fold, split
//                     -> no-swizzle K-major core-matrix slices
#include "ts_common.cuh"
#include "sm100_ptx.cuh"

namespace ts {
namespace fdft {

constexpr int NFFT = 512;
constexpr int FR = 128;                 // frames per CTA (TMEM lanes)
constexpr int KM = 160;                 // folded half window
constexpr int KSTEPS = KM / 16;
constexpr int SLICE = 128 * 16 * 2;     // one [128 rows x 16 k] fp16 slice in core-matrix layout: 4 KB
constexpr int B_BYTES = 4 * KSTEPS * SLICE;   // C_hi, C_lo, S_hi, S_lo halves: 160 KB
constexpr int A_STAGE = 4 * SLICE;      // ye_hi, ye_lo, yo_hi, yo_lo: 16 KB
constexpr int NSTA = 3;
constexpr int NBINS = 257;
constexpr int MAX_NFILT = 128;
constexpr int THREADS = 512;
constexpr int BUILD_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr float A_SCALE = 256.f, P_UNSCALE = 1.0f / 65536.f;
// shared memory: basis | A ring | mel tables (w0/w1 float2 [257], adv int [257]) | window taps 2 x 160 f32 | barriers
constexpr int TAB_BYTES = NBINS * 8 + ((NBINS * 4 + 15) & ~15) + 2 * KM * 4;
constexpr int SMEM_BYTES = B_BYTES + NSTA * A_STAGE + TAB_BYTES + 256 + 1024;

struct Params {
  const float* audio;
  int B, N, F, hop;
  float preemph;
  const float* wplus;        // [160] window[256 + m]
  const float* wminus;       // [160] window[255 - m]
  const uint4* basis;        // [2 ranks][B_BYTES / 16]: this rank's half of the split basis, already in smem layout
  const float2* mel_w;       // [257] weights of bin k for the current / next filter of the sliding window
  const int32_t* mel_adv;    // [257] filters completed BEFORE bin k is accumulated
  int nfilt;
  float* logmel;             // [B, nfilt, F]
  float* partials;           // optional [B, ceil(F/32), nfilt, 2]: sum / sum of squares over the valid frames of 32
  const int64_t* lengths;    // optional [B] audio samples (valid frames: f < len / hop + 1)
  int tiles_per_row, num_tiles, pslots;
  int dbg;                   // timing experiments: 1 = epilogue skips the mel walk, 2 = builders skip loads, 4 = no MMAs
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {   // release at cluster scope
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(ptx::smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}" ::"r"(ptx::smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mma2_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
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
// K-major, no swizzle: 8 x 16 B core matrices; LBO = distance between the two k-cores of a K16 step, SBO = distance
// between 8-row groups (cute::UMMA canonical INTERLEAVE layout: ((8,m),(T,2)):((1T,SBO),(1,LBO)))
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;    // LBO = 128 B
  d |= (uint64_t)((256u >> 4) & 0x3FFF) << 32;    // SBO = 256 B
  d |= (uint64_t)1 << 46;
  return d;
}

// pre-emphasised, reflect-padded signal sample i of one utterance (torch.stft(center=True) pads the PRE-EMPHASISED row;
// PreEmphasisFilter keeps y[0] = x[0], transform.py:136-144)
__device__ __forceinline__ float sig(const float* __restrict__ x, int i, int N, float pre) {
  int r = i < 0 ? -i : (i >= N ? 2 * (N - 1) - i : i);
  r = min(max(r, 0), N - 1);
  return r >= 1 ? fmaf(-pre, __ldg(x + r - 1), __ldg(x + r)) : __ldg(x);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
logmel_dft_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sBasis = smem;
  uint8_t* sA = sBasis + B_BYTES;
  float2* s_melw = reinterpret_cast<float2*>(sA + NSTA * A_STAGE);
  int* s_adv = reinterpret_cast<int*>(s_melw + NBINS);
  float* s_wp = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_adv) + ((NBINS * 4 + 15) & ~15));
  float* s_wm = s_wp + KM;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_wm + KM);   // leader's copy is used
  uint64_t* a_empty = a_full + NSTA;
  uint64_t* d_full = a_empty + NSTA;
  uint64_t* d_empty = d_full + 1;                               // leader's copy is used
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (tid == 0) {
    for (int s = 0; s < NSTA; ++s) {
      ptx::mbar_init(&a_full[s], 2 * BUILD_WARPS);   // one arrival per builder warp of both CTAs
      ptx::mbar_init(&a_empty[s], 1);
    }
    ptx::mbar_init(d_full, 1);
    ptx::mbar_init(d_empty, 2 * 128);                // epilogue threads of both CTAs
    ptx::fence_barrier_init();
  }
  // resident operands: this rank's half of the split basis (already in core-matrix layout), mel tables, window taps
  {
    const uint4* src = p.basis + (size_t)rank * (B_BYTES / 16);
    uint4* dst = reinterpret_cast<uint4*>(sBasis);
    for (int i = tid; i < B_BYTES / 16; i += THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < NBINS; i += THREADS) {
      s_melw[i] = p.mel_w[i];
      s_adv[i] = p.mel_adv[i];
    }
    for (int i = tid; i < KM; i += THREADS) {
      s_wp[i] = p.wplus[i] * A_SCALE;
      s_wm[i] = p.wminus[i] * A_SCALE;
    }
  }
  fence_proxy_async();
  cluster_sync_all();
  if (warp == 1) {
    tmem2_alloc(tmem_slot, TMEM_COLS);
    tmem2_relinquish();
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== MMA issuer (leader CTA) =====
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = ptx::umma_idesc_16(256, 256, 0, 0, 1);
      const uint32_t sb = ptx::smem_u32(sBasis), sa0 = ptx::smem_u32(sA);
      uint32_t cnt = 0, it = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs, ++it) {
        mbar_wait_cluster(d_empty, (it & 1) ^ 1);     // the epilogues of both CTAs have drained the accumulators
        ptx::tc_fence_after();
        for (int ks = 0; ks < KSTEPS; ++ks, ++cnt) {
          const int s = cnt % NSTA;
          mbar_wait_cluster(&a_full[s], (cnt / NSTA) & 1);
          ptx::tc_fence_after();
          const uint32_t a = sa0 + s * A_STAGE;
          const uint64_t ye_hi = desc_nosw(a), ye_lo = desc_nosw(a + SLICE), yo_hi = desc_nosw(a + 2 * SLICE),
                         yo_lo = desc_nosw(a + 3 * SLICE);
          const uint32_t b = sb + ks * SLICE;
          const uint64_t c_hi = desc_nosw(b), c_lo = desc_nosw(b + KSTEPS * SLICE), s_hi = desc_nosw(b + 2 * KSTEPS * SLICE),
                         s_lo = desc_nosw(b + 3 * KSTEPS * SLICE);
          const uint32_t acc = ks > 0 ? 1u : 0u;
          if (!(p.dbg & 4)) {
          mma2_f16_ss(tmem_base, ye_hi, c_hi, idesc, acc);
          mma2_f16_ss(tmem_base, ye_lo, c_hi, idesc, 1u);
          mma2_f16_ss(tmem_base, ye_hi, c_lo, idesc, 1u);
          mma2_f16_ss(tmem_base + 256, yo_hi, s_hi, idesc, acc);
          mma2_f16_ss(tmem_base + 256, yo_lo, s_hi, idesc, 1u);
          mma2_f16_ss(tmem_base + 256, yo_hi, s_lo, idesc, 1u);
          }
          mma2_commit_mcast(&a_empty[s]);
        }
        mma2_commit_mcast(d_full);
      }
    }
  } else if (warp >= 8) {
    // ===== builders: one warp step = 16 frames x 16 m (one k-step): lane = (frame j = lane & 15, k-core kc = lane >> 4) =====
    const int bw = warp - 8;                  // this warp owns frame rows [16 bw, 16 bw + 16)
    const int j = lane & 15, kc = lane >> 4;
    const int r = 16 * bw + j;                // row (frame) within the CTA's 128
    const uint32_t row_off = (uint32_t)((r >> 3) * 256 + kc * 128 + (r & 7) * 16);
    const bool vec_ok = (p.N % 4 == 0) && (p.hop % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.audio) & 15) == 0);
    uint32_t cnt = 0;
    for (int tile = pair; tile < p.num_tiles; tile += npairs) {
      const int b = tile / p.tiles_per_row, f0 = (tile - b * p.tiles_per_row) * (2 * FR) + (int)rank * FR;
      const int f = f0 + r;
      const float* x = p.audio + (size_t)b * p.N;
      const long long base = (long long)f * p.hop;       // y+[m] = sig(base + m), y-[m] = sig(base - 1 - m)
      const bool live = f < p.F;
      // fast path: every sample this row touches, and its predecessor, is inside the utterance
      const bool interior = live && base - KM - 1 >= 0 && base + KM < p.N;
      for (int ks = 0; ks < KSTEPS; ++ks, ++cnt) {
        const int s = cnt % NSTA;
        const int m0 = 16 * ks + 8 * kc;
        float yp[8], ym[8];
        if (p.dbg & 2) {
#pragma unroll
          for (int e = 0; e < 8; ++e) yp[e] = ym[e] = 1.f;
        } else if (interior && vec_ok) {       // 16-byte loads: base, m0 and the row start are multiples of 4 samples
          const float* xp = x + base + m0;
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(xp)), v1 = __ldg(reinterpret_cast<const float4*>(xp + 4));
          const float fw[9] = {__ldg(xp - 1), v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) yp[e] = fmaf(-p.preemph, fw[e], fw[e + 1]);
          const float* xq = x + base - m0 - 8;           // bk[t] = x[base - m0 - 9 + t], t = 0..8
          const float4 u0 = __ldg(reinterpret_cast<const float4*>(xq)), u1 = __ldg(reinterpret_cast<const float4*>(xq + 4));
          const float bk[9] = {__ldg(xq - 1), u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) ym[e] = fmaf(-p.preemph, bk[7 - e], bk[8 - e]);   // y-[m0 + e] = g[base - 1 - m0 - e]
        } else if (interior) {
          const float* xp = x + base + m0;               // forward run  x[base + m0 - 1 .. base + m0 + 7]
          const float* xm = x + base - 1 - m0;           // backward run x[base - 1 - m0 - 8 .. base - 1 - m0]
          float prev = __ldg(xp - 1);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float cur = __ldg(xp + e);
            yp[e] = fmaf(-p.preemph, prev, cur);
            prev = cur;
          }
          float nxt = __ldg(xm);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float below = __ldg(xm - e - 1);
            ym[e] = fmaf(-p.preemph, below, nxt);
            nxt = below;
          }
        } else if (live) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            yp[e] = sig(x, (int)(base + m0 + e), p.N, p.preemph);
            ym[e] = sig(x, (int)(base - 1 - m0 - e), p.N, p.preemph);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) yp[e] = ym[e] = 0.f;
        }
        uint32_t eh[4] = {0, 0, 0, 0}, el[4] = {0, 0, 0, 0}, oh[4] = {0, 0, 0, 0}, ol[4] = {0, 0, 0, 0};
        if (!(p.dbg & 8)) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float a0 = s_wp[m0 + e] * yp[e], a1 = s_wp[m0 + e + 1] * yp[e + 1];
          float b0 = s_wm[m0 + e] * ym[e], b1 = s_wm[m0 + e + 1] * ym[e + 1];
          const float ye0 = a0 + b0, ye1 = a1 + b1, yo0 = a0 - b0, yo1 = a1 - b1;
          const uint32_t h_e = pack_f16x2(ye0, ye1), h_o = pack_f16x2(yo0, yo1);
          const float2 fe = unpack_f16x2(h_e), fo = unpack_f16x2(h_o);
          eh[e >> 1] = h_e;
          oh[e >> 1] = h_o;
          el[e >> 1] = pack_f16x2(ye0 - fe.x, ye1 - fe.y);
          ol[e >> 1] = pack_f16x2(yo0 - fo.x, yo1 - fo.y);
        }
        }
        ptx::mbar_wait(&a_empty[s], ((cnt / NSTA) & 1) ^ 1);    // the MMAs that read this stage have retired
        uint8_t* st = sA + s * A_STAGE + row_off;
        if (!(p.dbg & 8)) {
        *reinterpret_cast<uint4*>(st) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
        *reinterpret_cast<uint4*>(st + SLICE) = make_uint4(el[0], el[1], el[2], el[3]);
        *reinterpret_cast<uint4*>(st + 2 * SLICE) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        *reinterpret_cast<uint4*>(st + 3 * SLICE) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
        fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&a_full[s], 0);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = frame =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint32_t it = 0;
    for (int tile = pair; tile < p.num_tiles; tile += npairs, ++it) {
      const int b = tile / p.tiles_per_row, f0 = (tile - b * p.tiles_per_row) * (2 * FR) + (int)rank * FR;
      const int f = f0 + r;
      int nvalid = p.F;
      if (p.lengths) {
        const long long len = p.lengths[b];
        const long long sq = (len >= 0 ? len / p.hop : -((-len + p.hop - 1) / p.hop)) + 1;
        nvalid = (int)(sq < 0 ? 0 : (sq > p.F ? p.F : sq));
      }
      const bool store = f < p.F, counted = f < nvalid;
      float* out = p.logmel + (size_t)b * p.nfilt * p.F + f;
      float* part = p.partials ? p.partials + (((size_t)b * p.pslots + (f0 + q * 32) / 32) * p.nfilt) * 2 : nullptr;
      ptx::mbar_wait(d_full, it & 1);
      ptx::tc_fence_after();
      float a0 = 0.f, a1 = 0.f, p256 = 0.f;
      int m = 0;   // filter accumulated in a0
      auto emit = [&]() {
        const float v = logf(a0 + 5.9604644775390625e-08f);   // log(x + 2^-24), transform.py:253
        if (m < p.nfilt) {
          if (store) out[(size_t)m * p.F] = v;
          if (part && f0 + q * 32 < p.F) {
            const float c = counted ? v : 0.f;
            const float s1 = warp_sum(c), s2 = warp_sum(c * c);
            if (lane == 0) *reinterpret_cast<float2*>(part + 2 * m) = make_float2(s1, s2);
          }
        }
        a0 = a1;
        a1 = 0.f;
        ++m;
      };
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t vr[32], vi[32];
        __syncwarp();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * c);
        if (!(p.dbg & 16)) {
        ptx::tmem_ld_32x32(taddr, vr);
        ptx::tmem_ld_32x32(taddr + 256, vi);
        ptx::tmem_ld_wait();
        }
        if (c == 7) {   // everything this thread needs has left TMEM
          ptx::tc_fence_before();
          mbar_arrive_cluster(d_empty, 0);
        }
        if (p.dbg & 1) continue;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int k = 32 * c + e;
          const float re = __uint_as_float(vr[e]), im = __uint_as_float(vi[e]);
          float pw;
          if (k == 0) {        // column 0 of the sine GEMM carries bin 256
            p256 = im * im * P_UNSCALE;
            pw = re * re * P_UNSCALE;
          } else {
            pw = fmaf(re, re, im * im) * P_UNSCALE;   // (the reference's sqrt-then-square differs by <= 2 ulp)
          }
          for (int adv = s_adv[k]; adv > 0; --adv) emit();   // uniform
          const float2 w = s_melw[k];
          a0 = fmaf(w.x, pw, a0);
          a1 = fmaf(w.y, pw, a1);
        }
      }
      for (int adv = s_adv[256]; adv > 0; --adv) emit();
      {
        const float2 w = s_melw[256];
        a0 = fmaf(w.x, p256, a0);
        a1 = fmaf(w.y, p256, a1);
      }
      while (m < p.nfilt) emit();
    }
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    tmem2_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace fdft
}  // namespace ts

using namespace ts;

extern "C" int ts_logmel_dft(const float* audio, int B, int N, int hop, float preemph, const float* wplus,
                             const float* wminus, const void* basis, const float* mel_w2, const int32_t* mel_adv,
                             int nfilt, float* logmel, float* partials, const int64_t* lengths, void* stream) {
  TS_REQUIRE(audio && wplus && wminus && basis && mel_w2 && mel_adv && logmel, TS_ERR_INVALID, "ts_logmel_dft: null pointer");
  TS_REQUIRE(B > 0 && hop > 0 && nfilt > 0 && nfilt <= fdft::MAX_NFILT, TS_ERR_INVALID, "ts_logmel_dft: bad sizes");
  TS_REQUIRE(N > fdft::NFFT / 2, TS_ERR_INVALID,
             "ts_logmel_dft: reflect padding needs N > n_fft/2 (N=%d), same as torch.stft(center=True)", N);
  TS_REQUIRE((reinterpret_cast<uintptr_t>(basis) & 15) == 0, TS_ERR_INVALID, "ts_logmel_dft: basis must be 16-byte aligned");
  fdft::Params p;
  memset(&p, 0, sizeof(p));
  p.audio = audio; p.B = B; p.N = N; p.hop = hop; p.preemph = preemph;
  p.F = 1 + N / hop;
  p.wplus = wplus; p.wminus = wminus;
  p.basis = reinterpret_cast<const uint4*>(basis);
  p.mel_w = reinterpret_cast<const float2*>(mel_w2);
  p.mel_adv = mel_adv;
  p.nfilt = nfilt;
  p.logmel = logmel; p.partials = partials; p.lengths = lengths;
  p.tiles_per_row = ceil_div(p.F, 2 * fdft::FR);
  p.pslots = ceil_div(p.F, 32);
  p.dbg = option_dbg();
  const long long nt = (long long)B * p.tiles_per_row;
  TS_REQUIRE(nt < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_logmel_dft: too many frame tiles");
  p.num_tiles = (int)nt;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    TS_CUDA(cudaGetDevice(&dev));
    TS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    TS_CUDA(cudaFuncSetAttribute(fdft::logmel_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fdft::SMEM_BYTES));
  }
  int pairs = num_sms / 2;
  if (p.num_tiles < pairs) pairs = p.num_tiles;
  fdft::logmel_dft_kernel<<<dim3(2 * pairs), dim3(fdft::THREADS), fdft::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  TS_LAUNCH_CHECK("logmel_dft_kernel");
  return TS_OK;
}
