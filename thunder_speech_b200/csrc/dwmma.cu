// Depthwise conv1d on the tensor cores: the FIR of one channel is a banded Toeplitz matrix product over time.
//
// For channel c (stride 1, dilation 1, odd K, "same" padding P = K/2) cut the time axis into windows of 64
// frames.  With input rows X[m, j] = x[64(m-1) + j] (row 0 = the zero padding on the left) and Toeplitz blocks
// Tq[j, r] = w[64q + j - 64 - r + P] (zero outside [0, K)):
//
//     Y[i, r] = y[64 i + r] = sum_{q=0..2} sum_{j<64} X[i + q, j] * Tq[j, r]
//
// i.e. three [128 x 64] x [64 x 64] bf16 MMAs accumulate one [128 windows x 64 frames] output tile in TMEM
// (fp32).  The 128 M-rows are the windows of up to floor(128 / (W + 2)) utterances stacked on top of each other
// (W = pitch / 64 windows per row, +2 halo rows), so a tile is 8192 frames of ONE channel.  A k-slice that only
// multiplies structural zeros of Tq is skipped (K = 33 needs 6 of the 12 MMAs).
//
// Operand layout: both operands are K-major, no swizzle, stored as 8 "columns" of 16-byte chunks
// (chunk kc of row r at col*COLSTRIDE + r*16): shifting a tile down by q rows is just +16 q bytes on the
// descriptor start address, which is what makes the three shifted A tiles free (one staged copy of the input).
// The input is staged with 16-byte cp.async whose src-size implements the length mask exactly
// (frames t >= len_in[b] are zero-filled: MaskedConv1d.mask_fill, quartznet/blocks.py:158-167).
//
// Warp roles (288 threads): warps 0-3 cp.async loaders, warps 4-7 epilogue (TMEM -> bf16 -> global, 128
// contiguous bytes per thread), warp 8 TMEM allocation + single-thread MMA issue.  One CTA = one channel and a
// slice of the batch; the Toeplitz blocks are built once per CTA from the fp32 taps (rounded to bf16).
#include "ts_common.cuh"
#include "sm100_ptx.cuh"

namespace ts {
namespace dwt {

constexpr int L = 64;                 // frames per window = MMA N
constexpr int MROWS = 128;            // MMA M
constexpr int AROWS = MROWS + 3;      // +2 for the q shifts, +1 so the column stride is an odd number of chunks
constexpr int COLB = AROWS * 16;      // byte stride between the 8 k-chunk columns of an A stage
constexpr int A_STAGE = 8 * COLB;     // 16768 B
constexpr int NSTAGE = 3;
constexpr int BQ = 8 * 64 * 16;       // one Toeplitz block: 8 columns x 64 rows x 16 B = 8 KB
constexpr int NQ = 3;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * L;  // 128
constexpr int NLOAD = 128;
constexpr int THREADS = 288;
constexpr int SMEM_BYTES = NSTAGE * A_STAGE + NQ * BQ + 256 + 128;

struct Params {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float* w;
  const int32_t* lens;  // may be null
  int B, C, T, pitch, K, P;
  int W, R, NB;         // windows per row, stacked rows per utterance (W + 2), utterances per tile
  int tiles_per_chan, tiles_per_cta;
};

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// no-swizzle K-major descriptor: 8-row groups 128 B apart (SBO), the two 16-byte k-chunks of one MMA `lbo` apart
__device__ __forceinline__ uint64_t desc_kmajor_noswz(uint32_t addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(THREADS, 2)
dw_mma_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + NSTAGE * A_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + NQ * BQ);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* acc_full = empty_bar + NSTAGE;
  uint64_t* acc_empty = acc_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x;
  const int tile0 = blockIdx.y * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.tiles_per_chan - tile0);

  // ---- Toeplitz blocks of this channel (all threads), barriers, TMEM --------------------------------
  {
    const float* wc = p.w + (size_t)c * p.K;
    for (int ci = tid; ci < NQ * 8 * 64; ci += THREADS) {
      const int q = ci / 512, col = (ci >> 6) & 7, r = ci & 63;
      const int d0 = 64 * q + 8 * col - 64 - r + p.P;
      uint32_t pk[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int da = d0 + 2 * h, db = da + 1;
        const float fa = (da >= 0 && da < p.K) ? wc[da] : 0.f;
        const float fb = (db >= 0 && db < p.K) ? wc[db] : 0.f;
        __nv_bfloat162 pr = __floats2bfloat162_rn(fa, fb);
        pk[h] = *reinterpret_cast<uint32_t*>(&pr);
      }
      *reinterpret_cast<uint4*>(sB + q * BQ + col * 1024 + r * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    // stage rows beyond the stacked utterances only feed accumulator rows that are never stored; zero them once
    const int r0 = p.NB * p.R, nz = AROWS - r0;  // rows the loaders never touch
    for (int ci = tid; ci < NSTAGE * 8 * nz; ci += THREADS) {
      const int s = ci / (8 * nz), col = (ci / nz) & 7, r = r0 + ci % nz;
      *reinterpret_cast<uint4*>(sA + s * A_STAGE + col * COLB + r * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) {
        ptx::mbar_init(&full_bar[s], NLOAD);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < ACC_STAGES; ++a) {
        ptx::mbar_init(&acc_full[a], 1);
        ptx::mbar_init(&acc_empty[a], 128);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===== loaders: 128 threads; thread geometry (row-in-utterance m, chunk column) is loop invariant =====
    const __nv_bfloat16* xc = p.x + (size_t)c * p.pitch;
    const size_t bstride = (size_t)p.C * p.pitch;
    const int rchunks = p.R * 8;  // 16-byte chunks per stacked utterance
    for (int n = 0; n < ntiles; ++n) {
      const int s = n % NSTAGE;
      ptx::mbar_wait(&empty_bar[s], ((n / NSTAGE) & 1) ^ 1);
      const int b0 = (tile0 + n) * p.NB;
      const uint32_t dst0 = ptx::smem_u32(sA + s * A_STAGE);
      for (int ch = tid; ch < rchunks; ch += NLOAD) {
        const int m = ch >> 3, col = ch & 7;
        const int t = 64 * (m - 1) + 8 * col;
        const uint32_t dst = dst0 + col * COLB + m * 16;
        const __nv_bfloat16* src = xc + (size_t)b0 * bstride + t;
#pragma unroll 3
        for (int bl = 0; bl < p.NB; ++bl) {
          const int b = b0 + bl;
          int nbytes = 0;
          if (b < p.B && t >= 0) {
            int lin = p.T;
            if (p.lens) lin = min(lin, max(__ldg(p.lens + b), 0));
            nbytes = min(max((lin - t) * 2, 0), 16);
          }
          cp_async_16_zfill(dst + bl * p.R * 16, nbytes > 0 ? (const void*)(src + bl * bstride) : (const void*)p.x,
                            nbytes);
        }
      }
      cp_async_commit();
      if (n > 0) {  // publish the previous tile once its copies have landed
        cp_async_wait<1>();
        fence_proxy_async();
        ptx::mbar_arrive(&full_bar[(n - 1) % NSTAGE]);
      }
    }
    if (ntiles > 0) {
      cp_async_wait<0>();
      fence_proxy_async();
      ptx::mbar_arrive(&full_bar[(ntiles - 1) % NSTAGE]);
    }
  } else if (warp == 8) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(MROWS, L, 0, 0);
      const uint32_t sb = ptx::smem_u32(sB);
      const int jlo = 64 - p.P, jhi = 127 + p.P;  // non-zero band of the stacked Toeplitz rows
      for (int n = 0; n < ntiles; ++n) {
        const int s = n % NSTAGE, a = n % ACC_STAGES;
        ptx::mbar_wait(&acc_empty[a], ((n / ACC_STAGES) & 1) ^ 1);
        ptx::mbar_wait(&full_bar[s], (n / NSTAGE) & 1);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(sA + s * A_STAGE);
        uint32_t acc = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int j0 = 64 * q + 16 * ks;
            if (j0 + 15 < jlo || j0 > jhi) continue;
            const uint64_t da = desc_kmajor_noswz(sa + (2 * ks) * COLB + q * 16, COLB);
            const uint64_t db = desc_kmajor_noswz(sb + q * BQ + (2 * ks) * 1024, 1024);
            ptx::mma_bf16_ss(tmem_base + a * L, da, db, idesc, acc);
            acc = 1;
          }
        }
        ptx::mma_commit(&empty_bar[s]);
        ptx::mma_commit(&acc_full[a]);
      }
    }
  } else {
    // ===== epilogue: warps 4..7, TMEM lane quarter = warp % 4 =====
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int bl = row / p.R, i = row - bl * p.R;
    const int t = 64 * i;
    const bool row_ok = bl < p.NB && i < p.W;
    for (int n = 0; n < ntiles; ++n) {
      const int a = n % ACC_STAGES;
      const int b = (tile0 + n) * p.NB + bl;
      int lout = p.T;  // fetched before the wait so its latency hides behind the MMA
      if (row_ok && b < p.B && p.lens) lout = min(lout, max(__ldg(p.lens + b), 0));
      ptx::mbar_wait(&acc_full[a], (n / ACC_STAGES) & 1);
      ptx::tc_fence_after();
      uint32_t v[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * L);
      ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc_empty[a]);  // accumulator stage can be overwritten
      if (row_ok && b < p.B) {
        uint4* o = reinterpret_cast<uint4*>(p.y + ((size_t)b * p.C + c) * p.pitch + t);
        if (t + 64 <= lout) {  // interior window: no masking
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            uint32_t pk[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              __nv_bfloat162 pr = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * h]),
                                                        __uint_as_float(v[g * 8 + 2 * h + 1]));
              pk[h] = *reinterpret_cast<uint32_t*>(&pr);
            }
            o[g] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        } else {  // window crossing (or beyond) the utterance end: frames >= lout are stored as zero
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            uint32_t pk[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int e = g * 8 + 2 * h;
              const float lo = (t + e < lout) ? __uint_as_float(v[e]) : 0.f;
              const float hi = (t + e + 1 < lout) ? __uint_as_float(v[e + 1]) : 0.f;
              __nv_bfloat162 pr = __floats2bfloat162_rn(lo, hi);
              pk[h] = *reinterpret_cast<uint32_t*>(&pr);
            }
            o[g] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      __syncwarp();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace dwt

// used by ts_dw_conv (dwconv.cu); returns TS_ERR_UNSUPPORTED when the shape is outside this kernel's domain
int launch_dw_mma(const __nv_bfloat16* x, int B, int C, int T, int pitch_in, const float* w, int K, int P,
                  const int32_t* lens, __nv_bfloat16* y, int pitch_out, cudaStream_t st) {
  if (!(K % 2 == 1 && P == K / 2 && K <= 129 && pitch_in == pitch_out && pitch_in % 64 == 0 && C <= 65535))
    return TS_ERR_UNSUPPORTED;
  dwt::Params p;
  p.x = x; p.y = y; p.w = w; p.lens = lens;
  p.B = B; p.C = C; p.T = T; p.pitch = pitch_in; p.K = K; p.P = P;
  p.W = pitch_in / 64;
  p.R = p.W + 2;
  if (p.R > dwt::MROWS) return TS_ERR_UNSUPPORTED;  // rows longer than 8064 frames: SIMT path
  p.NB = dwt::MROWS / p.R;
  p.tiles_per_chan = ceil_div(B, p.NB);
  // CTAs = C x groups; pick tiles-per-CTA for wave efficiency (2 CTAs/SM x 148 SMs per wave) while amortising the
  // per-CTA Toeplitz build (about one tile's worth of work) over several tiles
  {
    double best = -1.0;
    int best_tpc = p.tiles_per_chan;
    for (int tpc = 1; tpc <= p.tiles_per_chan; ++tpc) {
      const int g = ceil_div(p.tiles_per_chan, tpc);
      const double waves = (double)C * g / 296.0;
      const double eff = waves / (double)((long long)(waves + 0.999999)) * (double)p.tiles_per_chan / (double)(g * tpc) *
                         (double)tpc / (double)(tpc + 1);
      if (eff > best + 1e-9) {
        best = eff;
        best_tpc = tpc;
      }
    }
    p.tiles_per_cta = best_tpc;
  }
  const int groups = ceil_div(p.tiles_per_chan, p.tiles_per_cta);
  static bool attr_set = false;
  if (!attr_set) {
    TS_CUDA(cudaFuncSetAttribute(dwt::dw_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dwt::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(C, groups);
  dwt::dw_mma_kernel<<<grid, dwt::THREADS, dwt::SMEM_BYTES, st>>>(p);
  TS_LAUNCH_CHECK("dw_mma_kernel");
  return TS_OK;
}

}  // namespace ts
