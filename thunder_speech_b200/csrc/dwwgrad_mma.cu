// Depthwise-conv tap gradient on the tensor cores (stride 1):
//
//     dw[c, k] = sum_b sum_t da[b, c, t] * x[b, c, t + k*D - P]
//
// Per channel this is a correlation of two long signals; cut both into 64-frame windows w (over all utterances) and it
// becomes a small GEMM whose REDUCTION dimension is the window index:
//
//     G[i, e] = sum_w  da[64 w + i] * x[64 (w - HL) + e]          i in [0, 64),  e in [0, 64 NQ)
//     dw[k]   = sum_i  G[i, i + k*D - P + 64 HL]                  (a diagonal sum; HL = ceil(P / 64) halo windows)
//
// Both operands are staged by ONE 4-D TMA box each -- (64 frames, R windows, 1 channel, NB utterances), out-of-range
// windows / utterances zero-filled -- into SWIZZLE_128B rows of 128 bytes (one window per row).  Read "transposed" these
// rows are MN-major UMMA operands with the window index as K: A = da (64 useful rows; the second 64-row chunk of the
// M = 128 instruction re-reads the same rows and its output lanes are ignored -- M = 64 costs the same tensor time), B = x
// with N = 64 NQ <= 256: chunk q of B is the same staged copy addressed q rows further down (leading-dimension byte offset
// = 128: overlapping atoms; the swizzle XOR is a function of the absolute shared-memory address so shifted reads stay
// consistent).  So ONE tcgen05.mma per 16 windows.  One CTA owns a channel: it accumulates all utterance groups into TMEM
// (64 NQ columns), then the epilogue warps scatter G into a [64][130] shared tile (each (i, k) has exactly one source)
// and sum the 64 entries of every tap in a fixed order -- deterministic, no partial buffers, dw written once.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 epilogue; TMEM is double-buffered across channels.
#include "ts_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace ts {
namespace dwg {

constexpr int THREADS = 256;
constexpr int NSTAGE = 4;
constexpr int ACC = 2;
constexpr int MAX_ROWS = 128;                       // staged windows per operand per stage
constexpr int XS_BYTES = MAX_ROWS * 128 + 1024;     // + 8 zero rows behind the box for the row-shifted chunks
constexpr int DA_BYTES = MAX_ROWS * 128;
constexpr int STAGE_BYTES = XS_BYTES + DA_BYTES;
constexpr int MAX_K = 128;
constexpr int SI = 130;                              // S row stride (floats): SI - 1 odd -> conflict-free scatter
constexpr int S_BYTES = 64 * SI * 4;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + S_BYTES + 256 + 1024;
constexpr int TMEM_COLS = 512;

struct Params {
  CUtensorMap x, da;   // (64 frames, W windows, C, B), box (64, R, 1, NB)
  int B, C, K, P, D;
  int R, NB, HL, NQ, ACC, rows, groups;
  float* out;          // [C, K]
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
dw_wgrad_mma_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* S = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES + S_BYTES);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full = empty_bar + NSTAGE;
  uint64_t* tmem_empty = tmem_full + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int box_bytes = p.rows * 128;

  // zero the rows behind the x box of every stage once (never written by TMA, read by the shifted chunks)
  for (int i = threadIdx.x; i < NSTAGE * 64; i += THREADS) {
    const int s = i >> 6, off = (i & 63) << 4;
    *reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + box_bytes + off) = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.x);
    ptx::prefetch_tensormap(&p.da);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  // make the generic-proxy zero fill visible to the async proxy (tcgen05.mma reads shared memory through it)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int acc_cols = p.NQ * 64;
  pdl_launch_dependents();   // PDL: the prologue above overlapped the previous kernel's tail
  pdl_wait();                // its outputs (da, x) are complete and visible from here on

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    int it = 0;
    for (int c = blockIdx.x; c < p.C; c += gridDim.x) {
      for (int g = 0; g < p.groups; ++g, ++it) {
        const int s = it % NSTAGE;
        ptx::mbar_wait(&empty_bar[s], ((it / NSTAGE) & 1) ^ 1);
        uint8_t* xs = smem + s * STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * box_bytes);
        tma_load_4d(xs, &p.x, &full_bar[s], 0, -p.HL, c, g * p.NB);
        tma_load_4d(xs + XS_BYTES, &p.da, &full_bar[s], 0, 0, c, g * p.NB);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t n1 = (uint32_t)(p.NQ < 4 ? p.NQ : 4) * 64, n2 = (uint32_t)p.NQ * 64 - n1;
    const uint32_t idesc1 = ptx::umma_idesc_bf16(128, (int)n1, 1, 1);   // both operands MN-major
    const uint32_t idesc2 = ptx::umma_idesc_bf16(128, n2 ? (int)n2 : 64, 1, 1);
    const int ksteps = p.rows >> 4;
    int it = 0, ch = 0;
    for (int c = blockIdx.x; c < p.C; c += gridDim.x, ++ch) {
      const int acc = p.ACC == 2 ? (ch & 1) : 0;
      ptx::mbar_wait(&tmem_empty[acc], (p.ACC == 2 ? ((ch >> 1) & 1) : (ch & 1)) ^ 1);
      ptx::tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(acc * acc_cols);
      for (int g = 0; g < p.groups; ++g, ++it) {
        const int s = it % NSTAGE;
        ptx::mbar_wait(&full_bar[s], (it / NSTAGE) & 1);
        ptx::tc_fence_after();
        const uint32_t sx = ptx::smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sd = sx + XS_BYTES;
        for (int ks = 0; ks < ksteps; ++ks) {
          // rows ks*16 .. +15 are the K slice.  A: da rows (second M chunk = the same rows again, LBO 0 rows apart is
          // expressed as one full operand further: any finite data would do, those lanes are never read).
          const uint64_t da_ = ptx::umma_desc(sd + ks * 2048, 0, 1024);
          const uint64_t dx = ptx::umma_desc(sx + ks * 2048, 128, 1024);          // N chunks 0..3: +128 B each
          ptx::mma_bf16_ss(d0, da_, dx, idesc1, (g > 0 || ks > 0) ? 1u : 0u);
          if (n2) {
            const uint64_t dx2 = ptx::umma_desc(sx + ks * 2048 + 4 * 128, 128, 1024);
            ptx::mma_bf16_ss(d0 + n1, da_, dx2, idesc2, (g > 0 || ks > 0) ? 1u : 0u);
          }
        }
        ptx::mma_commit(&empty_bar[s]);
      }
      ptx::mma_commit(&tmem_full[acc]);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: diagonal sums
    const int q = warp & 3;
    const int i = q * 32 + lane;              // TMEM lane = row of G = frame within the da window (useful: 0..63)
    const int tid_e = threadIdx.x - 128;      // 0..127
    const int off = 64 * p.HL - p.P;
    int ch = 0;
    for (int c = blockIdx.x; c < p.C; c += gridDim.x, ++ch) {
      const int acc = p.ACC == 2 ? (ch & 1) : 0;
      ptx::mbar_wait(&tmem_full[acc], p.ACC == 2 ? ((ch >> 1) & 1) : (ch & 1));
      ptx::tc_fence_after();
      if (q < 2) {
        for (int h = 0; h < 2 * p.NQ; ++h) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_cols + h * 32), v);
          ptx::tmem_ld_wait();
          if (p.D == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int k = h * 32 + j - i - off;
              if (k >= 0 && k < p.K) S[i * SI + k] = __uint_as_float(v[j]);
            }
          } else if (p.D == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int kd = h * 32 + j - i - off;
              if (kd >= 0 && !(kd & 1) && (kd >> 1) < p.K) S[i * SI + (kd >> 1)] = __uint_as_float(v[j]);
            }
          } else {
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
              const int kd = h * 32 + j - i - off;
              const int k = kd / p.D;
              if (kd >= 0 && k * p.D == kd && k < p.K) S[i * SI + k] = __uint_as_float(v[j]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
      named_bar_sync(1, 128);
      if (tid_e < p.K) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int r = 0; r < 64; r += 4) {
          s0 += S[r * SI + tid_e];
          s1 += S[(r + 1) * SI + tid_e];
          s2 += S[(r + 2) * SI + tid_e];
          s3 += S[(r + 3) * SI + tid_e];
        }
        p.out[(size_t)c * p.K + tid_e] = (s0 + s1) + (s2 + s3);
      }
      named_bar_sync(1, 128);
    }
  }
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace dwg
}  // namespace ts

using namespace ts;

// returns TS_ERR_UNSUPPORTED when the shape is outside this kernel's envelope (caller falls back to the SIMT kernels).
// Requires rows that are zero in [T, pitch) and beyond each utterance's length (true for every producer in this library).
int launch_dw_wgrad_mma(const __nv_bfloat16* da, int pitch_out, const __nv_bfloat16* x, int pitch_in, int B, int C, int K,
                        int D, int P, float* out, cudaStream_t st) {
  if (pitch_in != pitch_out || pitch_in % 64 != 0 || K > dwg::MAX_K || D < 1 || P <= 0 || C > 65535)
    return TS_ERR_UNSUPPORTED;
  dwg::Params p;
  memset(&p, 0, sizeof(p));
  const int W = pitch_in / 64;
  p.B = B; p.C = C; p.K = K; p.P = P; p.D = D;
  p.HL = ceil_div(P, 64);
  const int emax = 63 + (K - 1) * D - P + 64 * p.HL;
  if (emax < 0) return TS_ERR_UNSUPPORTED;
  p.NQ = emax / 64 + 1;
  if (p.NQ > 8) return TS_ERR_UNSUPPORTED;
  p.ACC = p.NQ * 64 * 2 <= dwg::TMEM_COLS ? 2 : 1;   // double-buffer the accumulator across channels when it fits
  p.R = round_up(W + p.NQ - 1, 16);
  if (p.R > dwg::MAX_ROWS) return TS_ERR_UNSUPPORTED;
  p.NB = dwg::MAX_ROWS / p.R;
  p.rows = p.NB * p.R;
  p.groups = ceil_div(B, p.NB);
  p.out = out;
  int rc;
  cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[3] = {128, (cuuint64_t)pitch_in * 2, (cuuint64_t)C * pitch_in * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)p.R, 1, (cuuint32_t)p.NB};
  if ((rc = tma::encode(&p.x, x, 4, dims, strides, box)) != TS_OK) return rc;
  if ((rc = tma::encode(&p.da, da, 4, dims, strides, box)) != TS_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    TS_CUDA(cudaFuncSetAttribute(dwg::dw_wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dwg::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = C < 148 ? C : 148;
  TS_CUDA(launch_pdl(dwg::dw_wgrad_mma_kernel, dim3(grid), dim3(dwg::THREADS), dwg::SMEM_BYTES, st, (option_pdl() & 2) != 0, p));
  TS_LAUNCH_CHECK("dw_wgrad_mma_kernel");
  return TS_OK;
}
