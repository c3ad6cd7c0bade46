// Pointwise (1x1) convolution as a TMA-fed tcgen05 / TMEM bf16 GEMM.
//
//   out[b, m, t] = epilogue( sum_seg sum_k W_seg[m, k] * X_seg[b, k, t] )         m = output channel
//
// * NCW is kept: for each batch b the activation matrix X[b] is [C, Tp] with time contiguous, i.e. the
//   MN-major ("transposed") B operand of tcgen05.mma; weights [Cout, Cin] are the K-major A operand.
// * Up to two K segments share one TMEM accumulator: segment 0 is the sub-block's pointwise conv, segment 1
//   the block's residual 1x1 conv (QuartznetBlock.forward, quartznet/blocks.py:329-337).  Eval-mode BatchNorm
//   scales (quartznet/blocks.py:222) are folded into the bf16 weights on the host, the shifts are summed into
//   `shift[m]`, so residual-add + BN + ReLU cost nothing beyond the epilogue add.
// * Epilogue (thread = TMEM lane = output channel, 32 consecutive frames per tcgen05.ld): + shift, optional
//   SqueezeExcite handling (pool partial sums / scale-and-add of the main branch), ReLU, zero frames beyond
//   the utterance length (what the next MaskedConv1d would do at its input), bf16 or fp32 store.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue.  One 128 x BN output tile per CTA; 2 CTAs are co-resident per SM so one tile's
// epilogue overlaps the other's main loop.
#include "ts_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace ts {
namespace pw {

constexpr int BM = 128;  // output channels per tile (TMEM lanes)
constexpr int BK = 64;   // input channels per pipeline stage (one 128-byte swizzle atom of bf16)
constexpr int UMMA_K = 16;

template <int BN>
struct Tile {
  static constexpr int kStages = (BN == 128) ? 3 : 3;
  static constexpr int kABytes = BM * BK * 2;                // 16 KB
  static constexpr int kBBytes = BK * BN * 2;                // 16 KB (BN=128) / 32 KB (BN=256)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = BN;                       // power of two >= 32
};

struct Params {
  CUtensorMap a0, b0, a1, b1;  // segment 0 / 1: weights [Cout, K] (K-major), activations [B, K, Tp] (time-major)
  int kc0, kc1;                // number of 64-channel chunks per segment (kc1 == 0: no second segment)
  int Cout, T, B;
  const float* shift;          // [Cout] or null
  const int32_t* lens;         // [B] valid output frames, or null (no tail zeroing)
  void* out;                   // bf16 rows [B, Cout, out_pitch] or f32 [B, Cout, out_pitch]
  int out_pitch;
  int out_f32;
  int relu;
  unsigned long long* pool;    // SE squeeze: [B, Cout] fixed-point (2^-32) sums over t < T of the pre-activation output
  const float* se_scale;       // SE excite:  [B, Cout] sigmoid gate applied to y1 before the residual add
  const __nv_bfloat16* y1;     // main-branch output [B, Cout, y1_pitch] read by the SE-apply epilogue
  int y1_pitch;
  int f16;                     // operands / 16-bit output rows are IEEE fp16 instead of bf16
};

template <int BN>
__global__ void __launch_bounds__(256, 2)
pw_gemm_kernel(const __grid_constant__ Params p) {
  using T = Tile<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + T::kStages * T::kStageBytes);
  uint64_t* empty_bar = full_bar + T::kStages;
  uint64_t* tmem_full_bar = empty_bar + T::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int t0 = blockIdx.y * BN;
  const int b = blockIdx.z;
  const int num_k = p.kc0 + p.kc1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.a0);
    ptx::prefetch_tensormap(&p.b0);
    if (p.kc1 > 0) {
      ptx::prefetch_tensormap(&p.a1);
      ptx::prefetch_tensormap(&p.b1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < T::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, T::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int kc = 0; kc < num_k; ++kc) {
      const int s = kc % T::kStages;
      const uint32_t phase = (kc / T::kStages) & 1;
      ptx::mbar_wait(&empty_bar[s], phase ^ 1);
      uint8_t* sa = smem + s * T::kStageBytes;
      uint8_t* sb = sa + T::kABytes;
      ptx::mbar_arrive_expect_tx(&full_bar[s], T::kStageBytes);
      const bool seg1 = kc >= p.kc0;
      const CUtensorMap* ma = seg1 ? &p.a1 : &p.a0;
      const CUtensorMap* mb = seg1 ? &p.b1 : &p.b0;
      const int k0 = (seg1 ? kc - p.kc0 : kc) * BK;
      ptx::tma_load_2d(sa, ma, &full_bar[s], k0, m0);
#pragma unroll
      for (int j = 0; j < BN / 64; ++j)  // one [64 k x 64 t] SWIZZLE_128B atom column per copy
        ptx::tma_load_3d(sb + j * (BK * 64 * 2), mb, &full_bar[s], t0 + 64 * j, k0, b);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer (single thread) =====
    const uint32_t idesc = ptx::umma_idesc_16(BM, BN, /*a_mn_major=*/0, /*b_mn_major=*/1, p.f16);
    for (int kc = 0; kc < num_k; ++kc) {
      const int s = kc % T::kStages;
      const uint32_t phase = (kc / T::kStages) & 1;
      ptx::mbar_wait(&full_bar[s], phase);
      ptx::tc_fence_after();
      const uint32_t sa = ptx::smem_u32(smem + s * T::kStageBytes);
      const uint32_t sb = sa + T::kABytes;
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        // A: K-major SW128, 8-row groups 1024 B apart; 16 bf16 of K = 32 B inside the swizzle atom
        const uint64_t da = ptx::umma_desc(sa + k * 32, 0, 1024);
        // B: MN-major SW128, 64-frame atom columns BK*128 B apart (LBO), 8-k groups 1024 B apart (SBO);
        //    16 k rows = 2048 B
        const uint64_t db = ptx::umma_desc(sb + k * 2048, BK * 128, 1024);
        ptx::mma_bf16_ss(tmem_base, da, db, idesc, (kc > 0 || k > 0) ? 1u : 0u);
      }
      ptx::mma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
    }
    ptx::mma_commit(tmem_full_bar);    // accumulator complete
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane;  // output channel of this thread
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const bool m_ok = m < p.Cout;
    const float shift = (m_ok && p.shift) ? p.shift[m] : 0.f;
    const int len = p.lens ? p.lens[b] : p.T;
    const float gate = (m_ok && p.se_scale) ? p.se_scale[(size_t)b * p.Cout + m] : 0.f;
    float pooled = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      ptx::tmem_ld_wait();
      const int tb = t0 + c0;
      if (m_ok && tb < p.out_pitch) {
      float r[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]) + shift;
      if (p.pool) {
#pragma unroll
        for (int j = 0; j < 32; ++j) pooled += (tb + j < p.T) ? r[j] : 0.f;
      }
      if (p.y1) {
        const uint4* yp = reinterpret_cast<const uint4*>(p.y1 + ((size_t)b * p.Cout + m) * p.y1_pitch + tb);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 u = yp[g];
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 f = unpack16x2(w[h], p.f16 != 0);
            r[g * 8 + 2 * h] += gate * f.x;
            r[g * 8 + 2 * h + 1] += gate * f.y;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = r[j];
        if (p.relu) x = fmaxf(x, 0.f);
        if (tb + j >= len) x = 0.f;
        r[j] = x;
      }
      if (p.out_f32) {
        float* o = reinterpret_cast<float*>(p.out) + ((size_t)b * p.Cout + m) * p.out_pitch + tb;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (tb + j < p.out_pitch) o[j] = r[j];
      } else {
        // out_pitch is a multiple of 64 and tb a multiple of 32: the 32-frame chunk is all-in or all-out
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                            ((size_t)b * p.Cout + m) * p.out_pitch + tb);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            w[h] = pack16x2(r[g * 8 + 2 * h], r[g * 8 + 2 * h + 1], p.f16 != 0);
          }
          o[g] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      }  // m_ok
    }
    if (p.pool && m_ok) se_pool_add(p.pool + (size_t)b * p.Cout + m, pooled);
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, T::kTmemCols);
  }
}

}  // namespace pw
}  // namespace ts

namespace ts {
int launch_pw_gemm_big(const void* w0, const void* x0, int cin0, int x0_pitch, const void* w1, const void* x1, int cin1,
                       int x1_pitch, int B, int Cout, int T, const float* shift, const int32_t* lens, void* out,
                       int out_pitch, int relu, unsigned long long* pool, const float* se_scale, const void* y1, int y1_pitch,
                       cudaStream_t st);
int launch_pw_gemm_pair(const void* w0, const void* x0, int cin0, int x0_pitch, const void* w1, const void* x1, int cin1,
                        int x1_pitch, int B, int Cout, int T, const float* shift, const int32_t* lens, void* out,
                        int out_pitch, int relu, unsigned long long* pool, const float* se_scale, const void* y1, int y1_pitch,
                        cudaStream_t st, float* stats = nullptr, int f16 = 0, int wconst = 0);
int option_pw_pair();
}
using namespace ts;

// Training forward: z = W a (bf16 rows, no epilogue) plus per-block BatchNorm partial sums from the GEMM epilogue.
extern "C" int ts_pw_gemm_stats(const void* w, const void* x, int cin, int x_pitch, int B, int Cout, int T, void* out,
                                int out_pitch, float* stats, int slots, void* stream) {
  TS_REQUIRE(w && x && out && stats, TS_ERR_INVALID, "ts_pw_gemm_stats: null pointer");
  TS_REQUIRE(B > 0 && B <= 65535 && Cout > 0 && T > 0 && cin > 0 && cin % 8 == 0, TS_ERR_INVALID, "ts_pw_gemm_stats: bad sizes");
  TS_REQUIRE(x_pitch % 8 == 0 && x_pitch >= T && out_pitch % 64 == 0 && out_pitch >= T, TS_ERR_INVALID,
             "ts_pw_gemm_stats: bad pitches");
  TS_REQUIRE(slots == 2 * ceil_div(out_pitch, 256), TS_ERR_INVALID,
             "ts_pw_gemm_stats: slots must be 2 * ceil(out_pitch / 256) (one per 128-frame block)");
  if (!(option_pw_big() && option_pw_pair() > 0)) return TS_ERR_UNSUPPORTED;
  return launch_pw_gemm_pair(w, x, cin, x_pitch, nullptr, nullptr, 0, 0, B, Cout, T, nullptr, nullptr, out, out_pitch, 0,
                             nullptr, nullptr, nullptr, 0, (cudaStream_t)stream, stats);
}

// Host side: see include/thunder_b200.h for the contract.
extern "C" int ts_pw_gemm(const void* w0, const void* x0, int cin0, int x0_pitch, const void* w1, const void* x1,
                          int cin1, int x1_pitch, int B, int Cout, int T, const float* shift, const int32_t* lens,
                          void* out, int out_dtype, int out_pitch, int flags, int64_t* pool_fixed, const float* se_scale,
                          const void* y1, int y1_pitch, void* stream) {
  unsigned long long* pool = reinterpret_cast<unsigned long long*>(pool_fixed);
  const int relu = flags & TS_PW_RELU;
  const int f16 = (flags & TS_ROWS_F16) ? 1 : 0;
  TS_REQUIRE(w0 && x0 && out, TS_ERR_INVALID, "ts_pw_gemm: null pointer");
  TS_REQUIRE(B > 0 && Cout > 0 && T > 0 && cin0 > 0, TS_ERR_INVALID, "ts_pw_gemm: bad sizes");
  TS_REQUIRE(cin0 % 8 == 0 && (cin1 % 8 == 0), TS_ERR_UNSUPPORTED,
             "ts_pw_gemm: input channels must be multiples of 8 (16-byte TMA rows), got %d / %d", cin0, cin1);
  TS_REQUIRE(x0_pitch % 8 == 0 && x0_pitch >= T && (cin1 == 0 || (x1_pitch % 8 == 0 && x1_pitch >= T)), TS_ERR_INVALID,
             "ts_pw_gemm: activation pitch must be a multiple of 8 frames and >= T");
  TS_REQUIRE(out_dtype == TS_F32 || out_dtype == (f16 ? TS_F16 : TS_BF16), TS_ERR_INVALID,
             "ts_pw_gemm: out dtype must be TS_F32 or the operand format (TS_BF16, or TS_F16 with TS_ROWS_F16)");
  const bool out16 = out_dtype != TS_F32;
  TS_REQUIRE(out_pitch >= T && (!out16 || out_pitch % 64 == 0), TS_ERR_INVALID,
             "ts_pw_gemm: bf16 output pitch must be a multiple of 64 frames and >= T (got %d)", out_pitch);
  TS_REQUIRE(!y1 || (y1_pitch % 64 == 0 && se_scale), TS_ERR_INVALID, "ts_pw_gemm: y1 needs se_scale and a 64-multiple pitch");
  TS_REQUIRE((cin1 == 0) == (w1 == nullptr) && (cin1 == 0) == (x1 == nullptr), TS_ERR_INVALID,
             "ts_pw_gemm: segment 1 pointers and cin1 disagree");
  TS_REQUIRE(B <= 65535, TS_ERR_UNSUPPORTED, "ts_pw_gemm: B > 65535");

  // option pw_pair: 1 = CTA-pair kernel for K >= 1024, 2 = for every bf16-row GEMM with Cout > 128
  if (out16 && option_pw_big() && option_pw_pair() > 0 &&
      (option_pw_pair() >= 2 || cin0 + cin1 >= 1024 || f16)) {
    const int rc = launch_pw_gemm_pair(w0, x0, cin0, x0_pitch, w1, x1, cin1, x1_pitch, B, Cout, T, shift, lens, out,
                                       out_pitch, relu, pool, se_scale, y1, y1_pitch, (cudaStream_t)stream, nullptr, f16,
                                       (flags & TS_PW_CONST_WEIGHTS) ? 1 : 0);
    if (rc != TS_ERR_UNSUPPORTED) return rc;
  }
  if (out16 && !f16 && option_pw_big()) {
    const int rc = launch_pw_gemm_big(w0, x0, cin0, x0_pitch, w1, x1, cin1, x1_pitch, B, Cout, T, shift, lens, out,
                                      out_pitch, relu, pool, se_scale, y1, y1_pitch, (cudaStream_t)stream);
    if (rc != TS_ERR_UNSUPPORTED) return rc;
  }
  constexpr int BN = 128;
  pw::Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  // weights [Cout, cin] bf16 row-major: dims (k, m); box 64 x 128, SWIZZLE_128B
  if ((rc = tma::make_2d_bf16(&p.a0, w0, cin0, Cout, (uint64_t)cin0 * 2, pw::BK, pw::BM)) != TS_OK) return rc;
  // activations [B, cin, pitch] bf16: dims (t, k, b) with the time extent clipped to T (pad frames read as 0)
  if ((rc = tma::make_3d_bf16(&p.b0, x0, T, cin0, B, (uint64_t)x0_pitch * 2, (uint64_t)cin0 * x0_pitch * 2, 64, pw::BK,
                              1)) != TS_OK)
    return rc;
  p.kc0 = ceil_div(cin0, pw::BK);
  if (cin1 > 0) {
    if ((rc = tma::make_2d_bf16(&p.a1, w1, cin1, Cout, (uint64_t)cin1 * 2, pw::BK, pw::BM)) != TS_OK) return rc;
    if ((rc = tma::make_3d_bf16(&p.b1, x1, T, cin1, B, (uint64_t)x1_pitch * 2, (uint64_t)cin1 * x1_pitch * 2, 64,
                                pw::BK, 1)) != TS_OK)
      return rc;
    p.kc1 = ceil_div(cin1, pw::BK);
  }
  p.Cout = Cout;
  p.T = T;
  p.B = B;
  p.shift = shift;
  p.lens = lens;
  p.out = out;
  p.out_pitch = out_pitch;
  p.out_f32 = (out_dtype == TS_F32);
  p.relu = relu;
  p.pool = pool;
  p.se_scale = se_scale;
  p.y1 = reinterpret_cast<const __nv_bfloat16*>(y1);
  p.y1_pitch = y1_pitch;
  p.f16 = f16;

  auto kern = pw::pw_gemm_kernel<BN>;
  constexpr int smem = pw::Tile<BN>::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    TS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(Cout, pw::BM), ceil_div(T, BN), B);
  kern<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  TS_LAUNCH_CHECK("pw_gemm_kernel");
  return TS_OK;
}
