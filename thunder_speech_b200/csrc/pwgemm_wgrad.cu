// Pointwise-conv weight gradient on the tensor cores:
//
//     dW[co, ci] = sum_b sum_t dz[b, co, t] * a[b, ci, t]
//
// Both operands are activation rows [B, C, pitch] with time contiguous, i.e. BOTH are K-major for a GEMM whose
// reduction dimension is (batch, time): A = dz tile [128 co x 64 t], B = a tile [BN ci x 64 t], SWIZZLE_128B boxes
// straight from the 3-D tensor maps (frames >= T are zero-filled by TMA).  The reduction is cut into k-chunks of 64 frames
// numbered over (utterance, time chunk); grid = (Cout/128, Cin/BN, SPLIT) and each CTA reduces an even share of the
// k-chunks into a TMEM accumulator and writes an fp32 partial [SPLIT, Cout, Cin].  The output matrix is tiny next to the
// reduction length (256x256 vs 24k frames), so SPLIT is chosen to fill all 148 SMs; ts_pw_wgrad_reduce then sums the
// partials in a fixed order (deterministic, no atomics).  Same warp-specialised TMA / MMA / epilogue structure as pwgemm.cu.
#include "ts_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace ts {
namespace wg {

constexpr int BM = 128, BK = 64, UMMA_K = 16, BN = 256, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;

struct Params {
  CUtensorMap a, b;   // dz rows (t, co, b) box (64, 128, 1);  a rows (t, ci, b) box (64, 256, 1)
  int Cout, Cin, T, B;
  int nsplit;
  float* part;        // [SPLIT, Cout, Cin]
};

__global__ void __launch_bounds__(256, 1)
pw_wgrad_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, split = blockIdx.z;
  const int tchunks = (p.T + BK - 1) / BK;
  const int total_k = p.B * tchunks;
  const int k_begin = (int)((long long)total_k * split / p.nsplit);
  const int num_k = (int)((long long)total_k * (split + 1) / p.nsplit) - k_begin;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.a);
    ptx::prefetch_tensormap(&p.b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // PDL: the prologue above overlapped the previous kernel's tail
  pdl_wait();                // its outputs (dz, a) are complete and visible from here on

  if (warp == 0 && lane == 0) {
    for (int kc = 0; kc < num_k; ++kc) {
      const int s = kc % STAGES;
      ptx::mbar_wait(&empty_bar[s], ((kc / STAGES) & 1) ^ 1);
      uint8_t* sa = smem + s * STAGE_BYTES;
      ptx::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
      const int kg = k_begin + kc;
      const int b = kg / tchunks, t0 = (kg - b * tchunks) * BK;
      ptx::tma_load_3d(sa, &p.a, &full_bar[s], t0, m0, b);
      ptx::tma_load_3d(sa + A_BYTES, &p.b, &full_bar[s], t0, n0, b);
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, 0, 0);
    for (int kc = 0; kc < num_k; ++kc) {
      const int s = kc % STAGES;
      ptx::mbar_wait(&full_bar[s], (kc / STAGES) & 1);
      ptx::tc_fence_after();
      const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES);
      const uint32_t sb = sa + A_BYTES;
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        const uint64_t da = ptx::umma_desc(sa + k * 32, 0, 1024);
        const uint64_t db = ptx::umma_desc(sb + k * 32, 0, 1024);
        ptx::mma_bf16_ss(tmem_base, da, db, idesc, (kc > 0 || k > 0) ? 1u : 0u);
      }
      ptx::mma_commit(&empty_bar[s]);
    }
    ptx::mma_commit(tmem_full);
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    if (num_k > 0) {
      ptx::mbar_wait(tmem_full, 0);
      ptx::tc_fence_after();
    }
    float* orow = p.part + ((size_t)split * p.Cout + m) * p.Cin;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      __syncwarp();
      if (num_k > 0) {
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (m < p.Cout) {
        if ((p.Cin & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (n0 + c0 + j < p.Cin)
              *reinterpret_cast<uint4*>(orow + n0 + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < p.Cin) orow[n0 + c0 + j] = __uint_as_float(v[j]);
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, BN);
  }
}

// out[i] = sum_s part[s][i] in a fixed order; n4 = elements / 4.  A block owns 32 float4 columns; its 8 warps each sum a
// contiguous slice of the splits (all loads independent), then warp 0 adds the 8 slice sums in ascending order.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float4* __restrict__ part, int nsplit, long long n4, float4* __restrict__ out) {
  __shared__ float4 sm[8][32];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  const int per = (nsplit + 7) >> 3;
  const int s0 = w * per, s1 = min(nsplit, s0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    int s = s0;
    for (; s + 4 <= s1; s += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(part + (long long)(s + u) * n4 + i);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
    for (; s < s1; ++s) {
      const float4 v = __ldg(part + (long long)s * n4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  sm[w][lane] = acc;
  __syncthreads();
  if (w == 0 && i < n4) {
    float4 t = sm[0][lane];
#pragma unroll
    for (int u = 1; u < 8; ++u) {
      t.x += sm[u][lane].x; t.y += sm[u][lane].y; t.z += sm[u][lane].z; t.w += sm[u][lane].w;
    }
    out[i] = t;
  }
}

}  // namespace wg
}  // namespace ts

using namespace ts;

extern "C" int ts_pw_wgrad(const void* dz, int dz_pitch, const void* a, int a_pitch, int B, int Cout, int Cin, int T,
                           int nsplit, float* part, void* stream) {
  TS_REQUIRE(dz && a && part, TS_ERR_INVALID, "ts_pw_wgrad: null pointer");
  TS_REQUIRE(B > 0 && Cout > 0 && Cin > 0 && T > 0 && nsplit > 0, TS_ERR_INVALID, "ts_pw_wgrad: bad sizes");
  TS_REQUIRE(nsplit <= B * ceil_div(T, wg::BK), TS_ERR_INVALID, "ts_pw_wgrad: more splits than 64-frame chunks");
  TS_REQUIRE(dz_pitch % 8 == 0 && a_pitch % 8 == 0 && dz_pitch >= T && a_pitch >= T, TS_ERR_INVALID,
             "ts_pw_wgrad: pitches must be multiples of 8 frames and >= T");
  wg::Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = tma::make_3d_bf16(&p.a, dz, T, Cout, B, (uint64_t)dz_pitch * 2, (uint64_t)Cout * dz_pitch * 2, wg::BK, wg::BM,
                              1)) != TS_OK)
    return rc;
  if ((rc = tma::make_3d_bf16(&p.b, a, T, Cin, B, (uint64_t)a_pitch * 2, (uint64_t)Cin * a_pitch * 2, wg::BK, wg::BN,
                              1)) != TS_OK)
    return rc;
  p.Cout = Cout; p.Cin = Cin; p.T = T; p.B = B;
  p.nsplit = nsplit;
  p.part = part;
  static bool attr_set = false;
  if (!attr_set) {
    TS_CUDA(cudaFuncSetAttribute(wg::pw_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(ceil_div(Cout, wg::BM), ceil_div(Cin, wg::BN), nsplit);
  TS_CUDA(launch_pdl(wg::pw_wgrad_kernel, grid, dim3(256), wg::SMEM_BYTES, (cudaStream_t)stream, (option_pdl() & 2) != 0, p));
  TS_LAUNCH_CHECK("pw_wgrad_kernel");
  return TS_OK;
}

extern "C" int ts_pw_wgrad_reduce(const float* part, int nsplit, long long n, float* out, void* stream) {
  TS_REQUIRE(part && out, TS_ERR_INVALID, "ts_pw_wgrad_reduce: null pointer");
  TS_REQUIRE(nsplit > 0 && n > 0 && n % 4 == 0, TS_ERR_INVALID, "ts_pw_wgrad_reduce: n must be a positive multiple of 4");
  TS_REQUIRE((((uintptr_t)part | (uintptr_t)out) & 15) == 0, TS_ERR_INVALID, "ts_pw_wgrad_reduce: 16-byte alignment required");
  const long long n4 = n / 4;
  TS_CUDA(launch_pdl(wg::wgrad_reduce_kernel, dim3((unsigned)ceil_div64(n4, 32)), dim3(256), 0, (cudaStream_t)stream,
                     (option_pdl() & 2) != 0, reinterpret_cast<const float4*>(part), nsplit, n4, reinterpret_cast<float4*>(out)));
  TS_LAUNCH_CHECK("wgrad_reduce_kernel");
  return TS_OK;
}
