// Persistent 256 x 256 tile variant of the pointwise GEMM (see pwgemm.cu for the operand conventions).
//
// Why: a 128x128 tile needs 32 KB of shared-memory fill per 256 tensor-core cycles (128 B/clk/SM), more than an
// SM can ingest from L2 (~118 B/clk measured via l1tex__m_xbar2l1tex_read_bytes), and a one-tile-per-CTA launch pays
// TMEM allocation / barrier setup / pipeline fill for every 0.5-1 us of MMA work.  Here one persistent CTA per SM
// loops over 256(Cout) x 256(frames) tiles: 64 KB per k-chunk feed 1024 MMA cycles (64 B/clk/SM), the producer warp
// runs ahead across tile boundaries, and the epilogue leaves through shared memory + TMA stores (full 128-byte
// rows, no per-lane scattered 16-byte stores).
//
//   warp 0      TMA producer (A: one [256 x 64] SW128 K-major box, B: four [64 k x 64 t] SW128 MN-major boxes)
//   warp 1      MMA issuer: per k-step two tcgen05.mma M128 x N256 x K16, accumulators = TMEM columns [0,256), [256,512)
//   warp 2      TMEM allocation (all 512 columns)
//   warps 4-11  epilogue: warp (q, h) drains TMEM lane quarter q of accumulator h, 64 columns at a time:
//               +shift / SE pool / gate*y1 / ReLU / zero tail -> bf16 -> SW128 staging -> TMA store
#include "ts_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace ts {
int option_pw_bn();
namespace pw2 {

constexpr int BM = 256, BK = 64, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;          // 32 KB
constexpr int EPI_WARPS = 8;
constexpr int STG_BYTES = 32 * 128;           // one epilogue warp's staging tile: 32 rows x 64 bf16
constexpr int THREADS = 384;
constexpr int TMEM_COLS = 512;

// BN = 256: one accumulator set (2 x 256 columns), 64 B/clk/SM of operand fill -- for tensor-bound (large K) GEMMs.
// BN = 128: two accumulator sets, the epilogue of tile i overlaps the MMAs of tile i+1 -- for the HBM-bound
//           small-K QuartzNet layers where the epilogue is as long as the main loop.
template <int BN>
struct Cfg {
  static constexpr int ACC = TMEM_COLS / (2 * BN);
  static constexpr int B_BYTES = BK * BN * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES + 256 + 1024;
};

struct Params {
  CUtensorMap a0, b0, a1, b1, out;
  int kc0, kc1;
  int Cout, T, B;
  int m_tiles, n_tiles, num_tiles;
  const float* shift;
  const int32_t* lens;
  int out_pitch;
  int relu;
  unsigned long long* pool;
  const float* se_scale;
  const __nv_bfloat16* y1;
  int y1_pitch;
  int debug_no_loads;   // experiment: the producer signals the stages full without issuing any TMA load
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

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
pw_gemm_big_kernel(const __grid_constant__ Params p) {
  constexpr int STAGES = Cfg<BN>::STAGES;
  constexpr int STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
  constexpr int ACC = Cfg<BN>::ACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + EPI_WARPS * STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.kc0 + p.kc1;

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
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], EPI_WARPS * 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: runs ahead over this CTA's whole tile list =====
    uint32_t cnt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int mt = tile % p.m_tiles, rest = tile / p.m_tiles;
      const int nt = rest % p.n_tiles, b = rest / p.n_tiles;
      const int m0 = mt * BM, t0 = nt * BN;
      for (int kc = 0; kc < num_k; ++kc, ++cnt) {
        const int s = cnt % STAGES;
        ptx::mbar_wait(&empty_bar[s], ((cnt / STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        if (p.debug_no_loads) {
          ptx::mbar_arrive(&full_bar[s]);
          continue;
        }
        ptx::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        const bool seg1 = kc >= p.kc0;
        const CUtensorMap* ma = seg1 ? &p.a1 : &p.a0;
        const CUtensorMap* mb = seg1 ? &p.b1 : &p.b0;
        const int k0 = (seg1 ? kc - p.kc0 : kc) * BK;
        ptx::tma_load_2d(sa, ma, &full_bar[s], k0, m0);                   // [256 rows x 64 k]
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          ptx::tma_load_3d(sb + j * (BK * 128), mb, &full_bar[s], t0 + 64 * j, k0, b);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN, 0, 1);
    uint32_t cnt = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int a = it % ACC;
      ptx::mbar_wait(&tmem_empty[a], ((it / ACC) & 1) ^ 1);   // epilogue has drained this accumulator set
      ptx::tc_fence_after();
      for (int kc = 0; kc < num_k; ++kc, ++cnt) {
        const int s = cnt % STAGES;
        ptx::mbar_wait(&full_bar[s], (cnt / STAGES) & 1);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t db = ptx::umma_desc(sb + k * 2048, BK * 128, 1024);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t da = ptx::umma_desc(sa + h * (128 * BK * 2) + k * 32, 0, 1024);
            ptx::mma_bf16_ss(tmem_base + (a * 2 + h) * BN, da, db, idesc, (kc > 0 || k > 0) ? 1u : 0u);
          }
        }
        ptx::mma_commit(&empty_bar[s]);
      }
      ptx::mma_commit(&tmem_full[a]);
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int e = warp - 4;
    const int q = warp & 3;          // TMEM lane quarter accessible to this warp
    const int h = e >> 2;            // accumulator / M-block handled by this warp
    uint8_t* stg = stg_base + e * STG_BYTES;
    const uint32_t stg_row = ptx::smem_u32(stg) + lane * 128;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile % p.m_tiles, rest = tile / p.m_tiles;
      const int nt = rest % p.n_tiles, b = rest / p.n_tiles;
      const int t0 = nt * BN;
      const int mrow0 = mt * BM + h * 128 + q * 32;
      const int m = mrow0 + lane;
      const bool m_ok = m < p.Cout;
      const float shift = (m_ok && p.shift) ? p.shift[m] : 0.f;
      const int len = p.lens ? min(p.lens[b], p.T) : p.T;
      const float gate = (m_ok && p.se_scale) ? p.se_scale[(size_t)b * p.Cout + m] : 0.f;
      float pooled = 0.f;
      const int a = it % ACC;
      ptx::mbar_wait(&tmem_full[a], (it / ACC) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN / 64; ++cc) {
        uint32_t v[64];
        __syncwarp();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((a * 2 + h) * BN + cc * 64);
        ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        ptx::tmem_ld_wait();
        if (cc == BN / 64 - 1) {  // accumulators fully read: the MMA warp may start the next tile
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tmem_empty[a]);
        }
        const int tb = t0 + cc * 64;
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
                r[g * 8 + 2 * hh] += gate * __uint_as_float(w[hh] << 16);
                r[g * 8 + 2 * hh + 1] += gate * __uint_as_float(w[hh] & 0xFFFF0000u);
              }
            }
          }
          if (lane == 0) bulk_wait_read0();   // previous TMA store has finished reading the staging tile
          __syncwarp();
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 64; ++j) r[j] = fmaxf(r[j], 0.f);
          }
          if (tb + 64 > len) {   // block crosses the utterance end (warp-uniform): zero the tail
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (tb + j >= len) r[j] = 0.f;
          }
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            uint32_t pk[4];
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
              const int j = g * 8 + 2 * hh;
              __nv_bfloat162 pr = __floats2bfloat162_rn(r[j], r[j + 1]);
              pk[hh] = *reinterpret_cast<uint32_t*>(&pr);
            }
            // SWIZZLE_128B: 16-byte chunk g of row `lane` lives at chunk (g ^ (row & 7))
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row + ((g ^ (lane & 7)) << 4)),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                         : "memory");
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
    }
    if (lane == 0) bulk_wait0();   // all stores of this warp are complete before the CTA exits
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace pw2

// bf16-row outputs with Cout > 128; returns TS_ERR_UNSUPPORTED for anything else (caller falls back to pwgemm.cu)
int launch_pw_gemm_big(const void* w0, const void* x0, int cin0, int x0_pitch, const void* w1, const void* x1, int cin1,
                       int x1_pitch, int B, int Cout, int T, const float* shift, const int32_t* lens, void* out,
                       int out_pitch, int relu, unsigned long long* pool, const float* se_scale, const void* y1, int y1_pitch,
                       cudaStream_t st) {
  if (Cout <= 128 || out_pitch % 64 != 0) return TS_ERR_UNSUPPORTED;
  pw2::Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  // tile width: 256 frames for tensor-bound (K >= 1024) GEMMs, 128 with double-buffered accumulators otherwise
  int BN = (cin0 + cin1 >= 1024) ? 256 : 128;
  if ((option_pw_bn() & 0x1ff) == 128 || (option_pw_bn() & 0x1ff) == 256) BN = option_pw_bn() & 0x1ff;  // experiment
  p.debug_no_loads = (option_pw_bn() & 0x1000) ? 1 : 0;
  if ((rc = tma::make_2d_bf16(&p.a0, w0, cin0, Cout, (uint64_t)cin0 * 2, pw2::BK, pw2::BM)) != TS_OK) return rc;
  if ((rc = tma::make_3d_bf16(&p.b0, x0, T, cin0, B, (uint64_t)x0_pitch * 2, (uint64_t)cin0 * x0_pitch * 2, 64, pw2::BK,
                              1)) != TS_OK)
    return rc;
  p.kc0 = ceil_div(cin0, pw2::BK);
  if (cin1 > 0) {
    if ((rc = tma::make_2d_bf16(&p.a1, w1, cin1, Cout, (uint64_t)cin1 * 2, pw2::BK, pw2::BM)) != TS_OK) return rc;
    if ((rc = tma::make_3d_bf16(&p.b1, x1, T, cin1, B, (uint64_t)x1_pitch * 2, (uint64_t)cin1 * x1_pitch * 2, 64,
                                pw2::BK, 1)) != TS_OK)
      return rc;
    p.kc1 = ceil_div(cin1, pw2::BK);
  }
  // output rows [B, Cout, out_pitch]: the pad frames [T, out_pitch) are written too (as zeros), rows >= Cout clipped
  if ((rc = tma::make_3d_bf16(&p.out, out, out_pitch, Cout, B, (uint64_t)out_pitch * 2, (uint64_t)Cout * out_pitch * 2,
                              64, 32, 1)) != TS_OK)
    return rc;
  p.Cout = Cout; p.T = T; p.B = B;
  p.m_tiles = ceil_div(Cout, pw2::BM);
  p.n_tiles = ceil_div(out_pitch, BN);
  p.num_tiles = p.m_tiles * p.n_tiles * B;
  p.shift = shift; p.lens = lens; p.out_pitch = out_pitch; p.relu = relu;
  p.pool = pool; p.se_scale = se_scale; p.y1 = reinterpret_cast<const __nv_bfloat16*>(y1); p.y1_pitch = y1_pitch;

  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    TS_CUDA(cudaGetDevice(&dev));
    TS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    TS_CUDA(cudaFuncSetAttribute(pw2::pw_gemm_big_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 pw2::Cfg<128>::SMEM_BYTES));
    TS_CUDA(cudaFuncSetAttribute(pw2::pw_gemm_big_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 pw2::Cfg<256>::SMEM_BYTES));
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (BN == 256)
    pw2::pw_gemm_big_kernel<256><<<grid, pw2::THREADS, pw2::Cfg<256>::SMEM_BYTES, st>>>(p);
  else
    pw2::pw_gemm_big_kernel<128><<<grid, pw2::THREADS, pw2::Cfg<128>::SMEM_BYTES, st>>>(p);
  TS_LAUNCH_CHECK("pw_gemm_big_kernel");
  return TS_OK;
}

}  // namespace ts
