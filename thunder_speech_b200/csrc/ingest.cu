// Audio ingest before the feature kernel: AudioFileLoader.preprocess_audio
// (src/thunder/data/dataset.py:50-77) for a padded batch on the device.
//
//   ts_pcm_ingest   int16 / float32 PCM, [B, channels, N] planar or [B, N, channels] interleaved (the wav frame order)
//                   -> mono mix (mean over channels), int16 scaled by 1/32768 (torchaudio.load's normalisation), DC removal
//                   (minus the mean over the utterance's own len[b] samples), tail zeroed -> float32 [B, N]
//   ts_resample     torchaudio.functional.resample's windowed-sinc polyphase FIR (taps supplied by the caller, built once
//                   per rate pair): y[b, f*new' + p] = sum_k taps[p][k] * x[b, f*orig' + k - width], x zero outside [0, len)
//
// Both are streaming kernels.  ingest: pass 1 writes one fp64 partial sum per (utterance, 64 Ki-sample chunk), pass 2 adds
// the chunks in a fixed order (deterministic), and re-reads the PCM (L2-resident for int16 batches up to ~100 MB) to write
// the output; 16-byte loads (8 frames per thread) for int16 mono / stereo-interleaved and float mono.  resample: only the
// non-zero taps of every phase are applied (<= 2 width + 2 of the 2 width + orig' torchaudio allocates), read from a
// transposed table so that lanes with consecutive phases load consecutive addresses; 4 outputs per thread.
#include "ts_common.cuh"

namespace ts {
namespace ingest {

constexpr int CHUNK = 65536;

template <typename T>
__device__ __forceinline__ float load_mono(const T* p, long long base, int n, int channels, int N, bool interleaved,
                                           float scale) {
  float s = 0.f;
  if (interleaved) {
    for (int c = 0; c < channels; ++c) s += (float)p[base + (long long)n * channels + c];
  } else {
    for (int c = 0; c < channels; ++c) s += (float)p[base + (long long)c * N + n];
  }
  return s * scale;
}

// 8 consecutive mono-mixed samples starting at frame n (n % 8 == 0) with 16-byte loads where the layout allows it:
// int16 mono, int16 stereo interleaved, float mono; otherwise the generic per-sample path.  `fast` is CTA-uniform.
template <typename T>
__device__ __forceinline__ bool fast_layout(int channels, int N, bool interleaved, const T* pcm) {
  // every utterance row must start 16-byte aligned; the last (partial) group of 8 frames takes the generic path
  if ((reinterpret_cast<uintptr_t>(pcm) & 15) != 0 || (((long long)channels * N * (long long)sizeof(T)) & 15) != 0) return false;
  if (sizeof(T) == 2) return channels == 1 || (channels == 2 && interleaved);
  return channels == 1;
}
template <typename T>
__device__ __forceinline__ void load_mono8(const T* p, long long base, int n, int channels, int N, bool interleaved,
                                           float scale, bool fast, float (&v)[8]) {
  if (fast && sizeof(T) == 2 && channels == 1) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const int16_t*>(p) + base + n));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      v[2 * h] = (float)(int16_t)(w[h] & 0xFFFFu) * scale;
      v[2 * h + 1] = (float)(int16_t)(w[h] >> 16) * scale;
    }
  } else if (fast && sizeof(T) == 2) {   // stereo interleaved: 8 frames = 32 bytes
    const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const int16_t*>(p) + base + 2ll * n);
    const uint4 u0 = __ldg(q), u1 = __ldg(q + 1);
    const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
    for (int h = 0; h < 8; ++h) v[h] = ((float)(int16_t)(w[h] & 0xFFFFu) + (float)(int16_t)(w[h] >> 16)) * scale;
  } else if (fast) {                     // float mono
    const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + base + n);
    const float4 a = __ldg(q), b = __ldg(q + 1);
    v[0] = a.x * scale; v[1] = a.y * scale; v[2] = a.z * scale; v[3] = a.w * scale;
    v[4] = b.x * scale; v[5] = b.y * scale; v[6] = b.z * scale; v[7] = b.w * scale;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (n + j < N) ? load_mono(p, base, n + j, channels, N, interleaved, scale) : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
pcm_sum_kernel(const T* __restrict__ pcm, int B, int channels, int N, const int32_t* __restrict__ lens, int interleaved,
               float scale, double* __restrict__ part, int nchunks) {
  __shared__ double red[8];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, ch = blockIdx.x;
  const int len = lens ? min(max(lens[b], 0), N) : N;
  const long long base = (long long)b * channels * N;
  const int n0 = ch * CHUNK, n1 = min(len, n0 + CHUNK);
  const bool fast = fast_layout(channels, N, interleaved != 0, pcm);
  double acc = 0.0;
  for (int n = n0 + threadIdx.x * 8; n < n1; n += 256 * 8) {   // 8 samples per thread and iteration, summed in fp32
    float v[8];
    load_mono8(pcm, base, n, channels, N, interleaved != 0, scale, fast && n + 8 <= N, v);
    float s8 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s8 += (n + j < n1) ? v[j] : 0.f;
    acc += s8;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    part[(size_t)b * nchunks + ch] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
pcm_write_kernel(const T* __restrict__ pcm, int B, int channels, int N, const int32_t* __restrict__ lens, int interleaved,
                 float scale, const double* __restrict__ part, int nchunks, int remove_dc, float* __restrict__ out,
                 int out_pitch) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int len = lens ? min(max(lens[b], 0), N) : N;
  double s = 0.0;
  if (remove_dc)
    for (int c = 0; c < nchunks; ++c) s += part[(size_t)b * nchunks + c];
  const float mean = (remove_dc && len > 0) ? (float)(s / (double)len) : 0.f;
  const long long base = (long long)b * channels * N;
  const bool fast = fast_layout(channels, N, interleaved != 0, pcm);
  const int n = blockIdx.x * 2048 + threadIdx.x * 8;
  if (n >= out_pitch) return;
  float v[8];
  if (n < len) {
    load_mono8(pcm, base, n, channels, N, interleaved != 0, scale, fast && n + 8 <= N, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (n + j < len) ? v[j] - mean : 0.f;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
  }
  float* o = out + (size_t)b * out_pitch + n;
  if (n + 7 < out_pitch && (out_pitch & 3) == 0) {
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    for (int j = 0; j < 8 && n + j < out_pitch; ++j) o[j] = v[j];
  }
}

// Compressed taps: for phase p only taps k0[p] .. k0[p] + nt - 1 of torchaudio's [new'][2 width + orig'] kernel can be
// non-zero (the Hann window is clamped to exactly 0 beyond +-lowpass_filter_width zero crossings), nt <= 2 width + 2.
// tapsC [nt][newp] (transposed: lanes with consecutive phases read consecutive addresses); 4 outputs per thread,
// 256 apart, so that every thread has four independent FMA chains.
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, int N_in, int in_pitch, const int32_t* __restrict__ len_in, int origp,
                int newp, const float* __restrict__ tapsC, const int32_t* __restrict__ k0, int nt, int width,
                float* __restrict__ out, int N_out, int out_pitch) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int len = len_in ? min(max(len_in[b], 0), N_in) : N_in;
  const long long tgt_ll = ((long long)newp * len + origp - 1) / origp;   // ceil(new' * len / orig')
  const int target = tgt_ll < (long long)N_out ? (int)tgt_ll : N_out;
  const float* xb = x + (size_t)b * in_pitch;
  const int jb = blockIdx.x * 1024 + threadIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int xo[4], ph[4];
  bool live[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = jb + u * 256;
    live[u] = j < target;
    const int f = j / newp;
    ph[u] = j - f * newp;
    xo[u] = live[u] ? f * origp - width + __ldg(k0 + ph[u]) : 0;
    if (!live[u]) ph[u] = 0;
  }
  bool inside = true;   // all four windows fully inside [0, len): no per-tap bounds checks
#pragma unroll
  for (int u = 0; u < 4; ++u) inside = inside && live[u] && xo[u] >= 0 && xo[u] + nt <= len;
  if (inside) {
#pragma unroll 2
    for (int i = 0; i < nt; ++i) {
      const float* tp = tapsC + (size_t)i * newp;
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(__ldg(tp + ph[u]), __ldg(xb + xo[u] + i), acc[u]);
    }
  } else {
    for (int i = 0; i < nt; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = xo[u] + i;
        const float xv = (live[u] && n >= 0 && n < len) ? __ldg(xb + n) : 0.f;
        acc[u] = fmaf(__ldg(tapsC + (size_t)i * newp + ph[u]), xv, acc[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = jb + u * 256;
    if (j < out_pitch) out[(size_t)b * out_pitch + j] = acc[u];   // zero beyond the utterance's resampled length
  }
}

}  // namespace ingest
}  // namespace ts

using namespace ts;

extern "C" int ts_pcm_ingest(const void* pcm, int dtype, int B, int channels, int N, const int32_t* lens, int interleaved,
                             int remove_dc, float* out, int out_pitch, double* scratch, long long scratch_doubles,
                             void* stream) {
  TS_REQUIRE(pcm && out && scratch, TS_ERR_INVALID, "ts_pcm_ingest: null pointer");
  TS_REQUIRE(dtype == TS_F32 || dtype == TS_I16, TS_ERR_INVALID, "ts_pcm_ingest: dtype must be TS_F32 or TS_I16");
  TS_REQUIRE(B > 0 && B <= 65535 && channels > 0 && channels <= 64 && N > 0 && out_pitch >= N, TS_ERR_INVALID,
             "ts_pcm_ingest: bad sizes");
  const int nchunks = ceil_div(N, ingest::CHUNK);
  TS_REQUIRE(scratch_doubles >= (long long)B * nchunks, TS_ERR_INVALID,
             "ts_pcm_ingest: scratch too small (need B * ceil(N / 65536) doubles)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool pdl = (option_pdl() & 2) != 0;
  dim3 g1(nchunks, B), g2(ceil_div(out_pitch, 2048), B);
  if (dtype == TS_I16) {
    const float scale = 1.0f / 32768.0f / (float)channels;
    if (remove_dc)
      TS_CUDA(launch_pdl(ingest::pcm_sum_kernel<int16_t>, g1, dim3(256), 0, st, pdl, (const int16_t*)pcm, B, channels, N, lens,
                         interleaved, scale, scratch, nchunks));
    TS_CUDA(launch_pdl(ingest::pcm_write_kernel<int16_t>, g2, dim3(256), 0, st, pdl, (const int16_t*)pcm, B, channels, N, lens,
                       interleaved, scale, (const double*)scratch, nchunks, remove_dc, out, out_pitch));
  } else {
    const float scale = 1.0f / (float)channels;
    if (remove_dc)
      TS_CUDA(launch_pdl(ingest::pcm_sum_kernel<float>, g1, dim3(256), 0, st, pdl, (const float*)pcm, B, channels, N, lens,
                         interleaved, scale, scratch, nchunks));
    TS_CUDA(launch_pdl(ingest::pcm_write_kernel<float>, g2, dim3(256), 0, st, pdl, (const float*)pcm, B, channels, N, lens,
                       interleaved, scale, (const double*)scratch, nchunks, remove_dc, out, out_pitch));
  }
  TS_LAUNCH_CHECK("pcm_ingest kernels");
  return TS_OK;
}

extern "C" int ts_resample(const float* x, int B, int N_in, int in_pitch, const int32_t* len_in, int orig_p, int new_p,
                           const float* taps_c, const int32_t* k0, int nt, int width, float* out, int N_out, int out_pitch,
                           void* stream) {
  TS_REQUIRE(x && taps_c && k0 && out, TS_ERR_INVALID, "ts_resample: null pointer");
  TS_REQUIRE(B > 0 && B <= 65535 && N_in > 0 && in_pitch >= N_in && N_out > 0 && out_pitch >= N_out, TS_ERR_INVALID,
             "ts_resample: bad sizes");
  TS_REQUIRE(orig_p > 0 && new_p > 0 && nt > 0 && nt <= 2 * width + orig_p && width > 0, TS_ERR_INVALID,
             "ts_resample: bad tap table");
  TS_CUDA(launch_pdl(ingest::resample_kernel, dim3(ceil_div(out_pitch, 1024), B), dim3(256), 0, (cudaStream_t)stream,
                     (option_pdl() & 2) != 0, x, N_in, in_pitch, len_in, orig_p, new_p, taps_c, k0, nt, width, out, N_out,
                     out_pitch));
  TS_LAUNCH_CHECK("resample_kernel");
  return TS_OK;
}
