/*
 * thunder_b200.h -- C ABI of libthunder_b200.so: the B200 (sm_100a) implementation of the
 * thunder-speech ASR forward hot path.
 *
 * The reference (scart97/thunder-speech, pure Python) has no FFI; its drop-in boundary is the
 * nn.Module protocol `forward(x, lengths) -> (y, lengths)` (src/thunder/blocks.py:94-115).  The
 * Python package `thunder_speech_b200` mirrors those modules and binds the entry points below with
 * ctypes from `torch.library` custom ops.  Conventions for every entry point:
 *
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless named `h_*`;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no allocation, no host
 *     synchronisation, no ownership transfer; inputs are never written;
 *   - return value: TS_OK (0), a negative TS_ERR_* argument error, or a positive cudaError_t;
 *   - `ts_last_error()` returns a thread-local human readable message for the last failure.
 *
 * Activation layout ("rows"): the reference layout [B, C, T] (NCW) is kept, with the time axis of
 * every internal bf16 tensor padded to a pitch `Tp` (multiple of 64 frames) so that each (b, c) row
 * starts 128-byte aligned and TMA boxes never straddle rows.  Frames t >= T inside the pitch are
 * don't-care.  `ts_row_pitch(T)` gives the pitch.
 */
#ifndef THUNDER_B200_H_
#define THUNDER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TS_OK 0
#define TS_ERR_INVALID (-1)      /* bad argument (shape, null pointer, range)            */
#define TS_ERR_UNSUPPORTED (-2)  /* valid for the reference, not implemented by this build */
#define TS_ERR_NO_DEVICE (-3)    /* no sm_100 device / driver entry point missing         */

/* dtype tags for `void*` tensors */
#define TS_F32 0
#define TS_BF16 1

/* ---- library ------------------------------------------------------------------------------ */
const char* ts_version(void);
const char* ts_last_error(void);
/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
int64_t ts_launch_count(void);
/* pitch (in frames) of a padded activation row holding T frames */
int ts_row_pitch(int T);

/* ---- (1) feature front-end ---------------------------------------------------------------- */
/* Replaces FilterbankFeatures stages 1-3 in eval mode: PreEmphasisFilter.forward
 * (src/thunder/quartznet/transform.py:136-144), PowerSpectrum.forward (transform.py:186-208:
 * reflect-pad n_fft/2, frames of n_fft every `hop`, window zero-padded and centred in n_fft,
 * one-sided |STFT|^2) and MelScale.forward (transform.py:243-255: log(fb @ P + 2^-24)).
 *   audio      [B, N] f32 (row stride N)            window_full [n_fft] f32 (already centred/padded),
 *              non-zero only inside [win_lo, win_hi) (a [96,416) support selects the sparse-input FFT)
 *   twiddle    [n_fft] float2 = exp(-2*pi*i*e/n_fft)
 *   mel_start/mel_count/mel_off [nfilt] i32, mel_w [nnz] f32: row-compressed filter bank, filter m
 *              covers FFT bins [mel_start[m], mel_start[m]+mel_count[m]) with weights mel_w[mel_off[m]..]
 *   logmel     [B, nfilt, F] f32 out, F = 1 + N / hop
 * Only n_fft == 512 is implemented (TS_ERR_UNSUPPORTED otherwise); N must exceed n_fft/2 like
 * torch.stft's reflect padding.  */
int ts_logmel(const float* audio, int B, int N, int n_fft, int hop, float preemph,
              const float* window_full, int win_lo, int win_hi, const float* twiddle,
              const int32_t* mel_start, const int32_t* mel_count, const int32_t* mel_off,
              const float* mel_w, int nfilt, int nnz,
              float* logmel, void* stream);

/* Replaces PowerSpectrum.get_sequence_length (transform.py:182-184) + FeatureBatchNormalizer.forward
 * (transform.py:77-92 -> src/thunder/blocks.py:118-149): seq_len = floor(len/hop)+1; per (b, feature)
 * masked mean / biased std over t < seq_len, (x-mean)/(std+div_guard), zero for t >= seq_len.
 *   lengths [B] i64 (audio samples)    seq_len_out [B] i64 (may be NULL)
 *   out: f32 -> [B, nfilt, F] (out_pitch == F) or bf16 -> [B, nfilt, out_pitch] padded rows */
int ts_feature_normalize(const float* logmel, const int64_t* lengths, int B, int nfilt, int F, int hop,
                         float div_guard, void* out, int out_dtype, int out_pitch,
                         int64_t* seq_len_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* THUNDER_B200_H_ */
