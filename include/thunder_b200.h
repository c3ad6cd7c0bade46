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
#define TS_I16 2   /* int16 PCM (ts_pcm_ingest only) */
#define TS_FIX32 3 /* int64 fixed point in units of 2^-32 (SqueezeExcite pool sums: order-independent integer atomics) */
#define TS_F16 4   /* IEEE fp16 activation rows / GEMM operands ("half rows"): same bytes and tensor-core rate as bf16,
                    * 11-bit mantissa; conversions saturate to +-65504.  Inference entry points only. */
/* flag shared by the row-consuming inference entry points: the 16-bit rows (and 16-bit weights) are TS_F16, not TS_BF16 */
#define TS_ROWS_F16 2

/* ---- library ------------------------------------------------------------------------------ */
const char* ts_version(void);
const char* ts_last_error(void);
/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
int64_t ts_launch_count(void);
/* runtime switches for A/B measurements: "dw_mma" (default 1) = stride-1 depthwise convs on the tensor cores,
 * "pw_big" (default 1) = persistent 256x256-tile GEMM for bf16 outputs with Cout > 128, "pw_pair" (default 2) = run it
 * on CTA pairs (tcgen05 cta_group::2; 1 = only for K >= 1024, 0 = single-CTA kernel),
 * "pdl" (default 3: bit 0 = depthwise, bit 1 = GEMM) = programmatic dependent launch for the two hot kernels
 * (measured: +1.7 % on the QuartzNet forward, +6 % on the training step),
 * "dw_tma" (default 1) = TMA-fed Toeplitz kernel for pre-masked inputs, "dw_base_offset" (descriptor experiment) */
int ts_set_option(const char* name, int value);
/* diagnostics: install a DEVICE buffer of `slots` x 32 uint64 words; every following launch of the pair GEMM / Toeplitz
 * kernels takes the next slot and its first and last CTA stamp %globaltimer at fixed points of their life (layout:
 * csrc/ts_common.cuh, reader: tools/trace_chain.py).  buffer = NULL switches tracing off.  No reference counterpart
 * (the reference profiles with torch.profiler); used to break down the per-kernel latency of the captured graph. */
int ts_trace(void* buffer, int slots);
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
 * n_fft == 512 (the reference's models) runs the packed radix-8 FFT kernel; any other even n_fft <= 8192 (the reference
 * takes whatever torch.stft takes, transform.py:258-271) runs a direct DFT over the window support -- correct, not tuned.
 * N must exceed n_fft/2 like torch.stft's reflect padding.  */
int ts_logmel(const float* audio, int B, int N, int n_fft, int hop, float preemph,
              const float* window_full, int win_lo, int win_hi, const float* twiddle,
              const int32_t* mel_start, const int32_t* mel_count, const int32_t* mel_off,
              const float* mel_w, int nfilt, int nnz,
              float* logmel, void* stream);
/* ts_logmel with the train()-mode dither of DitherAudio.forward (src/thunder/quartznet/transform.py:109-118) fused in:
 * every sample becomes x + dither * n, n ~ N(0, 1), BEFORE the pre-emphasis, inside the kernel (no extra pass over the
 * audio).  n is counter-based (Philox4x32-10 keyed by `seed`, counter = (sample index / 4, utterance), Box-Muller): the
 * result is a pure function of (audio, seed) -- reproducible, independent of the launch geometry -- but NOT torch's
 * random stream (the reference draws torch.randn_like), so parity with the reference is statistical.  dither = 0 is
 * ts_logmel bit for bit.  seed_dev (nullable, DEVICE pointer to one u64) is XORed into `seed` by the kernel: a per-step
 * state the caller advances on the stream, so that a captured CUDA graph draws fresh noise on every replay. */
int ts_logmel_dither(const float* audio, int B, int N, int n_fft, int hop, float preemph,
                     const float* window_full, int win_lo, int win_hi, const float* twiddle,
                     const int32_t* mel_start, const int32_t* mel_count, const int32_t* mel_off,
                     const float* mel_w, int nfilt, int nnz,
                     float* logmel, float dither, unsigned long long seed, const unsigned long long* seed_dev,
                     void* stream);

/* The same three stages with the STFT as a DFT-MATRIX CONTRACTION on the tensor cores (tcgen05, CTA pairs; the
 * reference's own DFT-matrix STFT is src/thunder/blocks.py:38-91): n_fft = 512 with the window supported on [96, 416)
 * only.  The frame is folded about n = 255.5 (y+[m] = y[256+m], y-[m] = y[255-m], m < 160), so
 * |X[k]|^2 = (sum_m (y+ + y-)[m] cos(th))^2 + (sum_m (y+ - y-)[m] sin(th))^2, th = 2 pi k (m + 1/2) / 512: two real GEMMs
 * [frames x 160] x [160 x 256]; operands are fp16 hi/lo splits (3 MMAs per product, ~22 mantissa bits).
 *   wplus/wminus [160] f32 = window_full[256 + m] / window_full[255 - m]
 *   basis   2 x 163840 bytes: for CTA rank r in {0,1}: matrices C_hi, C_lo, S_hi, S_lo, each 10 k-steps of a
 *           [128 bins (128 r + n) x 16 m] fp16 slice in the no-swizzle K-major core-matrix layout
 *           (byte offset of (n, k) = (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2); C = cos(th), S = sin(th) except
 *           S[m][bin 0] = (-1)^m, which makes sine column 0 carry bin 256
 *   mel_w2  [257][2] f32, mel_adv [257] i32: the filter bank as a SLIDING two-filter window over the ordered bins: before
 *           bin k is accumulated, mel_adv[k] filters are complete and emitted; mel_w2[k] = weights of bin k for the current
 *           and the next filter (requires every bin to touch at most two consecutive filters)
 *   logmel  [B, nfilt, F] f32 out;  partials (nullable) [B, ceil(F/32), nfilt, 2] f32: sum / sum of squares of the
 *           log-mel values over the valid frames (f < lengths[b] / hop + 1; lengths nullable = all) of each 32-frame
 *           group, consumed by ts_feature_normalize's `partials` path
 * |window * audio| must stay below 255 (fp16 operand range after the 2^8 scaling): float audio in [-1, 1]. */
int ts_logmel_dft(const float* audio, int B, int N, int hop, float preemph, const float* wplus, const float* wminus,
                  const void* basis, const float* mel_w2, const int32_t* mel_adv, int nfilt, float* logmel,
                  float* partials, const int64_t* lengths, void* stream);

/* Replaces PowerSpectrum.get_sequence_length (transform.py:182-184) + FeatureBatchNormalizer.forward
 * (transform.py:77-92 -> src/thunder/blocks.py:118-149): seq_len = floor(len/hop)+1; per (b, feature)
 * masked mean / biased std over t < seq_len, (x-mean)/(std+div_guard), zero for t >= seq_len.
 *   lengths [B] i64 (audio samples)    seq_len_out [B] i64 (may be NULL)
 *   out: f32 -> [B, nfilt, F] (out_pitch == F) or bf16 / fp16 -> [B, nfilt, out_pitch] padded rows */
int ts_feature_normalize(const float* logmel, const int64_t* lengths, int B, int nfilt, int F, int hop,
                         float div_guard, void* out, int out_dtype, int out_pitch,
                         int64_t* seq_len_out, void* stream);

/* The same from the partial sums of ts_logmel_dft (`partials` [B, ceil(F/32), nfilt, 2] f32: sum / sum of squares of the
 * valid log-mel frames per 32-frame group): one streaming pass over the log-mel tensor, no reduction pass. */
int ts_feature_normalize_partials(const float* logmel, const float* partials, const int64_t* lengths, int B, int nfilt,
                                  int F, int hop, float div_guard, void* out, int out_dtype, int out_pitch,
                                  int64_t* seq_len_out, void* stream);

/* ---- (2) masked depthwise conv1d ----------------------------------------------------------- */
/* Replaces MaskedConv1d.forward with groups == channels (src/thunder/quartznet/blocks.py:158-182; built by
 * _get_conv_bn_layer, quartznet/blocks.py:195-201): zero frames t >= len_in[b], depthwise FIR, and zero
 * output frames t' >= get_seq_len(len_in[b]) (the mask the following pointwise MaskedConv1d applies).
 *   x  bf16 rows [B, C, pitch_in] holding T_in frames     w  [C, K] f32     len_in [B] i32 or NULL (= all valid)
 *   y  bf16 rows [B, C, pitch_out], T_out = floor((T_in + 2P - D(K-1) - 1)/S) + 1
 * TS_ERR_INVALID when both S > 1 and D > 1 (get_same_padding raises ValueError, src/thunder/blocks.py:192-193).
 *   flags: TS_DW_INPUT_PREMASKED = the caller guarantees x is already zero for t >= len_in[b] (true for rows
 *          produced by ts_pack_rows / ts_pw_gemm / ts_feature_normalize with lengths): enables the TMA-fed kernel;
 *          TS_ROWS_F16 = x and y are fp16 rows (the Toeplitz taps are then rounded to fp16 as well) */
#define TS_DW_INPUT_PREMASKED 1
int ts_dw_conv(const void* x, int B, int C, int T_in, int pitch_in, const float* w, int K, int S, int D, int P,
               const int32_t* len_in, int flags, void* y, int pitch_out, void* stream);

/* ---- (3) pointwise conv / residual / decoder GEMM (tcgen05 + TMEM + TMA) ------------------- */
/* out[b, m, t] = epi( W0[m, :] . X0[b, :, t] + W1[m, :] . X1[b, :, t] + shift[m] )
 * Replaces the 1x1 MaskedConv1d + Masked(BatchNorm1d) + ReLU of a sub-block (quartznet/blocks.py:202-228), the
 * residual branch `out + BN(conv1x1(x))` and `mout` ReLU of QuartznetBlock.forward (quartznet/blocks.py:329-337)
 * when segment 1 is given, and conv1d_decoder (src/thunder/blocks.py:199-216) with out_dtype = TS_F32.
 * Eval BatchNorm scales must be folded into the bf16 weights by the caller; `shift` is the sum of BN shifts / bias.
 *   w0 [Cout, cin0] bf16, x0 bf16 rows [B, cin0, x0_pitch]; segment 1 likewise or NULL/0
 *   lens [B] i32 or NULL: frames t >= lens[b] are stored as zero (input mask of the next MaskedConv1d)
 *   out  bf16 rows [B, Cout, out_pitch] (pitch % 64 == 0) or f32 [B, Cout, out_pitch]
 * SqueezeExcite (citrinet/blocks.py:70-83): `pool` [B, Cout] TS_FIX32 (int64, units of 2^-32; caller zeroes it)
 * accumulates sum_t (acc + shift) over t < T with INTEGER atomics, so the sums -- and with them the logits -- do not
 * depend on the order in which the tiles of an utterance finish; `se_scale` [B, Cout] + `y1` bf16 rows:
 * out = epi(acc + shift + se_scale * y1). */
#define TS_PW_RELU 1
/* the caller guarantees that w0 / w1 were written BEFORE anything still in flight on this stream (folded inference weights,
 * completed by a stream synchronisation when the plan was built): the weight-stationary kernel may then copy them into
 * tensor memory before griddepcontrol.wait, i.e. while the previous kernel is still running */
#define TS_PW_CONST_WEIGHTS 4
/* flags: TS_PW_RELU = ReLU in the epilogue; TS_ROWS_F16 = w*, x*, y1 (and a 16-bit `out`, out_dtype = TS_F16) are IEEE
 * fp16 instead of bf16 (same kernels: the tcgen05 kind::f16 instruction descriptor selects the operand format). */
int ts_pw_gemm(const void* w0, const void* x0, int cin0, int x0_pitch, const void* w1, const void* x1, int cin1,
               int x1_pitch, int B, int Cout, int T, const float* shift, const int32_t* lens, void* out,
               int out_dtype, int out_pitch, int flags, int64_t* pool, const float* se_scale, const void* y1,
               int y1_pitch, void* stream);

/* ---- (4) SqueezeExcite FC, greedy CTC ------------------------------------------------------- */
/* gate[b, :] = sigmoid(W2 relu(W1 (pool[b, :] / T)))  -- SqueezeExcite.fc + sigmoid (citrinet/blocks.py:63-83).
 *   pool [B, C] sums over all T frames: TS_FIX32 as ts_pw_gemm accumulates them, or TS_F32; w1 [H, C] f32,
 *   w2 [C, H] f32 (nn.Linear layouts, no bias), hid [B, H] f32 scratch (the ReLU'd hidden layer), gate [B, C] f32 out */
int ts_se_fc(const void* pool, int pool_dtype, int B, int C, int H, int T, const float* w1, const float* w2, float* hid,
             float* gate, void* stream);

/* out = relu(gate[b, c] * y1[b, c, t]) over bf16 rows: SqueezeExcite scale + `mout` ReLU for blocks without a
 * residual branch (citrinet/blocks.py:154,195-197); frames t >= lens[b] are stored as zero when lens != NULL */
/* flags: TS_PW_RELU, TS_ROWS_F16 (rows are fp16) */
int ts_se_apply(const void* y1, const float* gate, int B, int C, int pitch, const int32_t* lens, int flags,
                void* out, void* stream);

/* Greedy CTC: `pred.argmax(1)` (src/thunder/module.py:100; first maximal index, NaN maximal) followed by the
 * per-row torch.unique_consecutive of decode_prediction (src/thunder/text_processing/transform.py:107-110).
 *   logits [B, V, pitch] f32 or bf16 rows, T valid frames
 *   ids [B, T] i64 argmax; collapsed [B, T] i64 padded with -1; counts [B] i32
 *   drop_blank < 0 keeps blanks (the reference removes the blank token as a string, vocab.py:114-130) */
int ts_ctc_greedy(const void* logits, int dtype, int B, int V, int T, int pitch, int64_t* ids, int64_t* collapsed,
                  int32_t* counts, int drop_blank, void* stream);

/* y[b, c, t'] = x[b, c, S t'] (zero for S t' >= min(T_in, len_in[b]) and in the pad): the input gather of a strided
 * 1x1 residual conv, MaskedConv1d(kernel_size=1, stride=S) (citrinet/blocks.py:159-168, quartznet/blocks.py:301-311);
 * its channel mixing then runs in ts_pw_gemm.  T_out = (T_in - 1) / S + 1 */
int ts_gather_rows(const void* x, int B, int C, int T_in, int pitch_in, int S, const int32_t* len_in, void* y,
                   int pitch_out, void* stream);
/* im2col over 16-bit rows for NON-separable convolutions with kernel_size > 1 (the reference's QuartznetBlock default
 * `separable=False`, src/thunder/quartznet/blocks.py:212-219, built at :232-243):
 *   y[b, c*K + k, t] = x[b, c, t*S + k*D - P]  inside [0, min(T_in, len_in[b])), 0 elsewhere and in [T_out, pitch_out)
 * after which MaskedConv1d.forward (:158-182) is ONE ts_pw_gemm with Cin*K input channels whose weight is conv.weight
 * [Cout, Cin, K] viewed as [Cout, Cin*K].  x [B, C, pitch_in], y [B, rows_out >= C*K, pitch_out] (rows beyond C*K are the
 * caller's: zero them when the GEMM's K extent is padded), either 16-bit row format. */
int ts_im2col_rows(const void* x, int B, int C, int T_in, int pitch_in, int K, int S, int D, int P,
                   const int32_t* len_in, void* y, int rows_out, int pitch_out, void* stream);

/* ---- layout / length plumbing at module boundaries ------------------------------------------ */
/* contiguous [B, C, T] (TS_F32, TS_BF16 or TS_F16) -> 16-bit rows [B, C, pitch] of out_dtype TS_BF16 / TS_F16 (frames
 * >= min(T, lens[b]) zero; lens may be NULL) -- the MaskedConv1d.mask_fill of the first consumer
 * (quartznet/blocks.py:158-167) -- and back to f32 */
int ts_pack_rows(const void* in, int in_dtype, int B, int C, int T, const int32_t* lens, void* out, int out_dtype,
                 int pitch, void* stream);
int ts_unpack_rows(const void* in, int in_dtype, int pitch, int B, int C, int T, float* out, void* stream);
/* MaskedConv1d.get_seq_len (quartznet/blocks.py:142-156) on i32 lengths; i64 <-> i32 conversions */
int ts_conv_lengths(const int32_t* in, int32_t* out, int B, int K, int S, int D, int P, void* stream);
int ts_lengths_to_i32(const int64_t* in, int32_t* out, int B, void* stream);
int ts_lengths_to_i64(const int32_t* in, int64_t* out, int B, void* stream);

/* SpecAugment / SpecCutout (src/thunder/quartznet/spec_augment.py:23-110, training only): zero IN PLACE every element
 * (b, f, t) of feat ([B, C, pitch] f32 or bf16, valid frames < T) that lies in any of the n rectangles
 * rects[i] = {f0, f1, t0, t1} (half-open; device memory, int32).  The same rectangles for every utterance. */
int ts_spec_mask(void* feat, int dtype, int B, int C, int T, int pitch, const int32_t* rects, int n, void* stream);

/* ---- audio ingest before the path ---------------------------------------------------------- */
/* Reference: AudioFileLoader.preprocess_audio (src/thunder/data/dataset.py:50-77) -- mono mix, DC removal, resample
 * (torchaudio.functional.resample) -- applied to a padded batch on the device.
 * ts_pcm_ingest: pcm is int16 (TS_I16, scaled by 1/32768 like torchaudio.load) or float32 (TS_F32), laid out
 * [B, channels, N] (interleaved = 0) or [B, N, channels] (interleaved = 1, the wav frame order); lens[b] (nullable) =
 * valid samples.  out[b, n] = mean_c pcm - (remove_dc ? mean over n < lens[b] : 0) for n < lens[b], 0 up to out_pitch.
 * scratch: >= B * ceil(N / 65536) doubles. */
int ts_pcm_ingest(const void* pcm, int dtype, int B, int channels, int N, const int32_t* lens, int interleaved,
                  int remove_dc, float* out, int out_pitch, double* scratch, long long scratch_doubles, void* stream);
/* ts_resample: y[b, f * new_p + p] = sum_k taps[p][k] * x[b, f * orig_p + k - width] with x zero outside [0, len_in[b]),
 * for f * new_p + p < ceil(new_p * len_in[b] / orig_p), zero beyond (up to out_pitch).  orig_p / new_p are the rates divided
 * by their gcd; taps = torchaudio's _get_sinc_resample_kernel (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99),
 * passed COMPRESSED: k0[p] = first tap of phase p that can be non-zero, taps_c [nt][new_p] with
 * taps_c[i][p] = taps[p][k0[p] + i] (zero padded) -- the window is exactly 0 outside +-6 zero crossings. */
int ts_resample(const float* x, int B, int N_in, int in_pitch, const int32_t* len_in, int orig_p, int new_p,
                const float* taps_c, const int32_t* k0, int nt, int width, float* out, int N_out, int out_pitch,
                void* stream);

/* ---- training step: train()-mode BatchNorm and the backward pass ------------------------- */
/* Reference: BaseCTCModule.training_step (src/thunder/module.py:102-127) through train()-mode QuartznetBlock
 * (src/thunder/quartznet/blocks.py:231-338); nn.BatchNorm1d(eps=1e-3, momentum=0.1) then uses BATCH statistics over all
 * B x T positions (padded frames included).  Forward reuses ts_dw_conv / ts_pw_gemm (unfolded weights, no shift). */
/* stats[b, c] = (sum_t z, sum_t z^2) over t < T of bf16 rows; the caller sums over b */
int ts_row_stats(const void* z, int B, int C, int T, int pitch, float* stats, void* stream);
/* y = act(z * scale[c] + shift[c] (+ zr * scale_r[c] + shift_r[c])): BatchNorm affine (+ residual branch) + ReLU,
 * frames t >= lens[b] and the pad stored as zero (quartznet/blocks.py:222-228,332-337) */
int ts_bn_apply(const void* z, const float* scale, const float* shift, const void* zr, const float* scale_r,
                const float* shift_r, int B, int C, int T, int pitch, const int32_t* lens, int relu, void* y,
                void* stream);
/* sums[b, c] = (sum dym, sum dym*z, sum dym*zr) with dym = dy * (y > 0) when relu: the reductions of BatchNorm backward.
 * y may be NULL when relu: the mask is then rebuilt from z (and zr) with the forward scale / shift vectors exactly as
 * ts_bn_apply evaluated it (saves reading y); mask_* are ignored otherwise */
int ts_bn_bwd_reduce(const void* dy, const void* y, const void* z, const void* zr, int B, int C, int T, int pitch,
                     int relu, float* sums, const float* mask_scale, const float* mask_shift, const float* mask_scale_r,
                     const float* mask_shift_r, void* stream);
/* dz = coef[c,0] dym + coef[c,1] z + coef[c,2]  (and the same for the residual branch): BatchNorm + ReLU backward;
 * y / mask_* as in ts_bn_bwd_reduce */
int ts_bn_bwd_apply(const void* dy, const void* y, const void* z, const void* zr, const float* coef,
                    const float* coef_r, int B, int C, int T, int pitch, int relu, void* dz, void* dzr,
                    const float* mask_scale, const float* mask_shift, const float* mask_scale_r,
                    const float* mask_shift_r, void* stream);
/* pointwise-conv weight gradient dW[co, ci] = sum_{b,t} dz[b, co, t] a[b, ci, t] on the tensor cores;
 * part [nsplit, Cout, Cin] f32 partial sums over even shares of the B * ceil(T/64) reduction chunks */
int ts_pw_wgrad(const void* dz, int dz_pitch, const void* a, int a_pitch, int B, int Cout, int Cin, int T, int nsplit,
                float* part, void* stream);
/* out[i] = sum_s part[s, i] in ascending s (deterministic); n a multiple of 4, 16-byte aligned pointers */
int ts_pw_wgrad_reduce(const float* part, int nsplit, long long n, float* out, void* stream);
/* depthwise weight gradient: part[chunk, c, k] = sum_{b in chunk, t'} da[b, c, t'] xm[b, c, t' S + k D - P] with xm = x
 * masked to len_in (MaskedConv1d).  flags & TS_DW_INPUT_PREMASKED: both row tensors are already zero beyond the utterance
 * length and in the pad up to the pitch (true for every producer of this library); with stride 1, equal pitches and
 * bchunk >= B (a single chunk) this selects the tensor-core kernel (dwwgrad_mma.cu). */
int ts_dw_wgrad(const void* da, int T_out, int pitch_out, const void* x, int T_in, int pitch_in, const int32_t* len_in,
                int B, int C, int K, int S, int D, int P, int bchunk, int flags, float* part, void* stream);

/* Transpose of ts_gather_rows (backward of strided layers): dst[b,c,t] = src[b,c,t/S] where S divides t and t/S < T_src, else
 * 0, for t < T_dst (zero up to pitch_dst); `accumulate` adds onto the existing dst instead.  bf16 rows. */
int ts_scatter_rows(const void* src, int B, int C, int T_src, int pitch_src, int S, void* dst, int T_dst, int pitch_dst,
                    int accumulate, void* stream);
/* SqueezeExcite variants for Citrinet training (citrinet/blocks.py:48-83,177-197): the main branch of a block's last
 * sub-block is u = BN(z) scaled by gate[b, c] = sigmoid(W2 relu(W1 mean_t u)) before the residual add and ReLU.
 *   ts_bn_apply_se:      y = act(gate[b,c] * (z*scale+shift) (+ zr*scale_r+shift_r)), tail zeroed
 *   ts_bn_bwd_reduce_se: sums [B, C, 3] like ts_bn_bwd_reduce, the ReLU mask rebuilt WITH the gate
 *   ts_bn_bwd_apply_se:  dz = a * (dym * gate[b,c] + addc[b,c]) + b * z + c0 (addc = d loss / d mean_t(u) / T, every frame);
 *                        dzr = a_r * dym + b_r * zr + c0_r. */
int ts_bn_apply_se(const void* z, const float* scale, const float* shift, const void* zr, const float* scale_r,
                   const float* shift_r, const float* gate, int B, int C, int T, int pitch, const int32_t* lens, int relu,
                   void* y, void* stream);
int ts_bn_bwd_reduce_se(const void* dy, const void* z, const void* zr, int B, int C, int T, int pitch, int relu, float* sums,
                        const float* mask_scale, const float* mask_shift, const float* mask_scale_r,
                        const float* mask_shift_r, const float* gate, void* stream);
int ts_bn_bwd_apply_se(const void* dy, const void* z, const void* zr, const float* coef, const float* coef_r, int B, int C,
                       int T, int pitch, int relu, void* dz, void* dzr, const float* mask_scale, const float* mask_shift,
                       const float* mask_scale_r, const float* mask_shift_r, const float* gate, const float* addc,
                       void* stream);
/* torch.optim.AdamW's update for every parameter in ONE launch (BaseCTCModule's default optimizer, module.py:32,129-140).
 * grad / exp_avg / exp_avg_sq: flat f32 buffers in the same element order; table: device array of n_entries rows of 4 int64
 * {param pointer, offset in the flat buffers, elements, first tile} with 1024-element tiles.  bias_correction{1,2} =
 * 1 - beta^step (the caller counts steps). */
int ts_adamw(const long long* table, int n_entries, long long total_tiles, const float* grad, float* exp_avg,
             float* exp_avg_sq, float lr, float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
             float bias_correction2, void* stream);
/* Prepares every weight operand of a training step from the fp32 master weights in ONE launch.  `table` is a device array
 * of n_entries rows of 8 int64: {src, dst, dstT, rows, cols, ldT, kind, first_tile}; tiles are 32 x 32 elements and
 * first_tile is the running sum of ceil(rows/32) * ceil(cols/32).  kind 0 (pointwise / decoder weight [rows, cols] f32):
 * dst = bf16 copy, dstT = bf16 transpose with leading dimension ldT.  kind 1 (depthwise taps [rows, cols] f32): dst = the
 * taps rounded to bf16 (stored f32), dstT = the same flipped along cols (the taps of the input-gradient convolution). */
int ts_prep_weights(const long long* table, int n_entries, long long total_tiles, void* stream);
/* Training forward of a pointwise conv: out = W x as bf16 rows (no shift / activation / mask: x is zero beyond the
 * utterance lengths) on the CTA-pair tcgen05 kernel, with the BatchNorm partial sums of the STORED values produced by
 * the epilogue: stats [B, Cout, slots, 2] = (sum, sum of squares) per 128-frame block, slots = 2 * ceil(out_pitch / 256);
 * every slot is written exactly once (deterministic).  TS_ERR_UNSUPPORTED when Cout <= 128 (use ts_pw_gemm +
 * ts_row_stats). */
int ts_pw_gemm_stats(const void* w, const void* x, int cin, int x_pitch, int B, int Cout, int T, void* out, int out_pitch,
                     float* stats, int slots, void* stream);
/* BatchNorm train()-mode statistics from partial sums part [NB, C, slots, 2] (ts_row_stats: slots = 1, NB = B):
 * mean, biased variance over n = B*T positions -> scale = gamma * inv, shift = beta - mean * scale, mean, inv = rsqrt(var +
 * eps); running_mean / running_var (nullable) are updated in place with `momentum` and the UNBIASED variance, as
 * nn.BatchNorm1d does (quartznet/blocks.py:222-228).  Everything on the device: no host round trip, graph-capturable. */
int ts_bn_finalize(const float* part, int NB, int slots, int C, double n, const float* gamma, const float* beta, float eps,
                   float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                   float* inv, void* stream);
/* Fused ts_bn_finalize + ts_bn_apply (one CTA per channel reduces the partial sums, then streams the B rows):
 * y = act(BN(z) [+ BN_r(zr)]) with train()-mode batch statistics over n = B * T positions; stats_out [4, C] = (scale,
 * shift, mean, inv) per branch for the backward pass; running statistics (nullable) updated in place. */
int ts_bn_apply_fused(const void* z, const float* part, int NB, int slots, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float* stats_out, const void* zr, const float* part_r,
                      int NB_r, int slots_r, const float* gamma_r, const float* beta_r, float* running_mean_r,
                      float* running_var_r, float* stats_out_r, float eps, float momentum, int B, int C, int T, int pitch,
                      const int32_t* lens, int relu, void* y, void* stream);
/* Fused ts_bn_bwd_coef + ts_bn_bwd_apply: from sums [B, C, 3] (ts_bn_bwd_reduce) and the forward stats [4, C], writes
 * dgamma / dbeta [C] and dz (dzr) = a dym + b z + c0 with the ReLU mask rebuilt from z (zr). */
int ts_bn_bwd_apply_fused(const void* dy, const void* z, const float* gamma, const float* stats, float* dgamma,
                          float* dbeta, void* dz, const void* zr, const float* gamma_r, const float* stats_r,
                          float* dgamma_r, float* dbeta_r, void* dzr, const float* sums, int B, int C, int T, int pitch,
                          int relu, void* stream);
/* BatchNorm backward coefficients from the partial sums of ts_bn_bwd_reduce (part [NB, C, 3]; `which` = 1 for the main
 * branch (sum dym*z), 2 for the residual branch (sum dym*zr)): dbeta = sum dym, dgamma = inv (sum dym*z - mean sum dym),
 * coef[c] = (a, b, c0) with dz = a dym + b z + c0 (see ts_bn_bwd_apply) */
int ts_bn_bwd_coef(const float* part, int NB, int C, int which, double n, const float* gamma, const float* mean,
                   const float* inv, float* dgamma, float* dbeta, float* coef, void* stream);
/* CTC loss with fused log-softmax and its gradient w.r.t. the logits (calculate_ctc, src/thunder/ctc_loss.py:15-47:
 * log_softmax -> F.ctc_loss(blank, reduction="mean", zero_infinity=True)).  logits: f32 rows [B, V, pitch] (pitch >= T);
 * in_len: int32 [B] encoder output lengths; targets int64 [B, Lmax], tgt_len int64 [B].  loss[b] = nll_b / max(tgt_len_b,
 * 1) (0 when infinite); the reference's loss is mean_b loss[b].  grad: bf16 rows [B, Vp, grad_pitch] (Vp >= V; rows >= V
 * untouched; grad_pitch a multiple of 8; frames >= T zeroed) =
 * gscale * d mean_b(loss) / d logits.  scratch: >= 2*B*T*Sp + B*T + B floats with Sp = round_up(2*Lmax+1, 4). */
int ts_ctc_loss(const float* logits, int B, int V, int T, int pitch, const int32_t* in_len, const int64_t* targets,
                int Lmax, const int64_t* tgt_len, int blank, float gscale, float* scratch, long long scratch_floats,
                float* loss, void* grad, int Vp, int grad_pitch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* THUNDER_B200_H_ */
