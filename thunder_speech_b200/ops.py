"""``torch.library`` custom ops (namespace ``thunder_b200``) over the C ABI.

Each op takes/returns torch tensors, allocates outputs with torch (device memory and streams are
torch's job -- plumbing), and passes raw device pointers plus the CURRENT CUDA stream to
``libthunder_b200.so``.  Ops are registered for CUDA only; calling them with CPU tensors raises --
there is no CPU fallback.  Fake (meta) implementations give shapes to ``torch.jit.script`` /
``torch.compile`` / ``to_torchscript``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

NS = "thunder_b200"

#: when set to a list, every kernel-launching op appends (kernel name, meta dict, start event, end event);
#: used by bench.py's per-kernel roofline pass (CUDA events on the launching stream).
PROFILE = None


class _timed:
    def __init__(self, name: str, **meta):
        self.name, self.meta = name, meta

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.name, self.meta, self.e0, e1))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("thunder_b200 ops run on CUDA (sm_100a) tensors only; there is no CPU fallback")


def _ptr(t: Tensor) -> int:
    return t.data_ptr()


#: the two 16-bit row formats: bf16 (default) and IEEE fp16 ("half rows": 11-bit mantissa, saturating conversions)
ROW_DTYPES = (torch.bfloat16, torch.float16)


def _row_tag(t: Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return _lib.TS_BF16
    if t.dtype == torch.float16:
        return _lib.TS_F16
    raise TypeError(f"activation rows must be bfloat16 or float16, got {t.dtype}")


def _f16_flag(t: Tensor) -> int:
    return _lib.TS_ROWS_F16 if _row_tag(t) == _lib.TS_F16 else 0


# ------------------------------------------------------------------------------------------- features
@torch.library.custom_op(f"{NS}::filterbank", mutates_args=())
def filterbank(audio: Tensor, lengths: Tensor, window_full: Tensor, twiddle: Tensor, mel_start: Tensor,
               mel_count: Tensor, mel_off: Tensor, mel_w: Tensor, hop: int, preemph: float, win_lo: int,
               win_hi: int, div_guard: float, out_bf16_pitch: int, out_f16: bool = False, dither: float = 0.0,
               seed: int = 0, seed_state: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """``FilterbankFeatures`` (src/thunder/quartznet/transform.py:258-321); ``dither != 0`` adds the train()-mode
    ``DitherAudio`` noise inside the feature kernel (counter-based normals keyed by ``seed`` XOR the device-resident
    ``seed_state`` (int64 ``[1]``), which the caller advances per step -- graph replays then draw fresh noise).

    ``out_bf16_pitch == 0``: returns ``(features[B,nfilt,F] f32, feature_lengths[B] i64)`` exactly like the
    reference.  ``out_bf16_pitch > 0``: features are emitted as 16-bit padded rows ``[B,nfilt,pitch]`` for the
    encoder kernels (bf16, or fp16 with ``out_f16``)."""
    _need_cuda(audio, lengths, window_full, twiddle, mel_start, mel_count, mel_off, mel_w)
    if audio.dim() != 2:
        raise ValueError("audio must be [batch, time]")
    audio = audio.contiguous().float()
    B, N = audio.shape
    n_fft = window_full.numel()
    nfilt = mel_start.numel()
    F = 1 + N // hop
    lens64 = lengths.to(torch.int64).contiguous()
    L = _lib.lib()
    logmel = torch.empty((B, nfilt, F), device=audio.device, dtype=torch.float32)
    if dither != 0.0:
        _lib.check(L.ts_logmel_dither(_ptr(audio), B, N, n_fft, hop, preemph, _ptr(window_full), win_lo, win_hi,
                                      _ptr(twiddle), _ptr(mel_start), _ptr(mel_count), _ptr(mel_off), _ptr(mel_w), nfilt,
                                      mel_w.numel(), _ptr(logmel), float(dither), int(seed) & (2 ** 64 - 1),
                                      _ptr(seed_state) if seed_state is not None else None, _stream()),
                   "ts_logmel_dither")
    else:
        _lib.check(L.ts_logmel(_ptr(audio), B, N, n_fft, hop, preemph, _ptr(window_full), win_lo, win_hi,
                               _ptr(twiddle), _ptr(mel_start), _ptr(mel_count), _ptr(mel_off), _ptr(mel_w), nfilt,
                               mel_w.numel(), _ptr(logmel), _stream()), "ts_logmel")
    seq = torch.empty((B,), device=audio.device, dtype=torch.int64)
    if out_bf16_pitch > 0:
        out = torch.empty((B, nfilt, out_bf16_pitch), device=audio.device,
                          dtype=torch.float16 if out_f16 else torch.bfloat16)
        dt, pitch = (_lib.TS_F16 if out_f16 else _lib.TS_BF16), out_bf16_pitch
    else:
        out = torch.empty((B, nfilt, F), device=audio.device, dtype=torch.float32)
        dt, pitch = _lib.TS_F32, F
    _lib.check(L.ts_feature_normalize(_ptr(logmel), _ptr(lens64), B, nfilt, F, hop, div_guard, _ptr(out), dt,
                                      pitch, _ptr(seq), _stream()), "ts_feature_normalize")
    return out, seq


@filterbank.register_fake
def _(audio, lengths, window_full, twiddle, mel_start, mel_count, mel_off, mel_w, hop, preemph, win_lo, win_hi,
      div_guard, out_bf16_pitch, out_f16=False, dither=0.0, seed=0, seed_state=None):
    B, N = audio.shape
    nfilt = mel_start.numel()
    F = 1 + N // hop
    if out_bf16_pitch > 0:
        out = audio.new_empty((B, nfilt, out_bf16_pitch), dtype=torch.float16 if out_f16 else torch.bfloat16)
    else:
        out = audio.new_empty((B, nfilt, F), dtype=torch.float32)
    return out, audio.new_empty((B,), dtype=torch.int64)


@torch.library.custom_op(f"{NS}::filterbank_dft", mutates_args=())
def filterbank_dft(audio: Tensor, lengths: Tensor, wplus: Tensor, wminus: Tensor, basis: Tensor, mel_w2: Tensor,
                   mel_adv: Tensor, nfilt: int, hop: int, preemph: float, div_guard: float, out_bf16_pitch: int,
                   out_f16: bool = False) -> Tuple[Tensor, Tensor]:
    """Eval-mode ``FilterbankFeatures`` with the STFT as a DFT-matrix contraction on the tensor cores (``ts_logmel_dft``,
    n_fft 512 / window support [96, 416) only) and the normaliser fed by the partial sums that kernel emits
    (``ts_feature_normalize_partials``).  Same outputs as :func:`filterbank`."""
    _need_cuda(audio, lengths, wplus, wminus, basis, mel_w2, mel_adv)
    if audio.dim() != 2:
        raise ValueError("audio must be [batch, time]")
    audio = audio.contiguous().float()
    B, N = audio.shape
    F = 1 + N // hop
    lens64 = lengths.to(torch.int64).contiguous()
    L = _lib.lib()
    logmel = torch.empty((B, nfilt, F), device=audio.device, dtype=torch.float32)
    partials = torch.empty((B, (F + 31) // 32, nfilt, 2), device=audio.device, dtype=torch.float32)
    _lib.check(L.ts_logmel_dft(_ptr(audio), B, N, hop, preemph, _ptr(wplus), _ptr(wminus), _ptr(basis), _ptr(mel_w2),
                               _ptr(mel_adv), nfilt, _ptr(logmel), _ptr(partials), _ptr(lens64), _stream()), "ts_logmel_dft")
    seq = torch.empty((B,), device=audio.device, dtype=torch.int64)
    if out_bf16_pitch > 0:
        out = torch.empty((B, nfilt, out_bf16_pitch), device=audio.device,
                          dtype=torch.float16 if out_f16 else torch.bfloat16)
        dt, pitch = (_lib.TS_F16 if out_f16 else _lib.TS_BF16), out_bf16_pitch
    else:
        out = torch.empty((B, nfilt, F), device=audio.device, dtype=torch.float32)
        dt, pitch = _lib.TS_F32, F
    _lib.check(L.ts_feature_normalize_partials(_ptr(logmel), _ptr(partials), _ptr(lens64), B, nfilt, F, hop, div_guard,
                                               _ptr(out), dt, pitch, _ptr(seq), _stream()), "ts_feature_normalize_partials")
    return out, seq


@filterbank_dft.register_fake
def _(audio, lengths, wplus, wminus, basis, mel_w2, mel_adv, nfilt, hop, preemph, div_guard, out_bf16_pitch, out_f16=False):
    B, N = audio.shape
    F = 1 + N // hop
    if out_bf16_pitch > 0:
        out = audio.new_empty((B, nfilt, out_bf16_pitch), dtype=torch.float16 if out_f16 else torch.bfloat16)
    else:
        out = audio.new_empty((B, nfilt, F), dtype=torch.float32)
    return out, audio.new_empty((B,), dtype=torch.int64)


# ------------------------------------------------------------------------------------------- layout
def row_pitch(T: int) -> int:
    """Pitch (frames) of a padded bf16 activation row holding T frames (multiple of 64)."""
    return (int(T) + 63) // 64 * 64


@torch.library.custom_op(f"{NS}::pack_rows", mutates_args=())
def pack_rows(x: Tensor, lens: Optional[Tensor] = None, f16: bool = False) -> Tensor:
    """``[B, C, T]`` (f32 / bf16 / fp16, contiguous) -> 16-bit padded rows ``[B, C, row_pitch(T)]`` (bf16, or fp16 with
    ``f16``); pad frames and, when ``lens`` (i32 ``[B]``) is given, frames ``t >= lens[b]`` are zero
    (``MaskedConv1d.mask_fill``)."""
    _need_cuda(x)
    if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        x = x.float()
    x = x.contiguous()
    B, C, T = x.shape
    out = torch.empty((B, C, row_pitch(T)), device=x.device, dtype=torch.float16 if f16 else torch.bfloat16)
    dt = _lib.TS_F32 if x.dtype == torch.float32 else _row_tag(x)
    _lib.check(_lib.lib().ts_pack_rows(_ptr(x), dt, B, C, T, _ptr(lens) if lens is not None else None, _ptr(out),
                                       _row_tag(out), out.shape[2], _stream()), "ts_pack_rows")
    return out


@pack_rows.register_fake
def _(x, lens=None, f16=False):
    B, C, T = x.shape
    return x.new_empty((B, C, row_pitch(T)), dtype=torch.float16 if f16 else torch.bfloat16)


@torch.library.custom_op(f"{NS}::unpack_rows", mutates_args=())
def unpack_rows(x: Tensor, T: int) -> Tensor:
    """bf16 padded rows ``[B, C, pitch]`` -> contiguous f32 ``[B, C, T]``."""
    _need_cuda(x)
    B, C, pitch = x.shape
    out = torch.empty((B, C, T), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().ts_unpack_rows(_ptr(x), _row_tag(x), pitch, B, C, T, _ptr(out), _stream()), "ts_unpack_rows")
    return out


@unpack_rows.register_fake
def _(x, T):
    return x.new_empty((x.shape[0], x.shape[1], T), dtype=torch.float32)


def lengths_i32(lengths: Tensor) -> Tensor:
    """Any numeric lengths tensor -> int32 device tensor (``.type(torch.long)`` truncation first, like
    ``lengths_to_mask``, src/thunder/blocks.py:166)."""
    return lengths.to(torch.int64).to(torch.int32).contiguous()


# ------------------------------------------------------------------------------------------- depthwise
@torch.library.custom_op(f"{NS}::dw_conv", mutates_args=())
def dw_conv(x: Tensor, T_in: int, weight: Tensor, stride: int, dilation: int, padding: int,
            lens: Optional[Tensor], premasked: bool = False) -> Tensor:
    """Masked depthwise conv over bf16 rows.  ``weight`` is ``[C, K]`` f32; ``lens`` i32 ``[B]`` input lengths.
    ``premasked``: the caller guarantees ``x`` is already zero beyond ``lens`` (enables the TMA-fed kernel)."""
    _need_cuda(x, weight)
    B, C, pitch = x.shape
    K = weight.shape[1]
    T_out = (T_in + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    out = torch.empty((B, C, row_pitch(max(T_out, 1))), device=x.device, dtype=x.dtype)
    with _timed("dw_conv", bytes=2 * B * C * (T_in + T_out), flops=2 * B * C * T_out * K, K=K, C=C, T=T_out):
        _lib.check(_lib.lib().ts_dw_conv(_ptr(x), B, C, T_in, pitch, _ptr(weight), K, stride, dilation, padding,
                                         _ptr(lens) if lens is not None else None,
                                         (_lib.TS_DW_INPUT_PREMASKED if premasked else 0) | _f16_flag(x), _ptr(out),
                                         out.shape[2],
                                         _stream()), "ts_dw_conv")
    return out


@dw_conv.register_fake
def _(x, T_in, weight, stride, dilation, padding, lens, premasked=False):
    K = weight.shape[1]
    T_out = (T_in + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    return x.new_empty((x.shape[0], x.shape[1], row_pitch(max(T_out, 1))))


# ------------------------------------------------------------------------------------------- pointwise
@torch.library.custom_op(f"{NS}::pw_gemm", mutates_args=("pool",))
def pw_gemm(w0: Tensor, x0: Tensor, w1: Optional[Tensor], x1: Optional[Tensor], T: int, shift: Optional[Tensor],
            lens: Optional[Tensor], out_f32: bool, relu: bool, pool: Optional[Tensor], se_scale: Optional[Tensor],
            y1: Optional[Tensor], const_weights: bool = False) -> Tensor:
    """tcgen05 pointwise GEMM with fused BN-shift / residual segment / ReLU / tail mask / SE epilogues.
    ``const_weights``: ``w0`` / ``w1`` were complete before anything still in flight on the stream (folded inference plans,
    built and synchronised once) -- lets the weight-stationary kernel fetch them under the previous kernel's tail.
    ``w*`` bf16 ``[Cout, Cin]`` (BN scale folded in), ``x*`` bf16 rows ``[B, Cin, pitch]``.  ``pool`` (SqueezeExcite
    squeeze) is a zeroed int64 ``[B, Cout]``: fixed-point sums in units of 2^-32 (``se_pool_to_float``), accumulated with
    integer atomics so that the result does not depend on tile order."""
    _need_cuda(w0, x0)
    if pool is not None and (pool.dtype != torch.int64 or tuple(pool.shape) != (x0.shape[0], w0.shape[0])):
        raise TypeError("pw_gemm: pool must be an int64 [B, Cout] tensor (fixed-point sums, see se_pool_to_float)")
    B, cin0, p0 = x0.shape
    Cout = w0.shape[0]
    if out_f32:
        out = torch.empty((B, Cout, T), device=x0.device, dtype=torch.float32)
        dt, pitch = _lib.TS_F32, T
    else:
        out = torch.empty((B, Cout, row_pitch(T)), device=x0.device, dtype=x0.dtype)
        dt, pitch = _row_tag(x0), out.shape[2]
    for t in (w0, w1, x1, y1):
        if t is not None and t.dtype != x0.dtype:
            raise TypeError(f"pw_gemm: operands must share the row format of x0 ({x0.dtype}), got {t.dtype}")
    cin1 = x1.shape[1] if x1 is not None else 0
    p1 = x1.shape[2] if x1 is not None else 0

    def P(t):
        return _ptr(t) if t is not None else None

    kin = cin0 + cin1
    nbytes = 2 * B * T * kin + 2 * Cout * kin + (4 if out_f32 else 2) * B * T * Cout + (2 * B * T * Cout if y1 is not None else 0)
    with _timed("pw_gemm", bytes=nbytes, flops=2 * B * T * kin * Cout, K=kin, C=Cout, T=T):
        _lib.check(_lib.lib().ts_pw_gemm(_ptr(w0), _ptr(x0), cin0, p0, P(w1), P(x1), cin1, p1, B, Cout, T, P(shift),
                                         P(lens), _ptr(out), dt, pitch,
                                         (_lib.TS_PW_RELU if relu else 0) | _f16_flag(x0)
                                         | (_lib.TS_PW_CONST_WEIGHTS if const_weights else 0), P(pool), P(se_scale), P(y1),
                                         y1.shape[2] if y1 is not None else 0, _stream()), "ts_pw_gemm")
    return out


@pw_gemm.register_fake
def _(w0, x0, w1, x1, T, shift, lens, out_f32, relu, pool, se_scale, y1, const_weights=False):
    B, Cout = x0.shape[0], w0.shape[0]
    if out_f32:
        return x0.new_empty((B, Cout, T), dtype=torch.float32)
    return x0.new_empty((B, Cout, row_pitch(T)))


# ------------------------------------------------------------------------------------------- SE / CTC
@torch.library.custom_op(f"{NS}::se_fc", mutates_args=())
def se_fc(pool: Tensor, T: int, w1: Tensor, w2: Tensor) -> Tensor:
    """``sigmoid(W2 relu(W1 pool/T))`` -> gate ``[B, C]`` f32 (citrinet/blocks.py:63-83).  ``pool`` holds the sums over all T
    frames: int64 fixed point as ``pw_gemm`` accumulates them, or plain float32."""
    _need_cuda(pool, w1, w2)
    if pool.dtype not in (torch.int64, torch.float32):
        raise TypeError("se_fc: pool must be int64 (fixed point) or float32")
    B, C = pool.shape
    gate = torch.empty((B, C), device=pool.device, dtype=torch.float32)
    hid = torch.empty((B, w1.shape[0]), device=pool.device, dtype=torch.float32)
    dt = _lib.TS_FIX32 if pool.dtype == torch.int64 else _lib.TS_F32
    _lib.check(_lib.lib().ts_se_fc(_ptr(pool), dt, B, C, w1.shape[0], T, _ptr(w1), _ptr(w2), _ptr(hid), _ptr(gate),
                                   _stream()), "ts_se_fc")
    return gate


@se_fc.register_fake
def _(pool, T, w1, w2):
    return pool.new_empty(pool.shape, dtype=torch.float32)


def se_pool_to_float(pool: Tensor) -> Tensor:
    """The fixed-point SqueezeExcite sums of ``pw_gemm`` (int64, units of 2^-32) as float32."""
    return (pool.double() * 2.0 ** -32).float()


@torch.library.custom_op(f"{NS}::ctc_greedy", mutates_args=())
def ctc_greedy(logits: Tensor, T: int, drop_blank: int) -> Tuple[Tensor, Tensor, Tensor]:
    """``argmax(1)`` + per-row ``unique_consecutive``.  ``logits`` is f32 ``[B, V, T]`` or bf16 rows
    ``[B, V, pitch]``.  Returns ``(ids[B,T] i64, collapsed[B,T] i64 padded with -1, counts[B] i32)``."""
    _need_cuda(logits)
    logits = logits.contiguous()
    B, V, pitch = logits.shape
    dt = _lib.TS_F32 if logits.dtype == torch.float32 else _lib.TS_BF16
    ids = torch.empty((B, T), device=logits.device, dtype=torch.int64)
    col = torch.empty((B, T), device=logits.device, dtype=torch.int64)
    cnt = torch.empty((B,), device=logits.device, dtype=torch.int32)
    _lib.check(_lib.lib().ts_ctc_greedy(_ptr(logits), dt, B, V, T, pitch, _ptr(ids), _ptr(col), _ptr(cnt),
                                        drop_blank, _stream()), "ts_ctc_greedy")
    return ids, col, cnt


@ctc_greedy.register_fake
def _(logits, T, drop_blank):
    B = logits.shape[0]
    return (logits.new_empty((B, T), dtype=torch.int64), logits.new_empty((B, T), dtype=torch.int64),
            logits.new_empty((B,), dtype=torch.int32))


@torch.library.custom_op(f"{NS}::se_apply", mutates_args=())
def se_apply(y1: Tensor, gate: Tensor, lens: Optional[Tensor], relu: bool) -> Tensor:
    """``relu(gate[b,c] * y1)`` over bf16 rows (SE scale for blocks without a residual branch)."""
    _need_cuda(y1, gate)
    B, C, pitch = y1.shape
    out = torch.empty_like(y1)
    _lib.check(_lib.lib().ts_se_apply(_ptr(y1), _ptr(gate), B, C, pitch, _ptr(lens) if lens is not None else None,
                                      (_lib.TS_PW_RELU if relu else 0) | _f16_flag(y1), _ptr(out), _stream()),
               "ts_se_apply")
    return out


@se_apply.register_fake
def _(y1, gate, lens, relu):
    return torch.empty_like(y1)


@torch.library.custom_op(f"{NS}::conv_lengths", mutates_args=())
def conv_lengths(lens: Tensor, kernel_size: int, stride: int, dilation: int, padding: int) -> Tensor:
    """``MaskedConv1d.get_seq_len`` (quartznet/blocks.py:142-156) on an i32 device tensor."""
    _need_cuda(lens)
    out = torch.empty_like(lens)
    _lib.check(_lib.lib().ts_conv_lengths(_ptr(lens), _ptr(out), lens.numel(), kernel_size, stride, dilation,
                                          padding, _stream()), "ts_conv_lengths")
    return out


@conv_lengths.register_fake
def _(lens, kernel_size, stride, dilation, padding):
    return torch.empty_like(lens)


@torch.library.custom_op(f"{NS}::im2col_rows", mutates_args=())
def im2col_rows(x: Tensor, T_in: int, K: int, stride: int, dilation: int, padding: int, lens: Optional[Tensor]) -> Tensor:
    """``y[b, c*K + k, t] = x[b, c, t*stride + k*dilation - padding]`` (zero outside the utterance / the conv padding) over
    16-bit rows: the operand that turns a non-separable ``MaskedConv1d`` with ``kernel_size > 1`` into one ``pw_gemm``."""
    _need_cuda(x)
    B, C, pitch = x.shape
    T_out = (T_in + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    CK = (C * K + 7) // 8 * 8          # the GEMM's K extent in whole 16-byte groups (TMA row pitch); extra rows are zero
    out = torch.empty((B, CK, row_pitch(max(T_out, 1))), device=x.device, dtype=x.dtype)
    if CK != C * K:
        out[:, C * K:].zero_()
    _lib.check(_lib.lib().ts_im2col_rows(_ptr(x), B, C, T_in, pitch, K, stride, dilation, padding,
                                         _ptr(lens) if lens is not None else None, _ptr(out), CK, out.shape[2], _stream()),
               "ts_im2col_rows")
    return out


@im2col_rows.register_fake
def _(x, T_in, K, stride, dilation, padding, lens):
    T_out = (T_in + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    return x.new_empty((x.shape[0], (x.shape[1] * K + 7) // 8 * 8, row_pitch(max(T_out, 1))))


@torch.library.custom_op(f"{NS}::gather_rows", mutates_args=())
def gather_rows(x: Tensor, T_in: int, stride: int, lens: Optional[Tensor]) -> Tensor:
    """``y[b,c,t'] = x[b,c,stride*t']`` over bf16 rows, masked by the INPUT lengths: the input side of a strided
    1x1 residual conv."""
    _need_cuda(x)
    B, C, pitch = x.shape
    T_out = (T_in - 1) // stride + 1
    out = torch.empty((B, C, row_pitch(T_out)), device=x.device, dtype=x.dtype)
    _lib.check(_lib.lib().ts_gather_rows(_ptr(x), B, C, T_in, pitch, stride, _ptr(lens) if lens is not None else None,
                                         _ptr(out), out.shape[2], _stream()), "ts_gather_rows")
    return out


@gather_rows.register_fake
def _(x, T_in, stride, lens):
    return x.new_empty((x.shape[0], x.shape[1], row_pitch((T_in - 1) // stride + 1)))
