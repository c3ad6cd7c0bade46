"""ctypes binding of ``libthunder_b200.so`` (the C ABI declared in ``include/thunder_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C thunder_speech_b200/csrc``).
There is NO fallback: if the shared object is missing or a call fails, a ``RuntimeError`` /
``ValueError`` is raised -- the product path never routes around the CUDA kernels.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libthunder_b200.so")
if os.environ.get("THUNDER_B200_TRACE_BUILD") == "1":     # diagnostics: the build with the phase-trace hooks (`make trace`)
    LIB_PATH = os.path.join(_HERE, "libthunder_b200_trace.so")

TS_OK = 0
TS_ERR_INVALID = -1
TS_ERR_UNSUPPORTED = -2
TS_ERR_NO_DEVICE = -3
TS_F32 = 0
TS_BF16 = 1
TS_I16 = 2
TS_FIX32 = 3
TS_F16 = 4
TS_DW_INPUT_PREMASKED = 1
TS_ROWS_F16 = 2
TS_PW_RELU = 1
TS_PW_CONST_WEIGHTS = 4

_lib = None

# name -> (restype, argtypes); kept in one table so the CPU test-suite can check that every symbol the
# header declares is bound and exported.
SIGNATURES = {
    "ts_version": (c_char_p, []),
    "ts_last_error": (c_char_p, []),
    "ts_launch_count": (c_int64, []),
    "ts_row_pitch": (c_int, [c_int]),
    "ts_set_option": (c_int, [c_char_p, c_int]),
    "ts_trace": (c_int, [c_void_p, c_int]),
    "ts_logmel": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_int, c_void_p,
                          c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ts_logmel_dither": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_uint64,
                                 c_void_p, c_void_p]),
    "ts_logmel_dft": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "ts_feature_normalize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int,
                                     c_int, c_void_p, c_void_p]),
    "ts_feature_normalize_partials": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                              c_int, c_int, c_void_p, c_void_p]),
    "ts_dw_conv": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                           c_int, c_void_p, c_int, c_void_p]),
    "ts_pw_gemm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                           c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                           c_void_p]),
    "ts_se_fc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ts_se_apply": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ts_ctc_greedy": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p]),
    "ts_row_stats": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ts_bn_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                            c_void_p, c_int, c_void_p, c_void_p]),
    "ts_bn_bwd_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ts_bn_bwd_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ts_pw_wgrad": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ts_pw_wgrad_reduce": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "ts_dw_wgrad": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_int, c_int, c_void_p, c_void_p]),
    "ts_spec_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ts_pcm_ingest": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p,
                              c_longlong, c_void_p]),
    "ts_resample": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                            c_void_p, c_int, c_int, c_void_p]),
    "ts_scatter_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ts_bn_apply_se": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                               c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ts_bn_bwd_reduce_se": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ts_bn_bwd_apply_se": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ts_adamw": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float,
                         c_float, c_float, c_float, c_void_p]),
    "ts_prep_weights": (c_int, [c_void_p, c_int, c_longlong, c_void_p]),
    "ts_pw_gemm_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                 c_void_p]),
    "ts_bn_finalize": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ts_bn_apply_fused": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_float, c_float, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ts_bn_bwd_apply_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_int, c_void_p]),
    "ts_bn_bwd_coef": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "ts_ctc_loss": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_float,
                            c_void_p, c_longlong, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ts_gather_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "ts_im2col_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                               c_int, c_void_p]),
    "ts_pack_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ts_unpack_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ts_conv_lengths": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ts_lengths_to_i32": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "ts_lengths_to_i64": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
}


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C thunder_speech_b200/csrc). thunder_speech_b200 has no CPU or eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        # A/B switches for measurements, e.g. THUNDER_B200_OPTIONS="pw_pair=0,dw_tma=1"
        for kv in filter(None, os.environ.get("THUNDER_B200_OPTIONS", "").split(",")):
            k, v = kv.split("=")
            check(handle.ts_set_option(k.strip().encode(), int(v)), "ts_set_option")
    return _lib


def check(rc: int, what: str = "") -> None:
    """Turn a C-ABI return code into the exception the reference would raise for the same misuse."""
    if rc == TS_OK:
        return
    msg = lib().ts_last_error().decode("utf-8", "replace")
    if rc == TS_ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    if rc == TS_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: rc={rc}: {msg}")


def launch_count() -> int:
    return int(lib().ts_launch_count())


def row_pitch(T: int) -> int:
    return int(lib().ts_row_pitch(int(T)))


def set_option(name: str, value: int) -> None:
    check(lib().ts_set_option(name.encode(), int(value)), "ts_set_option")
