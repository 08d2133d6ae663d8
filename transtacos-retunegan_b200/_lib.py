"""ctypes binding of libspectral_b200.so (the C ABI in include/spectral_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# SB200_LIB: kernel-experiment builds (tools/variants.py); the default is the in-tree library
LIB_PATH = os.environ.get("SB200_LIB") or os.path.join(_HERE, "libspectral_b200.so")


class Config(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("win_length", C.c_int32),
                ("hop_length", C.c_int32), ("n_mel", C.c_int32), ("fmin", C.c_float), ("fmax", C.c_float),
                ("mel_htk", C.c_int32), ("window", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("B", C.c_int32), ("len", C.c_int64), ("stride", C.c_int64), ("sig_off", C.c_void_p),
                ("sig_len", C.c_void_p), ("frame_off", C.c_void_p), ("item_off", C.c_void_p),
                ("total_frames", C.c_int64), ("total_items", C.c_int64), ("total_samples", C.c_int64)]


class Scale(C.Structure):
    _fields_ = [("log", C.c_int32), ("a", C.c_float), ("b", C.c_float), ("floor", C.c_float)]


RAW = Scale(0, 1.0, 0.0, 0.0)

MAX_PEERS = 8


class PeerReduce(C.Structure):
    """sb200_peer_reduce: exchange buffers of the ranks of one box (in-kernel loss reduction under DDP)."""
    _fields_ = [("peer", C.c_void_p * MAX_PEERS), ("rank", C.c_int32), ("world", C.c_int32), ("loss_global", C.c_void_p)]


# name -> (restype, argtypes); every symbol declared in include/spectral_b200.h
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "sb200_version": (C.c_char_p, []),
    "sb200_last_error_string": (C.c_char_p, []),
    "sb200_launch_count": (_I64, []),
    "sb200_plan_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "sb200_plan_destroy": (C.c_int, [_P]),
    "sb200_plan_frames_per_pass": (C.c_int, [_P]),
    "sb200_plan_mel_basis_host": (C.c_int, [_P, _P]),
    "sb200_plan_window_host": (C.c_int, [_P, _P]),
    "sb200_stft_features": (C.c_int, [_P, _P, C.POINTER(Batch), _F, Scale, Scale, _P, _P, _P, _P]),
    "sb200_mel_project": (C.c_int, [_P, _P, _I64, Scale, _P, _P]),
    "sb200_mel_to_linear": (C.c_int, [_P, _P, _I64, _P, _P]),
    "sb200_spec_to_amplitude": (C.c_int, [_P, _I64, _I32, _F, _F, _F, _F, _P, _P]),
    "sb200_frame_stats": (C.c_int, [_P, C.POINTER(Batch), _I32, _I32, _P, _P, _P]),
    "sb200_trim_bounds": (C.c_int, [_P, _P, _I64, _I32, _F, _P, _P]),
    "sb200_yin": (C.c_int, [_P, C.POINTER(Batch), _I32, _F, _F, _I32, _I32, _F, _P, _P]),
    "sb200_pool_loss_workspace_bytes": (_I64, []),
    "sb200_pool_loss": (C.c_int, [_P, _P, _I32, _I64, _I32, _I32, _P, _P, _P, _P]),
    "sb200_preemphasis": (C.c_int, [_P, C.POINTER(Batch), _F, _P, _P]),
    "sb200_inv_preemphasis": (C.c_int, [_P, C.POINTER(Batch), _F, _P, _P]),
    "sb200_griffinlim_workspace_bytes": (_I64, [_P, _I64, _I32, _I32]),
    "sb200_istft": (C.c_int, [_P, _P, C.POINTER(Batch), _I64, _P, _P, _P]),
    "sb200_griffinlim": (C.c_int, [_P, _P, _P, C.POINTER(Batch), _I64, _I32, _F, _I32, _F, _P, _P, _P]),
    "sb200_stft_smp_workspace_bytes": (_I64, [_P, _I32, _I64]),
    "sb200_stft_smp_forward": (C.c_int, [_P, _P, _I32, _I64, _P, _P, _P, _P]),
    "sb200_stft_smp_backward": (C.c_int, [_P, _P, _I32, _I64, _P, _P, _P, _P, _P, _P]),
    "sb200_mstft_saved_bytes": (_I64, [C.POINTER(_P), _I32, _I32, _I64]),
    "sb200_mstft_workspace_bytes": (_I64, [C.POINTER(_P), _I32, _I32, _I64]),
    "sb200_mstft_forward": (C.c_int, [C.POINTER(_P), _I32, _P, _P, _I32, _I64, _I32, _P, C.POINTER(_P),
                                      C.POINTER(_P), _P, _P, _P]),
    "sb200_mstft_loss_and_grad": (C.c_int, [C.POINTER(_P), _I32, _P, _P, _I32, _I64, _P, _P, _P, _P]),
    "sb200_mstft_backward": (C.c_int, [C.POINTER(_P), _I32, _P, _I32, _I64, _I32, _P, C.POINTER(_P), _P, _P, _P, _P]),
    "sb200_peer_buffer_bytes": (_I64, []),
    "sb200_peer_buffer_create": (C.c_int, [C.POINTER(_P), _P]),
    "sb200_peer_buffer_open": (C.c_int, [_P, C.POINTER(_P)]),
    "sb200_peer_buffer_close": (C.c_int, [_P]),
    "sb200_peer_buffer_destroy": (C.c_int, [_P]),
    "sb200_mstft_forward_ddp": (C.c_int, [C.POINTER(_P), _I32, _P, _P, _I32, _I64, _I32, _P, C.POINTER(_P),
                                          C.POINTER(_P), _P, _P, C.POINTER(PeerReduce), _P]),
    "sb200_mstft_loss_and_grad_ddp": (C.c_int, [C.POINTER(_P), _I32, _P, _P, _I32, _I64, _P, _P, _P, C.POINTER(PeerReduce), _P]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  This package has no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    """Map C status codes to Python exceptions (SURVEY.md 8b error conventions)."""
    if rc == 0:
        return
    msg = load().sb200_last_error_string().decode() or what
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def launch_count() -> int:
    return int(load().sb200_launch_count())
