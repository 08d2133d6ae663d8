"""Drop-in for the spectral functions of ``retunegan/audio.py`` (same names, arguments, return arity).

    get_mag(y, clamp_low=True) -> ln|STFT| [F,T] f32              retunegan/audio.py:116-120
    get_mel(y, clamp_low=True) -> ln(mel @ |STFT|) [M,T] f32      retunegan/audio.py:123-128
    mag_to_mel(x)              -> mel_basis @ x                   retunegan/audio.py:20-21
    inv_mag(mag, wavlen=None)  -> Griffin-Lim wav (4 it, m=0.7)   retunegan/audio.py:131-147
    get_stft_torch(y, n_fft, win_length, hop_length) -> S, M, P   retunegan/audio.py:150-170

numpy in -> numpy out, torch in -> CUDA float32 tensors out.  Spectrograms are ``[F, T]`` views of
frame-major memory.  Lists / ``[B, L]`` batches are additions over the reference.
"""
from __future__ import annotations

import collections
import math
import threading

import numpy as np
import torch

from . import core
from .config import RETUNEGAN, SpectralConfig
from .transtacos_audio import _frame_feature, _is_np, _split_fm, _to_frame_major, griffin_lim_amplitude, phase_to_frame_major

hp: SpectralConfig = RETUNEGAN
eps = 1e-5
# Seeded initial phase (librosa.griffinlim random_state=int).  rand(F, T) is the first F*T values of ONE MT19937 stream, so a
# single device-resident prefix of that stream serves every T (grow-only, <= a few MB); the [T, F] frame-major layouts made
# from it are kept in a small LRU (a ragged corpus sees hundreds of distinct T: nothing is kept per length beyond that).
_phase_lock = threading.Lock()
_phase_stream = {}                       # (seed, device) -> float32 CUDA prefix of RandomState(seed).rand(n)
_phase_cache = collections.OrderedDict()  # (seed, F, T, device) -> float32 CUDA [T, F]
_PHASE_LRU = 16


def set_hparams(cfg) -> None:
    global hp
    hp = cfg if isinstance(cfg, SpectralConfig) else SpectralConfig.from_hparam(cfg)
    with _phase_lock:
        _phase_cache.clear()
        _phase_stream.clear()


def __getattr__(name):
    if name == "mel_basis":   # retunegan/audio.py:20 -- built lazily here (needs the device library)
        return core.get_plan(hp).mel_basis()
    raise AttributeError(name)


def ln_scale(clamp_low: bool) -> core.Scale:
    """np.log(S.clip(min=eps)) / np.log(S) as a*log2(max(floor, x)) + b."""
    return core.log_scale(math.log(2.0), 0.0, eps if clamp_low else 0.0)


def _features(y, want_mag, want_mel, clamp_low):
    as_np = _is_np(y)
    # hp.mel_scale is honoured by get_mel only (retunegan/audio.py:126); the magnitudes do not depend on the mel tables
    plan = core.get_plan(hp, htk=(want_mel and hp.mel_scale == "htk"))
    batch = core.SignalBatch(plan, y)
    sc = ln_scale(clamp_low)
    mag, mel, _ = core.stft_features(plan, batch, 0.0, sc, sc, want_mag, want_mel)
    single = not isinstance(y, (list, tuple)) and getattr(y, "ndim", 1) == 1
    S = _split_fm(mag, batch.frames, plan.F, as_np, np.float32, single) if want_mag else None
    M = _split_fm(mel, batch.frames, plan.n_mel, as_np, np.float32, single) if want_mel else None
    return S, M


def get_mag(y, clamp_low=True):
    return _features(y, True, False, clamp_low)[0]


def get_mel(y, clamp_low=True):
    return _features(y, False, True, clamp_low)[1]


def get_mag_mel(y, clamp_low=True):
    """Addition: both features from one launch (the reference computes the STFT twice)."""
    return _features(y, True, True, clamp_low)


def get_zcr(y):
    """retunegan/audio.py:98-100: ``librosa.feature.zero_crossing_rate(y, frame_length=win_length, hop_length=hop_length)[0]``."""
    return _frame_feature(y, hp.win_length, hp.hop_length, "zcr")


def get_c0(y):
    """retunegan/audio.py:103-105: ``librosa.feature.rms(y=y, frame_length=win_length, hop_length=hop_length)[0]``."""
    return _frame_feature(y, hp.win_length, hp.hop_length, "rms")


def get_uv(zcr, dyn):
    """retunegan/audio.py:108-113: unvoiced where ``zcr > 0.18 or dyn < 0.03`` (same dtype / container as ``zcr``)."""
    if isinstance(zcr, torch.Tensor):
        return ((zcr > 0.18) | (torch.as_tensor(dyn, device=zcr.device) < 0.03)).to(zcr.dtype)
    zcr = np.asarray(zcr)
    return ((zcr > 0.18) | (np.asarray(dyn) < 0.03)).astype(zcr.dtype)


def mag_to_mel(x):
    """np.dot(mel_basis, x) on whatever it is given, [F,T] -> [M,T] (retunegan/audio.py:21)."""
    plan = core.get_plan(hp)
    if x.shape[0] != plan.F:
        raise ValueError(f"shapes ({plan.n_mel},{plan.F}) and {tuple(x.shape)} not aligned")   # np.dot's complaint
    out = core.mel_project(plan, _to_frame_major(x))
    return out.cpu().numpy().T if isinstance(x, np.ndarray) else out.t()


def _seeded_phase(F: int, T: int) -> torch.Tensor:
    """librosa.griffinlim(random_state=int) draws RandomState(seed).rand(F, T) afresh on every call, so every
    utterance of a given length starts from the same phase: cache it on the device."""
    dev = torch.cuda.current_device()
    key = (hp.randseed, F, T, dev)
    cur = torch.cuda.current_stream()
    with _phase_lock:
        ph = _phase_cache.get(key)
        if ph is not None:
            _phase_cache.move_to_end(key)
            ph.record_stream(cur)      # may have been built on another thread's stream; keep its memory until this use is done
            return ph
        u = _phase_stream.get((hp.randseed, dev))
        if u is None or u.numel() < F * T:
            n = max(F * T, F * 600)            # corpus utterances have <= 524 frames (stats/DataBaker.stats:6-11)
            u = core.to_device_f32(np.random.RandomState(hp.randseed).rand(n))
            _phase_stream[(hp.randseed, dev)] = u
        ph = u[:F * T].view(F, T).t().contiguous()      # element (f, t) = draw f*T + t (C-order rand(F, T)), frame-major
        cur.synchronize()              # published to other threads / streams only once it is complete
        _phase_cache[key] = ph
        while len(_phase_cache) > _PHASE_LRU:
            _phase_cache.popitem(last=False)
    return ph


def _griffinlim(S, wavlen=None, init_phase=None):
    """retunegan/audio.py:131-136 on an amplitude spectrogram S [F,T]."""
    S_fm = core.spec_to_amplitude(_to_frame_major(S), 2, power=hp.gl_power) if hp.gl_power else _to_frame_major(S)
    T = S.shape[1]
    ph = _seeded_phase(S.shape[0], T) if init_phase is None else phase_to_frame_major(init_phase, S.shape[0], T, S_fm.device)
    y = griffin_lim_amplitude(S_fm, T, ph, hp.gl_iters, hp.gl_momentum, 1, wavlen, 0.0, hp)
    return y.cpu().numpy().astype(np.float32) if isinstance(S, np.ndarray) else y


def inv_mag(mag, wavlen=None, init_phase=None):
    """retunegan/audio.py:139-147: exp, prepend a zero DC row if F == n_freq-1, S**gl_power, fast Griffin-Lim."""
    F, T = mag.shape
    x = _to_frame_major(mag)
    S = core.spec_to_amplitude(x, 1, power=hp.gl_power if hp.gl_power else 1.0)
    if F == hp.n_freq - 1:
        S = torch.cat([torch.zeros(T, 1, device=S.device), S], dim=1).contiguous()
    elif F != hp.n_freq:
        raise ValueError(f"expected {hp.n_freq} or {hp.n_freq - 1} frequency rows, got {F}")
    ph = _seeded_phase(hp.n_freq, T) if init_phase is None else phase_to_frame_major(init_phase, hp.n_freq, T, S.device)
    y = griffin_lim_amplitude(S, T, ph, hp.gl_iters, hp.gl_momentum, 1, wavlen, 0.0, hp)
    if wavlen:
        assert y.numel() == wavlen
    return y.cpu().numpy().astype(np.float32) if isinstance(mag, np.ndarray) else y


def inv_mag_batch(mag_fm: torch.Tensor, frames, wavlens=None, init_phase="seeded"):
    """Addition: ``inv_mag`` for a ragged batch in ONE Griffin-Lim call (the corpus path of retunegan/data.py:60-76).

    mag_fm: float32 CUDA ``[sum(frames), n_freq]`` ln-magnitudes, frame-major, utterances back to back (what
    ``core.stft_features`` writes for a list input).  frames: per-utterance frame counts; wavlens: per-utterance output
    lengths (``1 + wavlen // hop == frames``) or None for ``hop * (T - 1)``; ``frames`` may also be a prebuilt
    ``core.FramesBatch`` (then ``wavlens`` is ignored).  init_phase: "seeded" = the reference's
    ``RandomState(randseed).rand(F, T)`` per utterance (host MT19937 draw, cached per length), "device" = on-device
    RNG (throughput mode), or a float32 CUDA ``[sum(frames), n_freq]`` tensor of draws in [0, 1).
    Returns (flat float32 CUDA wav, offsets int64 numpy [B + 1]).
    """
    plan = core.get_plan(hp)
    fb = None
    if isinstance(frames, core.FramesBatch):       # prebuilt descriptor (device offset tables already uploaded)
        fb, frames = frames, frames.frames
    frames = [int(t) for t in frames]
    if mag_fm.shape != (sum(frames), plan.F):
        raise ValueError(f"expected mag_fm of shape {(sum(frames), plan.F)}, got {tuple(mag_fm.shape)}")
    S = core.spec_to_amplitude(mag_fm.contiguous(), 1, power=hp.gl_power if hp.gl_power else 1.0)
    if isinstance(init_phase, str):
        if init_phase == "device":
            ph = torch.rand(S.shape, device=S.device, dtype=torch.float32)
        elif init_phase == "seeded":
            ph = torch.cat([_seeded_phase(plan.F, t) for t in frames])
        else:
            raise ValueError("init_phase must be 'seeded', 'device' or a tensor")
    else:
        ph = init_phase.to(device=S.device, dtype=torch.float32).contiguous()
        if ph.shape != S.shape:
            raise ValueError(f"init_phase must have shape {tuple(S.shape)}")
    if fb is None:
        fb = core.FramesBatch(plan, frames, wavlens, S.device)
    y = core.griffinlim(plan, S, ph, fb, hp.gl_iters, hp.gl_momentum, 1, 0.0)
    return y, fb.out_off


class _StftTorchFn(torch.autograd.Function):
    """S, M, P of get_stft_torch in ONE launch (``stft_smp_kernel``); backward = one launch of the mstft backward kernel on the
    upstream gradients of S, M and P plus the overlap-add (``sb200_stft_smp_backward``): no ATen op on either path."""

    @staticmethod
    def forward(ctx, y, plan):
        lib = core._lib.load()
        dev = core.require_cuda()
        yc = y.detach().to(device=dev, dtype=torch.float32).contiguous()
        B, T = yc.shape
        Tf = 1 + T // plan.hop_length
        S = torch.empty((B, Tf, plan.F), device=dev, dtype=torch.float32)
        P = torch.empty((B, Tf, plan.F), device=dev, dtype=torch.float32)
        M = torch.empty((B, Tf, plan.n_mel), device=dev, dtype=torch.float32)
        core.check(lib.sb200_stft_smp_forward(plan.handle, core.ptr(yc), B, T, core.ptr(S), core.ptr(M), core.ptr(P),
                                              core.stream_ptr()), "stft_smp_forward")
        ctx.plan, ctx.in_dtype, ctx.in_shape = plan, y.dtype, y.shape
        ctx.save_for_backward(yc)
        return S.transpose(1, 2), M.transpose(1, 2), P.transpose(1, 2)

    @staticmethod
    def backward(ctx, gS, gM, gP):
        lib = core._lib.load()
        (yc,) = ctx.saved_tensors
        plan = ctx.plan
        B, T = yc.shape

        def fm(g):   # [B, F, T'] upstream (normally a view of frame-major memory, then this is free) -> contiguous [B, T', F]
            return None if g is None else g.to(device=yc.device, dtype=torch.float32).transpose(1, 2).contiguous()
        gS, gM, gP = fm(gS), fm(gM), fm(gP)
        g_y = torch.empty((B, T), device=yc.device, dtype=torch.float32)
        ws = core._workspace(int(lib.sb200_stft_smp_workspace_bytes(plan.handle, B, T)), yc.device, "stft_smp")
        core.check(lib.sb200_stft_smp_backward(plan.handle, core.ptr(yc), B, T, core.ptr(gS), core.ptr(gM), core.ptr(gP),
                                               core.ptr(g_y), core.ptr(ws), core.stream_ptr()), "stft_smp_backward")
        return g_y.to(ctx.in_dtype).reshape(ctx.in_shape), None


def get_stft_torch(y, n_fft, win_length, hop_length):
    """S = |D + 1e-9|, M = mel_basis @ S, P = angle(D) for y [B, T] (retunegan/audio.py:150-170).

    Outputs are [B, F, T'] / [B, n_mel, T'] views of frame-major buffers, computed by one fused launch, and they are
    differentiable with respect to ``y`` like the reference's (torch.stft -> abs / matmul / angle under autograd).
    The mel basis is the Slaney one whatever ``hp.mel_scale`` says (audio.py:158).
    """
    if not isinstance(y, torch.Tensor):
        raise TypeError("get_stft_torch expects a torch.Tensor [B, T]")
    squeeze = y.dim() == 1
    if squeeze:
        y = y.unsqueeze(0)
    if y.shape[-1] <= n_fft // 2:
        raise RuntimeError(f"Argument #4: Padding size should be less than the corresponding input dimension, "
                           f"but got: padding ({n_fft // 2}, {n_fft // 2}) at dimension 2 of input")   # torch.stft
    plan = core.get_plan(hp, n_fft, win_length, hop_length)
    S, M, P = _StftTorchFn.apply(y, plan)
    return (S[0], M[0], P[0]) if squeeze else (S, M, P)
