"""Drop-in for the spectral functions of ``transtacos/audio.py`` (same names, arguments, return arity).

    get_specs(y) -> (mag_norm [F,T], mel_norm [M,T])          transtacos/audio.py:73-77
    inv_spec(spec) -> wav                                      transtacos/audio.py:93-97
    preemphasis / inv_preemphasis / spec_to_natural_scale / fix_zero_DC / align_wav

numpy in -> numpy out with the reference's dtypes (float64 features, float32 wav); torch in -> float32
CUDA tensors out.  Every spectrogram is returned as an ``[F, T]`` view of frame-major memory (strides
(1, F)), the layout ``librosa.stft`` itself allocates (order='F').  Additions over the reference:
batched input (``[B, L]`` tensor/array or a list of ragged utterances -> list of per-utterance results),
``init_phase=`` / ``n_iter=`` on ``inv_spec``.  Configuration lives in the module attribute ``hp``
(a ``SpectralConfig``; the reference reads a global ``hparam`` module) -- use ``set_hparams``.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import core
from .config import TRANSTACOS, SpectralConfig, hz_to_midi, note_to_hz

hp: SpectralConfig = TRANSTACOS
eps = 1e-5


def set_hparams(cfg) -> None:
    """Install a SpectralConfig or a reference-style ``hparam`` module."""
    global hp
    hp = cfg if isinstance(cfg, SpectralConfig) else SpectralConfig.from_hparam(cfg, gl_momentum=0.0)


def _is_np(x) -> bool:
    return isinstance(x, np.ndarray) or (isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], np.ndarray))


def db_norm_scale(cfg: SpectralConfig) -> core.Scale:
    """_normalize(_amp_to_db(x) - ref_level_db) as a*log2(max(1e-5, x)) + b (transtacos/audio.py:177-193)."""
    a = 2 * cfg.max_abs_value * 20 * math.log10(2.0) / -cfg.min_level_db
    b = 2 * cfg.max_abs_value * ((-cfg.ref_level_db - cfg.min_level_db) / -cfg.min_level_db) - cfg.max_abs_value
    return core.log_scale(a, b, 1e-5)


def _split_fm(t: torch.Tensor, frames, width: int, as_numpy: bool, np_dtype, single: bool):
    """[total_frames, width] frame-major -> per-utterance [width, T] views (numpy: Fortran-ordered)."""
    if as_numpy:
        host = t.cpu().numpy()
        if np_dtype is not None and host.dtype != np_dtype:
            host = host.astype(np_dtype)
    outs, o = [], 0
    for T in frames:
        T = int(T)
        outs.append((host[o:o + T].T if as_numpy else t[o:o + T].t()))
        o += T
    return outs[0] if single else outs


def align_wav(wav, r=None):
    """transtacos/audio.py:52-56 -- pad to a multiple of hop_length."""
    r = hp.hop_length if r is None else r
    d = len(wav) % r
    if d != 0:
        wav = np.pad(wav, (0, (r - d))) if isinstance(wav, np.ndarray) else torch.nn.functional.pad(wav, (0, r - d))
    return wav


def _frame_feature(y, frame_length, hop_length, which):
    """[L] -> [T]; [B, L] -> [B, T]; list of utterances -> list of [T_i].  numpy in -> float32 numpy out, torch in -> CUDA."""
    as_np = _is_np(y) or (isinstance(y, (list, tuple)) and len(y) > 0 and _is_np(y[0]))
    rms, zcr, frames = core.frame_stats(y, frame_length, hop_length, want_rms=which == "rms", want_zcr=which == "zcr")
    v = rms if which == "rms" else zcr
    if isinstance(y, (list, tuple)):
        parts = list(torch.split(v, [int(t) for t in frames]))
        return [p.cpu().numpy() for p in parts] if as_np else parts
    v = v.view(len(frames), int(frames[0]))
    if (y.ndim if _is_np(y) else y.dim()) == 1:
        v = v[0]
    return v.cpu().numpy() if as_np else v


def get_c0(y):
    """transtacos/audio.py:112-114: ``librosa.feature.rms(y, frame_length=win_length, hop_length=hop_length)[0]`` as float32."""
    return _frame_feature(y, hp.win_length, hp.hop_length, "rms")


def get_f0(y):
    """transtacos/audio.py:107-109: ``librosa.yin(y, fmin=rf0min, fmax=rf0max, frame_length=win_length, hop_length=hop_length)``
    as float32 ([L] -> [T]; [B, L] -> [B, T]; list -> list)."""
    as_np = _is_np(y) or (isinstance(y, (list, tuple)) and len(y) > 0 and _is_np(y[0]))
    f0, frames = core.yin(y, hp.sample_rate, note_to_hz(hp.rf0min), note_to_hz(hp.rf0max), hp.win_length, hp.hop_length)
    if isinstance(y, (list, tuple)):
        parts = list(torch.split(f0, [int(t) for t in frames]))
        return [p.cpu().numpy() for p in parts] if as_np else parts
    f0 = f0.view(len(frames), int(frames[0]))
    if (y.ndim if _is_np(y) else y.dim()) == 1:
        f0 = f0[0]
    return f0.cpu().numpy() if as_np else f0


def quantilize_f0(f0):
    """transtacos/audio.py:117-121: MIDI bins ``hz_to_midi(f) - floor(hz_to_midi(f0min))`` clipped to the bin range, int32."""
    f0 = np.asarray(f0.detach().cpu() if isinstance(f0, torch.Tensor) else f0)
    n_min = int(np.floor(hz_to_midi(hp.f0min)))
    n_bins = int(np.ceil(hz_to_midi(hp.f0max))) - n_min + 1
    return (hz_to_midi(f0) - n_min).clip(0, n_bins - 1).astype(np.int32)


def quantilize_c0(c0):
    """transtacos/audio.py:124-128: linear bins of [c0min, c0max] -> int32 in [0, n_c0_bins - 1] (host arithmetic)."""
    c0 = np.asarray(c0.detach().cpu() if isinstance(c0, torch.Tensor) else c0)
    q = (c0 - hp.c0min) / (hp.c0max - hp.c0min) * hp.n_c0_bins
    return q.clip(0, hp.n_c0_bins - 1).astype(np.int32)


def trim_silence(wav, frame_length=512, hop_length=128):
    """transtacos/audio.py:59-61: ``librosa.effects.trim(wav, top_db=trim_below_peak_db, frame_length, hop_length)[0]``.
    The RMS track comes from the GPU (one launch); the threshold / first-last scan is host arithmetic on T values.
    A list of utterances is trimmed with one launch for all of them."""
    many = isinstance(wav, (list, tuple))
    ws = list(wav) if many else [wav]
    rms, _, frames = core.frame_stats(ws, frame_length, hop_length, want_zcr=False)
    lens = [int(w.shape[-1]) for w in ws]
    out = [w[s:e] for w, (s, e) in zip(ws, core.trim_bounds(rms, frames, lens, hop_length, hp.trim_below_peak_db))]
    return out if many else out[0]


def preemphasis(x):
    """x[n] - k*x[n-1], zero initial state (transtacos/audio.py:64-66; scipy promotes to float64)."""
    t = core.to_device_f32(x)
    out = core.preemphasis(t.reshape(1, -1) if t.dim() == 1 else t, hp.preemphasis).reshape(t.shape)
    return out.cpu().numpy().astype(np.float64) if isinstance(x, np.ndarray) else out


def inv_preemphasis(x):
    """y[n] = x[n] + k*y[n-1] (transtacos/audio.py:69-70)."""
    t = core.to_device_f32(x)
    out = core.inv_preemphasis(t.reshape(1, -1) if t.dim() == 1 else t, hp.preemphasis).reshape(t.shape)
    return out.cpu().numpy().astype(np.float64) if isinstance(x, np.ndarray) else out


def _host_batch(y):
    """A uniform [B, L] batch living in host memory (numpy or CPU tensor) -> CPU float32 tensor, else None."""
    if isinstance(y, np.ndarray) and y.ndim == 2 and y.shape[0] > 1:
        return torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
    if isinstance(y, torch.Tensor) and not y.is_cuda and y.dim() == 2 and y.shape[0] > 1:
        return y.detach().to(torch.float32).contiguous()
    return None


def features_host(cfg: SpectralConfig, y_host: torch.Tensor, preemph, mag_scale, mel_scale, want_mag=True,
                  want_mel=True, out=None, chunk=16):
    """Host-resident [B, L] batch -> host features through the chunked copy/compute pipeline.
    ``out=(mag [B*T, F], mel [B*T, n_mel])`` lets the caller supply (pinned) destinations; returns CPU tensors."""
    plan = core.get_plan(cfg)
    B, L = y_host.shape
    if L < plan.n_fft // 4 + 1:
        raise ValueError("signal shorter than n_fft/4 + 1 samples: reflect padding undefined")
    T = 1 + L // plan.hop_length
    mag = mel = None
    if out is not None:
        mag, mel = out
    else:
        if want_mag:
            mag = torch.empty((B * T, plan.F), dtype=torch.float32)
        if want_mel:
            mel = torch.empty((B * T, plan.n_mel), dtype=torch.float32)
    core.host_feature_pipeline(plan, chunk).run(y_host, mag, mel, preemph, mag_scale, mel_scale)
    return mag, mel, T


def get_specs(y, out_dtype=None, out=None):
    """(normalised dB magnitude [F,T], normalised dB mel [M,T]) -- transtacos/audio.py:73-77.

    numpy in -> numpy out in float64, the dtype the reference returns (scipy's lfilter promotes, transtacos/audio.py:66;
    the features are np.save'd as float64, datasets/databaker.py:113-114).  The kernels compute in float32:
    ``out_dtype=np.float32`` skips the widening.  torch in -> float32 torch views out (CPU tensor in -> CPU out, CUDA in ->
    CUDA out).  ``[B, L]`` host batches stream through pinned-copy / launch / copy-back overlap; ``out=(mag [B*T, F],
    mel [B*T, M])`` supplies (pinned) float32 CPU destination tensors, which are then what is returned (no widening).
    """
    sc = db_norm_scale(hp)
    yh = _host_batch(y)
    if yh is not None:
        if isinstance(y, np.ndarray) and not np.isfinite(y).all():
            raise ValueError("Audio buffer is not finite everywhere")
        mag, mel, T = features_host(hp, yh, hp.preemphasis, sc, sc, out=out)
        B = yh.shape[0]
        if isinstance(y, np.ndarray):
            dt = (np.float32 if out is not None else np.float64) if out_dtype is None else out_dtype
            S = mag.numpy().astype(dt, copy=False).reshape(B, T, -1).transpose(0, 2, 1)
            M = mel.numpy().astype(dt, copy=False).reshape(B, T, -1).transpose(0, 2, 1)
            return S, M
        return mag.view(B, T, -1).transpose(1, 2), mel.view(B, T, -1).transpose(1, 2)
    as_np = _is_np(y)
    plan = core.get_plan(hp)
    batch = core.SignalBatch(plan, y)
    mag, mel, _ = core.stft_features(plan, batch, preemph=hp.preemphasis, mag_scale=sc, mel_scale=sc)
    single = not isinstance(y, (list, tuple)) and getattr(y, "ndim", 1) == 1
    dt = (np.float64 if out_dtype is None else out_dtype) if as_np else None
    S = _split_fm(mag, batch.frames, plan.F, as_np, dt, single)
    M = _split_fm(mel, batch.frames, plan.n_mel, as_np, dt, single)
    if not single and not isinstance(y, (list, tuple)):      # uniform [B, L] CUDA batch -> [B, F, T] views
        S = torch.stack(S) if not as_np else np.stack(S)
        M = torch.stack(M) if not as_np else np.stack(M)
    return S, M


def _amp_to_db(x):
    return 20 * np.log10(np.maximum(1e-5, x)) if isinstance(x, np.ndarray) else 20 * torch.log10(x.clamp_min(1e-5))


def _db_to_amp(x):
    return np.power(10.0, x * 0.05) if isinstance(x, np.ndarray) else torch.pow(10.0, x * 0.05)


def _normalize(S):
    return 2 * hp.max_abs_value * ((S - hp.min_level_db) / -hp.min_level_db) - hp.max_abs_value


def _denormalize(S):
    return ((S + hp.max_abs_value) * -hp.min_level_db) / (2 * hp.max_abs_value) + hp.min_level_db


def spec_to_natural_scale(spec):
    """transtacos/audio.py:80-82 (host-side element-wise helper, kept for API parity)."""
    return _db_to_amp(_denormalize(spec) + hp.ref_level_db)


def fix_zero_DC(S):
    """transtacos/audio.py:85-90 -- prepend a DC row S.min()*1e-2 when F == n_freq - 1."""
    F, T = S.shape
    if F == hp.n_freq - 1:
        if isinstance(S, np.ndarray):
            S = np.concatenate([np.ones([1, T]) * S.min() * 1e-2, S], axis=0)
        else:
            S = torch.cat([torch.ones(1, T, device=S.device, dtype=S.dtype) * S.min() * 1e-2, S], dim=0)
    return S


def _to_frame_major(spec) -> torch.Tensor:
    """[F, T] (numpy any order / torch any strides) -> contiguous float32 CUDA [T, F]."""
    if isinstance(spec, np.ndarray):
        return core.to_device_f32(np.ascontiguousarray(spec.T, dtype=np.float32))
    return spec.detach().to(device=core.require_cuda(), dtype=torch.float32).t().contiguous()


def draw_phase(F: int, T: int, seed=None) -> np.ndarray:
    """The reference's initial-phase draw: ``np.random.rand(F, T)`` from the global RNG
    (transtacos/audio.py:134) or a fresh ``RandomState(seed)`` (librosa.griffinlim random_state=int)."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    return rng.rand(F, T)


def phase_to_frame_major(init_phase, F: int, T: int, device, seed=None) -> torch.Tensor:
    """User-facing ``init_phase`` ([F, T] values in [0,1) like the reference's rand draw, None, or 'device')
    -> float32 CUDA [T, F]."""
    if init_phase is None:
        init_phase = draw_phase(F, T, seed)
    if isinstance(init_phase, str):
        if init_phase != "device":
            raise ValueError("init_phase must be an [F, T] array, None or 'device'")
        return torch.rand((T, F), device=device, dtype=torch.float32)   # throughput mode: on-device counter RNG
    if tuple(init_phase.shape) != (F, T):
        raise ValueError(f"init_phase must have shape {(F, T)}, got {tuple(init_phase.shape)}")
    return _to_frame_major(init_phase)


def griffin_lim_amplitude(S_fm: torch.Tensor, T: int, phase_fm: torch.Tensor, n_iter: int, momentum: float, form: int,
                          length, inv_preemph: float, cfg: SpectralConfig) -> torch.Tensor:
    """Shared driver: S_fm [T, F] amplitudes already raised to gl_power, phase_fm [T, F] -> wav (float32 CUDA)."""
    plan = core.get_plan(cfg)
    fb = core.FramesBatch(plan, [T], None if not length else [int(length)], S_fm.device)
    return core.griffinlim(plan, S_fm, phase_fm, fb, n_iter, momentum, form, inv_preemph)


def _griffin_lim(S, init_phase=None, n_iter=None):
    """transtacos/audio.py:130-140 -- 'angle' form, no momentum; returns the raw Griffin-Lim signal."""
    S_fm = _to_frame_major(np.abs(S) if isinstance(S, np.ndarray) else S.abs())
    ph = phase_to_frame_major(init_phase, S.shape[0], S.shape[1], S_fm.device)
    y = griffin_lim_amplitude(S_fm, S.shape[1], ph, hp.gl_iters if n_iter is None else n_iter, 0.0, 0, None, 0.0, hp)
    return y.cpu().numpy().astype(np.float64) if isinstance(S, np.ndarray) else y


def inv_spec(spec, init_phase=None, n_iter=None):
    """transtacos/audio.py:93-97 -- denormalise, fix DC, S**gl_power, Griffin-Lim (30 it), de-emphasis."""
    F, T = spec.shape
    x = _to_frame_major(spec)                                     # [T, F or F-1]
    if F == hp.n_freq:
        S = core.spec_to_amplitude(x, 0, hp.max_abs_value, hp.min_level_db, hp.ref_level_db, hp.gl_power)
    elif F == hp.n_freq - 1:                                      # fix_zero_DC acts on the natural scale, before the power
        S = core.spec_to_amplitude(x, 0, hp.max_abs_value, hp.min_level_db, hp.ref_level_db, 1.0)
        S = torch.cat([(S.min() * 1e-2).expand(T, 1), S], dim=1).contiguous()
        S = core.spec_to_amplitude(S, 2, power=hp.gl_power)
    else:
        raise ValueError(f"expected {hp.n_freq} or {hp.n_freq - 1} frequency rows, got {F}")
    ph = phase_to_frame_major(init_phase, hp.n_freq, T, S.device)
    wav = griffin_lim_amplitude(S, T, ph, hp.gl_iters if n_iter is None else n_iter, 0.0, 0, None, hp.preemphasis, hp)
    return wav.cpu().numpy().astype(np.float32) if isinstance(spec, np.ndarray) else wav


def _mel_to_linear(mel):
    """transtacos/audio.py:164-165: ``np.dot(_get_linear_basis(), mel)``, [n_mel, T] -> [n_freq, T]."""
    out = core.mel_to_linear(core.get_plan(hp), _to_frame_major(mel))
    return out.cpu().numpy().T if isinstance(mel, np.ndarray) else out.t()


def inv_mel(mel, init_phase=None, n_iter=None):
    """transtacos/audio.py:100-104 -- denormalise, back to linear with the pseudo-inverse basis, S**gl_power, Griffin-Lim,
    de-emphasis ("might have no use case" upstream; it shares every kernel with ``inv_spec``)."""
    M, T = mel.shape
    if M != hp.n_mel:
        raise ValueError(f"expected {hp.n_mel} mel rows, got {M}")
    x = core.spec_to_amplitude(_to_frame_major(mel), 0, hp.max_abs_value, hp.min_level_db, hp.ref_level_db, 1.0)
    S = core.spec_to_amplitude(core.mel_to_linear(core.get_plan(hp), x), 2, power=hp.gl_power)
    ph = phase_to_frame_major(init_phase, hp.n_freq, T, S.device)
    wav = griffin_lim_amplitude(S, T, ph, hp.gl_iters if n_iter is None else n_iter, 0.0, 0, None, hp.preemphasis, hp)
    return wav.cpu().numpy().astype(np.float32) if isinstance(mel, np.ndarray) else wav
