"""Immutable configuration of the spectral path.

Mirrors exactly the hparam.py attributes the hot path reads (transtacos/hparam.py:5-17,90-95;
retunegan/hparam.py:3-15,36-40,72-81,97).  The reference configures everything through a global
``hparam`` module; ``SpectralConfig.from_hparam(module)`` adapts such a module.
"""
from __future__ import annotations

import dataclasses
from typing import Tuple

PI = 3.14159265358979  # retunegan/utils.py:12
EPS = 1e-5             # transtacos/audio.py:13, retunegan/audio.py:19

_WINDOWS = {"hann": 0, "hamming": 1, "blackman": 2, "bartlett": 3}


@dataclasses.dataclass(frozen=True)
class SpectralConfig:
    sample_rate: int = 22050
    n_fft: int = 2048
    win_length: int = 1024
    hop_length: int = 256
    n_mel: int = 80
    n_freq: int = 1025
    preemphasis: float = 0.97
    ref_level_db: float = 20
    min_level_db: float = -100
    max_abs_value: float = 4
    fmin: float = 125
    fmax: float = 7600
    window_fn: str = "hann"
    mel_scale: str = "slaney"
    gl_iters: int = 30
    gl_power: float = 1.2
    gl_momentum: float = 0.0
    randseed: int = 114514
    rf0min: object = 'D2'                   # transtacos/hparam.py:18-19: YIN search range (note name or Hz)
    rf0max: object = 'D5'
    f0min: float = 73.25581359863281        # transtacos/hparam.py:24-25: range of the quantiser (= sr / 301, sr / 37)
    f0max: float = 595.9459228515625
    trim_below_peak_db: float = 35          # transtacos/hparam.py:15, retunegan/hparam.py:13
    c0min: float = 4.6309418394230306e-05   # transtacos/hparam.py:22-23,28 (quantilize_c0)
    c0max: float = 0.3751049339771271
    n_c0_bins: int = 32
    multi_stft_params: Tuple[Tuple[int, int, int], ...] = ((2048, 1024, 240), (1024, 512, 120), (512, 256, 60))
    phd_input: str = "stft"
    envelope_pool_k: int = 160              # retunegan/hparam.py:90

    def __post_init__(self):
        if self.window_fn not in _WINDOWS:
            # retunegan/hparam.py:36 also lists 'kaiser', which scipy.get_window cannot build without beta
            raise ValueError(f"unsupported window_fn {self.window_fn!r}; supported: {sorted(_WINDOWS)}")
        if self.mel_scale not in ("slaney", "htk"):
            raise ValueError("mel_scale must be 'slaney' or 'htk'")
        if self.n_freq != self.n_fft // 2 + 1:
            raise ValueError("n_freq must equal n_fft // 2 + 1")
        if not self.fmax < self.sample_rate // 2:
            raise ValueError("fmax must be < sample_rate // 2 (transtacos/audio.py:160)")

    @property
    def window_id(self) -> int:
        return _WINDOWS[self.window_fn]

    def plan_key(self, n_fft=None, win_length=None, hop_length=None, htk=False):
        """Key of the plan tables.  ``htk``: mel scale of the filterbank.  The reference honours ``hp.mel_scale`` ONLY in
        ``get_mel`` (retunegan/audio.py:126 ``htk=hp.mel_scale=='htk'``); ``mel_basis`` / ``mag_to_mel`` (audio.py:20-21) and
        ``get_stft_torch`` / ``multi_stft_loss`` (audio.py:158) always build the Slaney basis, so the flag is explicit and
        defaults to Slaney."""
        n_fft = self.n_fft if n_fft is None else int(n_fft)
        win_length = self.win_length if win_length is None else int(win_length)
        hop_length = self.hop_length if hop_length is None else int(hop_length)
        return (self.sample_rate, n_fft, win_length, hop_length, self.n_mel, float(self.fmin), float(self.fmax),
                int(bool(htk)), self.window_id)

    def replace(self, **kw) -> "SpectralConfig":
        return dataclasses.replace(self, **kw)

    @classmethod
    def from_hparam(cls, hp, **overrides) -> "SpectralConfig":
        """Build from a reference-style ``hparam`` module / namespace (missing attributes keep defaults)."""
        kw = {}
        for f in dataclasses.fields(cls):
            if hasattr(hp, f.name):
                v = getattr(hp, f.name)
                if f.name == "multi_stft_params":
                    v = tuple(tuple(int(a) for a in p) for p in v)
                kw[f.name] = v
        kw.update(overrides)
        return cls(**kw)


_NOTE = {'C': 0, 'D': 2, 'E': 4, 'F': 5, 'G': 7, 'A': 9, 'B': 11}


def note_to_hz(note) -> float:
    """librosa.note_to_hz for one note name ('D2', 'C#4', 'Bb3'; A4 = 440 Hz) or a number (returned as float)."""
    if not isinstance(note, str):
        return float(note)
    s = note.strip()
    pitch = _NOTE[s[0].upper()]
    i = 1
    while i < len(s) and s[i] in '#b!♯♭':
        pitch += 1 if s[i] in '#♯' else -1
        i += 1
    octave = int(s[i:]) if i < len(s) else 0
    midi = 12 * (octave + 1) + pitch
    return 440.0 * 2.0 ** ((midi - 69) / 12.0)


def hz_to_midi(f):
    """librosa.hz_to_midi."""
    import numpy as np
    return 12 * (np.log2(np.asanyarray(f)) - np.log2(440.0)) + 69


# transtacos/hparam.py:90-91 -- 30 iterations, angle form (no momentum)
TRANSTACOS = SpectralConfig(gl_iters=30, gl_power=1.2, gl_momentum=0.0)
# retunegan/hparam.py:38-40 -- 4 iterations, momentum 0.7
RETUNEGAN = SpectralConfig(gl_iters=4, gl_power=1.2, gl_momentum=0.7)
