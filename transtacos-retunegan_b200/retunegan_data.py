"""Batched preparation of RetuneGAN's per-utterance training tuples.

``retunegan/data.py:38-130`` (``Dataset.__getitem__``) recomputes and caches, one utterance at a time inside DataLoader workers:
``mag = get_mag(wav[:-1])``, ``mel = mag_to_mel(mag)``, ``wav_tmpl = pad(inv_mag(mag, wavlen - 1), (0, 1))`` (the Griffin-Lim
reference wav), optionally its first difference (``hp.ref_wav == 'dy'``) and, for the split generators (``hp.split_cv``), the
unvoiced mask from ``get_zcr`` / ``get_c0`` of ``wav_tmpl`` with the masked copies of ``mel`` and ``wav_tmpl``.
``prepare_batch`` does the same for a list of utterances with five launches in total (fused STFT, mel projection, Griffin-Lim
init + iterations + finish, frame statistics) and returns the tuples the dataset caches.  Augmentation
(``augment_wav`` / ``augment_spec``, random and applied once per utterance) stays with the caller: it is not spectral work.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import core
from . import retunegan_audio as A


def prepare_batch(wavs: Sequence[np.ndarray], split_cv: bool = False, ref_wav: str = 'y', as_numpy: bool = True) -> List[tuple]:
    """``wavs``: float32 utterances already loaded (and augmented, if training) and aligned to ``hop_length``
    (``A.align_wav``).  Returns, per utterance, ``(mel [M, T], wav, wav_tmpl)`` or, with ``split_cv``,
    ``(mel, wav, mel_c, mel_v, wav_tmpl_c, wav_tmpl_v, uv_ex)`` exactly as ``retunegan/data.py:119-122`` caches them."""
    hp = A.hp
    if ref_wav not in ('y', 'dy'):
        raise ValueError("ref_wav must be 'y' or 'dy' (retunegan/hparam.py:41)")
    wavs = [np.ascontiguousarray(w, np.float32) for w in wavs]
    for w in wavs:
        if len(w) % hp.hop_length:
            raise ValueError("utterances must be aligned to hop_length (retunegan/data.py:47 align_wav)")
    plan = core.get_plan(hp)
    batch = core.SignalBatch(plan, [w[:-1] for w in wavs])                 # `[:-1]` to avoid an extra trailing frame (:58)
    sc = A.ln_scale(True)
    mag, _, _ = core.stft_features(plan, batch, 0.0, sc, sc, True, False)   # ln-magnitudes [sum T, F]   (get_mag, :59)
    mel = core.mel_project(plan, mag)                                      # mag_to_mel on the ln-magnitudes (:66)
    frames = [int(t) for t in batch.frames]
    tmpl, off = A.inv_mag_batch(mag, frames, [len(w) - 1 for w in wavs])   # inv_mag(mag, wavlen - 1) (:75)
    out, fo = [], 0
    tm_list = []
    for i, w in enumerate(wavs):
        t = torch.nn.functional.pad(tmpl[int(off[i]):int(off[i + 1])], (0, 1))   # np.pad(wav_tmpl, (0, 1)) (:76)
        if ref_wav == 'dy':                                                       # first difference (:81-83)
            t = torch.nn.functional.pad(t, (0, 1))
            t = t[1:] - t[:-1]
        tm_list.append(t)
    uv = None
    if split_cv:                                                                  # u/v mask of the reference wav (:86-89)
        cuts = [t[:-1] for t in tm_list]
        dyn, zcr, _ = core.frame_stats(cuts, hp.win_length, hp.hop_length)
        uv = A.get_uv(zcr, dyn)
    for i, w in enumerate(wavs):
        T = frames[i]
        m = mel[fo:fo + T].t()
        t = tm_list[i]
        assert len(w) == t.numel() == T * hp.hop_length                           # retunegan/data.py:115
        if split_cv:
            u = uv[fo:fo + T]
            uv_ex = u.repeat_interleave(hp.hop_length)                            # np.repeat(uv, hop_length) (:93)
            mel_min = m.min()
            shift = m - mel_min
            item = (m, w, shift * u + mel_min, shift * (1 - u) + mel_min, t * uv_ex, t * (1 - uv_ex), uv_ex)
        else:
            item = (m, w, t)
        if as_numpy:
            item = tuple(x.cpu().numpy() if isinstance(x, torch.Tensor) else x for x in item)
        out.append(item)
        fo += T
    return out
