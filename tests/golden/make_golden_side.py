#!/usr/bin/env python3
"""Golden vectors for the widened rows, again by EXECUTING THE REFERENCE'S OWN SOURCE
(transtacos/audio.py, retunegan/audio.py imported unmodified from /root/reference; build container only).

Pinned to the reference's numpy code: ``_get_linear_basis`` / ``_mel_to_linear`` / ``inv_mel``
(transtacos/audio.py:100-104,164-175), ``quantilize_f0`` / ``quantilize_c0`` (:117-128), ``align_wav`` (:52-56),
``get_uv`` (retunegan/audio.py:108-113).  ``get_c0`` / ``get_f0`` / ``get_zcr`` / ``trim_silence`` call librosa, which is
absent: the shim routes them to the oracle's restated librosa layer, so those fixtures pin only the wrappers
(arguments, dtype casts); the librosa layer itself is cross-checked in tests/test_frame_features.py.

Usage:  python tests/golden/make_golden_side.py      (writes tests/golden/reference_vectors_side.npz)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G   # noqa: E402
from make_golden import O, REF   # noqa: E402


def main():
    G._install_stubs()
    L = sys.modules["librosa"]
    L.feature.rms = lambda y=None, frame_length=2048, hop_length=512, **kw: O.rms(y, frame_length, hop_length)[None, :]
    L.feature.zero_crossing_rate = lambda y, frame_length=2048, hop_length=512, **kw: \
        O.zero_crossing_rate(y, frame_length, hop_length)[None, :]
    L.yin = lambda y, fmin, fmax, sr=22050, frame_length=2048, win_length=None, hop_length=None, **kw: \
        O.yin(y, fmin, fmax, sr, frame_length, win_length, hop_length)
    effects = types.ModuleType("librosa.effects")

    def _trim(y, top_db=60, ref=np.max, frame_length=2048, hop_length=512):
        s, e = O.trim_bounds(y, top_db, frame_length, hop_length)
        return y[s:e], np.asarray([s, e])
    effects.trim = _trim
    L.effects = effects
    sys.modules["librosa.effects"] = effects
    out = {}
    L1 = 256 * 24 - 1
    y = O.synth_speechlike(L1, 114514)
    out["y_speech"] = y

    G._forget("hparam", "audio", "utils")
    TT = G._load("ref_tt_audio_side", f"{REF}/transtacos/audio.py", f"{REF}/transtacos")
    out["tt_linear_basis"] = TT._get_linear_basis()
    S_norm, M_norm = TT.get_specs(y)
    out["tt_mel_norm_speech"] = M_norm
    out["tt_mel_to_linear_speech"] = TT._mel_to_linear(TT.spec_to_natural_scale(M_norm))
    np.random.seed(114514)
    out["tt_inv_mel_phase"] = np.random.rand(1025, 24)
    np.random.seed(114514)
    out["tt_inv_mel_speech"] = TT.inv_mel(M_norm)
    out["tt_get_c0_speech"] = TT.get_c0(y)
    out["tt_get_f0_speech"] = TT.get_f0(y)
    out["tt_quantilize_c0"] = TT.quantilize_c0(out["tt_get_c0_speech"])
    out["tt_quantilize_f0"] = TT.quantilize_f0(out["tt_get_f0_speech"])
    out["tt_n_f0"] = np.asarray([TT.hp.n_f0_min, TT.hp.n_f0_bins])
    padded = np.concatenate([1e-4 * O.synth_noise(3000, 1), y, 1e-4 * O.synth_noise(2500, 2)]).astype(np.float32)
    out["tt_trim_in"] = padded
    out["tt_trim_silence"] = TT.trim_silence(padded)
    out["tt_align_wav"] = TT.align_wav(y[:1000])

    G._forget("hparam", "audio", "utils")
    RT = G._load("ref_rtg_audio_side", f"{REF}/retunegan/audio.py", f"{REF}/retunegan")
    out["rtg_get_zcr_speech"] = RT.get_zcr(y)
    out["rtg_get_c0_speech"] = RT.get_c0(y)
    out["rtg_get_uv_speech"] = RT.get_uv(out["rtg_get_zcr_speech"], out["rtg_get_c0_speech"])

    # waveform max-pool losses (retunegan/models/loss.py:66-82) and their autograd gradients, torch CPU
    import torch
    sys.modules["audio"] = RT
    sys.modules["hparam"] = RT.hp
    UT = G._load("utils", f"{REF}/retunegan/utils.py", f"{REF}/retunegan")
    sys.modules["utils"] = UT
    LS = G._load("ref_rtg_loss_side", f"{REF}/retunegan/models/loss.py", f"{REF}/retunegan")
    B, T = 3, 8192 + 77                                     # 77 trailing samples outside every window of 160
    rs = np.random.RandomState(114514)
    yy = np.stack([O.synth_speechlike(T, 114514 + b) for b in range(B)]).astype(np.float32)
    yg = np.tanh(yy * 1.1 + 0.02 * rs.randn(B, T)).astype(np.float32)
    yg[1, 320:480] = 0.25                                   # a constant window: arg-max == arg-min == first sample
    out["pool_y"], out["pool_yg"] = yy, yg
    for name, fn in (("envelope", LS.envelope_loss), ("dynamic", LS.dynamic_loss)):
        ty = torch.from_numpy(yy).unsqueeze(1)
        tg = torch.from_numpy(yg).unsqueeze(1).requires_grad_(True)
        loss = fn(ty, tg)
        (g,) = torch.autograd.grad(loss, tg)
        out[f"pool_{name}_loss"] = np.asarray(loss.item())
        out[f"pool_{name}_grad"] = g[:, 0].numpy()

    path = os.path.join(HERE, "reference_vectors_side.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB,", len(out), "arrays")


if __name__ == "__main__":
    main()
