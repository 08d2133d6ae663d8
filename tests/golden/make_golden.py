#!/usr/bin/env python3
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN SOURCE.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  The reference's audio.py / loss.py are imported unmodified from /root/reference;
the third-party modules they need but that are not installed here are provided as stand-ins:

  * ``librosa``  -> a shim whose stft / istft / griffinlim / filters.mel /
                    feature.melspectrogram are the oracle's restated librosa-0.8.1 layer
                    (oracle/spectral_oracle.py).  So these fixtures pin the audio.py / loss.py
                    logic (dB / normalise, pre-emphasis, DC fix, power 1.2, exp/log, clip, the
                    torch path of get_stft_torch, the loss reduction and its autograd gradient)
                    to the reference's code; the librosa layer itself is cross-checked against
                    torch.stft / torch.istft / torchaudio in tests/test_oracle.py.
  * ``seaborn``, ``matplotlib`` -> empty stubs (plotting only).
  * ``np.complex`` -> ``complex`` (removed in numpy >= 1.24; transtacos/audio.py:135).

Usage:  python tests/golden/make_golden.py      (writes tests/golden/*.npz)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import spectral_oracle as O  # noqa: E402


def _install_stubs():
    if not hasattr(np, "complex"):
        np.complex = complex  # noqa: NPY001
    L = types.ModuleType("librosa")

    def _stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True, dtype=None,
              pad_mode="reflect"):
        return O.stft(y, n_fft, hop_length, win_length, window, center, pad_mode)

    def _istft(stft_matrix, hop_length=None, win_length=None, window="hann", center=True, dtype=None,
               length=None):
        return O.istft(stft_matrix, hop_length, win_length, window, center, length)

    def _griffinlim(S, n_iter=32, hop_length=None, win_length=None, window="hann", center=True, dtype=None,
                    length=None, pad_mode="reflect", momentum=0.99, init="random", random_state=None):
        assert init == "random" and center
        return O.griffinlim(S, n_iter, hop_length, win_length, window, length, momentum, random_state)

    filters = types.ModuleType("librosa.filters")
    filters.mel = lambda sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney": \
        O.mel_filterbank(sr, n_fft, n_mels, fmin, fmax, htk)
    feature = types.ModuleType("librosa.feature")

    def _melspectrogram(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                        center=True, pad_mode="reflect", power=2.0, **kw):
        return O.melspectrogram(y, sr, n_fft, hop_length, win_length, kw["n_mels"], kw["fmin"], kw["fmax"],
                                window, power, kw.get("htk", False))

    feature.melspectrogram = _melspectrogram
    L.stft, L.istft, L.griffinlim = _stft, _istft, _griffinlim
    L.filters, L.feature = filters, feature
    _notes = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}
    L.note_to_hz = lambda n: 440.0 * 2.0 ** ((12 * (int(n[-1]) + 1) + _notes[n[0]] - 69) / 12)
    L.hz_to_midi = lambda f: 12 * (np.log2(np.asanyarray(f)) - np.log2(440.0)) + 69
    sys.modules["librosa"] = L
    sys.modules["librosa.filters"] = filters
    sys.modules["librosa.feature"] = feature
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "matplotlib.pylab"):
        m = types.ModuleType(name)
        m.use = lambda *a, **k: None
        sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]


def _load(name, path, extra_path):
    """Import a reference file under a private module name with its own 'hparam'/'audio'/'utils'."""
    sys.path.insert(0, extra_path)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(extra_path)
    return mod


def _forget(*names):
    for n in names:
        sys.modules.pop(n, None)


def main():
    _install_stubs()
    out = {}

    L1 = 256 * 24 - 1
    y_sp = O.synth_speechlike(L1, 114514)
    y_no = O.synth_noise(L1, 114515)
    out["y_speech"], out["y_noise"] = y_sp, y_no

    # ---------------- TransTacoS audio.py ----------------
    _forget("hparam", "audio", "utils")
    TT = _load("ref_tt_audio", f"{REF}/transtacos/audio.py", f"{REF}/transtacos")
    for tag, y in (("speech", y_sp), ("noise", y_no)):
        S, M = TT.get_specs(y)
        assert S.dtype == np.float64 and S.shape == (1025, 24) and M.shape == (80, 24)
        out[f"tt_get_specs_S_{tag}"], out[f"tt_get_specs_M_{tag}"] = S, M
    out["tt_preemphasis_speech"] = TT.preemphasis(y_sp)
    out["tt_inv_preemphasis_speech"] = TT.inv_preemphasis(y_sp)
    S_norm, _ = TT.get_specs(y_sp)
    out["tt_natural_speech"] = TT.spec_to_natural_scale(S_norm)
    np.random.seed(114514)
    phase = np.random.rand(1025, 24)
    np.random.seed(114514)
    out["tt_inv_spec_phase1025"] = phase
    out["tt_inv_spec_speech"] = TT.inv_spec(S_norm)                      # 30 it, F=1025
    np.random.seed(114514)
    out["tt_inv_spec_speech_F1024"] = TT.inv_spec(S_norm[1:])            # fix_zero_DC path (same rand stream)
    out["tt_fix_zero_DC"] = TT.fix_zero_DC(out["tt_natural_speech"][1:])[:1]

    # ---------------- RetuneGAN audio.py ----------------
    _forget("hparam", "audio", "utils")
    RT = _load("ref_rtg_audio", f"{REF}/retunegan/audio.py", f"{REF}/retunegan")
    for tag, y in (("speech", y_sp), ("noise", y_no)):
        out[f"rtg_get_mag_{tag}"] = RT.get_mag(y)
        out[f"rtg_get_mel_{tag}"] = RT.get_mel(y)
    out["rtg_get_mag_noclamp_speech"] = RT.get_mag(y_sp, clamp_low=False)
    mag = out["rtg_get_mag_speech"]
    out["rtg_mag_to_mel_speech"] = RT.mag_to_mel(mag)
    out["rtg_inv_mag_speech"] = RT.inv_mag(mag, wavlen=L1)               # 4 it, m=0.7, seed 114514
    out["rtg_inv_mag_speech_F1024"] = RT.inv_mag(mag[1:], wavlen=L1)
    out["rtg_inv_mag_speech_nolen"] = RT.inv_mag(mag)
    out["rtg_mel_basis"] = RT.mel_basis

    # get_stft_torch + multi_stft_loss (torch CPU; float32 like the reference, and float64)
    sys.modules["audio"] = RT
    sys.modules["hparam"] = RT.hp
    UT = _load("utils", f"{REF}/retunegan/utils.py", f"{REF}/retunegan")
    sys.modules["utils"] = UT
    LS = _load("ref_rtg_loss", f"{REF}/retunegan/models/loss.py", f"{REF}/retunegan")

    B, T = 2, 2048
    rs = np.random.RandomState(114514)
    y = np.stack([O.synth_speechlike(T, 114514 + b) for b in range(B)])
    yg = np.tanh(y + 0.05 * rs.randn(B, T)).astype(np.float32)
    out["loss_y"], out["loss_yg"] = y, yg
    for ri, (n_fft, win, hop) in enumerate(RT.hp.multi_stft_params):
        RT.mel_basis_torch.clear(); RT.window_fn_torch.clear()
        S, M, P = RT.get_stft_torch(torch.from_numpy(y), n_fft, win, hop)
        out[f"stft_torch_S_{ri}"], out[f"stft_torch_M_{ri}"], out[f"stft_torch_P_{ri}"] = \
            S.numpy(), M.numpy(), P.numpy()

    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        RT.mel_basis_torch.clear(); RT.window_fn_torch.clear()
        ty = torch.from_numpy(y).to(dt).unsqueeze(1)
        tg = torch.from_numpy(yg).to(dt).unsqueeze(1).requires_grad_(True)
        if dt == torch.float64:
            # the reference casts the mel filter with .float(); keep it but promote matmul operands
            orig = RT.get_stft_torch

            def patched(y_, n_fft, win, hop):
                if n_fft not in RT.mel_basis_torch:
                    orig(y_.float(), n_fft, win, hop)
                RT.mel_basis_torch[n_fft] = RT.mel_basis_torch[n_fft].to(y_.dtype)
                RT.window_fn_torch[win] = torch.hann_window(win, dtype=y_.dtype)   # exact f64 window for the f64 instrument
                return orig(y_, n_fft, win, hop)
            LS.get_stft_torch = patched
        loss = LS.multi_stft_loss(ty, tg, ret_loss=True)
        (g_only,) = torch.autograd.grad(loss, tg)
        out[f"loss_value_{tag}"] = np.asarray(loss.item())
        out[f"loss_grad_lossonly_{tag}"] = g_only.numpy()
        loss2, (sr_, sg_) = LS.multi_stft_loss(ty, tg, ret_loss=True, ret_specs=True)
        rs2 = np.random.RandomState(7)
        ups = [torch.from_numpy(rs2.randn(*s.shape)).to(dt) * 1e-3 for s in sg_]
        total = loss2 * 8 + sum((u * s).sum() for u, s in zip(ups, sg_))
        (g_all,) = torch.autograd.grad(total, tg)
        out[f"loss_grad_train_{tag}"] = g_all.numpy()
        if tag == "f32":   # upstream grads are regenerated in the tests: RandomState(7).randn(*shape) * 1e-3
            for ri in range(3):
                out[f"loss_specs_r_{ri}"] = sr_[ri].detach().numpy()
                out[f"loss_specs_g_{ri}"] = sg_[ri].detach().numpy()
    try:
        LS.multi_stft_loss(ty, tg)
        out["loss_noflag_raises"] = np.asarray(0)
    except RuntimeError:
        out["loss_noflag_raises"] = np.asarray(1)

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB,", len(out), "arrays")


if __name__ == "__main__":
    main()
