#!/usr/bin/env python3
"""Golden vectors on the REAL recordings the reference ships: img/gt_hfg.wav (int16, 76 293 samples),
img/y_tmpl.wav (float32, 71 663 samples) and img/gen_hfg.wav (the vocoded twin of gt_hfg, for the loss), all 22 050 Hz.

Like make_golden.py this EXECUTES THE REFERENCE'S OWN SOURCE (transtacos/audio.py, retunegan/audio.py,
retunegan/models/loss.py imported unmodified from /root/reference, with the librosa / plotting imports stubbed as
described there) and runs only in the build container.  The recordings themselves are committed inside the fixture
(they are inputs, 0.4 MB); the spectrogram outputs are kept at every 16th frame and in float32 to stay small.

Usage:  python tests/golden/make_golden_real.py      (writes tests/golden/reference_vectors_real.npz)
"""
import os
import sys

import numpy as np
import torch
from scipy.io import wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402

REF = G.REF
FRAME_STRIDE = 16
GL_SEG = (16384, 32768)          # samples of the Griffin-Lim output that are kept
LOSS_OFFS = (20000, 40000)       # 8192-sample segments (the reference's segment_size, retunegan/hparam.py) for the loss


def load_wav(path):
    """librosa.load(path, sr=None) semantics for these files: int16 -> float32 / 32768, float32 as is."""
    sr, x = wavfile.read(path)
    assert sr == 22050
    return (x.astype(np.float32) / 32768.0) if x.dtype == np.int16 else x.astype(np.float32)


def aligned(y, hop=256):
    """retunegan/data.py:60-62 / transtacos/datasets/databaker.py:108-111: align to the hop, then y[:-1]."""
    return y[:(len(y) // hop) * hop - 1]


def main():
    G._install_stubs()
    out = {}
    wavs = {"gt_hfg": load_wav(f"{REF}/img/gt_hfg.wav"), "y_tmpl": load_wav(f"{REF}/img/y_tmpl.wav")}
    out["wav_gt_hfg_int16"] = wavfile.read(f"{REF}/img/gt_hfg.wav")[1]
    out["wav_y_tmpl"] = wavs["y_tmpl"]

    G._forget("hparam", "audio", "utils")
    TT = G._load("ref_tt_audio", f"{REF}/transtacos/audio.py", f"{REF}/transtacos")
    for tag, w in wavs.items():
        S, M = TT.get_specs(aligned(w))
        out[f"tt_S_{tag}"] = S[:, ::FRAME_STRIDE].astype(np.float32)
        out[f"tt_M_{tag}"] = M.astype(np.float32)

    G._forget("hparam", "audio", "utils")
    RT = G._load("ref_rtg_audio", f"{REF}/retunegan/audio.py", f"{REF}/retunegan")
    for tag, w in wavs.items():
        y = aligned(w)
        mag = RT.get_mag(y)
        out[f"rtg_mag_{tag}"] = mag[:, ::FRAME_STRIDE]
        out[f"rtg_mel_{tag}"] = RT.get_mel(y)
        if tag == "gt_hfg":
            wav = RT.inv_mag(mag, wavlen=len(y))                  # the Griffin-Lim template of retunegan/data.py:76
            assert wav.dtype == np.float32 and len(wav) == len(y)
            out["rtg_inv_mag_gt_hfg_seg"] = wav[GL_SEG[0]:GL_SEG[1]]
            out["rtg_inv_mag_gt_hfg_norm"] = np.asarray(np.linalg.norm(wav.astype(np.float64)))

    # multi_stft_loss on real / vocoded segments (retunegan/train.py:139-193), float64 torch autograd of the reference graph
    sys.modules["audio"] = RT
    sys.modules["hparam"] = RT.hp
    UT = G._load("utils", f"{REF}/retunegan/utils.py", f"{REF}/retunegan")
    sys.modules["utils"] = UT
    LS = G._load("ref_rtg_loss", f"{REF}/retunegan/models/loss.py", f"{REF}/retunegan")
    gen = load_wav(f"{REF}/img/gen_hfg.wav")
    y = np.stack([wavs["gt_hfg"][o:o + 8192] for o in LOSS_OFFS])
    yg = np.stack([gen[o:o + 8192] for o in LOSS_OFFS])
    out["loss_y_real"], out["loss_yg_real"] = y, yg
    orig = RT.get_stft_torch

    def patched(y_, n_fft, win, hop):
        if n_fft not in RT.mel_basis_torch:
            orig(y_.float(), n_fft, win, hop)
        RT.mel_basis_torch[n_fft] = RT.mel_basis_torch[n_fft].to(y_.dtype)
        RT.window_fn_torch[win] = torch.hann_window(win, dtype=y_.dtype)
        return orig(y_, n_fft, win, hop)
    LS.get_stft_torch = patched
    RT.mel_basis_torch.clear(); RT.window_fn_torch.clear()
    ty = torch.from_numpy(y).double().unsqueeze(1)
    tg = torch.from_numpy(yg).double().unsqueeze(1).requires_grad_(True)
    loss = LS.multi_stft_loss(ty, tg, ret_loss=True)
    (g,) = torch.autograd.grad(loss, tg)
    out["loss_real_value_f64"] = np.asarray(loss.item())
    out["loss_real_grad_f64"] = g.numpy()[:, 0]

    path = os.path.join(HERE, "reference_vectors_real.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB,", len(out), "arrays")


if __name__ == "__main__":
    main()
