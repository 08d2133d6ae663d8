#!/usr/bin/env python3
"""Small shapes of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_cases.py
    compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize_cases.py

Shapes are a few frames per utterance so that the instrumented kernels finish in seconds, but cover every code path
that has its own synchronisation: the warp-specialised feature kernel (mbarrier hand-off, edge + interior + shared-sample
items, ragged tails), the single-role feature kernels (n_fft 1024 / 512), the complex-output STFT kernel, the tiled
Griffin-Lim kernels (per-iteration launches and the persistent cooperative kernel with its grid barrier, both forms),
ISTFT, the mstft forward / backward / fused kernels with and without spec stacks, YIN, frame statistics, the max-pool
losses, the mel projections and the (inverse) pre-emphasis scan.  `SB200_SANITIZE_ONLY=name[,name]` selects cases.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import transtacos_retunegan_b200 as sb  # noqa: E402

ta, ra = sb.transtacos_audio, sb.retunegan_audio
rs = np.random.RandomState(0)


def noise(L):
    return np.clip(0.1 * rs.randn(L), -0.999, 0.999).astype(np.float32)


def case_feat3():
    ta.get_specs(noise(256 * 12 - 1))                                        # edge + interior items, pre-emphasis, dB
    ta.get_specs([noise(256 * 9 - 1), noise(5000), noise(256 * 3)])          # ragged, tails inside an item
    ta.get_specs(torch.from_numpy(np.stack([noise(256 * 10 - 1) for _ in range(3)])).cuda())   # uniform, odd row offsets
    ra.get_mag_mel(noise(256 * 11 - 1))                                      # ln epilogue, no pre-emphasis
    ra.get_mag(noise(256 * 6 - 1))                                           # magnitude only
    ra.get_mel(noise(256 * 6 - 1))                                           # mel only
    old = ra.hp
    ra.set_hparams(old.replace(hop_length=240))                              # HS = 0 instantiation (no shared-sample path)
    ra.get_mag_mel(noise(240 * 9 - 1))
    ra.set_hparams(old)


def case_feat2():
    old = ra.hp
    for n_fft, win, hop in ((1024, 512, 128), (512, 256, 64), (1024, 512, 120)):
        ra.set_hparams(old.replace(n_fft=n_fft, win_length=win, hop_length=hop, n_freq=n_fft // 2 + 1))
        ra.get_mag_mel([noise(hop * 17 - 1), noise(3 * n_fft + 5)])
    ra.set_hparams(old)


def case_stft_complex():
    y = torch.from_numpy(np.stack([noise(4096), noise(4096)])).cuda()
    for n_fft, win, hop in sb.RETUNEGAN.multi_stft_params:
        ra.get_stft_torch(y, n_fft, win, hop)
    ra.mag_to_mel(np.abs(rs.randn(1025, 7)).astype(np.float32))
    ta._mel_to_linear(np.abs(rs.randn(80, 7)).astype(np.float32))


def case_gl():
    y = noise(256 * 20 - 1)
    mag = ra.get_mag(y)
    ra.inv_mag(mag, wavlen=len(y))                                           # persistent kernel, fast form, length given
    ra.inv_mag(mag[1:])                                                      # F = 1024 (zero DC row), length None
    S, _ = ta.get_specs(y)
    ta.inv_spec(S, n_iter=3)                                                 # persistent kernel, angle form + de-emphasis
    plan = sb.core.get_plan(ra.hp)
    frames = [20, 7, 33, 9]
    lens = [256 * t - 1 for t in frames]
    batch = sb.core.SignalBatch(plan, [noise(L) for L in lens])
    mag_fm, _, _ = sb.core.stft_features(plan, batch, 0.0, ra.ln_scale(True), None, True, False)
    ra.inv_mag_batch(mag_fm, frames, lens, init_phase="seeded")              # ragged batch
    D = torch.from_numpy((rs.randn(9, 1025) + 1j * rs.randn(9, 1025)).astype(np.complex64)).cuda()
    sb.core.istft(plan, D, sb.core.FramesBatch(plan, [9], None, D.device))


def case_gl_multilaunch():
    # per-iteration launches (what batches larger than one wave use): forced through the environment switch in a child
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, transtacos_retunegan_b200 as sb\n"
            "y = np.clip(0.1 * np.random.RandomState(1).randn(256 * 20 - 1), -0.999, 0.999).astype(np.float32)\n"
            "ra = sb.retunegan_audio; ta = sb.transtacos_audio\n"
            "ra.inv_mag(ra.get_mag(y), wavlen=len(y)); ta.inv_spec(ta.get_specs(y)[0], n_iter=2)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SB200_GL_PERSISTENT="0"))
    assert r.returncode == 0


def case_mstft():
    y = torch.from_numpy(np.stack([noise(4096), noise(4096)])).cuda().unsqueeze(1)
    yg = torch.tanh(y * 1.05 + 0.01).requires_grad_(True)
    sb.multi_stft_loss(y, yg, ret_loss=True).backward()                      # fused value + gradient
    loss, (sr, sg) = sb.multi_stft_loss(y, yg, ret_loss=True, ret_specs=True)   # forward with spec stacks
    ups = [torch.randn_like(s) * 1e-3 for s in sg]
    torch.autograd.backward([loss] + list(sg), [torch.ones_like(loss)] + ups)   # backward with dense spec gradients
    with torch.no_grad():
        sb.multi_stft_loss(y, yg, ret_loss=True)                             # forward only
    sb.loss.envelope_loss(y, yg).backward()
    sb.loss.dynamic_loss(y, yg).backward()


def case_side():
    y = noise(256 * 30 - 1)
    ta.get_f0(y)
    ta.get_c0(y)
    ra.get_zcr(y)
    ta.trim_silence(y)
    ta.inv_preemphasis(ta.preemphasis(y).astype(np.float32))
    ta.inv_preemphasis(noise(70000))                                         # multi-block scan


CASES = {"feat3": case_feat3, "feat2": case_feat2, "stft_complex": case_stft_complex, "gl": case_gl,
         "gl_multilaunch": case_gl_multilaunch, "mstft": case_mstft, "side": case_side}

if __name__ == "__main__":
    only = [s for s in os.environ.get("SB200_SANITIZE_ONLY", "").split(",") if s]
    for name, fn in CASES.items():
        if only and name not in only:
            continue
        fn()
        torch.cuda.synchronize()
        print("case", name, "done, launches so far", sb._lib.launch_count(), flush=True)
    print("all cases done")
