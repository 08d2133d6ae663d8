"""Frame-level side features that share the STFT framing: get_c0 (RMS), get_zcr, get_uv,
trim_silence, quantilize_c0 -- transtacos/audio.py:59-61,112-128, retunegan/audio.py:98-113.

CPU part: the oracle restatement of librosa 0.8.1 rms / zero_crossing_rate / effects.trim against an independent
torch.unfold formulation (librosa is not installable here: parity unpinned upstream) and against the reference's
own constants.  GPU part: the CUDA kernel through the reference-named API against the oracle.
Tolerances: RMS 1e-5 relative (fp32 sums of 1024 squares); zero-crossing rate and trim bounds exact.
"""
import numpy as np
import pytest
import torch

from oracle import spectral_oracle as O


def _wav(L, seed, kind="speech"):
    return O.synth_speechlike(L, seed) if kind == "speech" else O.synth_noise(L, seed)


def _torch_frames(y, frame_length, hop, mode):
    yp = torch.nn.functional.pad(torch.from_numpy(np.asarray(y, np.float64))[None, None], (frame_length // 2,) * 2,
                                 mode="reflect" if mode == "reflect" else "replicate")[0, 0]
    return yp.unfold(0, frame_length, hop)   # [T, frame_length]


@pytest.mark.parametrize("L,fl,hop", [(256 * 40 - 1, 1024, 256), (5000, 512, 128), (1025, 1024, 256)])
def test_oracle_rms_zcr_vs_torch_unfold(L, fl, hop):
    y = _wav(L, 7)
    fr = _torch_frames(y, fl, hop, "reflect")
    ref_rms = fr.pow(2).mean(1).sqrt().numpy()
    got = O.rms(y.astype(np.float64), fl, hop)
    assert got.shape == (1 + L // hop,)
    np.testing.assert_allclose(got, ref_rms, rtol=1e-12)
    fe = _torch_frames(y, fl, hop, "edge")
    fe = torch.where(fe.abs() <= 1e-10, torch.zeros_like(fe), fe)
    sb = torch.signbit(fe)
    ref_zcr = (sb[:, 1:] != sb[:, :-1]).double().sum(1).numpy() / fl
    np.testing.assert_array_equal(O.zero_crossing_rate(y, fl, hop), ref_zcr)


def test_oracle_trim_and_quantise():
    L = 22050
    y = np.zeros(L, np.float32)
    y[6000:15000] = _wav(9000, 3)
    s, e = O.trim_bounds(y, 35, 512, 128)
    assert 0 < s <= 6000 + 512 and 15000 - 512 <= e < L and s % 128 == 0
    assert s >= 6000 - 512 and e <= 15000 + 512
    assert O.trim_bounds(np.zeros(4096, np.float32)) == (0, 4096)   # all frames equal the (floored) reference: nothing is trimmed
    assert O.trim_bounds(np.r_[np.zeros(3000), 0.5 * np.ones(10), np.zeros(3000)].astype(np.float32))[0] > 0
    # quantilize_c0: the reference's own constants (transtacos/hparam.py:22-28) map c0min -> 0 and c0max -> last bin
    q = O.tt_quantilize_c0(np.array([4.6309418394230306e-05, 0.3751049339771271, 0.1, 1.0]))
    assert q.tolist() == [0, 31, 8, 31] and q.dtype == np.int32
    uv = O.rtg_get_uv(np.array([0.1, 0.2, 0.1], np.float32), np.array([0.5, 0.5, 0.01], np.float32))
    assert uv.tolist() == [0.0, 1.0, 1.0] and uv.dtype == np.float32


def test_oracle_yin_pins_and_pure_tones():
    # weak pin from the reference tree: transtacos/hparam.py:24-25 are the period limits of rf0min / rf0max
    assert np.isclose(22050 / np.ceil(22050 / O.note_to_hz('D2')), 73.25581359863281)
    assert np.isclose(22050 / np.floor(22050 / O.note_to_hz('D5')), 595.9459228515625)
    t = np.arange(22050) / 22050.0
    for f in (110.0, 220.0, 331.0, 440.0):
        y = (0.3 * np.sin(2 * np.pi * f * t) + 0.1 * np.sin(4 * np.pi * f * t)).astype(np.float32)
        f0 = O.tt_get_f0(y)
        assert f0.shape == (1 + len(y) // 256,) and f0.dtype == np.float32
        assert np.all(np.abs(f0[4:-4] / f - 1) < 5e-3), (f, f0[4:-4].min(), f0[4:-4].max())
    assert np.allclose(O.tt_get_f0(np.zeros(4096, np.float32)), 22050 / 37)    # silence: d' = 0 everywhere -> first index -> shortest period
    q = O.tt_quantilize_f0(np.array([73.26, 100.0, 440.0, 595.9]))
    assert q.tolist() == [0, 6, 32, 37] and q.dtype == np.int32


def test_oracle_side_features_match_reference_source(golden_side):
    """The audio.py layer of the widened rows, executed from the reference's own source (make_golden_side.py)."""
    g = golden_side
    y = g["y_speech"]
    lb = O.linear_basis()
    assert lb.shape == (1025, 80) and np.abs(lb - g["tt_linear_basis"]).max() <= 1e-12 * np.abs(g["tt_linear_basis"]).max()
    M = O.tt_spec_to_natural_scale(g["tt_mel_norm_speech"])
    np.testing.assert_allclose(O.tt_mel_to_linear(M), g["tt_mel_to_linear_speech"], rtol=1e-12, atol=1e-300)
    w = O.tt_inv_mel(g["tt_mel_norm_speech"], init_phase=g["tt_inv_mel_phase"])
    assert w.dtype == np.float32 and np.linalg.norm(w - g["tt_inv_mel_speech"]) <= 1e-6 * np.linalg.norm(g["tt_inv_mel_speech"])
    np.testing.assert_array_equal(O.tt_get_c0(y), g["tt_get_c0_speech"])
    np.testing.assert_array_equal(O.tt_get_f0(y), g["tt_get_f0_speech"])
    np.testing.assert_array_equal(O.rtg_get_zcr(y), g["rtg_get_zcr_speech"])
    np.testing.assert_array_equal(O.rtg_get_uv(g["rtg_get_zcr_speech"], g["rtg_get_c0_speech"]), g["rtg_get_uv_speech"])
    np.testing.assert_array_equal(O.tt_quantilize_c0(g["tt_get_c0_speech"]), g["tt_quantilize_c0"])
    np.testing.assert_array_equal(O.tt_quantilize_f0(g["tt_get_f0_speech"]), g["tt_quantilize_f0"])
    assert g["tt_n_f0"].tolist() == [37, 39]
    np.testing.assert_array_equal(O.tt_trim_silence(g["tt_trim_in"]), g["tt_trim_silence"])
    assert 0 < len(g["tt_trim_silence"]) < len(g["tt_trim_in"]) and len(g["tt_align_wav"]) == 1024


def test_oracle_pool_losses_match_reference_source(golden_side):
    g = golden_side
    y, yg = g["pool_y"], g["pool_yg"]
    assert abs(O.rtg_envelope_loss(y, yg) - g["pool_envelope_loss"]) < 1e-6 * abs(g["pool_envelope_loss"])
    assert abs(O.rtg_dynamic_loss(y, yg) - g["pool_dynamic_loss"]) < 1e-6 * abs(g["pool_dynamic_loss"])
    np.testing.assert_allclose(O.rtg_pool_loss_backward(y, yg, 0), g["pool_envelope_grad"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(O.rtg_pool_loss_backward(y, yg, 1), g["pool_dynamic_grad"], rtol=1e-6, atol=1e-12)
    assert np.count_nonzero(g["pool_envelope_grad"][:, -77:]) == 0


# ------------------------------------------------------------------------------------------------ GPU ----

@pytest.fixture(scope="module")
def sb():
    import transtacos_retunegan_b200 as sb
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    sb._lib.load()
    return sb


@pytest.mark.gpu
@pytest.mark.parametrize("L", [256 * 40 - 1, 110335, 1500, 600])
def test_c0_zcr_single(sb, L):
    y = _wav(L, 11)
    c0 = sb.transtacos_audio.get_c0(y)
    assert isinstance(c0, np.ndarray) and c0.dtype == np.float32 and c0.shape == (1 + L // 256,)
    np.testing.assert_allclose(c0, O.tt_get_c0(y), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(sb.retunegan_audio.get_c0(y), c0, rtol=0, atol=0)
    z = sb.retunegan_audio.get_zcr(y)
    assert z.dtype == np.float32 and z.shape == c0.shape
    np.testing.assert_array_equal(z, O.rtg_get_zcr(y))
    np.testing.assert_array_equal(sb.retunegan_audio.get_uv(z, c0), O.rtg_get_uv(O.rtg_get_zcr(y), O.tt_get_c0(y)))


@pytest.mark.gpu
def test_c0_zcr_batched_ragged_and_torch(sb):
    ys = [_wav(L, 20 + i, "noise" if i % 2 else "speech") for i, L in enumerate([9000, 30001, 4097, 256 * 101 - 1])]
    c0s = sb.transtacos_audio.get_c0(ys)
    zs = sb.retunegan_audio.get_zcr(ys)
    for y, c0, z in zip(ys, c0s, zs):
        np.testing.assert_allclose(c0, O.tt_get_c0(y), rtol=1e-5, atol=1e-9)
        np.testing.assert_array_equal(z, O.rtg_get_zcr(y))
    Y = np.stack([_wav(8191, 40 + i) for i in range(5)])
    c0 = sb.transtacos_audio.get_c0(torch.from_numpy(Y).cuda())
    assert c0.is_cuda and tuple(c0.shape) == (5, 32)
    np.testing.assert_allclose(c0.cpu().numpy(), np.stack([O.tt_get_c0(y) for y in Y]), rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_zcr_exact_zeros_and_threshold(sb):
    y = np.zeros(4096, np.float32)
    y[100:200] = 1e-11            # below librosa's threshold: counts as +0
    y[300:400:2] = -0.5           # alternating sign
    y[300:400][1::2] = 0.0        # exact zeros are positive
    y[1000] = -1e-9
    np.testing.assert_array_equal(sb.retunegan_audio.get_zcr(y), O.rtg_get_zcr(y))
    np.testing.assert_allclose(sb.transtacos_audio.get_c0(y), O.tt_get_c0(y), rtol=1e-5, atol=1e-12)


@pytest.mark.gpu
def test_trim_silence_and_quantise(sb):
    L = 30000
    y = (1e-4 * O.synth_noise(L, 5)).astype(np.float32)
    y[7000:21000] += _wav(14000, 6)
    got = sb.transtacos_audio.trim_silence(y)
    want = O.tt_trim_silence(y)
    assert got.shape == want.shape and 0 < len(got) < L
    np.testing.assert_array_equal(got, want)
    both = sb.transtacos_audio.trim_silence([y, y[:20000]])
    np.testing.assert_array_equal(both[0], want)
    np.testing.assert_array_equal(both[1], O.tt_trim_silence(y[:20000]))
    c0 = sb.transtacos_audio.get_c0(got[:len(got) // 256 * 256 - 1])
    np.testing.assert_array_equal(sb.transtacos_audio.quantilize_c0(c0), O.tt_quantilize_c0(c0))


@pytest.mark.gpu
def test_f0_yin(sb):
    """get_f0 against the oracle restatement of librosa.yin.  The period choice is a discrete decision on fp32 (here) vs
    fp64-FFT (librosa) difference functions, so a frame may legitimately flip between two troughs at a tie; required: every
    frame within 1e-4 relative of the oracle on tonal input, >= 99 % of the frames on speech-like / noisy input
    (measured on the B200: all frames, p99 of the relative deviation 1e-6)."""
    t = np.arange(30000) / 22050.0
    y = (0.3 * np.sin(2 * np.pi * 196.0 * t) + 0.1 * np.sin(4 * np.pi * 196.0 * t)).astype(np.float32)
    f0 = sb.transtacos_audio.get_f0(y)
    ref = O.tt_get_f0(y)
    assert f0.dtype == np.float32 and f0.shape == ref.shape
    assert np.max(np.abs(f0 / ref - 1)) < 1e-4
    ys = [_wav(L, 60 + i) for i, L in enumerate([256 * 80 - 1, 40001, 9000])]
    f0s = sb.transtacos_audio.get_f0(ys)
    for y, f in zip(ys, f0s):
        ref = O.tt_get_f0(y)
        assert f.shape == ref.shape
        ok = np.abs(f / ref - 1) < 1e-4
        assert ok.mean() >= 0.99, ok.mean()
        assert f.min() >= 22050 / 302 and f.max() <= 22050 / 36
    Y = np.stack([_wav(8191, 70 + i) for i in range(3)])
    fb = sb.transtacos_audio.get_f0(torch.from_numpy(Y).cuda())
    assert fb.is_cuda and tuple(fb.shape) == (3, 32)
    np.testing.assert_array_equal(sb.transtacos_audio.quantilize_f0(ref), O.tt_quantilize_f0(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("split_cv,ref_wav", [(False, 'y'), (True, 'y'), (False, 'dy')])
def test_retunegan_dataset_tuples(sb, split_cv, ref_wav):
    """retunegan/data.py:38-122 restated on the oracle vs retunegan_data.prepare_batch (one batch, five launches)."""
    wavs = []
    for i, T in enumerate([40, 57, 33]):
        wavs.append(_wav(256 * T, 80 + i))
    got = sb.retunegan_data.prepare_batch(wavs, split_cv=split_cv, ref_wav=ref_wav)
    for w, g in zip(wavs, got):
        mag = O.rtg_get_mag(w[:-1])
        mel = O.rtg_mag_to_mel(mag)
        tm = np.pad(O.rtg_inv_mag(mag, wavlen=len(w) - 1), (0, 1))
        if ref_wav == 'dy':
            tp = np.pad(tm, (0, 1))
            tm = tp[1:] - tp[:-1]
        assert g[0].shape == mel.shape and g[1] is not None and len(g[1]) == len(w)
        assert np.linalg.norm(g[0] - mel) / np.linalg.norm(mel) < 1e-4
        tg = g[2] if not split_cv else g[4] + g[5]
        assert np.linalg.norm(tg - tm) / np.linalg.norm(tm) < 1e-3                 # Griffin-Lim tolerance
        if split_cv:
            zcr, dyn = O.rtg_get_zcr(tm[:-1]), O.tt_get_c0(tm[:-1])
            uv = O.rtg_get_uv(zcr, dyn)
            uv_ex = np.repeat(uv, 256)
            agree = (g[6] == uv_ex).mean()                                         # thresholds on a 1e-3-accurate wav
            assert agree > 0.98, agree
            u = g[6][::256]
            mel_min = g[0].min()
            np.testing.assert_allclose(g[2], (g[0] - mel_min) * u + mel_min, rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(g[3], (g[0] - mel_min) * (1 - u) + mel_min, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_side_features_golden(sb, golden_side):
    """The CUDA path against the fixtures generated from the reference's own source."""
    g = golden_side
    y = g["y_speech"]
    TA, RA = sb.transtacos_audio, sb.retunegan_audio
    np.testing.assert_allclose(TA.get_c0(y), g["tt_get_c0_speech"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(RA.get_c0(y), g["rtg_get_c0_speech"], rtol=1e-5, atol=1e-9)
    np.testing.assert_array_equal(RA.get_zcr(y), g["rtg_get_zcr_speech"])
    np.testing.assert_array_equal(RA.get_uv(g["rtg_get_zcr_speech"], g["rtg_get_c0_speech"]), g["rtg_get_uv_speech"])
    f0 = TA.get_f0(y)
    assert (np.abs(f0 / g["tt_get_f0_speech"] - 1) < 1e-4).mean() >= 0.99
    np.testing.assert_array_equal(TA.quantilize_c0(g["tt_get_c0_speech"]), g["tt_quantilize_c0"])
    np.testing.assert_array_equal(TA.quantilize_f0(g["tt_get_f0_speech"]), g["tt_quantilize_f0"])
    np.testing.assert_array_equal(TA.trim_silence(g["tt_trim_in"]), g["tt_trim_silence"])
    np.testing.assert_array_equal(TA.align_wav(y[:1000]), g["tt_align_wav"])
    # inv_mel: pseudo-inverse basis product, then the Griffin-Lim path of inv_spec
    M = O.tt_spec_to_natural_scale(g["tt_mel_norm_speech"])
    S = TA._mel_to_linear(M.astype(np.float32))
    ref = g["tt_mel_to_linear_speech"]
    assert S.shape == ref.shape == (1025, 24)
    assert np.linalg.norm(S - ref) / np.linalg.norm(ref) < 1e-5 and np.abs(S - ref).max() < 1e-5 * np.abs(ref).max()
    w = TA.inv_mel(g["tt_mel_norm_speech"], init_phase=g["tt_inv_mel_phase"])
    assert w.dtype == np.float32 and w.shape == g["tt_inv_mel_speech"].shape
    assert np.linalg.norm(w - g["tt_inv_mel_speech"]) / np.linalg.norm(g["tt_inv_mel_speech"]) < 1e-3   # Griffin-Lim tolerance


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode", [("envelope", 0), ("dynamic", 1)])
def test_pool_losses_golden(sb, golden_side, name, mode):
    """envelope_loss / dynamic_loss (retunegan/models/loss.py:66-82) value and autograd gradient vs the reference's torch code."""
    g = golden_side
    y = torch.from_numpy(g["pool_y"]).cuda().unsqueeze(1)
    yg = torch.from_numpy(g["pool_yg"]).cuda().unsqueeze(1).requires_grad_(True)
    fn = sb.envelope_loss if mode == 0 else sb.dynamic_loss
    loss = fn(y, yg)
    (3.0 * loss).backward()
    ref = float(g[f"pool_{name}_loss"])
    assert abs(loss.item() - ref) < 1e-5 * abs(ref)                      # loss value tolerance
    grad = yg.grad[:, 0].cpu().numpy() / 3.0
    np.testing.assert_allclose(grad, g[f"pool_{name}_grad"], rtol=1e-5, atol=1e-10)   # same arg-max positions, same signs
    assert fn(y, yg.detach()).requires_grad is False
    with pytest.raises(RuntimeError):
        fn(y[..., :100], yg[..., :100])


@pytest.mark.gpu
@pytest.mark.parametrize("fl,hop,fmin,fmax", [(2048, 512, 65.0, 2093.0), (512, 128, 200.0, 1500.0), (1024, 240, 73.4, 587.3)])
def test_yin_other_framings(sb, fl, hop, fmin, fmax):
    """sb200_yin away from the reference's framing (librosa's default 2048 / 512, a short frame, a hop that does not divide the frame)."""
    y = _wav(30011, 91)
    f0, frames = sb.core.yin(y, 22050, fmin, fmax, fl, hop)
    ref = O.yin(y, fmin, fmax, 22050, fl, None, hop).astype(np.float32)
    assert int(frames[0]) == len(ref) == 1 + len(y) // hop
    ok = np.abs(f0.cpu().numpy() / ref - 1) < 1e-4
    assert ok.mean() >= 0.99, ok.mean()
    with pytest.raises(ValueError):
        sb.core.yin(y, 22050, fmax, fmin, fl, hop)          # fmin >= fmax


@pytest.mark.gpu
@pytest.mark.parametrize("fl,hop,L", [(1024, 256, 1023), (400, 160, 16000), (512, 128, 130)])
def test_frame_stats_other_framings(sb, fl, hop, L):
    y = _wav(max(L, 700), 92)[:L]
    rms, zcr, frames = sb.core.frame_stats(y, fl, hop)
    if L > fl // 2:     # np.pad(mode='reflect') needs the signal longer than the pad
        np.testing.assert_allclose(rms.cpu().numpy(), O.rms(y, fl, hop), rtol=1e-5, atol=1e-9)
    np.testing.assert_array_equal(zcr.cpu().numpy(), O.zero_crossing_rate(y, fl, hop).astype(np.float32))
    assert int(frames[0]) == 1 + L // hop


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,k", [(1, 1000, 100), (5, 4099, 37), (2, 160, 160)])
def test_pool_losses_other_shapes(sb, B, T, k):
    """Windows that are not a multiple of the warp size, a tail outside every window, a single window per row."""
    rs = np.random.RandomState(B * 1000 + k)
    y = (0.3 * rs.randn(B, T)).astype(np.float32)
    yg = np.tanh(y + 0.05 * rs.randn(B, T)).astype(np.float32)
    old = sb.loss.hp
    sb.loss.set_hparams(sb.RETUNEGAN.replace(envelope_pool_k=k))
    try:
        for mode, fn, ref in ((0, sb.envelope_loss, O.rtg_envelope_loss), (1, sb.dynamic_loss, O.rtg_dynamic_loss)):
            tg = torch.from_numpy(yg).cuda().requires_grad_(True)
            loss = fn(torch.from_numpy(y).cuda(), tg)
            loss.backward()
            want = ref(y, yg, k)
            assert abs(loss.item() - want) < 1e-5 * abs(want) + 1e-9
            np.testing.assert_allclose(tg.grad.cpu().numpy(), O.rtg_pool_loss_backward(y, yg, mode, k), rtol=1e-5, atol=1e-10)
    finally:
        sb.loss.set_hparams(old)
