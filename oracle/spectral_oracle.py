"""CPU oracle for the spectral front/back end of TransTacoS / RetuneGAN.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  The product package (``transtacos-retunegan_b200``) never
imports anything from ``oracle/`` and has no CPU fallback.

What it restates (all paths relative to /root/reference):
  * transtacos/audio.py:64-97,130-196   (preemphasis, get_specs, inv_spec, Griffin-Lim)
  * retunegan/audio.py:19-21,116-170    (get_mag, get_mel, mag_to_mel, inv_mag, get_stft_torch)
  * retunegan/models/loss.py:22-62      (multi_stft_loss) + its closed-form backward
  * the widened rows: transtacos/audio.py:59-61,100-128,164-175 (trim_silence, inv_mel, get_f0, get_c0,
    quantisers, linear basis), retunegan/audio.py:98-113 (get_zcr, get_c0, get_uv), retunegan/models/loss.py:66-82
    (envelope_loss, dynamic_loss + closed-form backward); librosa 0.8.1 ``feature.rms / feature.zero_crossing_rate /
    effects.trim / yin`` restated (parity unpinned upstream; the audio.py / loss.py layer is pinned by
    tests/golden/make_golden_side.py)
  * the third-party layer those files call and that is NOT vendored in the
    reference: librosa==0.8.1 (requirements.txt:1) ``stft / istft / griffinlim /
    filters.mel / feature.melspectrogram / filters.window_sumsquare``, restated
    here from the published librosa 0.8.x algorithms.

Parity pinning status: the reference ships no golden vectors or tests
 and librosa cannot be installed in the build container, so the
*librosa layer* is "parity unpinned" upstream; it is cross-checked against
independent implementations that are present (torch.stft / torch.istft on CPU,
torchaudio.functional.melscale_fbanks) in tests/test_oracle.py.  The *audio.py /
loss.py layer* IS pinned: tests/golden/make_golden.py executes the reference's own
unmodified audio.py / loss.py source (imported from /root/reference, with this
module's librosa-layer functions standing in for the missing ``librosa`` import)
and commits its outputs as fixtures; tests/test_oracle.py checks this restatement
against them.

All functions are numpy (float64 where the reference is float64).
"""
from __future__ import annotations

import numpy as np
from scipy import signal as _signal

# --------------------------------------------------------------------------------------
# Configuration mirror of hparam.py (transtacos/hparam.py:5-17 == retunegan/hparam.py:3-15)
# --------------------------------------------------------------------------------------


class HP:
    sample_rate = 22050
    n_fft = 2048
    win_length = 1024
    hop_length = 256
    n_mel = 80
    n_freq = 1025
    preemphasis = 0.97
    ref_level_db = 20
    min_level_db = -100
    max_abs_value = 4
    fmin = 125
    fmax = 7600
    # transtacos/hparam.py:90-91,95
    tt_gl_iters = 30
    tt_gl_power = 1.2
    randseed = 114514
    # retunegan/hparam.py:36-40
    window_fn = "hann"
    mel_scale = "slaney"
    rtg_gl_iters = 4
    rtg_gl_momentum = 0.7
    rtg_gl_power = 1.2
    # retunegan/hparam.py:72-81
    multi_stft_params = [(2048, 1024, 240), (1024, 512, 120), (512, 256, 60)]
    phd_input = "stft"


EPS = 1e-5  # transtacos/audio.py:13, retunegan/audio.py:19
PI = 3.14159265358979  # retunegan/utils.py:12

# --------------------------------------------------------------------------------------
# librosa 0.8.1 layer (restated)
# --------------------------------------------------------------------------------------


def get_window(name: str, win_length: int) -> np.ndarray:
    """scipy.signal.get_window(name, win_length, fftbins=True) -> float64 periodic window."""
    return _signal.get_window(name, win_length, fftbins=True)


def pad_center(w: np.ndarray, size: int) -> np.ndarray:
    """librosa.util.pad_center: zero-pad ``w`` centred to ``size`` (lpad = (size-n)//2)."""
    n = w.shape[0]
    lpad = (size - n) // 2
    out = np.zeros(size, dtype=w.dtype)
    out[lpad:lpad + n] = w
    return out


def _dtype_r2c(dt):
    return np.complex64 if np.dtype(dt) == np.float32 else np.complex128


def _dtype_c2r(dt):
    return np.float32 if np.dtype(dt) == np.complex64 else np.float64


def stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
         pad_mode="reflect"):
    """librosa.stft (0.8.1).  Returns [1+n_fft/2, T] complex, Fortran order.

    float64 window * frames -> rfft in float64 -> stored complex64 for f32 input,
    complex128 for f64 input (Appendix A.1).
    """
    y = np.asarray(y)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    if not np.all(np.isfinite(y)):
        raise ValueError("Audio buffer is not finite everywhere")  # librosa ParameterError
    w = pad_center(get_window(window, win_length), n_fft).reshape(-1, 1)
    if center:
        y = np.pad(y, n_fft // 2, mode=pad_mode)
    elif n_fft > y.shape[-1]:
        raise ValueError("n_fft is too large for input signal")
    n_frames = 1 + (len(y) - n_fft) // hop_length
    frames = np.lib.stride_tricks.as_strided(
        y, shape=(n_fft, n_frames), strides=(y.itemsize, hop_length * y.itemsize), writeable=False)
    out = np.empty((1 + n_fft // 2, n_frames), dtype=_dtype_r2c(y.dtype), order="F")
    out[:] = np.fft.rfft(w * frames, axis=0)
    return out


def window_sumsquare(window, n_frames, hop_length, win_length, n_fft, dtype=np.float32):
    """librosa.filters.window_sumsquare(norm=None)."""
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = pad_center(get_window(window, win_length) ** 2, n_fft)
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))]
    return x


def istft(D, hop_length=None, win_length=None, window="hann", center=True, length=None):
    """librosa.istft (0.8.1) (Appendix A.2)."""
    D = np.asarray(D)
    n_fft = 2 * (D.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    w = pad_center(get_window(window, win_length), n_fft)[:, None]
    if length:
        padded_length = length + n_fft if center else length
        n_frames = min(D.shape[1], int(np.ceil(padded_length / hop_length)))
    else:
        n_frames = D.shape[1]
    dtype = _dtype_c2r(D.dtype)
    y = np.zeros(n_fft + hop_length * (n_frames - 1), dtype=dtype)
    ytmp = w * np.fft.irfft(D[:, :n_frames], axis=0)
    for t in range(n_frames):
        s = t * hop_length
        y[s:s + n_fft] += ytmp[:, t]
    wss = window_sumsquare(window, n_frames, hop_length, win_length, n_fft, dtype=dtype)
    nz = wss > np.finfo(wss.dtype).tiny
    y[nz] /= wss[nz]
    if length is None:
        if center:
            y = y[n_fft // 2: -(n_fft // 2)]
    else:
        start = n_fft // 2 if center else 0
        y = y[start:]
        if len(y) > length:
            y = y[:length]
        elif len(y) < length:
            y = np.pad(y, (0, length - len(y)))
    return y


def hz_to_mel(f, htk=False):
    f = np.asanyarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz(m, htk=False):
    m = np.asanyarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk, norm='slaney') -> f32 [n_mels, 1+n_fft/2]."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin, htk), hz_to_mel(fmax, htk), n_mels + 2), htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights


def melspectrogram(y, sr, n_fft, hop_length, win_length, n_mels, fmin, fmax, window="hann",
                   power=1, htk=False):
    """librosa.feature.melspectrogram: filters.mel(...) @ |stft(y)|**power."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window)) ** power
    return np.dot(mel_filterbank(sr, n_fft, n_mels, fmin, fmax, htk), S)


def griffinlim(S, n_iter=32, hop_length=None, win_length=None, window="hann", length=None,
               momentum=0.99, random_state=None, init_angles=None):
    """librosa.griffinlim (0.8.1, init='random') -- the fast-GL/momentum form (Appendix A.3).

    ``init_angles`` (complex [F,T]) overrides the RandomState draw (same values the
    reference would draw when given ``2*pi*RandomState(seed).rand(F,T)``).
    """
    S = np.asarray(S)
    n_fft = 2 * (S.shape[0] - 1)
    angles = np.empty(S.shape, dtype=np.complex64)
    if init_angles is not None:
        angles[:] = init_angles
    else:
        rng = np.random.RandomState(seed=random_state) if isinstance(random_state, int) else np.random
        angles[:] = np.exp(2j * np.pi * rng.rand(*S.shape))
    rebuilt = 0.0
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = istft(S * angles, hop_length=hop_length, win_length=win_length, window=window, length=length)
        rebuilt = stft(inverse, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window)
        angles[:] = rebuilt - (momentum / (1 + momentum)) * tprev
        angles[:] /= np.abs(angles) + 1e-16
    return istft(S * angles, hop_length=hop_length, win_length=win_length, window=window, length=length)


# --------------------------------------------------------------------------------------
# TransTacoS audio.py restatement
# --------------------------------------------------------------------------------------


def tt_preemphasis(x, hp=HP):
    """transtacos/audio.py:64-66 -- lfilter([1,-k],[1]) with zero state; promotes to float64."""
    return _signal.lfilter([1, -hp.preemphasis], [1], x)


def tt_inv_preemphasis(x, hp=HP):
    """transtacos/audio.py:69-70."""
    return _signal.lfilter([1], [1, -hp.preemphasis], x)


def _tt_stft(y, hp=HP):
    """transtacos/audio.py:143-144."""
    return stft(y, n_fft=hp.n_fft, hop_length=hp.hop_length, win_length=hp.win_length)


def _tt_istft(D, hp=HP):
    """transtacos/audio.py:147-148."""
    return istft(D, hop_length=hp.hop_length, win_length=hp.win_length)


def _amp_to_db(x):
    """transtacos/audio.py:177-184."""
    return 20 * np.log10(np.maximum(1e-5, x))


def _db_to_amp(x):
    """transtacos/audio.py:186-188."""
    return np.power(10.0, x * 0.05)


def _normalize(S, hp=HP):
    """transtacos/audio.py:190-193 (no clipping)."""
    return 2 * hp.max_abs_value * ((S - hp.min_level_db) / -hp.min_level_db) - hp.max_abs_value


def _denormalize(S, hp=HP):
    """transtacos/audio.py:195-196."""
    return ((S + hp.max_abs_value) * -hp.min_level_db) / (2 * hp.max_abs_value) + hp.min_level_db


_mel_cache: dict = {}


def mel_basis(n_fft=None, hp=HP):
    """transtacos/audio.py:157-162 / retunegan/audio.py:20,158 (cached per n_fft)."""
    n_fft = hp.n_fft if n_fft is None else n_fft
    key = (hp.sample_rate, n_fft, hp.n_mel, hp.fmin, hp.fmax, getattr(hp, "mel_scale", "slaney"))
    if key not in _mel_cache:
        assert hp.fmax < hp.sample_rate // 2
        _mel_cache[key] = mel_filterbank(hp.sample_rate, n_fft, hp.n_mel, hp.fmin, hp.fmax,
                                         htk=getattr(hp, "mel_scale", "slaney") == "htk")
    return _mel_cache[key]


def tt_get_specs(y, hp=HP):
    """transtacos/audio.py:73-77 -> (mag_norm [F,T] f64, mel_norm [M,T] f64)."""
    D = np.abs(_tt_stft(tt_preemphasis(y, hp), hp))
    S = _amp_to_db(D) - hp.ref_level_db
    M = _amp_to_db(np.dot(mel_basis(hp=hp), D)) - hp.ref_level_db
    return _normalize(S, hp), _normalize(M, hp)


def tt_spec_to_natural_scale(spec, hp=HP):
    """transtacos/audio.py:80-82."""
    return _db_to_amp(_denormalize(spec, hp) + hp.ref_level_db)


def tt_fix_zero_DC(S, hp=HP):
    """transtacos/audio.py:85-90."""
    F, T = S.shape
    if F == hp.n_freq - 1:
        S = np.concatenate([np.ones([1, T]) * S.min() * 1e-2, S], axis=0)
    return S


def tt_griffin_lim(S, hp=HP, init_phase=None, n_iter=None):
    """transtacos/audio.py:130-140 -- angle form, complex128, no momentum.

    ``init_phase`` ([F,T] in [0,1), the values ``np.random.rand`` would return) makes
    the draw explicit; None draws from the global numpy RNG exactly like the reference.
    """
    n_iter = hp.tt_gl_iters if n_iter is None else n_iter
    r = np.random.rand(*S.shape) if init_phase is None else np.asarray(init_phase, dtype=np.float64)
    angles = np.exp(2j * np.pi * r)
    S_complex = np.abs(S).astype(complex)
    y = _tt_istft(S_complex * angles, hp)
    for _ in range(n_iter):
        angles = np.exp(1j * np.angle(_tt_stft(y, hp)))
        y = _tt_istft(S_complex * angles, hp)
    return y


def tt_inv_spec(spec, hp=HP, init_phase=None, n_iter=None):
    """transtacos/audio.py:93-97."""
    S = tt_spec_to_natural_scale(spec, hp)
    S = tt_fix_zero_DC(S, hp)
    wav = tt_inv_preemphasis(tt_griffin_lim(S ** hp.tt_gl_power, hp, init_phase, n_iter), hp)
    return wav.astype(np.float32)


# --------------------------------------------------------------------------------------
# RetuneGAN audio.py restatement
# --------------------------------------------------------------------------------------


def rtg_get_mag(y, clamp_low=True, hp=HP):
    """retunegan/audio.py:116-120 -> ln|STFT| f32 [F,T]."""
    D = stft(y, n_fft=hp.n_fft, hop_length=hp.hop_length, win_length=hp.win_length, window=hp.window_fn)
    S = np.abs(D)
    mag = np.log(S.clip(min=EPS) if clamp_low else S)
    return mag.astype(np.float32)


def rtg_get_mel(y, clamp_low=True, hp=HP):
    """retunegan/audio.py:123-128 -> ln(mel_basis @ |STFT|) f32 [M,T]."""
    M = melspectrogram(y, hp.sample_rate, hp.n_fft, hp.hop_length, hp.win_length, hp.n_mel, hp.fmin,
                       hp.fmax, window=hp.window_fn, power=1, htk=hp.mel_scale == "htk")
    mel = np.log(M.clip(min=EPS) if clamp_low else M)
    return mel.astype(np.float32)


def rtg_mag_to_mel(x, hp=HP):
    """retunegan/audio.py:20-21 -- np.dot(mel_basis, x) on whatever it is given."""
    return np.dot(mel_basis(hp=hp), x)


def rtg_griffinlim(S, wavlen=None, hp=HP, init_angles=None):
    """retunegan/audio.py:131-136."""
    if hp.rtg_gl_power:
        S = S ** hp.rtg_gl_power
    y = griffinlim(S, n_iter=hp.rtg_gl_iters, hop_length=hp.hop_length, win_length=hp.win_length,
                   window=hp.window_fn, length=wavlen, momentum=hp.rtg_gl_momentum,
                   random_state=hp.randseed, init_angles=init_angles)
    return y.astype(np.float32)


def rtg_inv_mag(mag, wavlen=None, hp=HP, init_angles=None):
    """retunegan/audio.py:139-147."""
    S = np.exp(mag)
    F, T = mag.shape
    if F == hp.n_freq - 1:
        S = np.concatenate([np.zeros([1, T]), S], axis=0)
    y = rtg_griffinlim(S, wavlen, hp, init_angles)
    if wavlen:
        assert len(y) == wavlen
    return y


def rtg_get_stft(y, n_fft, win_length, hop_length, hp=HP, dtype=np.float32):
    """retunegan/audio.py:150-170 (numpy restatement of get_stft_torch) for y [B,T].

    torch.stft(center=True, pad_mode='reflect', window=hann(win), onesided) has the same
    framing / window / sign conventions as librosa.stft (Appendix A.7).  Returns
    S=|D+1e-9| [B,F,T'], M=mel@S [B,80,T'], P=angle(D) [B,F,T'] plus the complex D.
    Computation is done in float64 and rounded to ``dtype`` (the torch reference computes
    in float32; the difference is the parity tolerance).
    """
    y = np.asarray(y, dtype=np.float64)
    Ds = np.stack([stft(yb, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=hp.window_fn)
                   for yb in y])
    S = np.abs(Ds + 1e-9)
    M = np.einsum("mf,bft->bmt", mel_basis(n_fft, hp).astype(np.float64), S)
    P = np.angle(Ds)
    return S.astype(dtype), M.astype(dtype), P.astype(dtype), Ds


def rtg_multi_stft_loss(y, y_g, ret_loss=False, ret_specs=False, hp=HP, dtype=np.float32):
    """retunegan/models/loss.py:22-62 (numpy, float64 internally)."""
    y, y_g = np.asarray(y), np.asarray(y_g)
    if y.ndim == 3:
        y, y_g = y[:, 0], y_g[:, 0]
    loss = 0.0
    stft_r, stft_g = [], []
    for n_fft, win, hop in hp.multi_stft_params:
        S, M, P, _ = rtg_get_stft(y, n_fft, win, hop, hp, np.float64)
        Sg, Mg, Pg, _ = rtg_get_stft(y_g, n_fft, win, hop, hp, np.float64)
        if ret_specs:
            if hp.phd_input == "stft":
                stft_r.append(np.stack([np.log(S), P / PI], axis=1).astype(dtype))
                stft_g.append(np.stack([np.log(Sg), Pg / PI], axis=1).astype(dtype))
            elif hp.phd_input == "phase":
                stft_r.append(np.stack([np.log(S), P / PI], axis=1).astype(dtype))
                stft_g.append(np.stack([np.log(S), Pg / PI], axis=1).astype(dtype))
            else:
                raise RuntimeError
        loss += np.mean(np.abs(M - Mg))
        loss += np.mean(np.abs(np.log(M) - np.log(Mg)))
    loss /= len(hp.multi_stft_params)
    if ret_loss and ret_specs:
        return loss, (stft_r, stft_g)
    elif ret_loss:
        return loss
    elif ret_specs:
        return stft_r, stft_g
    raise RuntimeError("multi_stft_loss: neither ret_loss nor ret_specs")  # loss.py:62 bare raise


def rtg_multi_stft_loss_backward(y, y_g, g_loss=1.0, g_specs_g=None, hp=HP, tie_rel=None):
    """Closed-form d/dy_g of multi_stft_loss (the autograd graph of retunegan/train.py:192).

    SURVEY.md §8a row L2.  ``g_specs_g``: optional list of upstream grads [B,2,F,T'] for
    the generated-side stacks (ln S_g, P_g/PI), one per resolution.  Returns g_yg [B,T] f64.

    ``tie_rel``: the L1 terms contribute sign(M_g - M), which is discontinuous where a generated mel cell ties with the real
    one; any float32 evaluation (this repo's kernels and the reference's own torch float32 graph alike) may land on the other
    side of a tie that is closer than its rounding error.  With ``tie_rel`` set, cells with |M_g - M| <= tie_rel * M
    contribute nothing, and the function returns ``(g, n_ties)``: ``g_full - g`` is then exactly the part of the gradient
    that hinges on those near-ties (the tests bound the float32 disagreement by tol * |g| + 2 * |g_full - g|).
    """
    y, y_g = np.asarray(y, np.float64), np.asarray(y_g, np.float64)
    if y.ndim == 3:
        y, y_g = y[:, 0], y_g[:, 0]
    B, T = y_g.shape
    g = np.zeros((B, T))
    n_ties = 0
    nres = len(hp.multi_stft_params)
    for ri, (n_fft, win, hop) in enumerate(hp.multi_stft_params):
        _, M, _, _ = rtg_get_stft(y, n_fft, win, hop, hp, np.float64)
        Sg, Mg, _, Dg = rtg_get_stft(y_g, n_fft, win, hop, hp, np.float64)
        basis = mel_basis(n_fft, hp).astype(np.float64)
        cnt = M.size
        gM = g_loss * (np.sign(Mg - M) + np.sign(np.log(Mg) - np.log(M)) / Mg) / (nres * cnt)
        if tie_rel is not None:
            tie = np.abs(Mg - M) <= tie_rel * M
            n_ties += int(tie.sum())
            gM = np.where(tie, 0.0, gM)
        gS = np.einsum("mf,bmt->bft", basis, gM)
        gD = np.zeros_like(Dg)
        if g_specs_g is not None and g_specs_g[ri] is not None:
            gs = np.asarray(g_specs_g[ri], np.float64)
            if hp.phd_input == "stft":
                gS = gS + gs[:, 0] / Sg
            absD2 = np.abs(Dg) ** 2
            gP = gs[:, 1] / PI
            with np.errstate(divide="ignore", invalid="ignore"):
                gD = gD + np.where(absD2 > 0, gP * 1j * Dg / absD2, 0)
        gD = gD + gS * (Dg + 1e-9) / Sg
        # adjoint of the one-sided rfft: g_frame[n] = Re sum_k gD[k] e^{+2 pi i k n / N}
        w = pad_center(get_window(hp.window_fn, win), n_fft)
        n = np.arange(n_fft)
        k = np.arange(n_fft // 2 + 1)
        E = np.exp(2j * np.pi * np.outer(n, k) / n_fft)  # [N, F]
        Tf = gD.shape[2]
        h = n_fft // 2
        for b in range(B):
            gf = np.real(E @ gD[b]) * w[:, None]  # [N, T']
            gp = np.zeros(T + n_fft)
            for t in range(Tf):
                gp[t * hop:t * hop + n_fft] += gf[:, t]
            gb = gp[h:h + T].copy()
            i = np.arange(h)
            np.add.at(gb, h - i, gp[i])
            np.add.at(gb, T - 2 - i, gp[h + T + i])
            g[b] += gb
    return g if tie_rel is None else (g, n_ties)


def spectral_convergence(S_target, y, hp=HP):
    """||  |STFT(y)| - S ||_F / ||S||_F  -- Griffin-Lim parity metric."""
    D = np.abs(stft(np.asarray(y, np.float64), n_fft=hp.n_fft, hop_length=hp.hop_length, win_length=hp.win_length))
    T = min(D.shape[1], S_target.shape[1])
    return float(np.linalg.norm(D[:, :T] - S_target[:, :T]) / np.linalg.norm(S_target[:, :T]))


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------


def synth_noise(L, seed):
    rs = np.random.RandomState(seed)
    return np.clip(0.1 * rs.randn(L), -0.999, 0.999).astype(np.float32)


def synth_speechlike(L, seed, sr=22050):
    rs = np.random.RandomState(seed)
    t = np.arange(L) / sr
    phi = rs.rand() * 2 * np.pi
    f0 = 120 + 60 * np.sin(2 * np.pi * 0.7 * t + phi)
    ph = 2 * np.pi * np.cumsum(f0) / sr
    y = np.zeros(L)
    for h in range(1, 30):
        y += np.sin(h * ph + rs.rand() * 2 * np.pi) / h
    env = 0.5 * (1 + np.sin(2 * np.pi * 1.3 * t)) ** 2 / 4 + 0.02
    y = 0.15 * y * env + 0.003 * rs.randn(L)
    return np.clip(y, -0.999, 0.999).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# Frame statistics: librosa 0.8.1 feature.rms / feature.zero_crossing_rate / effects.trim,
# called at transtacos/audio.py:59-61,112-114 and retunegan/audio.py:98-113.  librosa is not installable here:
# "parity unpinned" upstream; tests cross-check against an independent torch.unfold formulation.

def _frame(y, frame_length, hop_length):
    """librosa.util.frame(y, frame_length, hop_length) -> [frame_length, n_frames] (a strided view in librosa)."""
    n = 1 + (len(y) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n)[None, :]
    return y[idx]


def rms(y, frame_length=2048, hop_length=512):
    """librosa.feature.rms(y=y, frame_length, hop_length, center=True, pad_mode='reflect')[0] (dtype of y)."""
    y = np.asarray(y)
    yp = np.pad(y, int(frame_length // 2), mode="reflect")
    x = _frame(yp, frame_length, hop_length)
    power = np.mean(np.abs(x) ** 2, axis=0)
    return np.sqrt(power)


def zero_crossing_rate(y, frame_length=2048, hop_length=512):
    """librosa.feature.zero_crossing_rate(y, frame_length, hop_length, center=True)[0]: edge padding,
    zero_crossings(threshold=1e-10, zero_pos=True, pad=False) along the frame axis, mean over the frame."""
    y = np.asarray(y)
    yp = np.pad(y, int(frame_length // 2), mode="edge")
    x = _frame(yp, frame_length, hop_length).copy()
    x[np.abs(x) <= 1e-10] = 0
    sign = np.signbit(x)
    cross = np.zeros(x.shape, dtype=bool)
    cross[1:] = sign[1:] != sign[:-1]
    return np.mean(cross, axis=0)


def tt_get_c0(y, win_length=1024, hop_length=256):
    """transtacos/audio.py:112-114 (= retunegan/audio.py:103-105)."""
    return rms(np.asarray(y, np.float32), win_length, hop_length).astype(np.float32)


def rtg_get_zcr(y, win_length=1024, hop_length=256):
    """retunegan/audio.py:98-100."""
    return zero_crossing_rate(np.asarray(y, np.float32), win_length, hop_length).astype(np.float32)


def rtg_get_uv(zcr, dyn):
    """retunegan/audio.py:108-113 (the loop, vectorised)."""
    zcr = np.asarray(zcr)
    return ((zcr > 0.18) | (np.asarray(dyn) < 0.03)).astype(zcr.dtype)


def trim_bounds(y, top_db=35, frame_length=512, hop_length=128):
    """librosa.effects.trim(y, top_db, ref=np.max, frame_length, hop_length)[1] -> (start, end)."""
    y = np.asarray(y)
    mse = rms(y, frame_length, hop_length) ** 2
    amin = 1e-10
    db = 10.0 * np.log10(np.maximum(amin, mse)) - 10.0 * np.log10(np.maximum(amin, np.max(mse)))
    nz = np.flatnonzero(db > -top_db)
    if nz.size:
        return int(nz[0]) * hop_length, min(len(y), (int(nz[-1]) + 1) * hop_length)
    return 0, 0


def tt_trim_silence(y, top_db=35, frame_length=512, hop_length=128):
    """transtacos/audio.py:59-61."""
    s, e = trim_bounds(y, top_db, frame_length, hop_length)
    return np.asarray(y)[s:e]


def tt_quantilize_c0(c0, c0min=4.6309418394230306e-05, c0max=0.3751049339771271, n_c0_bins=32):
    """transtacos/audio.py:124-128."""
    c0 = (np.asarray(c0) - c0min) / (c0max - c0min)
    c0 = c0 * n_c0_bins
    return c0.clip(0, n_c0_bins - 1).astype(np.int32)


def _tiny(x):
    return np.finfo(np.asarray(x).dtype if np.issubdtype(np.asarray(x).dtype, np.floating) else np.float32).tiny


def yin(y, fmin, fmax, sr=22050, frame_length=2048, win_length=None, hop_length=None, trough_threshold=0.1):
    """librosa.yin 0.8.1 (center=True, pad_mode='reflect'): FFT autocorrelation, cumulative-sum energy terms, cumulative
    mean normalised difference, parabolic interpolation, first trough below the threshold else the global minimum.
    Weak pin from the reference tree: transtacos/hparam.py:24-25 f0min / f0max equal sr / 301 and sr / 37, the period
    limits this restatement derives from rf0min = 'D2', rf0max = 'D5'."""
    y = np.asarray(y)
    win_length = frame_length // 2 if win_length is None else win_length
    hop_length = frame_length // 4 if hop_length is None else hop_length
    yp = np.pad(y, frame_length // 2, mode="reflect")
    yf = _frame(yp, frame_length, hop_length)                      # [frame_length, T]
    min_period = max(int(np.floor(sr / fmax)), 1)
    max_period = min(int(np.ceil(sr / fmin)), frame_length - win_length - 1)
    a = np.fft.rfft(yf, frame_length, axis=0)
    b = np.fft.rfft(yf[win_length:0:-1, :], frame_length, axis=0)
    acf = np.fft.irfft(a * b, frame_length, axis=0)[win_length:]
    acf[np.abs(acf) < 1e-6] = 0
    energy = np.cumsum(yf ** 2, axis=0)
    energy = energy[win_length:, :] - energy[:-win_length, :]
    energy[np.abs(energy) < 1e-6] = 0
    d = energy[0, :] + energy - 2 * acf
    num = d[min_period:max_period + 1, :]
    tau = np.arange(1, max_period + 1)[:, None]
    cm = np.cumsum(d[1:max_period + 1, :], axis=0) / tau
    den = cm[min_period - 1:max_period, :]
    yn = num / (den + _tiny(den))
    shifts = np.zeros_like(yn)
    pa = (yn[:-2, :] + yn[2:, :] - 2 * yn[1:-1, :]) / 2
    pb = (yn[2:, :] - yn[:-2, :]) / 2
    shifts[1:-1, :] = -pb / (2 * pa + _tiny(pa))
    shifts[np.abs(shifts) > 1] = 0
    neg = -yn
    xp = np.pad(neg, [(1, 1), (0, 0)], mode="edge")               # librosa.util.localmax(-yn, axis=0)
    trough = (neg > xp[:-2]) & (neg >= xp[2:])
    trough[0, :] = yn[0, :] < yn[1, :]
    thr = np.logical_and(trough, yn < trough_threshold)
    gmin = np.argmin(yn, axis=0)
    per = np.argmax(thr, axis=0)
    none = np.all(~thr, axis=0)
    per[none] = gmin[none]
    period = min_period + per + shifts[per, range(yn.shape[1])]
    return sr / period


def note_to_hz(note):
    """librosa.note_to_hz for a single note name (A4 = 440 Hz)."""
    pc = {'C': 0, 'D': 2, 'E': 4, 'F': 5, 'G': 7, 'A': 9, 'B': 11}[note[0].upper()]
    i = 1
    while i < len(note) and note[i] in '#b':
        pc += 1 if note[i] == '#' else -1
        i += 1
    midi = 12 * (int(note[i:]) + 1) + pc
    return 440.0 * 2.0 ** ((midi - 69) / 12.0)


def tt_get_f0(y, sr=22050, win_length=1024, hop_length=256, rf0min='D2', rf0max='D5'):
    """transtacos/audio.py:107-109."""
    return yin(np.asarray(y, np.float32), note_to_hz(rf0min), note_to_hz(rf0max), sr, win_length, None, hop_length).astype(np.float32)


def tt_quantilize_f0(f0, f0min=73.25581359863281, f0max=595.9459228515625):
    """transtacos/audio.py:14-21,117-121."""
    h2m = lambda f: 12 * (np.log2(f) - np.log2(440.0)) + 69
    n_min = int(np.floor(h2m(f0min)))
    n_bins = int(np.ceil(h2m(f0max))) - n_min + 1
    q = np.asarray([h2m(f) - n_min for f in f0])
    return q.clip(0, n_bins - 1).astype(np.int32)


def linear_basis(hp=HP):
    """transtacos/audio.py:167-175 _get_linear_basis: m^T diag(1 / colsum(m m^T)) (float32 basis, float64 result)."""
    m = mel_basis(hp=hp)
    m_T = np.transpose(m)
    p = np.matmul(m, m_T)
    d = [1.0 / x if np.abs(x) > 1.0e-8 else x for x in np.sum(p, axis=0)]
    return np.matmul(m_T, np.diag(d))


def tt_mel_to_linear(mel, hp=HP):
    """transtacos/audio.py:164-165."""
    return np.dot(linear_basis(hp), mel)


def tt_inv_mel(mel, hp=HP, init_phase=None, n_iter=None):
    """transtacos/audio.py:100-104."""
    M = tt_spec_to_natural_scale(mel, hp)
    S = tt_mel_to_linear(M, hp)
    wav = tt_inv_preemphasis(tt_griffin_lim(S ** hp.tt_gl_power, hp, init_phase, n_iter), hp)
    return wav.astype(np.float32)


def _maxpool1d(x, k):
    """nn.MaxPool1d(k) on [B, T]: stride k, no padding, floor -> ([B, T // k] values, first arg-max index inside each window)."""
    B, T = x.shape
    n = T // k
    w = x[:, :n * k].reshape(B, n, k)
    return w.max(axis=2), w.argmax(axis=2)


def rtg_envelope_loss(y, y_g, k=160):
    """retunegan/models/loss.py:66-72 on [B, T] float arrays."""
    y, y_g = np.asarray(y, np.float64), np.asarray(y_g, np.float64)
    return (np.mean(np.abs(_maxpool1d(y, k)[0] - _maxpool1d(y_g, k)[0])) +
            np.mean(np.abs(_maxpool1d(-y, k)[0] - _maxpool1d(-y_g, k)[0])))


def rtg_dynamic_loss(y, y_g, k=160):
    """retunegan/models/loss.py:76-82."""
    y, y_g = np.asarray(y, np.float64), np.asarray(y_g, np.float64)
    dyn = np.abs(_maxpool1d(y, k)[0] + _maxpool1d(-y, k)[0])
    dyn_g = np.abs(_maxpool1d(y_g, k)[0] + _maxpool1d(-y_g, k)[0])
    return np.mean(np.abs(dyn - dyn_g))


def rtg_pool_loss_backward(y, y_g, mode, k=160):
    """d loss / d y_g of envelope_loss (mode 0) / dynamic_loss (mode 1) in closed form (what torch autograd returns:
    max-pool gradients go to the first arg-max, abs has gradient sign(x))."""
    y, y_g = np.asarray(y, np.float64), np.asarray(y_g, np.float64)
    B, T = y_g.shape
    hi, _ = _maxpool1d(y, k); lo, _ = _maxpool1d(-y, k)
    hig, ih = _maxpool1d(y_g, k); log_, il = _maxpool1d(-y_g, k)
    n = hi.size
    g = np.zeros((B, T))
    bi, wi = np.meshgrid(np.arange(B), np.arange(hi.shape[1]), indexing="ij")
    if mode == 0:
        np.add.at(g, (bi, wi * k + ih), -np.sign(hi - hig) / n)
        np.add.at(g, (bi, wi * k + il), np.sign(lo - log_) / n)
    else:
        c = -np.sign(np.abs(hi + lo) - np.abs(hig + log_)) * np.sign(hig + log_) / n
        np.add.at(g, (bi, wi * k + ih), c)
        np.add.at(g, (bi, wi * k + il), -c)
    return g
