"""Host-side plumbing over the C ABI: plan cache, tensor marshalling, batch descriptors.

PyTorch is used only for device memory and streams; every arithmetic step of the path is a kernel in
libspectral_b200.so.  All spectrogram-shaped tensors handled here are FRAME-MAJOR ``[frames, F]``; the
reference-facing modules hand out ``[F, T]`` views of them (strides (1, F), the memory order librosa's
``order='F'`` STFT has).
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import Batch, Config, Scale, RAW, check
from .config import SpectralConfig

ArrayLike = Union[np.ndarray, torch.Tensor]

_plans = {}
_plans_lock = threading.Lock()


_cuda_ok = False
_devices = {}


def require_cuda() -> torch.device:
    """The current CUDA device.  (`torch.cuda.is_available()` is asked until it first says yes, and the `torch.device` objects are
    kept: on the host-bound training step these two calls were a tenth of the host time.)"""
    global _cuda_ok
    if not _cuda_ok:
        if not torch.cuda.is_available():
            raise RuntimeError("transtacos-retunegan_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _cuda_ok = True
    i = torch.cuda.current_device()
    d = _devices.get(i)
    if d is None:
        d = _devices[i] = torch.device("cuda", i)
    return d


def raw_stream(device_index: int) -> int:
    """cudaStream_t of torch's current stream on that device as an integer (no Stream object is built)."""
    return torch._C._cuda_getCurrentRawStream(device_index)


class Plan:
    """Opaque handle of the immutable tables of one (sr, n_fft, win, hop, n_mel, fmin, fmax, scale, window)."""

    def __init__(self, key, device_index: int):
        lib = _lib.load()
        cfg = Config(*key)
        h = C.c_void_p()
        with torch.cuda.device(device_index):
            check(lib.sb200_plan_create(C.byref(cfg), C.byref(h)), "plan_create")
        self.handle = h
        self.key = key
        self.device_index = device_index
        self.sample_rate, self.n_fft, self.win_length, self.hop_length, self.n_mel = key[:5]
        self.F = self.n_fft // 2 + 1
        self.Q = int(lib.sb200_plan_frames_per_pass(h))
        self._mel = None

    def mel_basis(self) -> np.ndarray:
        """float32 [n_mel, F], equal to librosa.filters.mel(sr, n_fft, n_mel, fmin, fmax)."""
        if self._mel is None:
            out = np.empty((self.n_mel, self.F), np.float32)
            check(_lib.load().sb200_plan_mel_basis_host(self.handle, out.ctypes.data_as(C.c_void_p)))
            self._mel = out
        return self._mel

    def window(self) -> np.ndarray:
        out = np.empty(self.win_length, np.float32)
        check(_lib.load().sb200_plan_window_host(self.handle, out.ctypes.data_as(C.c_void_p)))
        return out


def get_plan(cfg: SpectralConfig, n_fft=None, win_length=None, hop_length=None, htk=False) -> Plan:
    """Plan cache keyed by configuration AND device (the reference pins its caches to the first device seen,
    retunegan/audio.py:153-159).  ``htk=True`` only from ``get_mel`` (the one place the reference reads ``hp.mel_scale``)."""
    dev = require_cuda()
    key = cfg.plan_key(n_fft, win_length, hop_length, htk)
    ck = (key, dev.index)
    p = _plans.get(ck)
    if p is None:
        with _plans_lock:
            p = _plans.get(ck)
            if p is None:
                p = Plan(key, dev.index)
                _plans[ck] = p
    return p


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(raw_stream(torch.cuda.current_device()))


def ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def to_device_f32(x: ArrayLike, pinned_copy: bool = False) -> torch.Tensor:
    """numpy / torch (any device, any float dtype) -> contiguous float32 CUDA tensor on the current device."""
    dev = require_cuda()
    if isinstance(x, np.ndarray):
        if not np.isfinite(x).all():
            raise ValueError("Audio buffer is not finite everywhere")   # librosa.util.valid_audio (ParameterError)
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return t.to(dev, non_blocking=False)
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"expected numpy.ndarray or torch.Tensor, got {type(x).__name__}")
    return x.detach().to(device=dev, dtype=torch.float32).contiguous()


class SignalBatch:
    """Device-side description of a batch of utterances (uniform [B, L] or ragged list)."""

    def __init__(self, plan: Plan, ys):
        self.plan = plan
        hop, Q = plan.hop_length, plan.Q
        if isinstance(ys, (list, tuple)):
            parts = [to_device_f32(y).reshape(-1) for y in ys]
            if not parts:
                raise ValueError("empty batch")
            lens = np.array([p.numel() for p in parts], np.int64)
            if lens.min() < plan.n_fft // 4 + 1:
                raise ValueError("signal shorter than n_fft/4 + 1 samples: reflect padding undefined")
            self.x = torch.cat(parts)
            self.B = len(parts)
            self.lens = lens
            frames = 1 + lens // hop
            items = (frames + Q - 1) // Q
            tbl = np.zeros((4, self.B + 1), np.int64)
            tbl[0, 1:] = np.cumsum(lens)
            tbl[1, :-1] = lens
            tbl[2, 1:] = np.cumsum(frames)
            tbl[3, 1:] = np.cumsum(items)
            self.frames = frames
            self.total_frames = int(tbl[2, -1])
            self.tables = torch.from_numpy(tbl).to(self.x.device)
            self.c = Batch(self.B, 0, 0, self.tables[0].data_ptr(), self.tables[1].data_ptr(),
                           self.tables[2].data_ptr(), self.tables[3].data_ptr(), self.total_frames, int(tbl[3, -1]),
                           int(self.x.numel()))
            self.uniform = False
        else:
            x = to_device_f32(ys)
            if x.dim() == 1:
                x = x.unsqueeze(0)
            if x.dim() != 2:
                raise ValueError(f"expected [L] or [B, L] samples, got shape {tuple(x.shape)}")
            self.x = x
            self.B, L = x.shape
            if L < plan.n_fft // 4 + 1:
                raise ValueError("signal shorter than n_fft/4 + 1 samples: reflect padding undefined")
            T = 1 + L // hop
            self.lens = np.full(self.B, L, np.int64)
            self.frames = np.full(self.B, T, np.int64)
            self.total_frames = self.B * T
            self.c = Batch(self.B, L, L, None, None, None, None, 0, 0)
            self.uniform = True


def log_scale(a: float, b: float, floor: float) -> Scale:
    return Scale(1, a, b, floor)


def stft_features(plan: Plan, batch: SignalBatch, preemph: float = 0.0, mag_scale: Optional[Scale] = None,
                  mel_scale: Optional[Scale] = None, want_mag=True, want_mel=True, want_spec=False):
    """One fused launch.  Returns (mag [frames, F] | None, mel [frames, n_mel] | None, spec [frames, F] complex64 | None)."""
    dev = batch.x.device
    n = batch.total_frames
    mag = torch.empty((n, plan.F), device=dev, dtype=torch.float32) if want_mag else None
    mel = torch.empty((n, plan.n_mel), device=dev, dtype=torch.float32) if want_mel else None
    spec = torch.empty((n, plan.F, 2), device=dev, dtype=torch.float32) if want_spec else None
    check(_lib.load().sb200_stft_features(plan.handle, ptr(batch.x), C.byref(batch.c), float(preemph),
                                          mag_scale or RAW, mel_scale or RAW, ptr(mag), ptr(mel), ptr(spec),
                                          stream_ptr()), "stft_features")
    return mag, mel, (torch.view_as_complex(spec) if want_spec else None)


def mel_project(plan: Plan, x_fm: torch.Tensor, scale: Optional[Scale] = None) -> torch.Tensor:
    """[frames, F] float32 (frame-major) -> [frames, n_mel]."""
    assert x_fm.is_cuda and x_fm.dtype == torch.float32 and x_fm.is_contiguous() and x_fm.shape[-1] == plan.F
    frames = x_fm.numel() // plan.F
    out = torch.empty((frames, plan.n_mel), device=x_fm.device, dtype=torch.float32)
    check(_lib.load().sb200_mel_project(plan.handle, ptr(x_fm), frames, scale or RAW, ptr(out), stream_ptr()))
    return out


def mel_to_linear(plan: Plan, mel_fm: torch.Tensor) -> torch.Tensor:
    """[frames, n_mel] -> [frames, F]: the reference's pseudo-inverse basis product (transtacos/audio.py:164-175)."""
    mel_fm = mel_fm.contiguous()
    out = torch.empty((mel_fm.shape[0], plan.F), device=mel_fm.device, dtype=torch.float32)
    check(_lib.load().sb200_mel_to_linear(plan.handle, ptr(mel_fm), mel_fm.shape[0], ptr(out), stream_ptr()), "mel_to_linear")
    return out


def spec_to_amplitude(x: torch.Tensor, mode: int, p0=0.0, p1=0.0, p2=0.0, power=1.0) -> torch.Tensor:
    out = torch.empty_like(x)
    check(_lib.load().sb200_spec_to_amplitude(ptr(x), x.numel(), mode, float(p0), float(p1), float(p2), float(power),
                                              ptr(out), stream_ptr()))
    return out


def _rows_batch(x: torch.Tensor) -> Batch:
    B, L = x.shape
    return Batch(B, L, L, None, None, None, None, 0, 0)


def preemphasis(x: torch.Tensor, k: float) -> torch.Tensor:
    out = torch.empty_like(x)
    check(_lib.load().sb200_preemphasis(ptr(x), C.byref(_rows_batch(x)), float(k), ptr(out), stream_ptr()))
    return out


def inv_preemphasis(x: torch.Tensor, k: float) -> torch.Tensor:
    out = torch.empty_like(x)
    check(_lib.load().sb200_inv_preemphasis(ptr(x), C.byref(_rows_batch(x)), float(k), ptr(out), stream_ptr()))
    return out


def frame_stats(ys, frame_length: int, hop_length: int, want_rms=True, want_zcr=True):
    """Per-frame RMS and zero-crossing rate with librosa's centred framing (one launch of ``frame_stats_kernel``).

    ``ys``: [L] / [B, L] samples (numpy or torch) or a list of ragged utterances.  Returns
    ``(rms, zcr, frames)``: float32 CUDA tensors of ``sum(frames)`` values (None if not wanted) and the frame count of
    every row (``1 + len // hop_length``)."""
    require_cuda()
    if isinstance(ys, (list, tuple)):
        parts = [to_device_f32(y).reshape(-1) for y in ys]
        if not parts:
            raise ValueError("empty batch")
        lens = np.array([p.numel() for p in parts], np.int64)
        if lens.min() < 1:
            raise ValueError("empty signal")
        x = torch.cat(parts)
        frames = 1 + lens // hop_length
        tbl = np.zeros((3, len(parts) + 1), np.int64)
        tbl[0, 1:] = np.cumsum(lens)
        tbl[1, :-1] = lens
        tbl[2, 1:] = np.cumsum(frames)
        tables = torch.from_numpy(tbl).to(x.device)
        total = int(tbl[2, -1])
        c = Batch(len(parts), 0, 0, tables[0].data_ptr(), tables[1].data_ptr(), tables[2].data_ptr(), None, total, 0)
    else:
        x = to_device_f32(ys)
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.dim() != 2 or x.shape[1] < 1:
            raise ValueError(f"expected [L] or [B, L] samples, got shape {tuple(x.shape)}")
        x = x.contiguous()
        B, L = x.shape
        frames = np.full(B, 1 + L // hop_length, np.int64)
        total = int(frames.sum())
        tables = None
        c = Batch(B, L, L, None, None, None, None, 0, 0)
    rms = torch.empty(total, device=x.device, dtype=torch.float32) if want_rms else None
    zcr = torch.empty(total, device=x.device, dtype=torch.float32) if want_zcr else None
    check(_lib.load().sb200_frame_stats(ptr(x), C.byref(c), int(frame_length), int(hop_length), ptr(rms), ptr(zcr),
                                        stream_ptr()), "frame_stats")
    del tables   # (kept alive until the launch has been issued; the stream orders its use)
    return rms, zcr, frames


def yin(ys, sample_rate: int, fmin: float, fmax: float, frame_length: int, hop_length: int, trough_threshold: float = 0.1):
    """librosa.yin on [L] / [B, L] / a list of utterances (one launch of ``yin_kernel``): (f0 float32 CUDA [sum frames], frames)."""
    require_cuda()
    if isinstance(ys, (list, tuple)):
        parts = [to_device_f32(y).reshape(-1) for y in ys]
        if not parts:
            raise ValueError("empty batch")
        lens = np.array([p.numel() for p in parts], np.int64)
        if lens.min() < frame_length // 2 + 1:
            raise ValueError("signal shorter than frame_length/2 + 1 samples: reflect padding undefined")
        x = torch.cat(parts)
        frames = 1 + lens // hop_length
        tbl = np.zeros((3, len(parts) + 1), np.int64)
        tbl[0, 1:] = np.cumsum(lens)
        tbl[1, :-1] = lens
        tbl[2, 1:] = np.cumsum(frames)
        tables = torch.from_numpy(tbl).to(x.device)
        total = int(tbl[2, -1])
        c = Batch(len(parts), 0, 0, tables[0].data_ptr(), tables[1].data_ptr(), tables[2].data_ptr(), None, total, 0)
    else:
        x = to_device_f32(ys)
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.dim() != 2 or x.shape[1] < 1:
            raise ValueError(f"expected [L] or [B, L] samples, got shape {tuple(x.shape)}")
        if x.shape[1] < frame_length // 2 + 1:
            raise ValueError("signal shorter than frame_length/2 + 1 samples: reflect padding undefined")
        x = x.contiguous()
        B, L = x.shape
        frames = np.full(B, 1 + L // hop_length, np.int64)
        total = int(frames.sum())
        tables = None
        c = Batch(B, L, L, None, None, None, None, 0, 0)
    f0 = torch.empty(total, device=x.device, dtype=torch.float32)
    check(_lib.load().sb200_yin(ptr(x), C.byref(c), int(sample_rate), float(fmin), float(fmax), int(frame_length),
                                int(hop_length), float(trough_threshold), ptr(f0), stream_ptr()), "yin")
    del tables
    return f0, frames


def trim_bounds(rms: torch.Tensor, frames, lens, hop_length: int, top_db: float):
    """librosa.effects.trim on a precomputed RMS track (ref = max, amin = 1e-10): [(start, end)] sample bounds per row.
    The threshold and the first / last scan run on the device (``trim_bounds_kernel``, one warp per row); only the two
    frame indices per row come back to the host (slicing host arrays needs host integers)."""
    frames = np.asarray(frames, np.int64)
    B = len(frames)
    off = np.zeros(B + 1, np.int64)
    off[1:] = np.cumsum(frames)
    off_d = torch.from_numpy(off).to(rms.device)
    out_d = torch.empty((B, 2), device=rms.device, dtype=torch.int64)
    check(_lib.load().sb200_trim_bounds(ptr(rms), ptr(off_d), 0, B, float(top_db), ptr(out_d), stream_ptr()), "trim_bounds")
    fl = out_d.cpu().numpy()
    return [(int(f) * hop_length, min(int(L), int(l) * hop_length)) for (f, l), L in zip(fl, lens)]


class FramesBatch:
    """Batch of spectrograms described by frame counts (ISTFT / Griffin-Lim direction)."""

    def __init__(self, plan: Plan, frames: Sequence[int], lengths: Optional[Sequence[int]], device):
        frames = np.asarray(frames, np.int64)
        hop, Q = plan.hop_length, plan.Q
        self.frames = frames
        self.B = len(frames)
        out_len = np.asarray(lengths, np.int64) if lengths is not None else hop * (frames - 1)
        self.out_len = out_len
        self.has_length = lengths is not None
        if lengths is not None and np.any(1 + out_len // hop != frames):
            raise ValueError("length inconsistent with the number of frames: need 1 + length // hop == n_frames")
        if out_len.min() < plan.n_fft // 4 + 1:
            raise ValueError("output signal shorter than n_fft/4 + 1 samples")
        self.total_frames = int(frames.sum())
        if np.all(frames == frames[0]) and np.all(out_len == out_len[0]):
            self.uniform = True
            self.c = Batch(self.B, int(frames[0]), int(out_len[0]), None, None, None, None, 0, 0)
            self.length_arg = int(out_len[0]) if lengths is not None else 0
            self.total_out = int(out_len[0]) * self.B
            self.out_off = np.arange(self.B + 1, dtype=np.int64) * int(out_len[0])
        else:
            self.uniform = False
            items = (frames + Q - 1) // Q
            tbl = np.zeros((4, self.B + 1), np.int64)
            tbl[0, 1:] = np.cumsum(out_len)
            tbl[1, :-1] = out_len
            tbl[2, 1:] = np.cumsum(frames)
            tbl[3, 1:] = np.cumsum(items)
            self.tables = torch.from_numpy(tbl).to(device)
            self.c = Batch(self.B, int(out_len.max()), 0, self.tables[0].data_ptr(), self.tables[1].data_ptr(),
                           self.tables[2].data_ptr(), self.tables[3].data_ptr(), self.total_frames, int(tbl[3, -1]))
            # ragged rows always carry explicit lengths (hop*(T-1) when none was asked for: same value librosa returns)
            self.length_arg = 1
            self.total_out = int(tbl[0, -1])
            self.out_off = tbl[0].copy()


_tls = threading.local()


def _workspace(nbytes: int, device, tag: str) -> torch.Tensor:
    """Grow-only scratch buffer (the caller-allocated workspace of the C ABI), private to the calling THREAD and to the
    CUDA STREAM the call is enqueued on.  A workspace holds state that lives across the kernels of one call (Griffin-Lim
    signal buffers and the grid-barrier counter of the persistent kernel, mstft gradient frames and partial sums): two
    host threads interleaving their enqueues on one stream, or two streams running at once, must never share it.  Flask
    request threads (retunegan/server.py:33-62, transtacos/server.py:59-101) each get their own; it dies with the thread."""
    cache = getattr(_tls, "ws", None)
    if cache is None:
        cache = _tls.ws = {}
    key = (device.index, raw_stream(device.index), tag)
    buf = cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), device=device, dtype=torch.uint8)
        cache[key] = buf
    return buf


def istft(plan: Plan, spec_fm: torch.Tensor, fb: FramesBatch) -> torch.Tensor:
    """spec_fm: complex64 [total_frames, F] frame-major -> flat float32 [sum out_len]."""
    lib = _lib.load()
    sp = torch.view_as_real(spec_fm.contiguous())
    y = torch.empty(fb.total_out, device=spec_fm.device, dtype=torch.float32)
    ws = _workspace(lib.sb200_griffinlim_workspace_bytes(plan.handle, fb.total_frames, fb.B, 0), spec_fm.device, "gl")
    check(lib.sb200_istft(plan.handle, ptr(sp), C.byref(fb.c), fb.length_arg, ptr(y), ptr(ws), stream_ptr()), "istft")
    return y


def griffinlim(plan: Plan, S_fm: torch.Tensor, phase_fm: torch.Tensor, fb: FramesBatch, n_iter: int, momentum: float,
               form: int, inv_preemph: float = 0.0) -> torch.Tensor:
    """S_fm, phase_fm: float32 [total_frames, F] -> flat float32 [sum out_len]."""
    lib = _lib.load()
    assert S_fm.is_contiguous() and phase_fm.is_contiguous() and S_fm.shape == phase_fm.shape == (fb.total_frames, plan.F)
    y = torch.empty(fb.total_out, device=S_fm.device, dtype=torch.float32)
    ws = _workspace(lib.sb200_griffinlim_workspace_bytes(plan.handle, fb.total_frames, fb.B, form), S_fm.device, "gl")
    pre_in_call = inv_preemph if fb.uniform else 0.0
    check(lib.sb200_griffinlim(plan.handle, ptr(S_fm), ptr(phase_fm), C.byref(fb.c), fb.length_arg, int(n_iter),
                               float(momentum), int(form), float(pre_in_call), ptr(y), ptr(ws), stream_ptr()),
          "griffinlim")
    if inv_preemph and not fb.uniform:
        b = Batch(fb.B, 0, 0, fb.tables[0].data_ptr(), fb.tables[1].data_ptr(), None, None, 0, 0)
        check(lib.sb200_inv_preemphasis(ptr(y), C.byref(b), float(inv_preemph), ptr(y), stream_ptr()))
    return y


class HostFeaturePipeline:
    """Host-resident batches: overlap H2D copies, the fused STFT+mel launch and D2H copies chunk by chunk.

    For a uniform ``[B, L]`` batch living in (ideally pinned) host memory.  Three CUDA streams and a ring of
    ``depth`` device slots; every chunk is one ``sb200_stft_features`` launch.  This is what the numpy-facing
    ``get_specs`` / ``get_mag`` use for 2-D host input, and what bench.py times as the end-to-end number.
    """

    def __init__(self, plan: Plan, chunk: int = 16, depth: int = 3):
        self.plan, self.chunk, self.depth = plan, int(chunk), int(depth)
        self.dev = require_cuda()
        self.L = 0                      # capacity of the slots in samples per row; grow-only
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.x, self.mag, self.mel = [], [], []
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.used = [False] * depth
        self.lock = threading.Lock()    # one batch at a time per device: the slots, streams and events are shared state

    def _reserve(self, L: int) -> None:
        """Slots sized for the longest row seen so far (a corpus / DataLoader changes the padded length every step: one set
        of buffers, streams and events serves every length, nothing is kept per length)."""
        if L <= self.L:
            return
        torch.cuda.synchronize(self.dev)            # nothing of the old slots is in flight any more
        plan, chunk, dev = self.plan, self.chunk, self.dev
        T = 1 + L // plan.hop_length
        self.x = self.mag = self.mel = None         # free before allocating the larger set
        self.x = [torch.empty(chunk * L, device=dev, dtype=torch.float32) for _ in range(self.depth)]
        self.mag = [torch.empty(chunk * T * plan.F, device=dev, dtype=torch.float32) for _ in range(self.depth)]
        self.mel = [torch.empty(chunk * T * plan.n_mel, device=dev, dtype=torch.float32) for _ in range(self.depth)]
        self.used = [False] * self.depth
        self.L = L

    def run(self, y_host: torch.Tensor, mag_host: Optional[torch.Tensor], mel_host: Optional[torch.Tensor],
            preemph: float, mag_scale: Scale, mel_scale: Scale) -> None:
        """y_host [B, L] float32 CPU; mag_host [B*T, F] / mel_host [B*T, n_mel] float32 CPU (None = not wanted).
        Returns after all copies have completed.  Thread-safe (serialised per device)."""
        with self.lock:
            self._run(y_host, mag_host, mel_host, preemph, mag_scale, mel_scale)

    def _run(self, y_host, mag_host, mel_host, preemph, mag_scale, mel_scale) -> None:
        lib = _lib.load()
        B, L = y_host.shape
        self._reserve(int(L))
        plan = self.plan
        T = 1 + L // plan.hop_length
        xs = [b[:self.chunk * L].view(self.chunk, L) for b in self.x]
        mags = [b[:self.chunk * T * plan.F].view(self.chunk * T, plan.F) for b in self.mag]
        mels = [b[:self.chunk * T * plan.n_mel].view(self.chunk * T, plan.n_mel) for b in self.mel]
        cur = torch.cuda.current_stream()
        self.s_in.wait_stream(cur)
        # ramp-up schedule: a quarter and a half chunk first (the copy-back, which bounds the pipeline, starts early), full
        # chunks afterwards (few launches / cross-stream waits)
        sched, b0 = [], 0
        for n in (max(1, self.chunk // 4), max(1, self.chunk // 2)):
            if b0 < B:
                sched.append((b0, min(n, B - b0)))
                b0 += sched[-1][1]
        while b0 < B:
            sched.append((b0, min(self.chunk, B - b0)))
            b0 += sched[-1][1]
        for c, (b0, n) in enumerate(sched):
            s = c % self.depth
            with torch.cuda.stream(self.s_in):
                if self.used[s]:
                    self.s_in.wait_event(self.ev_run[s])          # slot's previous launch has consumed x[s]
                xs[s][:n].copy_(y_host[b0:b0 + n], non_blocking=True)
                self.ev_in[s].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[s])
                if self.used[s]:
                    self.s_run.wait_event(self.ev_out[s])         # slot's previous outputs have left the device
                bc = Batch(n, L, L, None, None, None, None, 0, 0)
                check(lib.sb200_stft_features(plan.handle, ptr(xs[s]), C.byref(bc), float(preemph), mag_scale,
                                              mel_scale, ptr(mags[s]) if mag_host is not None else None,
                                              ptr(mels[s]) if mel_host is not None else None, None,
                                              C.c_void_p(self.s_run.cuda_stream)), "stft_features")
                self.ev_run[s].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[s])
                if mag_host is not None:
                    mag_host[b0 * T:(b0 + n) * T].copy_(mags[s][:n * T], non_blocking=True)
                if mel_host is not None:
                    mel_host[b0 * T:(b0 + n) * T].copy_(mels[s][:n * T], non_blocking=True)
                self.ev_out[s].record(self.s_out)
            self.used[s] = True
        self.s_out.synchronize()
        cur.wait_stream(self.s_out)


_pipelines = {}
_pipelines_lock = threading.Lock()


def host_feature_pipeline(plan: Plan, chunk: int = 16) -> HostFeaturePipeline:
    """One pipeline per (plan, device, chunk size); its slots grow to the longest row seen (no entry per length)."""
    key = (plan.key, plan.device_index, int(chunk))
    with _pipelines_lock:
        p = _pipelines.get(key)
        if p is None:
            p = HostFeaturePipeline(plan, chunk)
            _pipelines[key] = p
    return p
