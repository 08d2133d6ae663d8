#!/usr/bin/env python3
"""bench.py -- spectrogram-seconds per second of the spectral hot path on N B200s (one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stft_mel|griffinlim|griffinlim_batch|mstft|corpus ...]
                    [--impl reference] [--no-extra]

Default workload = BASELINE.json configs[2]: batched STFT + mel feature extraction (TransTacoS get_specs:
pre-emphasis, dB-normalise), 64 synthetic 5 s utterances per GPU, hparam.py shapes.  A "step" is one pass of the
hot path over one such batch.  `value` is timed with the batch resident in HBM (CUDA events on the launch
stream, max over ranks); `e2e` goes through the numpy/CPU-tensor-facing public API with pinned HOST buffers,
host<->device copies inside the timed region.  Utterances shard across ranks with no data-path collective (weak
scaling: every rank owns 64 utterances).  `--impl reference` times the CPU oracle restatement of the reference's
librosa/numpy path (the reference itself cannot be installed: librosa / TF are absent and there is no network) on
all host cores, the way the reference parallelises it (process pool over utterances, databaker.py:31).

`extra` (every N, not only N = 1) carries the other BASELINE.json configs on the same box in the same run, each timed
between barriers as the max over ranks, with its own clocks, HBM and FP32-pipe roofline fractions: Griffin-Lim (single
utterance both forms, 64-utterance batch, weak), the multi-resolution STFT loss forward + backward (loss-only and training
variant; under torchrun with the loss all-reduce inside the step and DDP's 11 MB gradient all-reduce timed beside it) and the
10 000-utterance corpus (STRONG scaling: the corpus is sharded over the ranks, per-rank load max / min reported).
Before anything is timed one row of every workload's output is compared with the CPU oracle (the checker, not timed).
"""
import argparse
import json
import os
import sys
import threading
import time

# one BLAS/OpenMP thread per worker process, as BASELINE.md section 3 prescribes for the CPU arm (the process pool
# supplies the parallelism); must be set before numpy is imported.  The torch CPU leg sets its own thread count.
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, N_FFT, HOP, WIN, N_MEL = 22050, 2048, 256, 1024, 80
F = N_FFT // 2 + 1
L5 = 431 * HOP - 1          # "5 s utterance": y[:-1] of 431 hops -> 431 frames
T5 = 431
METRIC = "spectrogram-seconds/sec (STFT+mel, Griffin-Lim) at 1/2/4/8 B200; % HBM roofline"
UNIT = "spectrogram-seconds/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- CPU reference arm ----

def _cpu_worker(args):
    kind, seed, L = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import spectral_oracle as O
    y = O.synth_noise(L, seed)
    if kind == "stft_mel":
        O.tt_get_specs(y)
    elif kind.startswith("griffinlim"):
        S = np.abs(O.stft(y, N_FFT, HOP, WIN)).astype(np.float32)
        O.rtg_griffinlim(S, wavlen=L)
    return L


def cpu_reference_rate(kind, n_utt, L, cores):
    """Audio seconds per wall second of the oracle on `cores` worker processes (bounded sample)."""
    from concurrent.futures import ProcessPoolExecutor
    if kind == "mstft":
        import torch
        from oracle import spectral_oracle as O   # noqa: F401
        torch.set_num_threads(cores)
        y = torch.randn(16, 22050) * 0.1
        yg = torch.tanh(y + 0.01 * torch.randn_like(y)).requires_grad_(True)
        t0 = time.perf_counter()
        reps = max(1, n_utt // 16)
        for _ in range(reps):
            _torch_ref_mstft(y, yg).backward()
        dt = time.perf_counter() - t0
        return reps * 16 * 22050 / SR / dt, dt
    jobs = [(kind, 114514 + i, L) for i in range(n_utt)]
    with ProcessPoolExecutor(max_workers=cores) as ex:
        list(ex.map(_cpu_worker, jobs[:cores]))          # warm the workers (imports)
        t0 = time.perf_counter()
        tot = sum(ex.map(_cpu_worker, jobs))
        dt = time.perf_counter() - t0
    return tot / SR / dt, dt


def _torch_ref_mstft(y, yg):
    """The reference's multi_stft_loss graph (retunegan/models/loss.py:22-62) in plain torch, on the device of its inputs
    (CPU for the reference arm; on the GPU it is the torch / cuFFT comparator of SURVEY.md 8d)."""
    import torch
    from oracle import spectral_oracle as O
    loss = 0
    for n_fft, win, hop in O.HP.multi_stft_params:
        w = torch.hann_window(win, device=y.device)
        mb = torch.from_numpy(O.mel_basis(n_fft)).to(y.device)

        def f(x):
            D = torch.stft(x, n_fft, hop, win, window=w, center=True, pad_mode="reflect", return_complex=True)
            return mb @ torch.abs(D + 1e-9)
        M, Mg = f(y), f(yg)
        loss = loss + (M - Mg).abs().mean() + (M.log() - Mg.log()).abs().mean()
    return loss / 3


# ----------------------------------------------------------------------------- clocks ---------------

class ClockSampler(threading.Thread):
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.ok = index, [], set(), False, False
        self.sm_max = None
        self.active = False
        # NVML queries take driver locks that kernel launches also need: poll sparsely (the timed loop is tens of milliseconds
        # of back-to-back launches; the continuation loop in measure() supplies the samples under load)
        self.period = float(os.environ.get("SB200_BENCH_SAMPLER_MS", "20")) / 1e3
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                if self.active:
                    self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    m = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if m & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, window):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "window": window}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "window": window}


# ----------------------------------------------------------------------------- workloads ------------

class Workload:
    """step(i): enqueue one pass on the current stream (device-resident rotating buffers).
    e2e(i): one pass through the public API from pinned host buffers and back."""
    name = ""
    units = 0.0            # audio seconds per step per rank
    alg_bytes = 0.0        # algorithmic HBM bytes of the dominant kernel per launch
    dominant = ""
    launches_dominant_per_step = 1
    h2d = d2h = 0
    note = ""
    flops = 0.0            # algorithmic FP32 FLOPs per step per rank
    load = None            # corpus: audio seconds of this rank's shard


def make_stft_mel(sb, torch, B=64, rot=6):
    w = Workload()
    w.name = f"stft_mel_{B}x5s"
    ta = sb.transtacos_audio
    plan = sb.core.get_plan(ta.hp)
    g = torch.Generator(device="cuda").manual_seed(114514 + int(os.environ.get("RANK", 0)))
    ys = [(0.1 * torch.randn(B, L5, device="cuda", generator=g)).clamp_(-0.999, 0.999) for _ in range(rot)]
    batches = [sb.core.SignalBatch(plan, y) for y in ys]
    mags = [torch.empty((B * T5, F), device="cuda") for _ in range(rot)]
    mels = [torch.empty((B * T5, N_MEL), device="cuda") for _ in range(rot)]
    sc = ta.db_norm_scale(ta.hp)
    lib = sb._lib.load()
    import ctypes as C

    def step(i):
        j = i % rot
        sb._lib.check(lib.sb200_stft_features(plan.handle, sb.core.ptr(batches[j].x), C.byref(batches[j].c),
                                              float(ta.hp.preemphasis), sc, sc, sb.core.ptr(mags[j]),
                                              sb.core.ptr(mels[j]), None, sb.core.stream_ptr()))
    yh = [torch.empty((B, L5), dtype=torch.float32).pin_memory() for _ in range(2)]
    for t in yh:
        t.copy_(ys[0].cpu())
    out_h = [(torch.empty((B * T5, F), dtype=torch.float32).pin_memory(),
              torch.empty((B * T5, N_MEL), dtype=torch.float32).pin_memory()) for _ in range(2)]

    def e2e(i):
        S, M = ta.get_specs(yh[i % 2], out=out_h[i % 2])     # public API: CPU tensor in -> CPU tensors out
        return S
    w.step, w.e2e = step, e2e
    w.units = B * L5 / SR
    w.alg_bytes = B * (4 * L5 + 4 * T5 * (F + N_MEL))          # 4L + 4T(F+M) per utterance
    w.dominant = "stft_feature3_kernel<2048,true,4>"
    w.h2d = B * L5 * 4
    w.d2h = B * T5 * (F + N_MEL) * 4
    w.note = (f"{rot} rotating input/output sets ({rot * (w.alg_bytes) / 1e6:.0f} MB) > 126 MB L2 between reuses; "
              "TransTacoS epilogue (pre-emphasis 0.97, dB-normalise)")
    w.flops = B * T5 * 66e3                                       # ~66 kFLOP per frame

    def check():
        """rows 0 and B-1 of the exact timed launch against the oracle (tolerance of the parity tests: 1e-4 relative)."""
        from oracle import spectral_oracle as O
        step(0)
        torch.cuda.synchronize()
        for b in (0, B - 1):
            So, Mo = O.tt_get_specs(ys[0][b].cpu().numpy())
            S = mags[0].view(B, T5, F)[b].t().cpu().numpy()
            M = mels[0].view(B, T5, N_MEL)[b].t().cpu().numpy()
            e1, e2 = np.linalg.norm(S - So) / np.linalg.norm(So), np.linalg.norm(M - Mo) / np.linalg.norm(Mo)
            if not (e1 < 1e-4 and e2 < 1e-4):
                return False, f"row {b}: rel-Frobenius mag {e1:.2e} mel {e2:.2e}"
        return True, "rows 0 and B-1 of the timed launch vs oracle tt_get_specs: rel-Frobenius < 1e-4"
    w.check = check
    return w


def make_griffinlim(sb, torch, B=1, form="rtg", rot=4):
    w = Workload()
    ra = sb.retunegan_audio if form == "rtg" else sb.transtacos_audio
    cfg = ra.hp
    plan = sb.core.get_plan(cfg)
    n_iter, mom, frm = (cfg.gl_iters, cfg.gl_momentum, 1) if form == "rtg" else (cfg.gl_iters, 0.0, 0)
    w.name = f"griffinlim_{form}_{B}x5s_{n_iter}it"
    g = torch.Generator(device="cuda").manual_seed(1234 + int(os.environ.get("RANK", 0)))
    y = (0.1 * torch.randn(B, L5, device="cuda", generator=g)).clamp_(-0.999, 0.999)
    S = sb.core.stft_features(plan, sb.core.SignalBatch(plan, y), want_mel=False)[0]
    Ss = [S.clone() for _ in range(rot)]
    ph = torch.rand((B * T5, F), device="cuda", generator=g)
    fb = sb.core.FramesBatch(plan, [T5] * B, [L5] * B if form == "rtg" else None, y.device)

    def step(i):
        return sb.core.griffinlim(plan, Ss[i % rot], ph, fb, n_iter, mom, frm, 0.0)
    Sh = torch.log(S.clamp_min(1e-5)).view(B, T5, F)[0].t().cpu().numpy() if form == "rtg" else None

    def e2e(i):
        return ra.inv_mag(Sh, wavlen=L5) if form == "rtg" else None
    w.step, w.e2e = step, (e2e if (B == 1 and form == "rtg") else None)
    w.units = B * L5 / SR
    per_iter = B * (4 * F * T5 + 8 * L5 + (16 * F * T5 if frm == 1 else 0))   # SURVEY.md 8d streaming model
    w.alg_bytes = per_iter
    w.dominant = f"gl2_kernel<2048,{2 + frm}>"
    w.launches_dominant_per_step = n_iter
    w.h2d, w.d2h = F * T5 * 4, L5 * 4
    w.note = f"{'fast form, momentum 0.7' if frm else 'angle form'}; {n_iter} iterations + initial/final ISTFT; state L2/HBM resident"
    w.flops = B * T5 * 134e3 * (n_iter + 0.5)                    # ~134 kFLOP per frame and iteration

    def check():
        """row 0 of the timed call against the oracle, same initial phase: waveform rel-L2 <= 1e-3 (parity-test tolerance)."""
        from oracle import spectral_oracle as O
        out = step(0).view(B, -1)[0].cpu().numpy()
        S0 = Ss[0].view(B, T5, F)[0].t().cpu().numpy().astype(np.float64)
        p0 = ph.view(B, T5, F)[0].t().cpu().numpy().astype(np.float64)
        if form == "rtg":
            ref = O.griffinlim(S0, n_iter=n_iter, hop_length=HOP, win_length=WIN, length=L5, momentum=mom,
                               init_angles=np.exp(2j * np.pi * p0))
        else:
            ref = O.tt_griffin_lim(S0, init_phase=p0, n_iter=n_iter)
        e = np.linalg.norm(out - ref) / np.linalg.norm(ref)
        return bool(e < 1e-3), f"row 0 vs oracle Griffin-Lim ({n_iter} it), waveform rel-L2 {e:.2e}"
    w.check = check
    return w


def graph_replay_ms(torch, step, n=100):
    """ms per replay of step(0) captured as one CUDA graph (the step must be capturable: no host synchronisation inside)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step(0)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step(0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del g
    return e0.elapsed_time(e1) / n


def make_mstft(sb, torch, B=16, T=22050, specs=False, rot=4, ddp=False):
    w = Workload()
    w.name = f"mstft_fwd_bwd_{B}x{T}" + ("_specs" if specs else "_lossonly")
    g = torch.Generator(device="cuda").manual_seed(77 + int(os.environ.get("RANK", 0)))
    ys = [(0.1 * torch.randn(B, 1, T, device="cuda", generator=g)).clamp_(-0.999, 0.999) for _ in range(rot)]
    # y_g = tanh(.) leaf.  A gain of 1.1 keeps the generated mel cells systematically apart from the
    # real ones: the L1 gradient sign(M_g - M) is discontinuous at ties, and with y_g = y + small noise ~1e-4 of the 824 k
    # cells tie to within float32 rounding, which no float32 implementation (the reference's included) resolves like float64
    ygs = [torch.tanh(1.1 * y).requires_grad_(True) for y in ys]
    ups = None
    if specs:
        hp = sb.loss.hp
        ups = [torch.randn(B, 2, n // 2 + 1, 1 + T // h, device="cuda", generator=g).transpose(2, 3).contiguous().transpose(2, 3) * 1e-3
               for n, _, h in hp.multi_stft_params]

    def step(i):
        j = i % rot
        ygs[j].grad = None
        if specs:
            loss, (sr, sg) = sb.multi_stft_loss(ys[j], ygs[j], ret_loss=True, ret_specs=True, ddp_reduce=ddp)
            torch.autograd.backward([loss] + list(sg), [torch.ones_like(loss)] + ups)
        else:
            sb.multi_stft_loss(ys[j], ygs[j], ret_loss=True, ddp_reduce=ddp).backward()
        return ygs[j].grad
    w.step, w.e2e = step, None
    w.units = B * T / SR
    w.alg_bytes = (4.23e6 if not specs else 113e6) * (B / 16) * (T / 22050)
    w.dominant = "mstft_fwd_kernel+mstft_bwd_kernel (3 resolutions)"
    w.note = ("loss-only" if not specs else "training variant: spec stacks written, dense upstream spec grads") + \
        ("; the reported loss is averaged over the ranks inside the step (multi_stft_loss(ddp_reduce=True): " +
         sb.loss.ddp_reduce_path() + ")" if ddp else "")
    w.flops = (0.49e9 + 0.25e9) * (B / 16) * (T / 22050)

    def check():
        """loss value and d loss / d y_g of the timed batch against the oracle (1e-5 / 1e-4, the parity-test tolerances)."""
        from oracle import spectral_oracle as O
        ygs[0].grad = None
        loss = sb.multi_stft_loss(ys[0], ygs[0], ret_loss=True)
        loss.backward()
        yn, gn = ys[0].cpu().numpy(), ygs[0].detach().cpu().numpy()
        lo = O.rtg_multi_stft_loss(yn, gn, ret_loss=True)
        R = min(4, B)      # the closed-form gradient is per row: the first rows, rescaled to the batch's 1 / B
        go = O.rtg_multi_stft_loss_backward(yn[:R], gn[:R], g_loss=R / B)
        gt, nt = O.rtg_multi_stft_loss_backward(yn[:R], gn[:R], g_loss=R / B, tie_rel=1e-5)
        e1 = abs(loss.item() - lo) / abs(lo)
        e2 = np.linalg.norm(ygs[0].grad[:R, 0].cpu().numpy() - go) / np.linalg.norm(go)
        slack = 2 * np.linalg.norm(go - gt) / np.linalg.norm(go)      # what hinges on mel cells tied to within 1e-5 (sign of |.|)
        ygs[0].grad = None
        return bool(e1 < 1e-5 and e2 < 1e-4 + slack), (f"loss rel {e1:.1e}, grad rel-L2 {e2:.1e} vs oracle (rows 0..{R - 1}; "
                                                        f"{nt} near-tie cells allow {slack:.1e})")
    w.check = check
    return w


def corpus_lengths(n=10000):
    """frame counts T_i = clip(round(N(307, 100)), 101, 524) from RandomState(114514),
    L_i = 256 T_i - 1 (min / mean / max of stats/DataBaker.stats:6-11)."""
    rs = np.random.RandomState(114514)
    T = np.clip(np.round(rs.normal(307, 100, n)), 101, 524).astype(np.int64)
    return T, 256 * T - 1


def make_corpus(sb, torch, rank, world, n_utt=10000, chunk=256, d2h=False):
    """BASELINE.json configs[4]: corpus-scale preprocessing, 10k synthetic utterances sharded over the ranks (length-
    balanced, no collective).  Per utterance, as retunegan/data.py:60-76 + transtacos get_specs: dB-normalised linear + mel
    features (A3) and the Griffin-Lim reference wav (R4: ln-magnitude -> exp -> ^1.2 -> 4 iterations, momentum 0.7).
    A step is one pass over this rank's shard, in ragged chunks of `chunk` utterances."""
    import ctypes as C
    w = Workload()
    w.name = f"corpus_{n_utt}utt_specs+griffinlim"
    ta, ra = sb.transtacos_audio, sb.retunegan_audio
    plan = sb.core.get_plan(ta.hp)
    T_all, L_all = corpus_lengths(n_utt)
    mine = sorted(sb.sharding.shard_utterances(L_all, world)[rank])
    T, L = T_all[mine], L_all[mine]
    g = torch.Generator(device="cuda").manual_seed(114514 + rank)
    lib = sb._lib.load()
    sc_db, sc_ln = ta.db_norm_scale(ta.hp), ra.ln_scale(True)
    chunks = []
    for c0 in range(0, len(mine), chunk):
        Lc, Tc = L[c0:c0 + chunk], T[c0:c0 + chunk]
        ys = [(0.1 * torch.randn(int(l), device="cuda", generator=g)).clamp_(-0.999, 0.999) for l in Lc]
        chunks.append((sb.core.SignalBatch(plan, ys), sb.core.FramesBatch(plan, Tc, Lc, torch.device("cuda")), None))
    fmax = max(b.total_frames for b, _, _ in chunks)
    lmax = max(int(b.x.numel()) for b, _, _ in chunks)
    nbuf = 2 if d2h else 1
    mags = [torch.empty((fmax, F), device="cuda") for _ in range(nbuf)]
    mels = [torch.empty((fmax, N_MEL), device="cuda") for _ in range(nbuf)]
    lnm = torch.empty((fmax, F), device="cuda")
    phase = torch.rand((fmax, F), device="cuda", generator=g)      # throughput mode: device RNG, drawn once
    if d2h:   # features and reference wavs leave the device chunk by chunk (double-buffered, copy stream, pinned host ring)
        h_mag = [torch.empty((fmax, F), pin_memory=True) for _ in range(2)]
        h_mel = [torch.empty((fmax, N_MEL), pin_memory=True) for _ in range(2)]
        h_wav = [torch.empty(lmax, pin_memory=True) for _ in range(2)]
        s_out = torch.cuda.Stream()
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_used = [False, False]

    def step(i):
        out = None
        cur = torch.cuda.current_stream()
        for ci, (batch, Tc, Lc) in enumerate(chunks):
            n = batch.total_frames
            b = ci % nbuf
            if d2h and ev_used[b]:
                cur.wait_event(ev_done[b])          # the copy of the chunk that used this buffer pair has finished
            mag, mel = mags[b], mels[b]
            sb._lib.check(lib.sb200_stft_features(plan.handle, sb.core.ptr(batch.x), C.byref(batch.c), float(ta.hp.preemphasis),
                                                  sc_db, sc_db, sb.core.ptr(mag), sb.core.ptr(mel), None, sb.core.stream_ptr()))
            sb._lib.check(lib.sb200_stft_features(plan.handle, sb.core.ptr(batch.x), C.byref(batch.c), 0.0,
                                                  sc_ln, sb.core.RAW, sb.core.ptr(lnm), None, None, sb.core.stream_ptr()))
            out, _ = ra.inv_mag_batch(lnm[:n], Tc, Lc, init_phase=phase[:n])
            if d2h:
                s_out.wait_stream(cur)
                with torch.cuda.stream(s_out):
                    h_mag[b][:n].copy_(mag[:n], non_blocking=True)
                    h_mel[b][:n].copy_(mel[:n], non_blocking=True)
                    h_wav[b][:out.numel()].copy_(out, non_blocking=True)
                    out.record_stream(s_out)
                    ev_done[b].record(s_out)
                ev_used[b] = True
        if d2h:
            cur.wait_stream(s_out)
        return out
    w.step, w.e2e = step, None
    w.units = float(L.sum()) / SR
    nf = int(T.sum())
    w.alg_bytes = 4.0 * L.sum() * 2 + 4.0 * nf * (F + N_MEL) + 4.0 * nf * F + nf * (4 * F * 5 + 16 * F * 4) + 8.0 * L.sum() * 5
    w.dominant = "gl2_kernel<2048,3> (4 per chunk) + stft_feature3_kernel (2 per chunk)"
    w.launches_dominant_per_step = 1
    w.note = (f"{len(mine)} of {n_utt} utterances on this rank ({nf} frames, {w.units / 3600:.2f} h), ragged chunks of {chunk}; "
              "features + ln-magnitude + Griffin-Lim (4 it, m 0.7, device-drawn initial phase); " +
              ("features and wavs copied to pinned host memory chunk by chunk (copy stream, double-buffered)" if d2h
               else "outputs stay in HBM"))
    w.flops = nf * (2 * 66e3 + 134e3 * 4.5)
    w.load = float(L.sum()) / SR

    def check():
        """first utterance of the first ragged chunk: features against the oracle."""
        from oracle import spectral_oracle as O
        ok = bool(torch.isfinite(step(0)).all().item())
        torch.cuda.synchronize()
        batch = chunks[0][0]
        l0, t0 = int(batch.lens[0]), int(batch.frames[0])
        y0 = batch.x[:l0].cpu().numpy()
        cur = torch.cuda.current_stream()
        sb._lib.check(lib.sb200_stft_features(plan.handle, sb.core.ptr(batch.x), C.byref(batch.c), float(ta.hp.preemphasis),
                                              sc_db, sc_db, sb.core.ptr(mags[0]), sb.core.ptr(mels[0]), None, sb.core.stream_ptr()))
        cur.synchronize()
        So, Mo = O.tt_get_specs(y0)
        e1 = np.linalg.norm(mags[0][:t0].t().cpu().numpy() - So) / np.linalg.norm(So)
        e2 = np.linalg.norm(mels[0][:t0].t().cpu().numpy() - Mo) / np.linalg.norm(Mo)
        return bool(ok and e1 < 1e-4 and e2 < 1e-4), f"utterance 0 of the ragged chunk vs oracle: mag {e1:.1e} mel {e2:.1e}; wavs finite"
    w.check = check
    return w


# ----------------------------------------------------------------------------- timing ---------------

def time_steps(torch, fn, steps, warmup, barrier):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    wall = time.perf_counter() - t0
    return e0.elapsed_time(e1) / 1e3, wall


def link_probe(torch, h2d_bytes, d2h_bytes, barrier, allmax, reps=6):
    """The host-link ceiling of the end-to-end step, measured with NO kernels: the step's own H2D and D2H byte counts moved
    between pinned host memory and HBM on two streams at once, every rank at the same time (max over ranks).  Returns
    the time such a step needs on the link alone plus the one-direction bandwidths."""
    dev_in = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dev_out = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    host_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    host_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(do_in, do_out):
        torch.cuda.synchronize()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if do_in:
                with torch.cuda.stream(s1):
                    dev_in.copy_(host_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    host_out.copy_(dev_out, non_blocking=True)
        torch.cuda.synchronize()
        return allmax(time.perf_counter() - t0) / reps
    run(True, True)
    t_in, t_out, t_both = run(True, False), run(False, True), run(True, True)
    return {"h2d_gbs": h2d_bytes / t_in / 1e9, "d2h_gbs": d2h_bytes / t_out / 1e9,
            "both_ms": 1e3 * t_both, "link_gbs": (h2d_bytes + d2h_bytes) / t_both / 1e9}


def fp32_peak_tflops(sm_mhz):
    """Non-tensor FP32 peak: 148 SMs x 128 FMA lanes x 2 FLOP at the maximum SM clock."""
    return 148 * 128 * 2 * (sm_mhz or 1965.0) * 1e6 / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="stft_mel", choices=["stft_mel", "griffinlim", "griffinlim_tt",
                                                               "griffinlim_batch", "mstft", "mstft_specs", "corpus", "corpus_d2h"])
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="profiling runs: skip e2e, extra workloads and the CPU baseline")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    cores = len(os.sched_getaffinity(0))
    cpu_kind = {"stft_mel": "stft_mel", "griffinlim": "griffinlim", "griffinlim_tt": "griffinlim",
                "griffinlim_batch": "griffinlim", "mstft": "mstft", "mstft_specs": "mstft", "corpus": "griffinlim",
                "corpus_d2h": "griffinlim"}[a.workload]
    workload_name = {"stft_mel": "stft_mel_64x5s", "griffinlim": "griffinlim_rtg_1x5s_4it",
                     "griffinlim_tt": "griffinlim_tt_1x5s_30it", "griffinlim_batch": "griffinlim_rtg_64x5s_4it",
                     "mstft": "mstft_fwd_bwd_16x22050_lossonly", "mstft_specs": "mstft_fwd_bwd_16x22050_specs",
                     "corpus": "corpus_10000utt_specs+griffinlim", "corpus_d2h": "corpus_10000utt_specs+griffinlim"}[a.workload]

    config = {"workload": workload_name, "per_gpu": workload_name, "sample_rate": SR, "n_fft": N_FFT, "hop": HOP, "win": WIN,
              "n_mel": N_MEL, "sharding": "utterances per rank, no data-path collective"}

    if a.impl == "reference":
        # The reference's own CPU implementation of the path (oracle port: librosa / TF are not installable) on
        # all host cores, parallelised as the reference does (process pool over utterances).  A step is a bounded sample of
        # the workload (the workload's own 64 utterances for stft_mel, rounded up to a whole number per core); at most 20
        # steps so the run ends within minutes.
        if rank != 0:
            return
        n_utt = {"stft_mel": cores * ((64 + cores - 1) // cores), "griffinlim": max(2 * cores, 16), "mstft": 64}[cpu_kind]
        cpu_reference_rate(cpu_kind, max(cores, 16), L5, cores)          # warm-up step (imports, page-in)
        vals, t_tot = [], 0.0
        for _ in range(max(1, min(a.steps, 20))):
            v, dt = cpu_reference_rate(cpu_kind, n_utt, L5, cores)
            vals.append(v)
            t_tot += dt
            if t_tot > 120.0:
                break
        v = float(np.median(vals))
        sample = (f"{n_utt} x 5 s utterances per step through the oracle restatement of the reference path, "
                  f"{cores} worker processes, {len(vals)} steps, {t_tot:.1f} s")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": len(vals),
            "warmup": 1, "ms_per_step": 1e3 * n_utt * L5 / SR / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if cpu_kind != "mstft" else "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch
    import transtacos_retunegan_b200 as sb
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    numa_bound = sb.sharding.bind_to_gpu_numa(local) if world > 1 else False   # before any pinned allocation
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sync_t = torch.zeros(1, device="cuda")

    def barrier():
        if dist is not None:
            dist.all_reduce(sync_t)

    def allred(x, op=None):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
        return float(t.item())

    def allsum(x):
        return allred(x)

    def allmax(x):
        return allred(x, dist.ReduceOp.MAX if dist is not None else None)

    def allmin(x):
        return allred(x, dist.ReduceOp.MIN if dist is not None else None)

    makers = {"stft_mel": lambda: make_stft_mel(sb, torch), "griffinlim": lambda: make_griffinlim(sb, torch, 1, "rtg"),
              "griffinlim_tt": lambda: make_griffinlim(sb, torch, 1, "tt"),
              "griffinlim_batch": lambda: make_griffinlim(sb, torch, 64, "rtg"),
              "mstft": lambda: make_mstft(sb, torch, ddp=world > 1),
              "mstft_specs": lambda: make_mstft(sb, torch, specs=True, ddp=world > 1),
              "corpus": lambda: make_corpus(sb, torch, rank, world),
              "corpus_d2h": lambda: make_corpus(sb, torch, rank, world, d2h=True)}
    sampler = ClockSampler(local)
    sampler.start()
    peak, peak_src = peaks()
    fp32_peak = fp32_peak_tflops(sampler.sm_max)

    def checked(mk):
        """Build a workload and compare one row of its output with the oracle (rank 0; never timed).  Agreed over the ranks:
        a step may contain a collective, so either every rank runs it or none does."""
        w, ok, msg = None, True, ""
        try:
            w = mk()
            if rank == 0 and not os.environ.get("SB200_BENCH_NO_CHECK"):   # (ablation builds compute garbage on purpose)
                ok, msg = w.check()
        except Exception as ex:
            ok, msg = False, repr(ex)[:300]
        if allmin(1.0 if ok else 0.0) < 1.0:
            raise RuntimeError(f"workload check failed on some rank (rank {rank}: {msg or 'ok'})")
        return w, msg

    def measure(w, steps, warmup, min_seconds=0.0):
        """Timed region of one workload: device time between barriers, max over ranks; clocks sampled over the same loop."""
        sampler.samples, sampler.reasons = [], set()
        sampler.active = True
        l0 = sb._lib.launch_count()
        dev_s, wall_s = time_steps(torch, w.step, steps, warmup, barrier)
        launches = sb._lib.launch_count() - l0
        window = "warmup+timed"
        # short timed region: keep the same loop running to catch the clocks under load.  The decision and the number of
        # extra steps are agreed over the ranks (a step may contain a collective: a rank-local loop would deadlock).
        dev_s = allmax(dev_s)
        if allmax(1.0 if len(sampler.samples) < 5 else 0.0) > 0 or dev_s < min_seconds:
            n_cont = int(min(20000, max(1, max(min_seconds, 0.5) / max(dev_s / steps, 1e-6))))
            for i in range(n_cont):
                w.step(i)
                if i % 64 == 63:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            window = "warmup+timed+continuation of the same loop"
        sampler.active = False
        return dev_s, wall_s, launches * steps // (steps + warmup), sampler.summary(window)

    w, check_msg = checked(makers[a.workload])
    dev_s, wall_s, timed_launches, clocks = measure(w, a.steps, a.warmup, 1.0)
    total_units = allsum(w.units)          # audio seconds per step over all ranks (corpus shards differ slightly)

    # dominant kernel duration, live: for single-kernel steps it is the step; otherwise the per-launch duration is the device
    # time of the step divided by its dominant launches (upper bound).
    kern_s = dev_s / a.steps / w.launches_dominant_per_step
    achieved = w.alg_bytes / kern_s / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(a.workload)
            traffic_src = "static: " + tj.get("_source", "profiles/traffic.json (ncu --set full capture, not measured in this run)")
        except Exception:
            traffic = None

    # end to end through the public API with host buffers; the link ceiling of the same byte counts is probed beside it
    e2e = None
    if w.e2e is not None and not a.kernel_only:
        K2 = max(3, min(a.steps, 20))
        for i in range(3):
            w.e2e(i)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(K2):
            w.e2e(i)
        torch.cuda.synchronize()
        e2e_s = allmax(time.perf_counter() - t0)
        e2e = {"value": world * w.units * K2 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h,
               "steps": K2, "ms_per_step": 1e3 * e2e_s / K2,
               "api": "transtacos_audio.get_specs(cpu_tensor[64,L], out=pinned)" if a.workload == "stft_mel"
               else "retunegan_audio.inv_mag(numpy)"}
        try:
            lp = link_probe(torch, int(w.h2d), int(w.d2h), barrier, allmax)
            e2e.update({"link_gbs": lp["link_gbs"], "link_ms_per_step": lp["both_ms"],
                        "frac_of_link": lp["both_ms"] / (1e3 * e2e_s / K2), "link_h2d_gbs": lp["h2d_gbs"],
                        "link_d2h_gbs": lp["d2h_gbs"],
                        "link_note": "the step's own H2D + D2H bytes between pinned host memory and HBM on two streams, no kernels, "
                                     "all ranks at once, max over ranks; frac_of_link = that time / the end-to-end step time"})
        except Exception as ex:
            e2e["link_error"] = repr(ex)[:200]

    def ddp_probe():
        # under DDP the generator's gradients (RefineGAN_small, ~2.75 M parameters, retunegan/hparam.py:50)
        # are all-reduced by DDP's buckets many layers after this loss; timed here on its own, next to the loss step
        gbuf = torch.zeros(2_750_000, device="cuda")
        for _ in range(5):
            dist.all_reduce(gbuf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(50):
            dist.all_reduce(gbuf)
        e1.record()
        torch.cuda.synchronize()
        return {"generator_grad_allreduce_ms": allmax(e0.elapsed_time(e1) / 50), "generator_grad_bytes": gbuf.numel() * 4,
                "loss_allreduce": "one fp32 scalar per step, inside the timed step (multi_stft_loss(ddp_reduce=True)): " +
                                  sb.loss.ddp_reduce_path()}

    ddp_info = ddp_probe() if (world > 1 and a.workload.startswith("mstft")) else None

    extra = {}
    if not a.no_extra and not a.kernel_only and a.workload == "stft_mel":
        ddp_cached = None
        for key, mkey, kmin in (("griffinlim_rtg_1x5s_4it", "griffinlim", 50),
                                ("griffinlim_tt_1x5s_30it", "griffinlim_tt", 10),
                                ("griffinlim_rtg_64x5s_4it", "griffinlim_batch", 5),
                                ("mstft_fwd_bwd_16x22050_lossonly", "mstft", 50),
                                ("mstft_fwd_bwd_16x22050_specs", "mstft_specs", 20),
                                ("corpus_10000utt_specs+griffinlim", "corpus", 2),
                                ("corpus_10000utt_specs+griffinlim_d2h", "corpus_d2h", 2)):
            try:
                ww, msg = checked(makers[mkey])
            except Exception as ex:   # secondary numbers must never break the contract line
                extra[key] = {"error": repr(ex)[:300]}
                continue
            d1, _ = time_steps(torch, ww.step, 1, 3, barrier)          # calibration: steps for >= 0.3 s of device time
            k = int(min(2000, max(kmin, 0.3 / max(allmax(d1), 1e-6))))
            d, _, nl, ck = measure(ww, k, 3)
            units = allsum(ww.units)
            strong = mkey.startswith("corpus")
            ent = {"value": units * k / d, "unit": UNIT, "ms_per_step": 1e3 * d / k, "steps": k, "n_gpus": world,
                   "scaling": "strong" if strong else "weak", "gpu_launches_per_step": nl // k,
                   "roofline_frac_hbm": ww.alg_bytes * ww.launches_dominant_per_step / (d / k) / 1e9 / peak,
                   "roofline_frac_fp32": ww.flops / (d / k) / 1e12 / fp32_peak,
                   "clocks": ck, "parity_check": msg, "note": ww.note}
            if strong:
                ent["load_audio_s_max"], ent["load_audio_s_min"] = allmax(ww.load), allmin(ww.load)
            if mkey.startswith("mstft") and world > 1:
                ddp_cached = ddp_cached or ddp_probe()
                ent["ddp"] = ddp_cached
            if mkey.startswith("mstft") and world == 1:
                # the eager step is bound by the host (Python, autograd engine, driver calls); the same step replayed as ONE CUDA
                # graph is its GPU time
                try:
                    ent["graph_replay_ms_per_step"] = graph_replay_ms(torch, ww.step)
                    ent["graph_replay_note"] = ("whole step (forward, backward, the autograd wrapper's ops) captured once with "
                                                "torch.cuda.graph and replayed 100 times: GPU time of the step; ms_per_step is the eager API")
                except Exception as ex:
                    ent["graph_replay_error"] = repr(ex)[:200]
            extra[key] = ent
            del ww
            torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            try:   # the reference's own torch formulation on the same GPU (cuFFT + ~60 small kernels): comparator only
                yt = (0.1 * torch.randn(16, 22050, device="cuda")).clamp_(-0.999, 0.999)
                ygt = torch.tanh(yt + 0.01 * torch.randn_like(yt)).requires_grad_(True)

                def torch_step(i):
                    ygt.grad = None
                    _torch_ref_mstft(yt, ygt).backward()
                    return ygt.grad
                d, _ = time_steps(torch, torch_step, 30, 5, lambda: None)
                extra["mstft_fwd_bwd_16x22050_lossonly_torch_cufft"] = {
                    "value": 16 * 22050 / SR * 30 / d, "unit": UNIT, "ms_per_step": 1e3 * d / 30,
                    "note": "comparator: the reference's multi_stft_loss graph through torch.stft (cuFFT) + autograd on the same GPU; "
                            "not this repo's code path"}
            except Exception as ex:
                extra["mstft_fwd_bwd_16x22050_lossonly_torch_cufft"] = {"error": repr(ex)[:200]}
            try:   # get_specs (transtacos/audio.py:73-77) written with torch ops on the same GPU: cuFFT + dense mel GEMM comparator
                from oracle import spectral_oracle as O
                yb = (0.1 * torch.randn(64, L5, device="cuda")).clamp_(-0.999, 0.999)
                wnd = torch.hann_window(WIN, device="cuda")
                mbt = torch.from_numpy(O.mel_basis(N_FFT)).cuda()

                def torch_specs(i):
                    x = torch.cat([yb[:, :1], yb[:, 1:] - 0.97 * yb[:, :-1]], dim=1)
                    D = torch.stft(x, N_FFT, HOP, WIN, window=wnd, center=True, pad_mode="reflect", return_complex=True).abs()
                    S = 8.0 * ((20.0 * torch.log10(D.clamp_min(1e-5)) - 20.0 + 100.0) / 100.0) - 4.0
                    M = 8.0 * ((20.0 * torch.log10((mbt @ D).clamp_min(1e-5)) - 20.0 + 100.0) / 100.0) - 4.0
                    return S, M
                d, _ = time_steps(torch, torch_specs, 30, 5, lambda: None)
                extra["stft_mel_64x5s_torch_cufft"] = {
                    "value": 64 * L5 / SR * 30 / d, "unit": UNIT, "ms_per_step": 1e3 * d / 30,
                    "note": "comparator: get_specs written with torch ops (torch.stft / cuFFT, dense mel GEMM, elementwise) on the same GPU; "
                            "not this repo's code path"}
                del yb
            except Exception as ex:
                extra["stft_mel_64x5s_torch_cufft"] = {"error": repr(ex)[:200]}
    sampler.stop_flag = True

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.kernel_only:
        # bounded sample: 64 utterances per core (10-30 CPU-seconds of the reference path in all)
        n_utt = {"stft_mel": max(64 * cores, 64), "griffinlim": max(4 * cores, 16), "mstft": 64}[cpu_kind]
        v, dt = cpu_reference_rate(cpu_kind, n_utt, L5, cores)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n_utt} x 5 s utterances, oracle restatement of the reference path, {cores} processes, {dt:.1f} s"}

    if rank == 0:
        strong = a.workload.startswith("corpus")
        out = {
            "metric": METRIC, "value": total_units * a.steps / dev_s, "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dev_s / a.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, workload=w.name, per_gpu=w.name, l2=w.note, parity_check=check_msg,
                           **({"ddp": ddp_info} if ddp_info else {}),
                           **({"numa_bound": bool(numa_bound)} if world > 1 else {})),
            "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(timed_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": w.dominant, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": w.alg_bytes, "kernel_us": kern_s * 1e6,
                         "fp32": {"achieved_tflops": w.flops / (dev_s / a.steps) / 1e12, "peak_tflops": fp32_peak,
                                  "frac": w.flops / (dev_s / a.steps) / 1e12 / fp32_peak,
                                  "note": "algorithmic FP32 FLOPs of the step (SURVEY.md 8d) / step time / (148 SMs x 128 lanes x 2 x max SM "
                                          "clock); the measured FMA-pipe utilisation is in profiles/"}},
            "cpu_baseline": cpu_baseline, "wall_s": wall_s,
        }
        if extra:
            out["extra"] = extra
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
