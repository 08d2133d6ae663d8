#!/usr/bin/env python3
"""bench.py -- spectrogram-seconds per second of the spectral hot path on N B200s (one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stft_mel|griffinlim|griffinlim_batch|mstft]
                    [--impl reference] [--no-extra]

Default workload = BASELINE.json configs[2]: batched STFT + mel feature extraction (TransTacoS get_specs:
pre-emphasis, dB-normalise), 64 synthetic 5 s utterances per GPU, hparam.py shapes.  A "step" is one pass of the
hot path over one such batch.  `value` is timed with the batch resident in HBM (CUDA events on the launch
stream, max over ranks); `e2e` goes through the numpy/CPU-tensor-facing public API with pinned HOST buffers,
host<->device copies inside the timed region.  Utterances shard across ranks with no data-path collective (weak
scaling: every rank owns 64 utterances).  `--impl reference` times the CPU oracle restatement of the reference's
librosa/numpy path (the reference itself cannot be installed: librosa / TF are absent and there is no network) on
all host cores, the way the reference parallelises it (process pool over utterances, databaker.py:31).
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
            time.sleep(0.01)

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
    w.check = lambda: (torch.isfinite(mags[0]).all().item() and torch.isfinite(mels[0]).all().item())
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
    w.check = lambda: torch.isfinite(step(0)).all().item()
    return w


def make_mstft(sb, torch, B=16, T=22050, specs=False, rot=4, ddp=False):
    w = Workload()
    w.name = f"mstft_fwd_bwd_{B}x{T}" + ("_specs" if specs else "_lossonly")
    g = torch.Generator(device="cuda").manual_seed(77 + int(os.environ.get("RANK", 0)))
    ys = [(0.1 * torch.randn(B, 1, T, device="cuda", generator=g)).clamp_(-0.999, 0.999) for _ in range(rot)]
    ygs = [torch.tanh(y + 0.01 * torch.randn(B, 1, T, device="cuda", generator=g)).requires_grad_(True) for y in ys]
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
        ("; the reported loss is averaged over the ranks with one NCCL all-reduce of a scalar inside the step" if ddp else "")
    w.check = lambda: torch.isfinite(step(0)).all().item()
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
    w.check = lambda: torch.isfinite(step(0)).all().item()
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="stft_mel", choices=["stft_mel", "griffinlim", "griffinlim_tt",
                                                               "griffinlim_batch", "mstft", "mstft_specs", "corpus"])
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="profiling runs: skip e2e, extra workloads and the CPU baseline")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    cores = len(os.sched_getaffinity(0))
    cpu_kind = {"stft_mel": "stft_mel", "griffinlim": "griffinlim", "griffinlim_tt": "griffinlim",
                "griffinlim_batch": "griffinlim", "mstft": "mstft", "mstft_specs": "mstft", "corpus": "griffinlim"}[a.workload]
    workload_name = {"stft_mel": "stft_mel_64x5s", "griffinlim": "griffinlim_rtg_1x5s_4it",
                     "griffinlim_tt": "griffinlim_tt_1x5s_30it", "griffinlim_batch": "griffinlim_rtg_64x5s_4it",
                     "mstft": "mstft_fwd_bwd_16x22050_lossonly", "mstft_specs": "mstft_fwd_bwd_16x22050_specs",
                     "corpus": "corpus_10000utt_specs+griffinlim"}[a.workload]

    config = {"workload": workload_name, "per_gpu": workload_name, "sample_rate": SR, "n_fft": N_FFT, "hop": HOP, "win": WIN,
              "n_mel": N_MEL, "sharding": "utterances per rank, no data-path collective"}

    if a.impl == "reference":
        # The reference's own CPU implementation of the path (oracle port: librosa / TF are not installable) on
        # all host cores, parallelised as the reference does (process pool over utterances).  A step is a bounded sample of
        # the workload: 8 utterances per core (~0.3 s); at most 20 steps so the run ends within minutes.
        if rank != 0:
            return
        n_utt = {"stft_mel": max(8 * cores, 64), "griffinlim": max(2 * cores, 16), "mstft": 64}[cpu_kind]
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

    def allsum(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    makers = {"stft_mel": lambda: make_stft_mel(sb, torch), "griffinlim": lambda: make_griffinlim(sb, torch, 1, "rtg"),
              "griffinlim_tt": lambda: make_griffinlim(sb, torch, 1, "tt"),
              "griffinlim_batch": lambda: make_griffinlim(sb, torch, 64, "rtg"),
              "mstft": lambda: make_mstft(sb, torch, ddp=world > 1),
              "mstft_specs": lambda: make_mstft(sb, torch, specs=True, ddp=world > 1),
              "corpus": lambda: make_corpus(sb, torch, rank, world),
              "corpus_d2h": lambda: make_corpus(sb, torch, rank, world, d2h=True)}
    w = makers[a.workload]()
    assert w.check(), "workload produced non-finite output"

    sampler = ClockSampler(local)
    sampler.start()
    # warm-up happens inside time_steps; sample clocks from the start of warm-up to the end of the timed region
    sampler.active = True
    l0 = sb._lib.launch_count()
    dev_s, wall_s = time_steps(torch, w.step, a.steps, a.warmup, barrier)
    launches = sb._lib.launch_count() - l0
    window = "warmup+timed"
    # very short timed region: keep the same loop running ~1 s to catch the clocks under load.  The decision and the number of
    # extra steps are agreed over the ranks (a step may contain a collective: a rank-local, time-based loop would deadlock).
    if allmax(1.0 if len(sampler.samples) < 5 else 0.0) > 0:
        n_cont = int(min(20000, max(1, 1.0 / max(allmax(dev_s) / a.steps, 1e-6))))
        for i in range(n_cont):
            w.step(i)
            if i % 64 == 63:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        window = "warmup+timed+1s continuation of the same loop"
    sampler.active = False
    dev_s = allmax(dev_s)
    total_units = allsum(w.units)          # audio seconds per step over all ranks (corpus shards differ slightly)
    timed_launches = launches * a.steps // (a.steps + a.warmup)

    # dominant kernel duration, live: for single-kernel steps it is the step; otherwise re-time it alone is not possible
    # from Python, so the per-launch duration is the device time of the step divided by its dominant launches (upper bound).
    kern_s = dev_s / a.steps / w.launches_dominant_per_step
    peak, peak_src = peaks()
    achieved = w.alg_bytes / kern_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(a.workload)
        except Exception:
            traffic = None

    # end to end through the public API with host buffers
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
               "steps": K2, "api": "transtacos_audio.get_specs(cpu_tensor[64,L], out=pinned)" if a.workload == "stft_mel"
               else "retunegan_audio.inv_mag(numpy)"}

    ddp_info = None
    if world > 1 and a.workload.startswith("mstft"):
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
        ddp_info = {"generator_grad_allreduce_ms": allmax(e0.elapsed_time(e1) / 50), "generator_grad_bytes": gbuf.numel() * 4,
                    "loss_allreduce": "one fp32 scalar per step, inside the timed step (multi_stft_loss(ddp_reduce=True))"}
        del gbuf

    extra = {}
    if world == 1 and not a.no_extra and not a.kernel_only and a.workload == "stft_mel":
        for key, mk, k in (("griffinlim_rtg_1x5s_4it", makers["griffinlim"], 50),
                           ("griffinlim_tt_1x5s_30it", makers["griffinlim_tt"], 10),
                           ("griffinlim_rtg_64x5s_4it", makers["griffinlim_batch"], 5),
                           ("mstft_fwd_bwd_16x22050_lossonly", makers["mstft"], 50),
                           ("mstft_fwd_bwd_16x22050_specs", makers["mstft_specs"], 20),
                           ("corpus_10000utt_specs+griffinlim", makers["corpus"], 2),
                           ("corpus_10000utt_specs+griffinlim_d2h", makers["corpus_d2h"], 2)):
            try:
                ww = mk()
                d, _ = time_steps(torch, ww.step, k, 3, lambda: None)
                extra[key] = {"value": ww.units * k / d, "unit": UNIT, "ms_per_step": 1e3 * d / k,
                              "roofline_frac_hbm": ww.alg_bytes * ww.launches_dominant_per_step / (d / k) / 1e9 / peak,
                              "note": ww.note}
                del ww
            except Exception as ex:   # secondary numbers must never break the contract line
                extra[key] = {"error": repr(ex)[:200]}
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
        out = {
            "metric": METRIC, "value": total_units * a.steps / dev_s, "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dev_s / a.steps, "higher_is_better": True,
            "scaling": "strong" if a.workload == "corpus" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, workload=w.name, per_gpu=w.name, l2=w.note, **({"ddp": ddp_info} if ddp_info else {}),
                           **({"numa_bound": bool(numa_bound)} if world > 1 else {})),
            "clocks": sampler.summary(window),
            "e2e": e2e, "gpu_launches": int(timed_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": w.dominant, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": w.alg_bytes, "kernel_us": kern_s * 1e6},
            "cpu_baseline": cpu_baseline, "wall_s": wall_s,
        }
        if extra:
            out["extra"] = extra
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
