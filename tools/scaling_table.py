#!/usr/bin/env python3
"""Rewrite the multi-GPU table of DESIGN.md section 5 from profiles/r02_bench_{1,2,4,8}gpu.json (the bench lines of the round)."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = {n: json.loads(open(os.path.join(ROOT, "profiles", f"r02_bench_{n}gpu.json")).read().strip().splitlines()[-1]) for n in (1, 2, 4, 8)}
g = lambda n, k: L[n]["extra"][k]
fmt = lambda v: f"{v / 1e6:.2f} M" if v >= 1e6 else f"{v / 1e3:.1f} k"
rows = [("STFT+mel 64 × 5 s per GPU (weak), device resident", [L[n]["value"] for n in (1, 2, 4, 8)]),
        ("same, end to end from / to pinned host memory", [L[n]["e2e"]["value"] for n in (1, 2, 4, 8)])]
for k, name in (("griffinlim_rtg_64x5s_4it", "Griffin-Lim 64 × 5 s × 4 it per GPU (weak)"),
                ("griffinlim_tt_1x5s_30it", "Griffin-Lim 1 × 5 s × 30 it per GPU (weak)"),
                ("mstft_fwd_bwd_16x22050_lossonly", "mstft loss-only, 16 × 1 s per GPU (weak; at N > 1 the loss is averaged over the ranks inside the step, in-kernel over NVLink peer memory)"),
                ("mstft_fwd_bwd_16x22050_specs", "mstft training variant (weak)"),
                ("corpus_10000utt_specs+griffinlim", "corpus 10 000 utterances (STRONG), resident"),
                ("corpus_10000utt_specs+griffinlim_d2h", "corpus (STRONG), features + wavs copied to pinned host memory")):
    rows.append((name, [g(n, k)["value"] for n in (1, 2, 4, 8)]))
t = "| workload (spectrogram-seconds / s, whole job) | 1 GPU | 2 GPUs | 4 GPUs | 8 GPUs | 8-GPU efficiency |\n|---|---|---|---|---|---|\n"
for name, v in rows:
    t += f"| {name} | {fmt(v[0])} | {fmt(v[1])} | {fmt(v[2])} | {fmt(v[3])} | {v[3] / v[0] / 8:.2f} |\n"
e = lambda n: L[n]["e2e"]
t += ("| host-link ceiling of the end-to-end step (`e2e.link_*`): D2H GB/s per GPU with all ranks copying; link time per step; `frac_of_link` | "
      + " | ".join(f"{e(n)['link_d2h_gbs']:.1f}; {e(n)['link_ms_per_step']:.2f} ms; {e(n)['frac_of_link']:.2f}" for n in (1, 2, 4, 8)) + " | |\n")
lo, sp = "mstft_fwd_bwd_16x22050_lossonly", "mstft_fwd_bwd_16x22050_specs"
t += ("| mstft step time loss-only / training; generator-gradient all-reduce (11 MB) beside it | "
      + f"{g(1, lo)['ms_per_step']:.3f} / {g(1, sp)['ms_per_step']:.3f} ms | "
      + " | ".join(f"{g(n, lo)['ms_per_step']:.3f} / {g(n, sp)['ms_per_step']:.3f} ms; {g(n, lo)['ddp']['generator_grad_allreduce_ms'] * 1e3:.0f} µs" for n in (2, 4, 8)) + " | |\n")
c = "corpus_10000utt_specs+griffinlim"
t += ("| corpus shard load, audio seconds max / min over ranks | | "
      + " | ".join(f"{g(n, c)['load_audio_s_max']:.2f} / {g(n, c)['load_audio_s_min']:.2f}" for n in (2, 4, 8)) + " | |\n")
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
s = re.sub(r"(<!-- scaling table: tools/scaling_table.py -->\n).*?(<!-- end scaling table -->)", lambda m: m.group(1) + t + m.group(2), s, flags=re.S)
open(p, "w").write(s)
print(t)
