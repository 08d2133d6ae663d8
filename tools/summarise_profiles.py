#!/usr/bin/env python3
"""Turn the CSV exports of tools/collect_profiles.sh (gpurun_out/r02/evidence/) into the committed evidence under profiles/:

    profiles/r02_ncu_<capture>.md      key metrics per captured launch, stall breakdown, dynamic instruction mix per opcode
    profiles/r02_launches_<workload>.csv   ncu launch lists of the bench commands (sb200 kernels only + a count of the rest)
    profiles/traffic.json              dram bytes per launch of the dominant kernel of every workload, FROM the ncu capture
    profiles/r02_sass_mix.txt          static SASS opcode histograms of the hot kernels (cuobjdump of the in-tree library)

usage: python tools/summarise_profiles.py [evidence_dir]
"""
import collections
import csv
import gzip
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02", "evidence")
OUT = os.path.join(ROOT, "profiles")

KEYS = [("time_us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("regs", "launch__registers_per_thread"), ("dyn_smem_KB", "launch__shared_mem_per_block_dynamic"),
        ("issue_active_%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("fma_pipe_%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("lsu_pipe_%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("xu_pipe_%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("alu_pipe_%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("smem_wavefronts_%", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("warp_insts", "smsp__inst_executed.sum"), ("warps_per_scheduler", "smsp__warps_active.avg.per_cycle_active"),
        ("dram_read_B", "dram__bytes_read.sum"), ("dram_write_B", "dram__bytes_write.sum"),
        ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")]
STALLS = ["long_scoreboard", "wait", "short_scoreboard", "not_selected", "mio_throttle", "math_pipe_throttle", "no_instruction",
          "dispatch_stall", "barrier", "branch_resolving", "lg_throttle", "drain", "membar", "sleeping"]


def to_num(v, unit=""):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    u = unit.lower()
    for pre, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("usecond", 1.0), ("msecond", 1e3), ("nsecond", 1e-3)):
        if u.startswith(pre):
            return x * m
    return x


def raw_rows(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for k, u, v in zip(h, units, r):
            d[k] = to_num(v, u)
        d["_name"] = r[h.index("Kernel Name")]
        out.append(d)
    return out


def source_summary(path):
    data = list(csv.reader(io.TextIOWrapper(gzip.open(path))))
    # one table per launch in the capture, separated by header rows; keep them apart
    tables, cur = [], None
    for r in data:
        if r and r[0] == "Address":
            cur = {"hdr": r, "rows": []}
            tables.append(cur)
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    outs = []
    for t in tables:
        ix = {h: i for i, h in enumerate(t["hdr"])}
        ops, stalls, wf = collections.Counter(), collections.Counter(), 0
        for r in t["rows"]:
            toks = r[ix["Source"]].split()
            if not toks:
                continue
            op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
            ops[op] += int(r[ix["Instructions Executed"]] or 0)
            wf += int(r[ix["L1 Wavefronts Shared"]] or 0)
            for h, i in ix.items():
                if h.startswith("stall_"):
                    stalls[h[6:]] += int(r[i] or 0)
        outs.append((len(t["rows"]), ops, stalls, wf))
    return outs


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("sb200::", "")


def main():
    os.makedirs(OUT, exist_ok=True)
    traffic = {"_source": "ncu --set full --clock-control none captures of round 2 (tools/collect_profiles.sh, summarised by "
                          "tools/summarise_profiles.py): dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel"}
    dominant = {"feat3": ("stft_mel", "stft_feature3"), "gl2_batch": ("griffinlim_batch", "gl2_kernel"),
                "mstft_fused": ("mstft", "mstft_multi_bwd"), "mstft_specs": ("mstft_specs", "mstft_multi")}
    for cap, (workload, pat) in dominant.items():
        raw = os.path.join(EV, f"ncu_{cap}_raw.csv")
        if not os.path.exists(raw):
            continue
        rows = raw_rows(raw)
        src = os.path.join(EV, f"ncu_{cap}_source.csv.gz")
        srcs = source_summary(src) if os.path.exists(src) else []
        lines = [f"# Round 2 -- ncu capture `{cap}` (bench.py --workload {workload} --steps 3 --warmup 3 --kernel-only; B200, "
                 "`--set full --clock-control none --import-source on`)", "",
                 "Generated by tools/summarise_profiles.py from the CSV exports of the report (the .ncu-rep itself is not committed).", ""]
        lines.append("| launch | " + " | ".join(k for k, _ in KEYS) + " |")
        lines.append("|---|" + "---|" * len(KEYS))
        for i, d in enumerate(rows):
            vals = []
            for k, m in KEYS:
                v = d.get(m, "-")
                if isinstance(v, float):
                    v = f"{v / 1024:.1f}" if k == "dyn_smem_KB" else (f"{v:.0f}" if abs(v) >= 1000 else f"{v:.2f}")
                vals.append(str(v))
            lines.append(f"| {i}: `{short(d['_name'])[:60]}` | " + " | ".join(vals) + " |")
        lines += ["", "Warps stalled per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):", ""]
        lines.append("| launch | " + " | ".join(STALLS) + " |")
        lines.append("|---|" + "---|" * len(STALLS))
        for i, d in enumerate(rows):
            vals = [d.get(f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio", "-") for s in STALLS]
            lines.append(f"| {i} | " + " | ".join(f"{v:.2f}" if isinstance(v, float) else str(v) for v in vals) + " |")
        for i, (n, ops, stalls, wf) in enumerate(srcs):
            tot = sum(ops.values())
            lines += ["", f"Launch {i}: {n} SASS instructions, {tot} warp instructions executed, {wf} shared-memory wavefronts "
                          "(source page).  Dynamic instruction mix:", "",
                      "  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(28)), "",
                      "  stall samples: " + "  ".join(f"{k}:{v}" for k, v in stalls.most_common(10))]
        open(os.path.join(OUT, f"r02_ncu_{cap}.md"), "w").write("\n".join(lines) + "\n")
        dom = [d for d in rows if pat in d["_name"]]
        if dom:
            # the dominant kernel of the workload = the longest matching launch
            d = max(dom, key=lambda x: x.get("gpu__time_duration.sum", 0))
            traffic[workload] = int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
            traffic[f"_{workload}_kernel"] = short(d["_name"])
    # launch lists: keep our kernels, count the rest
    for f in sorted(os.listdir(EV)):
        if not f.startswith("launches_"):
            continue
        rows = [r for r in csv.reader(open(os.path.join(EV, f))) if len(r) > 10]
        if not rows:
            continue
        h = rows[0]
        kn, mv = h.index("Kernel Name"), h.index("Metric Value")
        other = sum(1 for r in rows[1:] if "sb200" not in r[kn])
        keep = [r for r in rows[1:] if "sb200" in r[kn]]
        with open(os.path.join(OUT, "r02_" + f), "w", newline="") as fo:
            w = csv.writer(fo)
            w.writerow(["# ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --workload "
                        f"{f[9:-4]} --steps 3 --warmup 3 --kernel-only --no-extra`; {other} torch kernels that build the synthetic "
                        "input (outside the timed region) omitted"])
            w.writerow(["id", "kernel", "stream", "block", "grid", "gpu__time_duration.sum (ns)"])
            for r in keep:
                w.writerow([r[0], short(r[kn]), r[h.index("Stream")], r[h.index("Block Size")], r[h.index("Grid Size")], r[mv]])
    if len(traffic) > 1:
        json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    # static SASS histograms
    lib = os.path.join(ROOT, "transtacos-retunegan_b200", "libspectral_b200.so")
    if os.path.exists(lib):
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
        cur, funcs = None, collections.OrderedDict()
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                funcs[cur] = collections.Counter()
            elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
                toks = line.split()
                op = toks[2] if toks[1].startswith("@") else toks[1]
                funcs[cur][op.split(".")[0].rstrip(";")] += 1
        demangle = subprocess.run(["cu++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()
        with open(os.path.join(OUT, "r02_sass_mix.txt"), "w") as fo:
            fo.write("Static SASS opcode histograms (cuobjdump -sass libspectral_b200.so, sm_100a), hot kernels.  FFMA2 / FADD2 / FMUL2 = packed\n"
                     "fp32 (Blackwell), SYNCS = mbarrier, UTMALDG = bulk-tensor (TMA) load, USETMAXREG = setmaxnreg.\n\n")
            for (name, ops), dn in zip(funcs.items(), demangle):
                if not any(k in name for k in ("stft_feature3", "stft_feature2_kernelILi1024", "gl2_kernelILi2048ELi3", "gl2_persistent",
                                               "mstft_multi", "mstft_bwd_kernelILi2048", "mstft_fwd_kernelILi2048", "mstft_all", "stft_smp_kernelILi2048", "yin", "grad_ola")):
                    continue
                fo.write(f"{short(dn)}  [{sum(ops.values())} instructions]\n  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(36)) + "\n\n")
    print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("r02_") or f == "traffic.json"))


if __name__ == "__main__":
    main()
