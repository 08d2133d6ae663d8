#!/usr/bin/env python3
"""Instruction mix of one kernel in libspectral_b200.so:  python tools/sass_mix.py <substring of mangled name> [--dump]"""
import collections, re, subprocess, sys, os
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "transtacos-retunegan_b200", "libspectral_b200.so")
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
for name, lines in funcs.items():
    if pat not in name: continue
    ops = collections.Counter()
    for l in lines:
        toks = l.split()
        op = toks[1]
        if op.startswith("@"): op = toks[2]
        ops[op.split(".")[0].rstrip(";")] += 1
    print(name, len(lines))
    print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(40)))
    if "--dump" in sys.argv:
        open("/tmp/t/dump.sass", "w").write("\n".join(lines))
