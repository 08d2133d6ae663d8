#!/bin/bash
# usage: tools/run_variants.sh name1 name2 ...   (on the GPU box; "main" = the in-tree library)
for v in "$@"; do
  lib=$PWD/scratch/var_$v.so; [ "$v" = main ] && lib=
  SB200_LIB=$lib python bench.py --kernel-only --steps 200 --warmup 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'kernel_us', round(d['roofline']['kernel_us'], 2), 'frac', round(d['roofline']['frac'], 4), 'sm_mhz', (d.get('clocks') or {}).get('sm_mhz'))"
done
