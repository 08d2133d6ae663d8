# A/B of Griffin-Lim variants (tools/variants.py --tu tu_gl.cu name="-D..."): ms per step of the three Griffin-Lim workloads
for v in "" $@; do for w in griffinlim_batch griffinlim griffinlim_tt; do
  r=$(SB200_LIB=${v:+$PWD/scratch/var_$v.so} python bench.py --no-extra --kernel-only --workload $w --steps 100 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), d['config']['parity_check'][:60])")
  echo "lib=${v:-main} $w ms/step: $r"
done; done
