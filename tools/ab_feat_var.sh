# A/B of feature-kernel variants (tools/variants.py name="-D..."): us per 64 x 5 s launch (config 3), two runs each
for v in "" $@; do for i in 1 2; do
  r=$(SB200_LIB=${v:+$PWD/scratch/var_$v.so} python bench.py --no-extra --kernel-only --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), d['config']['parity_check'][:40])")
  echo "lib=${v:-main} us=$r"
done; done
