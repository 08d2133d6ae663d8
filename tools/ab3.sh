for v in "" scratch/var_baronly.so; do for i in 1 2; do
  r=$(SB200_LIB=${v:+$PWD/$v} python bench.py --no-extra --kernel-only --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), d['config']['parity_check'][:30])")
  echo "lib=${v:-main} us=$r"
done; done
