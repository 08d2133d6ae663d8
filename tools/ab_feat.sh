for v in "" scratch/var_nohint.so; do for tma in 0 1; do
  r=$(SB200_LIB=${v:+$PWD/$v} SB200_FEAT_TMA=$tma python bench.py --no-extra --kernel-only --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2))")
  echo "lib=${v:-main} tma=$tma us=$r"
done; done
