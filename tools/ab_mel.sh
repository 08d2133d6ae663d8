# A/B of the feat3 epilogue's mel: ELL filterbank (default) vs the segment formulation (-DSB200_MEL_SEGMENTS)
#   python tools/variants.py melseg="-DSB200_MEL_SEGMENTS"   first
for v in "" scratch/var_melseg.so; do for i in 1 2; do
  r=$(SB200_LIB=${v:+$PWD/$v} python bench.py --no-extra --kernel-only --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), d['config']['parity_check'][:30])")
  echo "lib=${v:-main (ELL)} us=$r"
done; done
