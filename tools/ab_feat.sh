# A/B of the n_fft 2048 feature kernels on config 3 (64 x 5 s): SB200_FEAT_KERNEL = 2 single role, 3 = 8 + 8 warps, 4 = 12 + 4 warps
for k in 4 3 2; do
  r=$(SB200_FEAT_KERNEL=$k python bench.py --no-extra --kernel-only --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), d['config']['parity_check'][:40])")
  echo "SB200_FEAT_KERNEL=$k us/step: $r"
done
