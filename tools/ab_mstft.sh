# A/B of the mstft step: single cooperative launch for all resolutions vs per-resolution launches on side streams
for s in 1 0; do for w in mstft mstft_specs; do
  r=$(SB200_MSTFT_SINGLE=$s python bench.py --no-extra --kernel-only --workload $w --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), d['gpu_launches'], d['config']['parity_check'])")
  echo "single=$s workload=$w us/step launches check: $r"
done; done
