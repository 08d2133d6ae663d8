# A/B of the mstft step: the resolutions' grids concatenated into one launch (default), per-resolution launches on side streams
# (SB200_MSTFT_STREAMS=1, round 1's formulation), and ONE cooperative launch with the tail folded in (SB200_MSTFT_SINGLE=1)
for e in "" SB200_MSTFT_STREAMS=1 SB200_MSTFT_SINGLE=1; do for w in mstft mstft_specs; do
  r=$(env $e python bench.py --no-extra --kernel-only --workload $w --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), d['gpu_launches'], d['config']['parity_check'])")
  echo "${e:-default (one grid)} workload=$w us/step launches check: $r"
done; done
