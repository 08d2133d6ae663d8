python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/probe_mstft_graph.py
for w in mstft mstft_specs; do
  r=$(python bench.py --no-extra --kernel-only --workload $w --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), d['gpu_launches'], d['config']['parity_check'])")
  echo "workload=$w us/step launches check: $r"
done
