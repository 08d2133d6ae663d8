mkdir -p gpurun_out/r02b
python -m pytest tests -m gpu -x -q -k "stft_loss or mstft or stft_torch or real_recording" 2>&1 | tail -3
for v in "" scratch/var_mu4.so scratch/var_mu1.so; do for w in mstft mstft_specs; do
  r=$(SB200_LIB=${v:+$PWD/$v} python bench.py --no-extra --kernel-only --workload $w --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), d['gpu_launches'], d['config']['parity_check'])")
  echo "lib=${v:-main} workload=$w us/step launches check: $r"
done; done
