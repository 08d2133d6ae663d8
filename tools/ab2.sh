for p in 3 20 200; do for i in 1 2; do
  r=$(SB200_BENCH_SAMPLER_MS=$p python bench.py --no-extra --kernel-only --steps 200 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), d['clocks']['samples'])")
  echo "sampler period ${p} ms run $i: us/step, samples = $r"
done; done
for v in noepi nomel; do
  r=$(SB200_BENCH_NO_CHECK=1 SB200_LIB=$PWD/scratch/var_$v.so python bench.py --no-extra --kernel-only --steps 300 2>/dev/null | python -c "import json,sys; print(round(json.loads(sys.stdin.read())['ms_per_step']*1e3,2))")
  echo "ablation $v: $r us"
done
python -m pytest tests -m gpu -q -x -k "stft_loss or mstft or get_stft_torch or concurrent or trim or real_recording" 2>&1 | tail -3
bash tools/ab_mstft.sh
