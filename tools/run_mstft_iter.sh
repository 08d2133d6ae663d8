# mstft iteration run: parity tests of the mstft family, step times, one ncu capture of the fused N=2048 kernel (source page to CSV)
O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "stft_loss or mstft or stft_torch or real_recording" 2>&1 | tail -3
for w in mstft mstft_specs; do
  r=$(python bench.py --no-extra --kernel-only --workload $w --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), d['gpu_launches'], d['config']['parity_check'])")
  echo "workload=$w us/step launches check: $r"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_mstft.csv python bench.py --workload mstft --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_mstft_specs.csv python bench.py --workload mstft_specs --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mstft_bwd -s 9 -c 1 -o /tmp/mf -f python bench.py --workload mstft --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
ncu -i /tmp/mf.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $O/ncu_mstft_fused_source.csv.gz
ncu -i /tmp/mf.ncu-rep --page raw --csv > $O/ncu_mstft_fused_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:mstft_fwd -s 9 -c 1 -o /tmp/mw -f python bench.py --workload mstft_specs --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
ncu -i /tmp/mw.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $O/ncu_mstft_fwd_source.csv.gz
grep -h "mstft\|grad_ola" $O/launches_mstft.csv | tail -6 | cut -d, -f2,5,6
grep -h "mstft\|grad_ola" $O/launches_mstft_specs.csv | tail -9 | cut -d, -f2,5,6
