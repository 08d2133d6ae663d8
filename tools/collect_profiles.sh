#!/bin/bash
# Round-2 evidence run (one GPU): full GPU test suite, the bench line, ncu launch lists of the bench commands and one
# `ncu --set full` capture of each hot kernel, exported on the box to CSV (raw metrics + per-instruction source page) because
# gpurun brings back at most 64 MiB.  Outputs under gpurun_out/r02/evidence/ (summarised into profiles/ by tools/summarise_profiles.py).
O=gpurun_out/r02/evidence; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"
for w in stft_mel griffinlim griffinlim_tt griffinlim_batch mstft mstft_specs; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1; echo "launch list $w rc=$?"
done
cap() {   # name, kernel regex, skip, count, workload
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/$1 -f python bench.py --workload $5 --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $O/ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/ncu_$1_source.csv.gz
  rm -f /tmp/$1.ncu-rep; echo "capture $1 done"
}
cap feat3 stft_feature3 3 1 stft_mel
cap gl2_batch gl2_kernel 8 2 griffinlim_batch
cap mstft_fused mstft_multi_bwd 3 1 mstft
cap mstft_specs mstft_multi 6 2 mstft_specs
# ablations of the feature kernel (variants built by tools/variants.py): what the analysis role costs on its own, and without the mel
for v in noepi nomel; do
  r=$(SB200_BENCH_NO_CHECK=1 SB200_LIB=$PWD/scratch/var_$v.so python bench.py --no-extra --kernel-only --steps 300 2>/dev/null | python -c "import json,sys; print(round(json.loads(sys.stdin.read())['ms_per_step']*1e3,2))" 2>/dev/null)
  echo "ablation $v: $r us" | tee -a $O/ablations.txt
done
du -sh $O
