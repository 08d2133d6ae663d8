#!/bin/bash
# Refresh of the Griffin-Lim part of the evidence (tools/collect_profiles.sh does everything): launch lists, one ncu capture of the
# batch kernel, the bench line.  Outputs under gpurun_out/r02/evidence/.
O=gpurun_out/r02/evidence; mkdir -p $O
python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"
for w in griffinlim griffinlim_tt griffinlim_batch; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1; echo "launch list $w rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:gl2_kernel -s 8 -c 2 -o /tmp/gl2_batch -f python bench.py --workload griffinlim_batch --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
ncu -i /tmp/gl2_batch.ncu-rep --page raw --csv > $O/ncu_gl2_batch_raw.csv 2>/dev/null
ncu -i /tmp/gl2_batch.ncu-rep --page source --csv 2>/dev/null | gzip > $O/ncu_gl2_batch_source.csv.gz
echo "capture gl2_batch done"
