O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "stft_loss or mstft or stft_torch or real_recording" 2>&1 | tail -2
python tools/probe_mstft_graph.py
for w in mstft mstft_specs; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_$w.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-(9 if "$w"=="mstft_specs" else 5):]: print(r[4][:60], r[8], r[-1])
PY
done
