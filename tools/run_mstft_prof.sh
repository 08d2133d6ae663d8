O=gpurun_out/r02b; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:mstft_multi_bwd -s 3 -c 1 -o /tmp/mf -f python bench.py --workload mstft --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
ncu -i /tmp/mf.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $O/ncu_mstft_multi_fused_source.csv.gz
ncu -i /tmp/mf.ncu-rep --page raw --csv > $O/ncu_mstft_multi_fused_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:mstft_multi_fwd -s 3 -c 1 -o /tmp/mw -f python bench.py --workload mstft_specs --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
ncu -i /tmp/mw.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $O/ncu_mstft_multi_fwd_source.csv.gz
ncu -i /tmp/mw.ncu-rep --page raw --csv > $O/ncu_mstft_multi_fwd_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:mstft_multi_bwd -s 3 -c 1 -o /tmp/mb -f python bench.py --workload mstft_specs --steps 3 --warmup 3 --kernel-only --no-extra > /dev/null 2>&1
ncu -i /tmp/mb.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $O/ncu_mstft_multi_bwd_source.csv.gz
ncu -i /tmp/mb.ncu-rep --page raw --csv > $O/ncu_mstft_multi_bwd_raw.csv 2>/dev/null
ls -la $O
