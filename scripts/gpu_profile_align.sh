KRE='regex:align_kernel|seg_tile'
DNB_SPLIT_ALIGN=1 ncu --set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on -k "$KRE" -s 3 -c 3 -o gpurun_out/r1c_full python scripts/quick_perf.py 500 10000 4 > gpurun_out/r1c_full.log 2>&1
ncu -i gpurun_out/r1c_full.ncu-rep --page raw --csv > gpurun_out/r1c_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r1c_full.ncu-rep --page source --csv --kernel-name regex:align_kernel > gpurun_out/r1c_align_source.csv 2>/dev/null
ncu -i gpurun_out/r1c_full.ncu-rep --page source --csv --kernel-name regex:seg_tile > gpurun_out/r1c_segtile_source.csv 2>/dev/null
tail -3 gpurun_out/r1c_full.log
