ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 1300 -c 150 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 4 --warmup 10 --no-train --no-parity --no-cpu-baseline > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log | cut -c1-300
