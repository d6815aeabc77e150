export PYTHONPATH=.
timeout 900 ncu --set full --clock-control none -k regex:conv_i8_kernel --launch-skip 21 --launch-count 21 -f -o /tmp/prof_i8_r2x python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity --no-train > gpurun_out/r2x_ncu_full.log 2>&1
tail -1 gpurun_out/r2x_ncu_full.log | cut -c1-100
ncu -i /tmp/prof_i8_r2x.ncu-rep --page raw --csv > gpurun_out/r2x_ncu_conv_i8_raw.csv 2>/dev/null
python tools/ncu_conv_table.py gpurun_out/r2x_ncu_conv_i8_raw.csv | cut -c1-200
timeout 800 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 6000 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-parity --no-train > gpurun_out/r2x_ncu_list.log 2>&1
python tools/launch_table.py gpurun_out/r2x_launches.csv 1 | tail -14
