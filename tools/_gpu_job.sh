export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "standalone" 2>&1 | tail -30
timeout 800 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 6000 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-parity --no-train > gpurun_out/r2w_ncu_list.log 2>&1
python tools/launch_table.py gpurun_out/r2w_launches.csv 1
