export PYTHONPATH=.
R=r2bj
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${R}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${R}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_err.log
python -c "
import json
d=json.loads(open('gpurun_out/${R}_bench.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline']['per_block_ms']['heads'], d['train']['ms_per_step'], d['parity']['mde_abs_diff'], d['clocks'])"
tail -3 gpurun_out/${R}_err.log
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 6000 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-parity --no-train > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${R}_train_launches.csv python bench.py --mode train --batch 16 --T 5 --steps 1 --warmup 1 --input-sets 1 --no-cpu-baseline --no-extras --no-parity > /dev/null 2>&1
ls -la gpurun_out/${R}_*
