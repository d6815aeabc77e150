export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "gradients or end_to_end or golden or smoke or contract" 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2ae_train_launches.csv python bench.py --mode train --batch 16 --T 5 --steps 1 --warmup 1 --input-sets 1 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2ae_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2ae_train_launches.csv 40 | grep -i "step total\|head"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_err.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2ae_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_block_ms']['heads'], d['train']['ms_per_step'])
P
