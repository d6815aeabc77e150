export PYTHONPATH=.
R=r2bt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_err.log
python -c "
import json
d=json.loads(open('gpurun_out/${R}_bench.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['roofline']['frac'], d['train']['ms_per_step'], d['cpu_baseline']['value'], d['parity']['mde_abs_diff'])"
tail -3 gpurun_out/${R}_err.log
timeout 600 python -m pytest tests -m gpu -x -q -k "graph or pipeline or smoke" 2>&1 | tail -3
