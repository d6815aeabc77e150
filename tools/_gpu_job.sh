export PYTHONPATH=.
R=r2bn
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${R}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${R}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_err.log
python -c "
import json
d=json.loads(open('gpurun_out/${R}_bench.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline']['per_block_ms']['heads'], d['train']['ms_per_step'], d['parity']['mde_abs_diff'], d['clocks']); print(d['analog_model']); print({k:round(v['event_frames_per_s']) for k,v in d['timestep_sweep']['results'].items()})"
tail -3 gpurun_out/${R}_err.log
