export PYTHONPATH=.
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ay_bench.json 2> gpurun_out/r2ay_err.log
tail -2 gpurun_out/r2ay_err.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2ay_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])
print(d['roofline']['per_block_ms'])
print(d['train']['ms_per_step'], d['train']['value'], d['parity']['mde_abs_diff'], d['parity']['teacher_forced'])
print({k:round(v['event_frames_per_s']) for k,v in d['timestep_sweep']['results'].items()}, d['sj_cupy_proxy']['speedup_vs_fp32'], d['sj_cupy_proxy']['speedup_vs_tf32_allowed'], d['other_state_policy']['value'], d['cpu_baseline']['value'])
P
