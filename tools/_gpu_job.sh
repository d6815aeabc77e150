export PYTHONPATH=.
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_err.log
tail -3 gpurun_out/r2y_err.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'])
print(d['roofline']['per_block_ms'])
print(d.get('train',{}).get('ms_per_step'), d.get('parity'))
print(d.get('cpu_baseline'))
P
