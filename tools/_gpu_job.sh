export PYTHONPATH=.
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2bf_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2bf_bench.json 2> gpurun_out/r2bf_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r2bf_bench.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['roofline']['frac'], d['per_block_ms']['heads'], d['train']['ms_per_step'], d['parity'])"
tail -3 gpurun_out/r2bf_err.log
