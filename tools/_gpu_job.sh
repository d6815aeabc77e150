export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "fold or bit_identical or benchmark_config or full_size" 2>&1 | tail -5
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-parity --no-train > gpurun_out/r2aw_bench.json 2> gpurun_out/r2aw_err.log
python - <<P
import json
d=json.loads(open('gpurun_out/r2aw_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['frac'], d['roofline']['per_block_ms'])
P
done
