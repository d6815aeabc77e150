export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "fold or bit_identical or benchmark_config" 2>&1 | tail -5
for v in 1 0; do
SS_FOLD_ROWS_BY_LIST=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-parity --no-train > gpurun_out/r2ap_bench_$v.json 2> gpurun_out/r2ap_err.log
python - <<P
import json
d=json.loads(open('gpurun_out/r2ap_bench_$v.json').read().strip().splitlines()[-1])
print($v, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
print(d['roofline']['per_block_ms'])
P
done
