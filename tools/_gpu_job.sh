export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "fold or bit_identical or benchmark_config or full_size or graphed" 2>&1 | tail -5
for v in 1 0 1 0; do
SS_FOLD_OVERLAP=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-parity --no-train > gpurun_out/r2ax_bench_$v.json 2> gpurun_out/r2ax_err.log
python - <<P
import json
d=json.loads(open('gpurun_out/r2ax_bench_$v.json').read().strip().splitlines()[-1])
print($v, round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['per_block_ms'])
P
done
