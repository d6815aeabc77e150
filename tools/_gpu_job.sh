export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "bit_identical or first_layer or teacher_forced" 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-parity --no-train > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_err.log
SS_EPW16=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-parity --no-train > gpurun_out/r2ai_bench_epw8.json 2>> gpurun_out/r2ai_err.log
python - <<'P'
import json
for f in ('gpurun_out/r2ai_bench.json','gpurun_out/r2ai_bench_epw8.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
    print(d['roofline']['per_block_ms'])
P
