python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 --no-train > gpurun_out/r2i_bench.json 2>gpurun_out/r2i_err.log
tail -3 gpurun_out/r2i_err.log
python - <<'PY'
import json
for f in ('gpurun_out/r2i_bench.json',):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])
    print(d['roofline']['per_block_ms'])
    print(d.get('parity'))
PY
