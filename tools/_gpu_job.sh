export PYTHONPATH=.
timeout 600 python -m pytest tests -m gpu -x -q -k "host_pipeline" 2>&1 | tail -5
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-train --no-parity > gpurun_out/r2bo_bench_$i.json 2> gpurun_out/r2bo_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r2bo_bench_$i.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['e2e']['ms_per_step'])"
done
tail -3 gpurun_out/r2bo_err.log
