export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "gradients or standalone or penal" 2>&1 | tail -3
for v in 1 1 0; do
SS_WGRAD_STREAM=$v timeout 600 python bench.py --mode train --batch 16 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2bb_train_$v.json 2> gpurun_out/r2bb_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r2bb_train_$v.json').read().strip().splitlines()[-1]); print($v, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3))"
done
