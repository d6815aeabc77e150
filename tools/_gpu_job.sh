export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "gradients or golden or standalone or fold or penal or firing" 2>&1 | tail -6
timeout 600 python bench.py --mode train --batch 16 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2ah_train.json 2> gpurun_out/r2ah_err.log
tail -2 gpurun_out/r2ah_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r2ah_train.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
