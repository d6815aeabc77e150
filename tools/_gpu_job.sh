export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_grad_umma.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q -k "gradients" 2>&1 | tail -4
python tools/wgrad_probe.py 2>&1 | grep "dbg"
SS_WGRAD_N64=0 python tools/wgrad_probe.py 2>&1 | grep "dbg"
timeout 600 python bench.py --mode train --batch 16 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2u_train.json 2> gpurun_out/r2u_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r2u_train.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
