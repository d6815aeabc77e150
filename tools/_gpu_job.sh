export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "gradients or golden" 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2af_train_launches.csv python bench.py --mode train --batch 16 --T 5 --steps 1 --warmup 1 --input-sets 1 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2af_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2af_train_launches.csv 40 | grep -i "step total\|head"
timeout 600 python bench.py --mode train --batch 16 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2af_train.json 2> gpurun_out/r2af_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r2af_train.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
