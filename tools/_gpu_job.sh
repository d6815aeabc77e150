export PYTHONPATH=.
R=r2bq
for i in 1 2; do
timeout 600 python bench.py --mode train --batch 16 --T 5 --steps 20 --warmup 10 --no-cpu-baseline --no-extras --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', round(d['ms_per_step'],3), round(d['value'],1))"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_err.log
python -c "
import json
d=json.loads(open('gpurun_out/${R}_bench.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['roofline']['frac'], d['train']['ms_per_step'], d['cpu_baseline']['value'], d['analog_model']['cpu_oracle_frames_per_s'])"
