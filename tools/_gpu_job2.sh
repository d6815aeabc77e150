export PYTHONPATH=.
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2aa_bench_n2.json 2> gpurun_out/r2aa_err.log
tail -3 gpurun_out/r2aa_err.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2aa_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
print(d.get('train'))
P
