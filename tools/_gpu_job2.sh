export PYTHONPATH=.
N=${NG:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2an_bench_n$N.json 2> gpurun_out/r2an_err_n$N.log
tail -3 gpurun_out/r2an_err_n$N.log
python - <<P
import json
d=json.loads(open('gpurun_out/r2an_bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
print(d.get('train'))
P
