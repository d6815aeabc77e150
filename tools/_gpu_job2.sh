export PYTHONPATH=.
N=${NG:-2}
for conf in "expandable_segments:True" "expandable_segments:False" "expandable_segments:True"; do
PYTORCH_CUDA_ALLOC_CONF=$conf timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-train --no-parity --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$conf', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), round(d['other_state_policy']['ms_per_step'],4))"
done
