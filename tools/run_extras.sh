#!/bin/bash
# Extra measurement points of BASELINE.json (run on the GPU box): training step (config 2), timestep sweep (config 4),
# and the reference's own GPU execution model (PyTorch eager) for comparison.  Output: gpurun_out/extras.jsonl
out=gpurun_out/extras.jsonl
: > $out
run() { echo "# $*" >> $out; timeout 600 python bench.py "$@" 2>gpurun_out/extras_err.log | tail -1 >> $out || tail -3 gpurun_out/extras_err.log >> $out; }
for T in 1 5 10 20; do run --batch 16 --T $T --steps 20 --warmup 3 --input-sets 2 --no-cpu-baseline; done
run --mode train --batch 16 --T 5 --steps 5 --warmup 2 --input-sets 2 --no-cpu-baseline
run --mode train --batch 16 --T 5 --planes 2 --steps 5 --warmup 2 --input-sets 2 --no-cpu-baseline
run --impl reference --reference-device cuda --batch 8 --T 5 --steps 10 --warmup 3
run --impl reference --reference-device cuda --mode train --batch 16 --T 5 --steps 5 --warmup 2
run --impl reference --mode train --steps 2 --warmup 1
