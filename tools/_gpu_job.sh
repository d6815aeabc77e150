export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "fold or graphed or benchmark_config or contract" 2>&1 | tail -3
python tools/bench_latency.py 2>&1 | grep "^{" > gpurun_out/r2ao_latency_graph.jsonl
cat gpurun_out/r2ao_latency_graph.jsonl
