export PYTHONPATH=.
timeout 600 python -m pytest tests -m gpu -x -q -k "heads_readout or single_step_contract or golden or smoke" 2>&1 | tail -15 | tee gpurun_out/r2bi_tests.log
for v in 1 1; do
SS_HEAD_TAPS_MMA=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-train > gpurun_out/r2bi_bench_$v.json 2> gpurun_out/r2bi_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r2bi_bench_$v.json').read().strip().splitlines()[-1]); print($v, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['roofline']['per_block_ms']['heads'], d['parity']['mde_abs_diff'])"
done
tail -3 gpurun_out/r2bi_err.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:head --launch-count 2 --csv --log-file gpurun_out/r2bi_heads_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity --no-train > /dev/null 2>&1
grep "gpu__time\|issue_active" gpurun_out/r2bi_heads_ncu.csv | cut -d, -f5,13- 
