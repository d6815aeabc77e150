export PYTHONPATH=.
for cfg in "128 4" "256 4" "128 2" "64 4"; do
set -- $cfg
echo "== SS_BAS_MIN_CIN=$1 SS_BAS_MIN_BATCH=$2"
SS_BAS_MIN_CIN=$1 SS_BAS_MIN_BATCH=$2 timeout 300 python tools/t1_sweep.py 16 8 32 2>&1 | grep -v "^$"
done | tee gpurun_out/r2bm_t1.log
timeout 600 python -m pytest tests -m gpu -x -q -k "independent_steps" 2>&1 | tail -3
