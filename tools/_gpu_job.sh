export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q -k "analog or standalone" 2>&1 | tail -30 | tee gpurun_out/r2bk_tests.log
