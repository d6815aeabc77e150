export PYTHONPATH=.
python tools/dgrad_determinism.py 2>&1 | tail -12
export STEREOSPIKE_B200_LIB=build/timing/libstereospike_b200.so
for d in 0 1 2 4 6 7; do SS_WG_DBG=$d python tools/wgrad_probe.py 2>&1 | grep dbg; done
for d in 0 1 2 4 6; do SS_WGRAD_N64=0 SS_WG_DBG=$d python tools/wgrad_probe.py 2>&1 | grep dbg; done
