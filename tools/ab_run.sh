# scratch driver for GPU-box experiments (edited per experiment)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wpt or WPT or packet" 2>&1 | tail -3 > gpurun_out/pytest_tail.log
python tools/sweep.py c3 "" wpt_inplace=0 "" wpt_inplace=0 wpt_tile=1024,wpt_threads=96 wpt_tile=4096,wpt_threads=288 > gpurun_out/ab_inpl.log 2>&1
cat gpurun_out/pytest_tail.log gpurun_out/ab_inpl.log
