# scratch driver for GPU-box runs (edited per experiment)
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
