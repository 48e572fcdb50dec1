timeout 600 python -m pytest tests/test_gpu_regnet.py tests/test_gpu_e2e.py tests/test_gpu_cost_volume.py -m gpu -x -q > gpurun_out/pytest13.log 2>&1; tail -15 gpurun_out/pytest13.log
