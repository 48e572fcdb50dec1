MVSB200_TC_VERBOSE=1 MVSB200_REGNET_PROFILE=1 timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 --out gpurun_out/stage9.json > gpurun_out/stage9.log 2>&1
timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 --out gpurun_out/stage9b.json > gpurun_out/stage9b.log 2>&1
