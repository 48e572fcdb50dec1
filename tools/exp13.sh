MVSB200_TC_VERBOSE=1 MVSB200_REGNET_PROFILE=1 timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 --out gpurun_out/tmp.json 2>&1 | grep "^\[tc\]\|\[regnet\]" | tail -25
