timeout 600 python -m pytest tests/test_gpu_regnet.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -2
MVSB200_REGNET_PROFILE=1 timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 --out gpurun_out/tmp.json 2>&1 | grep "\[regnet\]" | tail -12
timeout 600 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['stage_ms'])"
