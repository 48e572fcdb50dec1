# development aid: cost-volume + whole-path tests, then one bench line (stage times)
timeout 600 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['config']['stage_ms'], d['e2e']['depth_checksum'])"
