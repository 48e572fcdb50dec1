# development aid: regularizer + whole-path tests, then bench lines (stage times) with and without an env switch ($1)
timeout 600 python -m pytest tests/test_gpu_regnet.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
for V in "" "$1"; do
env $V timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$V]', d['value'], d['e2e']['value'], d['config']['stage_ms'], d['e2e']['depth_checksum'])"
env $V timeout 600 python bench.py --config cfg1 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$V] cfg1', d['value'], d['config']['stage_ms'])"
done
