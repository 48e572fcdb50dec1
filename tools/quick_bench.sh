# development aid: regression + regularizer + whole-path tests, per-layer regularizer times, then bench lines (stage
# times) with and without an env switch ($1)
timeout 600 python -m pytest tests/test_gpu_regress.py tests/test_gpu_regnet.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
MVSB200_REGNET_PROFILE=1 timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 2>&1 | grep -E "total|3dconv" | tail -12 | tr "\n" " " | sed "s/\[regnet\] 3dconv//g; s/ ms//g"; echo
for V in "" "$1"; do
env $V timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$V]', d['value'], d['e2e']['value'], d['config']['stage_ms'], d['e2e']['depth_checksum'])"
done
