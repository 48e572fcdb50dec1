for M in 0 3 2; do
MVSB200_CV_MINB=$M timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('minb $M', d['value'], d['config']['stage_ms']['cost_volume'])"
done
