for K in 2 4 8; do
MVSB200_CV_KDC=$K timeout 600 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('KDC',$K, d['value'], d['config']['stage_ms'])"
done
