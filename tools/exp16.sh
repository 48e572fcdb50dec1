for DBG in 0 8; do
echo "== DBG=$DBG"
MVSB200_TC_DBG=$DBG MVSB200_REGNET_PROFILE=1 timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 --out gpurun_out/tmp.json 2>&1 | grep "\[regnet\]" | tail -12 | tr '\n' ' '
echo
done
