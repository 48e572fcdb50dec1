python -m pytest tests/test_gpu_regnet.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/pytest12.log 2>&1; tail -3 gpurun_out/pytest12.log
export MVSB200_TC_NO_TMA=1
for L in 3dconv0_1 3dconv1_1 3dconv6_2; do
for ZF in 1 2 4; do
  for DBG in 0 2 3 7; do
  MVSB200_TC_ZF=$ZF MVSB200_TC_DBG=$DBG MVSB200_TC_VERBOSE=1 python tools/run_layer.py --layer $L --iters 4 2>&1 | tail -2 | sed "s/^/ZF=$ZF /"
  done
done; done > gpurun_out/exp2.log 2>&1
unset MVSB200_TC_ZF
python tools/stage_bench.py --skip-cv --layers --regnet bf16 --out gpurun_out/stage8.json > gpurun_out/stage8.log 2>&1
