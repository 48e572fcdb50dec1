set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest11.log 2>&1; tail -3 gpurun_out/pytest11.log
for L in 3dconv0_1 3dconv1_0; do
for ZF in 1 2 4; do for NOTMA in 0 1; do
  if [ $NOTMA = 1 ]; then export MVSB200_TC_NO_TMA=1; else unset MVSB200_TC_NO_TMA; fi
  for DBG in 0 1 2 3 7; do
  MVSB200_TC_ZF=$ZF MVSB200_TC_DBG=$DBG MVSB200_TC_VERBOSE=1 python tools/run_layer.py --layer $L --iters 4 2>&1 | tail -2 | sed "s/^/ZF=$ZF NOTMA=$NOTMA /"
  done
done; done; done > gpurun_out/exp1.log 2>&1
