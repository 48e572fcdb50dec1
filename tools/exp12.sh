export MVSB200_TC_LAYER=8,1,0 MVSB200_TC_VERBOSE=1 MVSB200_REGNET_PROFILE=1
run() { echo "== $*"; env "$@" timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 --out gpurun_out/tmp.json 2>&1 | grep "3dconv6_2\|Cout=1(+0)" | tail -2; }
run MVSB200_TC_DBG=0
run MVSB200_TC_DBG=1
run MVSB200_TC_DBG=2
run MVSB200_TC_DBG=4
run MVSB200_TC_DBG=3
run MVSB200_TC_DBG=6
run MVSB200_TC_DBG=7
run MVSB200_TC_TILE=12x9
run MVSB200_TC_TILE=24x9
run MVSB200_TC_TILE=24x19
run MVSB200_TC_ZF=2
run MVSB200_TC_ZF=1
