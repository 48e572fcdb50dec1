export MVSB200_REGNET_PROFILE=1 MVSB200_TC_VERBOSE=1
run() { L=$1; N=$2; shift 2; echo "== $L $*"; env MVSB200_TC_LAYER=$L "$@" timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 --out gpurun_out/tmp.json 2>&1 | grep "$N \|mode=0 Cin=${L%%,*} " | tail -2 | cut -c1-200; }
run 32,8,0 3dconv0_1 MVSB200_TC_TILE=30x8
run 32,8,0 3dconv0_1 MVSB200_TC_TILE=30x4
run 32,8,0 3dconv0_1 MVSB200_TC_TILE=14x8
run 32,8,0 3dconv0_1 MVSB200_TC_TILE=14x16 MVSB200_TC_ZF=2
run 32,8,0 3dconv0_1 MVSB200_TC_TILE=30x8 MVSB200_TC_ZF=2
run 32,8,0 3dconv0_1 MVSB200_TC_TILE=30x16 MVSB200_TC_ZF=2
run 16,16,0 3dconv1_1 MVSB200_TC_TILE=30x8
run 16,16,0 3dconv1_1 MVSB200_TC_TILE=30x4
run 16,16,0 3dconv1_1 MVSB200_TC_TILE=14x16
run 16,16,0 3dconv1_1 MVSB200_TC_TILE=14x8
run 16,16,0 3dconv1_1 MVSB200_TC_TILE=30x8 MVSB200_TC_ZF=1
run 8,1,0 3dconv6_2 MVSB200_TC_TILE=30x16
run 8,1,0 3dconv6_2 MVSB200_TC_TILE=30x8
run 8,1,0 3dconv6_2 MVSB200_TC_TILE=14x16
