export MVSB200_TC_VERBOSE=1 MVSB200_REGNET_PROFILE=1 MVSB200_TC_PROF=1
timeout 300 python tools/stage_bench.py --skip-cv --regnet bf16 --out gpurun_out/tmp.json 2>&1 | grep -A4 "^\[tc\] mode=0 Cin=32 Cout=8\|^\[tc\] mode=1 Cin=32 Cout=16\|^\[tc\] mode=0 Cin=8 " | grep -v CTAs | cut -c1-250 | tail -24
