timeout 300 python tools/infer_once.py --fast-features > gpurun_out/once.log 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cost_volume_c32 -c 1 -o gpurun_out/prof_cv python tools/infer_once.py --fast-features > gpurun_out/ncu_cv.log 2>&1
ncu -i gpurun_out/prof_cv.ncu-rep --page raw --csv > gpurun_out/prof_cv_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_cv.ncu-rep --page details > gpurun_out/prof_cv_details.txt 2>/dev/null
