# Round profile (run under gpurun, ONE GPU): (1) plain bench run, (2) ncu launch list of the same command, (3) full
# capture of the dominant kernel and of the regularizer's launches.  tools/summarize_profiles.py <tag> then turns
# gpurun_out/ into the tracked summaries under profiles/ (and fails if the captured kernel is not the one bench.py names).
set -x
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench_plain.json 2> gpurun_out/prof_bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/infer_once.py > gpurun_out/once.log 2>&1 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:cost_volume_window -c 1 -o gpurun_out/prof_cv python tools/infer_once.py > gpurun_out/ncu_cv.log 2>&1
ncu -i gpurun_out/prof_cv.ncu-rep --page raw --csv > gpurun_out/prof_cv_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_cv.ncu-rep --page details > gpurun_out/prof_cv_details.txt 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv3d_tc_kernel -c 12 -o gpurun_out/prof_conv python tools/infer_once.py > gpurun_out/ncu_conv.log 2>&1
ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_conv.ncu-rep --page details > gpurun_out/prof_conv_details.txt 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv2d_tc_kernel -s 5 -c 1 -o gpurun_out/prof_tower python tools/tower_bench.py --precisions bf16 > gpurun_out/ncu_tower.log 2>&1
ncu -i gpurun_out/prof_tower.ncu-rep --page raw --csv > gpurun_out/prof_tower_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_tower.ncu-rep --page details > gpurun_out/prof_tower_details.txt 2>/dev/null
timeout 300 python tools/tower_bench.py > gpurun_out/tower_bench.log 2>&1
# tensor-pipe counters that see tcgen05 (UTCHMMA): whatever this ncu names them
ncu --query-metrics 2>/dev/null | grep -i -E "tensor|tmem|utc|tcgen" > gpurun_out/ncu_tensor_metric_names.txt
ls -la gpurun_out/*.ncu-rep
