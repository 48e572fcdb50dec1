timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest14.log 2>&1; tail -4 gpurun_out/pytest14.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json; tail -3 gpurun_out/bench_r1b.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
