for C in cfg1 cfg5; do
timeout 600 python bench.py --config $C --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$C.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$C', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['stage_ms'])" || tail -5 gpurun_out/bench_$C.err
done
