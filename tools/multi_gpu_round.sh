# Multi-GPU round (run under `gpurun --gpus 8`): the 2-GPU D-slab test (incl. the abort word), then the bench modes the
# driver's SCALE run uses -- view-sharded config 2, config 3 (49 views sharded + gathered), config 5 as depth slabs.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_dslab.py tests/test_gpu_e2e.py -m gpu -q > gpurun_out/mg_tests.log 2>&1; tail -3 gpurun_out/mg_tests.log
for N in 2 4 8; do
  timeout 600 $TR --nproc-per-node $N --master-port 2960$N bench.py --gpus $N --config cfg5 --mode dslab --steps 20 --warmup 3 > gpurun_out/mg_cfg5_n$N.json 2> gpurun_out/mg_cfg5_n$N.err
done
timeout 600 python bench.py --config cfg5 --steps 20 --warmup 3 > gpurun_out/mg_cfg5_n1.json 2> gpurun_out/mg_cfg5_n1.err
timeout 600 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/mg_cfg2_n8.json 2> gpurun_out/mg_cfg2_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --config cfg3 > gpurun_out/mg_cfg3_n8.json 2> gpurun_out/mg_cfg3_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --config cfg4 --steps 5 --warmup 3 > gpurun_out/mg_cfg4_n8.json 2> gpurun_out/mg_cfg4_n8.err
for f in gpurun_out/mg_*.json; do echo "== $f"; tail -c 700 $f; echo; done
