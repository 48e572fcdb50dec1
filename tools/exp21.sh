export NCCL_DEBUG=WARN
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/dslab_check.py --config cfg5 --iid > gpurun_out/dslab_cfg5_g8.log 2>&1; echo "rc=$?"; grep "^{" gpurun_out/dslab_cfg5_g8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 tools/dslab_check.py --config cfg5 --iid > gpurun_out/dslab_cfg5_g4.log 2>&1; echo "rc=$?"; grep "^{" gpurun_out/dslab_cfg5_g4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?"; cat gpurun_out/bench_n8.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
