export NCCL_DEBUG=WARN
for C in small cfg5; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dslab_check.py --config $C --iid > gpurun_out/dslab_$C.log 2>&1; echo "rc=$?"; grep "^{" gpurun_out/dslab_$C.log
done
