mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dist.log
tail -25 gpurun_out/pytest_dist.log
for mode in peer nccl; do
PYL_TRANSPOSE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scratch/dist_breakdown.py 640 2>&1 | grep "N=640"
PYL_TRANSPOSE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 scratch/dist_breakdown.py 1024 2>&1 | grep "N=1024"
done
