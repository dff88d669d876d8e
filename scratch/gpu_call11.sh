mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pk.py tests/test_gpu_dist.py -x -q > gpurun_out/pytest_pk.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pk.log
tail -15 gpurun_out/pytest_pk.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scratch/dist_breakdown.py 640 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tee gpurun_out/dist_breakdown2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scratch/dist_breakdown.py 1024 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tee -a gpurun_out/dist_breakdown2.log
