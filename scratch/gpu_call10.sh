mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scratch/dist_breakdown.py 640 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tee gpurun_out/dist_breakdown.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scratch/dist_breakdown.py 1024 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tee -a gpurun_out/dist_breakdown.log
timeout 600 python -m pytest tests/test_gpu_ma.py -x -q -k "interp or stream" 2>&1 | tail -3
