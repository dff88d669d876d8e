python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scratch/dist_breakdown.py 1024 2>&1 | grep -v "OMP_NUM\|\*\*\*"
