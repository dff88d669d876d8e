set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/pytest_dist22.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dist22.log
tail -4 gpurun_out/pytest_dist22.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench22_n2.json 2> gpurun_out/bench22_n2.err; cat gpurun_out/bench22_n2.json; tail -3 gpurun_out/bench22_n2.err
