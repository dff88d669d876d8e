set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_dist.py -x -q -k "4-64-peer or 4-45-peer" > gpurun_out/pytest_dist24.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dist24.log
tail -4 gpurun_out/pytest_dist24.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench24_n4.json 2> gpurun_out/bench24_n4.err; cat gpurun_out/bench24_n4.json; tail -3 gpurun_out/bench24_n4.err
