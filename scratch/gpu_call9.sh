mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "8-64" > gpurun_out/pytest_dist8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dist8.log
tail -4 gpurun_out/pytest_dist8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench 8 exit $?"; cat gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
