mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/pytest_dist4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dist4.log
tail -5 gpurun_out/pytest_dist4.log
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
echo "bench $n exit $?"; cat gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
done
