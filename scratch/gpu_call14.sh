mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pk.py tests/test_gpu_dist.py -x -q > gpurun_out/pytest_pk.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pk.log
tail -5 gpurun_out/pytest_pk.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scratch/dist_breakdown.py 1024 2>&1 | grep -v "OMP_NUM\|\*\*\*" | grep -v "^$" | tail -3
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench14_n$n.json 2> gpurun_out/bench14_n$n.err
python - <<PY
import json; d=json.load(open('gpurun_out/bench14_n$n.json')); print($n, d['ms_per_step'], d['value']/1e9, d['e2e']['ms_per_step'], d['stages_ms'])
PY
done
