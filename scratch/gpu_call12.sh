mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pk.py tests/test_gpu_dist.py -x -q > gpurun_out/pytest_pk.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pk.log
tail -15 gpurun_out/pytest_pk.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scratch/dist_breakdown.py 1024 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tee gpurun_out/dist_breakdown3.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench12.json 2> gpurun_out/bench12.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench12.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages_ms'], d['roofline']['avg_launch_ms'])
PY
