set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all21.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_all21.log
tail -4 gpurun_out/pytest_gpu_all21.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke21.log 2>&1; tail -2 gpurun_out/smoke21.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench21.json 2> gpurun_out/bench21.err; cat gpurun_out/bench21.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench21_ref.json 2> gpurun_out/bench21_ref.err; cat gpurun_out/bench21_ref.json
timeout 100 python profiles/bench_siblings.py 512 7 2>&1 | grep -E "Pk bin|XPk bin"
