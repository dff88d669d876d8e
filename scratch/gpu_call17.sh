set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_pk_more.py -q > gpurun_out/pytest_more.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_more.log
tail -25 gpurun_out/pytest_more.log
timeout 120 python profiles/bench_siblings.py 512 5 > gpurun_out/siblings.md 2> gpurun_out/siblings.err; tail -20 gpurun_out/siblings.md; tail -5 gpurun_out/siblings.err
timeout 200 python -m pytest tests/test_gpu_pk.py -x -q > gpurun_out/pytest_pk.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pk.log
tail -4 gpurun_out/pytest_pk.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench17.json 2> gpurun_out/bench17.err; cat gpurun_out/bench17.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_siblings.csv python profiles/bench_siblings.py 256 1 > gpurun_out/launches_siblings.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_pk_more.py --deselect tests/test_gpu_pk.py > gpurun_out/pytest_rest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_rest.log
tail -4 gpurun_out/pytest_rest.log
