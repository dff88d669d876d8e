set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pk_more.py -q > gpurun_out/pytest_more3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_more3.log
tail -30 gpurun_out/pytest_more3.log
