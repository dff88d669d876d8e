set -x
mkdir -p gpurun_out
for m in 0 6 8; do
  PYL_PKBIN_MINB=$m timeout 100 python profiles/bench_siblings.py 512 7 2>&1 | grep -E "Pk bin|XPk bin" | sed "s/^/MINB=$m /"
done | tee gpurun_out/minb.txt
PYL_PKBIN_MINB=8 timeout 200 python -m pytest tests/test_gpu_pk.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_pk_minb8.log
