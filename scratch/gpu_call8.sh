mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ma.py -x -q > gpurun_out/pytest_ma.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ma.log
tail -4 gpurun_out/pytest_ma.log
python scratch/time_deposit.py 512 tiled NGP,CIC,TSC,PCS 2>&1 | tee gpurun_out/time8.log
python scratch/time_deposit.py 512 tiled CIC,PCS zeldovich 2>&1 | tee -a gpurun_out/time8.log
