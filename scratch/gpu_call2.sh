mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ma.py -x -q > gpurun_out/pytest_ma.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ma.log
tail -5 gpurun_out/pytest_ma.log
for fs in 0 3 5; do PYL_FILL_SHIFT=$fs python scratch/time_deposit.py 512 tiled NGP,CIC,PCS; done 2>&1 | tee gpurun_out/time_fill.log
python scratch/time_deposit.py 512 deterministic NGP,CIC 2>&1 | tee -a gpurun_out/time_fill.log
python scratch/time_deposit.py 512 tiled CIC,PCS zeldovich 2>&1 | tee -a gpurun_out/time_fill.log
PYL_FILL_SHIFT=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dep2.csv python profiles/run_stage.py deposit CIC tiled 512 2 > /dev/null 2>&1
grep -E "tile_" gpurun_out/launches_dep2.csv | awk -F'","' '{print $5, $NF}' | tail -8
