mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ma.py -x -q > gpurun_out/pytest_ma.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ma.log
tail -4 gpurun_out/pytest_ma.log
python scratch/time_deposit.py 512 tiled NGP,CIC,TSC,PCS 2>&1 | tee gpurun_out/time6.log
python scratch/time_deposit.py 512 tiled CIC,PCS zeldovich 2>&1 | tee -a gpurun_out/time6.log
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none --csv --log-file gpurun_out/launches_dep6.csv python profiles/run_stage.py deposit CIC tiled 512 2 > /dev/null 2>&1
grep -E "tile_" gpurun_out/launches_dep6.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $NF}' | tail -8
