mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ma.py -x -q > gpurun_out/pytest_ma.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ma.log
tail -4 gpurun_out/pytest_ma.log
python scratch/time_deposit.py 512 tiled NGP,CIC,TSC,PCS 2>&1 | tee gpurun_out/time5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_scatter|tile_deposit" -s 2 -c 2 -o gpurun_out/prof_tiled_cic2 -f python profiles/run_stage.py deposit CIC tiled 512 2 > gpurun_out/prof_tiled2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cat gpurun_out/bench5.json
