set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -4 gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench16.json 2> gpurun_out/bench16.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench16_ref.json 2> gpurun_out/bench16_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_step16.csv python profiles/run_stage.py step 512 6 > gpurun_out/launches_step16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_count|tile_scatter|tile_deposit" -s 3 -c 3 -o gpurun_out/prof_tiled_cic3 -f python profiles/run_stage.py deposit CIC tiled 512 2 > gpurun_out/prof_tiled3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_scatter|tile_deposit" -s 2 -c 2 -o gpurun_out/prof_tiled_pcs3 -f python profiles/run_stage.py deposit PCS tiled 512 2 > gpurun_out/prof_tiled3p.log 2>&1
cat gpurun_out/bench16.json
