set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_step.csv python profiles/run_stage.py step 512 6 > gpurun_out/launches_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_scatter|tile_deposit" -s 2 -c 2 -o gpurun_out/prof_tiled_cic -f python profiles/run_stage.py deposit CIC tiled 512 2 > gpurun_out/prof_tiled.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pk_bin" -s 1 -c 1 -o gpurun_out/prof_pkbin -f python profiles/run_stage.py pk 512 0 2 > gpurun_out/prof_pkbin.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench1.json
