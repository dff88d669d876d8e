mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dep7.csv python profiles/run_stage.py deposit CIC tiled 512 2 zeldovich > /dev/null 2>&1
grep -E "tile_" gpurun_out/launches_dep7.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -5
