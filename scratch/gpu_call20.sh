set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pk_bin_walk" -s 1 -c 1 -o gpurun_out/prof_pkbin2 -f python profiles/run_stage.py pk 512 0 2 > gpurun_out/prof_pkbin2.log 2>&1; tail -3 gpurun_out/prof_pkbin2.log
