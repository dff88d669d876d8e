set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pk_more.py -q > gpurun_out/pytest_more2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_more2.log
tail -15 gpurun_out/pytest_more2.log
timeout 120 python profiles/bench_siblings.py 512 5 > gpurun_out/siblings2.md 2> gpurun_out/siblings2.err; tail -14 gpurun_out/siblings2.md; tail -5 gpurun_out/siblings2.err
timeout 300 python profiles/run_config3.py 1024 3 zeldovich > gpurun_out/config3_zeldovich.json 2> gpurun_out/config3_zeldovich.err; cat gpurun_out/config3_zeldovich.json; tail -5 gpurun_out/config3_zeldovich.err
timeout 200 python profiles/run_config3.py 1024 3 uniform > gpurun_out/config3_uniform.json 2> gpurun_out/config3_uniform.err; cat gpurun_out/config3_uniform.json; tail -5 gpurun_out/config3_uniform.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"shell_bin_kernel|mode_kernel" -c 6 -o gpurun_out/prof_shell -f python profiles/run_stage.py shell 512 1 > gpurun_out/prof_shell.log 2>&1; tail -3 gpurun_out/prof_shell.log
