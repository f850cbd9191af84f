mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_error_metric.py -m gpu -q > gpurun_out/t_err.log 2>&1; tail -25 gpurun_out/t_err.log
timeout 600 python tools/profile_error_metric.py 2048 3 > gpurun_out/prof_err.log 2>&1; tail -6 gpurun_out/prof_err.log
timeout 600 python tools/profile_error_metric.py 2048 3 64 > gpurun_out/prof_err_64mb.log 2>&1; tail -4 gpurun_out/prof_err_64mb.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_error_metric.py > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 2500 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:object_space_kernel -c 1 -f -o gpurun_out/err_kernel python tools/profile_error_metric.py 1024 1 > gpurun_out/ncu_err.log 2>&1; tail -3 gpurun_out/ncu_err.log
