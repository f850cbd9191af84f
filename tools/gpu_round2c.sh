mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_error_metric.py -m gpu -q > gpurun_out/t_err.log 2>&1; tail -25 gpurun_out/t_err.log
timeout 600 python tools/profile_error_metric.py 4096 4 > gpurun_out/prof_err.log 2>&1; tail -6 gpurun_out/prof_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:object_space_kernel -c 1 -f -o gpurun_out/err_kernel_v2 python tools/profile_error_metric.py 1024 1 > gpurun_out/ncu_err.log 2>&1; tail -3 gpurun_out/ncu_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/err_launches.csv python tools/profile_error_metric.py 4096 2 > gpurun_out/ncu_err_launches.log 2>&1; tail -12 gpurun_out/err_launches.csv
