mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_error_metric.py -m gpu -q > gpurun_out/t_err.log 2>&1; tail -5 gpurun_out/t_err.log
timeout 600 python tools/profile_error_metric.py 4096 5 > gpurun_out/prof_err.log 2>&1; tail -5 gpurun_out/prof_err.log
ACLB200_ERROR_WARPS=6 timeout 600 python tools/profile_error_metric.py 4096 5 > gpurun_out/prof_err_w6.log 2>&1; tail -4 gpurun_out/prof_err_w6.log
timeout 600 python tools/profile_error_metric.py 4096 4 1024 > gpurun_out/prof_err_1g.log 2>&1; tail -3 gpurun_out/prof_err_1g.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/err_launches2.csv python tools/profile_error_metric.py 4096 2 > gpurun_out/ncu_err_launches.log 2>&1; grep object_space gpurun_out/err_launches2.csv | tail -2 | cut -c1-30,330-420
ACLB200_ERROR_WARPS=6 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/err_launches2_w6.csv python tools/profile_error_metric.py 4096 2 > gpurun_out/ncu_err_launches.log 2>&1; grep object_space gpurun_out/err_launches2_w6.csv | tail -2 | cut -c1-30,330-440
