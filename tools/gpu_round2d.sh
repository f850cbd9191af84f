mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_error_metric.py tests/test_cpp_shim.py -m gpu -q > gpurun_out/t_err.log 2>&1; tail -25 gpurun_out/t_err.log
timeout 600 python tools/profile_error_metric.py 4096 4 > gpurun_out/prof_err.log 2>&1; tail -6 gpurun_out/prof_err.log
for rows in 1 0; do
  ACLB200_BASE_ROWS=$rows timeout 600 python bench.py --workload c5 --no-e2e --no-cpu-baseline --no-extra --steps 100 --warmup 5 > gpurun_out/bench_c5_rows$rows.json 2> gpurun_out/bench_c5_rows$rows.err
  python -c "import json; d=json.load(open('gpurun_out/bench_c5_rows$rows.json')); print('c5 rows=$rows', d['roofline']['kernel_ms'], d['roofline']['frac'], d['other_math']['kernel_ms'])"
done
ACLB200_BASE_ROWS=0 timeout 600 python bench.py --workload c3 --no-e2e --no-cpu-baseline --no-extra --steps 50 --warmup 5 > gpurun_out/bench_c3_rows0.json 2> gpurun_out/bench_c3_rows0.err
python -c "import json; d=json.load(open('gpurun_out/bench_c3_rows0.json')); print('c3 rows=0', d['roofline']['kernel_ms'], d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:object_space_kernel -c 1 -f -o gpurun_out/err_kernel_v3 python tools/profile_error_metric.py 1024 1 > gpurun_out/ncu_err.log 2>&1; tail -2 gpurun_out/ncu_err.log
