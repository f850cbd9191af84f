mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_error_metric.py tests/test_cpp_shim.py -m gpu -q > gpurun_out/t_err.log 2>&1; tail -8 gpurun_out/t_err.log
timeout 600 python tools/profile_error_metric.py 4096 4 > gpurun_out/prof_err.log 2>&1; tail -6 gpurun_out/prof_err.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_error_metric.py --deselect tests/test_cpp_shim.py > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_error_metric.py -m gpu -x -q -k "many_clips or negative_scales or additive_base and c1 or local_to_object_space and mixed or scalar_compression and float3 or flags" > gpurun_out/sanitizer_error_metric_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer_error_metric_$tool.log; tail -5 gpurun_out/sanitizer_error_metric_$tool.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:object_space_kernel -c 1 -f -o gpurun_out/err_kernel_v4 python tools/profile_error_metric.py 1024 1 > gpurun_out/ncu_err.log 2>&1; tail -2 gpurun_out/ncu_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/err_launches.csv python tools/profile_error_metric.py 4096 2 > gpurun_out/ncu_err_launches.log 2>&1; tail -7 gpurun_out/err_launches.csv | cut -c1-60,150-420
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']); print(json.dumps(d['workloads']['error_metric'])[:1500])"
