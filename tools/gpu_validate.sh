#!/bin/sh
# What the builder runs on the GPU box before a round ends (one gpurun call): the whole GPU suite, the smoke test, compute-sanitizer over
# the error metric kernels, the driver's bench command, and the ncu launch list of a short bench run. Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -6 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_error_metric.py -m gpu -x -q -k "many_clips or negative_scales or additive_base and c1 or local_to_object_space and mixed or scalar_compression and float3 or flags or matrix_metric and (mixed or mirrored) or all_samples" > gpurun_out/sanitizer_error_metric_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer_error_metric_$tool.log; tail -4 gpurun_out/sanitizer_error_metric_$tool.log
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']); print({k:(v.get('kernel_ms', v.get('call_ms')), round(v['roofline']['frac'],4)) for k,v in d['workloads'].items()})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; grep -c . gpurun_out/launches.csv
