mkdir -p gpurun_out
for mode in 0 1; do
  ACLB200_PIPELINE=$mode timeout 600 python bench.py --workload c5 --no-e2e --no-cpu-baseline --no-extra --steps 100 --warmup 5 > gpurun_out/bench_c5_pipeline$mode.json 2> gpurun_out/bench_c5_pipeline$mode.err
  python -c "import json; d=json.load(open('gpurun_out/bench_c5_pipeline$mode.json')); print('c5 pipeline=$mode', d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']); print({k:(v.get('kernel_ms', v.get('call_ms')), round(v['roofline']['frac'],4)) for k,v in d['workloads'].items()})"
