#!/bin/sh
# development helper: bench every library variant under build_variants/ (extra bench.py arguments: $BENCH_ARGS)
for lib in build_variants/lib_*.so; do
  ACLB200_LIB=$PWD/$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); o=d.get('other_math') or {}
print('$lib', round(d['value']/1e9,2), 'G/s', round(d['roofline']['kernel_ms'],4), 'ms', round(d['roofline']['frac'],4), '| other math', o.get('math'), round(o.get('ms_per_step',0),4), 'ms')"
done
