mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transform_tracks_pipeline_kernel -s 3 -c 1 -f -o gpurun_out/c5_kernel python bench.py --workload c5 --no-e2e --no-cpu-baseline --no-extra --steps 2 --warmup 3 > gpurun_out/ncu_c5.log 2>&1; tail -3 gpurun_out/ncu_c5.log
