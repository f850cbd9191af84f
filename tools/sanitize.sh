#!/bin/sh
# compute-sanitizer over the library's kernels (run on the GPU box): memcheck + racecheck (shared memory hazards of the mbarrier /
# TMA pipeline) + synccheck on tools/sanitize_driver.py, memcheck on the small GPU parity tests. Logs land in gpurun_out/.
SAN=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 $SAN --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
timeout 900 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "skip_masks or per_request or scalar_tracks or unpacked or seek_integers" > gpurun_out/sanitizer_memcheck_pytest.log 2>&1
echo "memcheck pytest exit $?" >> gpurun_out/sanitizer_memcheck_pytest.log
tail -4 gpurun_out/sanitizer_memcheck_pytest.log
