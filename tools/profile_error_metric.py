"""Profiling driver for the 8(f1) path: aclb200_calculate_compression_error over N reference-compressed C2-like clips, a few calls.
Used under ncu (`ncu --set full -k regex:object_space_kernel -c 1 python tools/profile_error_metric.py 1024 1`); prints the per clip agreement with the reference on a small sample."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import acl_b200 as ab
from oracle import ref

num_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
chunk_mb = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
spec = ref.TransformSpec(num_tracks=100, num_samples=60, seed=2000)
buffer, offsets, sizes = ref.compress_transform_batch(spec, num_clips, num_threads=ref.usable_threads())
raw, parents, shells = ref.sample_raw_transform_batch(spec, num_clips)
ctx = ab.Context(0)
ctx.set_error_chunk_bytes(chunk_mb << 20)
clipset = ctx.upload_packed(buffer, offsets, sizes)
jobs = np.zeros(num_clips, dtype=ab.ERROR_JOB_DTYPE)
jobs["clip"] = np.arange(num_clips)
jobs["num_samples"] = spec.num_samples
jobs["sample_rate"] = spec.sample_rate
jobs["duration"] = ref.finite_duration(spec.num_samples, spec.sample_rate)
jobs["num_tracks"] = spec.num_tracks
jobs["first_raw_pose"] = np.arange(num_clips, dtype=np.uint64) * spec.num_samples
d_raw = torch.from_numpy(raw.reshape(-1)).cuda()
d_parents = torch.from_numpy(parents.view(np.int32)).cuda()
d_shells = torch.from_numpy(shells).cuda()
d_errors = torch.zeros(num_clips * 4, dtype=torch.int32, device="cuda")
options = ab.Options(normalization=ab.NORMALIZE_ALWAYS, per_track_rounding=1, multiple_rotation_formats=1, default_modes=(ab.DEFAULT_CONSTANT,) * 3,
                     constant_defaults=[0, 0, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0])
for call in range(calls):
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    ctx.calculate_compression_error(clipset, jobs, d_raw, d_parents, d_shells, options, d_errors)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop)
    print(f"call {call}: {ms:.3f} ms, {num_clips * spec.num_samples * spec.num_tracks / ms / 1e6:.2f} G bone-poses measured/s")
errors = d_errors.cpu().numpy().view(ab.TRACK_ERROR_DTYPE)
sample = min(num_clips, 64)
blobs = [buffer[int(o):int(o) + int(s)] for o, s in zip(offsets[:sample], sizes[:sample])]
seconds, cpu = ref.bench_transform_error(spec, blobs, ref.usable_threads())
print("max |error - reference| over", sample, "clips:", float(np.max(np.abs(errors["error"][:sample] - cpu["error"]))),
      "same worst track:", float(np.mean(errors["index"][:sample] == cpu["index"])), "flags:", int(np.count_nonzero(errors["flags"])))
