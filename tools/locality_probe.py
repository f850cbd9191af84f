"""Development probe: how much of the C2 launch time is input-side memory latency? Re-runs the C2 request list with the clip
index folded onto the first M clips (M = 10000 is the real workload): the output traffic is unchanged, the compressed working
set shrinks from 257 MB (DRAM) to L2-resident to L1-resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import acl_b200 as ab
import bench

w = bench.make_workload("c2", 0, None)
ctx = ab.Context(0)
clipset = ctx.upload_packed(w["buffer"], w["offsets"], w["sizes"])
n = len(w["req_clip"])
options = ab.Options(output_layout=ab.LAYOUT_QVV40)
d_out = torch.empty(n * clipset.max_tracks * 40, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
for m in (10000, 2000, 444, 148, 12, 1):
    requests = ab.make_requests((w["req_clip"] % m).astype(np.uint32), w["req_time"])
    d_req = torch.from_numpy(requests.view(np.uint8)).cuda()
    for _ in range(3):
        ctx.decompress_tracks(clipset, d_req, n, options, d_out, stream)
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); ctx.decompress_tracks(clipset, d_req, n, options, d_out, stream); b.record(stream)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    print(f"clips folded onto {m:6d}: median {np.median(times):.4f} ms")
