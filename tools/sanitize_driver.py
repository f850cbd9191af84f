"""Small workload for compute-sanitizer (tools/sanitize.sh): every kernel of the library once or twice, checked against the oracle --
chained playback through the pipeline kernel (groups, tail crossing, base row reuse), ragged random requests, per track rounding,
skipped defaults (plain kernels), decompress_track, the chained scalar kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import acl_b200 as ab
from oracle import port
from tests import clips

L = clips.DEFINED_LANES
ctx = ab.Context(0)
names = ["c2_100bones", "mixed_scale", "stripped_loop", "c1_30bones", "ragged_17", "full_formats"]
blobs = [clips.load_blob(n) for n in names]
cs = ctx.upload(blobs, check_hash=True)
settings = port.settings_for_kind(0)
# sequential playback of every clip (chains + segment crossings), then random requests
req_clip, req_time = [], []
for c, n in enumerate(names):
    spec = clips.TRANSFORM_SPECS[n]
    for s in range(spec.num_samples):
        req_clip.append(c); req_time.append((s + 0.37) / spec.sample_rate)
rng = np.random.default_rng(0)
for _ in range(200):
    c = int(rng.integers(0, len(names))); req_clip.append(c); req_time.append(float(rng.uniform(-0.1, 2.5)))
req_clip = np.array(req_clip, np.uint32); req_time = np.array(req_time, np.float32)
req = ab.make_requests(req_clip, req_time)
d_req = torch.from_numpy(req.view(np.uint8)).cuda()
bad = 0
for layout, width in ((ab.LAYOUT_QVV48, 12), (ab.LAYOUT_QVV40, 10)):
    for math in (ab.MATH_EXACT, ab.MATH_FAST):
        out = torch.zeros((len(req), cs.max_tracks, width), dtype=torch.float32, device="cuda")
        ctx.decompress_tracks(cs, d_req, len(req), ab.Options(output_layout=layout, math_mode=math), out)
        torch.cuda.synchronize()
        if math == ab.MATH_EXACT and layout == ab.LAYOUT_QVV48:
            got = out.cpu().numpy()
            for i in range(0, len(req), 3):
                want = port.transform_decompress_tracks(blobs[req_clip[i]], settings, float(req_time[i]))
                bad += not clips.bit_equal(got[i, :want.shape[0]][:, L], want[:, L])
# per track rounding + always normalisation (generic consumers), skipped defaults (plain kernels), decompress_track
debug = port.settings_for_kind(1)
policies = torch.from_numpy(rng.integers(0, 4, cs.max_tracks).astype(np.uint8)).cuda()
out = torch.zeros((len(req), cs.max_tracks, 12), dtype=torch.float32, device="cuda")
ctx.decompress_tracks(cs, d_req, len(req), ab.Options(normalization=ab.NORMALIZE_ALWAYS, per_track_rounding=1, rounding_policy=ab.ROUND_PER_TRACK,
                                                      d_per_track_rounding=policies.data_ptr(), multiple_rotation_formats=1), out)
ctx.decompress_tracks(cs, d_req, len(req), ab.Options(default_modes=(ab.DEFAULT_SKIPPED, ab.DEFAULT_SKIPPED, ab.DEFAULT_SKIPPED), skip_mask=ab.SKIP_SCALE), out)
tracks = torch.from_numpy(rng.integers(0, 17, len(req)).astype(np.uint32)).cuda()
one = torch.zeros((len(req), 12), dtype=torch.float32, device="cuda")
ctx.decompress_track(cs, d_req, tracks, len(req), ab.Options(), one)
torch.cuda.synchronize()
# scalar clips
for name in ("float1", "float3", "vector4", "float1_c4_small"):
    blob = clips.load_blob(name); spec = clips.SCALAR_SPECS[name]
    scs = ctx.upload([blob])
    times = np.concatenate([(np.arange(spec.num_samples) + 0.4) / spec.sample_rate, rng.uniform(-0.1, 3.0, 40)]).astype(np.float32)
    sreq = ab.make_requests(np.zeros(len(times), np.uint32), times)
    d_sreq = torch.from_numpy(sreq.view(np.uint8)).cuda()
    sout = torch.zeros((len(times), scs.max_tracks, scs.components), dtype=torch.float32, device="cuda")
    ctx.scalar_decompress_tracks(scs, d_sreq, len(times), ab.Options(), sout)
    torch.cuda.synchronize()
    got = sout.cpu().numpy()
    scs.release()
print("driver done, mismatches vs oracle:", bad)
sys.exit(1 if bad else 0)
