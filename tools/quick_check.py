"""Ad-hoc GPU sanity check used during development: CUDA vs the C port on a few reference-compressed clips."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import acl_b200 as ab
from oracle import ref, port

M = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10]
ctx = ab.Context(0)
specs = {
    "C1": ref.TransformSpec(num_tracks=30, num_samples=60, seed=1000),
    "C2": ref.TransformSpec(),
    "mixed": ref.TransformSpec(num_tracks=57, num_samples=75, seed=7, rot_default_pct=10, rot_constant_pct=30, trans_default_pct=20, trans_constant_pct=40, scale_default_pct=60, scale_constant_pct=20, partial_activity_pct=30, noisy_pct=10),
    "single": ref.TransformSpec(num_tracks=33, num_samples=20, seed=9, rot_constant_pct=30, trans_constant_pct=40, scale_default_pct=50, scale_constant_pct=20),
    "strip": ref.TransformSpec(num_tracks=40, num_samples=120, seed=13, strip_proportion=0.4, trans_constant_pct=50, looping_content=1),
    "full": ref.TransformSpec(num_tracks=21, num_samples=50, seed=14, strip_trivial=0, rotation_format=ref.QUATF_FULL, translation_format=ref.VECTOR3F_FULL, scale_format=ref.VECTOR3F_FULL, scale_default_pct=50, rot_constant_pct=20),
}
for name, spec in specs.items():
    blob = ref.compress_transform(spec)
    cs = ctx.upload([blob], check_hash=True)
    dur = (spec.num_samples - 1) / spec.sample_rate
    times = np.concatenate([np.linspace(0, dur, 41), [-0.5, dur + 1, 0.3333]]).astype(np.float32)
    n = len(times)
    req = ab.make_requests(np.zeros(n, np.uint32), times)
    d_req = torch.from_numpy(req.view(np.uint8)).cuda()
    for kind in (0, 1, 3):
        s = port.settings_for_kind(kind)
        for rounding in (0, 1, 2, 3):
            opts = ab.Options(rounding_policy=rounding, normalization=s.c.normalization, per_track_rounding=s.c.per_track_rounding,
                              multiple_rotation_formats=s.c.multiple_rotation_formats)
            out = torch.zeros((n, cs.max_tracks, 12), dtype=torch.float32, device="cuda")
            ctx.decompress_tracks(cs, d_req, n, opts, out)
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            bad = 0
            for i, t in enumerate(times):
                exp = port.transform_decompress_tracks(blob, s, float(t), rounding)
                if not np.array_equal(exp[:, M].view(np.uint32), got[i][:, M].view(np.uint32)):
                    bad += 1
                    if bad == 1:
                        d = np.abs(exp[:, M] - got[i][:, M]); j = np.unravel_index(np.argmax(d), d.shape)
                        print("   first mismatch t", t, "maxdiff", d.max(), "bone", j[0], exp[j[0]], got[i][j[0]])
            print(name, "kind", kind, "round", rounding, "bad", bad, "/", n)
print("launches", ctx.launch_count)
