"""GPU parity of the BENCHMARKED workloads, exhaustively: the exact request lists `bench.py` times (BASELINE.json configs
C2, C3, C4, C5 at full size, same seeds) are decoded in one launch through the C ABI and EVERY request is compared with the
unmodified reference (oracle/_ref: acl::decompression_context<benchmark settings>::seek + decompress_tracks into a
debug_track_writer style pose, oracle/ref_tool.cpp aclref_bench_transform / aclref_bench_scalar with an output buffer).
Mirrors what the reference's own validation walks (tools/acl_compressor/sources/validate_tracks.cpp:92-260,328-511).

Bar: ACLB200_MATH_EXACT bit-identical on every defined lane; ACLB200_MATH_FAST rotations <= 1e-5 absolute, translations and
scales bit-identical. Needs the compiled reference (it travels to the GPU box prebuilt); skipped without it.
"""
import numpy as np
import pytest

from tests import clips

pytestmark = pytest.mark.gpu

LANES = clips.DEFINED_LANES
FAST_MATH_TOLERANCE = 1e-5


@pytest.fixture(scope="module")
def env():
    import torch
    import acl_b200 as ab
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libaclref.so not built (needs /root/reference at build time)")
    ref.lib()
    return dict(torch=torch, ab=ab, ref=ref, ctx=ab.Context(0))


def _blobs(w):
    return [w["buffer"][int(o):int(o) + int(s)] for o, s in zip(w["offsets"], w["sizes"])]


def _compare_transform(env, w, clipset, slice_requests):
    """Decodes the whole request list in ONE launch per arithmetic mode (the launch bench.py times), then walks it slice by
    slice against the reference. Returns (requests compared, worst fast-math rotation error)."""
    torch, ab, ref, ctx = env["torch"], env["ab"], env["ref"], env["ctx"]
    n_req, tracks = len(w["req_clip"]), clipset.max_tracks
    requests = ab.make_requests(w["req_clip"], w["req_time"])
    d_requests = torch.from_numpy(requests.view(np.uint8)).cuda()
    d_exact = torch.full((n_req, tracks, 12), float("nan"), dtype=torch.float32, device="cuda")
    ctx.decompress_tracks(clipset, d_requests, n_req, ab.Options(output_layout=ab.LAYOUT_QVV48, math_mode=ab.MATH_EXACT), d_exact)
    d_fast = torch.full((n_req, tracks, 12), float("nan"), dtype=torch.float32, device="cuda")
    ctx.decompress_tracks(clipset, d_requests, n_req, ab.Options(output_layout=ab.LAYOUT_QVV48, math_mode=ab.MATH_FAST), d_fast)
    # the 40 byte layout the bench measures carries the same bits
    d_40 = torch.full((n_req, tracks, 10), float("nan"), dtype=torch.float32, device="cuda")
    ctx.decompress_tracks(clipset, d_requests, n_req, ab.Options(output_layout=ab.LAYOUT_QVV40, math_mode=ab.MATH_EXACT), d_40)
    torch.cuda.synchronize()
    assert torch.equal(d_40.view(torch.int32), d_exact[:, :, LANES].contiguous().view(torch.int32))
    del d_40

    blobs = _blobs(w)
    lanes = torch.tensor(LANES, device="cuda")
    vector_lanes = torch.tensor([4, 5, 6, 8, 9, 10], device="cuda")
    worst_fast = 0.0
    for begin in range(0, n_req, slice_requests):
        end = min(begin + slice_requests, n_req)
        want = ref.decode_requests(blobs, w["req_clip"][begin:end], w["req_time"][begin:end], tracks)
        d_want = torch.from_numpy(want).cuda()
        got = d_exact[begin:end]
        same = torch.equal(got.index_select(2, lanes).view(torch.int32), d_want.index_select(2, lanes).view(torch.int32))
        if not same:
            diff = (got.index_select(2, lanes).view(torch.int32) != d_want.index_select(2, lanes).view(torch.int32)).nonzero()
            r, bone, lane = (int(v) for v in diff[0])
            raise AssertionError(f"request {begin + r} (clip {int(w['req_clip'][begin + r])}, t={float(w['req_time'][begin + r])}) bone {bone} lane {LANES[lane]}: "
                                 f"got {got[r, bone].tolist()} want {d_want[r, bone].tolist()} ({diff.shape[0]} differing values in this slice)")
        fast = d_fast[begin:end]
        worst_fast = max(worst_fast, float((fast[:, :, :4] - d_want[:, :, :4]).abs().max()))
        assert torch.equal(fast.index_select(2, vector_lanes).view(torch.int32), d_want.index_select(2, vector_lanes).view(torch.int32))
        del d_want
    assert worst_fast <= FAST_MATH_TOLERANCE, worst_fast
    return n_req, worst_fast


@pytest.mark.parametrize("name, slice_requests", [("c2", 60000), ("c3", 6000), ("c5", 125000)])
def test_bench_workload_every_request_vs_reference(env, name, slice_requests):
    import bench
    w = bench.make_workload(name, 0, None)
    assert w["distinct"], "the reference compressor is needed for the bench clips"
    clipset = env["ctx"].upload_packed(w["buffer"], w["offsets"], w["sizes"], check_hash=True)
    assert clipset.max_tracks == w["num_tracks"]
    compared, worst_fast = _compare_transform(env, w, clipset, slice_requests)
    assert compared == len(w["req_clip"])
    print(f"{name}: {compared} requests x {w['num_tracks']} bones bit-identical to the reference; fast math worst rotation error {worst_fast:.2e}")
    clipset.release()


def test_bench_workload_c4_every_request_vs_reference(env):
    """C4: scalar float1f 4096 tracks x 1024 samples replicated x64, the 65 536 requests bench.py times."""
    import bench
    torch, ab, ref, ctx = env["torch"], env["ab"], env["ref"], env["ctx"]
    w = bench.make_workload("c4", 0, None)
    clipset = ctx.upload_packed(w["buffer"], w["offsets"], w["sizes"], check_hash=True)
    n_req, tracks = len(w["req_clip"]), clipset.max_tracks
    assert tracks == 4096 and clipset.components == 1
    requests = ab.make_requests(w["req_clip"], w["req_time"])
    d_requests = torch.from_numpy(requests.view(np.uint8)).cuda()
    d_out = torch.full((n_req, tracks), float("nan"), dtype=torch.float32, device="cuda")
    ctx.scalar_decompress_tracks(clipset, d_requests, n_req, ab.Options(), d_out)
    torch.cuda.synchronize()
    blobs = _blobs(w)
    for begin in range(0, n_req, 4096):
        end = min(begin + 4096, n_req)
        want = ref.decode_requests(blobs, w["req_clip"][begin:end], w["req_time"][begin:end], tracks, scalar=True)[:, :, 0]
        d_want = torch.from_numpy(np.ascontiguousarray(want)).cuda()
        assert torch.equal(d_out[begin:end].view(torch.int32), d_want.view(torch.int32)), f"requests {begin}..{end}"
    clipset.release()
