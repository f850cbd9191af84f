"""GPU parity tests: the CUDA path, called through the C ABI (acl_b200.api -> libaclb200.so), against
  * the oracle (oracle/acl_oracle.c, itself pinned to the reference by tests/test_oracle_vs_reference.py), bit for bit, and
  * the golden vectors the reference itself produced (tests/golden/), bit for bit.
Float gate: decompress_tracks is expected bit-identical (0 ulp); decompress_track rotations are gated at 1e-5 absolute per
component because the reference itself normalises them with a CPU-dependent rsqrt estimate."""
import numpy as np
import pytest

from tests import clips

pytestmark = pytest.mark.gpu

LANES = clips.DEFINED_LANES
SINGLE_TRACK_TOLERANCE = 1e-5


@pytest.fixture(scope="module")
def gpu():
    import torch
    import acl_b200 as ab
    from oracle import port
    port.lib()
    return dict(torch=torch, ab=ab, port=port, ctx=ab.Context(0))


def _to_device(gpu, array):
    return gpu["torch"].from_numpy(np.ascontiguousarray(array).view(np.uint8).reshape(-1)).cuda()


def _kinds_for(spec):
    from oracle import ref
    is_full = spec.rotation_format == ref.QUATF_FULL
    default_ok = spec.rotation_format == ref.QUATF_DROP_W_VARIABLE and spec.translation_format == ref.VECTOR3F_VARIABLE \
        and spec.scale_format == ref.VECTOR3F_VARIABLE
    return [1, 3, 4] + ([0] if default_ok else []) + ([5] if is_full else [])


def _options(gpu, settings, **kw):
    ab = gpu["ab"]
    s = settings.c
    return ab.Options(normalization=s.normalization, per_track_rounding=s.per_track_rounding, wrapping=s.wrapping,
                      clamp_sample_time=s.clamp_sample_time, multiple_rotation_formats=s.multiple_rotation_formats,
                      default_modes=(s.default_rotation_mode, s.default_translation_mode, s.default_scale_mode),
                      constant_defaults=list(s.constant_defaults), **kw)


def _decode(gpu, clipset, req_clip, req_time, options, prefill=None):
    """Runs aclb200_decompress_tracks and returns float32 [n, max_tracks, 12] (QVV48) or [n, max_tracks, 10] (QVV40)."""
    torch, ab, ctx = gpu["torch"], gpu["ab"], gpu["ctx"]
    n = len(req_clip)
    width = 12 if options.output_layout == ab.LAYOUT_QVV48 else 10
    requests = ab.make_requests(req_clip, req_time)
    d_requests = _to_device(gpu, requests)
    if prefill is None:
        d_out = torch.full((n, clipset.max_tracks, width), float("nan"), dtype=torch.float32, device="cuda")
    else:
        d_out = torch.from_numpy(prefill).cuda()
    ctx.decompress_tracks(clipset, d_requests, n, options, d_out)
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


@pytest.mark.parametrize("name", list(clips.TRANSFORM_SPECS))
def test_decompress_tracks_bit_exact_vs_oracle(gpu, name):
    port, ab, ctx = gpu["port"], gpu["ab"], gpu["ctx"]
    spec = clips.TRANSFORM_SPECS[name]
    blob = clips.load_blob(name)
    clipset = ctx.upload([blob], check_hash=True)
    times = clips.sample_times(spec)
    zeros = np.zeros(len(times), dtype=np.uint32)
    n = clipset.max_tracks
    for kind in _kinds_for(spec):
        settings = port.settings_for_kind(kind)
        for looping in (ab.LOOP_AS_COMPRESSED, ab.LOOP_CLAMP, ab.LOOP_WRAP):
            for rounding in (0, 1, 2, 3):
                got = _decode(gpu, clipset, zeros, times, _options(gpu, settings, rounding_policy=rounding, looping_policy=looping))
                for i, t in enumerate(times):
                    want = port.transform_decompress_tracks(blob, settings, float(t), rounding, looping)
                    assert clips.bit_equal(got[i, :n][:, LANES], want[:, LANES]), (name, kind, looping, rounding, float(t))
    clipset.release()


@pytest.mark.parametrize("name", list(clips.TRANSFORM_SPECS))
def test_decompress_tracks_bit_exact_vs_reference_goldens(gpu, name):
    port, ab, ctx = gpu["port"], gpu["ab"], gpu["ctx"]
    blob = clips.load_blob(name)
    g = np.load(clips.golden_path(name, "golden.npz"))
    clipset = ctx.upload([blob])
    times = g["times"]
    zeros = np.zeros(len(times), dtype=np.uint32)
    for ci, (kind, rounding) in enumerate(g["combos"]):
        settings = port.settings_for_kind(int(kind))
        for layout in (ab.LAYOUT_QVV48, ab.LAYOUT_QVV40):
            got = _decode(gpu, clipset, zeros, times, _options(gpu, settings, rounding_policy=int(rounding), output_layout=layout))
            got = got[:, :, LANES] if layout == ab.LAYOUT_QVV48 else got
            assert clips.bit_equal(got, g["poses"][ci]), (name, kind, rounding, layout)
    clipset.release()


@pytest.mark.parametrize("name", ["mixed_scale", "ragged_17", "full_formats", "stripped_loop"])
def test_default_modes_and_per_track_rounding(gpu, name):
    """track_writer default sub-track modes (skipped / constant / variable / legacy) and per track rounding
    (validate_tracks.cpp:189-229 checks the same equivalences on the CPU)."""
    port, ab, ctx, torch = gpu["port"], gpu["ab"], gpu["ctx"], gpu["torch"]
    spec = clips.TRANSFORM_SPECS[name]
    blob = clips.load_blob(name)
    clipset = ctx.upload([blob])
    n = clipset.max_tracks
    rng = np.random.default_rng(99)
    policies = rng.integers(0, 4, size=n).astype(np.uint8)
    constant_defaults = rng.normal(size=12).astype(np.float32)
    variable_defaults = rng.normal(size=(n, 12)).astype(np.float32)
    d_policies = _to_device(gpu, policies)
    d_variable = _to_device(gpu, variable_defaults)
    times = clips.sample_times(spec)[::2]
    zeros = np.zeros(len(times), dtype=np.uint32)
    for writer in range(4):
        settings = port.settings_for_kind(1, default_modes=port.writer_modes(writer), constant_defaults=constant_defaults,
                                          variable_defaults=variable_defaults, per_track_policies=policies)
        for rounding in (0, 1, 2, 3, 4):
            prefill = rng.normal(size=(len(times), n, 12)).astype(np.float32)
            options = _options(gpu, settings, rounding_policy=rounding)
            options.d_variable_defaults = d_variable.data_ptr()
            options.d_per_track_rounding = d_policies.data_ptr()
            got = _decode(gpu, clipset, zeros, times, options, prefill=prefill.copy())
            for i, t in enumerate(times):
                want = port.transform_decompress_tracks(blob, settings, float(t), rounding, out=prefill[i].copy())
                assert clips.bit_equal(got[i][:, LANES], want[:, LANES]), (name, writer, rounding, float(t))
    clipset.release()


@pytest.mark.parametrize("name", list(clips.TRANSFORM_SPECS))
def test_seek_integers_bit_exact(gpu, name):
    """key frames, segment choice, bit offsets and interpolation alpha of seek() (integer stage, bit-exact)."""
    port, ab, ctx, torch = gpu["port"], gpu["ab"], gpu["ctx"], gpu["torch"]
    from acl_b200 import api
    spec = clips.TRANSFORM_SPECS[name]
    blob = clips.load_blob(name)
    clipset = ctx.upload([blob])
    duration = max(spec.num_samples - 1, 0) / spec.sample_rate
    times = np.concatenate([clips.sample_times(spec), np.linspace(-0.1, duration + 0.2, 61).astype(np.float32)])
    requests = ab.make_requests(np.zeros(len(times), dtype=np.uint32), times)
    d_requests = _to_device(gpu, requests)
    settings = port.settings_for_kind(1)
    for looping in (0, 1, 2):
        for rounding in (0, 1, 2, 3, 4):
            d_out = torch.zeros(len(times) * api.SEEK_STATE_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            ctx.debug_seek(clipset, d_requests, len(times), _options(gpu, settings, rounding_policy=rounding, looping_policy=looping), d_out)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy().view(api.SEEK_STATE_DTYPE)
            for i, t in enumerate(times):
                st = port.transform_seek(blob, settings, float(t), rounding, looping)
                if st.sample_time < 0:
                    assert got[i]["sample_time"] < 0
                    continue
                key = (name, looping, rounding, float(t))
                assert np.float32(st.sample_time).view(np.uint32) == got[i]["sample_time"].view(np.uint32), key
                assert np.float32(st.interpolation_alpha).view(np.uint32) == got[i]["interpolation_alpha"].view(np.uint32), key
                assert list(st.key_frame_bit_offsets) == list(got[i]["key_frame_bit_offsets"]), key
                assert list(st.segment_indices) == list(got[i]["segment_indices"]), key
                assert list(st.animated_offsets) == list(got[i]["animated_offsets"]), key
                assert list(st.format_offsets) == list(got[i]["format_offsets"]), key
                assert list(st.range_offsets) == list(got[i]["range_offsets"]), key
                assert st.uses_single_segment == got[i]["uses_single_segment"] and st.looping_policy == got[i]["looping_policy"], key
    clipset.release()


@pytest.mark.parametrize("name", ["c2_100bones", "mixed_scale", "noisy_raw", "full_formats", "drop_w_full", "mixed_formats", "stripped_loop", "single_segment"])
def test_unpacked_integers_bit_exact(gpu, name):
    """The quantised integers pulled out of the variable bit rate stream (format decode, bit-exact)."""
    port, ab, ctx, torch = gpu["port"], gpu["ab"], gpu["ctx"], gpu["torch"]
    spec = clips.TRANSFORM_SPECS[name]
    blob = clips.load_blob(name)
    clipset = ctx.upload([blob])
    times = clips.sample_times(spec)
    requests = ab.make_requests(np.zeros(len(times), dtype=np.uint32), times)
    d_requests = _to_device(gpu, requests)
    settings = port.settings_for_kind(1)
    header = blob[32:84].view(np.uint32)
    has_scale = int(blob[28:32].view(np.uint32)[0]) & 1
    total = int(header[2]) + int(header[3]) + (int(header[4]) if has_scale else 0)
    if total == 0:
        pytest.skip("no animated sub-track")
    raw_marker = 31
    for which in (0, 1):
        d_out = torch.zeros((len(times), total, 4), dtype=torch.int32, device="cuda")
        ctx.debug_unpack(clipset, d_requests, len(times), _options(gpu, settings), which, total, d_out)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint32)
        for i, t in enumerate(times):
            st = port.transform_seek(blob, settings, float(t))
            want = port.transform_key_frame_ints(blob, st, which)
            assert np.array_equal(got[i][:, :3], want[:, :3]), (name, which, float(t))
            # the entry code: stored bit count, with the raw marker (31) reported as 32 | 0x80 and full formats likewise
            stored = want[:, 3]
            code = np.where((stored == raw_marker) | (stored == 0xFFFFFFFF), 32 | 0x80, stored)
            assert np.array_equal(got[i][:, 3], code), (name, which, float(t))
    clipset.release()


@pytest.mark.parametrize("name", ["mixed_scale", "single_segment", "full_formats", "stripped_loop", "c2_100bones"])
def test_decompress_track_vs_oracle(gpu, name):
    port, ab, ctx, torch = gpu["port"], gpu["ab"], gpu["ctx"], gpu["torch"]
    spec = clips.TRANSFORM_SPECS[name]
    blob = clips.load_blob(name)
    clipset = ctx.upload([blob])
    n = clipset.max_tracks
    times = clips.sample_times(spec)[::3]
    bones = np.arange(n, dtype=np.uint32)
    req_time = np.repeat(times, n).astype(np.float32)
    req_bone = np.tile(bones, len(times)).astype(np.uint32)
    requests = ab.make_requests(np.zeros(len(req_time), dtype=np.uint32), req_time)
    d_requests = _to_device(gpu, requests)
    d_bones = _to_device(gpu, req_bone)
    worst = 0.0
    for kind in _kinds_for(spec):
        settings = port.settings_for_kind(kind)
        for rounding in (0, 1, 2, 3):
            d_out = torch.full((len(req_time), 12), float("nan"), dtype=torch.float32, device="cuda")
            ctx.decompress_track(clipset, d_requests, d_bones, len(req_time), _options(gpu, settings, rounding_policy=rounding), d_out)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy()
            for ti, t in enumerate(times):
                for bone in range(n):
                    want = port.transform_decompress_track(blob, settings, float(t), bone, rounding)[bone]
                    row = got[ti * n + bone]
                    # translation / scale use the same IEEE operations: bit-exact
                    assert clips.bit_equal(row[[4, 5, 6, 8, 9, 10]], want[[4, 5, 6, 8, 9, 10]]), (name, kind, rounding, float(t), bone)
                    diff = float(np.abs(row[:4] - want[:4]).max())
                    worst = max(worst, diff)
                    assert diff <= SINGLE_TRACK_TOLERANCE, (name, kind, rounding, float(t), bone, diff)
    clipset.release()
    print(f"{name}: worst decompress_track rotation difference {worst:.3e}")


@pytest.mark.parametrize("name", list(clips.SCALAR_SPECS))
def test_scalar_tracks_bit_exact(gpu, name):
    port, ab, ctx, torch = gpu["port"], gpu["ab"], gpu["ctx"], gpu["torch"]
    spec = clips.SCALAR_SPECS[name]
    blob = clips.load_blob(name)
    clipset = ctx.upload([blob], check_hash=True)
    n, nc = clipset.max_tracks, clipset.components
    times = clips.sample_times(spec)
    requests = ab.make_requests(np.zeros(len(times), dtype=np.uint32), times)
    d_requests = _to_device(gpu, requests)
    policies = np.random.default_rng(5).integers(0, 4, size=n).astype(np.uint8)
    d_policies = _to_device(gpu, policies)
    g = np.load(clips.golden_path(name, "golden.npz"))
    for per_track in (False, True):
        settings = port.SettingsBuilder(per_track_rounding=per_track, per_track_policies=policies if per_track else None)
        for rounding in ((0, 1, 2, 3, 4) if per_track else (0, 1, 2, 3)):
            for looping in (0, 1, 2):
                options = ab.Options(rounding_policy=rounding, looping_policy=looping, per_track_rounding=int(per_track))
                if per_track:
                    options.d_per_track_rounding = d_policies.data_ptr()
                d_out = torch.full((len(times), n, nc), float("nan"), dtype=torch.float32, device="cuda")
                ctx.scalar_decompress_tracks(clipset, d_requests, len(times), options, d_out)
                torch.cuda.synchronize()
                got = d_out.cpu().numpy()
                for i, t in enumerate(times):
                    want = port.scalar_decompress(blob, settings, float(t), rounding, looping)[:, :nc]
                    assert clips.bit_equal(got[i], want), (name, per_track, rounding, looping, float(t))
                # single track flavour
                tracks = np.array([0, n // 2, n - 1], dtype=np.uint32)
                req1 = ab.make_requests(np.zeros(len(times) * 3, dtype=np.uint32), np.repeat(times, 3))
                d_req1 = _to_device(gpu, req1)
                d_tracks = _to_device(gpu, np.tile(tracks, len(times)))
                d_out1 = torch.full((len(times) * 3, nc), float("nan"), dtype=torch.float32, device="cuda")
                ctx.scalar_decompress_track(clipset, d_req1, d_tracks, len(times) * 3, options, d_out1)
                torch.cuda.synchronize()
                got1 = d_out1.cpu().numpy()
                for i, t in enumerate(times):
                    for j, track in enumerate(tracks):
                        want1 = port.scalar_decompress(blob, settings, float(t), rounding, looping, track=int(track))[int(track), :nc]
                        assert clips.bit_equal(got1[i * 3 + j], want1), (name, per_track, rounding, looping, float(t), int(track))
    # and straight against the reference's golden outputs
    for rounding in range(4):
        for looping in range(3):
            gt = g["times"]
            reqg = ab.make_requests(np.zeros(len(gt), dtype=np.uint32), gt)
            d_outg = torch.zeros((len(gt), n, nc), dtype=torch.float32, device="cuda")
            ctx.scalar_decompress_tracks(clipset, _to_device(gpu, reqg), len(gt), ab.Options(rounding_policy=rounding, looping_policy=looping), d_outg)
            torch.cuda.synchronize()
            assert clips.bit_equal(d_outg.cpu().numpy(), g["values"][rounding, looping]), (name, rounding, looping)
    clipset.release()


def test_ragged_clip_set_and_invalid_requests(gpu):
    """Many clips of different skeleton sizes in one clip set, shuffled requests, out-of-range clip indices."""
    port, ab, ctx = gpu["port"], gpu["ab"], gpu["ctx"]
    names = [n for n, s in clips.TRANSFORM_SPECS.items() if s.rotation_format == 3 and s.translation_format == 1 and s.scale_format == 1]
    blobs = [clips.load_blob(n) for n in names]
    clipset = ctx.upload(blobs, check_hash=True)
    assert clipset.num_clips == len(blobs) and clipset.max_tracks == 540 and clipset.min_tracks == 1
    rng = np.random.default_rng(7)
    count = 700
    req_clip = rng.integers(0, len(blobs), size=count).astype(np.uint32)
    req_time = rng.uniform(-0.2, 4.5, size=count).astype(np.float32)
    req_clip[::97] = len(blobs) + 5             # invalid clip index: nothing may be written
    settings = port.settings_for_kind(0)
    sentinel = np.full((count, clipset.max_tracks, 12), 12345.0, dtype=np.float32)
    got = _decode(gpu, clipset, req_clip, req_time, _options(gpu, settings), prefill=sentinel.copy())
    for i in range(count):
        if req_clip[i] >= len(blobs):
            assert np.all(got[i] == 12345.0)
            continue
        blob = blobs[req_clip[i]]
        want = port.transform_decompress_tracks(blob, settings, float(req_time[i]))
        n = want.shape[0]
        assert clips.bit_equal(got[i, :n][:, LANES], want[:, LANES]), (i, names[req_clip[i]], float(req_time[i]))
        assert np.all(got[i, n:] == 12345.0), "rows past the clip's bone count must stay untouched"
    clipset.release()


def test_host_buffer_api_matches_device_api(gpu):
    port, ab, ctx = gpu["port"], gpu["ab"], gpu["ctx"]
    blobs = [clips.load_blob(n) for n in ("c1_30bones", "c5_30x32", "looping")]
    clipset = ctx.upload(blobs)
    rng = np.random.default_rng(3)
    count = 333
    req_clip = rng.integers(0, 3, size=count).astype(np.uint32)
    req_time = rng.uniform(0, 2.0, size=count).astype(np.float32)
    for layout, width in ((ab.LAYOUT_QVV48, 12), (ab.LAYOUT_QVV40, 10)):
        options = ab.Options(output_layout=layout)
        device = _decode(gpu, clipset, req_clip, req_time, options, prefill=np.zeros((count, clipset.max_tracks, width), dtype=np.float32))
        host = np.full((count, clipset.max_tracks, width), 7.0, dtype=np.float32)     # rows nobody writes come back as zero
        ctx.decompress_tracks_host(clipset, ab.make_requests(req_clip, req_time), options, host)
        assert np.array_equal(host.view(np.uint32), device.view(np.uint32))
    clipset.release()


def test_upload_rejects_what_initialize_rejects(gpu):
    """decompression_context::initialize() returns false for these (decompress.impl.h:66-83)."""
    ab, ctx = gpu["ab"], gpu["ctx"]
    from oracle import ref
    good = clips.load_blob("c1_30bones")

    def status_of(blob, check_hash=False):
        try:
            ctx.upload([blob], check_hash=check_hash).release()
            return 0
        except ab.AclB200Error as e:
            return e.status

    assert status_of(good, True) == 0
    bad = good.copy(); bad[8] ^= 0xFF
    assert status_of(ref.aligned_blob(bad)) == 2            # tag
    bad = good.copy(); bad[12] = 3
    assert status_of(ref.aligned_blob(bad)) == 2            # version
    bad = good.copy(); bad[14] = 1
    assert status_of(ref.aligned_blob(bad)) == 2            # algorithm
    bad = good.copy(); bad[300] ^= 1
    assert status_of(ref.aligned_blob(bad), True) == 2      # hash
    assert status_of(ref.aligned_blob(bad), False) == 0
    bad = good.copy(); bad[29] |= 1
    assert status_of(ref.aligned_blob(bad)) == 2            # database flag without a database header: corrupt
    assert status_of(ref.aligned_blob(good[:200].copy())) == 2      # truncated
    with pytest.raises(ab.AclB200Error) as err:
        ctx.upload([good, clips.load_blob("float1")])
    assert err.value.status == 3                            # mixed track types


@pytest.mark.parametrize("golden_name", ["database_c1_30bones", "database_mixed_scale"])
def test_database_clips_decode_from_their_resident_key_frames(gpu, golden_name):
    """SURVEY 8(f2), first step: clips bound to a streaming database (acl::build_database moved their movable key frames out) are accepted
    and decoded from the key frames that stay in the clip, like decompression_context<settings with database support>::initialize(tracks)
    with no database bound (decompress.impl.h:67-83): bit for bit against the reference's own output (golden) and the oracle, in the same
    launch as an ordinary clip, through the pipeline kernel and through decompress_track."""
    torch, ab, ctx, port = gpu["torch"], gpu["ab"], gpu["ctx"], gpu["port"]
    blob = clips.load_blob(golden_name)
    other = clips.load_blob("ragged_17")
    g = np.load(clips.golden_path(golden_name, "golden.npz"))
    clipset = ctx.upload([other, blob], check_hash=True)
    n = port.num_tracks_of(blob)
    times = g["times"]
    settings = port.settings_for_kind(1)
    req_clip = np.concatenate([np.ones(len(times), np.uint32), np.zeros(4, np.uint32)])
    req_time = np.concatenate([times, np.array([0.0, 0.4, 0.9, 1.3], np.float32)])
    for rounding in range(4):
        got = _decode(gpu, clipset, req_clip, req_time, _options(gpu, settings, rounding_policy=rounding))
        for i, t in enumerate(times):
            assert clips.bit_equal(got[i, :n][:, LANES], g["poses"][rounding, i]), (golden_name, rounding, float(t))
        for looping in (ab.LOOP_CLAMP, ab.LOOP_WRAP):
            got = _decode(gpu, clipset, req_clip, req_time, _options(gpu, settings, rounding_policy=rounding, looping_policy=looping))
            for i, t in enumerate(times):
                want = port.transform_decompress_tracks(blob, settings, float(t), rounding, looping)
                assert clips.bit_equal(got[i, :n][:, LANES], want[:, LANES]), (golden_name, rounding, looping, float(t))
    # decompress_track: vectors bit-exact, rotations within the single track tolerance
    bones = np.array([0, n // 2, n - 1], dtype=np.uint32)
    for bone in bones:
        requests = ab.make_requests(np.ones(len(times), np.uint32), times)
        d_out = torch.zeros((len(times), 12), dtype=torch.float32, device="cuda")
        ctx.decompress_track(clipset, _to_device(gpu, requests), _to_device(gpu, np.full(len(times), bone, np.uint32)), len(times), _options(gpu, settings), d_out)
        torch.cuda.synchronize()
        single = d_out.cpu().numpy()
        for i, t in enumerate(times):
            want = port.transform_decompress_tracks(blob, settings, float(t), 0)[bone]
            assert np.max(np.abs(single[i, :4] - want[:4])) <= SINGLE_TRACK_TOLERANCE, (golden_name, bone, float(t))
            assert clips.bit_equal(single[i, [4, 5, 6, 8, 9, 10]], want[[4, 5, 6, 8, 9, 10]]), (golden_name, bone, float(t))
    clipset.release()


FAST_MATH_TOLERANCE = 1e-5      # BASELINE.json north star: "within 1e-5 on the float QVV components"


@pytest.mark.parametrize("name", ["c1_30bones", "c2_100bones", "c5_30x32", "mixed_scale", "looping", "stripped_loop", "noisy_raw", "half_turn", "paragon_like"])
def test_fast_math_within_tolerance(gpu, name):
    """ACLB200_MATH_FAST (hardware sqrt / rsqrt, fused multiply-adds after the W reconstruction input): rotations within 1e-5 absolute
    of the reference on every component -- including `half_turn`, whose W crosses 0 -- translations and scales still bit-exact."""
    port, ab, ctx = gpu["port"], gpu["ab"], gpu["ctx"]
    spec = clips.TRANSFORM_SPECS[name]
    blob = clips.load_blob(name)
    clipset = ctx.upload([blob], check_hash=True)
    duration = (spec.num_samples - 1) / spec.sample_rate
    times = np.concatenate([clips.sample_times(spec), np.linspace(0.0, duration, 97, dtype=np.float32)])
    zeros = np.zeros(len(times), dtype=np.uint32)
    n = clipset.max_tracks
    worst = 0.0
    for kind in (0, 3):     # default settings (lerp_only) and never-normalise: the two policies the fast arithmetic serves
        settings = port.settings_for_kind(kind)
        for rounding in (0, 3):
            for layout in (ab.LAYOUT_QVV48, ab.LAYOUT_QVV40):
                got = _decode(gpu, clipset, zeros, times, _options(gpu, settings, rounding_policy=rounding, math_mode=ab.MATH_FAST, output_layout=layout))
                if layout == ab.LAYOUT_QVV40:
                    got = got[:, :, :10]
                for i, t in enumerate(times):
                    want = port.transform_decompress_tracks(blob, settings, float(t), rounding)
                    want = want[:, LANES] if layout == ab.LAYOUT_QVV40 else want
                    rot_error = float(np.abs(got[i, :n, :4] - want[:, :4]).max())
                    worst = max(worst, rot_error)
                    assert rot_error <= FAST_MATH_TOLERANCE, (name, kind, rounding, float(t), rot_error)
                    vec = slice(4, 10) if layout == ab.LAYOUT_QVV40 else [4, 5, 6, 8, 9, 10]
                    assert clips.bit_equal(got[i, :n][:, vec], want[:, vec]), (name, kind, rounding, float(t))
    assert worst < 2e-6, worst      # what the approximations actually cost on unit quaternions
    clipset.release()


@pytest.fixture(scope="module")
def c2_full(gpu):
    """BASELINE.json configs[1] at full size: 10 000 clips x 100 bones x 60 samples, 600 000 requests. Clips come from the reference
    compressor when oracle/_ref is present, else the committed golden clip is replicated."""
    import bench
    w = bench.make_workload("c2", 0, None)
    clipset = gpu["ctx"].upload_packed(w["buffer"], w["offsets"], w["sizes"])
    return w, clipset


def test_full_size_c2_properties(gpu, c2_full):
    """Size-independent properties at BASELINE's full size + bit-exact spot checks of 3000 random requests against the oracle."""
    port, ab, ctx, torch = gpu["port"], gpu["ab"], gpu["ctx"], gpu["torch"]
    w, clipset = c2_full
    n_req, bones = len(w["req_clip"]), w["num_tracks"]
    requests = ab.make_requests(w["req_clip"], w["req_time"])
    d_requests = _to_device(gpu, requests)
    options = ab.Options(output_layout=ab.LAYOUT_QVV40)
    d_out = torch.zeros((n_req, bones, 10), dtype=torch.float32, device="cuda")
    ctx.decompress_tracks(clipset, d_requests, n_req, options, d_out)
    torch.cuda.synchronize()

    # (1) every rotation is a unit quaternion, everything is finite
    assert bool(torch.isfinite(d_out).all())
    norms = d_out[:, :, :4].square().sum(dim=-1)
    assert float((norms - 1.0).abs().max()) < 1e-5

    # (2) idempotence / determinism: a second launch reproduces every bit
    d_out2 = torch.zeros_like(d_out)
    ctx.decompress_tracks(clipset, d_requests, n_req, options, d_out2)
    torch.cuda.synchronize()
    assert torch.equal(d_out.view(torch.int32), d_out2.view(torch.int32))

    # (3) permutation equivariance: shuffled requests give the shuffled poses (no cross-request state)
    perm = np.random.default_rng(1).permutation(n_req)
    d_req_perm = _to_device(gpu, requests[perm])
    ctx.decompress_tracks(clipset, d_req_perm, n_req, options, d_out2)
    torch.cuda.synchronize()
    d_perm = torch.from_numpy(perm).cuda()
    assert torch.equal(d_out2.view(torch.int32), d_out[d_perm].view(torch.int32))
    del d_out2

    # (4) the 48 byte layout carries the same values
    d_out48 = torch.zeros((n_req, bones, 12), dtype=torch.float32, device="cuda")
    ctx.decompress_tracks(clipset, d_requests, n_req, ab.Options(output_layout=ab.LAYOUT_QVV48), d_out48)
    torch.cuda.synchronize()
    assert torch.equal(d_out48[:, :, LANES].contiguous().view(torch.int32), d_out.view(torch.int32))
    del d_out48

    # (4b) fast arithmetic: every component of all 60 M bone-poses within 1e-5 of the bit-exact result, vectors identical
    d_fast = torch.zeros_like(d_out)
    ctx.decompress_tracks(clipset, d_requests, n_req, ab.Options(output_layout=ab.LAYOUT_QVV40, math_mode=ab.MATH_FAST), d_fast)
    torch.cuda.synchronize()
    assert float((d_fast[:, :, :4] - d_out[:, :, :4]).abs().max()) <= FAST_MATH_TOLERANCE
    assert torch.equal(d_fast[:, :, 4:].contiguous().view(torch.int32), d_out[:, :, 4:].contiguous().view(torch.int32))
    del d_fast

    # (5) bit-exact against the oracle on a random sample + the very first and last requests
    rng = np.random.default_rng(2)
    sample = np.unique(np.concatenate([rng.integers(0, n_req, size=3000), [0, n_req - 1]]))
    got = d_out[torch.from_numpy(sample).cuda()].cpu().numpy()
    settings = port.settings_for_kind(0)
    for j, r in enumerate(sample):
        clip = int(w["req_clip"][r])
        blob = w["buffer"][int(w["offsets"][clip]):int(w["offsets"][clip]) + int(w["sizes"][clip])]
        want = port.transform_decompress_tracks(blob, settings, float(w["req_time"][r]))
        assert clips.bit_equal(got[j], want[:, LANES]), (int(r), clip)


def test_skip_masks_leave_sub_tracks_untouched(gpu):
    """track_writer::skip_all_*() / skip_track_*(track) (core/track_writer.h:181-191): a skipped sub-track is not written --
    the caller's buffer keeps its bytes -- everything else is bit-exact."""
    port, ab, ctx, torch = gpu["port"], gpu["ab"], gpu["ctx"], gpu["torch"]
    name = "mixed_scale"
    spec, blob = clips.TRANSFORM_SPECS[name], clips.load_blob(name)
    clipset = ctx.upload([blob], check_hash=True)
    times = clips.sample_times(spec)
    n = clipset.max_tracks
    settings = port.settings_for_kind(0)
    rng = np.random.default_rng(4)
    per_track = rng.integers(0, 8, size=n).astype(np.uint8)
    d_per_track = _to_device(gpu, per_track)
    marker = np.float32(-12345.5)
    lane_kind = {0: 0, 1: 0, 2: 0, 3: 0, 4: 1, 5: 1, 6: 1, 8: 2, 9: 2, 10: 2}
    for skip_all, use_tracks in ((ab.SKIP_ROTATION, False), (ab.SKIP_TRANSLATION | ab.SKIP_SCALE, False), (0, True), (ab.SKIP_SCALE, True)):
        options = _options(gpu, settings, skip_mask=skip_all, d_skip_track_mask=d_per_track.data_ptr() if use_tracks else None)
        prefill = np.full((len(times), n, 12), marker, dtype=np.float32)
        got = _decode(gpu, clipset, np.zeros(len(times), np.uint32), times, options, prefill=prefill)
        for i, t in enumerate(times):
            want = port.transform_decompress_tracks(blob, settings, float(t))
            for track in range(n):
                bits = skip_all | (int(per_track[track]) if use_tracks else 0)
                for lane in LANES:
                    if (bits >> lane_kind[lane]) & 1:
                        assert got[i, track, lane] == marker, (skip_all, use_tracks, track, lane)
                    else:
                        assert got[i, track, lane].view(np.uint32) == want[track, lane].view(np.uint32), (skip_all, use_tracks, float(t), track, lane)
    clipset.release()


def test_per_request_rounding_and_looping(gpu):
    """seek(t, rounding) + set_looping_policy(policy) per request (decompress.h:147-160): one launch mixing every pair equals the
    per policy launches."""
    port, ab, ctx = gpu["port"], gpu["ab"], gpu["ctx"]
    names = ["looping", "stripped_loop", "c1_30bones", "mixed_scale"]
    blobs = [clips.load_blob(nm) for nm in names]
    clipset = ctx.upload(blobs, check_hash=True)
    rng = np.random.default_rng(6)
    count = 240
    req_clip = rng.integers(0, len(names), count).astype(np.uint32)
    req_time = np.array([rng.uniform(-0.2, (clips.TRANSFORM_SPECS[names[c]].num_samples + 2) / 30.0) for c in req_clip], dtype=np.float32)
    policies = np.stack([rng.integers(0, 4, count), rng.integers(0, 3, count)], axis=1).astype(np.uint8)
    settings = port.settings_for_kind(0)
    d_policies = _to_device(gpu, policies)
    got = _decode(gpu, clipset, req_clip, req_time, _options(gpu, settings, d_request_policies=d_policies.data_ptr()))
    for i in range(count):
        want = port.transform_decompress_tracks(blobs[req_clip[i]], settings, float(req_time[i]), int(policies[i, 0]), int(policies[i, 1]))
        nt = want.shape[0]
        assert clips.bit_equal(got[i, :nt][:, LANES], want[:, LANES]), (i, names[req_clip[i]], float(req_time[i]), policies[i].tolist())
    clipset.release()
