"""Regenerates tests/golden/*: reference-compressed clips + the reference's own decompression outputs.

Run where oracle/_ref/libaclref.so exists (i.e. where /root/reference is mounted and `make -C oracle` was run):

    python tests/golden/make_golden.py

For every named clip of tests/clips.py this writes
    <name>.acl.bin        the compressed_tracks blob produced by acl::compress_track_list
    <name>.golden.npz     outputs of acl::decompression_context (seek + decompress_tracks / decompress_track) and the
                          integers seek() leaves in the context, for a fixed list of times / settings / rounding policies
The outputs come from the UNMODIFIED reference (oracle/ref_tool.cpp); nothing here involves the port or the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from oracle import ref  # noqa: E402
from tests import clips  # noqa: E402

# (settings kind, rounding policy) pairs stored for every transform clip
COMBOS = [(0, 0), (0, 1), (0, 2), (0, 3), (1, 0), (1, 1), (1, 2), (1, 3), (3, 0), (4, 0)]
COMBOS_FULL = COMBOS + [(5, 0), (5, 1)]     # quatf_full clips also run the raw-only settings
BIG = {"paragon_like"}                      # keep the large skeleton light


def golden_times(spec) -> np.ndarray:
    t = clips.sample_times(spec)
    # 8 of them: before start, key frame, mid frame, end, past the end ...
    pick = [0, 2, 4, 8, 9, 10, 11, 13]
    return t[[i for i in pick if i < len(t)]]


def make_transform(name: str, spec) -> None:
    blob = ref.compress_transform(spec)
    with open(clips.golden_path(name, "acl.bin"), "wb") as f:
        f.write(blob.tobytes())

    times = golden_times(spec)
    is_full = spec.rotation_format == ref.QUATF_FULL
    supports_default = spec.rotation_format == ref.QUATF_DROP_W_VARIABLE and spec.translation_format == ref.VECTOR3F_VARIABLE \
        and spec.scale_format == ref.VECTOR3F_VARIABLE
    combos = COMBOS_FULL if is_full else COMBOS
    if not supports_default:
        combos = [c for c in combos if c[0] not in (0, 2)]     # default settings only decode the variable formats
    if name in BIG:
        combos = [(0, 0), (1, 3)]
        times = times[:5]

    n = ref.num_tracks_of(blob)
    poses = np.zeros((len(combos), len(times), n, 10), dtype=np.float32)
    for ci, (kind, rounding) in enumerate(combos):
        for ti, t in enumerate(times):
            poses[ci, ti] = ref.decompress_tracks(blob, float(t), rounding, settings=kind)[:, clips.DEFINED_LANES]

    # decompress_track for a handful of bones (debug settings, nearest + none)
    bones = sorted(set([0, n // 3, n // 2, n - 1]))
    single = np.zeros((2, len(times), len(bones), 10), dtype=np.float32)
    single_kind = 1
    for ri, rounding in enumerate((0, 3)):
        for ti, t in enumerate(times):
            for bi, bone in enumerate(bones):
                single[ri, ti, bi] = ref.decompress_track(blob, float(t), bone, rounding, settings=single_kind)[bone, clips.DEFINED_LANES]

    # seek integers for clamp / wrap / as_compressed x none / nearest
    seek_rows = []
    for looping in (0, 1, 2):
        for rounding in (0, 3):
            for t in times:
                info = ref.seek_info(blob, float(t), rounding, looping)
                seek_rows.append([looping, rounding, np.float32(t).view(np.uint32), np.float32(info.sample_time).view(np.uint32),
                                  np.float32(info.interpolation_alpha).view(np.uint32),
                                  info.key_frame_bit_offsets[0], info.key_frame_bit_offsets[1],
                                  info.animated_offsets[0], info.animated_offsets[1],
                                  info.format_offsets[0], info.format_offsets[1], info.range_offsets[0], info.range_offsets[1],
                                  info.uses_single_segment, info.looping_policy])
    np.savez_compressed(clips.golden_path(name, "golden.npz"), times=times, combos=np.array(combos, dtype=np.int32), poses=poses,
                        bones=np.array(bones, dtype=np.int32), single=single, seek=np.array(seek_rows, dtype=np.uint32))


def make_scalar(name: str, spec) -> None:
    blob = ref.compress_scalar(spec)
    with open(clips.golden_path(name, "acl.bin"), "wb") as f:
        f.write(blob.tobytes())
    times = golden_times(spec)
    n = ref.num_tracks_of(blob)
    nc = min(spec.track_type + 1, 4)
    values = np.zeros((4, 3, len(times), n, nc), dtype=np.float32)
    for rounding in range(4):
        for looping in range(3):
            for ti, t in enumerate(times):
                values[rounding, looping, ti] = ref.scalar_decompress(blob, float(t), rounding, looping, 0)[:, :nc]
    np.savez_compressed(clips.golden_path(name, "golden.npz"), times=times, values=values)


# clips bound to a streaming database (acl::build_database), decoded by the reference WITHOUT the database (first step)
DATABASE_GOLDEN = {"database_c1_30bones": ("c1_30bones", 0.0, 0.5), "database_mixed_scale": ("mixed_scale", 0.3, 0.3)}


def make_database(golden_name: str, clip_name: str, medium: float, low: float) -> None:
    spec = clips.TRANSFORM_SPECS[clip_name]
    blob = ref.compress_transform_database(spec, medium, low)
    with open(clips.golden_path(golden_name, "acl.bin"), "wb") as f:
        f.write(blob.tobytes())
    times = clips.sample_times(spec)
    poses = np.zeros((4, len(times), spec.num_tracks, 10), dtype=np.float32)
    for rounding in range(4):
        for ti, t in enumerate(times):
            poses[rounding, ti] = ref.decompress_tracks_without_database(blob, float(t), rounding)[:, clips.DEFINED_LANES]
    np.savez_compressed(clips.golden_path(golden_name, "golden.npz"), times=times, poses=poses)


def main() -> None:
    os.makedirs(clips.GOLDEN_DIR, exist_ok=True)
    for golden_name, (clip_name, medium, low) in DATABASE_GOLDEN.items():
        make_database(golden_name, clip_name, medium, low)
        print("database", golden_name)
    for name, spec in clips.TRANSFORM_SPECS.items():
        make_transform(name, spec)
        print("transform", name)
    for name, spec in clips.SCALAR_SPECS.items():
        make_scalar(name, spec)
        print("scalar", name)
    total = sum(os.path.getsize(os.path.join(clips.GOLDEN_DIR, f)) for f in os.listdir(clips.GOLDEN_DIR))
    print(f"golden directory: {total / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
