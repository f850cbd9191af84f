"""ctypes binding of oracle/_ref/libaclref.so -- the UNMODIFIED reference compiled from /root/reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU baseline /
`--impl reference` legs. The product package (acl_b200/) never imports this module.

The library is built by oracle/Makefile (see oracle/ref_tool.cpp). On the GPU box the prebuilt .so
travels with the repository snapshot; /root/reference is never read at run time.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field, asdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libaclref.so")

# acl::rotation_format8 / vector_format8 (core/track_formats.h)
QUATF_FULL, QUATF_DROP_W_FULL, QUATF_DROP_W_VARIABLE = 0, 2, 3
VECTOR3F_FULL, VECTOR3F_VARIABLE = 0, 1
# acl::compression_level8 (compression/compression_level.h)
LEVEL_LOWEST, LEVEL_LOW, LEVEL_MEDIUM, LEVEL_HIGH, LEVEL_HIGHEST, LEVEL_AUTOMATIC = 0, 1, 2, 3, 4, 100
# acl::sample_rounding_policy (core/sample_rounding_policy.h)
ROUND_NONE, ROUND_FLOOR, ROUND_CEIL, ROUND_NEAREST, ROUND_PER_TRACK = 0, 1, 2, 3, 4
# acl::sample_looping_policy (core/sample_looping_policy.h)
LOOP_CLAMP, LOOP_WRAP, LOOP_AS_COMPRESSED = 0, 1, 2
# acl::track_type8 (core/track_types.h)
TRACK_FLOAT1F, TRACK_FLOAT2F, TRACK_FLOAT3F, TRACK_FLOAT4F, TRACK_VECTOR4F, TRACK_QVVF = 0, 1, 2, 3, 4, 12

# settings kinds understood by ref_tool.cpp
SETTINGS_DEFAULT, SETTINGS_DEBUG, SETTINGS_BENCHMARK, SETTINGS_NEVER, SETTINGS_ALL_LERP, SETTINGS_RAW_ONLY = 0, 1, 2, 3, 4, 5
# writer modes understood by ref_tool.cpp
WRITER_LEGACY, WRITER_SKIPPED, WRITER_CONSTANT, WRITER_VARIABLE = 0, 1, 2, 3


class _TransformSpec(C.Structure):
    _fields_ = [
        ("num_tracks", C.c_uint32), ("num_samples", C.c_uint32), ("sample_rate", C.c_float), ("seed", C.c_uint32),
        ("rot_default_pct", C.c_uint32), ("rot_constant_pct", C.c_uint32),
        ("trans_default_pct", C.c_uint32), ("trans_constant_pct", C.c_uint32),
        ("scale_default_pct", C.c_uint32), ("scale_constant_pct", C.c_uint32),
        ("partial_activity_pct", C.c_uint32), ("noisy_pct", C.c_uint32), ("looping_content", C.c_uint32),
        ("translation_range", C.c_float), ("precision", C.c_float), ("shell_distance", C.c_float),
        ("rotation_format", C.c_uint32), ("translation_format", C.c_uint32), ("scale_format", C.c_uint32),
        ("level", C.c_uint32), ("optimize_loops", C.c_uint32), ("strip_trivial", C.c_uint32),
        ("strip_proportion", C.c_float), ("strip_threshold", C.c_float), ("rotation_offset", C.c_float),
        ("negative_scale_pct", C.c_uint32),
    ]


class _ScalarSpec(C.Structure):
    _fields_ = [
        ("num_tracks", C.c_uint32), ("num_samples", C.c_uint32), ("sample_rate", C.c_float), ("seed", C.c_uint32),
        ("track_type", C.c_uint32), ("constant_pct", C.c_uint32), ("noisy_pct", C.c_uint32), ("precision", C.c_float),
    ]


class SeekInfo(C.Structure):
    _fields_ = [
        ("sample_time", C.c_float), ("interpolation_alpha", C.c_float),
        ("key_frame_bit_offsets", C.c_uint32 * 2), ("segment_offsets", C.c_uint32 * 2),
        ("format_offsets", C.c_uint32 * 2), ("range_offsets", C.c_uint32 * 2), ("animated_offsets", C.c_uint32 * 2),
        ("uses_single_segment", C.c_uint32), ("clip_duration", C.c_float), ("looping_policy", C.c_uint32),
    ]


@dataclass
class TransformSpec:
    """Synthetic clip recipe: the default values are the humanoid of config C2."""
    num_tracks: int = 100
    num_samples: int = 60
    sample_rate: float = 30.0
    seed: int = 2000
    rot_default_pct: int = 0
    rot_constant_pct: int = 0
    trans_default_pct: int = 0
    trans_constant_pct: int = 90
    scale_default_pct: int = 100
    scale_constant_pct: int = 0
    partial_activity_pct: int = 0
    noisy_pct: int = 0
    looping_content: int = 0
    translation_range: float = 10.0
    precision: float = 0.01
    shell_distance: float = 3.0
    rotation_format: int = QUATF_DROP_W_VARIABLE
    translation_format: int = VECTOR3F_VARIABLE
    scale_format: int = VECTOR3F_VARIABLE
    level: int = LEVEL_AUTOMATIC
    optimize_loops: int = 1
    strip_trivial: int = 1
    strip_proportion: float = 0.0
    strip_threshold: float = 0.0
    rotation_offset: float = 0.0
    negative_scale_pct: int = 0

    def to_c(self) -> _TransformSpec:
        return _TransformSpec(**asdict(self))


@dataclass
class ScalarSpec:
    num_tracks: int = 4096
    num_samples: int = 1024
    sample_rate: float = 30.0
    seed: int = 42
    track_type: int = TRACK_FLOAT1F
    constant_pct: int = 12
    noisy_pct: int = 0
    precision: float = 0.001

    def to_c(self) -> _ScalarSpec:
        return _ScalarSpec(**asdict(self))


_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{_LIB_PATH} is missing: run `make -C oracle` where /root/reference exists")
        l = C.CDLL(_LIB_PATH)
        l.aclref_version.restype = C.c_char_p
        l.aclref_hardware_threads.restype = C.c_uint32
        l.aclref_compress_transform.argtypes = [C.POINTER(_TransformSpec), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
        l.aclref_compress_scalar.argtypes = [C.POINTER(_ScalarSpec), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
        l.aclref_free.argtypes = [C.c_void_p]
        l.aclref_is_valid.argtypes = [C.c_void_p, C.c_uint32]
        l.aclref_decompress_tracks.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.aclref_decompress_track.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.aclref_scalar_decompress.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32,
                                               C.c_void_p, C.c_void_p]
        l.aclref_seek_info_transform.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.POINTER(SeekInfo)]
        for name in ("aclref_bench_transform", "aclref_bench_scalar"):
            fn = getattr(l, name)
            fn.restype = C.c_double
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        l.aclref_compress_transform_batch.restype = C.c_uint64
        l.aclref_compress_transform_batch.argtypes = [C.POINTER(_TransformSpec), C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64,
                                                      C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def aligned_blob(data: bytes | np.ndarray, slack: int = 64) -> np.ndarray:
    """Copy a blob into a 64-byte aligned uint8 array with `slack` zero bytes after it (the reference's
    `_unsafe` unpackers read up to 15 bytes past the bit stream, compress.transform.impl.h:387-396)."""
    src = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    raw = np.zeros(src.size + slack + 64, dtype=np.uint8)
    shift = (-raw.ctypes.data) % 64
    out = raw[shift:shift + src.size + slack]
    out[:src.size] = src
    return out[:src.size]


def _take(ptr: C.c_void_p, size: int) -> np.ndarray:
    buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(size,))
    out = aligned_blob(buf.copy())
    lib().aclref_free(ptr)
    return out


def compress_transform(spec: TransformSpec) -> np.ndarray:
    """Synthesise + compress one transform clip with the reference compressor. Returns the blob bytes."""
    c_spec = spec.to_c()
    ptr, size = C.c_void_p(), C.c_uint32()
    rc = lib().aclref_compress_transform(C.byref(c_spec), C.byref(ptr), C.byref(size))
    if rc != 0:
        raise RuntimeError(f"reference compression failed ({rc}) for {spec}")
    return _take(ptr, size.value)


def compress_scalar(spec: ScalarSpec) -> np.ndarray:
    c_spec = spec.to_c()
    ptr, size = C.c_void_p(), C.c_uint32()
    rc = lib().aclref_compress_scalar(C.byref(c_spec), C.byref(ptr), C.byref(size))
    if rc != 0:
        raise RuntimeError(f"reference compression failed ({rc}) for {spec}")
    return _take(ptr, size.value)


def compress_transform_batch(spec: TransformSpec, num_clips: int, num_threads: int = 0, bytes_per_clip_hint: int = 0):
    """Compress `num_clips` clips (seed, seed+1, ...) on `num_threads` host threads.
    Returns (buffer uint8[...], offsets uint64[num_clips], sizes uint32[num_clips]); every blob starts on a
    64 byte boundary inside `buffer`."""
    if bytes_per_clip_hint <= 0:
        probe = compress_transform(spec)
        bytes_per_clip_hint = int(probe.size * 1.5) + 256
    capacity = num_clips * bytes_per_clip_hint + 4096
    raw = np.zeros(capacity + 64, dtype=np.uint8)
    shift = (-raw.ctypes.data) % 64
    buffer = raw[shift:shift + capacity]
    offsets = np.zeros(num_clips, dtype=np.uint64)
    sizes = np.zeros(num_clips, dtype=np.uint32)
    c_spec = spec.to_c()
    used = lib().aclref_compress_transform_batch(C.byref(c_spec), num_clips, num_threads, buffer.ctypes.data, capacity,
                                                 offsets.ctypes.data, sizes.ctypes.data)
    if used == 0:
        raise RuntimeError("reference batch compression failed (or the size hint was too small)")
    return buffer[:used], offsets, sizes


def _ptr(a):
    return None if a is None else a.ctypes.data


def num_tracks_of(blob: np.ndarray) -> int:
    return int(blob[16:20].view(np.uint32)[0])


def decompress_tracks(blob: np.ndarray, t: float, rounding: int = ROUND_NONE, looping: int = LOOP_AS_COMPRESSED,
                      settings: int = SETTINGS_DEFAULT, writer: int = WRITER_LEGACY,
                      per_track_rounding: np.ndarray | None = None, constant_defaults: np.ndarray | None = None,
                      variable_defaults: np.ndarray | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """Reference seek(t, rounding) + decompress_tracks(writer). Returns float32 [num_tracks, 12]
    (rotation xyzw, translation xyz + pad, scale xyz + pad). `out` lets the caller pre-fill the pose (skipped mode)."""
    n = num_tracks_of(blob)
    if out is None:
        out = np.zeros((n, 12), dtype=np.float32)
    rc = lib().aclref_decompress_tracks(blob.ctypes.data, t, rounding, looping, settings, writer,
                                        _ptr(per_track_rounding), _ptr(constant_defaults), _ptr(variable_defaults), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"reference decompress_tracks failed ({rc})")
    return out


def decompress_track(blob: np.ndarray, t: float, track_index: int, rounding: int = ROUND_NONE, looping: int = LOOP_AS_COMPRESSED,
                     settings: int = SETTINGS_DEFAULT, writer: int = WRITER_LEGACY,
                     per_track_rounding: np.ndarray | None = None, constant_defaults: np.ndarray | None = None,
                     variable_defaults: np.ndarray | None = None, out: np.ndarray | None = None) -> np.ndarray:
    n = num_tracks_of(blob)
    if out is None:
        out = np.zeros((n, 12), dtype=np.float32)
    rc = lib().aclref_decompress_track(blob.ctypes.data, t, rounding, looping, settings, writer, track_index,
                                       _ptr(per_track_rounding), _ptr(constant_defaults), _ptr(variable_defaults), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"reference decompress_track failed ({rc})")
    return out


def scalar_decompress(blob: np.ndarray, t: float, rounding: int = ROUND_NONE, looping: int = LOOP_AS_COMPRESSED,
                      settings: int = 0, track_index: int = -1, per_track_rounding: np.ndarray | None = None) -> np.ndarray:
    """Reference scalar decompress_tracks (track_index < 0) or decompress_track. Returns float32 [num_tracks, 4]."""
    n = num_tracks_of(blob)
    out = np.zeros((n, 4), dtype=np.float32)
    rc = lib().aclref_scalar_decompress(blob.ctypes.data, t, rounding, looping, settings, track_index, _ptr(per_track_rounding), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"reference scalar decompress failed ({rc})")
    return out


def seek_info(blob: np.ndarray, t: float, rounding: int = ROUND_NONE, looping: int = LOOP_AS_COMPRESSED) -> SeekInfo:
    info = SeekInfo()
    rc = lib().aclref_seek_info_transform(blob.ctypes.data, t, rounding, looping, C.byref(info))
    if rc != 0:
        raise RuntimeError("reference seek_info failed")
    return info


def bench(blobs: list[np.ndarray], request_clip: np.ndarray, request_time: np.ndarray, max_tracks: int,
          num_threads: int, repeats: int = 1, scalar: bool = False, out: np.ndarray | None = None) -> float:
    """Seconds taken by the reference CPU path (fastest of `repeats` passes) for the request list. `out` (optional):
    float32 [requests][max_tracks][12] (transform: rtm::qvvf as debug_track_writer stores it) / [requests][max_tracks][4] (scalar)."""
    ptrs = (C.c_void_p * len(blobs))(*[b.ctypes.data for b in blobs])
    request_clip = np.ascontiguousarray(request_clip, dtype=np.uint32)
    request_time = np.ascontiguousarray(request_time, dtype=np.float32)
    fn = lib().aclref_bench_scalar if scalar else lib().aclref_bench_transform
    if out is not None:
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == request_clip.size * max_tracks * (4 if scalar else 12)
    return float(fn(C.cast(ptrs, C.c_void_p), request_clip.ctypes.data, request_time.ctypes.data, request_clip.size,
                    max_tracks, num_threads, repeats, None if out is None else out.ctypes.data))


def decode_requests(blobs: list[np.ndarray], request_clip: np.ndarray, request_time: np.ndarray, max_tracks: int,
                    num_threads: int = 0, scalar: bool = False) -> np.ndarray:
    """Every request through the unmodified reference (acl::decompression_context<benchmark settings>, seek(t, none) +
    decompress_tracks): float32 [requests][max_tracks][12] (scalar clips: [requests][max_tracks][4], one float per float1f track)."""
    out = np.zeros((len(request_clip), max_tracks, 4 if scalar else 12), dtype=np.float32)
    bench(blobs, request_clip, request_time, max_tracks, num_threads or usable_threads(), 1, scalar=scalar, out=out)
    return out


def usable_threads() -> int:
    """Host threads this process may really use: the scheduler affinity, bounded by the cgroup CPU quota (a container lease often
    exposes every core of the box through hardware_concurrency() while the quota grants a fraction of them)."""
    import math
    import os
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            text = open(path).read().split()
            if path.endswith("cpu.max"):
                if text[0] != "max":
                    threads = min(threads, max(1, math.ceil(int(text[0]) / int(text[1]))))
            else:
                quota = int(text[0])
                period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if quota > 0:
                    threads = min(threads, max(1, math.ceil(quota / period)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, threads)


# the reference's calculate_compression_error and the values it works on (oracle/ref_tool.cpp)
class TrackError(C.Structure):
    _fields_ = [("index", C.c_uint32), ("error", C.c_float), ("sample_time", C.c_float)]


def finite_duration(num_samples: int, sample_rate: float) -> float:
    """calculate_finite_duration (core/impl/time_utils.impl.h:105-114) of a raw clip, in float arithmetic like the reference."""
    if num_samples <= 1:
        return 0.0
    return float(np.float32(num_samples - 1) / np.float32(sample_rate))


def transform_error(spec: TransformSpec, blob: np.ndarray, settings: int = SETTINGS_DEBUG) -> dict:
    """calculate_compression_error(allocator, raw clip of `spec`, decompression_context<settings>(blob), qvvf_transform_error_metric).
    Returns the track_error plus the function's inputs and intermediates recomputed with the reference's own classes:
    raw_poses / lossy_poses float32 [num_samples][num_tracks][12], object_poses [2][num_samples][num_tracks][12],
    errors [num_samples][num_tracks], parents uint32 [num_tracks], shell_distances float32 [num_tracks], rounding."""
    n, m = spec.num_tracks, spec.num_samples
    out = dict(raw_poses=np.zeros((m, n, 12), np.float32), lossy_poses=np.zeros((m, n, 12), np.float32),
               object_poses=np.zeros((2, m, n, 12), np.float32), errors=np.zeros((m, n), np.float32),
               parents=np.zeros(n, np.uint32), shell_distances=np.zeros(n, np.float32))
    result, rounding = TrackError(), C.c_uint32()
    fn = lib().aclref_transform_error
    fn.argtypes = [C.POINTER(_TransformSpec), C.c_void_p, C.c_uint32, C.POINTER(TrackError), C.POINTER(C.c_uint32)] + [C.c_void_p] * 6
    c_spec = spec.to_c()
    rc = fn(C.byref(c_spec), blob.ctypes.data, settings, C.byref(result), C.byref(rounding), out["raw_poses"].ctypes.data,
            out["lossy_poses"].ctypes.data, out["object_poses"].ctypes.data, out["errors"].ctypes.data, out["parents"].ctypes.data,
            out["shell_distances"].ctypes.data)
    if rc != 0:
        raise RuntimeError(f"aclref_transform_error failed ({rc})")
    out.update(index=int(result.index), error=float(result.error), sample_time=float(result.sample_time), rounding=int(rounding.value),
               sample_rate=float(spec.sample_rate), duration=finite_duration(m, spec.sample_rate))
    return out


def scalar_error(spec: ScalarSpec, blob: np.ndarray) -> dict:
    """calculate_compression_error(allocator, raw clip of `spec`, decompression_context<default_scalar>(blob)) + the raw samples
    float32 [num_samples][num_tracks][4] it compares with."""
    raw = np.zeros((spec.num_samples, spec.num_tracks, 4), np.float32)
    result, rounding = TrackError(), C.c_uint32()
    fn = lib().aclref_scalar_error
    fn.argtypes = [C.POINTER(_ScalarSpec), C.c_void_p, C.POINTER(TrackError), C.POINTER(C.c_uint32), C.c_void_p]
    c_spec = spec.to_c()
    rc = fn(C.byref(c_spec), blob.ctypes.data, C.byref(result), C.byref(rounding), raw.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"aclref_scalar_error failed ({rc})")
    return dict(index=int(result.index), error=float(result.error), sample_time=float(result.sample_time), rounding=int(rounding.value),
                raw_values=raw, sample_rate=float(spec.sample_rate), duration=finite_duration(spec.num_samples, spec.sample_rate))


def bench_transform_error(spec: TransformSpec, blobs: list[np.ndarray], num_threads: int) -> tuple[float, np.ndarray]:
    """Seconds the reference's calculate_compression_error takes for clips spec.seed .. spec.seed + len(blobs) - 1 on `num_threads`
    host threads (raw clips rebuilt outside the timed region), and the per clip (index, error, sample_time) records."""
    ptrs = (C.c_void_p * len(blobs))(*[b.ctypes.data for b in blobs])
    errors = np.zeros(len(blobs), dtype=np.dtype([("index", np.uint32), ("error", np.float32), ("sample_time", np.float32)]))
    fn = lib().aclref_bench_transform_error
    fn.restype = C.c_double
    fn.argtypes = [C.POINTER(_TransformSpec), C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    c_spec = spec.to_c()
    seconds = float(fn(C.byref(c_spec), C.cast(ptrs, C.c_void_p), len(blobs), num_threads, errors.ctypes.data))
    return seconds, errors


def sample_raw_transform_batch(spec: TransformSpec, num_clips: int, num_threads: int = 0):
    """Raw poses of clips spec.seed .. spec.seed + num_clips - 1 as calculate_compression_error samples them (nearest):
    (float32 [num_clips][num_samples][num_tracks][12], parents uint32 [num_tracks], shell_distances float32 [num_tracks])."""
    raw = np.zeros((num_clips, spec.num_samples, spec.num_tracks, 12), np.float32)
    parents, shells = np.zeros(spec.num_tracks, np.uint32), np.zeros(spec.num_tracks, np.float32)
    fn = lib().aclref_sample_raw_transform_batch
    fn.argtypes = [C.POINTER(_TransformSpec), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    c_spec = spec.to_c()
    rc = fn(C.byref(c_spec), num_clips, num_threads or usable_threads(), raw.ctypes.data, parents.ctypes.data, shells.ctypes.data)
    if rc != 0:
        raise RuntimeError("aclref_sample_raw_transform_batch failed")
    return raw, parents, shells


ADDITIVE_NONE, ADDITIVE_RELATIVE, ADDITIVE_ADDITIVE0, ADDITIVE_ADDITIVE1 = 0, 1, 2, 3     # acl::additive_clip_format8 (core/additive_utils.h:42-66)


def transform_error_additive(spec: TransformSpec, blob: np.ndarray, base_spec: TransformSpec, additive_format: int) -> dict:
    """calculate_compression_error(allocator, raw clip of `spec`, context(blob), additive_qvvf_transform_error_metric<format>, raw clip of
    `base_spec`) with debug settings. As transform_error, plus base_poses float32 [num_samples][num_tracks][12] (the base clip sampled at the
    matching times) and errors measured after apply_additive_to_base."""
    n, m = spec.num_tracks, spec.num_samples
    out = dict(raw_poses=np.zeros((m, n, 12), np.float32), lossy_poses=np.zeros((m, n, 12), np.float32), base_poses=np.zeros((m, n, 12), np.float32),
               errors=np.zeros((m, n), np.float32), parents=np.zeros(n, np.uint32), shell_distances=np.zeros(n, np.float32))
    result, rounding = TrackError(), C.c_uint32()
    fn = lib().aclref_transform_error_additive
    fn.argtypes = [C.POINTER(_TransformSpec), C.c_void_p, C.POINTER(_TransformSpec), C.c_uint32, C.POINTER(TrackError), C.POINTER(C.c_uint32)] + [C.c_void_p] * 6
    c_spec, c_base = spec.to_c(), base_spec.to_c()
    rc = fn(C.byref(c_spec), blob.ctypes.data, C.byref(c_base), additive_format, C.byref(result), C.byref(rounding), out["raw_poses"].ctypes.data,
            out["lossy_poses"].ctypes.data, out["base_poses"].ctypes.data, out["errors"].ctypes.data, out["parents"].ctypes.data,
            out["shell_distances"].ctypes.data)
    if rc != 0:
        raise RuntimeError(f"aclref_transform_error_additive failed ({rc})")
    out.update(index=int(result.index), error=float(result.error), sample_time=float(result.sample_time), rounding=int(rounding.value),
               sample_rate=float(spec.sample_rate), duration=finite_duration(m, spec.sample_rate), additive_format=additive_format)
    return out


METRIC_QVVF, METRIC_QVVF_MATRIX3X4F = 0, 1


def transform_error_matrix(spec: TransformSpec, blob: np.ndarray) -> dict:
    """calculate_compression_error with qvvf_matrix3x4f_transform_error_metric (debug settings): index / error / sample_time and the per bone
    errors float32 [num_samples][num_tracks] replayed with the metric's own functions. The poses are transform_error(spec, blob, 1)'s."""
    errors = np.zeros((spec.num_samples, spec.num_tracks), np.float32)
    result = TrackError()
    fn = lib().aclref_transform_error_matrix
    fn.argtypes = [C.POINTER(_TransformSpec), C.c_void_p, C.POINTER(TrackError), C.c_void_p]
    c_spec = spec.to_c()
    rc = fn(C.byref(c_spec), blob.ctypes.data, C.byref(result), errors.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"aclref_transform_error_matrix failed ({rc})")
    return dict(index=int(result.index), error=float(result.error), sample_time=float(result.sample_time), errors=errors)


def compress_transform_database(spec: TransformSpec, medium_proportion: float = 0.0, low_proportion: float = 0.5) -> np.ndarray:
    """The raw clip of `spec` compressed with database support and split by acl::build_database: returns the compressed_tracks blob bound
    to the database (its movable key frames moved out; the database itself is dropped)."""
    c_spec = spec.to_c()
    ptr, size = C.c_void_p(), C.c_uint32()
    fn = lib().aclref_compress_transform_database
    fn.argtypes = [C.POINTER(_TransformSpec), C.c_float, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
    rc = fn(C.byref(c_spec), medium_proportion, low_proportion, C.byref(ptr), C.byref(size))
    if rc != 0:
        raise RuntimeError(f"reference database compression failed ({rc}) for {spec}")
    return _take(ptr, size.value)


def decompress_tracks_without_database(blob: np.ndarray, t: float, rounding: int = ROUND_NONE, looping: int = LOOP_AS_COMPRESSED,
                                       writer: int = WRITER_LEGACY) -> np.ndarray:
    """decompression_context<debug settings + database support>::initialize(tracks) with NO database bound, seek, decompress_tracks."""
    n = num_tracks_of(blob)
    out = np.zeros((n, 12), dtype=np.float32)
    fn = lib().aclref_decompress_tracks_without_database
    fn.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    rc = fn(blob.ctypes.data, t, rounding, looping, writer, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"reference decompress_tracks (database clip, no database) failed ({rc})")
    return out
