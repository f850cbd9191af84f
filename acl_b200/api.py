"""ctypes binding of libaclb200.so (include/aclb200.h). No decode logic lives here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("ACLB200_LIB", os.path.join(_HERE, "libaclb200.so"))

ROUND_NONE, ROUND_FLOOR, ROUND_CEIL, ROUND_NEAREST, ROUND_PER_TRACK = 0, 1, 2, 3, 4
LOOP_CLAMP, LOOP_WRAP, LOOP_AS_COMPRESSED = 0, 1, 2
NORMALIZE_NEVER, NORMALIZE_LERP_ONLY, NORMALIZE_ALWAYS = 0, 1, 2
DEFAULT_SKIPPED, DEFAULT_CONSTANT, DEFAULT_VARIABLE, DEFAULT_LEGACY = 0, 1, 2, 3
LAYOUT_QVV48, LAYOUT_QVV40 = 0, 1
MATH_EXACT, MATH_FAST = 0, 1
SKIP_ROTATION, SKIP_TRANSLATION, SKIP_SCALE = 1, 2, 4
TRACK_QVVF = 12

# numpy view of aclb200_request {uint32 clip; float sample_time}
REQUEST_DTYPE = np.dtype([("clip", np.uint32), ("sample_time", np.float32)])
SEEK_STATE_DTYPE = np.dtype([
    ("sample_time", np.float32), ("interpolation_alpha", np.float32),
    ("key_frame_bit_offsets", np.uint32, 2), ("segment_indices", np.uint32, 2), ("animated_offsets", np.uint32, 2),
    ("format_offsets", np.uint32, 2), ("range_offsets", np.uint32, 2),
    ("uses_single_segment", np.uint32), ("looping_policy", np.uint32),
])


# numpy views of aclb200_error_job / aclb200_track_error
ERROR_JOB_DTYPE = np.dtype([("clip", np.uint32), ("num_samples", np.uint32), ("sample_rate", np.float32), ("duration", np.float32),
                            ("num_tracks", np.uint32), ("skeleton_offset", np.uint32), ("first_raw_pose", np.uint64),
                            ("additive_format", np.uint32), ("error_metric", np.uint32), ("first_base_pose", np.uint64)])
METRIC_QVVF, METRIC_QVVF_MATRIX3X4F = 0, 1
ADDITIVE_NONE, ADDITIVE_RELATIVE, ADDITIVE_ADDITIVE0, ADDITIVE_ADDITIVE1 = 0, 1, 2, 3
TRACK_ERROR_DTYPE = np.dtype([("index", np.uint32), ("error", np.float32), ("sample_time", np.float32), ("flags", np.uint32)])
ERROR_FLAG_NEGATIVE_SCALE, ERROR_FLAG_INVALID_SKELETON = 1, 2


class AclB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"aclb200 status {status}: {message}")
        self.status = status


class Options(C.Structure):
    """aclb200_options"""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("rounding_policy", C.c_uint32), ("looping_policy", C.c_uint32),
        ("normalization", C.c_uint32), ("per_track_rounding", C.c_uint32), ("wrapping", C.c_uint32),
        ("clamp_sample_time", C.c_uint32), ("multiple_rotation_formats", C.c_uint32),
        ("default_rotation_mode", C.c_uint32), ("default_translation_mode", C.c_uint32), ("default_scale_mode", C.c_uint32),
        ("constant_defaults", C.c_float * 12),
        ("d_variable_defaults", C.c_void_p), ("d_per_track_rounding", C.c_void_p),
        ("output_layout", C.c_uint32), ("math_mode", C.c_uint32),
        ("pose_stride_bytes", C.c_uint64),
        ("skip_mask", C.c_uint32), ("d_skip_track_mask", C.c_void_p), ("d_request_policies", C.c_void_p),
    ]

    def __init__(self, **kw):
        super().__init__()
        _lib().aclb200_default_options(C.byref(self))
        for key, value in kw.items():
            if key == "constant_defaults":
                for i, v in enumerate(np.asarray(value, dtype=np.float32).reshape(12)):
                    self.constant_defaults[i] = float(v)
            elif key == "default_modes":
                self.default_rotation_mode, self.default_translation_mode, self.default_scale_mode = value
            else:
                if not hasattr(self, key):
                    raise AttributeError(key)
                setattr(self, key, value)

    @property
    def bone_bytes(self) -> int:
        return 48 if self.output_layout == LAYOUT_QVV48 else 40


class _ClipsetInfo(C.Structure):
    _fields_ = [("num_clips", C.c_uint32), ("track_type", C.c_uint32), ("max_tracks", C.c_uint32), ("min_tracks", C.c_uint32),
                ("blob_bytes", C.c_uint64), ("index_bytes", C.c_uint64)]


class _ClipInfo(C.Structure):
    _fields_ = [("num_tracks", C.c_uint32), ("num_samples", C.c_uint32), ("sample_rate", C.c_float), ("duration", C.c_float),
                ("num_segments", C.c_uint32), ("looping_policy", C.c_uint32), ("hash", C.c_uint32), ("size", C.c_uint32)]


_lib_handle = None


def library_path() -> str:
    return _LIB_PATH


def _lib():
    global _lib_handle
    if _lib_handle is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(acl_b200/csrc/build.sh). There is no CPU fallback.")
        l = C.CDLL(_LIB_PATH)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        l.aclb200_version_string.restype = C.c_char_p
        l.aclb200_status_string.restype = C.c_char_p
        l.aclb200_status_string.argtypes = [C.c_int]
        l.aclb200_default_options.argtypes = [C.POINTER(Options)]
        l.aclb200_create.argtypes = [C.c_int, C.POINTER(vp)]
        l.aclb200_destroy.argtypes = [vp]
        l.aclb200_last_error.argtypes = [vp]
        l.aclb200_last_error.restype = C.c_char_p
        l.aclb200_upload_clips.argtypes = [vp, vp, vp, u32, u32, C.POINTER(vp), C.POINTER(u32)]
        l.aclb200_upload_clips_packed.argtypes = [vp, vp, vp, vp, u32, u32, C.POINTER(vp), C.POINTER(u32)]
        l.aclb200_release_clipset.argtypes = [vp, vp]
        l.aclb200_clipset_get_info.argtypes = [vp, C.POINTER(_ClipsetInfo)]
        l.aclb200_clipset_get_clip_info.argtypes = [vp, u32, C.POINTER(_ClipInfo)]
        l.aclb200_decompress_tracks.argtypes = [vp, vp, vp, u32, C.POINTER(Options), vp, vp]
        l.aclb200_decompress_track.argtypes = [vp, vp, vp, vp, u32, C.POINTER(Options), vp, vp]
        l.aclb200_scalar_decompress_tracks.argtypes = [vp, vp, vp, u32, C.POINTER(Options), vp, vp]
        l.aclb200_scalar_decompress_track.argtypes = [vp, vp, vp, vp, u32, C.POINTER(Options), vp, vp]
        l.aclb200_decompress_tracks_host.argtypes = [vp, vp, vp, u32, C.POINTER(Options), vp, C.c_size_t]
        l.aclb200_debug_seek.argtypes = [vp, vp, vp, u32, C.POINTER(Options), vp, vp]
        l.aclb200_debug_unpack.argtypes = [vp, vp, vp, u32, C.POINTER(Options), u32, u32, vp, vp]
        l.aclb200_debug_set_trace.argtypes = [vp, vp, u32, u32]
        l.aclb200_device_malloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        l.aclb200_device_free.argtypes = [vp, vp]
        l.aclb200_device_free.restype = None
        l.aclb200_copy_to_device.argtypes = [vp, vp, vp, C.c_size_t]
        l.aclb200_copy_to_host.argtypes = [vp, vp, vp, C.c_size_t]
        l.aclb200_calculate_compression_error.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp, vp, C.POINTER(Options), vp, vp, vp]
        l.aclb200_set_error_chunk_bytes.argtypes = [vp, u64]
        l.aclb200_decompress_all_samples.argtypes = [vp, vp, vp, u32, C.POINTER(Options), vp, vp]
        l.aclb200_local_to_object_space.argtypes = [vp, vp, vp, u64, u32, u64, vp, vp, vp]
        l.aclb200_launch_count.argtypes = [vp]
        l.aclb200_launch_count.restype = u64
        _lib_handle = l
    return _lib_handle


def exported_symbols() -> list[str]:
    """Every function include/aclb200.h declares (used by the CPU-side symbol test)."""
    return [
        "aclb200_version_string", "aclb200_status_string", "aclb200_default_options", "aclb200_create", "aclb200_destroy",
        "aclb200_last_error", "aclb200_upload_clips", "aclb200_upload_clips_packed", "aclb200_release_clipset",
        "aclb200_clipset_get_info", "aclb200_clipset_get_clip_info", "aclb200_decompress_tracks", "aclb200_decompress_track",
        "aclb200_scalar_decompress_tracks", "aclb200_scalar_decompress_track", "aclb200_decompress_tracks_host",
        "aclb200_debug_seek", "aclb200_debug_unpack", "aclb200_debug_set_trace", "aclb200_launch_count",
        "aclb200_device_malloc", "aclb200_device_free", "aclb200_copy_to_device", "aclb200_copy_to_host",
        "aclb200_calculate_compression_error", "aclb200_set_error_chunk_bytes", "aclb200_local_to_object_space",
        "aclb200_decompress_all_samples",
    ]


def make_requests(clips, times) -> np.ndarray:
    """(clip index, sample time) arrays -> aclb200_request[]"""
    clips = np.asarray(clips, dtype=np.uint32)
    times = np.asarray(times, dtype=np.float32)
    out = np.empty(clips.shape[0], dtype=REQUEST_DTYPE)
    out["clip"] = clips
    out["sample_time"] = times
    return out


def _device_ptr(x) -> int:
    """Device pointer of a torch CUDA tensor, or an int passed through."""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    return x.data_ptr()


def _stream_ptr(stream) -> int:
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


class ClipSet:
    def __init__(self, context: "Context", handle: int):
        self._context = context
        self._handle = handle
        info = _ClipsetInfo()
        _lib().aclb200_clipset_get_info(handle, C.byref(info))
        self.num_clips = info.num_clips
        self.track_type = info.track_type
        self.max_tracks = info.max_tracks
        self.min_tracks = info.min_tracks
        self.blob_bytes = info.blob_bytes
        self.index_bytes = info.index_bytes

    @property
    def components(self) -> int:
        """floats per scalar track sample"""
        return self.track_type + 1 if self.track_type <= 3 else 4

    def clip_info(self, clip: int) -> _ClipInfo:
        info = _ClipInfo()
        if _lib().aclb200_clipset_get_clip_info(self._handle, clip, C.byref(info)) != 0:
            raise IndexError(clip)
        return info

    def release(self) -> None:
        if self._handle:
            _lib().aclb200_release_clipset(self._context._handle, self._handle)
            self._handle = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Context:
    """aclb200_context: one per (thread, device)."""

    def __init__(self, device: int = 0):
        handle = C.c_void_p()
        status = _lib().aclb200_create(device, C.byref(handle))
        if status != 0:
            raise AclB200Error(status, _lib().aclb200_status_string(status).decode())
        self._handle = handle.value
        self.device = device

    def close(self) -> None:
        if self._handle:
            _lib().aclb200_destroy(self._handle)
            self._handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status: int) -> None:
        if status != 0:
            raise AclB200Error(status, _lib().aclb200_last_error(self._handle).decode())

    @property
    def launch_count(self) -> int:
        return int(_lib().aclb200_launch_count(self._handle))

    # ---- upload (decompression_context::initialize for many clips) ----
    def upload(self, blobs: list[np.ndarray], check_hash: bool = False) -> ClipSet:
        n = len(blobs)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in blobs])
        sizes = np.array([b.size for b in blobs], dtype=np.uint32)
        handle, failed = C.c_void_p(), C.c_uint32(0xFFFFFFFF)
        status = _lib().aclb200_upload_clips(self._handle, C.cast(ptrs, C.c_void_p), sizes.ctypes.data, n, int(check_hash),
                                             C.byref(handle), C.byref(failed))
        self._check(status)
        return ClipSet(self, handle.value)

    def upload_packed(self, buffer: np.ndarray, offsets: np.ndarray, sizes: np.ndarray, check_hash: bool = False) -> ClipSet:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        handle, failed = C.c_void_p(), C.c_uint32(0xFFFFFFFF)
        status = _lib().aclb200_upload_clips_packed(self._handle, buffer.ctypes.data, offsets.ctypes.data, sizes.ctypes.data,
                                                    sizes.size, int(check_hash), C.byref(handle), C.byref(failed))
        self._check(status)
        return ClipSet(self, handle.value)

    # ---- device entry points: every pointer is a torch CUDA tensor (or a raw device address) ----
    def decompress_tracks(self, clipset: ClipSet, d_requests, num_requests: int, options: Options, d_out, stream=None) -> None:
        self._check(_lib().aclb200_decompress_tracks(self._handle, clipset._handle, _device_ptr(d_requests), num_requests,
                                                     C.byref(options), _device_ptr(d_out), _stream_ptr(stream)))

    def decompress_track(self, clipset: ClipSet, d_requests, d_track_indices, num_requests: int, options: Options, d_out, stream=None) -> None:
        self._check(_lib().aclb200_decompress_track(self._handle, clipset._handle, _device_ptr(d_requests), _device_ptr(d_track_indices),
                                                    num_requests, C.byref(options), _device_ptr(d_out), _stream_ptr(stream)))

    def scalar_decompress_tracks(self, clipset: ClipSet, d_requests, num_requests: int, options: Options, d_out, stream=None) -> None:
        self._check(_lib().aclb200_scalar_decompress_tracks(self._handle, clipset._handle, _device_ptr(d_requests), num_requests,
                                                            C.byref(options), _device_ptr(d_out), _stream_ptr(stream)))

    def scalar_decompress_track(self, clipset: ClipSet, d_requests, d_track_indices, num_requests: int, options: Options, d_out, stream=None) -> None:
        self._check(_lib().aclb200_scalar_decompress_track(self._handle, clipset._handle, _device_ptr(d_requests), _device_ptr(d_track_indices),
                                                           num_requests, C.byref(options), _device_ptr(d_out), _stream_ptr(stream)))

    def debug_seek(self, clipset: ClipSet, d_requests, num_requests: int, options: Options, d_out, stream=None) -> None:
        self._check(_lib().aclb200_debug_seek(self._handle, clipset._handle, _device_ptr(d_requests), num_requests, C.byref(options),
                                              _device_ptr(d_out), _stream_ptr(stream)))

    def debug_unpack(self, clipset: ClipSet, d_requests, num_requests: int, options: Options, which: int, max_sub_tracks: int, d_out, stream=None) -> None:
        self._check(_lib().aclb200_debug_unpack(self._handle, clipset._handle, _device_ptr(d_requests), num_requests, C.byref(options),
                                                which, max_sub_tracks, _device_ptr(d_out), _stream_ptr(stream)))

    def debug_set_trace(self, d_trace, num_blocks: int, num_iterations: int) -> None:
        self._check(_lib().aclb200_debug_set_trace(self._handle, _device_ptr(d_trace), num_blocks, num_iterations))

    # ---- compression error measurement and the object space walk, poses stay on the device ----
    def calculate_compression_error(self, clipset: ClipSet, jobs: np.ndarray, d_raw_poses, d_parent_indices, d_shell_distances,
                                    options: Options, d_out_errors, d_output_indices=None, d_out_error_matrix=None, d_base_poses=None,
                                    stream=None) -> None:
        jobs = np.ascontiguousarray(jobs)
        assert jobs.dtype == ERROR_JOB_DTYPE
        self._check(_lib().aclb200_calculate_compression_error(
            self._handle, clipset._handle, jobs.ctypes.data, jobs.shape[0], _device_ptr(d_raw_poses), _device_ptr(d_parent_indices),
            _device_ptr(d_shell_distances), _device_ptr(d_output_indices), _device_ptr(d_base_poses), C.byref(options), _device_ptr(d_out_errors),
            _device_ptr(d_out_error_matrix), _stream_ptr(stream)))

    def decompress_all_samples(self, clipset: ClipSet, jobs: np.ndarray, options: Options, d_out, stream=None) -> None:
        jobs = np.ascontiguousarray(jobs)
        assert jobs.dtype == ERROR_JOB_DTYPE
        self._check(_lib().aclb200_decompress_all_samples(self._handle, clipset._handle, jobs.ctypes.data, jobs.shape[0], C.byref(options),
                                                          _device_ptr(d_out), _stream_ptr(stream)))

    def set_error_chunk_bytes(self, num_bytes: int) -> None:
        self._check(_lib().aclb200_set_error_chunk_bytes(self._handle, num_bytes))

    def local_to_object_space(self, d_local_poses, d_object_poses, num_poses: int, num_tracks: int, d_parent_indices,
                              pose_stride_bytes: int = 0, d_out_flags=None, stream=None) -> None:
        self._check(_lib().aclb200_local_to_object_space(self._handle, _device_ptr(d_local_poses), _device_ptr(d_object_poses), num_poses,
                                                         num_tracks, pose_stride_bytes, _device_ptr(d_parent_indices),
                                                         _device_ptr(d_out_flags), _stream_ptr(stream)))

    # ---- host buffers in, host buffers out (the call the C++ header shim uses) ----
    def decompress_tracks_host(self, clipset: ClipSet, requests: np.ndarray, options: Options, out: np.ndarray) -> np.ndarray:
        requests = np.ascontiguousarray(requests)
        assert requests.dtype == REQUEST_DTYPE
        self._check(_lib().aclb200_decompress_tracks_host(self._handle, clipset._handle, requests.ctypes.data, requests.shape[0],
                                                          C.byref(options), out.ctypes.data, out.nbytes))
        return out
