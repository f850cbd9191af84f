"""acl_b200 -- B200-native (sm_100a) batched decompression of nfrechette/acl `compressed_tracks`.

The product is the C-ABI shared library `acl_b200/libaclb200.so` (sources in acl_b200/csrc/, interface in
include/aclb200.h, C++ header shim in include/acl_b200/decompress.h). This Python package is only the thin
ctypes binding the tests and bench.py drive it through; it holds no decode logic and has no CPU fallback:
importing it without the built library, or creating a Context without a CUDA device, raises.
"""
from .api import (  # noqa: F401
    AclB200Error, Context, ClipSet, Options, library_path, make_requests,
    ROUND_NONE, ROUND_FLOOR, ROUND_CEIL, ROUND_NEAREST, ROUND_PER_TRACK,
    LOOP_CLAMP, LOOP_WRAP, LOOP_AS_COMPRESSED,
    NORMALIZE_NEVER, NORMALIZE_LERP_ONLY, NORMALIZE_ALWAYS,
    DEFAULT_SKIPPED, DEFAULT_CONSTANT, DEFAULT_VARIABLE, DEFAULT_LEGACY,
    LAYOUT_QVV48, LAYOUT_QVV40, MATH_EXACT, MATH_FAST, TRACK_QVVF,
    SKIP_ROTATION, SKIP_TRANSLATION, SKIP_SCALE,
    ERROR_JOB_DTYPE, TRACK_ERROR_DTYPE, ERROR_FLAG_NEGATIVE_SCALE, ERROR_FLAG_INVALID_SKELETON,
)
