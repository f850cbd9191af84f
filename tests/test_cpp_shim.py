"""include/acl_b200/decompress.h: the C++ header shim that keeps the reference's decompression_context call sequence
(includes/acl/decompression/decompress.h:90-172). tests/cpp/shim_decode.cpp is written like a reference call site:
initialize -> seek -> decompress_tracks(writer)."""
import os
import subprocess

import numpy as np
import pytest

from . import clips

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_shim_program(tmp_path, name="shim_decode", cuda_runtime=False):
    exe = str(tmp_path / name)
    lib_dir = os.path.join(ROOT, "acl_b200")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
           "-L" + lib_dir, "-laclb200", "-Wl,-rpath," + lib_dir]
    if cuda_runtime:
        cmd += ["-I/usr/local/cuda/include", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_shim_compiles_and_has_no_cpu_fallback(tmp_path):
    """The shim is plain C++14 over the C ABI; without a GPU the program must fail with NO_DEVICE, not decode on the CPU."""
    import torch
    exe = build_shim_program(tmp_path)
    result = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "c1_30bones.acl.bin"), "default", "0.1"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert result.returncode == 0
    else:
        assert result.returncode == 3, (result.returncode, result.stderr)
        assert result.stdout == ""


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_30bones", "mixed_scale", "single_segment", "stripped_loop", "full_formats", "ragged_17"])
def test_shim_decode_matches_oracle(tmp_path, oracle_port, name):
    """decompression_context<default settings>: seek + decompress_tracks through a track_writer, bit-exact vs the oracle."""
    exe = build_shim_program(tmp_path)
    blob = clips.load_blob(name)
    times = [float(t) for t in clips.sample_times(clips.TRANSFORM_SPECS[name])[::2]]
    # the default settings only support the variable formats (decompression_settings.h:211-232), like the reference's: the full
    # precision clip goes through the debug settings (oracle kind 1)
    debug = name == "full_formats"
    result = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", name + ".acl.bin"), "debug" if debug else "default"] + [repr(t) for t in times],
                            capture_output=True, text=True, check=True)
    rows = [line.split() for line in result.stdout.strip().splitlines()]
    settings = oracle_port.settings_for_kind(1 if debug else 0)
    num_tracks = max(int(r[1]) for r in rows) + 1
    got = np.zeros((len(times), num_tracks, 12), np.uint32)
    for r in rows:
        got[int(r[0]), int(r[1])] = [int(w, 16) for w in r[2:]]
    for i, t in enumerate(times):
        expected = oracle_port.transform_decompress_tracks(blob, settings, np.float32(t), 0)
        assert np.array_equal(expected[:, clips.DEFINED_LANES].view(np.uint32), got[i][:, clips.DEFINED_LANES]), (name, t)


def test_batch_shim_compiles(tmp_path):
    build_shim_program(tmp_path, "shim_batch", cuda_runtime=True)


@pytest.mark.gpu
def test_batch_context_matches_oracle(tmp_path, oracle_port):
    """batch_context<default settings>: bind several clips, one launch for a ragged request list in device memory."""
    exe = build_shim_program(tmp_path, "shim_batch", cuda_runtime=True)
    names = ["c1_30bones", "ragged_17", "mixed_scale", "one_bone"]
    blobs = [clips.load_blob(n) for n in names]
    specs = [clips.TRANSFORM_SPECS[n] for n in names]
    rng = np.random.default_rng(3)
    req_clip = rng.integers(0, len(names), 24)
    req_time = [float(np.float32(rng.uniform(-0.1, (specs[c].num_samples - 1) / specs[c].sample_rate + 0.1))) for c in req_clip]
    args = [exe, str(len(names))] + [os.path.join(ROOT, "tests", "golden", n + ".acl.bin") for n in names]
    for c, t in zip(req_clip, req_time):
        args += [str(int(c)), repr(t)]
    result = subprocess.run(args, capture_output=True, text=True, check=True)
    max_tracks = max(s.num_tracks for s in specs)
    got = np.zeros((len(req_clip), max_tracks, 12), np.uint32)
    for line in result.stdout.strip().splitlines():
        r = line.split()
        got[int(r[0]), int(r[1])] = [int(w, 16) for w in r[2:]]
    settings = oracle_port.settings_for_kind(0)
    for i, (c, t) in enumerate(zip(req_clip, req_time)):
        expected = oracle_port.transform_decompress_tracks(blobs[c], settings, np.float32(t), 0)
        n = expected.shape[0]
        assert np.array_equal(expected[:, clips.DEFINED_LANES].view(np.uint32), got[i][:n][:, clips.DEFINED_LANES]), (names[c], t)
        assert not got[i][n:].any()       # bones past the clip's own track count are left untouched


REFERENCE_INCLUDES = ["/root/reference/includes", "/root/reference/external/rtm/includes"]
CALLSITE_EXE = os.path.join(ROOT, "tests", "cpp", "_build", "shim_reference_callsite")


TRACK_ERROR_EXE = os.path.join(ROOT, "tests", "cpp", "_build", "shim_track_error")


def build_reference_callsite():
    """tests/cpp/shim_reference_callsite.cpp and shim_track_error.cpp need the reference's headers: they are built where /root/reference
    exists (__graft_entry__.build() does it too) and travel to the GPU box prebuilt, like oracle/_ref."""
    os.makedirs(os.path.dirname(CALLSITE_EXE), exist_ok=True)
    for exe in (CALLSITE_EXE, TRACK_ERROR_EXE):
        cmd = ["g++", "-std=c++14", "-O2", "-msse4.1", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror"] + ["-I" + d for d in REFERENCE_INCLUDES] + [
            "-o", exe, os.path.join(ROOT, "tests", "cpp", os.path.basename(exe) + ".cpp"),
            "-L" + os.path.join(ROOT, "acl_b200"), "-laclb200", "-Wl,-rpath,$ORIGIN/../../../acl_b200"]
        subprocess.run(cmd, check=True, capture_output=True, text=True)


@pytest.mark.skipif(not all(os.path.isdir(d) for d in REFERENCE_INCLUDES), reason="needs the reference's headers")
def test_reference_callsite_compiles_against_the_shim():
    """The reference's own benchmark loop body (benchmark.cpp:246-258) with acl::debug_track_writer, acl::compressed_tracks, settings
    derived from acl::default_transform_decompression_settings: instantiated with acl::decompression_context and with
    acl_b200::decompression_context. Without a GPU the program must stop with NO_DEVICE (no CPU fallback)."""
    import torch
    build_reference_callsite()
    result = subprocess.run([CALLSITE_EXE, os.path.join(ROOT, "tests", "golden", "c1_30bones.acl.bin")], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert result.returncode == 3, (result.returncode, result.stdout, result.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_30bones", "c2_100bones", "mixed_scale", "stripped_loop", "float1", "float3", "vector4"])
def test_reference_callsite_matches_the_reference(name):
    """Same call site, both classes, on the GPU box (prebuilt binary): decompress_tracks bit-identical, decompress_track rotations
    within 1e-5, relocated / is_bound_to / initialize answer like the reference."""
    if not os.path.exists(CALLSITE_EXE):
        pytest.skip("tests/cpp/_build/shim_reference_callsite was not built (needs /root/reference at build time)")
    result = subprocess.run([CALLSITE_EXE, os.path.join(ROOT, "tests", "golden", name + ".acl.bin"), os.path.join(ROOT, "tests", "golden", "ragged_17.acl.bin")],
                            capture_output=True, text=True)
    assert result.returncode == 0 and "PASS" in result.stdout, (result.stdout, result.stderr)


def test_binding_shim_compiles(tmp_path):
    build_shim_program(tmp_path, "shim_binding")


@pytest.mark.gpu
def test_binding_semantics_follow_the_reference(tmp_path):
    """relocated() compares hashes, is_bound_to() address + hash, failed initialize leaves the context unbound."""
    exe = build_shim_program(tmp_path, "shim_binding")
    golden = os.path.join(ROOT, "tests", "golden")
    result = subprocess.run([exe, os.path.join(golden, "c1_30bones.acl.bin"), os.path.join(golden, "mixed_scale.acl.bin")], capture_output=True, text=True)
    assert result.returncode == 0 and "PASS" in result.stdout, (result.stdout, result.stderr)


@pytest.mark.skipif(not all(os.path.isdir(d) for d in REFERENCE_INCLUDES), reason="needs the reference's headers")
def test_track_error_callsite_compiles_against_the_shim():
    """acl_b200/track_error.h: calculate_compression_error(allocator, raw_tracks, context[, error_metric]) with the reference's signature
    and types (compression/track_error.h:64-91). Without a GPU the program must stop with NO_DEVICE."""
    import torch
    build_reference_callsite()
    result = subprocess.run([TRACK_ERROR_EXE], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert result.returncode == 3, (result.returncode, result.stdout, result.stderr)


@pytest.mark.gpu
def test_track_error_callsite_matches_the_reference():
    """Raw clips compressed by the reference's compressor, measured by acl::calculate_compression_error with acl::decompression_context
    and by acl_b200::calculate_compression_error with acl_b200::decompression_context (bind pose that is not the identity, scale, a
    stripped track, full precision formats, stripped key frames, scalar float3f): same worst track and sample time, error within 5e-5
    (scalar: exact)."""
    if not os.path.exists(TRACK_ERROR_EXE):
        pytest.skip("tests/cpp/_build/shim_track_error was not built (needs /root/reference at build time)")
    result = subprocess.run([TRACK_ERROR_EXE], capture_output=True, text=True)
    assert result.returncode == 0 and "PASS" in result.stdout, (result.stdout, result.stderr)
