"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what include/aclb200.h declares,
refuses to run without a GPU (no CPU fallback) and never reaches into oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "aclb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aclb200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from acl_b200 import api
    lib = api._lib()
    declared = _header_functions()
    assert declared, "no functions found in include/aclb200.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/aclb200.h but not exported by libaclb200.so"
    assert sorted(api.exported_symbols()) == declared


def test_default_options_match_reference_defaults():
    import acl_b200 as ab
    o = ab.Options()
    assert o.struct_size == C.sizeof(ab.Options)
    # default_transform_decompression_settings (decompression_settings.h:211-232) + track_writer defaults (track_writer.h:170-186)
    assert (o.normalization, o.per_track_rounding, o.wrapping, o.clamp_sample_time) == (ab.NORMALIZE_LERP_ONLY, 0, 1, 1)
    assert (o.default_rotation_mode, o.default_translation_mode, o.default_scale_mode) == (ab.DEFAULT_CONSTANT, ab.DEFAULT_CONSTANT, ab.DEFAULT_LEGACY)
    assert list(o.constant_defaults) == [0, 0, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0]
    assert (o.rounding_policy, o.looping_policy, o.output_layout, o.math_mode) == (ab.ROUND_NONE, ab.LOOP_AS_COMPRESSED, ab.LAYOUT_QVV48, ab.MATH_EXACT)


def test_no_cpu_fallback():
    import torch
    import acl_b200 as ab
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ab.AclB200Error) as err:
        ab.Context(0)
    assert err.value.status == 4        # ACLB200_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    for folder in ("acl_b200", "include"):
        for base, _, files in os.walk(os.path.join(ROOT, folder)):
            for name in files:
                if name.endswith((".py", ".h", ".hpp", ".cpp", ".cu", ".sh")):
                    text = open(os.path.join(base, name), errors="ignore").read()
                    assert "oracle" not in text.lower().replace("the oracle", "").replace("test oracle", ""), f"{base}/{name} mentions the oracle"


def test_request_and_seek_state_layouts():
    from acl_b200 import api
    assert api.REQUEST_DTYPE.itemsize == 8
    assert api.SEEK_STATE_DTYPE.itemsize == 56
