"""Known-answer vectors the reference's own unit tests hold for this path, replayed against the oracle.

tests/sources/core/test_interpolation_utils.cpp:225-331 -- find_linear_interpolation_samples_with_sample_rate, the function seek_v0
calls (decompression.transform.h:218-226): (num_samples, sample_rate, sample_time, rounding, looping) -> (key0, key1, alpha +- 1e-6)."""
import numpy as np
import pytest

NONE, FLOOR, CEIL, NEAREST = 0, 1, 2, 3
CLAMP, WRAP = 0, 1
f = np.float32

# the table of test_interpolation_utils.cpp, transcribed as data: sample times are float32 expressions exactly as written there
KNOWN_ANSWERS = [
    # clamped looping policy, 31 samples at 30 Hz (:233-275)
    (31, 30.0, f(0.0), NONE, CLAMP, 0, 1, 0.0),
    (31, 30.0, f(1.0) / f(30.0), NONE, CLAMP, 1, 2, 0.0),
    (31, 30.0, f(2.5) / f(30.0), NONE, CLAMP, 2, 3, 0.5),
    (31, 30.0, f(1.0), NONE, CLAMP, 30, 30, 0.0),
    (31, 30.0, f(2.5) / f(30.0), FLOOR, CLAMP, 2, 3, 0.0),
    (31, 30.0, f(2.5) / f(30.0), CEIL, CLAMP, 2, 3, 1.0),
    (31, 30.0, f(2.4) / f(30.0), NEAREST, CLAMP, 2, 3, 0.0),
    (31, 30.0, f(2.6) / f(30.0), NEAREST, CLAMP, 2, 3, 1.0),
    # wrapping looping policy, 30 samples at 30 Hz (:287-329)
    (30, 30.0, f(0.0), NONE, WRAP, 0, 1, 0.0),
    (30, 30.0, f(1.0) / f(30.0), NONE, WRAP, 1, 2, 0.0),
    (30, 30.0, f(2.5) / f(30.0), NONE, WRAP, 2, 3, 0.5),
    (30, 30.0, f(1.0), NONE, WRAP, 0, 0, 0.0),
    (30, 30.0, f(2.5) / f(30.0), FLOOR, WRAP, 2, 3, 0.0),
    (30, 30.0, f(2.5) / f(30.0), CEIL, WRAP, 2, 3, 1.0),
    (30, 30.0, f(2.4) / f(30.0), NEAREST, WRAP, 2, 3, 0.0),
    (30, 30.0, f(2.6) / f(30.0), NEAREST, WRAP, 2, 3, 1.0),
]


@pytest.mark.parametrize("case", KNOWN_ANSWERS)
def test_find_key_frames_known_answers(oracle_port, case):
    num_samples, sample_rate, sample_time, rounding, looping, key0, key1, alpha = case
    got = oracle_port.find_key_frames(num_samples, sample_rate, float(sample_time), rounding, looping)
    assert got[0] == key0 and got[1] == key1
    assert abs(got[2] - alpha) <= 1e-6      # the reference's own error_threshold (:35)
