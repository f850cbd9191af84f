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


# external/rtm/tests/sources/test_qvv.cpp:225-257 -- rtm::qvv_mul_point3 and rtm::qvv_mul, the two functions the error metric is built from
# (transform_error_metrics.h:289-358), with the test's own transforms and expected points. quat_from_euler (quatf.h:1439-1456) of
# (0, 90 deg, 0) is a quarter turn about Z, of (0, 0, 90 deg) a quarter turn about -X.
def _qvv(rotation, translation, scale):
    return np.array(list(rotation) + list(translation) + [0.0] + list(scale) + [0.0], dtype=np.float32)


def test_qvv_mul_and_mul_point3_known_answers(oracle_port):
    import ctypes as C
    lib = oracle_port.lib()
    lib.aclo_test_qvv_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.aclo_test_qvv_mul_point3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]

    def mul_point3(point, qvv):
        point = np.array(point, dtype=np.float32)
        out = np.zeros(3, dtype=np.float32)
        lib.aclo_test_qvv_mul_point3(point.ctypes.data, qvv.ctypes.data, out.ctypes.data)
        return out

    def mul(lhs, rhs):
        out = np.zeros(12, dtype=np.float32)
        lib.aclo_test_qvv_mul(lhs.ctypes.data, rhs.ctypes.data, 1, out.ctypes.data)
        return out

    threshold = 1.0e-4          # test_qvv.cpp: `const FloatType threshold = FloatType(1.0E-4)` for the float flavour
    h = np.float32(np.sin(np.pi / 4))
    x_axis, y_axis = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0)
    transform_a = _qvv((0.0, 0.0, h, h), x_axis, (1.2, 1.2, 1.2))        # rotation_around_z, translation x, scale 1.2
    transform_b = _qvv((-h, 0.0, 0.0, h), y_axis, (1.2, 1.2, 1.2))       # rotation_around_x, translation y
    near = lambda got, want: np.all(np.abs(got - np.array(want, dtype=np.float32)) <= threshold)
    assert near(mul_point3(x_axis, transform_a), (1.0, 1.2, 0.0))          # :232-233
    assert near(mul_point3(y_axis, transform_a), (-0.2, 0.0, 0.0))         # :234-235
    assert near(mul_point3(x_axis, transform_b), (1.2, 1.0, 0.0))          # :239-240
    assert near(mul_point3(y_axis, transform_b), (0.0, 1.0, -1.2))         # :241-242
    transform_ab, transform_ba = mul(transform_a, transform_b), mul(transform_b, transform_a)      # :244-245
    assert near(mul_point3(x_axis, transform_ab), (1.2, 1.0, -1.44))       # :246-247
    assert near(mul_point3(x_axis, transform_ab), mul_point3(mul_point3(x_axis, transform_a), transform_b))     # :248
    assert near(mul_point3(y_axis, transform_ab), (-0.24, 1.0, 0.0))       # :249-250
    assert near(mul_point3(y_axis, transform_ab), mul_point3(mul_point3(y_axis, transform_a), transform_b))     # :251
    assert near(mul_point3(x_axis, transform_ba), (-0.2, 1.44, 0.0))       # :252-253
    assert near(mul_point3(x_axis, transform_ba), mul_point3(mul_point3(x_axis, transform_b), transform_a))     # :254
    assert near(mul_point3(y_axis, transform_ba), (-0.2, 0.0, -1.44))      # :255-256
    assert near(mul_point3(y_axis, transform_ba), mul_point3(mul_point3(y_axis, transform_b), transform_a))     # :257
