/* oracle/acl_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement ("port") of the reference's uniformly-sampled decompression path
 * (nfrechette/acl @ 0f855f0, includes/acl/decompression/...). It exists to check the CUDA product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it. The product
 * (acl_b200/, include/, libaclb200.so) never links, imports or executes anything in oracle/.
 *
 * Parity is pinned: tests/test_oracle_vs_reference.py compares this port bit-for-bit with the
 * reference itself (oracle/_ref/libaclref.so, built from /root/reference) and with the committed
 * golden vectors in tests/golden/ that the reference generated (tests/golden/make_golden.py).
 */
#ifndef ACL_ORACLE_H
#define ACL_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* acl::sample_rounding_policy, core/sample_rounding_policy.h:47-107 */
enum { ACLO_ROUND_NONE = 0, ACLO_ROUND_FLOOR = 1, ACLO_ROUND_CEIL = 2, ACLO_ROUND_NEAREST = 3, ACLO_ROUND_PER_TRACK = 4 };
/* acl::sample_looping_policy, core/sample_looping_policy.h:56-82 */
enum { ACLO_LOOP_CLAMP = 0, ACLO_LOOP_WRAP = 1, ACLO_LOOP_AS_COMPRESSED = 2 };
/* acl::rotation_normalization_policy_t, decompression/decompression_settings.h:52-62 */
enum { ACLO_NORMALIZE_NEVER = 0, ACLO_NORMALIZE_LERP_ONLY = 1, ACLO_NORMALIZE_ALWAYS = 2 };
/* acl::default_sub_track_mode, core/track_writer.h:49-74 */
enum { ACLO_DEFAULT_SKIPPED = 0, ACLO_DEFAULT_CONSTANT = 1, ACLO_DEFAULT_VARIABLE = 2, ACLO_DEFAULT_LEGACY = 3 };

/* What the reference bakes into `decompression_settings` + `track_writer` at compile time. */
typedef struct aclo_settings
{
	uint32_t normalization;				/* ACLO_NORMALIZE_* : get_rotation_normalization_policy() */
	uint32_t per_track_rounding;		/* is_per_track_rounding_supported() */
	uint32_t wrapping;					/* is_wrapping_supported() */
	uint32_t clamp_sample_time;			/* clamp_sample_time() */
	uint32_t multiple_rotation_formats;	/* num_supported_rotation_formats() > 1, decompression_context.transform.h:145-151 */
	uint32_t default_rotation_mode;		/* ACLO_DEFAULT_* (legacy is not valid for rotation / translation) */
	uint32_t default_translation_mode;
	uint32_t default_scale_mode;
	float    constant_defaults[12];		/* rotation xyzw, translation xyz -, scale xyz - */
	const float* variable_defaults;		/* [num_tracks][12] or NULL */
	const uint8_t* per_track_rounding_policies;	/* [num_tracks] or NULL: writer.get_rounding_policy() when seek used per_track */
} aclo_settings;

/* Result of seek(), mirrors persistent_transform_decompression_context_v0
 * (decompression_context.transform.h:53-116) with offsets instead of pointers. */
typedef struct aclo_seek_state
{
	float    sample_time;				/* < 0 => no valid seek */
	float    interpolation_alpha;
	float    clip_duration;
	uint32_t looping_policy;
	uint32_t rounding_policy;
	uint32_t key_frames[2];				/* clip relative, after stripped key frame remapping */
	uint32_t segment_indices[2];
	uint32_t segment_key_frames[2];		/* index of the stored key frame inside its segment */
	uint32_t key_frame_bit_offsets[2];
	uint32_t segment_offsets[2];		/* byte offset of the segment header from the blob start */
	uint32_t format_offsets[2];			/* byte offsets from the blob start */
	uint32_t range_offsets[2];
	uint32_t animated_offsets[2];
	uint32_t uses_single_segment;
} aclo_seek_state;

/* Fills `settings` with default_transform_decompression_settings + track_writer defaults
 * (decompression_settings.h:211-232, track_writer.h:170-186). */
void aclo_default_settings(aclo_settings* settings);

/* 0 when the buffer would be accepted by decompression_context::initialize
 * (decompress.impl.h:66-83 -> compressed_tracks::is_valid, compressed_tracks.impl.h:278-301),
 * negative otherwise. `check_hash` != 0 also verifies the FNV-1a hash (core/hash.h). */
int aclo_validate(const void* blob, size_t size, int check_hash);

uint32_t aclo_hash32(const void* data, size_t size);

/* find_linear_interpolation_samples_with_sample_rate, core/impl/interpolation_utils.impl.h:143-201 */
void aclo_find_key_frames(uint32_t num_samples, float sample_rate, float sample_time, uint32_t rounding_policy, uint32_t looping_policy,
	uint32_t* out_index0, uint32_t* out_index1, float* out_alpha);

/* Transform (qvvf) path. All functions return 0 on success, <0 on invalid input. */
int aclo_transform_seek(const void* blob, const aclo_settings* settings, float sample_time,
	uint32_t rounding_policy, uint32_t looping_policy, aclo_seek_state* out_state);

/* out = [num_tracks][12] floats: rotation xyzw, translation xyz (w untouched), scale xyz (w untouched).
 * Sub-tracks in `skipped` default mode are left untouched. */
int aclo_transform_decompress_tracks(const void* blob, const aclo_settings* settings, const aclo_seek_state* state, float* out);
int aclo_transform_decompress_track(const void* blob, const aclo_settings* settings, const aclo_seek_state* state, uint32_t track_index, float* out);

/* Extracted integers of one animated key frame (bit-exact parity of the format decode).
 * out_ints = [num animated sub-tracks (rot, trans, scale order)][4]: x, y, z (raw quantised integers or
 * raw float bits), num_bits as stored in the per-track format byte (0xFFFFFFFF for full formats). */
int aclo_transform_extract_key_frame(const void* blob, const aclo_seek_state* state, uint32_t which, uint32_t* out_ints);

/* Scalar (float1f..float4f, vector4f) path. */
typedef struct aclo_scalar_seek_state
{
	float    sample_time;
	float    interpolation_alpha;
	float    duration;
	uint32_t looping_policy;
	uint32_t rounding_policy;
	uint32_t key_frames[2];
	uint32_t key_frame_bit_offsets[2];
} aclo_scalar_seek_state;

int aclo_scalar_seek(const void* blob, const aclo_settings* settings, float sample_time,
	uint32_t rounding_policy, uint32_t looping_policy, aclo_scalar_seek_state* out_state);
/* out = [num_tracks][4] floats, only the first N components of each row are written. */
int aclo_scalar_decompress_tracks(const void* blob, const aclo_settings* settings, const aclo_scalar_seek_state* state, float* out);
int aclo_scalar_decompress_track(const void* blob, const aclo_settings* settings, const aclo_scalar_seek_state* state, uint32_t track_index, float* out);

/* The reference's own "bytes touched by one decompress_tracks call" accounting
 * (compress.transform.impl.h:336-338, write_stats.h:115-120), split so a batch can de-duplicate:
 * out[0] clip-constant bytes (headers, sub-track types, constants, clip range),
 * out[1] per-segment metadata bytes (segment header + format + segment range) of ONE segment,
 * out[2] bytes of one animated key frame (ceil(animated_pose_bit_size / 8)) of segment `segment_index`,
 * out[3] output bytes of one pose (num_tracks * 40, tools/acl_decompressor/sources/benchmark.cpp:146-147). */
int aclo_transform_touched_bytes(const void* blob, uint32_t segment_index, uint64_t out[4]);
int aclo_scalar_touched_bytes(const void* blob, uint64_t out[4]);

/* the reference's compression error measurement (compression/impl/track_error.impl.h:166-392) with the
 * qvvf_transform_error_metric (compression/transform_error_metrics.h:281-385), restated over already sampled poses.
 * normalize_mode 0 = rtm::quat_normalize as the reference's SSE2 build runs it (rsqrtss + 2 Newton-Raphson steps: bit-identical to
 * the reference on the CPU both run on; tests pin it there), 1 = IEEE 1 / sqrt (what the CUDA path computes, bit for bit). */
typedef struct aclo_track_error		/* acl::track_error, compression/track_error.h:48-62 */
{
	uint32_t index;
	float    error;
	float    sample_time;
} aclo_track_error;

/* poses are [num_tracks][12] floats (rtm::qvvf). Returns < 0 when a parent does not precede its child, 1 when a negative scale took
 * rtm::qvv_mul through its matrix branch (qvvf.h:320-345, restated as well; informational), else 0. */
int aclo_local_to_object_space(const float* local_pose, const uint32_t* parent_indices, uint32_t num_tracks, int normalize_mode, float* out_object_pose);
float aclo_calculate_error(const float* raw_object_bone, const float* lossy_object_bone, float shell_distance);
int aclo_transform_track_error(const float* raw_poses, const float* lossy_poses, uint32_t num_samples, uint32_t num_tracks,
	float sample_rate, float duration, const uint32_t* parent_indices, const float* shell_distances, int normalize_mode,
	aclo_track_error* out_error, float* out_errors, float* scratch_object_poses /* [4][num_tracks][12] */,
	const float* base_poses /* optional additive base, [num_samples][num_tracks][12] */, uint32_t additive_format,
	uint32_t metric /* 0 qvvf_transform_error_metric, 1 qvvf_matrix3x4f_transform_error_metric (transform_error_metrics.h:389-464; no CPU specific step) */);
/* acl::apply_additive_to_base (core/additive_utils.h:131-167) over a pose, in place on `pose`; format = acl::additive_clip_format8 */
int aclo_apply_additive_to_base(uint32_t additive_format, const float* base_pose, float* pose, uint32_t num_tracks, int normalize_mode);
int aclo_scalar_track_error(const float* raw_values, const float* lossy_values, uint32_t num_samples, uint32_t num_tracks, uint32_t components,
	float sample_rate, float duration, aclo_track_error* out_error);

/* rtm::qvv_mul / rtm::qvv_mul_point3 as restated for the error metric, exposed for the known-answer table of the reference's own math
 * tests (external/rtm/tests/sources/test_qvv.cpp:225-257, replayed by tests/test_reference_known_answers.py) */
void aclo_test_qvv_mul(const float* lhs, const float* rhs, int normalize_mode, float* out);
void aclo_test_qvv_mul_point3(const float* point, const float* qvv, float* out);

/* Single-threaded timing helper for the "port" CPU baseline: decodes `num_requests` (clip, time)
 * requests with default settings and returns the elapsed seconds. */
double aclo_bench_transform(const void* const* blobs, const uint32_t* request_clip, const float* request_time,
	uint32_t num_requests, uint32_t max_tracks);

#ifdef __cplusplus
}
#endif

#endif /* ACL_ORACLE_H */
