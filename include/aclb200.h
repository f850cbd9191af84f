/* include/aclb200.h -- C ABI of libaclb200.so: batched, B200-native (sm_100a) decompression of
 * nfrechette/acl `compressed_tracks` blobs.
 *
 * ACL (reference @ 0f855f0) has no FFI layer of its own: its "operator API" for this path is the
 * header-only C++ class acl::decompression_context<settings> (includes/acl/decompression/decompress.h:76-201)
 * driven by a duck-typed acl::track_writer (includes/acl/core/track_writer.h:82-216). The entry points
 * below are what a binding of that path needs, one (clip, sample_time) request per pose, many requests
 * per call. Each one names the reference interface it replaces. include/acl_b200/decompress.h is the C++
 * header shim that keeps the reference's class/method names on top of these functions, and
 * INTEGRATION.md shows the binding a maintainer would add on the reference side.
 *
 * Conventions: POD only, no C++ types, no exceptions; every function returns an aclb200_status;
 * device pointers are plain `void*` / typed pointers into CUDA device memory of the context's device;
 * `stream` is a `cudaStream_t` passed as `void*` (NULL = the legacy default stream). Calls are
 * asynchronous on `stream` unless stated otherwise. There is NO CPU fallback: without a CUDA device
 * aclb200_create fails with ACLB200_ERR_NO_DEVICE.
 */
#ifndef ACLB200_H
#define ACLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
	#define ACLB200_API __declspec(dllexport)
#else
	#define ACLB200_API __attribute__((visibility("default")))
#endif

#define ACLB200_VERSION_MAJOR 0
#define ACLB200_VERSION_MINOR 3

typedef enum aclb200_status
{
	ACLB200_OK = 0,
	ACLB200_ERR_INVALID_ARGUMENT = 1,
	ACLB200_ERR_INVALID_CLIP = 2,		/* what decompression_context::initialize() reports by returning false (decompress.impl.h:66-83) */
	ACLB200_ERR_UNSUPPORTED = 3,		/* valid ACL data this build refuses: mixed track types in one clip set, clip images past 4 GiB */
	ACLB200_ERR_NO_DEVICE = 4,
	ACLB200_ERR_CUDA = 5,
	ACLB200_ERR_OUT_OF_MEMORY = 6
} aclb200_status;

/* acl::sample_rounding_policy (core/sample_rounding_policy.h:47-107) */
enum { ACLB200_ROUND_NONE = 0, ACLB200_ROUND_FLOOR = 1, ACLB200_ROUND_CEIL = 2, ACLB200_ROUND_NEAREST = 3, ACLB200_ROUND_PER_TRACK = 4 };
/* acl::sample_looping_policy (core/sample_looping_policy.h:56-82) */
enum { ACLB200_LOOP_CLAMP = 0, ACLB200_LOOP_WRAP = 1, ACLB200_LOOP_AS_COMPRESSED = 2 };
/* acl::rotation_normalization_policy_t (decompression/decompression_settings.h:52-62) */
enum { ACLB200_NORMALIZE_NEVER = 0, ACLB200_NORMALIZE_LERP_ONLY = 1, ACLB200_NORMALIZE_ALWAYS = 2 };
/* acl::default_sub_track_mode (core/track_writer.h:49-74) */
enum { ACLB200_DEFAULT_SKIPPED = 0, ACLB200_DEFAULT_CONSTANT = 1, ACLB200_DEFAULT_VARIABLE = 2, ACLB200_DEFAULT_LEGACY = 3 };
/* acl::track_type8 (core/track_types.h:53-68) */
enum { ACLB200_TRACK_FLOAT1F = 0, ACLB200_TRACK_FLOAT2F = 1, ACLB200_TRACK_FLOAT3F = 2, ACLB200_TRACK_FLOAT4F = 3, ACLB200_TRACK_VECTOR4F = 4, ACLB200_TRACK_QVVF = 12 };

/* Output layouts of the transform path (the device-side stand-in for track_writer::write_rotation/translation/scale). */
enum
{
	ACLB200_LAYOUT_QVV48 = 0,	/* rtm::qvvf as debug_track_writer stores it: rotation xyzw, translation xyz + 0, scale xyz + 0 (48 B / bone) */
	ACLB200_LAYOUT_QVV40 = 1	/* rotation xyzw, translation xyz, scale xyz (40 B / bone) == the reference's own "pose size"
								 * (tools/acl_decompressor/sources/benchmark.cpp:146-147) */
};

/* Arithmetic of the float stage. The integer / format decode is bit-exact in both modes. */
enum
{
	ACLB200_MATH_EXACT = 0,		/* IEEE-754 mul/add/sqrt/div in the reference's operation order, never fused: bit-identical
								 * to the reference's SSE2/AVX/scalar builds for decompress_tracks */
	ACLB200_MATH_FAST = 1		/* decompress_tracks on variable bit rate rotations: x, y, z and the W reconstruction input stay exact,
								 * then hardware sqrt / rsqrt and fused multiply-adds: rotations <= 1e-5 absolute from EXACT (measured
								 * < 2e-6, including W ~ 0), translations / scales / every other path unchanged (bit-exact) */
};

typedef struct aclb200_context aclb200_context;
typedef struct aclb200_clipset aclb200_clipset;

/* One decompression request == one `context.initialize(clip); context.seek(sample_time, policy);
 * context.decompress_tracks(writer);` sequence of the reference (decompress.h:90-172). */
typedef struct aclb200_request
{
	uint32_t clip;				/* index into the clip set */
	float    sample_time;		/* seconds, clamped to the clip like seek() does (decompression.transform.h:215-216) */
} aclb200_request;

/* Everything the reference bakes into `decompression_settings` (decompression_settings.h:74-166) and
 * `track_writer` (track_writer.h:82-216) at compile time, plus the seek() arguments shared by the batch.
 * Zero-initialise then call aclb200_default_options(). */
typedef struct aclb200_options
{
	uint32_t struct_size;					/* sizeof(aclb200_options), for forward compatibility */

	/* seek(sample_time, rounding_policy) + set_looping_policy(policy) (decompress.h:147-160) */
	uint32_t rounding_policy;				/* ACLB200_ROUND_* */
	uint32_t looping_policy;				/* ACLB200_LOOP_* */

	/* decompression_settings */
	uint32_t normalization;					/* get_rotation_normalization_policy() */
	uint32_t per_track_rounding;			/* is_per_track_rounding_supported() */
	uint32_t wrapping;						/* is_wrapping_supported() */
	uint32_t clamp_sample_time;				/* clamp_sample_time() */
	uint32_t multiple_rotation_formats;		/* more than one is_rotation_format_supported(): only changes the result for quatf_full
											 * clips sampled exactly on a key frame (decompression_context.transform.h:191-200) */

	/* track_writer */
	uint32_t default_rotation_mode;			/* get_default_rotation_mode(): ACLB200_DEFAULT_* (legacy is scale only) */
	uint32_t default_translation_mode;
	uint32_t default_scale_mode;
	float    constant_defaults[12];			/* get_constant_default_rotation/translation/scale(): xyzw, xyz-, xyz- */
	const float* d_variable_defaults;		/* get_variable_default_*(track): device [max_tracks][12] floats, or NULL */
	const uint8_t* d_per_track_rounding;	/* get_rounding_policy(per_track, track): device [max_tracks] ACLB200_ROUND_*, or NULL */

	/* output */
	uint32_t output_layout;					/* ACLB200_LAYOUT_* */
	uint32_t math_mode;						/* ACLB200_MATH_* */
	uint64_t pose_stride_bytes;				/* distance between the poses of consecutive requests; 0 = max_tracks * bone size */

	/* track_writer::skip_all_rotations/translations/scales() and skip_track_rotation/translation/scale(track)
	 * (core/track_writer.h:181-191, honoured per sub-track at decompression.transform.h:626-665 and in every unpack pass):
	 * a skipped sub-track is NOT written, the output buffer keeps what it held. Transform clip sets only. */
	uint32_t skip_mask;						/* ACLB200_SKIP_* bits: sub-track kinds skipped for every track */
	const uint8_t* d_skip_track_mask;		/* device [max_tracks] bytes of ACLB200_SKIP_* bits: sub-tracks skipped per track, or NULL */

	/* seek(sample_time, rounding_policy) / set_looping_policy(policy) PER REQUEST (the reference takes both per call,
	 * decompress.h:147-160): device [num_requests][2] bytes { ACLB200_ROUND_* (none..nearest), ACLB200_LOOP_* }, or NULL to use
	 * rounding_policy / looping_policy above for the whole batch. Values out of range read as none / as_compressed;
	 * ACLB200_ROUND_PER_TRACK stays a batch wide choice (rounding_policy + d_per_track_rounding). */
	const uint8_t* d_request_policies;
} aclb200_options;

enum { ACLB200_SKIP_ROTATION = 1, ACLB200_SKIP_TRANSLATION = 2, ACLB200_SKIP_SCALE = 4 };

typedef struct aclb200_clipset_info
{
	uint32_t num_clips;
	uint32_t track_type;			/* ACLB200_TRACK_*: a clip set holds transform clips or scalar clips of one type, never both */
	uint32_t max_tracks;
	uint32_t min_tracks;
	uint64_t blob_bytes;			/* device bytes holding the compressed clips */
	uint64_t index_bytes;			/* device bytes of the acceleration index built at upload */
} aclb200_clipset_info;

typedef struct aclb200_clip_info
{
	uint32_t num_tracks;
	uint32_t num_samples;
	float    sample_rate;
	float    duration;				/* compressed_tracks::get_finite_duration() with the clip's own looping policy */
	uint32_t num_segments;
	uint32_t looping_policy;		/* compressed_tracks::get_looping_policy() */
	uint32_t hash;					/* compressed_tracks::get_hash() */
	uint32_t size;
} aclb200_clip_info;

/* Device-side result of seek(), exposed for integer parity checks against the reference
 * (persistent_transform_decompression_context_v0, decompression_context.transform.h:53-116). */
typedef struct aclb200_seek_state
{
	float    sample_time;			/* clamped; < 0 when the request was invalid */
	float    interpolation_alpha;
	uint32_t key_frame_bit_offsets[2];
	uint32_t segment_indices[2];
	uint32_t animated_offsets[2];	/* byte offsets of the animated bit streams relative to the start of the clip */
	uint32_t format_offsets[2];
	uint32_t range_offsets[2];
	uint32_t uses_single_segment;
	uint32_t looping_policy;
} aclb200_seek_state;

ACLB200_API const char* aclb200_version_string(void);
ACLB200_API const char* aclb200_status_string(aclb200_status status);

/* Fills `options` with default_transform_decompression_settings + acl::track_writer defaults
 * (decompression_settings.h:211-232, track_writer.h:170-186), rounding none, looping as_compressed,
 * layout QVV48, exact math. */
ACLB200_API void aclb200_default_options(aclb200_options* options);

/* Replaces make_decompression_context (decompress.h:205-209): binds a context to CUDA device `device`. */
ACLB200_API aclb200_status aclb200_create(int device, aclb200_context** out_context);
ACLB200_API void aclb200_destroy(aclb200_context* context);
/* Text of the last error raised on this context (never NULL). */
ACLB200_API const char* aclb200_last_error(const aclb200_context* context);

/* Replaces decompression_context::initialize(const compressed_tracks&) (decompress.h:90-101,
 * decompress.impl.h:66-83) for `num_clips` clips at once: validates every blob exactly like
 * compressed_tracks::is_valid(check_hash) + is_version_supported (v02_00_00 .. v02_01_00), copies the blobs to device memory (>= 64 bytes of tail slack for the `_unsafe` unaligned reads,
 * compress.transform.impl.h:387-396) and builds the acceleration index. Synchronous; the host blobs may be
 * freed when it returns. `out_failed_clip` (optional) receives the index of the first rejected clip.
 * A clip bound to a streaming database (compressed_tracks::has_database, what acl::build_database returns) is accepted and decodes from
 * the key frames that stay resident in the clip: what decompression_context<settings with database support>::initialize(tracks) gives
 * with no database bound, or with a database none of whose tiers is streamed in (decompress.impl.h:67-83, decompression.transform.h:
 * 262-265). Streaming the medium / low importance tiers in is not implemented. */
ACLB200_API aclb200_status aclb200_upload_clips(aclb200_context* context, const void* const* blobs, const uint32_t* sizes, uint32_t num_clips,
	uint32_t check_hash, aclb200_clipset** out_clipset, uint32_t* out_failed_clip);

/* Same, for clips stored back to back in one host buffer: clip i is buffer[offsets[i] .. offsets[i] + sizes[i]). */
ACLB200_API aclb200_status aclb200_upload_clips_packed(aclb200_context* context, const void* buffer, const uint64_t* offsets, const uint32_t* sizes,
	uint32_t num_clips, uint32_t check_hash, aclb200_clipset** out_clipset, uint32_t* out_failed_clip);

ACLB200_API void aclb200_release_clipset(aclb200_context* context, aclb200_clipset* clipset);
ACLB200_API aclb200_status aclb200_clipset_get_info(const aclb200_clipset* clipset, aclb200_clipset_info* out_info);
/* compressed_tracks accessors (compressed_tracks.h:60-140) */
ACLB200_API aclb200_status aclb200_clipset_get_clip_info(const aclb200_clipset* clipset, uint32_t clip, aclb200_clip_info* out_info);

/* Replaces seek() + decompress_tracks(writer) (decompress.h:147-166; seek_v0 + decompress_tracks_v0,
 * decompression.transform.h:206-563,1526-1737) for `num_requests` requests in one fused kernel.
 * `d_requests` and `d_out` are device pointers; pose r starts at d_out + r * pose_stride_bytes and holds
 * one bone every 48 or 40 bytes (options->output_layout). Sub-tracks whose default mode is `skipped` are not
 * written. Transform clip sets only. */
ACLB200_API aclb200_status aclb200_decompress_tracks(aclb200_context* context, const aclb200_clipset* clipset,
	const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options* options, void* d_out, void* stream);

/* Replaces seek() + decompress_track(track_index, writer) (decompress.h:168-172; decompress_track_v0,
 * decompression.transform.h:1753-2050): request r decodes bone d_track_indices[r] only and writes ONE bone
 * (48 / 40 bytes) at d_out + r * bone size. */
ACLB200_API aclb200_status aclb200_decompress_track(aclb200_context* context, const aclb200_clipset* clipset,
	const aclb200_request* d_requests, const uint32_t* d_track_indices, uint32_t num_requests,
	const aclb200_options* options, void* d_out, void* stream);

/* Scalar clip sets (float1f..float4f, vector4f): seek_v0 + decompress_tracks_v0 of decompression.scalar.h:181-481.
 * Request r writes num_tracks rows of `components` floats (write_float1..4 / write_vector4) at
 * d_out + r * pose_stride_bytes (0 = max_tracks * components * 4). */
ACLB200_API aclb200_status aclb200_scalar_decompress_tracks(aclb200_context* context, const aclb200_clipset* clipset,
	const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options* options, void* d_out, void* stream);

/* decompress_track_v0 of decompression.scalar.h:483-705: one track per request, `components` floats each. */
ACLB200_API aclb200_status aclb200_scalar_decompress_track(aclb200_context* context, const aclb200_clipset* clipset,
	const aclb200_request* d_requests, const uint32_t* d_track_indices, uint32_t num_requests,
	const aclb200_options* options, void* d_out, void* stream);

/* Host-buffer convenience over aclb200_decompress_tracks / aclb200_scalar_decompress_tracks: copies `num_requests`
 * requests from host memory, decodes, copies the poses back to `out` (host) and waits. This is the call the C++
 * header shim uses to replay results into a host-side track_writer. Pinned host memory makes the copies faster
 * but is not required. */
ACLB200_API aclb200_status aclb200_decompress_tracks_host(aclb200_context* context, const aclb200_clipset* clipset,
	const aclb200_request* requests, uint32_t num_requests, const aclb200_options* options, void* out, size_t out_bytes);

/* ---- the nearest callers of the decode path, on the device (acl_b200/csrc/error_metric.cu) ---------------- */

/* acl::track_error (compression/track_error.h:48-62) + what the measurement met on the way */
typedef struct aclb200_track_error
{
	uint32_t index;					/* track with the worst error (0xFFFFFFFF when nothing was measured) */
	float    error;
	float    sample_time;
	uint32_t flags;					/* ACLB200_ERROR_FLAG_* */
} aclb200_track_error;

/* acl::itransform_error_metric implementations (compression/transform_error_metrics.h) */
enum
{
	ACLB200_METRIC_QVVF = 0,				/* qvvf_transform_error_metric (:281-385), and additive_qvvf_transform_error_metric<format> (:470-526) for the
											 * jobs that carry an additive_format */
	ACLB200_METRIC_QVVF_MATRIX3X4F = 1		/* qvvf_matrix3x4f_transform_error_metric (:389-464): transforms as 3x4 matrices (for rigs with shear);
											 * every operation is IEEE exact: bit-identical to the reference on any CPU */
};

enum
{
	ACLB200_ERROR_FLAG_NEGATIVE_SCALE = 1,		/* informational: a negative scale took rtm::qvv_mul through its matrix branch (qvvf.h:320-345) somewhere */
	ACLB200_ERROR_FLAG_INVALID_SKELETON = 2		/* a parent index does not precede its child (the reference reads an unwritten transform there):
												 * the bone was treated as a root */
};

/* One clip to measure == one `calculate_compression_error(allocator, raw_tracks, context, error_metric)` call of the reference
 * (compression/track_error.h:64-121). The raw clip arrives already sampled: pose s of the job is what
 * `raw_tracks.sample_tracks(min(s / sample_rate, duration), rounding, writer)` writes (track_error.impl.h:337-338,504-507), where rounding is
 * nearest, or none when the compressed clip has stripped key frames (:556-559). */
typedef struct aclb200_error_job
{
	uint32_t clip;					/* index into the clip set */
	uint32_t num_samples;			/* raw_tracks.get_num_samples_per_track() */
	float    sample_rate;			/* raw_tracks.get_sample_rate() */
	float    duration;				/* raw_tracks.get_finite_duration() */
	uint32_t num_tracks;			/* raw_tracks.get_num_tracks() */
	uint32_t skeleton_offset;		/* first entry of this clip's skeleton in d_parent_indices / d_shell_distances / d_output_indices */
	uint64_t first_raw_pose;		/* pose index of sample 0 in d_raw_poses */
	uint32_t additive_format;		/* acl::additive_clip_format8 (core/additive_utils.h:42-66): 0 none, 1 relative, 2 additive0, 3 additive1 */
	uint32_t error_metric;			/* ACLB200_METRIC_* (transform clip sets) */
	uint64_t first_base_pose;		/* additive jobs: pose index of sample 0 in d_base_poses */
} aclb200_error_job;

/* Replaces calculate_compression_error (compression/impl/track_error.impl.h:400-571: calculate_transform_track_error :225-392 with the
 * qvvf_transform_error_metric of compression/transform_error_metrics.h:281-385, calculate_scalar_track_error :166-223) for `num_jobs` clips:
 * every sample of every clip is decoded on the device (decompress_tracks, rounding as above, `options` = the settings / writer of the context
 * the reference would be handed: pass the bind pose as constant or variable defaults), taken to object space and compared with the raw pose.
 *   jobs              HOST array
 *   d_raw_poses       device: pose p at p * pose_stride_bytes (options, 0 = max_tracks * 48): rtm::qvvf per bone (transform clip sets),
 *                     `components` floats per track (scalar clip sets: the layout aclb200_scalar_decompress_tracks writes)
 *   d_parent_indices  device: track_desc_transformf::parent_index per track (0xFFFFFFFF = root; a parent precedes its children),
 *   d_shell_distances device: track_desc_transformf::shell_distance per track; both unused (NULL) for scalar clip sets
 *   d_output_indices  device, optional: track_desc::output_index per raw track (0xFFFFFFFF = stripped from the compressed clip: the raw
 *                     value stands in, track_error.impl.h:522-532); NULL = every raw track i is output i
 *   d_base_poses      device, optional: the additive base of the jobs whose additive_format is not 0 (the calculate_compression_error
 *                     overload with additive_base_tracks, track_error.impl.h:573-680, + additive_qvvf_transform_error_metric<format>,
 *                     transform_error_metrics.h:470-526): pose s of a job is `additive_base_tracks.sample_tracks(t_base(s), rounding, writer)`
 *                     with t_base = (t / duration) * base_duration, or 0 when the base has one sample (:352-356); it is applied to the raw
 *                     and to the decoded pose with acl::apply_additive_to_base (core/additive_utils.h:147-157) before the hierarchy walk
 *   d_out_errors      device: one aclb200_track_error per job
 *   d_out_error_matrix device, optional: the error of every bone of every pose, row (poses of the earlier jobs + s) of
 *                     pose_stride_bytes / 48 floats (= max_tracks by default; scalar clip sets: tracks per row)
 * rtm::quat_normalize's rsqrtss estimate is CPU specific: errors agree with a given CPU's
 * within 5e-5 on poses tens of units across, not bit for bit (see error_metric.cu). Asynchronous on `stream`; uses scratch owned by the context. */
ACLB200_API aclb200_status aclb200_calculate_compression_error(aclb200_context* context, const aclb200_clipset* clipset, const aclb200_error_job* jobs,
	uint32_t num_jobs, const void* d_raw_poses, const uint32_t* d_parent_indices, const float* d_shell_distances,
	const uint32_t* d_output_indices, const void* d_base_poses, const aclb200_options* options, aclb200_track_error* d_out_errors,
	float* d_out_error_matrix, void* stream);

/* The sampling loop of acl::convert_track_list(allocator, const compressed_tracks&, track_array&) (compression/convert.h,
 * compression/impl/convert.impl.h:146-232) and of the error measurement above: EVERY sample of every listed clip in one launch sequence.
 * Job j contributes num_samples poses, sample i sought at min(i / sample_rate, duration) with options->rounding_policy (the reference
 * uses `nearest` "to land directly on a sample") and decoded like aclb200_decompress_tracks / aclb200_scalar_decompress_tracks would:
 * the poses of the jobs follow one another in d_out (pose stride and layout from `options`). Only clip, num_samples, sample_rate and
 * duration of a job are read. jobs is a HOST array; asynchronous on `stream`; uses scratch owned by the context. */
ACLB200_API aclb200_status aclb200_decompress_all_samples(aclb200_context* context, const aclb200_clipset* clipset, const aclb200_error_job* jobs,
	uint32_t num_jobs, const aclb200_options* options, void* d_out, void* stream);

/* Decoded poses held per chunk of clips by aclb200_calculate_compression_error (default 1 GiB). */
ACLB200_API aclb200_status aclb200_set_error_chunk_bytes(aclb200_context* context, uint64_t bytes);

/* Replaces qvvf_transform_error_metric::local_to_object_space (compression/transform_error_metrics.h:289-310: obj[i] =
 * qvv_normalize(qvv_mul(local[i], obj[parent[i]])), roots copied) for `num_poses` poses of one skeleton: rtm::qvvf rows in, rtm::qvvf rows
 * out (48 byte bones, pose_stride_bytes 0 = num_tracks * 48; the two buffers may be the same). d_out_flags (optional, device uint32):
 * ACLB200_ERROR_FLAG_* met on the way. */
ACLB200_API aclb200_status aclb200_local_to_object_space(aclb200_context* context, const void* d_local_poses, void* d_object_poses, uint64_t num_poses,
	uint32_t num_tracks, uint64_t pose_stride_bytes, const uint32_t* d_parent_indices, uint32_t* d_out_flags, void* stream);

/* Parity / debugging hooks (integer stages of the decode, bit-exact against the reference):
 *  - aclb200_debug_seek: the state seek_v0 computes, one aclb200_seek_state per request (device output).
 *  - aclb200_debug_unpack: for request r and key frame `which` (0/1), writes one uint4 per animated sub-track
 *    (rotations, translations, scales order): x, y, z quantised integers (or raw float bits) and the stored
 *    per-track bit count, at d_out + r * max_animated_sub_tracks * 16 bytes. */
ACLB200_API aclb200_status aclb200_debug_seek(aclb200_context* context, const aclb200_clipset* clipset,
	const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options* options, aclb200_seek_state* d_out, void* stream);
ACLB200_API aclb200_status aclb200_debug_unpack(aclb200_context* context, const aclb200_clipset* clipset,
	const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options* options, uint32_t which,
	uint32_t max_animated_sub_tracks, uint32_t* d_out, void* stream);

/* Profiling hook of the pipeline kernel (only builds compiled with -DACLB200_PIPE_TRACE=1 write anything): the first
 * `num_blocks` blocks record 8 clock64() stamps per batch they decode, for their first `num_iterations` batches, at
 * d_trace[(block * num_iterations + iteration) * 8 + k] (uint64). NULL switches it off. tools/pipe_trace.py reads it. */
ACLB200_API aclb200_status aclb200_debug_set_trace(aclb200_context* context, void* d_trace, uint32_t num_blocks, uint32_t num_iterations);

/* Plain device memory helpers so that a host language without the CUDA runtime can hand device arrays (requests, variable
 * defaults, per track policies, skip masks, outputs) to the entry points above. Synchronous. */
ACLB200_API aclb200_status aclb200_device_malloc(aclb200_context* context, size_t bytes, void** out_device_pointer);
ACLB200_API void aclb200_device_free(aclb200_context* context, void* device_pointer);
ACLB200_API aclb200_status aclb200_copy_to_device(aclb200_context* context, void* device_destination, const void* host_source, size_t bytes);
ACLB200_API aclb200_status aclb200_copy_to_host(aclb200_context* context, void* host_destination, const void* device_source, size_t bytes);

/* Number of kernels launched by this context so far (bench.py reports it as `gpu_launches`). */
ACLB200_API uint64_t aclb200_launch_count(const aclb200_context* context);

#ifdef __cplusplus
}
#endif

#endif /* ACLB200_H */
