/* oracle/acl_oracle.c -- TEST INFRASTRUCTURE ONLY (see acl_oracle.h).
 *
 * Plain-C restatement of the reference's uniformly-sampled decompression path. Every function cites
 * the reference file:line it follows (paths relative to /root/reference/includes/acl unless noted).
 * The structure is deliberately NOT the reference's (no SIMD groups of 4, no caches, no prefetches):
 * each sub-track is decoded on its own with scalar IEEE arithmetic in the same operation order, so the
 * results are bit-identical to the reference's SSE path (verified by tests/test_oracle_vs_reference.py).
 *
 * Build: gcc -std=c11 -O2 -msse4.1 -ffp-contract=off -fno-fast-math (oracle/Makefile). x86-64 only
 * (the decompress_track rotation path uses the same rsqrtss estimate as rtm::quat_normalize).
 */
#define _POSIX_C_SOURCE 200809L
#include "acl_oracle.h"

#include <math.h>
#include <string.h>
#include <time.h>
#include <xmmintrin.h>

/* ------------------------------------------------------------------------------------------------
 * Binary layout (core/impl/compressed_headers.h:51-58,61-131,140-165,171-197,227-325)
 * ---------------------------------------------------------------------------------------------- */

#define ACLO_TAG 0xac11ac11u							/* core/buffer_tag.h:49 */
#define ACLO_VERSION_FIRST 7u							/* v02_00_00, core/compressed_tracks_version.h:75 */
#define ACLO_VERSION_RAW31 9u							/* v02_01_99_1: raw bit rate stored as 31 */
#define ACLO_VERSION_LATEST 10u
#define ACLO_TRACK_QVVF 12u								/* core/track_types.h:68 */
#define ACLO_INVALID_OFFSET 0xFFFFFFFFu					/* core/ptr_offset.h */

#define ACLO_ROT_FULL 0u								/* core/track_formats.h:48-53 */
#define ACLO_ROT_DROP_W_FULL 2u
#define ACLO_ROT_DROP_W_VARIABLE 3u
#define ACLO_VEC_FULL 0u
#define ACLO_VEC_VARIABLE 1u

static uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd_u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static float rd_f32(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }
static float u32_as_f32(uint32_t u) { float v; memcpy(&v, &u, 4); return v; }
static uint32_t f32_as_u32(float f) { uint32_t v; memcpy(&v, &f, 4); return v; }

/* core/memory_utils.h byte_swap + unaligned_load: big-endian reads of the bit stream */
static uint32_t rd_be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
static uint64_t rd_be64(const uint8_t* p) { return ((uint64_t)rd_be32(p) << 32) | (uint64_t)rd_be32(p + 4); }

static uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) & ~(a - 1); }

typedef struct clip_view
{
	const uint8_t* blob;
	uint32_t size;
	uint32_t version;
	uint32_t track_type;
	uint32_t num_tracks;
	uint32_t num_samples;
	float    sample_rate;
	uint32_t misc;

	/* transform only */
	const uint8_t* th;						/* transform_tracks_header, compressed_tracks.impl.h:53-60 */
	uint32_t num_segments;
	uint32_t num_animated_variable_sub_tracks;
	uint32_t num_animated[3];				/* rotation, translation, scale */
	uint32_t num_constant[3];
	uint32_t segment_headers_offset;
	uint32_t sub_track_types_offset;
	uint32_t constant_data_offset;
	uint32_t clip_range_offset;
	uint32_t rotation_format;
	uint32_t translation_format;
	uint32_t scale_format;
	uint32_t has_scale;
	uint32_t default_scale;
	uint32_t has_stripped_keyframes;
	uint32_t has_database;
	uint32_t is_wrap_optimized;
	uint32_t segment_header_size;
} clip_view;

uint32_t aclo_hash32(const void* data, size_t size)
{
	/* core/hash.h:44-84 FNV-1a 32 */
	const uint8_t* bytes = (const uint8_t*)data;
	uint32_t acc = 2166136261u;
	for (size_t i = 0; i < size; ++i)
		acc = (acc ^ bytes[i]) * 16777619u;
	return acc;
}

static int view_clip(const void* blob, clip_view* v)
{
	const uint8_t* p = (const uint8_t*)blob;
	memset(v, 0, sizeof(*v));
	v->blob = p;
	v->size = rd_u32(p + 0);								/* raw_buffer_header::size */
	/* tracks_header at +8, compressed_headers.h:61-86 */
	if (rd_u32(p + 8) != ACLO_TAG)
		return -1;
	v->version = rd_u16(p + 12);
	v->track_type = p[15];
	v->num_tracks = rd_u32(p + 16);
	v->num_samples = rd_u32(p + 20);
	v->sample_rate = rd_f32(p + 24);
	v->misc = rd_u32(p + 28);
	v->is_wrap_optimized = (v->misc >> 30) & 1;			/* compressed_headers.h:127 */

	if (v->track_type == ACLO_TRACK_QVVF)
	{
		const uint8_t* th = p + 32;
		v->th = th;
		/* transform_tracks_header, compressed_headers.h:227-262 */
		v->num_segments = rd_u32(th + 0);
		v->num_animated_variable_sub_tracks = rd_u32(th + 4);
		v->num_animated[0] = rd_u32(th + 8);
		v->num_animated[1] = rd_u32(th + 12);
		v->num_animated[2] = rd_u32(th + 16);
		v->num_constant[0] = rd_u32(th + 20);
		v->num_constant[1] = rd_u32(th + 24);
		v->num_constant[2] = rd_u32(th + 28);
		/* th + 32: database_header_offset */
		v->segment_headers_offset = rd_u32(th + 36);
		v->sub_track_types_offset = rd_u32(th + 40);
		v->constant_data_offset = rd_u32(th + 44);
		v->clip_range_offset = rd_u32(th + 48);
		/* misc_packed accessors, compressed_headers.h:109-124 */
		v->has_scale = v->misc & 1;
		v->default_scale = (v->misc >> 1) & 1;
		v->scale_format = (v->misc >> 2) & 1;
		v->translation_format = (v->misc >> 3) & 1;
		v->rotation_format = (v->misc >> 4) & 15;
		v->has_database = (v->misc >> 8) & 1;
		v->has_stripped_keyframes = (v->misc >> 10) & 1;
		/* segment_header is 16 bytes, stripped_segment_header_t adds `sample_indices` (compressed_headers.h:171-197) */
		v->segment_header_size = (v->has_stripped_keyframes || v->has_database) ? 20u : 16u;
	}
	return 0;
}

int aclo_validate(const void* blob, size_t size, int check_hash)
{
	/* compressed_tracks::is_valid, core/impl/compressed_tracks.impl.h:278-301, then
	 * decompression_version_selector::is_version_supported */
	if (blob == NULL || size < 32)
		return -1;
	if (((uintptr_t)blob & 15u) != 0)
		return -2;											/* alignof(compressed_tracks) == 16 */
	const uint8_t* p = (const uint8_t*)blob;
	if (rd_u32(p + 8) != ACLO_TAG)
		return -3;
	if (p[14] != 0)
		return -4;											/* algorithm_type8::uniformly_sampled == 0 */
	const uint32_t version = rd_u16(p + 12);
	if (version < ACLO_VERSION_FIRST || version > ACLO_VERSION_LATEST)
		return -5;
	const uint32_t stored_size = rd_u32(p);
	if (stored_size > size)
		return -6;
	if (check_hash)
	{
		if (aclo_hash32(p + 8, stored_size - 8) != rd_u32(p + 4))
			return -7;
	}
	if (p[15] == ACLO_TRACK_QVVF && ((rd_u32(p + 28) >> 8) & 1))
	{
		/* A clip bound to a database decodes from the key frames resident in the clip, like a context initialised without
		 * its database (decompress.impl.h:67-83). Its transform header names the database metadata (compressed_headers.h:245-246): a
		 * flag without that header is a corrupt clip. */
		if (stored_size < 32 + 52 || rd_u32(p + 32 + 32) == ACLO_INVALID_OFFSET || rd_u32(p + 32 + 32) >= stored_size - 32)
			return -8;
	}
	return 0;
}

void aclo_default_settings(aclo_settings* s)
{
	memset(s, 0, sizeof(*s));
	s->normalization = ACLO_NORMALIZE_LERP_ONLY;			/* decompression_settings.h:226 */
	s->per_track_rounding = 0;								/* :231 */
	s->wrapping = 1;										/* :153 */
	s->clamp_sample_time = 1;								/* :80 */
	s->multiple_rotation_formats = 0;						/* :219 only quatf_drop_w_variable */
	s->default_rotation_mode = ACLO_DEFAULT_CONSTANT;		/* track_writer.h:170-172 */
	s->default_translation_mode = ACLO_DEFAULT_CONSTANT;
	s->default_scale_mode = ACLO_DEFAULT_LEGACY;
	s->constant_defaults[3] = 1.0f;							/* identity rotation, zero translation, one scale (:174-176) */
	s->constant_defaults[8] = s->constant_defaults[9] = s->constant_defaults[10] = 1.0f;
}

/* ------------------------------------------------------------------------------------------------
 * Seek math
 * ---------------------------------------------------------------------------------------------- */

/* core/impl/interpolation_utils.impl.h:261-278 */
static float apply_rounding_policy(float alpha, uint32_t policy)
{
	switch (policy)
	{
	default:
	case ACLO_ROUND_NONE:
	case ACLO_ROUND_PER_TRACK:
		return alpha;
	case ACLO_ROUND_FLOOR:
		return 0.0f;
	case ACLO_ROUND_CEIL:
		return 1.0f;
	case ACLO_ROUND_NEAREST:
		return floorf(alpha + 0.5f);
	}
}

/* core/impl/interpolation_utils.impl.h:143-201 */
static void find_linear_interpolation_samples_with_sample_rate(uint32_t num_samples, float sample_rate, float sample_time,
	uint32_t rounding_policy, uint32_t looping_policy, uint32_t* out_index0, uint32_t* out_index1, float* out_alpha)
{
	const uint32_t last_sample_index = num_samples - 1;
	float sample_index = sample_time * sample_rate;
	uint32_t sample_index0 = (uint32_t)sample_index;
	const uint32_t next_sample_index = sample_index0 + 1;
	uint32_t sample_index1;
	if (looping_policy == ACLO_LOOP_CLAMP)
		sample_index1 = next_sample_index < last_sample_index ? next_sample_index : last_sample_index;
	else
	{
		if (sample_index0 > last_sample_index)
		{
			sample_index = 0.0f;
			sample_index0 = 0;
			sample_index1 = 0;
		}
		else
			sample_index1 = next_sample_index >= num_samples ? 0 : next_sample_index;
	}
	const float alpha = sample_index - (float)sample_index0;
	*out_index0 = sample_index0;
	*out_index1 = sample_index1;
	*out_alpha = apply_rounding_policy(alpha, rounding_policy);
}

/* exported so that the reference's own known-answer table (tests/sources/core/test_interpolation_utils.cpp:225-331) can be replayed */
void aclo_find_key_frames(uint32_t num_samples, float sample_rate, float sample_time, uint32_t rounding_policy, uint32_t looping_policy,
	uint32_t* out_index0, uint32_t* out_index1, float* out_alpha)
{
	find_linear_interpolation_samples_with_sample_rate(num_samples, sample_rate, sample_time, rounding_policy, looping_policy, out_index0, out_index1, out_alpha);
}

/* core/impl/interpolation_utils.impl.h:224-253 */
static float find_linear_interpolation_alpha(float sample_index, uint32_t index0, uint32_t index1, uint32_t rounding_policy)
{
	if (rounding_policy == ACLO_ROUND_FLOOR)
		return 0.0f;
	if (rounding_policy == ACLO_ROUND_CEIL)
		return 1.0f;
	if (index0 == index1)
		return 0.0f;
	float alpha;
	if (index0 < index1)
		alpha = (sample_index - (float)index0) / (float)(index1 - index0);
	else
		alpha = sample_index - (float)index0;
	if (rounding_policy == ACLO_ROUND_NONE || rounding_policy == ACLO_ROUND_PER_TRACK)
		return alpha;
	return floorf(alpha + 0.5f);
}

/* compressed_tracks::get_looping_policy / get_finite_duration, core/impl/compressed_tracks.impl.h:102-141
 * + calculate_finite_duration, core/impl/time_utils.impl.h:105-114 */
static uint32_t clip_looping_policy(const clip_view* v)
{
	if (v->version <= ACLO_VERSION_FIRST)
		return ACLO_LOOP_CLAMP;
	return v->is_wrap_optimized ? ACLO_LOOP_WRAP : ACLO_LOOP_CLAMP;
}

static float clip_finite_duration(const clip_view* v, uint32_t looping_policy)
{
	if (looping_policy == ACLO_LOOP_AS_COMPRESSED)
		looping_policy = clip_looping_policy(v);
	uint32_t num_samples = v->num_samples;
	if (looping_policy == ACLO_LOOP_WRAP && num_samples != 0)
		num_samples++;
	if (num_samples <= 1)
		return 0.0f;
	return (float)(num_samples - 1) / v->sample_rate;
}

/* initialize_v0 + set_looping_policy_v0, decompression/impl/decompression.transform.h:120-129,186-204 */
static void resolve_looping(const clip_view* v, const aclo_settings* settings, uint32_t requested, uint32_t* out_policy, float* out_duration)
{
	if (!settings->wrapping)
	{
		*out_policy = ACLO_LOOP_CLAMP;
		*out_duration = clip_finite_duration(v, ACLO_LOOP_CLAMP);
		return;
	}
	uint32_t policy = requested == ACLO_LOOP_AS_COMPRESSED ? clip_looping_policy(v) : requested;
	*out_policy = policy;
	*out_duration = clip_finite_duration(v, policy);
}

/* rtm::scalar_clamp == min(max(x, lo), hi) with SSE min/max operand order */
static float clampf(float x, float lo, float hi)
{
	float m = x > lo ? x : lo;		/* _mm_max_ss(x, lo): returns lo when x is NaN or equal */
	return m < hi ? m : hi;
}

static uint32_t ctz32(uint32_t v) { return v != 0 ? (uint32_t)__builtin_ctz(v) : 32u; }	/* core/bit_manip_utils.h:163-175 */
static uint32_t clz32(uint32_t v) { return v != 0 ? (uint32_t)__builtin_clz(v) : 32u; }	/* :142-160 */
static uint32_t popc32(uint32_t v) { return (uint32_t)__builtin_popcount(v); }

/* transform_tracks_header::get_segment_data, core/impl/compressed_headers.h:310-324 */
static void segment_data_offsets(const clip_view* v, uint32_t segment_header_offset_from_th, uint32_t* format_off, uint32_t* range_off, uint32_t* animated_off)
{
	const uint8_t* header = v->th + segment_header_offset_from_th;
	const uint32_t segment_data = rd_u32(header + 12);				/* ptr_offset32 relative to the transform header */
	const uint32_t base = 32u + segment_data;						/* -> relative to the blob */
	/* align_to() aligns the ADDRESS; the blob is 16 byte aligned so aligning blob offsets is equivalent */
	const uint32_t range = align_up(base + v->num_animated_variable_sub_tracks, 2);
	const uint32_t range_size = v->num_segments > 1 ? 6u * v->num_animated_variable_sub_tracks : 0u;
	*format_off = base;
	*range_off = range;
	*animated_off = align_up(range + range_size, 4);
}

int aclo_transform_seek(const void* blob, const aclo_settings* settings, float sample_time,
	uint32_t rounding_policy, uint32_t looping_policy, aclo_seek_state* st)
{
	/* seek_v0, decompression/impl/decompression.transform.h:206-563 with no database bound (db == nullptr: the tier metadata branches
	 * :289-305,326-353 fall away, a database clip is sought like a clip with stripped key frames, :262-265) */
	clip_view v;
	if (view_clip(blob, &v) != 0 || v.track_type != ACLO_TRACK_QVVF)
		return -1;
	memset(st, 0, sizeof(*st));
	st->sample_time = -1.0f;
	if (v.num_tracks == 0)
		return 0;

	uint32_t policy;
	float duration;
	resolve_looping(&v, settings, looping_policy, &policy, &duration);
	st->clip_duration = duration;
	st->looping_policy = policy;

	if (settings->clamp_sample_time)
		sample_time = clampf(sample_time, 0.0f, duration);				/* :215-216 */
	st->sample_time = sample_time;

	uint32_t key_frame0, key_frame1;
	float alpha;
	find_linear_interpolation_samples_with_sample_rate(v.num_samples, v.sample_rate, sample_time, rounding_policy, policy, &key_frame0, &key_frame1, &alpha);
	st->rounding_policy = rounding_policy;

	uint32_t segment_key_frame0, segment_key_frame1;
	uint32_t segment_index0 = 0, segment_index1 = 0;
	const uint32_t headers = v.segment_headers_offset;
	const uint32_t hsize = v.segment_header_size;

	if (v.num_segments == 1)
	{
		if (v.has_stripped_keyframes || v.has_database)
		{
			/* :272-362 */
			const uint32_t sample_indices0 = rd_u32(v.th + headers + 16);
			const float sample_index = alpha + (float)key_frame0;
			const uint32_t candidate_indices0 = sample_indices0 & (0xFFFFFFFFu << (31 - key_frame0));
			key_frame0 = 31 - ctz32(candidate_indices0);
			const uint32_t candidate_indices1 = sample_indices0 & (0xFFFFFFFFu >> key_frame1);
			key_frame1 = clz32(candidate_indices1);
			alpha = find_linear_interpolation_alpha(sample_index, key_frame0, key_frame1, ACLO_ROUND_NONE);
			segment_key_frame0 = popc32(~(0xFFFFFFFFu >> key_frame0) & sample_indices0);	/* and_not(a, b) == ~a & b */
			segment_key_frame1 = popc32(~(0xFFFFFFFFu >> key_frame1) & sample_indices0);
		}
		else
		{
			segment_key_frame0 = key_frame0;
			segment_key_frame1 = key_frame1;
		}
	}
	else
	{
		/* :372-520; segment_start_indices follow the 52 byte header, compressed_headers.h:271-272 */
		const uint8_t* start_indices = v.th + 52;
		const uint32_t approx_num_samples_per_segment = v.num_samples / v.num_segments;
		const uint32_t approx_segment_index = key_frame0 / approx_num_samples_per_segment;
		const uint32_t start_segment_index = approx_segment_index > 0 ? approx_segment_index - 1 : 0;
		const uint32_t end_segment_index = start_segment_index + 4;
		for (uint32_t segment_index = start_segment_index; segment_index < end_segment_index; ++segment_index)
		{
			const uint32_t start = rd_u32(start_indices + 4 * segment_index);
			if (key_frame0 < start)
			{
				segment_index0 = segment_index - 1;
				if (settings->wrapping && key_frame1 == 0)
					segment_index1 = 0;
				else
					segment_index1 = key_frame1 < start ? segment_index0 : segment_index;
				break;
			}
		}
		const uint32_t start0 = rd_u32(start_indices + 4 * segment_index0);
		const uint32_t start1 = rd_u32(start_indices + 4 * segment_index1);
		segment_key_frame0 = key_frame0 - start0;
		segment_key_frame1 = key_frame1 - start1;

		if (v.has_stripped_keyframes || v.has_database)
		{
			/* :411-515 */
			const uint32_t sample_indices0 = rd_u32(v.th + headers + hsize * segment_index0 + 16);
			const uint32_t sample_indices1 = rd_u32(v.th + headers + hsize * segment_index1 + 16);
			const float sample_index = alpha + (float)key_frame0;
			const uint32_t candidate_indices0 = sample_indices0 & (0xFFFFFFFFu << (31 - segment_key_frame0));
			segment_key_frame0 = 31 - ctz32(candidate_indices0);
			const uint32_t candidate_indices1 = sample_indices1 & (0xFFFFFFFFu >> segment_key_frame1);
			segment_key_frame1 = clz32(candidate_indices1);
			const uint32_t clip_key_frame0 = start0 + segment_key_frame0;
			const uint32_t clip_key_frame1 = start1 + segment_key_frame1;
			alpha = find_linear_interpolation_alpha(sample_index, clip_key_frame0, clip_key_frame1, ACLO_ROUND_NONE);
			key_frame0 = clip_key_frame0;
			key_frame1 = clip_key_frame1;
			segment_key_frame0 = popc32(~(0xFFFFFFFFu >> segment_key_frame0) & sample_indices0);
			segment_key_frame1 = popc32(~(0xFFFFFFFFu >> segment_key_frame1) & sample_indices1);
		}
	}

	st->interpolation_alpha = alpha;
	st->key_frames[0] = key_frame0;
	st->key_frames[1] = key_frame1;
	st->segment_indices[0] = segment_index0;
	st->segment_indices[1] = segment_index1;
	st->segment_key_frames[0] = segment_key_frame0;
	st->segment_key_frames[1] = segment_key_frame1;
	st->uses_single_segment = segment_index0 == segment_index1;				/* :530 */

	for (int i = 0; i < 2; ++i)
	{
		const uint32_t header_off = headers + hsize * st->segment_indices[i];
		segment_data_offsets(&v, header_off, &st->format_offsets[i], &st->range_offsets[i], &st->animated_offsets[i]);
		st->segment_offsets[i] = 32u + header_off;
		const uint32_t animated_pose_bit_size = rd_u32(v.th + header_off + 0);
		st->key_frame_bit_offsets[i] = (i == 0 ? segment_key_frame0 : segment_key_frame1) * animated_pose_bit_size;	/* :558-559 */
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Bit stream unpacking (math/vector4_packing.h, math/scalar_packing.h)
 * ---------------------------------------------------------------------------------------------- */

/* unpack_vector3_uXX_unsafe integer part, math/vector4_packing.h:947-971 (one component) */
static uint32_t unpack_bits(const uint8_t* data, uint32_t bit_offset, uint32_t num_bits)
{
	const uint32_t bit_shift = 32 - num_bits;
	const uint32_t mask = (1u << num_bits) - 1;
	const uint32_t v = rd_be32(data + bit_offset / 8);
	return (v >> (bit_shift - (bit_offset % 8))) & mask;
}

/* unpack_vector3_96_unsafe / unpack_scalarf_32_unsafe integer part, math/vector4_packing.h:482-503 */
static uint32_t unpack_raw32(const uint8_t* data, uint32_t bit_offset)
{
	uint64_t v = rd_be64(data + bit_offset / 8);
	v <<= bit_offset % 8;
	v >>= 32;
	return (uint32_t)v;
}

/* PackedTableEntry::max_value, math/vector4_packing.h:927-929 (a float division evaluated in float) */
static float inv_max_value(uint32_t num_bits)
{
	return num_bits == 0 ? 1.0f : 1.0f / (float)((1 << num_bits) - 1);
}

static uint32_t stored_bits_to_stream_bits(uint32_t num_bits, uint32_t raw_marker)
{
	return num_bits == raw_marker ? 32u : num_bits;	/* animated_track_cache.transform.h:589-599,1176 */
}

/* ------------------------------------------------------------------------------------------------
 * Quaternion helpers (math/quatf.h SOA versions; the scalar lanes are identical)
 * ---------------------------------------------------------------------------------------------- */

/* quat_from_positive_w4, math/quatf.h:135-147: w = sqrt(|((1 - x*x) - y*y) - z*z|) */
static float quat_w_from_xyz(float x, float y, float z)
{
	float r = 1.0f - x * x;
	r = r - y * y;
	r = r - z * z;
	return sqrtf(fabsf(r));
}

/* quat_normalize4, math/quatf.h:200-211 */
static void quat_normalize4(float q[4])
{
	float dot = q[0] * q[0];
	dot = q[1] * q[1] + dot;
	dot = q[2] * q[2] + dot;
	dot = q[3] * q[3] + dot;
	const float len = sqrtf(dot);
	const float inv_len = 1.0f / len;
	q[0] = q[0] * inv_len;
	q[1] = q[1] * inv_len;
	q[2] = q[2] * inv_len;
	q[3] = q[3] * inv_len;
}

/* quat_lerp_no_normalization4, math/quatf.h:170-196 */
static void quat_lerp_no_normalization4(const float s[4], const float e[4], float alpha, float out[4])
{
	float dot = s[0] * e[0];
	dot = s[1] * e[1] + dot;
	dot = s[2] * e[2] + dot;
	dot = s[3] * e[3] + dot;
	const uint32_t bias = f32_as_u32(dot) & 0x80000000u;
	for (int i = 0; i < 4; ++i)
	{
		const float e_biased = u32_as_f32(f32_as_u32(e[i]) ^ bias);
		out[i] = e_biased * alpha + (s[i] - s[i] * alpha);
	}
}

/* rtm::quat_normalize SSE2 path, external/rtm/includes/rtm/quatf.h:917-953 (rsqrtss + 2 Newton-Raphson) */
static void rtm_quat_normalize(float q[4])
{
	const float x2 = q[0] * q[0], y2 = q[1] * q[1], z2 = q[2] * q[2], w2 = q[3] * q[3];
	const float dot = (x2 + z2) + (y2 + w2);
	const float half = 0.5f;
	const float input_half = dot * half;
	const float x0 = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(dot)));
	float x1 = x0 * x0;
	x1 = half - input_half * x1;
	x1 = x0 * x1 + x0;
	float x2_ = x1 * x1;
	x2_ = half - input_half * x2_;
	x2_ = x1 * x2_ + x1;
	for (int i = 0; i < 4; ++i)
		q[i] = q[i] * x2_;
}

/* sign of _mm_dp_ps(start, end, 0xFF): the dpps micro-code sums (x+y)+(z+w) */
static uint32_t dot_bias_dpps(const float s[4], const float e[4])
{
	const float dot = (s[0] * e[0] + s[1] * e[1]) + (s[2] * e[2] + s[3] * e[3]);
	return f32_as_u32(dot) & 0x80000000u;
}

/* rtm::quat_lerp SSE4 path, external/rtm/includes/rtm/quatf.h:1006-1075 */
static void rtm_quat_lerp(const float s[4], const float e[4], float alpha, float out[4])
{
	const uint32_t bias = dot_bias_dpps(s, e);
	for (int i = 0; i < 4; ++i)
		out[i] = (s[i] - alpha * s[i]) + alpha * u32_as_f32(f32_as_u32(e[i]) ^ bias);
	rtm_quat_normalize(out);
}

/* acl::quat_lerp_no_normalization SSE4 path, math/quatf.h:40-82 */
static void acl_quat_lerp_no_normalization(const float s[4], const float e[4], float alpha, float out[4])
{
	const uint32_t bias = dot_bias_dpps(s, e);
	for (int i = 0; i < 4; ++i)
		out[i] = (s[i] - alpha * s[i]) + alpha * u32_as_f32(f32_as_u32(e[i]) ^ bias);
}

/* rtm::vector_lerp, external/rtm/includes/rtm/vector4f.h:2417-2421 */
static float lerpf(float start, float end, float alpha)
{
	return end * alpha + (start - start * alpha);
}

/* should_interpolate_samples, decompression/impl/decompression_context.transform.h:191-200 */
static int should_interpolate(const aclo_settings* settings, uint32_t rotation_format, float alpha)
{
	if (settings->multiple_rotation_formats)
		return 1;
	return rotation_format == ACLO_ROT_FULL ? (alpha > 0.0f && alpha < 1.0f) : 1;
}

/* ------------------------------------------------------------------------------------------------
 * Transform decode
 * ---------------------------------------------------------------------------------------------- */

typedef struct segment_cursor
{
	/* one per key frame: where the three sub-track kinds start in this segment
	 * (animated_track_cache_v0::initialize, animated_track_cache.transform.h:1224-1314) */
	const uint8_t* format[3];
	const uint8_t* range[3];
	const uint8_t* animated;
	uint32_t bit_offset[3];
} segment_cursor;

typedef struct transform_decoder
{
	clip_view v;
	const aclo_settings* settings;
	const aclo_seek_state* st;
	int variable[3];				/* is the rotation / translation / scale format variable? */
	uint32_t raw_marker;			/* per-track format value meaning "raw": 31 (>= v02_01_99_1) or 32 */
	const uint8_t* clip_range[3];
	segment_cursor seg[2];
	const uint8_t* constant[3];
	int has_segments;
} transform_decoder;

static void decoder_init(transform_decoder* d, const void* blob, const aclo_settings* settings, const aclo_seek_state* st)
{
	view_clip(blob, &d->v);
	const clip_view* v = &d->v;
	d->settings = settings;
	d->st = st;
	d->variable[0] = v->rotation_format == ACLO_ROT_DROP_W_VARIABLE;
	d->variable[1] = v->translation_format == ACLO_VEC_VARIABLE;
	d->variable[2] = v->scale_format == ACLO_VEC_VARIABLE;
	d->raw_marker = v->version >= ACLO_VERSION_RAW31 ? 31u : 32u;			/* animated_track_cache.transform.h:523 */
	d->has_segments = v->num_segments > 1;

	const uint32_t padded_rotations = align_up(v->num_animated[0], 4);		/* :1257 */

	/* clip range: rotations (24 B each when variable), translations, scales, no padding (:1259-1262,1293-1294) */
	d->clip_range[0] = v->th + v->clip_range_offset;
	d->clip_range[1] = d->clip_range[0] + (d->variable[0] ? 24u * v->num_animated[0] : 0u);
	d->clip_range[2] = d->clip_range[1] + (d->variable[1] ? 24u * v->num_animated[1] : 0u);

	for (int k = 0; k < 2; ++k)
	{
		segment_cursor* c = &d->seg[k];
		const uint8_t* header = v->blob + st->segment_offsets[k];
		const uint32_t rotation_bit_size = rd_u32(header + 4);
		const uint32_t translation_bit_size = rd_u32(header + 8);
		c->animated = v->blob + st->animated_offsets[k];
		c->format[0] = v->blob + st->format_offsets[k];
		c->range[0] = v->blob + st->range_offsets[k];
		c->bit_offset[0] = st->key_frame_bit_offsets[k];
		/* rotation metadata / segment range are padded to groups of 4 (:1264-1276) */
		c->format[1] = c->format[0] + (d->variable[0] ? padded_rotations : 0u);
		c->range[1] = c->range[0] + (d->variable[0] ? 6u * padded_rotations : 0u);
		c->bit_offset[1] = c->bit_offset[0] + rotation_bit_size;				/* :1282-1283 */
		c->format[2] = c->format[1] + (d->variable[1] ? v->num_animated[1] : 0u);	/* :1296-1302 */
		c->range[2] = c->range[1] + (d->variable[1] ? 6u * v->num_animated[1] : 0u);
		c->bit_offset[2] = c->bit_offset[1] + translation_bit_size;			/* :1307-1308 */
	}

	/* constant_track_cache_v0::initialize, constant_track_cache.transform.h:96-110 */
	const uint32_t packed_rotation_size = v->rotation_format == ACLO_ROT_FULL ? 16u : 12u;
	d->constant[0] = v->th + v->constant_data_offset;
	d->constant[1] = d->constant[0] + packed_rotation_size * v->num_constant[0];
	d->constant[2] = d->constant[1] + 12u * v->num_constant[1];
}

/* Bit offset of animated sub-track `index` of kind `kind` inside key frame `k`: the running sum the
 * reference keeps in segment_animated_sampling_context_v0::animated_track_data_bit_offset. */
static uint32_t animated_bit_offset(const transform_decoder* d, int k, int kind, uint32_t index)
{
	const segment_cursor* c = &d->seg[k];
	if (!d->variable[kind])
	{
		const uint32_t sample_bits = (kind == 0 && d->v.rotation_format == ACLO_ROT_FULL) ? 128u : 96u;
		return c->bit_offset[kind] + index * sample_bits;
	}
	uint32_t offset = c->bit_offset[kind];
	for (uint32_t j = 0; j < index; ++j)
		offset += 3u * stored_bits_to_stream_bits(c->format[kind][j], d->raw_marker);
	return offset;
}

/* One animated rotation sample of key frame k as it is after range expansion and W reconstruction.
 * `soa_path` selects the decompress_tracks flavour (unpack_animated_quat + remap_*_range_data4 +
 * quat_from_positive_w4, animated_track_cache.transform.h:515-687,302-350,391-466) or the decompress_track
 * flavour (unpack_single_animated_quat, :689-869, where ignored ranges are skipped instead of multiplied by 1). */
static void decode_animated_rotation(const transform_decoder* d, int k, uint32_t index, int soa_path, float out[4])
{
	const clip_view* v = &d->v;
	const segment_cursor* c = &d->seg[k];
	const uint32_t group = index / 4;
	const uint32_t lane = index % 4;
	const uint32_t remaining = v->num_animated[0] - group * 4;
	const uint32_t group_size = remaining < 4 ? remaining : 4;
	const uint32_t bit_offset = animated_bit_offset(d, k, 0, index);

	float x, y, z;
	int ignore_segment = 0, ignore_clip = 0;

	if (d->variable[0])
	{
		const uint32_t num_bits = c->format[0][index];
		if (num_bits == 0)
		{
			/* constant within the segment: 16 bits per component spread over the SOA range bytes (:552-587) */
			const uint8_t* r = c->range[0] + group * 24 + lane;
			const uint32_t xi = ((uint32_t)r[0] << 8) | r[4];
			const uint32_t yi = ((uint32_t)r[8] << 8) | r[12];
			const uint32_t zi = ((uint32_t)r[16] << 8) | r[20];
			const float scale = 1.0f / 65535.0f;
			x = (float)xi * scale; y = (float)yi * scale; z = (float)zi * scale;
			ignore_segment = 1;
		}
		else if (num_bits == d->raw_marker)
		{
			x = u32_as_f32(unpack_raw32(c->animated, bit_offset));
			y = u32_as_f32(unpack_raw32(c->animated, bit_offset + 32));
			z = u32_as_f32(unpack_raw32(c->animated, bit_offset + 64));
			ignore_segment = 1;
			ignore_clip = 1;
		}
		else
		{
			const float inv_max = inv_max_value(num_bits);
			x = (float)unpack_bits(c->animated, bit_offset, num_bits) * inv_max;
			y = (float)unpack_bits(c->animated, bit_offset + num_bits, num_bits) * inv_max;
			z = (float)unpack_bits(c->animated, bit_offset + num_bits * 2, num_bits) * inv_max;
		}

		if (d->has_segments && (soa_path || !ignore_segment))
		{
			/* unpack_segment_range_data + remap_segment_range_data4 (:157-298,302-350). Key frame 1 reads
			 * segment 1's range only when two segments are in use (:1392). */
			const uint8_t* r = c->range[0] + group * 24 + lane;
			const float n = 1.0f / 255.0f;
			float min_x = (float)r[0] * n, min_y = (float)r[4] * n, min_z = (float)r[8] * n;
			float ext_x = (float)r[12] * n, ext_y = (float)r[16] * n, ext_z = (float)r[20] * n;
			if (ignore_segment)
			{
				min_x = min_y = min_z = 0.0f;
				ext_x = ext_y = ext_z = 1.0f;
			}
			x = x * ext_x + min_x;
			y = y * ext_y + min_y;
			z = z * ext_z + min_z;
		}

		if (soa_path || !ignore_clip)
		{
			/* remap_clip_range_data4 (:391-466): SOA per group, the last group holds `group_size` lanes */
			const uint8_t* r = d->clip_range[0] + group * 96 + lane * 4;
			const uint32_t stride = group_size * 4;
			float min_x = rd_f32(r + stride * 0), min_y = rd_f32(r + stride * 1), min_z = rd_f32(r + stride * 2);
			float ext_x = rd_f32(r + stride * 3), ext_y = rd_f32(r + stride * 4), ext_z = rd_f32(r + stride * 5);
			if (ignore_clip)
			{
				min_x = min_y = min_z = 0.0f;
				ext_x = ext_y = ext_z = 1.0f;
			}
			x = x * ext_x + min_x;
			y = y * ext_y + min_y;
			z = z * ext_z + min_z;
		}
	}
	else if (v->rotation_format == ACLO_ROT_FULL)
	{
		/* unpack_vector4_128_unsafe, math/vector4_packing.h:59-92 */
		out[0] = u32_as_f32(unpack_raw32(c->animated, bit_offset));
		out[1] = u32_as_f32(unpack_raw32(c->animated, bit_offset + 32));
		out[2] = u32_as_f32(unpack_raw32(c->animated, bit_offset + 64));
		out[3] = u32_as_f32(unpack_raw32(c->animated, bit_offset + 96));
		return;
	}
	else
	{
		x = u32_as_f32(unpack_raw32(c->animated, bit_offset));
		y = u32_as_f32(unpack_raw32(c->animated, bit_offset + 32));
		z = u32_as_f32(unpack_raw32(c->animated, bit_offset + 64));
	}

	out[0] = x; out[1] = y; out[2] = z;
	out[3] = quat_w_from_xyz(x, y, z);
}

/* unpack_animated_vector3 / unpack_single_animated_vector3, animated_track_cache.transform.h:871-990,992-1102 */
static void decode_animated_vector3(const transform_decoder* d, int k, int kind, uint32_t index, float out[3])
{
	const segment_cursor* c = &d->seg[k];
	const uint32_t bit_offset = animated_bit_offset(d, k, kind, index);

	if (!d->variable[kind])
	{
		out[0] = u32_as_f32(unpack_raw32(c->animated, bit_offset));
		out[1] = u32_as_f32(unpack_raw32(c->animated, bit_offset + 32));
		out[2] = u32_as_f32(unpack_raw32(c->animated, bit_offset + 64));
		return;
	}

	const uint32_t num_bits = c->format[kind][index];
	const uint8_t* range = c->range[kind] + index * 6;			/* AOS: 6 bytes per sub-track when segmented */
	float x, y, z;
	int ignore_segment = 0, ignore_clip = 0;

	if (num_bits == 0)
	{
		/* unpack_vector3_u48_unsafe, math/vector4_packing.h:628-653 (native little-endian u16 x3) */
		const float scale = 1.0f / 65535.0f;
		x = (float)rd_u16(range + 0) * scale;
		y = (float)rd_u16(range + 2) * scale;
		z = (float)rd_u16(range + 4) * scale;
		ignore_segment = 1;
	}
	else if (num_bits == d->raw_marker)
	{
		x = u32_as_f32(unpack_raw32(c->animated, bit_offset));
		y = u32_as_f32(unpack_raw32(c->animated, bit_offset + 32));
		z = u32_as_f32(unpack_raw32(c->animated, bit_offset + 64));
		ignore_segment = 1;
		ignore_clip = 1;
	}
	else
	{
		const float inv_max = inv_max_value(num_bits);
		x = (float)unpack_bits(c->animated, bit_offset, num_bits) * inv_max;
		y = (float)unpack_bits(c->animated, bit_offset + num_bits, num_bits) * inv_max;
		z = (float)unpack_bits(c->animated, bit_offset + num_bits * 2, num_bits) * inv_max;
	}

	if (d->has_segments && !ignore_segment)
	{
		/* unpack_vector3_u24_unsafe min then extent, math/vector4_packing.h:781-818 */
		const float n = 1.0f / 255.0f;
		const float min_x = (float)range[0] * n, min_y = (float)range[1] * n, min_z = (float)range[2] * n;
		const float ext_x = (float)range[3] * n, ext_y = (float)range[4] * n, ext_z = (float)range[5] * n;
		x = x * ext_x + min_x;
		y = y * ext_y + min_y;
		z = z * ext_z + min_z;
	}

	if (!ignore_clip)
	{
		const uint8_t* r = d->clip_range[kind] + index * 24;		/* min xyz, extent xyz (:949-958) */
		x = x * rd_f32(r + 12) + rd_f32(r + 0);
		y = y * rd_f32(r + 16) + rd_f32(r + 4);
		z = z * rd_f32(r + 20) + rd_f32(r + 8);
	}

	out[0] = x; out[1] = y; out[2] = z;
}

/* Constant rotation sample `index` (constant_track_cache_v0::unpack_rotation_group, constant_track_cache.transform.h:112-205) */
static void decode_constant_rotation(const transform_decoder* d, uint32_t index, int use_rtm_normalize, float out[4])
{
	const clip_view* v = &d->v;
	if (v->rotation_format == ACLO_ROT_FULL)
	{
		const uint8_t* p = d->constant[0] + index * 16;
		out[0] = rd_f32(p); out[1] = rd_f32(p + 4); out[2] = rd_f32(p + 8); out[3] = rd_f32(p + 12);
		return;
	}
	const uint32_t group = index / 4;
	const uint32_t lane = index % 4;
	const uint32_t remaining = v->num_constant[0] - group * 4;
	const uint32_t group_size = remaining < 4 ? remaining : 4;
	const uint8_t* p = d->constant[0] + group * 48 + lane * 4;
	const float x = rd_f32(p + group_size * 4 * 0);
	const float y = rd_f32(p + group_size * 4 * 1);
	const float z = rd_f32(p + group_size * 4 * 2);
	out[0] = x; out[1] = y; out[2] = z;
	out[3] = quat_w_from_xyz(x, y, z);
	if (d->settings->normalization == ACLO_NORMALIZE_ALWAYS)
	{
		if (use_rtm_normalize)
			rtm_quat_normalize(out);		/* unpack_rotation_within_group, :232-264 */
		else
			quat_normalize4(out);
	}
}

static uint32_t sub_track_type(const clip_view* v, uint32_t kind, uint32_t track)
{
	/* packed_sub_track_types: 2 bits per sub-track, MSB first, rotations | translations | scales
	 * (compressed_headers.h:214-224, decompression.transform.h:1551-1564) */
	const uint32_t num_entries = (v->num_tracks + 15) / 16;
	const uint8_t* types = v->th + v->sub_track_types_offset + 4 * (kind * num_entries + track / 16);
	return (rd_u32(types) >> ((15 - track % 16) * 2)) & 3;
}

static uint32_t track_rounding_policy(const aclo_settings* settings, uint32_t seek_policy, uint32_t track)
{
	/* track_writer::get_rounding_policy, core/track_writer.h:90 (+ the per-track override a writer may provide) */
	if (seek_policy != ACLO_ROUND_PER_TRACK || settings->per_track_rounding_policies == NULL)
		return seek_policy;
	return settings->per_track_rounding_policies[track];
}

static void write_default(const aclo_settings* settings, uint32_t mode, uint32_t kind, uint32_t track, float legacy_scale, float* out_track)
{
	/* unpack_default_*_sub_tracks, decompression.transform.h:574-675,881-983,1201-1310 */
	const uint32_t n = kind == 0 ? 4 : 3;
	float* dst = out_track + kind * 4;
	if (mode == ACLO_DEFAULT_SKIPPED)
		return;
	for (uint32_t i = 0; i < n; ++i)
	{
		if (mode == ACLO_DEFAULT_CONSTANT)
			dst[i] = settings->constant_defaults[kind * 4 + i];
		else if (mode == ACLO_DEFAULT_VARIABLE)
			dst[i] = settings->variable_defaults[track * 12 + kind * 4 + i];
		else
			dst[i] = legacy_scale;
	}
}

static uint32_t default_mode_of(const aclo_settings* settings, uint32_t kind)
{
	return kind == 0 ? settings->default_rotation_mode : (kind == 1 ? settings->default_translation_mode : settings->default_scale_mode);
}

int aclo_transform_decompress_tracks(const void* blob, const aclo_settings* settings, const aclo_seek_state* st, float* out)
{
	/* decompress_tracks_v0, decompression.transform.h:1526-1737. The reference walks nine passes; every
	 * sub-track is independent so we walk bones once and keep the six running indices instead. */
	transform_decoder d;
	decoder_init(&d, blob, settings, st);
	const clip_view* v = &d.v;
	if (v->track_type != ACLO_TRACK_QVVF)
		return -1;
	if (v->num_tracks == 0 || st->sample_time < 0.0f)
		return 0;

	const float alpha = st->interpolation_alpha;
	const float legacy_scale = (float)v->default_scale;					/* :1548 */
	const int interpolate = should_interpolate(settings, v->rotation_format, alpha);
	uint32_t constant_index[3] = { 0, 0, 0 };
	uint32_t animated_index[3] = { 0, 0, 0 };

	for (uint32_t track = 0; track < v->num_tracks; ++track)
	{
		float* out_track = out + (size_t)track * 12;
		const uint32_t policy = settings->per_track_rounding ? track_rounding_policy(settings, st->rounding_policy, track) : ACLO_ROUND_NONE;

		/* ---- rotation ---- */
		const uint32_t rot_type = sub_track_type(v, 0, track);
		if (rot_type == 0)
			write_default(settings, settings->default_rotation_mode, 0, track, 0.0f, out_track);
		else if (rot_type == 1)
			decode_constant_rotation(&d, constant_index[0]++, 0, out_track);
		else
		{
			/* animated_track_cache_v0::unpack_rotation_group, animated_track_cache.transform.h:1316-1662 */
			float s0[4], s1[4], result[4];
			const uint32_t index = animated_index[0]++;
			decode_animated_rotation(&d, 0, index, 1, s0);
			decode_animated_rotation(&d, 1, index, 1, s1);

			if (v->rotation_format != ACLO_ROT_FULL && settings->normalization == ACLO_NORMALIZE_ALWAYS)
			{
				if (settings->per_track_rounding || !interpolate)			/* :1463-1474 */
				{
					quat_normalize4(s0);
					quat_normalize4(s1);
				}
			}

			if (settings->per_track_rounding)
			{
				/* :1481-1596 then consume_rotation(policy) :1767-1772 */
				if (policy == ACLO_ROUND_FLOOR)
					memcpy(result, s0, sizeof(result));
				else if (policy == ACLO_ROUND_CEIL)
					memcpy(result, s1, sizeof(result));
				else if (policy == ACLO_ROUND_NEAREST)
					memcpy(result, alpha < 0.5f ? s0 : s1, sizeof(result));
				else
				{
					quat_lerp_no_normalization4(s0, s1, alpha, result);
					if (settings->normalization >= ACLO_NORMALIZE_LERP_ONLY)
						quat_normalize4(result);
				}
			}
			else if (interpolate)
			{
				quat_lerp_no_normalization4(s0, s1, alpha, result);			/* :1604-1616 */
				if (settings->normalization >= ACLO_NORMALIZE_LERP_ONLY)
					quat_normalize4(result);
			}
			else
				memcpy(result, alpha <= 0.0f ? s0 : s1, sizeof(result));		/* :1617-1627 */

			memcpy(out_track, result, sizeof(result));
		}

		/* ---- translation, scale ---- */
		for (uint32_t kind = 1; kind <= 2; ++kind)
		{
			float* dst = out_track + kind * 4;
			if (kind == 2 && !v->has_scale)
			{
				/* no scale in the clip: every bone gets the default (decompression.transform.h:1653-1680) */
				write_default(settings, settings->default_scale_mode, 2, track, legacy_scale, out_track);
				continue;
			}

			const uint32_t type = sub_track_type(v, kind, track);
			if (type == 0)
				write_default(settings, default_mode_of(settings, kind), kind, track, legacy_scale, out_track);
			else if (type == 1)
			{
				const uint8_t* p = d.constant[kind] + 12u * constant_index[kind]++;
				dst[0] = rd_f32(p); dst[1] = rd_f32(p + 4); dst[2] = rd_f32(p + 8);
			}
			else
			{
				/* unpack_translation_group / unpack_scale_group, animated_track_cache.transform.h:1774-1836,1896-1958 */
				float s0[3], s1[3];
				const uint32_t index = animated_index[kind]++;
				decode_animated_vector3(&d, 0, (int)kind, index, s0);
				decode_animated_vector3(&d, 1, (int)kind, index, s1);
				for (int i = 0; i < 3; ++i)
				{
					if (settings->per_track_rounding && policy == ACLO_ROUND_FLOOR)
						dst[i] = s0[i];
					else if (settings->per_track_rounding && policy == ACLO_ROUND_CEIL)
						dst[i] = s1[i];
					else if (settings->per_track_rounding && policy == ACLO_ROUND_NEAREST)
						dst[i] = alpha < 0.5f ? s0[i] : s1[i];
					else
						dst[i] = lerpf(s0[i], s1[i], alpha);
				}
			}
		}
	}
	return 0;
}

int aclo_transform_decompress_track(const void* blob, const aclo_settings* settings, const aclo_seek_state* st, uint32_t track_index, float* out)
{
	/* decompress_track_v0, decompression.transform.h:1753-2050. Writes row `track_index` of out[num_tracks][12]. */
	transform_decoder d;
	decoder_init(&d, blob, settings, st);
	const clip_view* v = &d.v;
	if (v->track_type != ACLO_TRACK_QVVF)
		return -1;
	if (v->num_tracks == 0 || st->sample_time < 0.0f || track_index >= v->num_tracks)
		return 0;

	float* out_track = out + (size_t)track_index * 12;
	const float legacy_scale = (float)v->default_scale;

	/* rank of this bone among constant / animated sub-tracks of each kind (:1873-1891) */
	uint32_t constant_index[3] = { 0, 0, 0 };
	uint32_t animated_index[3] = { 0, 0, 0 };
	for (uint32_t kind = 0; kind < (v->has_scale ? 3u : 2u); ++kind)
		for (uint32_t track = 0; track < track_index; ++track)
		{
			const uint32_t type = sub_track_type(v, kind, track);
			constant_index[kind] += type == 1;
			animated_index[kind] += type == 2;
		}

	float alpha = st->interpolation_alpha;
	if (settings->per_track_rounding)
		alpha = apply_rounding_policy(alpha, track_rounding_policy(settings, st->rounding_policy, track_index));	/* :1975-1983 */

	/* rotation */
	const uint32_t rot_type = sub_track_type(v, 0, track_index);
	if (rot_type == 0)
		write_default(settings, settings->default_rotation_mode, 0, track_index, 0.0f, out_track);
	else if (rot_type == 1)
		decode_constant_rotation(&d, constant_index[0], 1, out_track);
	else
	{
		/* unpack_rotation_within_group, animated_track_cache.transform.h:1709-1765 */
		float s0[4], s1[4], result[4];
		decode_animated_rotation(&d, 0, animated_index[0], 0, s0);
		decode_animated_rotation(&d, 1, animated_index[0], 0, s1);
		if (should_interpolate(settings, v->rotation_format, alpha))
		{
			if (settings->normalization >= ACLO_NORMALIZE_LERP_ONLY)
				rtm_quat_lerp(s0, s1, alpha, result);
			else
				acl_quat_lerp_no_normalization(s0, s1, alpha, result);
		}
		else
		{
			memcpy(result, alpha <= 0.0f ? s0 : s1, sizeof(result));
			if (settings->normalization == ACLO_NORMALIZE_ALWAYS && v->rotation_format != ACLO_ROT_FULL)
				rtm_quat_normalize(result);
		}
		memcpy(out_track, result, sizeof(result));
	}

	/* translation, scale */
	for (uint32_t kind = 1; kind <= 2; ++kind)
	{
		float* dst = out_track + kind * 4;
		const uint32_t type = (kind == 2 && !v->has_scale) ? 0u : sub_track_type(v, kind, track_index);	/* :1806-1822 */
		if (type == 0)
			write_default(settings, default_mode_of(settings, kind), kind, track_index, legacy_scale, out_track);
		else if (type == 1)
		{
			const uint8_t* p = d.constant[kind] + 12u * constant_index[kind];
			dst[0] = rd_f32(p); dst[1] = rd_f32(p + 4); dst[2] = rd_f32(p + 8);
		}
		else
		{
			float s0[3], s1[3];
			decode_animated_vector3(&d, 0, (int)kind, animated_index[kind], s0);
			decode_animated_vector3(&d, 1, (int)kind, animated_index[kind], s1);
			for (int i = 0; i < 3; ++i)
				dst[i] = lerpf(s0[i], s1[i], alpha);						/* :1879-1887 */
		}
	}
	return 0;
}

int aclo_transform_extract_key_frame(const void* blob, const aclo_seek_state* st, uint32_t which, uint32_t* out_ints)
{
	aclo_settings settings;
	aclo_default_settings(&settings);
	transform_decoder d;
	decoder_init(&d, blob, &settings, st);
	const clip_view* v = &d.v;
	if (v->track_type != ACLO_TRACK_QVVF || which > 1)
		return -1;
	const segment_cursor* c = &d.seg[which];
	uint32_t row = 0;
	for (int kind = 0; kind < (v->has_scale ? 3 : 2); ++kind)
	{
		for (uint32_t index = 0; index < v->num_animated[kind]; ++index, ++row)
		{
			uint32_t* dst = out_ints + (size_t)row * 4;
			const uint32_t bit_offset = animated_bit_offset(&d, (int)which, kind, index);
			if (!d.variable[kind])
			{
				dst[0] = unpack_raw32(c->animated, bit_offset);
				dst[1] = unpack_raw32(c->animated, bit_offset + 32);
				dst[2] = unpack_raw32(c->animated, bit_offset + 64);
				dst[3] = 0xFFFFFFFFu;
				continue;
			}
			const uint32_t num_bits = c->format[kind][index];
			dst[3] = num_bits;
			if (num_bits == 0)
			{
				if (kind == 0)
				{
					const uint8_t* r = c->range[0] + (index / 4) * 24 + index % 4;
					dst[0] = ((uint32_t)r[0] << 8) | r[4];
					dst[1] = ((uint32_t)r[8] << 8) | r[12];
					dst[2] = ((uint32_t)r[16] << 8) | r[20];
				}
				else
				{
					const uint8_t* r = c->range[kind] + index * 6;
					dst[0] = rd_u16(r); dst[1] = rd_u16(r + 2); dst[2] = rd_u16(r + 4);
				}
			}
			else if (num_bits == d.raw_marker)
			{
				dst[0] = unpack_raw32(c->animated, bit_offset);
				dst[1] = unpack_raw32(c->animated, bit_offset + 32);
				dst[2] = unpack_raw32(c->animated, bit_offset + 64);
			}
			else
			{
				dst[0] = unpack_bits(c->animated, bit_offset, num_bits);
				dst[1] = unpack_bits(c->animated, bit_offset + num_bits, num_bits);
				dst[2] = unpack_bits(c->animated, bit_offset + num_bits * 2, num_bits);
			}
		}
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Scalar tracks (decompression/impl/decompression.scalar.h)
 * ---------------------------------------------------------------------------------------------- */

static const uint8_t k_bit_rate_num_bits_v0[] = { 0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 32 };	/* core/impl/variable_bit_rates.h:39-40 */
static const uint8_t k_bit_rate_num_bits[] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 32 };	/* :42-43 */

static uint32_t scalar_num_components(uint32_t track_type)
{
	return track_type <= 3 ? track_type + 1 : 4;	/* float1f..float4f, vector4f: core/track_types.h */
}

int aclo_scalar_seek(const void* blob, const aclo_settings* settings, float sample_time,
	uint32_t rounding_policy, uint32_t looping_policy, aclo_scalar_seek_state* st)
{
	/* seek_v0, decompression.scalar.h:181-210 */
	clip_view v;
	if (view_clip(blob, &v) != 0 || v.track_type == ACLO_TRACK_QVVF)
		return -1;
	memset(st, 0, sizeof(*st));
	st->sample_time = -1.0f;
	if (v.num_samples == 0)
		return 0;

	uint32_t policy;
	float duration;
	resolve_looping(&v, settings, looping_policy, &policy, &duration);
	st->duration = duration;
	st->looping_policy = policy;
	if (settings->clamp_sample_time)
		sample_time = clampf(sample_time, 0.0f, duration);
	st->sample_time = sample_time;
	st->rounding_policy = rounding_policy;

	find_linear_interpolation_samples_with_sample_rate(v.num_samples, v.sample_rate, sample_time, rounding_policy, policy,
		&st->key_frames[0], &st->key_frames[1], &st->interpolation_alpha);

	const uint32_t num_bits_per_frame = rd_u32(v.blob + 32);		/* scalar_tracks_header, compressed_headers.h:140-150 */
	st->key_frame_bit_offsets[0] = st->key_frames[0] * num_bits_per_frame;
	st->key_frame_bit_offsets[1] = st->key_frames[1] * num_bits_per_frame;
	return 0;
}

static int scalar_decode(const void* blob, const aclo_settings* settings, const aclo_scalar_seek_state* st, int32_t only_track, float* out)
{
	/* decompress_tracks_v0 / decompress_track_v0, decompression.scalar.h:212-481,483-705 */
	clip_view v;
	if (view_clip(blob, &v) != 0 || v.track_type == ACLO_TRACK_QVVF)
		return -1;
	if (v.num_tracks == 0 || st->sample_time < 0.0f)
		return 0;
	if (only_track >= 0 && (uint32_t)only_track >= v.num_tracks)
		return 0;

	const uint8_t* sh = v.blob + 32;
	const uint8_t* metadata = sh + rd_u32(sh + 4);
	const uint8_t* constant_values = sh + rd_u32(sh + 8);
	const uint8_t* range_values = sh + rd_u32(sh + 12);
	const uint8_t* animated_values = sh + rd_u32(sh + 16);
	const uint8_t* num_bits_table = v.version == ACLO_VERSION_FIRST ? k_bit_rate_num_bits_v0 : k_bit_rate_num_bits;
	const uint32_t nc = scalar_num_components(v.track_type);

	uint32_t bit_offset0 = st->key_frame_bit_offsets[0];
	uint32_t bit_offset1 = st->key_frame_bit_offsets[1];

	for (uint32_t track = 0; track < v.num_tracks; ++track)
	{
		const uint32_t num_bits = num_bits_table[metadata[track]];
		const int wanted = only_track < 0 || (uint32_t)only_track == track;

		float alpha = st->interpolation_alpha;
		if (settings->per_track_rounding)
			alpha = apply_rounding_policy(st->interpolation_alpha, track_rounding_policy(settings, st->rounding_policy, track));

		if (num_bits == 0)
		{
			if (wanted)
				for (uint32_t c = 0; c < nc; ++c)
					out[(size_t)track * 4 + c] = rd_f32(constant_values + 4 * c);
			constant_values += 4 * nc;
		}
		else
		{
			if (wanted)
			{
				for (uint32_t c = 0; c < nc; ++c)
				{
					float v0, v1;
					if (num_bits == 32)
					{
						v0 = u32_as_f32(unpack_raw32(animated_values, bit_offset0 + 32 * c));
						v1 = u32_as_f32(unpack_raw32(animated_values, bit_offset1 + 32 * c));
					}
					else
					{
						const float inv_max = inv_max_value(num_bits);
						v0 = (float)unpack_bits(animated_values, bit_offset0 + num_bits * c, num_bits) * inv_max;
						v1 = (float)unpack_bits(animated_values, bit_offset1 + num_bits * c, num_bits) * inv_max;
						const float range_min = rd_f32(range_values + 4 * c);
						const float range_extent = rd_f32(range_values + 4 * (nc + c));
						v0 = v0 * range_extent + range_min;
						v1 = v1 * range_extent + range_min;
					}
					out[(size_t)track * 4 + c] = lerpf(v0, v1, alpha);
				}
			}
			if (num_bits != 32)
				range_values += 8 * nc;
			bit_offset0 += num_bits * nc;
			bit_offset1 += num_bits * nc;
		}

		if (only_track >= 0 && (uint32_t)only_track == track)
			break;
	}
	return 0;
}

int aclo_scalar_decompress_tracks(const void* blob, const aclo_settings* settings, const aclo_scalar_seek_state* st, float* out)
{
	return scalar_decode(blob, settings, st, -1, out);
}

int aclo_scalar_decompress_track(const void* blob, const aclo_settings* settings, const aclo_scalar_seek_state* st, uint32_t track_index, float* out)
{
	return scalar_decode(blob, settings, st, (int32_t)track_index, out);
}

/* ------------------------------------------------------------------------------------------------
 * Touched bytes accounting and timing
 * ---------------------------------------------------------------------------------------------- */

int aclo_transform_touched_bytes(const void* blob, uint32_t segment_index, uint64_t out[4])
{
	/* SURVEY.md 8(d) formula == the reference's decomp_touched_bytes (compression/impl/write_stats.h:115-120)
	 * split into clip / segment / key frame parts */
	clip_view v;
	if (view_clip(blob, &v) != 0 || v.track_type != ACLO_TRACK_QVVF)
		return -1;
	if (segment_index >= v.num_segments)
		return -2;
	const uint32_t num_entries = (v.num_tracks + 15) / 16;
	const int rot_variable = v.rotation_format == ACLO_ROT_DROP_W_VARIABLE;
	uint64_t clip_bytes = 84;																/* raw_buffer_header + tracks_header + transform_tracks_header */
	if (v.num_segments > 1)
		clip_bytes += 4ull * (v.num_segments + 1);											/* segment_start_indices + sentinel */
	clip_bytes += 4ull * num_entries * (v.has_scale ? 3 : 2);								/* sub-track types */
	clip_bytes += (v.rotation_format == ACLO_ROT_FULL ? 16ull : 12ull) * v.num_constant[0];
	clip_bytes += 12ull * (v.num_constant[1] + v.num_constant[2]);
	if (rot_variable)
		clip_bytes += 24ull * v.num_animated[0];
	if (v.translation_format == ACLO_VEC_VARIABLE)
		clip_bytes += 24ull * v.num_animated[1];
	if (v.has_scale && v.scale_format == ACLO_VEC_VARIABLE)
		clip_bytes += 24ull * v.num_animated[2];
	out[0] = clip_bytes;
	uint64_t segment_bytes = v.segment_header_size + v.num_animated_variable_sub_tracks;
	if (v.num_segments > 1)
		segment_bytes += 6ull * v.num_animated_variable_sub_tracks;
	out[1] = segment_bytes;
	const uint32_t pose_bits = rd_u32(v.th + v.segment_headers_offset + v.segment_header_size * segment_index);
	out[2] = (pose_bits + 7) / 8;
	out[3] = 40ull * v.num_tracks;
	return 0;
}

int aclo_scalar_touched_bytes(const void* blob, uint64_t out[4])
{
	clip_view v;
	if (view_clip(blob, &v) != 0 || v.track_type == ACLO_TRACK_QVVF)
		return -1;
	const uint8_t* sh = v.blob + 32;
	const uint8_t* metadata = sh + rd_u32(sh + 4);
	const uint8_t* table = v.version == ACLO_VERSION_FIRST ? k_bit_rate_num_bits_v0 : k_bit_rate_num_bits;
	const uint32_t nc = scalar_num_components(v.track_type);
	uint64_t num_constant = 0, num_ranged = 0;
	for (uint32_t track = 0; track < v.num_tracks; ++track)
	{
		const uint32_t bits = table[metadata[track]];
		num_constant += bits == 0;
		num_ranged += bits != 0 && bits != 32;
	}
	out[0] = 52ull + v.num_tracks + 4ull * nc * num_constant + 8ull * nc * num_ranged;
	out[1] = 0;
	out[2] = (rd_u32(sh) + 7) / 8;
	out[3] = 4ull * nc * v.num_tracks;
	return 0;
}

double aclo_bench_transform(const void* const* blobs, const uint32_t* request_clip, const float* request_time,
	uint32_t num_requests, uint32_t max_tracks)
{
	aclo_settings settings;
	aclo_default_settings(&settings);
	static float scratch[4096 * 12];
	if (max_tracks > 4096)
		return -1.0;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (uint32_t i = 0; i < num_requests; ++i)
	{
		aclo_seek_state st;
		const void* blob = blobs[request_clip[i]];
		aclo_transform_seek(blob, &settings, request_time[i], ACLO_ROUND_NONE, ACLO_LOOP_AS_COMPRESSED, &st);
		aclo_transform_decompress_tracks(blob, &settings, &st, scratch);
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------------
 * calculate_compression_error (compression/impl/track_error.impl.h:166-392) with the
 * qvvf_transform_error_metric (compression/transform_error_metrics.h:281-385). rtm paths are relative to
 * /root/reference/external/rtm/includes/rtm, SSE2 code paths (what oracle/_ref is compiled with).
 * ---------------------------------------------------------------------------------------------- */

/* rtm::quat_mul, quatf.h:498-545 (SSE2): ((r.w*l + s0*(r.x*l_wzyx)) + (s1*(r.y*l_zwxy) + s2*(r.z*l_yxwz))), signs applied by xor */
static void rtm_quat_mul(const float l[4], const float r[4], float out[4])
{
	const float a0 = r[3] * l[0], a1 = r[3] * l[1], a2 = r[3] * l[2], a3 = r[3] * l[3];
	const float b0 = r[0] * l[3], b1 = -(r[0] * l[2]), b2 = r[0] * l[1], b3 = -(r[0] * l[0]);
	const float c0 = r[1] * l[2], c1 = r[1] * l[3], c2 = -(r[1] * l[0]), c3 = -(r[1] * l[1]);
	const float d0 = -(r[2] * l[1]), d1 = r[2] * l[0], d2 = r[2] * l[3], d3 = -(r[2] * l[2]);
	out[0] = (a0 + b0) + (c0 + d0);
	out[1] = (a1 + b1) + (c1 + d1);
	out[2] = (a2 + b2) + (c2 + d2);
	out[3] = (a3 + b3) + (c3 + d3);
}

/* rtm::quat_mul_vector3, quatf.h:616-668 (SSE2): temp = conj(r) * (v, 0) without the W terms, result = temp * r */
static void rtm_quat_mul_vector3(const float v[3], const float r[4], float out[3])
{
	const float nx = -r[0], ny = -r[1], nz = -r[2];
	const float t0 = (v[0] * r[3] + v[1] * nz) + v[2] * r[1];
	const float t1 = (v[0] * r[2] + v[1] * r[3]) + v[2] * nx;
	const float t2 = (v[0] * ny + v[1] * r[0]) + v[2] * r[3];
	const float t3 = (v[0] * r[0] + v[1] * r[1]) + v[2] * r[2];
	out[0] = (r[3] * t0 + r[0] * t3) + (r[1] * t2 + nz * t1);
	out[1] = (r[3] * t1 + nx * t2) + (r[1] * t3 + r[2] * t0);
	out[2] = (r[3] * t2 + r[0] * t1) + (ny * t0 + r[2] * t3);
}

/* rtm::quat_normalize, quatf.h:917-953. mode 0: the SSE2 code (rsqrtss estimate + 2 Newton-Raphson steps; the estimate is CPU specific,
 * so this is bit-identical to the reference only on the CPU both run on); mode 1: IEEE 1 / sqrt (what the CUDA path computes). */
static void metric_quat_normalize(float q[4], int normalize_mode)
{
	if (normalize_mode == 0)
	{
		rtm_quat_normalize(q);
		return;
	}
	const float x2 = q[0] * q[0], y2 = q[1] * q[1], z2 = q[2] * q[2], w2 = q[3] * q[3];
	const float dot = (x2 + z2) + (y2 + w2);
	const float inv_len = 1.0f / sqrtf(dot);
	for (int i = 0; i < 4; ++i)
		q[i] = q[i] * inv_len;
}

/* rtm::matrix_from_qvv, matrix3x4f.h:134-159: rows x_axis, y_axis, z_axis (xyz), w_axis = translation */
static void rtm_matrix_from_qvv(const float q[12], float m[4][3])
{
	const float x2 = q[0] + q[0], y2 = q[1] + q[1], z2 = q[2] + q[2];
	const float xx = q[0] * x2, xy = q[0] * y2, xz = q[0] * z2;
	const float yy = q[1] * y2, yz = q[1] * z2, zz = q[2] * z2;
	const float wx = q[3] * x2, wy = q[3] * y2, wz = q[3] * z2;
	m[0][0] = (1.0f - (yy + zz)) * q[8];	m[0][1] = (xy + wz) * q[8];				m[0][2] = (xz - wy) * q[8];
	m[1][0] = (xy - wz) * q[9];				m[1][1] = (1.0f - (xx + zz)) * q[9];	m[1][2] = (yz + wx) * q[9];
	m[2][0] = (xz + wy) * q[10];			m[2][1] = (yz - wx) * q[10];			m[2][2] = (1.0f - (xx + yy)) * q[10];
	m[3][0] = q[4];							m[3][1] = q[5];							m[3][2] = q[6];
}

/* rtm_impl::quat_from_matrix, impl/matrix_affine_common.h:153-227 (float overloads of scalar_sqrt_reciprocal / scalar_reciprocal:
 * 1 / sqrt, 1 / x, scalarf.h:294-325), ending in rtm::quat_normalize */
static void rtm_quat_from_matrix(float m[3][3], int normalize_mode, float out[4])
{
	for (int axis = 0; axis < 3; ++axis)
		if (fabsf(m[axis][0]) <= 0.00001f && fabsf(m[axis][1]) <= 0.00001f && fabsf(m[axis][2]) <= 0.00001f)
		{
			out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; out[3] = 1.0f;		/* Zero scale not supported, return the identity */
			return;
		}
	const float trace = m[0][0] + m[1][1] + m[2][2];
	float q[4];
	if (trace > 0.0f)
	{
		const float inv_trace = 1.0f / sqrtf(trace + 1.0f);
		const float half_inv_trace = inv_trace * 0.5f;
		q[0] = (m[1][2] - m[2][1]) * half_inv_trace;
		q[1] = (m[2][0] - m[0][2]) * half_inv_trace;
		q[2] = (m[0][1] - m[1][0]) * half_inv_trace;
		q[3] = (1.0f / inv_trace) * 0.5f;
	}
	else
	{
		int best = 0;
		if (m[1][1] > m[0][0])
			best = 1;
		if (m[2][2] > m[best][best])
			best = 2;
		const int next = (best + 1) % 3;
		const int next_next = (next + 1) % 3;
		const float pseudo_trace = 1.0f + m[best][best] - m[next][next] - m[next_next][next_next];
		const float inv_pseudo_trace = 1.0f / sqrtf(pseudo_trace);
		const float half_inv_pseudo_trace = inv_pseudo_trace * 0.5f;
		q[best] = (1.0f / inv_pseudo_trace) * 0.5f;
		q[next] = half_inv_pseudo_trace * (m[best][next] + m[next][best]);
		q[next_next] = half_inv_pseudo_trace * (m[best][next_next] + m[next_next][best]);
		q[3] = half_inv_pseudo_trace * (m[next][next_next] - m[next_next][next]);
	}
	metric_quat_normalize(q, normalize_mode);
	memcpy(out, q, sizeof(q));
}

/* rtm::matrix_mul(lhs, rhs), matrix3x4f.h:298-321: per row tmp = l.x * r.x_axis; tmp = l.y * r.y_axis + tmp; tmp = l.z * r.z_axis + tmp
 * (vector_mul_add = (v0 * v1) + v2 on SSE2); the translation row adds r.w_axis */
static void rtm_matrix_mul(float l[4][3], float r[4][3], float out[4][3])
{
	float m[4][3];
	for (int row = 0; row < 4; ++row)
		for (int c = 0; c < 3; ++c)
		{
			float tmp = l[row][0] * r[0][c];
			tmp = l[row][1] * r[1][c] + tmp;
			tmp = l[row][2] * r[2][c] + tmp;
			m[row][c] = row == 3 ? r[3][c] + tmp : tmp;
		}
	memcpy(out, m, sizeof(m));
}

/* The negative scale branch of rtm::qvv_mul, qvvf.h:320-345: through matrices (matrix_mul matrix3x4f.h:298-321 with
 * vector_mul_add = (v0 * v1) + v2 on SSE2, matrix_remove_scale :636-644 = vector_normalize3(axis, axis, 1e-8) vector4f.h:2310-2318,
 * the result's sign bits xor-ed onto the axes) */
static void qvv_mul_negative_scale(const float lhs[12], const float rhs[12], int normalize_mode, float out[12])
{
	float l[4][3], r[4][3], m[4][3];
	rtm_matrix_from_qvv(lhs, l);
	rtm_matrix_from_qvv(rhs, r);
	rtm_matrix_mul(l, r, m);
	float scale[3];
	for (int i = 0; i < 3; ++i)
		scale[i] = lhs[8 + i] * rhs[8 + i];
	for (int axis = 0; axis < 3; ++axis)
	{
		const float len_sq = (m[axis][0] * m[axis][0] + m[axis][1] * m[axis][1]) + m[axis][2] * m[axis][2];
		if (len_sq >= 1.0e-8f)
		{
			const float inv_len = 1.0f / sqrtf(len_sq);
			for (int c = 0; c < 3; ++c)
				m[axis][c] = m[axis][c] * inv_len;
		}
		const uint32_t sign = f32_as_u32(scale[axis]) & 0x80000000u;
		for (int c = 0; c < 3; ++c)
			m[axis][c] = u32_as_f32(f32_as_u32(m[axis][c]) ^ sign);
	}
	float result[12];
	rtm_quat_from_matrix(m, normalize_mode, result);
	for (int i = 0; i < 3; ++i)
	{
		result[4 + i] = m[3][i];
		result[8 + i] = scale[i];
	}
	result[7] = 0.0f;
	result[11] = 0.0f;
	memcpy(out, result, sizeof(result));
}

/* rtm::qvv_normalize(rtm::qvv_mul(local, parent_object)), qvvf.h:315-355,426-430. Returns 1 when the negative scale branch (through
 * matrices) was taken. */
static int qvv_mul_normalize(const float local[12], const float parent[12], int normalize_mode, float out[12])
{
	int negative = 0;
	for (int i = 0; i < 3; ++i)
	{
		const float min_scale = local[8 + i] < parent[8 + i] ? local[8 + i] : parent[8 + i];	/* _mm_min_ps(lhs, rhs) */
		negative |= min_scale < 0.0f;
	}
	float result[12];
	if (negative)
		qvv_mul_negative_scale(local, parent, normalize_mode, result);
	else
	{
		rtm_quat_mul(local + 0, parent + 0, result);
		const float scaled[3] = { local[4] * parent[8], local[5] * parent[9], local[6] * parent[10] };
		float rotated[3];
		rtm_quat_mul_vector3(scaled, parent + 0, rotated);
		for (int i = 0; i < 3; ++i)
		{
			result[4 + i] = rotated[i] + parent[4 + i];
			result[8 + i] = local[8 + i] * parent[8 + i];
		}
		result[7] = 0.0f;
		result[11] = 0.0f;
	}
	metric_quat_normalize(result, normalize_mode);		/* qvv_normalize, qvvf.h:426-430 */
	memcpy(out, result, sizeof(result));
	return negative;
}

/* rtm::qvv_mul(lhs, rhs), qvvf.h:315-355, no normalisation on top. Returns 1 when the negative scale branch was taken. */
static int qvv_mul_plain(const float lhs[12], const float rhs[12], int normalize_mode, float out[12])
{
	int negative = 0;
	for (int i = 0; i < 3; ++i)
	{
		const float min_scale = lhs[8 + i] < rhs[8 + i] ? lhs[8 + i] : rhs[8 + i];
		negative |= min_scale < 0.0f;
	}
	float result[12];
	if (negative)
		qvv_mul_negative_scale(lhs, rhs, normalize_mode, result);
	else
	{
		rtm_quat_mul(lhs + 0, rhs + 0, result);
		const float scaled[3] = { lhs[4] * rhs[8], lhs[5] * rhs[9], lhs[6] * rhs[10] };
		float rotated[3];
		rtm_quat_mul_vector3(scaled, rhs + 0, rotated);
		for (int i = 0; i < 3; ++i)
		{
			result[4 + i] = rotated[i] + rhs[4 + i];
			result[8 + i] = lhs[8 + i] * rhs[8 + i];
		}
		result[7] = 0.0f;
		result[11] = 0.0f;
	}
	memcpy(out, result, sizeof(result));
	return negative;
}

/* acl::apply_additive_to_base(format, base, additive), core/additive_utils.h:131-167, over a pose, in place on `pose` (the additive one):
 * 0 none, 1 relative = qvv_mul(additive, base), 2 additive0 (scale = additive * base), 3 additive1 (scale = (1 + additive) * base).
 * Returns 1 when `relative` went through rtm::qvv_mul's negative scale branch (whose quat_from_matrix normalises: normalize_mode). */
int aclo_apply_additive_to_base(uint32_t additive_format, const float* base_pose, float* pose, uint32_t num_tracks, int normalize_mode)
{
	int negative = 0;
	for (uint32_t bone = 0; bone < num_tracks; ++bone)
	{
		const float* base = base_pose + (size_t)bone * 12;
		float* additive = pose + (size_t)bone * 12;
		if (additive_format == 1)
			negative |= qvv_mul_plain(additive, base, normalize_mode, additive);
		else if (additive_format == 2 || additive_format == 3)
		{
			float rotation[4];
			rtm_quat_mul(additive + 0, base + 0, rotation);
			for (int i = 0; i < 4; ++i)
				additive[i] = rotation[i];
			for (int i = 0; i < 3; ++i)
			{
				additive[4 + i] = additive[4 + i] + base[4 + i];
				additive[8 + i] = additive_format == 2 ? additive[8 + i] * base[8 + i] : (1.0f + additive[8 + i]) * base[8 + i];
			}
		}
	}
	return negative;
}

/* qvvf_transform_error_metric::local_to_object_space, transform_error_metrics.h:289-310 (all transforms dirty, in index order) */
int aclo_local_to_object_space(const float* local_pose, const uint32_t* parent_indices, uint32_t num_tracks, int normalize_mode, float* out_object_pose)
{
	int negative = 0;
	for (uint32_t bone = 0; bone < num_tracks; ++bone)
	{
		const uint32_t parent = parent_indices[bone];
		if (parent == 0xFFFFFFFFu)
			memcpy(out_object_pose + (size_t)bone * 12, local_pose + (size_t)bone * 12, 12 * sizeof(float));
		else if (parent >= bone)
			return -1;		/* the reference reads a stale / unwritten parent here: not a valid skeleton order */
		else
			negative |= qvv_mul_normalize(local_pose + (size_t)bone * 12, out_object_pose + (size_t)parent * 12, normalize_mode, out_object_pose + (size_t)bone * 12);
	}
	return negative;
}

/* rtm::qvv_mul_point3, qvvf.h:370-373 */
static void qvv_mul_point3(const float point[3], const float qvv[12], float out[3])
{
	const float scaled[3] = { qvv[8] * point[0], qvv[9] * point[1], qvv[10] * point[2] };
	float rotated[3];
	rtm_quat_mul_vector3(scaled, qvv, rotated);
	for (int i = 0; i < 3; ++i)
		out[i] = rotated[i] + qvv[4 + i];
}

/* rtm::vector_distance3_as_scalar, vector4f.h:2260-2264 -> vector_dot3_as_scalar :1899-1906 (SSE2: (x2 + y2) + z2), scalar_sqrt = sqrtss */
static float vector_distance3(const float a[3], const float b[3])
{
	const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
	return sqrtf((dx * dx + dy * dy) + dz * dz);
}

static float sse_max_ss(float a, float b) { return a > b ? a : b; }		/* _mm_max_ss(a, b): b when unordered */

/* qvvf_transform_error_metric::calculate_error, transform_error_metrics.h:335-358 with construct_sphere_shell :261-266 */
float aclo_calculate_error(const float* raw_object_bone, const float* lossy_object_bone, float shell_distance)
{
	float errors[3];
	for (int axis = 0; axis < 3; ++axis)
	{
		float point[3] = { 0.0f, 0.0f, 0.0f };
		point[axis] = shell_distance;
		float raw_vtx[3], lossy_vtx[3];
		qvv_mul_point3(point, raw_object_bone, raw_vtx);
		qvv_mul_point3(point, lossy_object_bone, lossy_vtx);
		errors[axis] = vector_distance3(raw_vtx, lossy_vtx);
	}
	return sse_max_ss(sse_max_ss(errors[0], errors[1]), errors[2]);
}

/* qvvf_matrix3x4f_transform_error_metric (transform_error_metrics.h:389-464): convert_transforms = matrix_from_qvv per bone,
 * local_to_object_space = matrix_mul(local, parent object) (roots copied), into out_object [num_tracks][4][3] */
static int matrix_local_to_object_space(const float* local_pose, const uint32_t* parent_indices, uint32_t num_tracks, float* out_object)
{
	for (uint32_t bone = 0; bone < num_tracks; ++bone)
	{
		float local[4][3];
		rtm_matrix_from_qvv(local_pose + (size_t)bone * 12, local);
		float (*object)[3] = (float (*)[3])(out_object + (size_t)bone * 12);
		const uint32_t parent = parent_indices[bone];
		if (parent == 0xFFFFFFFFu)
			memcpy(object, local, sizeof(local));
		else if (parent >= bone)
			return -1;
		else
			rtm_matrix_mul(local, (float (*)[3])(out_object + (size_t)parent * 12), object);
	}
	return 0;
}

/* calculate_error of the matrix metric (:438-463): rtm::matrix_mul_point3 (matrix3x4f.h:326-336: tmp0 = p.x * x_axis; tmp0 = p.y * y_axis + tmp0;
 * tmp1 = p.z * z_axis + w_axis; tmp0 + tmp1) of the three shell points, distances, the largest */
static float matrix_calculate_error(const float* raw_object, const float* lossy_object, float shell_distance)
{
	float errors[3];
	for (int axis = 0; axis < 3; ++axis)
	{
		float point[3] = { 0.0f, 0.0f, 0.0f };
		point[axis] = shell_distance;
		float vtx[2][3];
		for (int stream = 0; stream < 2; ++stream)
		{
			const float (*m)[3] = (const float (*)[3])(stream == 0 ? raw_object : lossy_object);
			for (int c = 0; c < 3; ++c)
			{
				float tmp0 = point[0] * m[0][c];
				tmp0 = point[1] * m[1][c] + tmp0;
				const float tmp1 = point[2] * m[2][c] + m[3][c];
				vtx[stream][c] = tmp0 + tmp1;
			}
		}
		errors[axis] = vector_distance3(vtx[0], vtx[1]);
	}
	return sse_max_ss(sse_max_ss(errors[0], errors[1]), errors[2]);
}

/* The sample loop of calculate_transform_track_error (track_error.impl.h:225-392) once both pose streams are sampled:
 * raw_poses = raw_tracks.sample_tracks(t_i), lossy_poses = seek(t_i) + decompress_tracks (already remapped, :341), both
 * [num_samples][num_tracks][12]; t_i = min(i / sample_rate, duration) (:337). base_poses (optional): the additive base sampled at the matching
 * times (:352-356), applied to both poses with `additive_format` (:358-359). out_errors (optional):
 * [num_samples][num_tracks]. Returns < 0 for an invalid skeleton order, 1 if a negative scale took rtm::qvv_mul through matrices (informational). */
int aclo_transform_track_error(const float* raw_poses, const float* lossy_poses, uint32_t num_samples, uint32_t num_tracks,
	float sample_rate, float duration, const uint32_t* parent_indices, const float* shell_distances, int normalize_mode,
	aclo_track_error* out_error, float* out_errors, float* scratch_object_poses /* [4][num_tracks][12] */,
	const float* base_poses /* [num_samples][num_tracks][12] or NULL */, uint32_t additive_format, uint32_t metric /* 0 qvvf, 1 qvvf_matrix3x4f */)
{
	out_error->index = 0xFFFFFFFFu;		/* track_error(), track_error.h:48-62 */
	out_error->error = 0.0f;
	out_error->sample_time = 0.0f;
	if (num_samples == 0 || num_tracks == 0)
		return 0;						/* :229-235 */
	out_error->error = -1.0f;			/* :333 */

	float* raw_object = scratch_object_poses;
	float* lossy_object = scratch_object_poses + (size_t)num_tracks * 12;
	int negative = 0;
	for (uint32_t sample = 0; sample < num_samples; ++sample)
	{
		const float t = (float)sample / sample_rate;
		const float sample_time = t < duration ? t : duration;		/* rtm::scalar_min */
		const size_t pose = (size_t)sample * num_tracks * 12;
		const float* raw_local = raw_poses + pose;
		const float* lossy_local = lossy_poses + pose;
		if (base_poses != NULL && additive_format != 0)
		{
			/* apply_additive_to_base on both poses before the walk (:358-359) */
			float* raw_applied = scratch_object_poses + (size_t)num_tracks * 24;
			float* lossy_applied = scratch_object_poses + (size_t)num_tracks * 36;
			memcpy(raw_applied, raw_local, (size_t)num_tracks * 12 * sizeof(float));
			memcpy(lossy_applied, lossy_local, (size_t)num_tracks * 12 * sizeof(float));
			negative |= aclo_apply_additive_to_base(additive_format, base_poses + pose, raw_applied, num_tracks, normalize_mode);
			negative |= aclo_apply_additive_to_base(additive_format, base_poses + pose, lossy_applied, num_tracks, normalize_mode);
			raw_local = raw_applied;
			lossy_local = lossy_applied;
		}
		const int r0 = metric == 1 ? matrix_local_to_object_space(raw_local, parent_indices, num_tracks, raw_object)
			: aclo_local_to_object_space(raw_local, parent_indices, num_tracks, normalize_mode, raw_object);
		const int r1 = metric == 1 ? matrix_local_to_object_space(lossy_local, parent_indices, num_tracks, lossy_object)
			: aclo_local_to_object_space(lossy_local, parent_indices, num_tracks, normalize_mode, lossy_object);
		if (r0 < 0 || r1 < 0)
			return -1;
		negative |= r0 | r1;
		for (uint32_t bone = 0; bone < num_tracks; ++bone)
		{
			const float error = metric == 1 ? matrix_calculate_error(raw_object + (size_t)bone * 12, lossy_object + (size_t)bone * 12, shell_distances[bone])
				: aclo_calculate_error(raw_object + (size_t)bone * 12, lossy_object + (size_t)bone * 12, shell_distances[bone]);
			if (out_errors != NULL)
				out_errors[(size_t)sample * num_tracks + bone] = error;
			if (error > out_error->error)		/* :367-372 */
			{
				out_error->error = error;
				out_error->index = bone;
				out_error->sample_time = sample_time;
			}
		}
	}
	return negative;
}

/* calculate_scalar_track_error, track_error.impl.h:166-223 with get_scalar_track_error :51-101: rows are [num_tracks][4] floats of which
 * the first `components` are compared (float1f..float4f / vector4f). */
int aclo_scalar_track_error(const float* raw_values, const float* lossy_values, uint32_t num_samples, uint32_t num_tracks, uint32_t components,
	float sample_rate, float duration, aclo_track_error* out_error)
{
	out_error->index = 0xFFFFFFFFu;
	out_error->error = 0.0f;
	out_error->sample_time = 0.0f;
	if (num_samples == 0 || num_tracks == 0)
		return 0;
	out_error->error = -1.0f;
	for (uint32_t sample = 0; sample < num_samples; ++sample)
	{
		const float t = (float)sample / sample_rate;
		const float sample_time = t < duration ? t : duration;
		for (uint32_t track = 0; track < num_tracks; ++track)
		{
			const size_t row = ((size_t)sample * num_tracks + track) * 4;
			/* abs(raw - lossy), unused lanes zeroed, vector_get_max_component = max(max(x, y), max(z, w)) (vector4f.h, SSE2 shuffles) */
			float e[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
			for (uint32_t c = 0; c < components; ++c)
				e[c] = fabsf(raw_values[row + c] - lossy_values[row + c]);
			const float xz = sse_max_ss(e[0], e[2]), yw = sse_max_ss(e[1], e[3]);
			const float max_error = components == 1 ? e[0] : sse_max_ss(xz, yw);
			if (max_error > out_error->error)
			{
				out_error->error = max_error;
				out_error->index = track;
				out_error->sample_time = sample_time;
			}
		}
	}
	return 0;
}

/* Test hooks for the known-answer table of the reference's own math tests (external/rtm/tests/sources/test_qvv.cpp:225-257):
 * rtm::qvv_mul (either branch) and rtm::qvv_mul_point3 as restated above. */
void aclo_test_qvv_mul(const float* lhs, const float* rhs, int normalize_mode, float* out)
{
	qvv_mul_plain(lhs, rhs, normalize_mode, out);
}

void aclo_test_qvv_mul_point3(const float* point, const float* qvv, float* out)
{
	qvv_mul_point3(point, qvv, out);
}
