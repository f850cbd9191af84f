// acl_b200/csrc/kernels.cu -- sm_100a kernels of the batched ACL decompression path.
//
// One launch decodes `num_requests` (clip, sample_time) requests == that many
//   context.seek(t, policy); context.decompress_tracks(writer);
// sequences of the reference (includes/acl/decompression/decompress.h:147-166). The kernel is fused: the seek
// (key frame / segment lookup, seek_v0, decompression/impl/decompression.transform.h:206-563), the variable bit
// rate unpack (unpack_animated_quat / unpack_animated_vector3, animated_track_cache.transform.h:515-687,871-990),
// the segment + clip range expansion (:157-350,391-466), the quaternion W reconstruction, the key frame
// interpolation and normalisation (math/quatf.h:135-211) all happen in one pass, and the pose is written once.
//
// Work decomposition (not the reference's: the CPU walks nine serial passes with running cursors):
//   thread block  = `requests_per_block` whole requests; their seek runs once, on one thread each, into shared memory
//   thread        = one bone of one request; the acceleration index built at upload (layout.h) gives it its
//                   constant / animated ranks and bit offsets, so no thread depends on another one.
//
// Arithmetic contract (EXACT mode): every float operation is an IEEE-754 round-to-nearest mul/add/sub/sqrt/rcp
// issued in the reference's order through __fmul_rn/__fadd_rn/... intrinsics, which nvcc never contracts into FMAs
// (the reference never fuses either: external/rtm/includes/rtm/impl/macros.vector4.impl.h:67,93,122). The results
// are bit-identical to the reference's SSE2/AVX/scalar builds for decompress_tracks.
#include "context.h"

namespace aclb200
{
	namespace
	{
		constexpr uint32_t k_threads_per_block = 256;
		constexpr uint32_t k_max_requests_per_block = 64;
		constexpr uint32_t k_target_poses_per_block = 2048;

		// ---------------------------------------------------------------------------------------------------
		// exact float helpers
		// ---------------------------------------------------------------------------------------------------
		__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
		__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
		__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
		// rtm::vector_mul_add(a, b, c) == (a * b) + c, two roundings
		__device__ __forceinline__ float fmuladd(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
		// rtm::vector_neg_mul_sub(a, b, c) == c - (a * b)
		__device__ __forceinline__ float fnegmulsub(float a, float b, float c) { return __fsub_rn(c, __fmul_rn(a, b)); }
		__device__ __forceinline__ float u2f(uint32_t v) { return __uint2float_rn(v); }

		// ---------------------------------------------------------------------------------------------------
		// per request state, written by one thread, read by every bone thread of the request
		// ---------------------------------------------------------------------------------------------------
		struct ReqState
		{
			const uint8_t* blob;
			const uint8_t* index;
			uint8_t* out;
			float    alpha;
			uint32_t num_tracks;			// 0 => nothing to decode (invalid request or empty clip)
			uint32_t clip_flags;
			uint32_t single_segment;
			uint32_t kf_bit[2];				// key_frame_bit_offsets
			uint32_t anim_off[2];
			uint32_t entries_off[2];
			uint32_t range_off[2][3];
			uint32_t const_off[3];
			uint32_t clip_range_off[3];
			uint32_t num_animated[3];
			uint32_t num_constant_rot;
			uint32_t bone_table_off;
			// extras reported by the debug seek kernel
			float    sample_time;
			uint32_t segment_index[2];
			uint32_t format_off[2];
			uint32_t looping_policy;
		};

		// apply_rounding_policy, core/impl/interpolation_utils.impl.h:261-278
		__device__ __forceinline__ float apply_rounding_policy(float alpha, uint32_t policy)
		{
			if (policy == ACLB200_ROUND_FLOOR) return 0.0f;
			if (policy == ACLB200_ROUND_CEIL) return 1.0f;
			if (policy == ACLB200_ROUND_NEAREST) return floorf(fadd(alpha, 0.5f));
			return alpha;
		}

		// find_linear_interpolation_alpha, core/impl/interpolation_utils.impl.h:224-253 with rounding_policy == none
		__device__ __forceinline__ float interpolation_alpha_none(float sample_index, uint32_t index0, uint32_t index1)
		{
			if (index0 == index1)
				return 0.0f;
			if (index0 < index1)
				return __fdiv_rn(fsub(sample_index, u2f(index0)), u2f(index1 - index0));
			return fsub(sample_index, u2f(index0));
		}

		// Looping policy + clamp duration: initialize_v0 / set_looping_policy_v0, decompression.transform.h:120-129,186-204
		__device__ __forceinline__ void resolve_looping(const DecodeParams& p, const ClipDesc& clip, uint32_t& policy, float& duration)
		{
			if (!p.wrapping)
				policy = ACLB200_LOOP_CLAMP;
			else if (p.looping_policy == ACLB200_LOOP_AS_COMPRESSED)
				policy = (clip.flags & k_clip_wrap) ? ACLB200_LOOP_WRAP : ACLB200_LOOP_CLAMP;
			else
				policy = p.looping_policy;
			duration = policy == ACLB200_LOOP_WRAP ? clip.duration_wrap : clip.duration_clamp;
		}

		// find_linear_interpolation_samples_with_sample_rate, core/impl/interpolation_utils.impl.h:143-201
		__device__ __forceinline__ void find_key_frames(uint32_t num_samples, float sample_rate, float sample_time, uint32_t rounding_policy,
			uint32_t looping_policy, uint32_t& key_frame0, uint32_t& key_frame1, float& alpha)
		{
			const uint32_t last_sample_index = num_samples - 1;
			float sample_index = fmul(sample_time, sample_rate);
			uint32_t sample_index0 = __float2uint_rz(sample_index);
			const uint32_t next_sample_index = sample_index0 + 1;
			uint32_t sample_index1;
			if (looping_policy == ACLB200_LOOP_CLAMP)
				sample_index1 = min(next_sample_index, last_sample_index);
			else if (sample_index0 > last_sample_index)
			{
				sample_index = 0.0f;
				sample_index0 = 0;
				sample_index1 = 0;
			}
			else
				sample_index1 = next_sample_index >= num_samples ? 0 : next_sample_index;
			key_frame0 = sample_index0;
			key_frame1 = sample_index1;
			alpha = apply_rounding_policy(fsub(sample_index, u2f(sample_index0)), rounding_policy);
		}

		// seek_v0 for transform clips, decompression.transform.h:206-563 (database branches do not exist here: such clips
		// are refused at upload)
		__device__ void seek_transform(const DecodeParams& p, uint32_t request_index, ReqState& rs)
		{
			rs.num_tracks = 0;
			rs.sample_time = -1.0f;
			const aclb200_request request = p.requests[request_index];
			if (request.clip >= p.num_clips)
				return;
			const ClipDesc& clip = p.clips[request.clip];
			if (clip.num_tracks == 0)
				return;

			const uint8_t* blob = p.blobs + clip.blob_offset;
			const uint8_t* index = p.index + clip.index_offset;

			uint32_t looping_policy;
			float duration;
			resolve_looping(p, clip, looping_policy, duration);

			float sample_time = request.sample_time;
			if (p.clamp_sample_time)
				sample_time = fminf(fmaxf(sample_time, 0.0f), duration);		// rtm::scalar_clamp, :215-216

			uint32_t key_frame0, key_frame1;
			float alpha;
			find_key_frames(clip.num_samples, clip.sample_rate, sample_time, p.rounding_policy, looping_policy, key_frame0, key_frame1, alpha);

			const SegDesc* segs = reinterpret_cast<const SegDesc*>(index + clip.seg_table_offset);
			const bool stripped = (clip.flags & k_clip_stripped) != 0;
			uint32_t segment_index0 = 0, segment_index1 = 0;
			uint32_t segment_key_frame0, segment_key_frame1;

			if (clip.num_segments == 1)
			{
				if (stripped)
				{
					// :272-362
					const uint32_t sample_indices = segs[0].sample_indices;
					const float sample_index = fadd(alpha, u2f(key_frame0));
					const uint32_t candidates0 = sample_indices & (0xFFFFFFFFu << (31 - key_frame0));
					key_frame0 = 31 - (__ffs(candidates0) - 1);						// count_trailing_zeros
					const uint32_t candidates1 = sample_indices & (0xFFFFFFFFu >> key_frame1);
					key_frame1 = __clz(candidates1);
					alpha = interpolation_alpha_none(sample_index, key_frame0, key_frame1);
					segment_key_frame0 = __popc(~(0xFFFFFFFFu >> key_frame0) & sample_indices);
					segment_key_frame1 = __popc(~(0xFFFFFFFFu >> key_frame1) & sample_indices);
				}
				else
				{
					segment_key_frame0 = key_frame0;
					segment_key_frame1 = key_frame1;
				}
			}
			else
			{
				// :372-409, segment_start_indices ends with a 0xFFFFFFFF sentinel (compression/impl/write_segment_data.h:48-65)
				const uint32_t* start_indices = reinterpret_cast<const uint32_t*>(blob + clip.start_indices_offset);
				const uint32_t approx_segment_index = key_frame0 / clip.samples_per_segment;
				const uint32_t start_segment_index = approx_segment_index > 0 ? approx_segment_index - 1 : 0;
				uint32_t found_start = 0;
				for (uint32_t i = 0; i < 4; ++i)
				{
					const uint32_t segment_index = start_segment_index + i;
					const uint32_t start = start_indices[segment_index];
					if (key_frame0 < start)
					{
						segment_index0 = segment_index - 1;
						if (p.wrapping && key_frame1 == 0)
							segment_index1 = 0;
						else
							segment_index1 = key_frame1 < start ? segment_index0 : segment_index;
						found_start = 1;
						break;
					}
				}
				(void)found_start;
				const uint32_t start0 = start_indices[segment_index0];
				const uint32_t start1 = start_indices[segment_index1];
				segment_key_frame0 = key_frame0 - start0;
				segment_key_frame1 = key_frame1 - start1;

				if (stripped)
				{
					// :411-515
					const uint32_t sample_indices0 = segs[segment_index0].sample_indices;
					const uint32_t sample_indices1 = segs[segment_index1].sample_indices;
					const float sample_index = fadd(alpha, u2f(key_frame0));
					const uint32_t candidates0 = sample_indices0 & (0xFFFFFFFFu << (31 - segment_key_frame0));
					segment_key_frame0 = 31 - (__ffs(candidates0) - 1);
					const uint32_t candidates1 = sample_indices1 & (0xFFFFFFFFu >> segment_key_frame1);
					segment_key_frame1 = __clz(candidates1);
					alpha = interpolation_alpha_none(sample_index, start0 + segment_key_frame0, start1 + segment_key_frame1);
					segment_key_frame0 = __popc(~(0xFFFFFFFFu >> segment_key_frame0) & sample_indices0);
					segment_key_frame1 = __popc(~(0xFFFFFFFFu >> segment_key_frame1) & sample_indices1);
				}
			}

			const SegDesc seg0 = segs[segment_index0];
			const SegDesc seg1 = segs[segment_index1];

			rs.blob = blob;
			rs.index = index;
			rs.alpha = alpha;
			rs.num_tracks = clip.num_tracks;
			rs.clip_flags = clip.flags;
			rs.single_segment = segment_index0 == segment_index1;
			rs.kf_bit[0] = segment_key_frame0 * seg0.pose_bit_size;				// :558-559
			rs.kf_bit[1] = segment_key_frame1 * seg1.pose_bit_size;
			rs.anim_off[0] = seg0.animated_offset;
			rs.anim_off[1] = seg1.animated_offset;
			rs.entries_off[0] = seg0.entries_offset;
			rs.entries_off[1] = seg1.entries_offset;
			for (int k = 0; k < 3; ++k)
			{
				rs.range_off[0][k] = seg0.range_offset[k];
				rs.range_off[1][k] = seg1.range_offset[k];
				rs.const_off[k] = clip.constant_offset[k];
				rs.clip_range_off[k] = clip.clip_range_offset[k];
				rs.num_animated[k] = clip.num_animated[k];
			}
			rs.num_constant_rot = clip.num_constant[0];
			rs.bone_table_off = clip.bone_table_offset;
			rs.sample_time = sample_time;
			rs.segment_index[0] = segment_index0;
			rs.segment_index[1] = segment_index1;
			rs.format_off[0] = seg0.format_offset;
			rs.format_off[1] = seg1.format_offset;
			rs.looping_policy = looping_policy;
		}

		// ---------------------------------------------------------------------------------------------------
		// bit stream reads. The animated stream of a segment starts on a 4 byte boundary of a 16 byte aligned
		// blob, so it can be read as big-endian 32 bit words.
		// ---------------------------------------------------------------------------------------------------
		__device__ __forceinline__ uint32_t load_be_word(const uint32_t* words, uint32_t word_index)
		{
			return __byte_perm(__ldg(words + word_index), 0, 0x0123);
		}

		// The 32 bits that start at bit `bit_offset` of the stream (unpack_vector3_96_unsafe, math/vector4_packing.h:482-503)
		__device__ __forceinline__ uint32_t read_bits32(const uint32_t* words, uint32_t bit_offset)
		{
			const uint32_t word_index = bit_offset >> 5;
			const uint32_t hi = load_be_word(words, word_index);
			const uint32_t lo = load_be_word(words, word_index + 1);
			return __funnelshift_l(lo, hi, bit_offset & 31);
		}

		// `num_bits` (1..23) bits at `bit_offset` (unpack_vector3_uXX_unsafe, math/vector4_packing.h:947-971)
		__device__ __forceinline__ uint32_t read_bits(const uint32_t* words, uint32_t bit_offset, uint32_t num_bits)
		{
			return read_bits32(words, bit_offset) >> (32 - num_bits);
		}

		// PackedTableEntry::max_value: 1.0F / float((1 << n) - 1) evaluated in float == correctly rounded reciprocal
		__device__ __forceinline__ float inv_max_value(uint32_t num_bits)
		{
			return __frcp_rn(u2f((1u << num_bits) - 1u));
		}

		// quat_from_positive_w4, math/quatf.h:135-147
		__device__ __forceinline__ float quat_w(float x, float y, float z)
		{
			float r = fnegmulsub(x, x, 1.0f);
			r = fnegmulsub(y, y, r);
			r = fnegmulsub(z, z, r);
			return __fsqrt_rn(fabsf(r));
		}

		// quat_normalize4, math/quatf.h:200-211
		__device__ __forceinline__ void quat_normalize(float q[4])
		{
			float dot = fmul(q[0], q[0]);
			dot = fmuladd(q[1], q[1], dot);
			dot = fmuladd(q[2], q[2], dot);
			dot = fmuladd(q[3], q[3], dot);
			const float len = __fsqrt_rn(dot);
			const float inv_len = __frcp_rn(len);		// vector_div(1.0, len): a correctly rounded reciprocal
			q[0] = fmul(q[0], inv_len);
			q[1] = fmul(q[1], inv_len);
			q[2] = fmul(q[2], inv_len);
			q[3] = fmul(q[3], inv_len);
		}

		// quat_lerp_no_normalization4, math/quatf.h:170-196
		__device__ __forceinline__ void quat_lerp(const float s[4], const float e[4], float alpha, float out[4])
		{
			float dot = fmul(s[0], e[0]);
			dot = fmuladd(s[1], e[1], dot);
			dot = fmuladd(s[2], e[2], dot);
			dot = fmuladd(s[3], e[3], dot);
			const uint32_t bias = __float_as_uint(dot) & 0x80000000u;
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				const float e_biased = __uint_as_float(__float_as_uint(e[i]) ^ bias);
				out[i] = fmuladd(e_biased, alpha, fnegmulsub(s[i], alpha, s[i]));
			}
		}

		// rtm::vector_lerp, external/rtm/includes/rtm/vector4f.h:2417-2421
		__device__ __forceinline__ float lerp(float start, float end, float alpha)
		{
			return fmuladd(end, alpha, fnegmulsub(start, alpha, start));
		}

		// rtm::quat_normalize (external/rtm/includes/rtm/quatf.h:917-953). On x86 this is rsqrtss + two Newton-Raphson steps,
		// whose result depends on the CPU's estimate table; we use the correctly rounded rsqrt, which every such estimate
		// converges to within 2 ulp (hence the 1e-5 gate on decompress_track rotations).
		__device__ __forceinline__ void rtm_quat_normalize(float q[4])
		{
			const float x2 = fmul(q[0], q[0]), y2 = fmul(q[1], q[1]), z2 = fmul(q[2], q[2]), w2 = fmul(q[3], q[3]);
			const float dot = fadd(fadd(x2, z2), fadd(y2, w2));
			const float inv_len = __frsqrt_rn(dot);
#pragma unroll
			for (int i = 0; i < 4; ++i)
				q[i] = fmul(q[i], inv_len);
		}

		// rtm::quat_lerp / acl::quat_lerp_no_normalization, SSE4 flavour (dpps sums (x+y)+(z+w)),
		// external/rtm/includes/rtm/quatf.h:1006-1075, math/quatf.h:40-82
		__device__ __forceinline__ void rtm_quat_lerp(const float s[4], const float e[4], float alpha, bool normalize, float out[4])
		{
			const float dot = fadd(fadd(fmul(s[0], e[0]), fmul(s[1], e[1])), fadd(fmul(s[2], e[2]), fmul(s[3], e[3])));
			const uint32_t bias = __float_as_uint(dot) & 0x80000000u;
#pragma unroll
			for (int i = 0; i < 4; ++i)
				out[i] = fadd(fsub(s[i], fmul(alpha, s[i])), fmul(alpha, __uint_as_float(__float_as_uint(e[i]) ^ bias)));
			if (normalize)
				rtm_quat_normalize(out);
		}

		// ---------------------------------------------------------------------------------------------------
		// sub-track decoders
		// ---------------------------------------------------------------------------------------------------

		// Raw integers of one animated sample: x, y, z (quantised integers or raw float bits), shared by the decode and by
		// the parity hook. Returns the entry code.
		__device__ __forceinline__ uint32_t unpack_sample_ints(const ReqState& rs, int k, uint32_t kind, uint32_t rank, uint32_t entry_index,
			uint32_t& xi, uint32_t& yi, uint32_t& zi, uint32_t& wi)
		{
			const uint32_t entry = __ldg(reinterpret_cast<const uint32_t*>(rs.index + rs.entries_off[k]) + entry_index);
			const uint32_t code = entry & 0xFFu;
			const uint32_t bit_offset = (entry >> 8) + rs.kf_bit[k];
			const uint32_t* words = reinterpret_cast<const uint32_t*>(rs.blob + rs.anim_off[k]);
			wi = 0;
			if (code == 0)
			{
				if (kind == 0)
				{
					// constant inside the segment: 16 bits per component spread over the SOA range bytes of the group
					// (unpack_animated_quat, animated_track_cache.transform.h:552-587)
					const uint8_t* r = rs.blob + rs.range_off[k][0] + (rank >> 2) * 24 + (rank & 3);
					xi = (uint32_t(__ldg(r + 0)) << 8) | __ldg(r + 4);
					yi = (uint32_t(__ldg(r + 8)) << 8) | __ldg(r + 12);
					zi = (uint32_t(__ldg(r + 16)) << 8) | __ldg(r + 20);
				}
				else
				{
					// unpack_vector3_u48_unsafe, math/vector4_packing.h:628-653: three native u16
					const uint16_t* r = reinterpret_cast<const uint16_t*>(rs.blob + rs.range_off[k][kind] + rank * 6);
					xi = __ldg(r + 0);
					yi = __ldg(r + 1);
					zi = __ldg(r + 2);
				}
			}
			else if (code & k_entry_raw)
			{
				xi = read_bits32(words, bit_offset);
				yi = read_bits32(words, bit_offset + 32);
				zi = read_bits32(words, bit_offset + 64);
				if (kind == 0 && (rs.clip_flags & k_clip_rot_full))
					wi = read_bits32(words, bit_offset + 96);
			}
			else
			{
				xi = read_bits(words, bit_offset, code);
				yi = read_bits(words, bit_offset + code, code);
				zi = read_bits(words, bit_offset + code * 2, code);
			}
			return code;
		}

		// One animated rotation sample after range expansion and W reconstruction.
		// SINGLE == false: decompress_tracks flavour (unpack_animated_quat + remap_segment_range_data4 + remap_clip_range_data4,
		//                  animated_track_cache.transform.h:515-687,302-350,391-466): ignored ranges still multiply by 1 and add 0.
		// SINGLE == true : decompress_track flavour (unpack_single_animated_quat, :689-869): ignored ranges are skipped.
		template<bool SINGLE>
		__device__ __forceinline__ void decode_animated_rotation(const ReqState& rs, int k, uint32_t rank, float out[4])
		{
			uint32_t xi, yi, zi, wi;
			const uint32_t code = unpack_sample_ints(rs, k, 0, rank, rank, xi, yi, zi, wi);

			if (!(rs.clip_flags & k_clip_rot_variable))
			{
				out[0] = __uint_as_float(xi);
				out[1] = __uint_as_float(yi);
				out[2] = __uint_as_float(zi);
				out[3] = (rs.clip_flags & k_clip_rot_full) ? __uint_as_float(wi) : quat_w(out[0], out[1], out[2]);
				return;
			}

			float x, y, z;
			bool ignore_segment = false, ignore_clip = false;
			if (code == 0)
			{
				const float scale = 1.0f / 65535.0f;
				x = fmul(u2f(xi), scale); y = fmul(u2f(yi), scale); z = fmul(u2f(zi), scale);
				ignore_segment = true;
			}
			else if (code & k_entry_raw)
			{
				x = __uint_as_float(xi); y = __uint_as_float(yi); z = __uint_as_float(zi);
				ignore_segment = true;
				ignore_clip = true;
			}
			else
			{
				const float inv_max = inv_max_value(code);
				x = fmul(u2f(xi), inv_max); y = fmul(u2f(yi), inv_max); z = fmul(u2f(zi), inv_max);
			}

			if ((rs.clip_flags & k_clip_has_segments) && (!SINGLE || !ignore_segment))
			{
				float min_x = 0.0f, min_y = 0.0f, min_z = 0.0f, ext_x = 1.0f, ext_y = 1.0f, ext_z = 1.0f;
				if (!ignore_segment)
				{
					// unpack_segment_range_data, :157-298: SOA bytes of the group of 4: min.xxxx min.yyyy min.zzzz extent.xxxx ...
					const uint8_t* r = rs.blob + rs.range_off[k][0] + (rank >> 2) * 24 + (rank & 3);
					const float n = 1.0f / 255.0f;
					min_x = fmul(u2f(__ldg(r + 0)), n); min_y = fmul(u2f(__ldg(r + 4)), n); min_z = fmul(u2f(__ldg(r + 8)), n);
					ext_x = fmul(u2f(__ldg(r + 12)), n); ext_y = fmul(u2f(__ldg(r + 16)), n); ext_z = fmul(u2f(__ldg(r + 20)), n);
				}
				x = fmuladd(x, ext_x, min_x);
				y = fmuladd(y, ext_y, min_y);
				z = fmuladd(z, ext_z, min_z);
			}

			if (!SINGLE || !ignore_clip)
			{
				float min_x = 0.0f, min_y = 0.0f, min_z = 0.0f, ext_x = 1.0f, ext_y = 1.0f, ext_z = 1.0f;
				if (!ignore_clip)
				{
					// remap_clip_range_data4, :391-466: SOA per group of 4, the last group holds `group_size` lanes
					const uint32_t group = rank >> 2;
					const uint32_t group_size = min(rs.num_animated[0] - group * 4, 4u);
					const float* r = reinterpret_cast<const float*>(rs.blob + rs.clip_range_off[0] + group * 96) + (rank & 3);
					min_x = __ldg(r + group_size * 0); min_y = __ldg(r + group_size * 1); min_z = __ldg(r + group_size * 2);
					ext_x = __ldg(r + group_size * 3); ext_y = __ldg(r + group_size * 4); ext_z = __ldg(r + group_size * 5);
				}
				x = fmuladd(x, ext_x, min_x);
				y = fmuladd(y, ext_y, min_y);
				z = fmuladd(z, ext_z, min_z);
			}

			out[0] = x; out[1] = y; out[2] = z;
			out[3] = quat_w(x, y, z);
		}

		// unpack_animated_vector3 / unpack_single_animated_vector3, animated_track_cache.transform.h:871-990,992-1102
		__device__ __forceinline__ void decode_animated_vector3(const ReqState& rs, int k, uint32_t kind, uint32_t rank, float out[3])
		{
			const uint32_t entry_index = rs.num_animated[0] + (kind == 2 ? rs.num_animated[1] : 0u) + rank;
			uint32_t xi, yi, zi, wi;
			const uint32_t code = unpack_sample_ints(rs, k, kind, rank, entry_index, xi, yi, zi, wi);
			const bool variable = (rs.clip_flags & (kind == 1 ? k_clip_trans_variable : k_clip_scale_variable)) != 0;

			if (!variable || (code & k_entry_raw))
			{
				out[0] = __uint_as_float(xi); out[1] = __uint_as_float(yi); out[2] = __uint_as_float(zi);
				return;
			}

			float x, y, z;
			if (code == 0)
			{
				const float scale = 1.0f / 65535.0f;
				x = fmul(u2f(xi), scale); y = fmul(u2f(yi), scale); z = fmul(u2f(zi), scale);
			}
			else
			{
				const float inv_max = inv_max_value(code);
				x = fmul(u2f(xi), inv_max); y = fmul(u2f(yi), inv_max); z = fmul(u2f(zi), inv_max);
				if (rs.clip_flags & k_clip_has_segments)
				{
					// unpack_vector3_u24_unsafe min then extent, math/vector4_packing.h:781-818
					const uint8_t* r = rs.blob + rs.range_off[k][kind] + rank * 6;
					const float n = 1.0f / 255.0f;
					x = fmuladd(x, fmul(u2f(__ldg(r + 3)), n), fmul(u2f(__ldg(r + 0)), n));
					y = fmuladd(y, fmul(u2f(__ldg(r + 4)), n), fmul(u2f(__ldg(r + 1)), n));
					z = fmuladd(z, fmul(u2f(__ldg(r + 5)), n), fmul(u2f(__ldg(r + 2)), n));
				}
			}

			// clip range: min xyz then extent xyz, 24 bytes per sub-track (:949-958)
			const float* r = reinterpret_cast<const float*>(rs.blob + rs.clip_range_off[kind] + rank * 24);
			out[0] = fmuladd(x, __ldg(r + 3), __ldg(r + 0));
			out[1] = fmuladd(y, __ldg(r + 4), __ldg(r + 1));
			out[2] = fmuladd(z, __ldg(r + 5), __ldg(r + 2));
		}

		// constant_track_cache_v0::unpack_rotation_group / unpack_rotation_within_group, constant_track_cache.transform.h:112-205,232-264
		template<int NORM, bool SINGLE>
		__device__ __forceinline__ void decode_constant_rotation(const ReqState& rs, uint32_t rank, float out[4])
		{
			if (rs.clip_flags & k_clip_rot_full)
			{
				const float* r = reinterpret_cast<const float*>(rs.blob + rs.const_off[0]) + rank * 4;
				out[0] = __ldg(r + 0); out[1] = __ldg(r + 1); out[2] = __ldg(r + 2); out[3] = __ldg(r + 3);
				return;
			}
			const uint32_t group = rank >> 2;
			const uint32_t group_size = min(rs.num_constant_rot - group * 4, 4u);
			const float* r = reinterpret_cast<const float*>(rs.blob + rs.const_off[0] + group * 48) + (rank & 3);
			const float x = __ldg(r + group_size * 0);
			const float y = __ldg(r + group_size * 1);
			const float z = __ldg(r + group_size * 2);
			out[0] = x; out[1] = y; out[2] = z;
			out[3] = quat_w(x, y, z);
			if (NORM == ACLB200_NORMALIZE_ALWAYS)
			{
				if (SINGLE)
					rtm_quat_normalize(out);
				else
					quat_normalize(out);
			}
		}

		// should_interpolate_samples, decompression_context.transform.h:191-200
		__device__ __forceinline__ bool should_interpolate(const DecodeParams& p, uint32_t clip_flags, float alpha)
		{
			if (p.multiple_rotation_formats)
				return true;
			return (clip_flags & k_clip_rot_full) ? (alpha > 0.0f && alpha < 1.0f) : true;
		}

		__device__ __forceinline__ uint32_t track_rounding_policy(const DecodeParams& p, uint32_t track)
		{
			// track_writer::get_rounding_policy(seek_policy, track_index), core/track_writer.h:90
			if (p.rounding_policy != ACLB200_ROUND_PER_TRACK || p.per_track_policies == nullptr)
				return p.rounding_policy;
			return p.per_track_policies[track];
		}

		// Default sub-tracks: unpack_default_*_sub_tracks, decompression.transform.h:574-675,881-983,1201-1310
		__device__ __forceinline__ bool default_value(const DecodeParams& p, uint32_t kind, uint32_t track, uint32_t clip_flags, float out[4])
		{
			const uint32_t mode = p.default_mode[kind];
			if (mode == ACLB200_DEFAULT_SKIPPED)
				return false;
			if (mode == ACLB200_DEFAULT_VARIABLE && p.variable_defaults != nullptr)
			{
				const float* v = p.variable_defaults + size_t(track) * 12 + kind * 4;
				out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; out[3] = v[3];
			}
			else if (mode == ACLB200_DEFAULT_LEGACY && kind == 2)
			{
				const float s = (clip_flags & k_clip_default_scale_one) ? 1.0f : 0.0f;	// float(header.get_default_scale()), :1548
				out[0] = s; out[1] = s; out[2] = s; out[3] = s;
			}
			else
			{
				out[0] = p.constant_defaults[kind * 4 + 0]; out[1] = p.constant_defaults[kind * 4 + 1];
				out[2] = p.constant_defaults[kind * 4 + 2]; out[3] = p.constant_defaults[kind * 4 + 3];
			}
			return true;
		}

		// One bone of one request: the three sub-tracks of decompress_tracks_v0 (decompression.transform.h:1526-1737) or
		// decompress_track_v0 (:1753-2050, SINGLE).
		template<int NORM, bool PER_TRACK, bool SINGLE>
		__device__ __forceinline__ void decode_bone(const DecodeParams& p, const ReqState& rs, uint32_t bone, uint8_t* out_bone)
		{
			const uint64_t desc = __ldg(reinterpret_cast<const unsigned long long*>(rs.index + rs.bone_table_off) + bone);
			const uint32_t policy = PER_TRACK ? track_rounding_policy(p, bone) : ACLB200_ROUND_NONE;
			const float alpha = (SINGLE && PER_TRACK) ? apply_rounding_policy(rs.alpha, policy) : rs.alpha;	// :1975-1983

			float rotation[4];
			float translation[4];
			float scale[4];
			bool write_rotation = true, write_translation = true, write_scale = true;

			// ---- rotation ----
			{
				const uint32_t type = uint32_t(desc) & 3;
				const uint32_t rank = (uint32_t(desc) >> 2) & k_bone_index_mask;
				if (type == 0)
					write_rotation = default_value(p, 0, bone, rs.clip_flags, rotation);
				else if (type == 1)
					decode_constant_rotation<NORM, SINGLE>(rs, rank, rotation);
				else
				{
					float s0[4], s1[4];
					decode_animated_rotation<SINGLE>(rs, 0, rank, s0);
					decode_animated_rotation<SINGLE>(rs, 1, rank, s1);
					const bool interpolate = should_interpolate(p, rs.clip_flags, alpha);

					if (SINGLE)
					{
						// unpack_rotation_within_group, animated_track_cache.transform.h:1709-1765
						if (interpolate)
							rtm_quat_lerp(s0, s1, alpha, NORM >= ACLB200_NORMALIZE_LERP_ONLY, rotation);
						else
						{
#pragma unroll
							for (int i = 0; i < 4; ++i)
								rotation[i] = alpha <= 0.0f ? s0[i] : s1[i];
							if (NORM == ACLB200_NORMALIZE_ALWAYS && !(rs.clip_flags & k_clip_rot_full))
								rtm_quat_normalize(rotation);
						}
					}
					else
					{
						// unpack_rotation_group, :1463-1474,1477-1661
						if (NORM == ACLB200_NORMALIZE_ALWAYS && !(rs.clip_flags & k_clip_rot_full) && (PER_TRACK || !interpolate))
						{
							quat_normalize(s0);
							quat_normalize(s1);
						}

						if (PER_TRACK && policy == ACLB200_ROUND_FLOOR)
						{
#pragma unroll
							for (int i = 0; i < 4; ++i) rotation[i] = s0[i];
						}
						else if (PER_TRACK && policy == ACLB200_ROUND_CEIL)
						{
#pragma unroll
							for (int i = 0; i < 4; ++i) rotation[i] = s1[i];
						}
						else if (PER_TRACK && policy == ACLB200_ROUND_NEAREST)
						{
#pragma unroll
							for (int i = 0; i < 4; ++i) rotation[i] = alpha < 0.5f ? s0[i] : s1[i];
						}
						else if (PER_TRACK || interpolate)
						{
							quat_lerp(s0, s1, alpha, rotation);
							if (NORM >= ACLB200_NORMALIZE_LERP_ONLY)
								quat_normalize(rotation);
						}
						else
						{
#pragma unroll
							for (int i = 0; i < 4; ++i) rotation[i] = alpha <= 0.0f ? s0[i] : s1[i];
						}
					}
				}
			}

			// ---- translation, scale ----
#pragma unroll
			for (uint32_t kind = 1; kind <= 2; ++kind)
			{
				float* value = kind == 1 ? translation : scale;
				bool written = true;
				const uint32_t bits = uint32_t(desc >> (k_bone_kind_shift * kind));
				// clips without scale: every bone takes the default (decompression.transform.h:1653-1680,1806-1822)
				const uint32_t type = (kind == 2 && !(rs.clip_flags & k_clip_has_scale)) ? 0u : (bits & 3);
				const uint32_t rank = (bits >> 2) & k_bone_index_mask;
				if (type == 0)
					written = default_value(p, kind, bone, rs.clip_flags, value);
				else if (type == 1)
				{
					const float* r = reinterpret_cast<const float*>(rs.blob + rs.const_off[kind]) + rank * 3;
					value[0] = __ldg(r + 0); value[1] = __ldg(r + 1); value[2] = __ldg(r + 2);
					value[3] = 0.0f;
				}
				else
				{
					float s0[3], s1[3];
					decode_animated_vector3(rs, 0, kind, rank, s0);
					decode_animated_vector3(rs, 1, kind, rank, s1);
#pragma unroll
					for (int i = 0; i < 3; ++i)
					{
						// unpack_translation_group / consume_translation, animated_track_cache.transform.h:1774-1836,1889-1894
						if (!SINGLE && PER_TRACK && policy == ACLB200_ROUND_FLOOR)
							value[i] = s0[i];
						else if (!SINGLE && PER_TRACK && policy == ACLB200_ROUND_CEIL)
							value[i] = s1[i];
						else if (!SINGLE && PER_TRACK && policy == ACLB200_ROUND_NEAREST)
							value[i] = alpha < 0.5f ? s0[i] : s1[i];
						else
							value[i] = lerp(s0[i], s1[i], alpha);
					}
					value[3] = 0.0f;
				}
				if (kind == 1) write_translation = written; else write_scale = written;
			}

			// ---- the device track_writer: write_rotation / write_translation / write_scale ----
			if (p.layout == ACLB200_LAYOUT_QVV48)
			{
				float4* dst = reinterpret_cast<float4*>(out_bone);
				if (write_rotation) dst[0] = make_float4(rotation[0], rotation[1], rotation[2], rotation[3]);
				if (write_translation) dst[1] = make_float4(translation[0], translation[1], translation[2], 0.0f);
				if (write_scale) dst[2] = make_float4(scale[0], scale[1], scale[2], 0.0f);
			}
			else
			{
				float2* dst = reinterpret_cast<float2*>(out_bone);		// 40 byte bones are 8 byte aligned
				if (write_rotation && write_translation && write_scale)
				{
					dst[0] = make_float2(rotation[0], rotation[1]);
					dst[1] = make_float2(rotation[2], rotation[3]);
					dst[2] = make_float2(translation[0], translation[1]);
					dst[3] = make_float2(translation[2], scale[0]);
					dst[4] = make_float2(scale[1], scale[2]);
				}
				else
				{
					float* f = reinterpret_cast<float*>(out_bone);
					if (write_rotation) { f[0] = rotation[0]; f[1] = rotation[1]; f[2] = rotation[2]; f[3] = rotation[3]; }
					if (write_translation) { f[4] = translation[0]; f[5] = translation[1]; f[6] = translation[2]; }
					if (write_scale) { f[7] = scale[0]; f[8] = scale[1]; f[9] = scale[2]; }
				}
			}
		}

		// ---------------------------------------------------------------------------------------------------
		// kernels
		// ---------------------------------------------------------------------------------------------------
		template<int NORM, bool PER_TRACK>
		__global__ void __launch_bounds__(k_threads_per_block)
		transform_decompress_tracks_kernel(const DecodeParams p)
		{
			__shared__ ReqState s_req[k_max_requests_per_block];

			const uint32_t first_request = blockIdx.x * p.requests_per_block;
			const uint32_t num_requests = min(p.requests_per_block, p.num_requests - first_request);

			if (threadIdx.x < num_requests)
			{
				ReqState rs;
				seek_transform(p, first_request + threadIdx.x, rs);
				rs.out = p.out + uint64_t(first_request + threadIdx.x) * p.pose_stride;
				s_req[threadIdx.x] = rs;
			}
			__syncthreads();

			const uint32_t num_slots = num_requests * p.max_tracks;
			for (uint32_t slot = threadIdx.x; slot < num_slots; slot += k_threads_per_block)
			{
				const uint32_t local_request = p.max_tracks_magic != 0 ? __umulhi(slot, p.max_tracks_magic) : slot;
				const uint32_t bone = slot - local_request * p.max_tracks;
				const ReqState& rs = s_req[local_request];
				if (bone >= rs.num_tracks)
					continue;
				decode_bone<NORM, PER_TRACK, false>(p, rs, bone, rs.out + size_t(bone) * p.bone_stride);
			}
		}

		template<int NORM, bool PER_TRACK>
		__global__ void __launch_bounds__(128)
		transform_decompress_track_kernel(const DecodeParams p)
		{
			const uint32_t request = blockIdx.x * blockDim.x + threadIdx.x;
			if (request >= p.num_requests)
				return;
			ReqState rs;
			seek_transform(p, request, rs);
			const uint32_t bone = p.track_indices[request];
			if (bone >= rs.num_tracks)
				return;		// decompress_track_v0 :1766-1768: invalid track index, nothing is written
			decode_bone<NORM, PER_TRACK, true>(p, rs, bone, p.out + uint64_t(request) * p.bone_stride);
		}

		__global__ void __launch_bounds__(128)
		transform_debug_seek_kernel(const DecodeParams p, aclb200_seek_state* __restrict__ out)
		{
			const uint32_t request = blockIdx.x * blockDim.x + threadIdx.x;
			if (request >= p.num_requests)
				return;
			ReqState rs;
			seek_transform(p, request, rs);
			aclb200_seek_state st = {};
			st.sample_time = rs.sample_time;
			if (rs.num_tracks != 0)
			{
				st.interpolation_alpha = rs.alpha;
				st.uses_single_segment = rs.single_segment;
				st.looping_policy = rs.looping_policy;
				for (int k = 0; k < 2; ++k)
				{
					st.key_frame_bit_offsets[k] = rs.kf_bit[k];
					st.segment_indices[k] = rs.segment_index[k];
					st.animated_offsets[k] = rs.anim_off[k];
					st.format_offsets[k] = rs.format_off[k];
					st.range_offsets[k] = rs.range_off[k][0];
				}
			}
			out[request] = st;
		}

		__global__ void __launch_bounds__(128)
		transform_debug_unpack_kernel(const DecodeParams p, uint32_t* __restrict__ out)
		{
			__shared__ ReqState s_req;
			const uint32_t request = blockIdx.x;
			if (threadIdx.x == 0)
				seek_transform(p, request, s_req);
			__syncthreads();
			const ReqState& rs = s_req;
			if (rs.num_tracks == 0)
				return;
			const int k = int(p.debug_which);
			const uint32_t total = rs.num_animated[0] + rs.num_animated[1] + rs.num_animated[2];
			for (uint32_t j = threadIdx.x; j < total && j < p.debug_max_sub_tracks; j += blockDim.x)
			{
				uint32_t kind = 0, rank = j;
				if (rank >= rs.num_animated[0]) { rank -= rs.num_animated[0]; kind = 1; }
				if (kind == 1 && rank >= rs.num_animated[1]) { rank -= rs.num_animated[1]; kind = 2; }
				uint32_t xi, yi, zi, wi;
				const uint32_t code = unpack_sample_ints(rs, k, kind, rank, j, xi, yi, zi, wi);
				uint32_t* dst = out + (size_t(request) * p.debug_max_sub_tracks + j) * 4;
				dst[0] = xi; dst[1] = yi; dst[2] = zi;
				dst[3] = code;
			}
		}

		// ---------------------------------------------------------------------------------------------------
		// scalar tracks: seek_v0 + decompress_tracks_v0 / decompress_track_v0, decompression/impl/decompression.scalar.h:181-705
		// ---------------------------------------------------------------------------------------------------
		struct ScalarReqState
		{
			const uint8_t* blob;
			const ScalarTrackDesc* tracks;
			uint8_t* out;
			float    alpha;
			uint32_t num_tracks;
			uint32_t kf_bit[2];
			uint32_t constant_off;
			uint32_t range_off;
			uint32_t animated_off;
		};

		__device__ void seek_scalar(const DecodeParams& p, uint32_t request_index, ScalarReqState& rs)
		{
			rs.num_tracks = 0;
			const aclb200_request request = p.requests[request_index];
			if (request.clip >= p.num_clips)
				return;
			const ClipDesc& clip = p.clips[request.clip];
			if (clip.num_tracks == 0 || clip.num_samples == 0)
				return;

			uint32_t looping_policy;
			float duration;
			resolve_looping(p, clip, looping_policy, duration);
			float sample_time = request.sample_time;
			if (p.clamp_sample_time)
				sample_time = fminf(fmaxf(sample_time, 0.0f), duration);

			uint32_t key_frame0, key_frame1;
			float alpha;
			find_key_frames(clip.num_samples, clip.sample_rate, sample_time, p.rounding_policy, looping_policy, key_frame0, key_frame1, alpha);

			rs.blob = p.blobs + clip.blob_offset;
			rs.tracks = reinterpret_cast<const ScalarTrackDesc*>(p.index + clip.index_offset + clip.bone_table_offset);
			rs.alpha = alpha;
			rs.num_tracks = clip.num_tracks;
			rs.kf_bit[0] = key_frame0 * clip.num_constant[0];		// num_bits_per_frame, decompression.scalar.h:208-209
			rs.kf_bit[1] = key_frame1 * clip.num_constant[0];
			rs.constant_off = clip.constant_offset[0];
			rs.range_off = clip.constant_offset[1];
			rs.animated_off = clip.constant_offset[2];
		}

		// The animated values of scalar clips start at an arbitrary byte: read through the enclosing aligned words.
		__device__ __forceinline__ uint32_t read_bits32_unaligned(const uint8_t* base, uint32_t bit_offset)
		{
			const uintptr_t address = reinterpret_cast<uintptr_t>(base) + (bit_offset >> 3);
			const uint32_t* words = reinterpret_cast<const uint32_t*>(address & ~uintptr_t(3));
			const uint32_t shift = uint32_t(address & 3) * 8 + (bit_offset & 7);		// 0..31
			const uint32_t w0 = __byte_perm(__ldg(words + 0), 0, 0x0123);
			const uint32_t w1 = __byte_perm(__ldg(words + 1), 0, 0x0123);
			return __funnelshift_l(w1, w0, shift);
		}

		template<int COMPONENTS, bool PER_TRACK>
		__device__ __forceinline__ void decode_scalar_track(const DecodeParams& p, const ScalarReqState& rs, uint32_t track, float* out)
		{
			const ScalarTrackDesc desc = rs.tracks[track];
			const uint32_t num_bits = desc.value_index_and_bits & 0xFFu;
			const uint32_t value_index = desc.value_index_and_bits >> 8;
			float alpha = rs.alpha;
			if (PER_TRACK)
				alpha = apply_rounding_policy(rs.alpha, track_rounding_policy(p, track));	// decompression.scalar.h:235-247,273-280

			if (num_bits == 0)
			{
				const float* constants = reinterpret_cast<const float*>(rs.blob + rs.constant_off) + value_index;
#pragma unroll
				for (int c = 0; c < COMPONENTS; ++c)
					out[c] = __ldg(constants + c);
				return;
			}

			const uint8_t* animated = rs.blob + rs.animated_off;
			const uint32_t bit0 = rs.kf_bit[0] + desc.bit_offset;
			const uint32_t bit1 = rs.kf_bit[1] + desc.bit_offset;
			if (num_bits == 32)
			{
#pragma unroll
				for (int c = 0; c < COMPONENTS; ++c)
				{
					const float v0 = __uint_as_float(read_bits32_unaligned(animated, bit0 + 32 * c));
					const float v1 = __uint_as_float(read_bits32_unaligned(animated, bit1 + 32 * c));
					out[c] = lerp(v0, v1, alpha);
				}
				return;
			}

			const float inv_max = inv_max_value(num_bits);
			const float* range = reinterpret_cast<const float*>(rs.blob + rs.range_off) + value_index;
#pragma unroll
			for (int c = 0; c < COMPONENTS; ++c)
			{
				const uint32_t i0 = read_bits32_unaligned(animated, bit0 + num_bits * c) >> (32 - num_bits);
				const uint32_t i1 = read_bits32_unaligned(animated, bit1 + num_bits * c) >> (32 - num_bits);
				const float range_min = __ldg(range + c);
				const float range_extent = __ldg(range + COMPONENTS + c);
				const float v0 = fmuladd(fmul(u2f(i0), inv_max), range_extent, range_min);
				const float v1 = fmuladd(fmul(u2f(i1), inv_max), range_extent, range_min);
				out[c] = lerp(v0, v1, alpha);
			}
		}

		template<int COMPONENTS, bool PER_TRACK>
		__global__ void __launch_bounds__(k_threads_per_block)
		scalar_decompress_tracks_kernel(const DecodeParams p)
		{
			__shared__ ScalarReqState s_req[k_max_requests_per_block];

			const uint32_t first_request = blockIdx.x * p.requests_per_block;
			const uint32_t num_requests = min(p.requests_per_block, p.num_requests - first_request);
			if (threadIdx.x < num_requests)
			{
				ScalarReqState rs;
				seek_scalar(p, first_request + threadIdx.x, rs);
				rs.out = p.out + uint64_t(first_request + threadIdx.x) * p.pose_stride;
				s_req[threadIdx.x] = rs;
			}
			__syncthreads();

			const uint32_t num_slots = num_requests * p.max_tracks;
			for (uint32_t slot = threadIdx.x; slot < num_slots; slot += k_threads_per_block)
			{
				const uint32_t local_request = p.max_tracks_magic != 0 ? __umulhi(slot, p.max_tracks_magic) : slot;
				const uint32_t track = slot - local_request * p.max_tracks;
				const ScalarReqState& rs = s_req[local_request];
				if (track >= rs.num_tracks)
					continue;
				float value[COMPONENTS];
				decode_scalar_track<COMPONENTS, PER_TRACK>(p, rs, track, value);
				float* dst = reinterpret_cast<float*>(rs.out) + size_t(track) * COMPONENTS;
#pragma unroll
				for (int c = 0; c < COMPONENTS; ++c)
					dst[c] = value[c];
			}
		}

		template<int COMPONENTS, bool PER_TRACK>
		__global__ void __launch_bounds__(128)
		scalar_decompress_track_kernel(const DecodeParams p)
		{
			const uint32_t request = blockIdx.x * blockDim.x + threadIdx.x;
			if (request >= p.num_requests)
				return;
			ScalarReqState rs;
			seek_scalar(p, request, rs);
			const uint32_t track = p.track_indices[request];
			if (track >= rs.num_tracks)
				return;
			float value[COMPONENTS];
			decode_scalar_track<COMPONENTS, PER_TRACK>(p, rs, track, value);
			float* dst = reinterpret_cast<float*>(p.out) + size_t(request) * COMPONENTS;
#pragma unroll
			for (int c = 0; c < COMPONENTS; ++c)
				dst[c] = value[c];
		}

		template<template<int, bool> class Launcher>
		cudaError_t dispatch_transform(const DecodeParams& params, cudaStream_t stream)
		{
			const bool per_track = params.per_track_rounding != 0;
			switch (params.normalization)
			{
			case ACLB200_NORMALIZE_NEVER: return per_track ? Launcher<0, true>::launch(params, stream) : Launcher<0, false>::launch(params, stream);
			case ACLB200_NORMALIZE_LERP_ONLY: return per_track ? Launcher<1, true>::launch(params, stream) : Launcher<1, false>::launch(params, stream);
			default: return per_track ? Launcher<2, true>::launch(params, stream) : Launcher<2, false>::launch(params, stream);
			}
		}

		template<int NORM, bool PER_TRACK>
		struct TracksLauncher
		{
			static cudaError_t launch(const DecodeParams& params, cudaStream_t stream)
			{
				const uint32_t blocks = (params.num_requests + params.requests_per_block - 1) / params.requests_per_block;
				transform_decompress_tracks_kernel<NORM, PER_TRACK><<<blocks, k_threads_per_block, 0, stream>>>(params);
				return cudaGetLastError();
			}
		};

		template<int NORM, bool PER_TRACK>
		struct TrackLauncher
		{
			static cudaError_t launch(const DecodeParams& params, cudaStream_t stream)
			{
				const uint32_t blocks = (params.num_requests + 127) / 128;
				transform_decompress_track_kernel<NORM, PER_TRACK><<<blocks, 128, 0, stream>>>(params);
				return cudaGetLastError();
			}
		};
	}

	// requests_per_block and the division magic for a launch over `max_tracks` wide poses
	void plan_launch(DecodeParams& params)
	{
		const uint32_t max_tracks = params.max_tracks == 0 ? 1 : params.max_tracks;
		uint32_t requests_per_block = k_target_poses_per_block / max_tracks;
		if (requests_per_block < 1) requests_per_block = 1;
		if (requests_per_block > k_max_requests_per_block) requests_per_block = k_max_requests_per_block;
		params.requests_per_block = requests_per_block;
		// floor(slot / max_tracks) == mulhi(slot, magic) for slot < 2^16 * ... (slot < requests_per_block * max_tracks <= 2^18 + 2048)
		params.max_tracks_magic = max_tracks == 1 ? 0u : uint32_t((uint64_t(1) << 32) / max_tracks) + 1u;
	}

	cudaError_t launch_transform_decompress_tracks(const DecodeParams& params, uint32_t /*math_mode*/, cudaStream_t stream)
	{
		return dispatch_transform<TracksLauncher>(params, stream);
	}

	cudaError_t launch_transform_decompress_track(const DecodeParams& params, uint32_t /*math_mode*/, cudaStream_t stream)
	{
		return dispatch_transform<TrackLauncher>(params, stream);
	}

	cudaError_t launch_transform_debug_seek(const DecodeParams& params, aclb200_seek_state* d_out, cudaStream_t stream)
	{
		const uint32_t blocks = (params.num_requests + 127) / 128;
		transform_debug_seek_kernel<<<blocks, 128, 0, stream>>>(params, d_out);
		return cudaGetLastError();
	}

	cudaError_t launch_transform_debug_unpack(const DecodeParams& params, uint32_t* d_out, cudaStream_t stream)
	{
		transform_debug_unpack_kernel<<<params.num_requests, 128, 0, stream>>>(params, d_out);
		return cudaGetLastError();
	}

	template<bool PER_TRACK>
	static cudaError_t launch_scalar_tracks(const DecodeParams& params, uint32_t components, cudaStream_t stream)
	{
		const uint32_t blocks = (params.num_requests + params.requests_per_block - 1) / params.requests_per_block;
		switch (components)
		{
		case 1: scalar_decompress_tracks_kernel<1, PER_TRACK><<<blocks, k_threads_per_block, 0, stream>>>(params); break;
		case 2: scalar_decompress_tracks_kernel<2, PER_TRACK><<<blocks, k_threads_per_block, 0, stream>>>(params); break;
		case 3: scalar_decompress_tracks_kernel<3, PER_TRACK><<<blocks, k_threads_per_block, 0, stream>>>(params); break;
		default: scalar_decompress_tracks_kernel<4, PER_TRACK><<<blocks, k_threads_per_block, 0, stream>>>(params); break;
		}
		return cudaGetLastError();
	}

	template<bool PER_TRACK>
	static cudaError_t launch_scalar_track(const DecodeParams& params, uint32_t components, cudaStream_t stream)
	{
		const uint32_t blocks = (params.num_requests + 127) / 128;
		switch (components)
		{
		case 1: scalar_decompress_track_kernel<1, PER_TRACK><<<blocks, 128, 0, stream>>>(params); break;
		case 2: scalar_decompress_track_kernel<2, PER_TRACK><<<blocks, 128, 0, stream>>>(params); break;
		case 3: scalar_decompress_track_kernel<3, PER_TRACK><<<blocks, 128, 0, stream>>>(params); break;
		default: scalar_decompress_track_kernel<4, PER_TRACK><<<blocks, 128, 0, stream>>>(params); break;
		}
		return cudaGetLastError();
	}

	cudaError_t launch_scalar_decompress_tracks(const DecodeParams& params, cudaStream_t stream)
	{
		const uint32_t components = params.bone_stride / 4;
		return params.per_track_rounding ? launch_scalar_tracks<true>(params, components, stream) : launch_scalar_tracks<false>(params, components, stream);
	}

	cudaError_t launch_scalar_decompress_track(const DecodeParams& params, cudaStream_t stream)
	{
		const uint32_t components = params.bone_stride / 4;
		return params.per_track_rounding ? launch_scalar_track<true>(params, components, stream) : launch_scalar_track<false>(params, components, stream);
	}
}
