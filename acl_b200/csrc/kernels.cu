// acl_b200/csrc/kernels.cu -- sm_100a kernels of the batched ACL decompression path.
//
// One launch decodes `num_requests` (clip, sample_time) requests == that many
//   context.seek(t, policy); context.decompress_tracks(writer);
// sequences of the reference (includes/acl/decompression/decompress.h:147-166). The kernel is fused: the seek
// (key frame / segment lookup, seek_v0, decompression/impl/decompression.transform.h:206-563), the variable bit
// rate unpack (unpack_animated_quat / unpack_animated_vector3, animated_track_cache.transform.h:515-687,871-990),
// the segment + clip range expansion (:157-350,391-466), the quaternion W reconstruction, the key frame
// interpolation and normalisation (math/quatf.h:135-211) all happen in one pass, and every pose byte is written once.
//
// Work decomposition (not the reference's: the CPU walks nine serial passes with running cursors):
//   thread block = `requests_per_block` whole requests.
//     phase 1  one thread per request runs the seek, stores the request state in shared memory and asks the TMA unit
//              (cp.async.bulk + mbarrier) to stage the request's two key frames -- a few hundred contiguous bytes of the
//              packed segment stream each -- in shared memory.
//     phase 2  while those copies fly: one thread per (request, bone) writes the constant and default sub-tracks.
//     phase 3  one thread per (request, animated rotation sub-track): unpack both key frames from shared memory, expand,
//              reconstruct W, lerp, normalise, store the quaternion.
//     phase 4  one thread per (request, animated translation / scale sub-track).
//   The passes are compacted per sub-track class, so warps do not diverge between animated and constant bones; the clip
//   image built at upload (layout.h) gives every thread its operands with 16 byte loads and no dependency on other threads.
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
		constexpr uint32_t k_target_items_per_block = 2048;

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
		// TMA bulk copy + mbarrier (PTX ISA: cp.async.bulk, mbarrier)
		// ---------------------------------------------------------------------------------------------------
		__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

		__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
		{
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}

		__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
		{
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
		}

		__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
		{
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
		}

		__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
		{
			uint32_t done;
			do
			{
				asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
					: "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
			} while (!done);
		}

		// 1-D bulk tensor-less TMA copy global -> shared (SASS: UBLKCP); dst, src and bytes are multiples of 16
		__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
		{
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
		}

		// ---------------------------------------------------------------------------------------------------
		// per request state, written by one thread, read by every item thread of the request
		// ---------------------------------------------------------------------------------------------------
		struct alignas(16) ReqState
		{
			const uint8_t* image;
			uint8_t* out;
			float    alpha;
			uint32_t num_tracks;			// 0 => nothing to decode (invalid request or empty clip)
			uint32_t clip_flags;
			uint32_t single_segment;
			uint32_t entries_off[2];		// image relative Entry tables of the two key frames' segments
			uint32_t stream_off[2];			// image relative streams
			uint32_t bit_base[2];			// staged: bit of the key frame inside its shared memory window; else key_frame_bit_offsets
			uint32_t word_base[2];			// staged: first word of the window inside the block's staging area
			uint32_t anim_off;
			uint32_t bone_table_off;
			uint32_t const_rot_off;
			uint32_t const_vec_off;
			uint32_t num_animated[3];
			uint32_t num_constant_trans;
			// extras reported by the seek parity hook
			float    sample_time;
			uint32_t kf_bit[2];
			uint32_t segment_index[2];
			uint32_t blob_format_off[2];
			uint32_t blob_range_off[2];
			uint32_t blob_animated_off[2];
			uint32_t pose_bits[2];
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

			const uint8_t* image = p.data + clip.data_offset;

			uint32_t looping_policy;
			float duration;
			resolve_looping(p, clip, looping_policy, duration);

			float sample_time = request.sample_time;
			if (p.clamp_sample_time)
				sample_time = fminf(fmaxf(sample_time, 0.0f), duration);		// rtm::scalar_clamp, :215-216

			uint32_t key_frame0, key_frame1;
			float alpha;
			find_key_frames(clip.num_samples, clip.sample_rate, sample_time, p.rounding_policy, looping_policy, key_frame0, key_frame1, alpha);

			const SegDesc* segs = reinterpret_cast<const SegDesc*>(image + clip.seg_table_offset);
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
				const uint32_t* start_indices = reinterpret_cast<const uint32_t*>(image + clip.start_indices_offset);
				const uint32_t approx_segment_index = key_frame0 / clip.samples_per_segment;
				const uint32_t start_segment_index = approx_segment_index > 0 ? approx_segment_index - 1 : 0;
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
						break;
					}
				}
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

			rs.image = image;
			rs.alpha = alpha;
			rs.num_tracks = clip.num_tracks;
			rs.clip_flags = clip.flags;
			rs.single_segment = segment_index0 == segment_index1;
			rs.kf_bit[0] = segment_key_frame0 * seg0.pose_bit_size;				// :558-559
			rs.kf_bit[1] = segment_key_frame1 * seg1.pose_bit_size;
			rs.bit_base[0] = rs.kf_bit[0];
			rs.bit_base[1] = rs.kf_bit[1];
			rs.word_base[0] = rs.word_base[1] = 0;
			rs.stream_off[0] = seg0.stream_offset;
			rs.stream_off[1] = seg1.stream_offset;
			rs.entries_off[0] = seg0.entries_offset;
			rs.entries_off[1] = seg1.entries_offset;
			rs.pose_bits[0] = seg0.pose_bit_size;
			rs.pose_bits[1] = seg1.pose_bit_size;
			rs.anim_off = clip.anim_table_offset;
			rs.bone_table_off = clip.bone_table_offset;
			rs.const_rot_off = clip.const_rot_offset;
			rs.const_vec_off = clip.const_vec_offset;
			rs.num_constant_trans = clip.num_constant[1];
			for (int k = 0; k < 3; ++k)
				rs.num_animated[k] = clip.num_animated[k];
			rs.sample_time = sample_time;
			rs.segment_index[0] = segment_index0;
			rs.segment_index[1] = segment_index1;
			rs.blob_format_off[0] = seg0.blob_format_offset;
			rs.blob_format_off[1] = seg1.blob_format_offset;
			rs.blob_range_off[0] = seg0.blob_range_offset;
			rs.blob_range_off[1] = seg1.blob_range_offset;
			rs.blob_animated_off[0] = seg0.blob_animated_offset;
			rs.blob_animated_off[1] = seg1.blob_animated_offset;
			rs.looping_policy = looping_policy;
		}

		// ---------------------------------------------------------------------------------------------------
		// bit stream reads. Streams are stored as byte-swapped 32-bit words (clipset.cpp append_stream): word i holds the
		// stream bits [32 i, 32 i + 32) MSB first, so the 32 bits that start at any bit are one funnel shift of two words.
		// ---------------------------------------------------------------------------------------------------
		template<bool STAGED>
		__device__ __forceinline__ uint32_t read_bits32(const ReqState& rs, const uint32_t* s_stage, int k, uint32_t bit_offset)
		{
			// unpack_vector3_96_unsafe, math/vector4_packing.h:482-503
			const uint32_t bit = rs.bit_base[k] + bit_offset;
			uint32_t hi, lo;
			if (STAGED)
			{
				const uint32_t* w = s_stage + rs.word_base[k] + (bit >> 5);
				hi = w[0];
				lo = w[1];
			}
			else
			{
				const uint32_t* w = reinterpret_cast<const uint32_t*>(rs.image + rs.stream_off[k]) + (bit >> 5);
				hi = __ldg(w);
				lo = __ldg(w + 1);
			}
			return __funnelshift_l(lo, hi, bit & 31);
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
		__device__ __forceinline__ Entry load_entry(const ReqState& rs, int k, uint32_t slot)
		{
			const uint4 v = __ldg(reinterpret_cast<const uint4*>(rs.image + rs.entries_off[k]) + slot);
			Entry e;
			e.offset_code = v.x; e.range_lo = v.y; e.range_hi = v.z; e.inv_max = __uint_as_float(v.w);
			return e;
		}

		// Raw integers of one animated sample: x, y, z (quantised integers or raw float bits), shared by the decode and by the
		// parity hook (unpack_animated_quat / unpack_animated_vector3 integer stage).
		template<bool STAGED>
		__device__ __forceinline__ void unpack_sample_ints(const ReqState& rs, const uint32_t* s_stage, int k, const Entry& e, bool four_components,
			uint32_t& xi, uint32_t& yi, uint32_t& zi, uint32_t& wi)
		{
			const uint32_t code = e.offset_code & 0xFFu;
			const uint32_t bit_offset = e.offset_code >> 8;
			wi = 0;
			if (code == 0)
			{
				// constant inside the segment: the 3 x 16 bit sample was gathered from the segment range bytes at upload
				// (animated_track_cache.transform.h:552-587; unpack_vector3_u48_unsafe, math/vector4_packing.h:628-653)
				xi = e.range_lo & 0xFFFFu;
				yi = e.range_lo >> 16;
				zi = e.range_hi & 0xFFFFu;
			}
			else if (code & k_entry_raw)
			{
				xi = read_bits32<STAGED>(rs, s_stage, k, bit_offset);
				yi = read_bits32<STAGED>(rs, s_stage, k, bit_offset + 32);
				zi = read_bits32<STAGED>(rs, s_stage, k, bit_offset + 64);
				if (four_components)
					wi = read_bits32<STAGED>(rs, s_stage, k, bit_offset + 96);
			}
			else
			{
				// unpack_vector3_uXX_unsafe, math/vector4_packing.h:947-971
				const uint32_t shift = 32 - code;
				xi = read_bits32<STAGED>(rs, s_stage, k, bit_offset) >> shift;
				yi = read_bits32<STAGED>(rs, s_stage, k, bit_offset + code) >> shift;
				zi = read_bits32<STAGED>(rs, s_stage, k, bit_offset + code * 2) >> shift;
			}
		}

		// One animated rotation sample after range expansion and W reconstruction.
		// SINGLE == false: decompress_tracks flavour (unpack_animated_quat + remap_segment_range_data4 + remap_clip_range_data4,
		//                  animated_track_cache.transform.h:515-687,302-350,391-466): ignored ranges still multiply by 1 and add 0.
		// SINGLE == true : decompress_track flavour (unpack_single_animated_quat, :689-869): ignored ranges are skipped.
		template<bool SINGLE, bool STAGED>
		__device__ __forceinline__ void decode_animated_rotation(const ReqState& rs, const uint32_t* s_stage, int k, const Entry& e,
			const float4& clip_extent, const float4& clip_min, float out[4])
		{
			const bool rot_full = (rs.clip_flags & k_clip_rot_full) != 0;
			uint32_t xi, yi, zi, wi;
			unpack_sample_ints<STAGED>(rs, s_stage, k, e, rot_full, xi, yi, zi, wi);
			const uint32_t code = e.offset_code & 0xFFu;

			if (!(rs.clip_flags & k_clip_rot_variable))
			{
				out[0] = __uint_as_float(xi);
				out[1] = __uint_as_float(yi);
				out[2] = __uint_as_float(zi);
				out[3] = rot_full ? __uint_as_float(wi) : quat_w(out[0], out[1], out[2]);
				return;
			}

			float x, y, z;
			const bool is_raw = (code & k_entry_raw) != 0;
			const bool ignore_segment = code == 0 || is_raw;
			const bool ignore_clip = is_raw;
			if (is_raw)
			{
				x = __uint_as_float(xi); y = __uint_as_float(yi); z = __uint_as_float(zi);
			}
			else
			{
				// code 0: 1 / 65535, else 1 / (2^code - 1) -- both stored in the entry
				x = fmul(u2f(xi), e.inv_max); y = fmul(u2f(yi), e.inv_max); z = fmul(u2f(zi), e.inv_max);
			}

			if ((rs.clip_flags & k_clip_has_segments) && (!SINGLE || !ignore_segment))
			{
				float min_x = 0.0f, min_y = 0.0f, min_z = 0.0f, ext_x = 1.0f, ext_y = 1.0f, ext_z = 1.0f;
				if (!ignore_segment)
				{
					// unpack_segment_range_data, :157-298: u8 * (1 / 255)
					const float n = 1.0f / 255.0f;
					min_x = fmul(u2f(e.range_lo & 0xFFu), n);
					min_y = fmul(u2f((e.range_lo >> 8) & 0xFFu), n);
					min_z = fmul(u2f((e.range_lo >> 16) & 0xFFu), n);
					ext_x = fmul(u2f(e.range_lo >> 24), n);
					ext_y = fmul(u2f(e.range_hi & 0xFFu), n);
					ext_z = fmul(u2f((e.range_hi >> 8) & 0xFFu), n);
				}
				x = fmuladd(x, ext_x, min_x);
				y = fmuladd(y, ext_y, min_y);
				z = fmuladd(z, ext_z, min_z);
			}

			if (!SINGLE || !ignore_clip)
			{
				// remap_clip_range_data4, :391-466
				const float ext_x = ignore_clip ? 1.0f : clip_extent.x, ext_y = ignore_clip ? 1.0f : clip_extent.y, ext_z = ignore_clip ? 1.0f : clip_extent.z;
				const float min_x = ignore_clip ? 0.0f : clip_min.x, min_y = ignore_clip ? 0.0f : clip_min.y, min_z = ignore_clip ? 0.0f : clip_min.z;
				x = fmuladd(x, ext_x, min_x);
				y = fmuladd(y, ext_y, min_y);
				z = fmuladd(z, ext_z, min_z);
			}

			out[0] = x; out[1] = y; out[2] = z;
			out[3] = quat_w(x, y, z);
		}

		// unpack_animated_vector3 / unpack_single_animated_vector3, animated_track_cache.transform.h:871-990,992-1102
		template<bool STAGED>
		__device__ __forceinline__ void decode_animated_vector3(const ReqState& rs, const uint32_t* s_stage, int k, const Entry& e, bool variable,
			const float4& clip_extent, const float4& clip_min, float out[3])
		{
			uint32_t xi, yi, zi, wi;
			unpack_sample_ints<STAGED>(rs, s_stage, k, e, false, xi, yi, zi, wi);
			const uint32_t code = e.offset_code & 0xFFu;

			if (!variable || (code & k_entry_raw))
			{
				out[0] = __uint_as_float(xi); out[1] = __uint_as_float(yi); out[2] = __uint_as_float(zi);
				return;
			}

			float x = fmul(u2f(xi), e.inv_max), y = fmul(u2f(yi), e.inv_max), z = fmul(u2f(zi), e.inv_max);
			if (code != 0 && (rs.clip_flags & k_clip_has_segments))
			{
				// unpack_vector3_u24_unsafe min then extent, math/vector4_packing.h:781-818
				const float n = 1.0f / 255.0f;
				x = fmuladd(x, fmul(u2f(e.range_lo >> 24), n), fmul(u2f(e.range_lo & 0xFFu), n));
				y = fmuladd(y, fmul(u2f(e.range_hi & 0xFFu), n), fmul(u2f((e.range_lo >> 8) & 0xFFu), n));
				z = fmuladd(z, fmul(u2f((e.range_hi >> 8) & 0xFFu), n), fmul(u2f((e.range_lo >> 16) & 0xFFu), n));
			}
			// clip range (:949-958)
			out[0] = fmuladd(x, clip_extent.x, clip_min.x);
			out[1] = fmuladd(y, clip_extent.y, clip_min.y);
			out[2] = fmuladd(z, clip_extent.z, clip_min.z);
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

		// ---- the device track_writer: write_rotation / write_translation / write_scale ----
		__device__ __forceinline__ void write_rotation(uint32_t layout, uint8_t* bone, const float q[4])
		{
			if (layout == ACLB200_LAYOUT_QVV48)
				*reinterpret_cast<float4*>(bone) = make_float4(q[0], q[1], q[2], q[3]);
			else
			{
				float2* dst = reinterpret_cast<float2*>(bone);		// 40 byte bones are 8 byte aligned
				dst[0] = make_float2(q[0], q[1]);
				dst[1] = make_float2(q[2], q[3]);
			}
		}

		__device__ __forceinline__ void write_vector(uint32_t layout, uint8_t* bone, uint32_t kind, const float v[3])
		{
			if (layout == ACLB200_LAYOUT_QVV48)
				*reinterpret_cast<float4*>(bone + 16 * kind) = make_float4(v[0], v[1], v[2], 0.0f);
			else if (kind == 1)
			{
				*reinterpret_cast<float2*>(bone + 16) = make_float2(v[0], v[1]);
				*reinterpret_cast<float*>(bone + 24) = v[2];
			}
			else
			{
				*reinterpret_cast<float*>(bone + 28) = v[0];
				*reinterpret_cast<float2*>(bone + 32) = make_float2(v[1], v[2]);
			}
		}

		// Interpolation of two decoded rotation samples: unpack_rotation_group, animated_track_cache.transform.h:1463-1474,1477-1661
		// (SINGLE: unpack_rotation_within_group, :1709-1765)
		template<int NORM, bool PER_TRACK, bool SINGLE>
		__device__ __forceinline__ void interpolate_rotation(const DecodeParams& p, uint32_t clip_flags, float s0[4], float s1[4], float alpha, uint32_t policy, float rotation[4])
		{
			const bool interpolate = should_interpolate(p, clip_flags, alpha);
			if (SINGLE)
			{
				if (interpolate)
					rtm_quat_lerp(s0, s1, alpha, NORM >= ACLB200_NORMALIZE_LERP_ONLY, rotation);
				else
				{
#pragma unroll
					for (int i = 0; i < 4; ++i)
						rotation[i] = alpha <= 0.0f ? s0[i] : s1[i];
					if (NORM == ACLB200_NORMALIZE_ALWAYS && !(clip_flags & k_clip_rot_full))
						rtm_quat_normalize(rotation);
				}
				return;
			}

			if (NORM == ACLB200_NORMALIZE_ALWAYS && !(clip_flags & k_clip_rot_full) && (PER_TRACK || !interpolate))
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

		// One animated rotation sub-track of one request (phase 3)
		template<int NORM, bool PER_TRACK, bool SINGLE, bool STAGED>
		__device__ __forceinline__ uint32_t animated_rotation(const DecodeParams& p, const ReqState& rs, const uint32_t* s_stage, uint32_t rank, float alpha_in, float rotation[4])
		{
			const float4* anim = reinterpret_cast<const float4*>(rs.image + rs.anim_off) + size_t(rank) * 2;
			const float4 clip_extent = __ldg(anim + 0);		// .w carries the bone index
			const float4 clip_min = __ldg(anim + 1);
			const uint32_t bone = __float_as_uint(clip_extent.w);
			const Entry e0 = load_entry(rs, 0, rank);
			const Entry e1 = rs.single_segment ? e0 : load_entry(rs, 1, rank);

			float s0[4], s1[4];
			decode_animated_rotation<SINGLE, STAGED>(rs, s_stage, 0, e0, clip_extent, clip_min, s0);
			decode_animated_rotation<SINGLE, STAGED>(rs, s_stage, 1, e1, clip_extent, clip_min, s1);

			const uint32_t policy = PER_TRACK ? track_rounding_policy(p, bone) : ACLB200_ROUND_NONE;
			const float alpha = (SINGLE && PER_TRACK) ? apply_rounding_policy(alpha_in, policy) : alpha_in;	// :1975-1983
			interpolate_rotation<NORM, PER_TRACK, SINGLE>(p, rs.clip_flags, s0, s1, alpha, policy, rotation);
			return bone;
		}

		// One animated translation (kind 1) or scale (kind 2) sub-track of one request (phase 4):
		// unpack_translation_group / consume_translation, animated_track_cache.transform.h:1774-1836,1889-1894
		template<bool PER_TRACK, bool SINGLE, bool STAGED>
		__device__ __forceinline__ uint32_t animated_vector(const DecodeParams& p, const ReqState& rs, const uint32_t* s_stage, uint32_t kind, uint32_t rank, float alpha_in, float value[3])
		{
			const uint32_t slot = rs.num_animated[0] + (kind == 2 ? rs.num_animated[1] : 0u) + rank;
			const float4* anim = reinterpret_cast<const float4*>(rs.image + rs.anim_off) + size_t(slot) * 2;
			const float4 clip_extent = __ldg(anim + 0);
			const float4 clip_min = __ldg(anim + 1);
			const uint32_t bone = __float_as_uint(clip_extent.w);
			const Entry e0 = load_entry(rs, 0, slot);
			const Entry e1 = rs.single_segment ? e0 : load_entry(rs, 1, slot);
			const bool variable = (rs.clip_flags & (kind == 1 ? k_clip_trans_variable : k_clip_scale_variable)) != 0;

			float s0[3], s1[3];
			decode_animated_vector3<STAGED>(rs, s_stage, 0, e0, variable, clip_extent, clip_min, s0);
			decode_animated_vector3<STAGED>(rs, s_stage, 1, e1, variable, clip_extent, clip_min, s1);

			const uint32_t policy = PER_TRACK ? track_rounding_policy(p, bone) : ACLB200_ROUND_NONE;
			const float alpha = (SINGLE && PER_TRACK) ? apply_rounding_policy(alpha_in, policy) : alpha_in;
#pragma unroll
			for (int i = 0; i < 3; ++i)
			{
				if (!SINGLE && PER_TRACK && policy == ACLB200_ROUND_FLOOR)
					value[i] = s0[i];
				else if (!SINGLE && PER_TRACK && policy == ACLB200_ROUND_CEIL)
					value[i] = s1[i];
				else if (!SINGLE && PER_TRACK && policy == ACLB200_ROUND_NEAREST)
					value[i] = alpha < 0.5f ? s0[i] : s1[i];
				else
					value[i] = lerp(s0[i], s1[i], alpha);
			}
			return bone;
		}

		// Constant and default sub-tracks of one bone (phase 2): unpack_default_* / unpack_constant_*_sub_tracks,
		// decompression.transform.h:574-748,881-1072,1201-1430; constant rotations had their W reconstructed (and normalised) at upload
		template<int NORM, bool SINGLE>
		__device__ __forceinline__ void constant_sub_tracks(const DecodeParams& p, const ReqState& rs, uint32_t bone, uint64_t desc, uint8_t* out_bone)
		{
			// rotation
			{
				const uint32_t type = uint32_t(desc) & 3;
				const uint32_t rank = (uint32_t(desc) >> 2) & k_bone_index_mask;
				float q[4];
				if (type == 0)
				{
					if (default_value(p, 0, bone, rs.clip_flags, q))
						write_rotation(p.layout, out_bone, q);
				}
				else if (type == 1)
				{
					const float4* table = reinterpret_cast<const float4*>(rs.image + rs.const_rot_off) + size_t(rank) * 2;
					if (SINGLE && NORM == ACLB200_NORMALIZE_ALWAYS && !(rs.clip_flags & k_clip_rot_full))
					{
						// unpack_rotation_within_group normalises with rtm::quat_normalize, constant_track_cache.transform.h:255-258
						const float4 v = __ldg(table);
						q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
						rtm_quat_normalize(q);
					}
					else
					{
						const float4 v = __ldg(table + (NORM == ACLB200_NORMALIZE_ALWAYS ? 1 : 0));
						q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
					}
					write_rotation(p.layout, out_bone, q);
				}
			}
			// translation, scale
#pragma unroll
			for (uint32_t kind = 1; kind <= 2; ++kind)
			{
				const uint32_t bits = uint32_t(desc >> (k_bone_kind_shift * kind));
				// clips without scale: every bone takes the default (decompression.transform.h:1653-1680,1806-1822)
				const uint32_t type = (kind == 2 && !(rs.clip_flags & k_clip_has_scale)) ? 0u : (bits & 3);
				const uint32_t rank = (bits >> 2) & k_bone_index_mask;
				float v[4];
				if (type == 0)
				{
					if (default_value(p, kind, bone, rs.clip_flags, v))
						write_vector(p.layout, out_bone, kind, v);
				}
				else if (type == 1)
				{
					const float4 c = __ldg(reinterpret_cast<const float4*>(rs.image + rs.const_vec_off) + (kind == 2 ? rs.num_constant_trans : 0u) + rank);
					v[0] = c.x; v[1] = c.y; v[2] = c.z;
					write_vector(p.layout, out_bone, kind, v);
				}
			}
		}

		__device__ __forceinline__ uint32_t fast_div(uint32_t value, uint32_t magic)
		{
			return magic != 0 ? __umulhi(value, magic) : value;
		}

		// ---------------------------------------------------------------------------------------------------
		// kernels
		// ---------------------------------------------------------------------------------------------------
		template<int NORM, bool PER_TRACK, bool STAGED>
		__global__ void __launch_bounds__(k_threads_per_block)
		transform_decompress_tracks_kernel(const DecodeParams p)
		{
			extern __shared__ __align__(16) uint8_t s_dynamic[];
			__shared__ ReqState s_req[k_max_requests_per_block];
			__shared__ __align__(8) uint64_t s_barrier;
			const uint32_t* s_stage = reinterpret_cast<const uint32_t*>(s_dynamic);

			const uint32_t first_request = blockIdx.x * p.requests_per_block;
			const uint32_t num_requests = min(p.requests_per_block, p.num_requests - first_request);

			if (STAGED)
			{
				if (threadIdx.x == 0)
					mbar_init(&s_barrier, num_requests);
				__syncthreads();
			}

			// ---- phase 1: seek + stage the two key frames ----
			if (threadIdx.x < num_requests)
			{
				ReqState rs;
				seek_transform(p, first_request + threadIdx.x, rs);
				rs.out = p.out + uint64_t(first_request + threadIdx.x) * p.pose_stride;
				if (STAGED)
				{
					if (rs.num_tracks != 0 && (rs.num_animated[0] | rs.num_animated[1] | rs.num_animated[2]) != 0)
					{
						uint32_t src_byte[2], bytes[2];
#pragma unroll
						for (int k = 0; k < 2; ++k)
						{
							src_byte[k] = (rs.kf_bit[k] >> 3) & ~15u;
							rs.bit_base[k] = rs.kf_bit[k] - src_byte[k] * 8;
							rs.word_base[k] = (threadIdx.x * 2 + k) * (p.stage_bytes >> 2);
							bytes[k] = min((((rs.bit_base[k] + rs.pose_bits[k] + 7) >> 3) + 8 + 15) & ~15u, p.stage_bytes);
						}
						mbar_arrive_expect_tx(&s_barrier, bytes[0] + bytes[1]);
#pragma unroll
						for (int k = 0; k < 2; ++k)
							bulk_copy_g2s(s_dynamic + size_t(rs.word_base[k]) * 4, rs.image + rs.stream_off[k] + src_byte[k], bytes[k], &s_barrier);
					}
					else
						mbar_arrive(&s_barrier);
				}
				s_req[threadIdx.x] = rs;
			}
			__syncthreads();

			// ---- phase 2: constant and default sub-tracks, one thread per (request, bone) ----
			{
				const uint32_t num_slots = num_requests * p.max_tracks;
				for (uint32_t slot = threadIdx.x; slot < num_slots; slot += k_threads_per_block)
				{
					const uint32_t local_request = fast_div(slot, p.magic_tracks);
					const uint32_t bone = slot - local_request * p.max_tracks;
					const ReqState& rs = s_req[local_request];
					if (bone >= rs.num_tracks)
						continue;
					const uint64_t desc = __ldg(reinterpret_cast<const unsigned long long*>(rs.image + rs.bone_table_off) + bone);
					constant_sub_tracks<NORM, false>(p, rs, bone, desc, rs.out + size_t(bone) * p.bone_stride);
				}
			}

			if (STAGED)
				mbar_wait(&s_barrier, 0);

			// ---- phase 3: animated rotations, one thread per (request, animated rotation sub-track) ----
			if (p.max_animated[0] != 0)
			{
				const uint32_t num_slots = num_requests * p.max_animated[0];
				for (uint32_t slot = threadIdx.x; slot < num_slots; slot += k_threads_per_block)
				{
					const uint32_t local_request = fast_div(slot, p.magic_rot);
					const uint32_t rank = slot - local_request * p.max_animated[0];
					const ReqState& rs = s_req[local_request];
					if (rs.num_tracks == 0 || rank >= rs.num_animated[0])
						continue;
					float rotation[4];
					const uint32_t bone = animated_rotation<NORM, PER_TRACK, false, STAGED>(p, rs, s_stage, rank, rs.alpha, rotation);
					write_rotation(p.layout, rs.out + size_t(bone) * p.bone_stride, rotation);
				}
			}

			// ---- phase 4: animated translations then scales ----
			const uint32_t max_vectors = p.max_animated[1] + p.max_animated[2];
			if (max_vectors != 0)
			{
				const uint32_t num_slots = num_requests * max_vectors;
				for (uint32_t slot = threadIdx.x; slot < num_slots; slot += k_threads_per_block)
				{
					const uint32_t local_request = fast_div(slot, p.magic_vec);
					uint32_t rank = slot - local_request * max_vectors;
					const ReqState& rs = s_req[local_request];
					uint32_t kind = 1;
					if (rank >= p.max_animated[1])
					{
						rank -= p.max_animated[1];
						kind = 2;
					}
					if (rs.num_tracks == 0 || rank >= rs.num_animated[kind])
						continue;
					float value[3];
					const uint32_t bone = animated_vector<PER_TRACK, false, STAGED>(p, rs, s_stage, kind, rank, rs.alpha, value);
					write_vector(p.layout, rs.out + size_t(bone) * p.bone_stride, kind, value);
				}
			}
		}

		// decompress_track_v0, decompression.transform.h:1753-2050: one thread per request, one bone each
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
				return;		// :1766-1768: invalid track index, nothing is written
			uint8_t* out_bone = p.out + uint64_t(request) * p.bone_stride;
			const uint64_t desc = __ldg(reinterpret_cast<const unsigned long long*>(rs.image + rs.bone_table_off) + bone);
			constant_sub_tracks<NORM, true>(p, rs, bone, desc, out_bone);

			if ((uint32_t(desc) & 3) == 2)
			{
				float rotation[4];
				animated_rotation<NORM, PER_TRACK, true, false>(p, rs, nullptr, (uint32_t(desc) >> 2) & k_bone_index_mask, rs.alpha, rotation);
				write_rotation(p.layout, out_bone, rotation);
			}
#pragma unroll
			for (uint32_t kind = 1; kind <= 2; ++kind)
			{
				const uint32_t bits = uint32_t(desc >> (k_bone_kind_shift * kind));
				if ((bits & 3) == 2 && (kind == 1 || (rs.clip_flags & k_clip_has_scale)))
				{
					float value[3];
					animated_vector<PER_TRACK, true, false>(p, rs, nullptr, kind, (bits >> 2) & k_bone_index_mask, rs.alpha, value);
					write_vector(p.layout, out_bone, kind, value);
				}
			}
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
					st.animated_offsets[k] = rs.blob_animated_off[k];
					st.format_offsets[k] = rs.blob_format_off[k];
					st.range_offsets[k] = rs.blob_range_off[k];
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
			for (uint32_t slot = threadIdx.x; slot < total && slot < p.debug_max_sub_tracks; slot += blockDim.x)
			{
				const Entry e = load_entry(rs, k, slot);
				const bool four = slot < rs.num_animated[0] && (rs.clip_flags & k_clip_rot_full);
				uint32_t xi, yi, zi, wi;
				unpack_sample_ints<false>(rs, nullptr, k, e, four, xi, yi, zi, wi);
				uint32_t* dst = out + (size_t(request) * p.debug_max_sub_tracks + slot) * 4;
				dst[0] = xi; dst[1] = yi; dst[2] = zi;
				dst[3] = e.offset_code & 0xFFu;
			}
		}

		// ---------------------------------------------------------------------------------------------------
		// scalar tracks: seek_v0 + decompress_tracks_v0 / decompress_track_v0, decompression/impl/decompression.scalar.h:181-705
		// ---------------------------------------------------------------------------------------------------
		struct ScalarReqState
		{
			const uint8_t* image;
			const ScalarTrackDesc* tracks;
			uint8_t* out;
			float    alpha;
			uint32_t num_tracks;
			uint32_t kf_bit[2];
			uint32_t constant_off;
			uint32_t range_off;
			uint32_t stream_off;
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

			rs.image = p.data + clip.data_offset;
			rs.tracks = reinterpret_cast<const ScalarTrackDesc*>(rs.image + clip.bone_table_offset);
			rs.alpha = alpha;
			rs.num_tracks = clip.num_tracks;
			rs.kf_bit[0] = key_frame0 * clip.num_animated_total;		// num_bits_per_frame, decompression.scalar.h:208-209
			rs.kf_bit[1] = key_frame1 * clip.num_animated_total;
			rs.constant_off = clip.const_rot_offset;
			rs.range_off = clip.const_vec_offset;
			rs.stream_off = clip.seg_table_offset;
		}

		__device__ __forceinline__ uint32_t read_stream32(const uint32_t* words, uint32_t bit)
		{
			const uint32_t hi = __ldg(words + (bit >> 5));
			const uint32_t lo = __ldg(words + (bit >> 5) + 1);
			return __funnelshift_l(lo, hi, bit & 31);
		}

		template<int COMPONENTS, bool PER_TRACK>
		__device__ __forceinline__ void decode_scalar_track(const DecodeParams& p, const ScalarReqState& rs, uint32_t track, float* out)
		{
			const uint4 raw = __ldg(reinterpret_cast<const uint4*>(rs.tracks) + track);
			const uint32_t num_bits = raw.y & 0xFFu;
			const uint32_t value_index = raw.y >> 8;
			const float inv_max = __uint_as_float(raw.z);
			float alpha = rs.alpha;
			if (PER_TRACK)
				alpha = apply_rounding_policy(rs.alpha, track_rounding_policy(p, track));	// decompression.scalar.h:235-247,273-280

			if (num_bits == 0)
			{
				const float* constants = reinterpret_cast<const float*>(rs.image + rs.constant_off) + value_index;
#pragma unroll
				for (int c = 0; c < COMPONENTS; ++c)
					out[c] = __ldg(constants + c);
				return;
			}

			const uint32_t* words = reinterpret_cast<const uint32_t*>(rs.image + rs.stream_off);
			const uint32_t bit0 = rs.kf_bit[0] + raw.x;
			const uint32_t bit1 = rs.kf_bit[1] + raw.x;
			if (num_bits == 32)
			{
#pragma unroll
				for (int c = 0; c < COMPONENTS; ++c)
				{
					const float v0 = __uint_as_float(read_stream32(words, bit0 + 32 * c));
					const float v1 = __uint_as_float(read_stream32(words, bit1 + 32 * c));
					out[c] = lerp(v0, v1, alpha);
				}
				return;
			}

			const float* range = reinterpret_cast<const float*>(rs.image + rs.range_off) + value_index;
#pragma unroll
			for (int c = 0; c < COMPONENTS; ++c)
			{
				const uint32_t i0 = read_stream32(words, bit0 + num_bits * c) >> (32 - num_bits);
				const uint32_t i1 = read_stream32(words, bit1 + num_bits * c) >> (32 - num_bits);
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
				const uint32_t local_request = fast_div(slot, p.magic_tracks);
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

		uint32_t division_magic(uint32_t divisor)
		{
			// floor(v / d) == mulhi(v, magic) for v < 2^32 / d, which every slot index here satisfies (v < 2^18 + 2048 items, d <= 2^18)
			return divisor <= 1 ? 0u : uint32_t((uint64_t(1) << 32) / divisor) + 1u;
		}

		template<int NORM, bool PER_TRACK>
		cudaError_t launch_tracks(const DecodeParams& params, cudaStream_t stream)
		{
			const uint32_t blocks = (params.num_requests + params.requests_per_block - 1) / params.requests_per_block;
			if (params.stage_bytes != 0)
				transform_decompress_tracks_kernel<NORM, PER_TRACK, true><<<blocks, k_threads_per_block, params.smem_bytes, stream>>>(params);
			else
				transform_decompress_tracks_kernel<NORM, PER_TRACK, false><<<blocks, k_threads_per_block, 0, stream>>>(params);
			return cudaGetLastError();
		}

		template<int NORM, bool PER_TRACK>
		cudaError_t launch_track(const DecodeParams& params, cudaStream_t stream)
		{
			const uint32_t blocks = (params.num_requests + 127) / 128;
			transform_decompress_track_kernel<NORM, PER_TRACK><<<blocks, 128, 0, stream>>>(params);
			return cudaGetLastError();
		}

		template<int NORM, bool PER_TRACK>
		cudaError_t set_smem_attribute(int optin_limit, int& min_available)
		{
			// the opt-in limit covers static + dynamic shared memory
			cudaFuncAttributes attributes;
			cudaError_t error = cudaFuncGetAttributes(&attributes, transform_decompress_tracks_kernel<NORM, PER_TRACK, true>);
			if (error != cudaSuccess)
				return error;
			const int available = optin_limit - int(attributes.sharedSizeBytes);
			if (available < min_available)
				min_available = available;
			return cudaFuncSetAttribute(transform_decompress_tracks_kernel<NORM, PER_TRACK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, available);
		}
	}

	// Called once per context: lets the staged kernels use large dynamic shared memory windows. `max_dynamic_smem` comes in as the
	// device's opt-in limit and goes out as what a launch may actually request.
	cudaError_t configure_kernels(int& max_dynamic_smem)
	{
		const int optin_limit = max_dynamic_smem - 1024;
		int available = optin_limit;
		cudaError_t error = set_smem_attribute<0, false>(optin_limit, available);
		if (error == cudaSuccess) error = set_smem_attribute<0, true>(optin_limit, available);
		if (error == cudaSuccess) error = set_smem_attribute<1, false>(optin_limit, available);
		if (error == cudaSuccess) error = set_smem_attribute<1, true>(optin_limit, available);
		if (error == cudaSuccess) error = set_smem_attribute<2, false>(optin_limit, available);
		if (error == cudaSuccess) error = set_smem_attribute<2, true>(optin_limit, available);
		max_dynamic_smem = available;
		return error;
	}

	// requests_per_block, the division magics and the shared memory staging of a launch
	void plan_launch(DecodeParams& params, uint32_t max_key_frame_bytes, int max_dynamic_smem)
	{
		const uint32_t max_tracks = params.max_tracks == 0 ? 1 : params.max_tracks;
		uint32_t requests_per_block = k_target_items_per_block / max_tracks;
		if (requests_per_block < 1) requests_per_block = 1;
		if (requests_per_block > k_max_requests_per_block) requests_per_block = k_max_requests_per_block;

		// bytes per staged key frame: up to 15 bytes of alignment skew + the key frame + one extra word for the funnel shift
		uint32_t stage_bytes = (max_key_frame_bytes + 48 + 15) & ~15u;
		if (max_key_frame_bytes == 0)
			stage_bytes = 0;
		// keep several blocks resident per SM: at most ~40 KB of staging per block, else fewer requests per block
		const uint32_t staging_budget = 40u * 1024u;
		while (stage_bytes != 0 && requests_per_block > 1 && requests_per_block * 2 * stage_bytes > staging_budget)
			requests_per_block = (requests_per_block + 1) / 2;
		if (stage_bytes != 0 && uint64_t(requests_per_block) * 2 * stage_bytes > uint64_t(max_dynamic_smem > 4096 ? max_dynamic_smem - 4096 : 0))
			stage_bytes = 0;		// a single key frame pair does not fit: read the streams from global memory instead

		params.requests_per_block = requests_per_block;
		params.stage_bytes = stage_bytes;
		params.smem_bytes = requests_per_block * 2 * stage_bytes;
		params.magic_tracks = division_magic(max_tracks);
		params.magic_rot = division_magic(params.max_animated[0]);
		params.magic_vec = division_magic(params.max_animated[1] + params.max_animated[2]);
	}

	cudaError_t launch_transform_decompress_tracks(const DecodeParams& params, uint32_t /*math_mode*/, cudaStream_t stream)
	{
		const bool per_track = params.per_track_rounding != 0;
		switch (params.normalization)
		{
		case ACLB200_NORMALIZE_NEVER: return per_track ? launch_tracks<0, true>(params, stream) : launch_tracks<0, false>(params, stream);
		case ACLB200_NORMALIZE_LERP_ONLY: return per_track ? launch_tracks<1, true>(params, stream) : launch_tracks<1, false>(params, stream);
		default: return per_track ? launch_tracks<2, true>(params, stream) : launch_tracks<2, false>(params, stream);
		}
	}

	cudaError_t launch_transform_decompress_track(const DecodeParams& params, uint32_t /*math_mode*/, cudaStream_t stream)
	{
		const bool per_track = params.per_track_rounding != 0;
		switch (params.normalization)
		{
		case ACLB200_NORMALIZE_NEVER: return per_track ? launch_track<0, true>(params, stream) : launch_track<0, false>(params, stream);
		case ACLB200_NORMALIZE_LERP_ONLY: return per_track ? launch_track<1, true>(params, stream) : launch_track<1, false>(params, stream);
		default: return per_track ? launch_track<2, true>(params, stream) : launch_track<2, false>(params, stream);
		}
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
