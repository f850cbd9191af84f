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
#include "device_common.cuh"

namespace aclb200
{
	using namespace dev;

	namespace
	{
		// ---------------------------------------------------------------------------------------------------
		// kernels
		// ---------------------------------------------------------------------------------------------------
		// STAGED    : the two key frames of every request are staged in shared memory by the TMA unit (else read from global memory)
		// OUT_STAGED: poses are assembled in shared memory and written out with full-line coalesced 16 byte stores (else every
		//             phase stores its sub-tracks straight to global memory: needed when `skipped` default sub-tracks must keep
		//             what the caller's buffer holds, or when a pose does not fit in shared memory)
		template<int NORM, bool PER_TRACK, bool STAGED, bool OUT_STAGED>
		__global__ void __launch_bounds__(k_threads_per_block)
		transform_decompress_tracks_kernel(const DecodeParams p)
		{
			// dynamic shared memory: ReqState[requests_per_block] | key frame windows | pose staging
			extern __shared__ __align__(16) uint8_t s_dynamic[];
			__shared__ __align__(8) uint64_t s_barrier;
			ReqState* s_req = reinterpret_cast<ReqState*>(s_dynamic);
			uint8_t* s_stage_bytes = s_dynamic + p.smem_stage_offset;
			const uint32_t* s_stage = reinterpret_cast<const uint32_t*>(s_stage_bytes);
			uint8_t* s_out = s_dynamic + p.smem_out_offset;

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
							bulk_copy_g2s(s_stage_bytes + size_t(rs.word_base[k]) * 4, rs.image + rs.stream_off[k] + src_byte[k], bytes[k], &s_barrier);
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
					uint8_t* pose = OUT_STAGED ? s_out + size_t(local_request) * p.smem_pose_bytes : rs.out;
					constant_sub_tracks<NORM, false>(p, rs, bone, desc, pose + size_t(bone) * p.bone_stride);
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
					uint8_t* pose = OUT_STAGED ? s_out + size_t(local_request) * p.smem_pose_bytes : rs.out;
					write_rotation(p.layout, pose + size_t(bone) * p.bone_stride, rotation);
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
					uint8_t* pose = OUT_STAGED ? s_out + size_t(local_request) * p.smem_pose_bytes : rs.out;
					write_vector(p.layout, pose + size_t(bone) * p.bone_stride, kind, value);
				}
			}

			// ---- phase 5: the assembled poses leave shared memory as full, coalesced 16 byte (or 8 byte) stores ----
			if (OUT_STAGED)
			{
				__syncthreads();
				const uint32_t chunks_per_pose = p.smem_pose_bytes >> 4;
				const uint32_t num_chunks = num_requests * chunks_per_pose;
				for (uint32_t slot = threadIdx.x; slot < num_chunks; slot += k_threads_per_block)
				{
					const uint32_t local_request = fast_div(slot, p.magic_chunks);
					const uint32_t byte = (slot - local_request * chunks_per_pose) << 4;
					const uint32_t row_bytes = s_req[local_request].num_tracks * p.bone_stride;
					if (byte >= row_bytes)
						continue;
					const uint8_t* src = s_out + size_t(local_request) * p.smem_pose_bytes + byte;
					uint8_t* dst = p.out + uint64_t(first_request + local_request) * p.pose_stride + byte;
					if (p.out_vector16 && byte + 16 <= row_bytes)
						*reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
					else
					{
						*reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
						if (byte + 8 < row_bytes)
							*reinterpret_cast<uint2*>(dst + 8) = *reinterpret_cast<const uint2*>(src + 8);
					}
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
			const bool staged = params.stage_bytes != 0;
			const bool out_staged = params.smem_pose_bytes != 0;
			if (staged && out_staged)
				transform_decompress_tracks_kernel<NORM, PER_TRACK, true, true><<<blocks, k_threads_per_block, params.smem_bytes, stream>>>(params);
			else if (staged)
				transform_decompress_tracks_kernel<NORM, PER_TRACK, true, false><<<blocks, k_threads_per_block, params.smem_bytes, stream>>>(params);
			else if (out_staged)
				transform_decompress_tracks_kernel<NORM, PER_TRACK, false, true><<<blocks, k_threads_per_block, params.smem_bytes, stream>>>(params);
			else
				transform_decompress_tracks_kernel<NORM, PER_TRACK, false, false><<<blocks, k_threads_per_block, params.smem_bytes, stream>>>(params);
			return cudaGetLastError();
		}

		template<int NORM, bool PER_TRACK>
		cudaError_t launch_track(const DecodeParams& params, cudaStream_t stream)
		{
			const uint32_t blocks = (params.num_requests + 127) / 128;
			transform_decompress_track_kernel<NORM, PER_TRACK><<<blocks, 128, 0, stream>>>(params);
			return cudaGetLastError();
		}

		template<int NORM, bool PER_TRACK, bool STAGED, bool OUT_STAGED>
		cudaError_t set_smem_attribute_one(int optin_limit, int& min_available)
		{
			// the opt-in limit covers static + dynamic shared memory
			cudaFuncAttributes attributes;
			cudaError_t error = cudaFuncGetAttributes(&attributes, transform_decompress_tracks_kernel<NORM, PER_TRACK, STAGED, OUT_STAGED>);
			if (error != cudaSuccess)
				return error;
			const int available = optin_limit - int(attributes.sharedSizeBytes);
			if (available < min_available)
				min_available = available;
			return cudaFuncSetAttribute(transform_decompress_tracks_kernel<NORM, PER_TRACK, STAGED, OUT_STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, available);
		}

		template<int NORM, bool PER_TRACK>
		cudaError_t set_smem_attribute(int optin_limit, int& min_available)
		{
			cudaError_t error = set_smem_attribute_one<NORM, PER_TRACK, true, true>(optin_limit, min_available);
			if (error == cudaSuccess) error = set_smem_attribute_one<NORM, PER_TRACK, true, false>(optin_limit, min_available);
			if (error == cudaSuccess) error = set_smem_attribute_one<NORM, PER_TRACK, false, true>(optin_limit, min_available);
			if (error == cudaSuccess) error = set_smem_attribute_one<NORM, PER_TRACK, false, false>(optin_limit, min_available);
			return error;
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
		if (error == cudaSuccess) error = configure_pipeline_kernels(optin_limit, available);
		max_dynamic_smem = available;
		return error;
	}

	// requests_per_block, the division magics and the shared memory carve-up of a launch
	void plan_launch(DecodeParams& params, uint32_t max_key_frame_bytes, int max_dynamic_smem, bool allow_output_staging)
	{
		const uint32_t max_tracks = params.max_tracks == 0 ? 1 : params.max_tracks;
		const uint32_t budget = uint32_t(max_dynamic_smem > 0 ? max_dynamic_smem : 0);
		// ~28 KB of shared memory per block keeps 8 blocks resident per SM
		const uint32_t block_budget = budget < 28u * 1024u ? budget : 28u * 1024u;

		// bytes per staged key frame: alignment skew + the key frame + the extra word the funnel shift reads, 16 byte granular
		uint32_t stage_bytes = max_key_frame_bytes != 0 ? ((max_key_frame_bytes + 48 + 15) & ~15u) : 0u;
		uint32_t pose_bytes = allow_output_staging ? ((max_tracks * params.bone_stride + 15) & ~15u) : 0u;

		uint32_t requests_per_block = k_target_items_per_block / max_tracks;
		if (requests_per_block < 1) requests_per_block = 1;
		if (requests_per_block > k_max_requests_per_block) requests_per_block = k_max_requests_per_block;

		auto bytes_needed = [&](uint32_t requests) { return requests * (uint32_t(sizeof(ReqState)) + 2 * stage_bytes + pose_bytes); };
		while (requests_per_block > 1 && bytes_needed(requests_per_block) > block_budget)
			--requests_per_block;
		// a single request that does not fit: give up output staging first, then key frame staging
		if (bytes_needed(requests_per_block) > budget)
			pose_bytes = 0;
		if (bytes_needed(requests_per_block) > budget)
			stage_bytes = 0;

		params.requests_per_block = requests_per_block;
		params.stage_bytes = stage_bytes;
		params.smem_pose_bytes = pose_bytes;
		params.smem_stage_offset = (requests_per_block * uint32_t(sizeof(ReqState)) + 15) & ~15u;
		params.smem_out_offset = params.smem_stage_offset + requests_per_block * 2 * stage_bytes;
		params.smem_bytes = params.smem_out_offset + requests_per_block * pose_bytes;
		params.magic_tracks = division_magic(max_tracks);
		params.magic_rot = division_magic(params.max_animated[0]);
		params.magic_vec = division_magic(params.max_animated[1] + params.max_animated[2]);
		params.magic_chunks = division_magic(pose_bytes >> 4);
		params.out_vector16 = ((uint64_t(reinterpret_cast<uintptr_t>(params.out)) | params.pose_stride) & 15) == 0 ? 1u : 0u;
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
