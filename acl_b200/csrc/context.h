// acl_b200/csrc/context.h -- host-side objects behind the opaque handles of include/aclb200.h.
#pragma once

#include <cuda_runtime.h>

#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/aclb200.h"
#include "layout.h"

struct aclb200_context
{
	int device = 0;
	int num_sms = 0;
	int max_dynamic_smem = 0;
	std::string last_error;
	uint64_t launch_count = 0;

	// scratch of the host-buffer convenience call
	void* d_scratch_requests = nullptr;
	size_t scratch_requests_bytes = 0;
	void* d_scratch_out = nullptr;
	size_t scratch_out_bytes = 0;
	cudaStream_t host_stream = nullptr;
	cudaStream_t copy_stream = nullptr;			// device -> host copies of the host-buffer call overlap the next chunk's decode
	cudaEvent_t chunk_done[2] = { nullptr, nullptr };

	// aclb200_calculate_compression_error (error_metric.cu): job table, arg max accumulators, requests and the decoded poses of one chunk
	void* d_error_scratch = nullptr;
	size_t error_scratch_bytes = 0;
	uint64_t error_chunk_bytes = 1024ull << 20;	// decoded poses per chunk (aclb200_set_error_chunk_bytes)

	// aclb200_debug_set_trace
	unsigned long long* d_trace = nullptr;
	uint32_t trace_blocks = 0;
	uint32_t trace_iterations = 0;
};

namespace aclb200
{
	// What a clip's base pose row (constant + default sub-tracks in the output layout, pipeline.cu) depends on besides the clip
	struct BasePoseKey
	{
		uint32_t layout;
		uint32_t normalize_always;
		uint32_t default_mode[3];
		float    constant_defaults[12];
	};

	struct BasePoseRows
	{
		BasePoseKey key;
		uint8_t* d_rows = nullptr;			// [num_clips][row_stride]
		uint32_t row_stride = 0;
		cudaEvent_t ready = nullptr;		// recorded after the build kernel on the stream that asked for it first
		cudaEvent_t last_launch = nullptr;	// recorded after the latest launch that reads the rows
		uint32_t users = 0;					// launches being set up with these rows (between acquire and release)
		uint64_t last_use = 0;
	};
}

struct aclb200_clipset
{
	int device = 0;
	aclb200_clipset_info info = {};
	std::vector<aclb200::ClipDesc> host_clips;		// host mirror for the info queries
	std::vector<uint32_t> host_looping;				// compressed_tracks::get_looping_policy() per clip
	uint32_t max_animated[3] = { 0, 0, 0 };			// largest number of animated rotation / translation / scale sub-tracks of a clip
	uint32_t max_animated_total = 0;
	uint32_t max_key_frame_bytes = 0;				// largest ceil(animated_pose_bit_size / 8) of any segment
	bool all_tracks_even = true;					// every clip has an even number of tracks (40 byte bones then give 16 byte granular rows)

	uint8_t* d_data = nullptr;
	aclb200::ClipDesc* d_clips = nullptr;

	// base pose rows built on first use per (layout, normalisation, default modes, default values), a handful kept
	mutable std::mutex base_mutex;
	mutable std::vector<aclb200::BasePoseRows> base_rows;
	mutable uint64_t base_clock = 0;
};

namespace aclb200
{
	aclb200_status set_error(aclb200_context* context, aclb200_status status, const std::string& message);
	aclb200_status check_cuda(aclb200_context* context, cudaError_t error, const char* what);

	// One launch description shared by every kernel of the transform / scalar paths.
	struct DecodeParams
	{
		const uint8_t* data;
		const ClipDesc* clips;
		const aclb200_request* requests;
		const uint32_t* track_indices;		// decompress_track only
		uint32_t num_requests;
		uint32_t num_clips;
		uint32_t max_tracks;
		uint32_t max_animated[3];
		uint32_t requests_per_block;		// whole requests handled by one thread block
		uint32_t magic_tracks;				// floor(2^32 / d) + 1 for d = max_tracks / max_animated[0] / max(max_animated[1] + [2]):
		uint32_t magic_rot;					// turns slot / d into a mulhi (0 when d == 1)
		uint32_t magic_vec;
		uint32_t magic_chunks;
		uint32_t stage_bytes;				// shared memory bytes reserved per staged key frame (0 = read the streams from global memory)
		uint32_t smem_pose_bytes;			// shared memory bytes reserved per assembled pose (0 = phases store straight to global memory)
		uint32_t smem_stage_offset;			// carve-up of the dynamic shared memory: ReqState[] | key frame windows | poses
		uint32_t smem_out_offset;
		uint32_t smem_bytes;				// dynamic shared memory of the launch
		uint32_t out_vector16;				// poses can leave shared memory with 16 byte stores
		uint32_t out_bulk;					// pipeline: every pose row is 16 byte granular, rows leave shared memory as TMA bulk stores
		uint32_t grid_blocks;				// pipeline: persistent grid size
		uint32_t smem_stage_size;			// pipeline: bytes of one stage (key frame windows + poses)
		uint32_t contiguous_batches;		// pipeline: every block takes one contiguous range of batches (else batches strided by the grid)
		uint32_t hot_slot_bytes;			// pipeline: bytes of one slot of the ReqHot ring (records + group words)
		uint32_t smem_tag_offset;			// pipeline: base row tags (which clip's base pose each pose row holds)
		unsigned long long* trace;			// pipeline, ACLB200_PIPE_TRACE builds: clock stamps per (block, iteration), see aclb200_debug_set_trace
		uint32_t trace_blocks;
		uint32_t trace_iterations;
		float    one;						// 1.0f the compiler cannot see (keeps f32x2 mul + add unfused, see pipeline.cu)
		const uint8_t* base_poses;			// pipeline: base pose row per clip (nullptr: phase A runs in the kernel)
		uint32_t base_stride;
		uint8_t* out;
		uint64_t pose_stride;
		uint32_t bone_stride;				// 48 or 40 (transform), components * 4 (scalar)

		uint32_t rounding_policy;
		uint32_t looping_policy;
		uint32_t normalization;
		uint32_t per_track_rounding;
		uint32_t wrapping;
		uint32_t clamp_sample_time;
		uint32_t multiple_rotation_formats;
		uint32_t default_mode[3];
		float    constant_defaults[12];
		const float* variable_defaults;
		const uint8_t* per_track_policies;
		uint32_t skip_all;					// ACLB200_SKIP_* bits skipped for every track
		const uint8_t* skip_tracks;			// [max_tracks] ACLB200_SKIP_* bits per track, or nullptr
		const uint8_t* request_policies;	// [num_requests][2] { rounding, looping } per request, or nullptr
		uint32_t layout;
		uint32_t debug_which;
		uint32_t debug_max_sub_tracks;
	};

	// kernels.cu
	void plan_launch(DecodeParams& params, uint32_t max_key_frame_bytes, int max_dynamic_smem, bool allow_output_staging);
	void plan_scalar_launch(DecodeParams& params, uint32_t max_key_frame_bytes);
	cudaError_t launch_transform_decompress_tracks(const DecodeParams& params, uint32_t math_mode, cudaStream_t stream);
	cudaError_t launch_transform_decompress_track(const DecodeParams& params, uint32_t math_mode, cudaStream_t stream);
	cudaError_t launch_transform_debug_seek(const DecodeParams& params, aclb200_seek_state* d_out, cudaStream_t stream);
	cudaError_t launch_transform_debug_unpack(const DecodeParams& params, uint32_t* d_out, cudaStream_t stream);
	cudaError_t launch_scalar_decompress_tracks(const DecodeParams& params, cudaStream_t stream);
	cudaError_t launch_scalar_decompress_track(const DecodeParams& params, cudaStream_t stream);
	cudaError_t configure_kernels(int& max_dynamic_smem);
	// error_metric.cu
	cudaError_t configure_error_kernels(int optin_limit);
	// pipeline.cu
	cudaError_t configure_pipeline_kernels(int optin_limit, int& min_available);
	bool plan_pipeline(DecodeParams& params, uint32_t max_key_frame_bytes, int max_dynamic_smem, int num_sms);
	cudaError_t launch_transform_pipeline(const DecodeParams& params, uint32_t math_mode, cudaStream_t stream);
	void acquire_base_poses(const aclb200_clipset* clipset, DecodeParams& params, cudaStream_t stream);
	void release_base_poses_use(const aclb200_clipset* clipset, const DecodeParams& params, cudaStream_t stream);
	void release_base_poses(aclb200_clipset* clipset);
}
