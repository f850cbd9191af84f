// acl_b200/csrc/api.cpp -- the extern "C" surface declared in include/aclb200.h.
#include "context.h"

#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>

namespace aclb200
{
	aclb200_status build_clipset(aclb200_context* context, const std::function<const uint8_t*(uint32_t)>& get_blob, const uint32_t* sizes,
		uint32_t num_clips, bool check_hash, aclb200_clipset** out_clipset, uint32_t* out_failed_clip);

	namespace
	{
		uint32_t scalar_components(uint32_t track_type)
		{
			return track_type <= 3 ? track_type + 1 : 4;
		}

		aclb200_status make_params(aclb200_context* context, const aclb200_clipset* clipset, const aclb200_request* d_requests,
			uint32_t num_requests, const aclb200_options* options, void* d_out, bool want_transform, bool single_track, DecodeParams& params)
		{
			if (context == nullptr || clipset == nullptr || options == nullptr)
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "null context / clipset / options");
			if (options->struct_size != sizeof(aclb200_options))
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "options.struct_size does not match this library, call aclb200_default_options()");
			if (num_requests != 0 && (d_requests == nullptr || d_out == nullptr))
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "null request / output pointer");
			if (clipset->device != context->device)
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "clip set lives on another device");
			const bool is_transform = clipset->info.track_type == ACLB200_TRACK_QVVF;
			if (is_transform != want_transform)
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, want_transform ? "not a transform clip set" : "not a scalar clip set");
			if (options->rounding_policy > ACLB200_ROUND_PER_TRACK || options->looping_policy > ACLB200_LOOP_AS_COMPRESSED
				|| options->normalization > ACLB200_NORMALIZE_ALWAYS || options->output_layout > ACLB200_LAYOUT_QVV40
				|| options->default_rotation_mode > ACLB200_DEFAULT_VARIABLE || options->default_translation_mode > ACLB200_DEFAULT_VARIABLE
				|| options->default_scale_mode > ACLB200_DEFAULT_LEGACY)
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "option value out of range");
			// ACL_ASSERT(rounding_policy != per_track || is_per_track_rounding_supported()), decompress.impl.h:211
			if (options->rounding_policy == ACLB200_ROUND_PER_TRACK && !options->per_track_rounding)
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "per track rounding must be enabled to seek with the per_track policy");

			// the per track rounding kernels read one batch wide seek policy for the tracks that do not override it
			if (options->d_request_policies != nullptr && options->per_track_rounding)
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "per request policies (d_request_policies) need per_track_rounding == 0");

			std::memset(&params, 0, sizeof(params));
			params.data = clipset->d_data;
			params.clips = clipset->d_clips;
			params.requests = d_requests;
			params.num_requests = num_requests;
			params.num_clips = clipset->info.num_clips;
			params.max_tracks = clipset->info.max_tracks;
			for (int k = 0; k < 3; ++k)
				params.max_animated[k] = clipset->max_animated[k];
			params.out = static_cast<uint8_t*>(d_out);
			if (is_transform)
				params.bone_stride = options->output_layout == ACLB200_LAYOUT_QVV48 ? 48u : 40u;
			else
				params.bone_stride = scalar_components(clipset->info.track_type) * 4u;
			params.pose_stride = options->pose_stride_bytes != 0 ? options->pose_stride_bytes : uint64_t(params.max_tracks) * params.bone_stride;
			if (!single_track && params.pose_stride < uint64_t(params.max_tracks) * params.bone_stride)
				return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "pose_stride_bytes is smaller than one pose");
			if (is_transform)
			{
				const uint64_t alignment = options->output_layout == ACLB200_LAYOUT_QVV48 ? 16 : 8;
				if ((params.pose_stride % alignment) != 0 || (reinterpret_cast<uintptr_t>(d_out) % alignment) != 0)
					return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "the output pointer and pose_stride_bytes must keep bones 16 (QVV48) / 8 (QVV40) byte aligned");
			}
			params.rounding_policy = options->rounding_policy;
			params.looping_policy = options->looping_policy;
			params.normalization = options->normalization;
			params.per_track_rounding = options->per_track_rounding != 0;
			params.wrapping = options->wrapping != 0;
			params.clamp_sample_time = options->clamp_sample_time != 0;
			params.multiple_rotation_formats = options->multiple_rotation_formats != 0;
			params.default_mode[0] = options->default_rotation_mode;
			params.default_mode[1] = options->default_translation_mode;
			params.default_mode[2] = options->default_scale_mode;
			std::memcpy(params.constant_defaults, options->constant_defaults, sizeof(params.constant_defaults));
			params.variable_defaults = options->d_variable_defaults;
			params.per_track_policies = options->d_per_track_rounding;
			params.skip_all = is_transform ? (options->skip_mask & 7u) : 0u;
			params.skip_tracks = is_transform ? options->d_skip_track_mask : nullptr;
			params.request_policies = options->d_request_policies;
			params.layout = options->output_layout;
			// `skipped` default sub-tracks must keep what the caller's buffer holds: those launches store sub-tracks straight to
			// global memory instead of assembling whole poses in shared memory
			const bool any_skipped = options->default_rotation_mode == ACLB200_DEFAULT_SKIPPED || options->default_translation_mode == ACLB200_DEFAULT_SKIPPED
				|| options->default_scale_mode == ACLB200_DEFAULT_SKIPPED;
			// ... and so must skipped sub-tracks (track_writer::skip_*)
			const bool any_masked = (options->skip_mask & 7u) != 0 || options->d_skip_track_mask != nullptr;
			const bool tracks_launch = is_transform && !single_track;
			plan_launch(params, tracks_launch ? clipset->max_key_frame_bytes : 0u, context->max_dynamic_smem, tracks_launch && !any_skipped && !any_masked);
			return ACLB200_OK;
		}

		aclb200_status finish_launch(aclb200_context* context, cudaError_t error, const char* what)
		{
			if (error == cudaSuccess)
				context->launch_count++;
			return check_cuda(context, error, what);
		}
	}
}

using namespace aclb200;

extern "C"
{
	const char* aclb200_version_string(void)
	{
		return "aclb200 0.3 (sm_100a; ACL compressed_tracks v02_00_00..v02_01_00)";
	}

	const char* aclb200_status_string(aclb200_status status)
	{
		switch (status)
		{
		case ACLB200_OK: return "ok";
		case ACLB200_ERR_INVALID_ARGUMENT: return "invalid argument";
		case ACLB200_ERR_INVALID_CLIP: return "invalid compressed_tracks buffer";
		case ACLB200_ERR_UNSUPPORTED: return "unsupported clip";
		case ACLB200_ERR_NO_DEVICE: return "no CUDA device";
		case ACLB200_ERR_CUDA: return "CUDA error";
		case ACLB200_ERR_OUT_OF_MEMORY: return "out of device memory";
		default: return "unknown status";
		}
	}

	void aclb200_default_options(aclb200_options* options)
	{
		if (options == nullptr)
			return;
		std::memset(options, 0, sizeof(*options));
		options->struct_size = sizeof(aclb200_options);
		options->rounding_policy = ACLB200_ROUND_NONE;
		options->looping_policy = ACLB200_LOOP_AS_COMPRESSED;
		options->normalization = ACLB200_NORMALIZE_LERP_ONLY;		// default_transform_decompression_settings, decompression_settings.h:226
		options->per_track_rounding = 0;							// :231
		options->wrapping = 1;										// :153
		options->clamp_sample_time = 1;								// :80
		options->multiple_rotation_formats = 0;						// :219
		options->default_rotation_mode = ACLB200_DEFAULT_CONSTANT;	// track_writer.h:170-172
		options->default_translation_mode = ACLB200_DEFAULT_CONSTANT;
		options->default_scale_mode = ACLB200_DEFAULT_LEGACY;
		options->constant_defaults[3] = 1.0f;						// :174-176
		options->constant_defaults[8] = options->constant_defaults[9] = options->constant_defaults[10] = 1.0f;
		options->output_layout = ACLB200_LAYOUT_QVV48;
		options->math_mode = ACLB200_MATH_EXACT;
	}

	aclb200_status aclb200_create(int device, aclb200_context** out_context)
	{
		if (out_context == nullptr)
			return ACLB200_ERR_INVALID_ARGUMENT;
		*out_context = nullptr;
		int device_count = 0;
		if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0)
			return ACLB200_ERR_NO_DEVICE;		// no CPU fallback exists
		if (device < 0 || device >= device_count)
			return ACLB200_ERR_INVALID_ARGUMENT;
		cudaDeviceProp prop;
		if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
			return ACLB200_ERR_CUDA;

		aclb200_context* context = new (std::nothrow) aclb200_context();
		if (context == nullptr)
			return ACLB200_ERR_OUT_OF_MEMORY;
		context->device = device;
		context->num_sms = prop.multiProcessorCount;
		context->max_dynamic_smem = int(prop.sharedMemPerBlockOptin);
		if (prop.major < 10 || cudaSetDevice(device) != cudaSuccess || configure_kernels(context->max_dynamic_smem) != cudaSuccess)
		{
			// the kernels are compiled for sm_100a only
			delete context;
			return ACLB200_ERR_NO_DEVICE;
		}
		*out_context = context;
		return ACLB200_OK;
	}

	void aclb200_destroy(aclb200_context* context)
	{
		if (context == nullptr)
			return;
		cudaSetDevice(context->device);
		cudaFree(context->d_scratch_requests);
		cudaFree(context->d_scratch_out);
		cudaFree(context->d_error_scratch);
		if (context->host_stream != nullptr)
			cudaStreamDestroy(context->host_stream);
		if (context->copy_stream != nullptr)
			cudaStreamDestroy(context->copy_stream);
		for (cudaEvent_t event : context->chunk_done)
			if (event != nullptr)
				cudaEventDestroy(event);
		delete context;
	}

	const char* aclb200_last_error(const aclb200_context* context)
	{
		return context != nullptr ? context->last_error.c_str() : "null context";
	}

	aclb200_status aclb200_upload_clips(aclb200_context* context, const void* const* blobs, const uint32_t* sizes, uint32_t num_clips,
		uint32_t check_hash, aclb200_clipset** out_clipset, uint32_t* out_failed_clip)
	{
		if (blobs == nullptr)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "upload_clips: null blob array");
		return build_clipset(context, [blobs](uint32_t clip) { return static_cast<const uint8_t*>(blobs[clip]); }, sizes, num_clips,
			check_hash != 0, out_clipset, out_failed_clip);
	}

	aclb200_status aclb200_upload_clips_packed(aclb200_context* context, const void* buffer, const uint64_t* offsets, const uint32_t* sizes,
		uint32_t num_clips, uint32_t check_hash, aclb200_clipset** out_clipset, uint32_t* out_failed_clip)
	{
		if (buffer == nullptr || offsets == nullptr)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "upload_clips_packed: null buffer / offsets");
		const uint8_t* base = static_cast<const uint8_t*>(buffer);
		return build_clipset(context, [base, offsets](uint32_t clip) { return base + offsets[clip]; }, sizes, num_clips,
			check_hash != 0, out_clipset, out_failed_clip);
	}

	void aclb200_release_clipset(aclb200_context* context, aclb200_clipset* clipset)
	{
		if (clipset == nullptr)
			return;
		cudaSetDevice(clipset->device);
		release_base_poses(clipset);
		cudaFree(clipset->d_data);
		cudaFree(clipset->d_clips);
		delete clipset;
		(void)context;
	}

	aclb200_status aclb200_clipset_get_info(const aclb200_clipset* clipset, aclb200_clipset_info* out_info)
	{
		if (clipset == nullptr || out_info == nullptr)
			return ACLB200_ERR_INVALID_ARGUMENT;
		*out_info = clipset->info;
		return ACLB200_OK;
	}

	aclb200_status aclb200_clipset_get_clip_info(const aclb200_clipset* clipset, uint32_t clip, aclb200_clip_info* out_info)
	{
		if (clipset == nullptr || out_info == nullptr || clip >= clipset->info.num_clips)
			return ACLB200_ERR_INVALID_ARGUMENT;
		const ClipDesc& desc = clipset->host_clips[clip];
		out_info->num_tracks = desc.num_tracks;
		out_info->num_samples = desc.num_samples;
		out_info->sample_rate = desc.sample_rate;
		out_info->looping_policy = clipset->host_looping[clip];
		out_info->duration = out_info->looping_policy == ACLB200_LOOP_WRAP ? desc.duration_wrap : desc.duration_clamp;
		out_info->num_segments = desc.num_segments;
		out_info->hash = desc.hash;
		out_info->size = desc.size;
		return ACLB200_OK;
	}

	aclb200_status aclb200_decompress_tracks(aclb200_context* context, const aclb200_clipset* clipset,
		const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options* options, void* d_out, void* stream)
	{
		DecodeParams params;
		const aclb200_status status = make_params(context, clipset, d_requests, num_requests, options, d_out, true, false, params);
		if (status != ACLB200_OK || num_requests == 0)
			return status;
		cudaSetDevice(context->device);

		// Main path: the persistent TMA pipeline (pipeline.cu). It assembles whole poses in shared memory, so launches that must
		// leave `skipped` default sub-tracks untouched, or whose poses do not fit in shared memory, use the plain kernels instead.
		const bool any_skipped = options->default_rotation_mode == ACLB200_DEFAULT_SKIPPED || options->default_translation_mode == ACLB200_DEFAULT_SKIPPED
			|| options->default_scale_mode == ACLB200_DEFAULT_SKIPPED;
		const bool any_masked = (options->skip_mask & 7u) != 0 || options->d_skip_track_mask != nullptr;
		// ACLB200_PIPELINE=0 sends everything through the plain kernels (tuning / A-B measurements)
		static const char* const override_pipeline = std::getenv("ACLB200_PIPELINE");
		const bool pipeline_allowed = override_pipeline == nullptr || override_pipeline[0] != '0';
		if (pipeline_allowed && !any_skipped && !any_masked && clipset->max_key_frame_bytes != 0)
		{
			DecodeParams pipeline_params = params;
			if (plan_pipeline(pipeline_params, clipset->max_key_frame_bytes, context->max_dynamic_smem, context->num_sms))
			{
				const bool rows_16 = options->output_layout == ACLB200_LAYOUT_QVV48 || clipset->all_tracks_even;
				pipeline_params.out_bulk = rows_16 && ((uint64_t(reinterpret_cast<uintptr_t>(d_out)) | pipeline_params.pose_stride) & 15) == 0 ? 1u : 0u;
				pipeline_params.trace = context->d_trace;
				pipeline_params.trace_blocks = context->trace_blocks;
				pipeline_params.trace_iterations = context->trace_iterations;
				acquire_base_poses(clipset, pipeline_params, static_cast<cudaStream_t>(stream));
				const cudaError_t launched = launch_transform_pipeline(pipeline_params, options->math_mode, static_cast<cudaStream_t>(stream));
				release_base_poses_use(clipset, pipeline_params, static_cast<cudaStream_t>(stream));
				return finish_launch(context, launched, "decompress_tracks (pipeline)");
			}
		}
		return finish_launch(context, launch_transform_decompress_tracks(params, options->math_mode, static_cast<cudaStream_t>(stream)), "decompress_tracks");
	}

	aclb200_status aclb200_decompress_track(aclb200_context* context, const aclb200_clipset* clipset,
		const aclb200_request* d_requests, const uint32_t* d_track_indices, uint32_t num_requests,
		const aclb200_options* options, void* d_out, void* stream)
	{
		DecodeParams params;
		const aclb200_status status = make_params(context, clipset, d_requests, num_requests, options, d_out, true, true, params);
		if (status != ACLB200_OK || num_requests == 0)
			return status;
		if (d_track_indices == nullptr)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "null track index pointer");
		params.track_indices = d_track_indices;
		cudaSetDevice(context->device);
		return finish_launch(context, launch_transform_decompress_track(params, options->math_mode, static_cast<cudaStream_t>(stream)), "decompress_track");
	}

	aclb200_status aclb200_scalar_decompress_tracks(aclb200_context* context, const aclb200_clipset* clipset,
		const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options* options, void* d_out, void* stream)
	{
		DecodeParams params;
		const aclb200_status status = make_params(context, clipset, d_requests, num_requests, options, d_out, false, false, params);
		if (status != ACLB200_OK || num_requests == 0)
			return status;
		cudaSetDevice(context->device);
		plan_scalar_launch(params, clipset->max_key_frame_bytes);
		return finish_launch(context, launch_scalar_decompress_tracks(params, static_cast<cudaStream_t>(stream)), "scalar_decompress_tracks");
	}

	aclb200_status aclb200_scalar_decompress_track(aclb200_context* context, const aclb200_clipset* clipset,
		const aclb200_request* d_requests, const uint32_t* d_track_indices, uint32_t num_requests,
		const aclb200_options* options, void* d_out, void* stream)
	{
		DecodeParams params;
		const aclb200_status status = make_params(context, clipset, d_requests, num_requests, options, d_out, false, true, params);
		if (status != ACLB200_OK || num_requests == 0)
			return status;
		if (d_track_indices == nullptr)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "null track index pointer");
		params.track_indices = d_track_indices;
		cudaSetDevice(context->device);
		return finish_launch(context, launch_scalar_decompress_track(params, static_cast<cudaStream_t>(stream)), "scalar_decompress_track");
	}

	aclb200_status aclb200_decompress_tracks_host(aclb200_context* context, const aclb200_clipset* clipset,
		const aclb200_request* requests, uint32_t num_requests, const aclb200_options* options, void* out, size_t out_bytes)
	{
		if (context == nullptr || clipset == nullptr || options == nullptr)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "null context / clipset / options");
		if (num_requests == 0)
			return ACLB200_OK;
		if (requests == nullptr || out == nullptr)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "null request / output pointer");

		const bool is_transform = clipset->info.track_type == ACLB200_TRACK_QVVF;
		const uint32_t bone_stride = is_transform ? (options->output_layout == ACLB200_LAYOUT_QVV48 ? 48u : 40u) : scalar_components(clipset->info.track_type) * 4u;
		const uint64_t pose_stride = options->pose_stride_bytes != 0 ? options->pose_stride_bytes : uint64_t(clipset->info.max_tracks) * bone_stride;
		const size_t needed_out = size_t(pose_stride) * num_requests;
		if (out_bytes < needed_out)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "output buffer too small");
		const size_t needed_requests = sizeof(aclb200_request) * size_t(num_requests);

		cudaError_t error = cudaSetDevice(context->device);
		if (error == cudaSuccess && context->host_stream == nullptr)
			error = cudaStreamCreateWithFlags(&context->host_stream, cudaStreamNonBlocking);
		if (error == cudaSuccess && context->scratch_requests_bytes < needed_requests)
		{
			cudaFree(context->d_scratch_requests);
			context->d_scratch_requests = nullptr;
			context->scratch_requests_bytes = 0;
			error = cudaMalloc(&context->d_scratch_requests, needed_requests);
			if (error == cudaSuccess) context->scratch_requests_bytes = needed_requests;
		}
		if (error == cudaSuccess && context->scratch_out_bytes < needed_out)
		{
			cudaFree(context->d_scratch_out);
			context->d_scratch_out = nullptr;
			context->scratch_out_bytes = 0;
			error = cudaMalloc(&context->d_scratch_out, needed_out);
			if (error == cudaSuccess) context->scratch_out_bytes = needed_out;
		}
		if (error != cudaSuccess)
			return check_cuda(context, error, "decompress_tracks_host: scratch allocation");

		cudaStream_t stream = context->host_stream;
		if (context->copy_stream == nullptr)
			error = cudaStreamCreateWithFlags(&context->copy_stream, cudaStreamNonBlocking);
		for (int i = 0; i < 2 && error == cudaSuccess; ++i)
			if (context->chunk_done[i] == nullptr)
				error = cudaEventCreateWithFlags(&context->chunk_done[i], cudaEventDisableTiming);
		if (error != cudaSuccess)
			return check_cuda(context, error, "decompress_tracks_host: streams");

		// `skipped` default sub-tracks keep what the caller's buffer held: bring the buffer in first in that case
		const bool keeps_input = is_transform && (options->default_rotation_mode == ACLB200_DEFAULT_SKIPPED
			|| options->default_translation_mode == ACLB200_DEFAULT_SKIPPED || options->default_scale_mode == ACLB200_DEFAULT_SKIPPED
			|| (options->skip_mask & 7u) != 0 || options->d_skip_track_mask != nullptr);
		// Rows no request writes (clips shorter than the widest one, requests naming a clip outside the set) read as zero. Clearing the
		// scratch costs a pass over it, so it only happens when such rows can exist.
		bool needs_clear = !keeps_input && (clipset->info.min_tracks != clipset->info.max_tracks || pose_stride != uint64_t(clipset->info.max_tracks) * bone_stride);
		if (!keeps_input && !needs_clear)
			for (uint32_t r = 0; r < num_requests && !needs_clear; ++r)
				needs_clear = requests[r].clip >= clipset->info.num_clips;

		error = cudaMemcpyAsync(context->d_scratch_requests, requests, needed_requests, cudaMemcpyHostToDevice, stream);
		if (error == cudaSuccess && keeps_input)
			error = cudaMemcpyAsync(context->d_scratch_out, out, needed_out, cudaMemcpyHostToDevice, stream);
		else if (error == cudaSuccess && needs_clear)
			error = cudaMemsetAsync(context->d_scratch_out, 0, needed_out, stream);
		if (error != cudaSuccess)
			return check_cuda(context, error, "decompress_tracks_host: upload");

		// Decode in a few chunks on one stream while the previous chunk's poses cross PCIe on another: the copy is the long pole
		// (tens of milliseconds per GB against well under a millisecond of decode), so it starts as early as possible and never waits
		// for the whole batch.
		const uint32_t num_chunks = num_requests >= 65536 ? 8u : (num_requests >= 4096 ? 2u : 1u);
		const aclb200_request* d_requests = static_cast<const aclb200_request*>(context->d_scratch_requests);
		uint8_t* d_out = static_cast<uint8_t*>(context->d_scratch_out);
		for (uint32_t chunk = 0; chunk < num_chunks; ++chunk)
		{
			const uint32_t first = uint32_t(uint64_t(num_requests) * chunk / num_chunks);
			const uint32_t last = uint32_t(uint64_t(num_requests) * (chunk + 1) / num_chunks);
			if (first == last)
				continue;
			uint8_t* d_chunk = d_out + size_t(first) * pose_stride;
			const aclb200_status status = is_transform
				? aclb200_decompress_tracks(context, clipset, d_requests + first, last - first, options, d_chunk, stream)
				: aclb200_scalar_decompress_tracks(context, clipset, d_requests + first, last - first, options, d_chunk, stream);
			if (status != ACLB200_OK)
			{
				cudaStreamSynchronize(stream);
				cudaStreamSynchronize(context->copy_stream);
				return status;
			}
			cudaEvent_t done = context->chunk_done[chunk & 1];
			error = cudaEventRecord(done, stream);
			if (error == cudaSuccess) error = cudaStreamWaitEvent(context->copy_stream, done, 0);
			if (error == cudaSuccess)
				error = cudaMemcpyAsync(static_cast<uint8_t*>(out) + size_t(first) * pose_stride, d_chunk, size_t(last - first) * pose_stride, cudaMemcpyDeviceToHost, context->copy_stream);
			if (error != cudaSuccess)
				break;
		}
		const cudaError_t sync_decode = cudaStreamSynchronize(stream);
		const cudaError_t sync_copy = cudaStreamSynchronize(context->copy_stream);
		if (error == cudaSuccess) error = sync_decode;
		if (error == cudaSuccess) error = sync_copy;
		return check_cuda(context, error, "decompress_tracks_host: download");
	}

	aclb200_status aclb200_debug_seek(aclb200_context* context, const aclb200_clipset* clipset,
		const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options* options, aclb200_seek_state* d_out, void* stream)
	{
		DecodeParams params;
		const aclb200_status status = make_params(context, clipset, d_requests, num_requests, options, d_out, true, true, params);
		if (status != ACLB200_OK || num_requests == 0)
			return status;
		cudaSetDevice(context->device);
		return finish_launch(context, launch_transform_debug_seek(params, d_out, static_cast<cudaStream_t>(stream)), "debug_seek");
	}

	aclb200_status aclb200_debug_unpack(aclb200_context* context, const aclb200_clipset* clipset,
		const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options* options, uint32_t which,
		uint32_t max_animated_sub_tracks, uint32_t* d_out, void* stream)
	{
		DecodeParams params;
		const aclb200_status status = make_params(context, clipset, d_requests, num_requests, options, d_out, true, true, params);
		if (status != ACLB200_OK || num_requests == 0)
			return status;
		if (which > 1)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "key frame selector must be 0 or 1");
		params.debug_which = which;
		params.debug_max_sub_tracks = max_animated_sub_tracks;
		cudaSetDevice(context->device);
		return finish_launch(context, launch_transform_debug_unpack(params, d_out, static_cast<cudaStream_t>(stream)), "debug_unpack");
	}

	aclb200_status aclb200_debug_set_trace(aclb200_context* context, void* d_trace, uint32_t num_blocks, uint32_t num_iterations)
	{
		if (context == nullptr)
			return ACLB200_ERR_INVALID_ARGUMENT;
		context->d_trace = static_cast<unsigned long long*>(d_trace);
		context->trace_blocks = d_trace != nullptr ? num_blocks : 0;
		context->trace_iterations = d_trace != nullptr ? num_iterations : 0;
		return ACLB200_OK;
	}

	aclb200_status aclb200_device_malloc(aclb200_context* context, size_t bytes, void** out_device_pointer)
	{
		if (context == nullptr || out_device_pointer == nullptr)
			return ACLB200_ERR_INVALID_ARGUMENT;
		*out_device_pointer = nullptr;
		cudaSetDevice(context->device);
		return check_cuda(context, cudaMalloc(out_device_pointer, bytes == 0 ? 1 : bytes), "device_malloc");
	}

	void aclb200_device_free(aclb200_context* context, void* device_pointer)
	{
		if (context == nullptr || device_pointer == nullptr)
			return;
		cudaSetDevice(context->device);
		cudaFree(device_pointer);
	}

	aclb200_status aclb200_copy_to_device(aclb200_context* context, void* device_destination, const void* host_source, size_t bytes)
	{
		if (context == nullptr || (bytes != 0 && (device_destination == nullptr || host_source == nullptr)))
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "copy_to_device: null pointer");
		cudaSetDevice(context->device);
		return check_cuda(context, cudaMemcpy(device_destination, host_source, bytes, cudaMemcpyHostToDevice), "copy_to_device");
	}

	aclb200_status aclb200_copy_to_host(aclb200_context* context, void* host_destination, const void* device_source, size_t bytes)
	{
		if (context == nullptr || (bytes != 0 && (host_destination == nullptr || device_source == nullptr)))
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "copy_to_host: null pointer");
		cudaSetDevice(context->device);
		return check_cuda(context, cudaMemcpy(host_destination, device_source, bytes, cudaMemcpyDeviceToHost), "copy_to_host");
	}

	uint64_t aclb200_launch_count(const aclb200_context* context)
	{
		return context != nullptr ? context->launch_count : 0;
	}
}
