// include/acl_b200/track_error.h -- the reference's compression error measurement with the reference's own signature,
// for a context of this library: acl_b200::calculate_compression_error(allocator, raw_tracks, context[, error_metric]) is what a
// call site of acl::calculate_compression_error (includes/acl/compression/track_error.h:64-121) changes its namespace to.
//
// Needs the reference's headers on the include path (acl::track_array, acl::track_error, acl::itransform_error_metric are its types).
// What runs where: the raw clip is sampled on the host by the reference's own track_array::sample_tracks (the raw data lives in host
// memory and is the caller's), every sample is DECODED, taken to object space and measured ON THE DEVICE by
// aclb200_calculate_compression_error (include/aclb200.h), one acl::track_error comes back. For many clips at once call the C entry
// point directly (one launch sequence for the lot).
//
// Differences from the reference, all reported (never silent):
//   * the reference's three metrics run on the device (qvvf_transform_error_metric, qvvf_matrix3x4f_transform_error_metric,
//     additive_qvvf_transform_error_metric<format> with the additive base overload); a user defined metric throws acl_b200::error
//     (ACLB200_ERR_UNSUPPORTED);
//   * rtm::quat_normalize starts from the CPU's rsqrtss estimate in the reference: errors agree within 5e-5 on poses tens of units
//     across (measured 1e-5), the worst track and its sample time are the reference's whenever its lead exceeds that.
#pragma once

#include "decompress.h"

#if !ACLB200_WITH_ACL_HEADERS
	#error "acl_b200/track_error.h works on the reference's types: put <acl>/includes and <rtm>/includes on the include path"
#endif

#include <acl/compression/track_array.h>
#include <acl/compression/track_error.h>
#include <acl/compression/transform_error_metrics.h>
#include <acl/core/impl/debug_track_writer.h>

namespace acl_b200
{
	namespace shim_impl
	{
		// additive_qvvf_transform_error_metric<format>::get_name() (transform_error_metrics.h:475-485) -> additive_clip_format8, 0 for the plain
		// qvvf_transform_error_metric, ~0 for a metric the device does not implement
		inline uint32_t additive_format_of(const acl::itransform_error_metric& error_metric)
		{
			const char* name = error_metric.get_name();
			if (std::strcmp(name, "qvvf_transform_error_metric") == 0 || std::strcmp(name, "additive_qvvf_transform_error_metric<none>") == 0) return 0;
			if (std::strcmp(name, "additive_qvvf_transform_error_metric<relative>") == 0) return 1;
			if (std::strcmp(name, "additive_qvvf_transform_error_metric<additive0>") == 0) return 2;
			if (std::strcmp(name, "additive_qvvf_transform_error_metric<additive1>") == 0) return 3;
			return ~0u;
		}

		inline bool is_matrix_metric(const acl::itransform_error_metric& error_metric)
		{
			return std::strcmp(error_metric.get_name(), "qvvf_matrix3x4f_transform_error_metric") == 0;
		}

		template<class decompression_context_type>
		inline acl::track_error measure_on_device(acl::iallocator& allocator, const acl::track_array& raw_tracks, decompression_context_type& context,
			const acl::track_array_qvvf* additive_base_tracks = nullptr, uint32_t additive_format = 0, uint32_t error_metric = ACLB200_METRIC_QVVF)
		{
			using settings_type = typename decompression_context_type::settings_type;
			const acl::compressed_tracks& tracks = *context.get_compressed_tracks();
			const bool is_transform = raw_tracks.get_track_type() == acl::track_type8::qvvf;
			const uint32_t num_tracks = raw_tracks.get_num_tracks();
			const uint32_t num_samples = raw_tracks.get_num_samples_per_track();
			if (num_samples == 0 || num_tracks == 0)
				return acl::track_error();		// Cannot measure any error (track_error.impl.h:170-176,229-235)

			batch_decompressor& batch = context.device_batch();
			device_context& device = batch.device();
			const aclb200_clipset_info& info = batch.info();
			const uint32_t components = is_transform ? 12u : (info.track_type <= 3 ? info.track_type + 1 : 4u);
			const uint32_t row_tracks = info.max_tracks > num_tracks ? info.max_tracks : num_tracks;
			const size_t pose_floats = size_t(row_tracks) * components;

			// We use the nearest sample to accurately measure the loss that happened, if any but only if all data is loaded (track_error.impl.h:552-559)
			const acl::sample_rounding_policy rounding = (tracks.has_database() || tracks.has_stripped_keyframes()) ? acl::sample_rounding_policy::none : acl::sample_rounding_policy::nearest;
			const float sample_rate = raw_tracks.get_sample_rate();
			const float duration = raw_tracks.get_finite_duration();

			// sample_tracks0 (track_error.impl.h:504-507 / :415-418): the raw clip, sampled by the reference's own code on the host
			std::vector<float> raw_poses(pose_floats * num_samples, 0.0F);
			{
				acl::acl_impl::debug_track_writer writer(allocator, raw_tracks.get_track_type(), num_tracks);
				for (uint32_t sample = 0; sample < num_samples; ++sample)
				{
					const float sample_time = rtm::scalar_min(float(sample) / sample_rate, duration);
					raw_tracks.sample_tracks(sample_time, rounding, writer);
					float* row = raw_poses.data() + pose_floats * sample;
					if (is_transform)
						std::memcpy(row, writer.tracks_typed.qvvf, sizeof(rtm::qvvf) * num_tracks);
					else
					{
						const uint32_t stride = components * sizeof(float);		// debug_track_writer stores scalar tracks as arrays of their own type
						std::memcpy(row, writer.tracks_typed.any, size_t(stride) * num_tracks);
					}
				}
			}

			aclb200_options options;
			std::memset(&options, 0, sizeof(options));
			aclb200_default_options(&options);
			options.normalization = static_cast<uint32_t>(settings_type::get_rotation_normalization_policy());
			options.per_track_rounding = settings_type::is_per_track_rounding_supported() ? 1u : 0u;
			options.wrapping = settings_type::is_wrapping_supported() ? 1u : 0u;
			options.clamp_sample_time = settings_type::clamp_sample_time() ? 1u : 0u;
			options.multiple_rotation_formats = supports_multiple_rotation_formats<settings_type>() ? 1u : 0u;
			options.looping_policy = static_cast<uint32_t>(context.get_looping_policy());
			options.pose_stride_bytes = pose_floats * sizeof(float);

			device_buffer d_raw, d_bind, d_parents, d_shells, d_outputs, d_result, d_base;		// released when the call returns
			std::vector<uint32_t> parents, outputs;
			std::vector<float> shells;
			if (is_transform)
			{
				// initialize_bind_pose (track_error.impl.h:497-501, debug_track_writer.h:75-101): default sub-tracks read the bind pose
				std::vector<float> bind_pose(size_t(info.max_tracks) * 12, 0.0F);
				parents.resize(num_tracks);
				outputs.resize(num_tracks);
				shells.resize(num_tracks);
				const acl::track_array_qvvf& transforms = acl::track_array_cast<acl::track_array_qvvf>(raw_tracks);
				for (uint32_t track_index = 0; track_index < num_tracks; ++track_index)
				{
					const acl::track_desc_transformf& desc = transforms[track_index].get_description();
					parents[track_index] = desc.parent_index;
					shells[track_index] = desc.shell_distance;
					outputs[track_index] = desc.output_index;
					if (desc.output_index != acl::k_invalid_track_index && desc.output_index < info.max_tracks)
						std::memcpy(bind_pose.data() + size_t(desc.output_index) * 12, &desc.default_value, sizeof(rtm::qvvf));
				}
				options.default_rotation_mode = options.default_translation_mode = options.default_scale_mode = ACLB200_DEFAULT_VARIABLE;
				options.d_variable_defaults = static_cast<const float*>(d_bind.get(device, bind_pose.size() * sizeof(float)));
				device.check(aclb200_copy_to_device(device.get(), d_bind.get(device, bind_pose.size() * sizeof(float)), bind_pose.data(), bind_pose.size() * sizeof(float)), "bind pose upload");
				device.check(aclb200_copy_to_device(device.get(), d_parents.get(device, num_tracks * 4), parents.data(), num_tracks * 4), "skeleton upload");
				device.check(aclb200_copy_to_device(device.get(), d_shells.get(device, num_tracks * 4), shells.data(), num_tracks * 4), "skeleton upload");
				device.check(aclb200_copy_to_device(device.get(), d_outputs.get(device, num_tracks * 4), outputs.data(), num_tracks * 4), "skeleton upload");
			}
			device.check(aclb200_copy_to_device(device.get(), d_raw.get(device, raw_poses.size() * sizeof(float)), raw_poses.data(), raw_poses.size() * sizeof(float)), "raw pose upload");

			// sample_tracks_base (track_error.impl.h:350-356,658-661): the additive base at the matching time of every sample
			const bool has_base = is_transform && additive_base_tracks != nullptr && !additive_base_tracks->is_empty() && additive_format != 0;
			if (has_base)
			{
				const uint32_t base_num_samples = additive_base_tracks->get_num_samples_per_track();
				const float base_duration = additive_base_tracks->get_finite_duration();
				std::vector<float> base_poses(pose_floats * num_samples, 0.0F);
				acl::acl_impl::debug_track_writer writer(allocator, acl::track_type8::qvvf, num_tracks);
				for (uint32_t sample = 0; sample < num_samples; ++sample)
				{
					const float sample_time = rtm::scalar_min(float(sample) / sample_rate, duration);
					const float normalized_sample_time = base_num_samples > 1 ? (sample_time / duration) : 0.0F;
					const float additive_sample_time = base_num_samples > 1 ? (normalized_sample_time * base_duration) : 0.0F;
					additive_base_tracks->sample_tracks(additive_sample_time, rounding, writer);
					std::memcpy(base_poses.data() + pose_floats * sample, writer.tracks_typed.qvvf, sizeof(rtm::qvvf) * num_tracks);
				}
				device.check(aclb200_copy_to_device(device.get(), d_base.get(device, base_poses.size() * sizeof(float)), base_poses.data(), base_poses.size() * sizeof(float)), "base pose upload");
			}

			aclb200_error_job job;
			std::memset(&job, 0, sizeof(job));
			job.clip = 0;
			job.num_samples = num_samples;
			job.sample_rate = sample_rate;
			job.duration = duration;
			job.num_tracks = num_tracks;
			job.additive_format = has_base ? additive_format : 0u;
			job.error_metric = is_transform ? error_metric : 0u;
			aclb200_track_error* d_out = static_cast<aclb200_track_error*>(d_result.get(device, sizeof(aclb200_track_error)));
			device.check(aclb200_calculate_compression_error(device.get(), batch.clipset(), &job, 1, d_raw.get(device, 0),
				is_transform ? static_cast<const uint32_t*>(d_parents.get(device, 0)) : nullptr, is_transform ? static_cast<const float*>(d_shells.get(device, 0)) : nullptr,
				is_transform ? static_cast<const uint32_t*>(d_outputs.get(device, 0)) : nullptr, has_base ? d_base.get(device, 0) : nullptr, &options, d_out, nullptr, nullptr), "aclb200_calculate_compression_error");
			aclb200_track_error result;
			device.check(aclb200_copy_to_host(device.get(), &result, d_out, sizeof(result)), "result download");
			if ((result.flags & ACLB200_ERROR_FLAG_INVALID_SKELETON) != 0)
				throw error(ACLB200_ERR_INVALID_ARGUMENT, "calculate_compression_error: a parent track does not precede its child");

			acl::track_error out;
			out.index = result.index;
			out.error = result.error;
			out.sample_time = result.sample_time;
			return out;
		}
	}

	// calculate_compression_error(allocator, raw_tracks, context), compression/track_error.h:64-75: scalar tracks only
	template<class settings_type>
	inline acl::track_error calculate_compression_error(acl::iallocator& allocator, const acl::track_array& raw_tracks, decompression_context<settings_type>& context)
	{
		ACL_ASSERT(raw_tracks.is_valid().empty(), "Raw tracks are invalid");
		ACL_ASSERT(context.is_initialized(), "Context isn't initialized");
		if (raw_tracks.get_track_type() == acl::track_type8::qvvf)
			return acl::acl_impl::invalid_track_error();	// Only supports scalar tracks (track_error.impl.h:408-409)
		return shim_impl::measure_on_device(allocator, raw_tracks, context);
	}

	// calculate_compression_error(allocator, raw_tracks, context, error_metric), compression/track_error.h:77-91: scalar and transform tracks
	template<class settings_type>
	inline acl::track_error calculate_compression_error(acl::iallocator& allocator, const acl::track_array& raw_tracks, decompression_context<settings_type>& context,
		const acl::itransform_error_metric& error_metric)
	{
		ACL_ASSERT(raw_tracks.is_valid().empty(), "Raw tracks are invalid");
		ACL_ASSERT(context.is_initialized(), "Context isn't initialized");
		if (shim_impl::is_matrix_metric(error_metric))
			return shim_impl::measure_on_device(allocator, raw_tracks, context, nullptr, 0, ACLB200_METRIC_QVVF_MATRIX3X4F);
		if (raw_tracks.get_track_type() == acl::track_type8::qvvf && shim_impl::additive_format_of(error_metric) == ~0u)
			throw error(ACLB200_ERR_UNSUPPORTED, std::string("calculate_compression_error: not one of the reference's error metrics: ") + error_metric.get_name());
		return shim_impl::measure_on_device(allocator, raw_tracks, context);
	}

	// calculate_compression_error(allocator, raw_tracks, context, error_metric, additive_base_tracks), compression/track_error.h:93-107:
	// the metric's additive format (additive_qvvf_transform_error_metric<format>) says how the base applies
	template<class settings_type>
	inline acl::track_error calculate_compression_error(acl::iallocator& allocator, const acl::track_array_qvvf& raw_tracks, decompression_context<settings_type>& context,
		const acl::itransform_error_metric& error_metric, const acl::track_array_qvvf& additive_base_tracks)
	{
		ACL_ASSERT(raw_tracks.is_valid().empty(), "Raw tracks are invalid");
		ACL_ASSERT(context.is_initialized(), "Context isn't initialized");
		const uint32_t additive_format = shim_impl::additive_format_of(error_metric);
		if (additive_format == ~0u)
			throw error(ACLB200_ERR_UNSUPPORTED, std::string("calculate_compression_error: only the qvvf_transform_error_metric family runs on the device, not ") + error_metric.get_name());
		return shim_impl::measure_on_device(allocator, raw_tracks, context, &additive_base_tracks, additive_format);
	}

	// calculate_compression_error(allocator, context0, context1), compression/track_error.h:109-121 (impl/track_error.impl.h:690-750): the worst
	// difference between two compressed clips, scalar tracks only. Composition of two device calls: every sample of context0's clip is decoded
	// into device memory (aclb200_decompress_all_samples) and plays the raw side of the measurement of context1's clip. Both contexts must
	// live on the same device.
	template<class settings_type0, class settings_type1>
	inline acl::track_error calculate_compression_error(acl::iallocator& allocator, decompression_context<settings_type0>& context0, decompression_context<settings_type1>& context1)
	{
		(void)allocator;
		ACL_ASSERT(context0.is_initialized(), "Context isn't initialized");
		ACL_ASSERT(context1.is_initialized(), "Context isn't initialized");
		const acl::compressed_tracks* tracks0 = context0.get_compressed_tracks();
		const acl::compressed_tracks* tracks1 = context1.get_compressed_tracks();
		if (tracks0->get_track_type() == acl::track_type8::qvvf)
			return acl::acl_impl::invalid_track_error();	// Only supports scalar tracks
		const uint32_t num_samples = tracks0->get_num_samples_per_track();
		const uint32_t num_tracks = tracks0->get_num_tracks();
		if (num_samples == 0 || num_tracks == 0)
			return acl::track_error();

		batch_decompressor& batch0 = context0.device_batch();
		batch_decompressor& batch1 = context1.device_batch();
		device_context& device = batch1.device();
		if (&batch0.device() != &device)
			throw error(ACLB200_ERR_INVALID_ARGUMENT, "calculate_compression_error(context0, context1): the contexts live on different devices");
		if (batch0.info().track_type != batch1.info().track_type || tracks1->get_num_tracks() != num_tracks)
			throw error(ACLB200_ERR_INVALID_ARGUMENT, "calculate_compression_error(context0, context1): the clips do not hold the same tracks");

		aclb200_options options;
		std::memset(&options, 0, sizeof(options));
		aclb200_default_options(&options);
		// nearest, unless either clip misses key frames (track_error.impl.h:741-745)
		const bool interpolate = tracks0->has_database() || tracks1->has_database() || tracks0->has_stripped_keyframes() || tracks1->has_stripped_keyframes();
		options.rounding_policy = interpolate ? ACLB200_ROUND_NONE : ACLB200_ROUND_NEAREST;

		aclb200_error_job job;
		std::memset(&job, 0, sizeof(job));
		job.clip = 0;
		job.num_samples = num_samples;
		job.sample_rate = tracks0->get_sample_rate();
		job.duration = tracks0->get_finite_duration();
		job.num_tracks = num_tracks;

		const uint32_t components = batch0.info().track_type <= 3 ? batch0.info().track_type + 1 : 4u;
		shim_impl::device_buffer d_first, d_result;
		void* d_samples = d_first.get(device, size_t(num_samples) * num_tracks * components * sizeof(float));
		options.looping_policy = static_cast<uint32_t>(context0.get_looping_policy());
		device.check(aclb200_decompress_all_samples(device.get(), batch0.clipset(), &job, 1, &options, d_samples, nullptr), "aclb200_decompress_all_samples");
		aclb200_track_error* d_out = static_cast<aclb200_track_error*>(d_result.get(device, sizeof(aclb200_track_error)));
		options.rounding_policy = ACLB200_ROUND_NONE;		// the measurement picks the policy itself, from the clip it decodes (scalar clips never
																	// strip key frames, so both sides are sought with `nearest` like the reference's)
		options.looping_policy = static_cast<uint32_t>(context1.get_looping_policy());
		device.check(aclb200_calculate_compression_error(device.get(), batch1.clipset(), &job, 1, d_samples, nullptr, nullptr, nullptr, nullptr, &options, d_out, nullptr, nullptr),
			"aclb200_calculate_compression_error");
		aclb200_track_error result;
		device.check(aclb200_copy_to_host(device.get(), &result, d_out, sizeof(result)), "result download");
		acl::track_error out;
		out.index = result.index;
		out.error = result.error;
		out.sample_time = result.sample_time;
		return out;
	}
}
