// tests/cpp/shim_track_error.cpp -- a REFERENCE call site of calculate_compression_error compiled against both namespaces.
//
// Built with the reference's headers on the include path. A raw clip is synthesised and compressed with the reference's own
// compressor (acl::compress_track_list); `measure()` is what tools/acl_compressor does after compressing
// (calculate_compression_error(allocator, raw_tracks, context, error_metric)), instantiated once with acl::decompression_context +
// acl::calculate_compression_error and once with acl_b200::decompression_context + acl_b200::calculate_compression_error: only the
// namespace differs.
//
// usage: shim_track_error            exit 0 = PASS (errors within 5e-5 -- a different worst track is a tie within that tolerance --;
//                                    scalar clips and the matrix metric: exact, same track and sample time),
//                                    3 = no usable GPU (the library has no CPU fallback), 1 = mismatch
#include <acl/core/ansi_allocator.h>
#include <acl/compression/compress.h>
#include <acl/compression/track_array.h>
#include <acl/compression/track_error.h>
#include <acl/compression/transform_error_metrics.h>
#include <acl/decompression/decompress.h>

#include "../../include/acl_b200/track_error.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace
{
	acl::ansi_allocator g_allocator;

	acl::track_array_qvvf make_transform_clip(uint32_t num_tracks, uint32_t num_samples, uint32_t seed, bool with_scale, bool strip_one_track)
	{
		acl::track_array_qvvf tracks(g_allocator, num_tracks);
		uint32_t state = seed * 2654435761u + 12345u;
		const auto next = [&state]() { state = state * 1664525u + 1013904223u; return float((state >> 8) & 0xFFFF) / 65535.0F; };
		uint32_t output_index = 0;
		for (uint32_t bone = 0; bone < num_tracks; ++bone)
		{
			acl::track_desc_transformf desc;
			desc.parent_index = bone == 0 ? acl::k_invalid_track_index : (bone - 1) / 3;
			desc.precision = 0.01F;
			desc.shell_distance = 1.0F + 3.0F * next();
			const bool stripped = strip_one_track && bone == num_tracks - 2;
			desc.output_index = stripped ? acl::k_invalid_track_index : output_index++;
			// a bind pose that is not the identity: default sub-tracks must read it
			desc.default_value = rtm::qvv_set(rtm::quat_identity(), rtm::vector_set(0.5F, -0.25F, 0.125F, 0.0F), rtm::vector_set(1.0F));

			const float axis_x = next() - 0.5F, axis_y = next() - 0.5F, axis_z = next() + 0.1F;
			const rtm::vector4f axis = rtm::vector_normalize3(rtm::vector_set(axis_x, axis_y, axis_z, 0.0F), rtm::vector_set(0.0F, 0.0F, 1.0F, 0.0F));
			const float frequency = 1.0F + 3.0F * next(), phase = 6.0F * next();
			const rtm::vector4f base_translation = rtm::vector_set(10.0F * (next() - 0.5F), 10.0F * (next() - 0.5F), 10.0F * (next() - 0.5F), 0.0F);
			const bool animated_translation = (bone % 4) == 0;
			const bool default_translation = (bone % 7) == 3;
			const bool constant_rotation = (bone % 5) == 2;

			acl::track_qvvf track = acl::track_qvvf::make_reserve(desc, g_allocator, num_samples, 30.0F);
			for (uint32_t sample = 0; sample < num_samples; ++sample)
			{
				const float t = float(sample) / 30.0F;
				rtm::qvvf value = desc.default_value;
				value.rotation = rtm::quat_from_axis_angle(axis, constant_rotation ? 0.7F : 0.8F * std::sin(frequency * t + phase));
				if (!default_translation)
					value.translation = animated_translation ? rtm::vector_add(base_translation, rtm::vector_set(std::sin(t * 2.0F + phase), std::cos(t * 3.0F), 0.3F * t, 0.0F)) : base_translation;
				if (with_scale && (bone % 3) == 1)
					value.scale = rtm::vector_set(1.0F + 0.25F * std::sin(t + phase), 1.1F, 0.9F + 0.05F * t, 0.0F);
				track[sample] = value;
			}
			tracks[bone] = std::move(track);
		}
		return tracks;
	}

	acl::track_array_float3f make_scalar_clip(uint32_t num_tracks, uint32_t num_samples, float precision = 0.001F)
	{
		acl::track_array_float3f tracks(g_allocator, num_tracks);
		for (uint32_t index = 0; index < num_tracks; ++index)
		{
			acl::track_desc_scalarf desc;
			desc.output_index = index;
			desc.precision = precision;
			acl::track_float3f track = acl::track_float3f::make_reserve(desc, g_allocator, num_samples, 30.0F);
			for (uint32_t sample = 0; sample < num_samples; ++sample)
			{
				const float t = float(sample) / 30.0F;
				track[sample] = rtm::float3f{ std::sin(t * (1.0F + float(index))), (index % 3) == 0 ? 2.0F : std::cos(t + float(index)), 0.1F * t * float(index) };
			}
			tracks[index] = std::move(track);
		}
		return tracks;
	}

	// the call site: any context class + the calculate_compression_error found next to it
	template<class context_type>
	acl::track_error measure(const acl::compressed_tracks& compressed, const acl::track_array& raw_tracks, bool acl_b200_namespace)
	{
		context_type context;
		if (!context.initialize(compressed))
		{
			std::fprintf(stderr, "initialize failed\n");
			std::exit(1);
		}
		const acl::qvvf_transform_error_metric error_metric;
		(void)acl_b200_namespace;
		return calculate_compression_error(g_allocator, raw_tracks, context, error_metric);		// argument dependent lookup picks the namespace of `context`
	}

	bool check(const char* what, const acl::track_error& reference, const acl::track_error& ours, float tolerance)
	{
		const bool same_error = std::fabs(reference.error - ours.error) <= tolerance;
		// Errors that agree within the tolerance at two different places are a tie within that tolerance (the reference itself moves between
		// CPU models there: its quat_normalize starts from the CPU's rsqrtss estimate); an exact contract (tolerance 0) leaves no such room:
		// equal errors at different places would mean one side did not keep the FIRST maximum.
		const bool same_place = (reference.index == ours.index && reference.sample_time == ours.sample_time) || tolerance > 0.0F;
		std::printf("%s: reference (track %u, error %.9g, t %.6g) ours (track %u, error %.9g, t %.6g) %s\n", what, reference.index, double(reference.error),
			double(reference.sample_time), ours.index, double(ours.error), double(ours.sample_time), same_error && same_place ? "ok" : "MISMATCH");
		return same_error && same_place;
	}
}

int main()
{
	// SHIM_TRACK_ERROR_REFERENCE_ONLY=1: only the reference half runs (checks the call sites themselves on a machine without a GPU)
	const bool reference_only = std::getenv("SHIM_TRACK_ERROR_REFERENCE_ONLY") != nullptr;
	bool ok = true;
	try
	{
		struct { const char* name; uint32_t tracks, samples, seed; bool scale, strip; acl::rotation_format8 rotation_format; float strip_proportion; } cases[] = {
			{ "transform 40 x 50", 40, 50, 1, false, false, acl::rotation_format8::quatf_drop_w_variable, 0.0F },
			{ "transform with scale 57 x 75", 57, 75, 2, true, false, acl::rotation_format8::quatf_drop_w_variable, 0.0F },
			{ "transform with a stripped track 33 x 20", 33, 20, 3, false, true, acl::rotation_format8::quatf_drop_w_variable, 0.0F },
			{ "transform full precision 21 x 40", 21, 40, 4, true, false, acl::rotation_format8::quatf_full, 0.0F },
			{ "transform with stripped key frames 40 x 60", 40, 60, 5, false, false, acl::rotation_format8::quatf_drop_w_variable, 0.5F },
		};
		for (const auto& c : cases)
		{
			const acl::track_array_qvvf raw = make_transform_clip(c.tracks, c.samples, c.seed, c.scale, c.strip);
			acl::qvvf_transform_error_metric error_metric;
			acl::compression_settings settings = acl::get_default_compression_settings();
			settings.error_metric = &error_metric;
			settings.rotation_format = c.rotation_format;
			if (c.rotation_format == acl::rotation_format8::quatf_full)
			{
				settings.translation_format = settings.scale_format = acl::vector_format8::vector3f_full;
				settings.keyframe_stripping.strip_trivial = false;		// raw tracks have no contributing error to strip by (compress.transform.impl.h:169-172)
			}
			settings.keyframe_stripping.proportion = c.strip_proportion;
			acl::compressed_tracks* compressed = nullptr;
			acl::output_stats stats;
			const acl::error_result result = acl::compress_track_list(g_allocator, raw, settings, compressed, stats);
			if (result.any())
			{
				std::fprintf(stderr, "compression failed: %s\n", result.c_str());
				return 1;
			}
			const acl::track_error reference = measure<acl::decompression_context<acl::debug_transform_decompression_settings>>(*compressed, raw, false);
			const acl::track_error ours = reference_only ? reference : measure<acl_b200::decompression_context<acl::debug_transform_decompression_settings>>(*compressed, raw, true);
			ok = check(c.name, reference, ours, 5.0e-5F) && ok;

			// the matrix metric has no CPU specific step: exact
			{
				acl::decompression_context<acl::debug_transform_decompression_settings> reference_context;
				acl_b200::decompression_context<acl::debug_transform_decompression_settings> our_context;
				if (!reference_context.initialize(*compressed) || (!reference_only && !our_context.initialize(*compressed)))
					return 1;
				const acl::qvvf_matrix3x4f_transform_error_metric matrix_metric;
				const acl::track_error matrix_reference = acl::calculate_compression_error(g_allocator, raw, reference_context, matrix_metric);
				const acl::track_error matrix_ours = reference_only ? matrix_reference : acl_b200::calculate_compression_error(g_allocator, raw, our_context, matrix_metric);
				const bool exact = matrix_reference.index == matrix_ours.index && matrix_reference.error == matrix_ours.error && matrix_reference.sample_time == matrix_ours.sample_time;
				std::printf("%s, qvvf_matrix3x4f metric: reference (track %u, error %.9g) ours (track %u, error %.9g) %s\n", c.name, matrix_reference.index,
					double(matrix_reference.error), matrix_ours.index, double(matrix_ours.error), exact ? "ok" : "MISMATCH");
				ok = exact && ok;
			}
			g_allocator.deallocate(compressed, compressed->get_size());
		}

		// the additive overload: the clip is measured on top of a base clip of another length (track_error.h:93-107)
		{
			const acl::track_array_qvvf raw = make_transform_clip(36, 40, 7, true, false);
			const acl::track_array_qvvf base = make_transform_clip(36, 23, 8, true, false);
			acl::qvvf_transform_error_metric plain_metric;
			acl::compression_settings settings = acl::get_default_compression_settings();
			settings.error_metric = &plain_metric;
			acl::compressed_tracks* compressed = nullptr;
			acl::output_stats stats;
			if (acl::compress_track_list(g_allocator, raw, settings, compressed, stats).any())
				return 1;
			acl::decompression_context<acl::debug_transform_decompression_settings> reference_context;
			acl_b200::decompression_context<acl::debug_transform_decompression_settings> our_context;
			if (!reference_context.initialize(*compressed) || (!reference_only && !our_context.initialize(*compressed)))
				return 1;
			const acl::additive_qvvf_transform_error_metric<acl::additive_clip_format8::relative> relative_metric;
			const acl::additive_qvvf_transform_error_metric<acl::additive_clip_format8::additive0> additive0_metric;
			const acl::additive_qvvf_transform_error_metric<acl::additive_clip_format8::additive1> additive1_metric;
			const acl::itransform_error_metric* metrics[] = { &relative_metric, &additive0_metric, &additive1_metric };
			for (const acl::itransform_error_metric* metric : metrics)
			{
				const acl::track_error reference = acl::calculate_compression_error(g_allocator, raw, reference_context, *metric, base);
				const acl::track_error ours = reference_only ? reference : acl_b200::calculate_compression_error(g_allocator, raw, our_context, *metric, base);
				ok = check(metric->get_name(), reference, ours, 5.0e-5F + 1.0e-4F * reference.error) && ok;
			}
			g_allocator.deallocate(compressed, compressed->get_size());
		}

		{
			const acl::track_array_float3f raw = make_scalar_clip(19, 45);
			acl::compressed_tracks* compressed = nullptr;
			acl::output_stats stats;
			if (acl::compress_track_list(g_allocator, raw, acl::compression_settings(), compressed, stats).any())
				return 1;
			acl::decompression_context<acl::default_scalar_decompression_settings> reference_context;
			acl_b200::decompression_context<acl::default_scalar_decompression_settings> our_context;
			if (!reference_context.initialize(*compressed) || (!reference_only && !our_context.initialize(*compressed)))
				return 1;
			const acl::track_error reference = acl::calculate_compression_error(g_allocator, raw, reference_context);
			const acl::track_error ours = reference_only ? reference : acl_b200::calculate_compression_error(g_allocator, raw, our_context);
			ok = check("scalar float3f 19 x 45", reference, ours, 0.0F) && ok;

			// calculate_compression_error(allocator, context0, context1) (track_error.h:109-121): the same raw clip compressed coarsely
			const acl::track_array_float3f coarse_raw = make_scalar_clip(19, 45, 0.05F);
			acl::compressed_tracks* coarse = nullptr;
			if (acl::compress_track_list(g_allocator, coarse_raw, acl::compression_settings(), coarse, stats).any())
				return 1;
			acl::decompression_context<acl::default_scalar_decompression_settings> reference_coarse;
			acl_b200::decompression_context<acl::default_scalar_decompression_settings> our_coarse;
			if (!reference_coarse.initialize(*coarse) || (!reference_only && !our_coarse.initialize(*coarse)))
				return 1;
			const acl::track_error pair_reference = acl::calculate_compression_error(g_allocator, reference_context, reference_coarse);
			const acl::track_error pair_ours = reference_only ? pair_reference : acl_b200::calculate_compression_error(g_allocator, our_context, our_coarse);
			ok = check("scalar float3f, fine clip against coarse clip", pair_reference, pair_ours, 0.0F) && ok;
			g_allocator.deallocate(coarse, coarse->get_size());
			g_allocator.deallocate(compressed, compressed->get_size());
		}
	}
	catch (const acl_b200::error& failure)
	{
		std::fprintf(stderr, "%s\n", failure.what());
		return failure.status == ACLB200_ERR_NO_DEVICE ? 3 : 1;
	}
	std::printf(ok ? "PASS\n" : "FAIL\n");
	return ok ? 0 : 1;
}
