// tests/cpp/shim_reference_callsite.cpp -- a REFERENCE call site compiled against both classes.
//
// Built with the reference's headers on the include path (-I<acl>/includes -I<rtm>/includes), so acl_b200/decompress.h takes the
// reference's own types. `play()` below is the body of the reference's benchmark loop
// (tools/acl_decompressor/sources/benchmark.cpp:246-258: seek(sample_time, none) then decompress_tracks(pose_writer), or
// decompress_track(bone_index, pose_writer) for every bone) with acl::acl_impl::debug_track_writer, instantiated once with
// acl::decompression_context<settings> and once with acl_b200::decompression_context<settings>: only the namespace differs.
//
// usage: shim_reference_callsite <clip.acl.bin> [other_clip.acl.bin]
//   exit 0 = every value bit-identical (decompress_track rotations: <= 1e-5, the reference normalises them with a CPU dependent
//   rsqrt estimate), 3 = no usable GPU (the library has no CPU fallback), 1 = mismatch.
#include <acl/core/ansi_allocator.h>
#include <acl/core/impl/debug_track_writer.h>
#include <acl/decompression/decompress.h>

#include "../../include/acl_b200/decompress.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <vector>

#if !ACLB200_WITH_ACL_HEADERS
	#error "this test must be compiled with the reference's headers on the include path"
#endif

namespace
{
	// tools/acl_decompressor/sources/benchmark.cpp:94-101
	struct benchmark_transform_decompression_settings final : public acl::default_transform_decompression_settings
	{
		static constexpr acl::compressed_tracks_version16 version_supported() { return acl::compressed_tracks_version16::latest; }
	};

	enum class DecompressionFunction { DecompressPose, DecompressBone };

	std::vector<char> read_file(const char* path)
	{
		std::ifstream file(path, std::ios::binary);
		return std::vector<char>((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
	}

	// 16 byte aligned copy of a compressed_tracks buffer (the reference requires the alignment)
	struct aligned_clip
	{
		explicit aligned_clip(const std::vector<char>& bytes) : storage(bytes.size() + 64 + 16)
		{
			char* base = storage.data();
			base += (16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15;
			std::memcpy(base, bytes.data(), bytes.size());
			tracks = reinterpret_cast<const acl::compressed_tracks*>(base);
		}
		std::vector<char> storage;
		const acl::compressed_tracks* tracks;
	};

	// The reference's benchmark loop body, for any context class with the reference's interface
	template<class context_type>
	void play(const acl::compressed_tracks& compressed_tracks, const std::vector<float>& sample_times, DecompressionFunction decompression_function,
		acl::acl_impl::debug_track_writer& pose_writer, std::vector<float>& out_values, size_t floats_per_pose)
	{
		context_type context;
		if (!context.initialize(compressed_tracks))
		{
			std::fprintf(stderr, "initialize failed\n");
			std::exit(1);
		}
		const uint32_t num_tracks = compressed_tracks.get_num_tracks();
		for (const float sample_time : sample_times)
		{
			// every sub-track the writer's `skipped` default mode leaves alone keeps this pattern
			float* pose = static_cast<float*>(pose_writer.tracks_typed.any);
			for (size_t i = 0; i < floats_per_pose; ++i)
				pose[i] = -7.25F;

			// Interpolate as this is the most common scenario
			context.seek(sample_time, acl::sample_rounding_policy::none);

			switch (decompression_function)
			{
			case DecompressionFunction::DecompressPose:
				context.decompress_tracks(pose_writer);
				break;
			case DecompressionFunction::DecompressBone:
				for (uint32_t bone_index = 0; bone_index < num_tracks; ++bone_index)
					context.decompress_track(bone_index, pose_writer);
				break;
			}
			out_values.insert(out_values.end(), pose, pose + floats_per_pose);
		}
	}

	bool compare(const std::vector<float>& reference, const std::vector<float>& ours, bool transform, float rotation_tolerance, const char* what)
	{
		if (reference.size() != ours.size())
			return false;
		size_t mismatches = 0;
		float worst = 0.0F;
		for (size_t i = 0; i < reference.size(); ++i)
		{
			const size_t lane = transform ? i % 12 : 0;
			if (transform && (lane == 7 || lane == 11))
				continue;		// translation.w / scale.w are unspecified in the reference (animated_track_cache.transform.h:964)
			uint32_t a, b;
			std::memcpy(&a, &reference[i], 4);
			std::memcpy(&b, &ours[i], 4);
			if (a == b)
				continue;
			const float difference = std::fabs(reference[i] - ours[i]);
			if (transform && lane < 4 && difference <= rotation_tolerance)
			{
				worst = difference > worst ? difference : worst;
				continue;
			}
			if (mismatches++ < 5)
				std::fprintf(stderr, "%s: value %zu differs: reference %.9g ours %.9g\n", what, i, double(reference[i]), double(ours[i]));
		}
		std::printf("%s: %zu values, %zu mismatches, worst tolerated rotation difference %.3g\n", what, reference.size(), mismatches, double(worst));
		return mismatches == 0;
	}
}

int main(int argc, char** argv)
{
	if (argc < 2)
		return 2;
	const std::vector<char> bytes = read_file(argv[1]);
	if (bytes.empty())
		return 2;
	aligned_clip clip(bytes);
	const acl::compressed_tracks& tracks = *clip.tracks;
	const bool transform = tracks.get_track_type() == acl::track_type8::qvvf;
	const uint32_t num_tracks = tracks.get_num_tracks();
	const float duration = tracks.get_finite_duration();

	std::vector<float> sample_times;
	for (int i = 0; i <= 24; ++i)
		sample_times.push_back(duration * float(i) / 24.0F);
	sample_times.push_back(-1.0F);
	sample_times.push_back(duration + 1.0F);

	acl::ansi_allocator allocator;
	bool ok = true;
	try
	{
		if (transform)
		{
			using settings = benchmark_transform_decompression_settings;
			const size_t floats_per_pose = size_t(num_tracks) * 12;
			for (DecompressionFunction function : { DecompressionFunction::DecompressPose, DecompressionFunction::DecompressBone })
			{
				acl::acl_impl::debug_track_writer pose_writer(allocator, acl::track_type8::qvvf, num_tracks);
				std::vector<float> reference, ours;
				play<acl::decompression_context<settings>>(tracks, sample_times, function, pose_writer, reference, floats_per_pose);
				play<acl_b200::decompression_context<settings>>(tracks, sample_times, function, pose_writer, ours, floats_per_pose);
				const bool pose = function == DecompressionFunction::DecompressPose;
				ok = compare(reference, ours, true, pose ? 0.0F : 1.0e-5F, pose ? "decompress_tracks" : "decompress_track") && ok;
			}
		}
		else
		{
			using settings = acl::debug_scalar_decompression_settings;
			const uint32_t components = tracks.get_track_type() == acl::track_type8::float1f ? 1 : tracks.get_track_type() == acl::track_type8::float2f ? 2
				: tracks.get_track_type() == acl::track_type8::float3f ? 3 : 4;
			const size_t floats_per_pose = size_t(num_tracks) * components;
			for (DecompressionFunction function : { DecompressionFunction::DecompressPose, DecompressionFunction::DecompressBone })
			{
				acl::acl_impl::debug_track_writer pose_writer(allocator, tracks.get_track_type(), num_tracks);
				std::vector<float> reference, ours;
				play<acl::decompression_context<settings>>(tracks, sample_times, function, pose_writer, reference, floats_per_pose);
				play<acl_b200::decompression_context<settings>>(tracks, sample_times, function, pose_writer, ours, floats_per_pose);
				ok = compare(reference, ours, false, 0.0F, function == DecompressionFunction::DecompressPose ? "scalar decompress_tracks" : "scalar decompress_track") && ok;
			}
		}

		// ---- initialize / relocated / is_bound_to follow the reference (decompress.impl.h:66-205, decompression.transform.h:134-176) ----
		{
			using settings = acl::debug_transform_decompression_settings;
			acl::decompression_context<settings> reference_context;
			acl_b200::decompression_context<settings> context;
			aligned_clip moved(bytes);
			bool same = true;
			same = same && reference_context.relocated(*moved.tracks) == context.relocated(*moved.tracks);		// not initialised: false
			if (transform)
			{
				same = same && reference_context.initialize(tracks) && context.initialize(tracks);
				same = same && reference_context.is_bound_to(tracks) == context.is_bound_to(tracks);
				same = same && reference_context.is_bound_to(*moved.tracks) == context.is_bound_to(*moved.tracks);		// another address: false
				same = same && reference_context.relocated(*moved.tracks) == context.relocated(*moved.tracks);			// same clip elsewhere: true
				same = same && context.get_compressed_tracks() == moved.tracks && context.is_bound_to(*moved.tracks);
				if (argc > 2)
				{
					const std::vector<char> other_bytes = read_file(argv[2]);
					aligned_clip other(other_bytes);
					const bool reference_result = reference_context.relocated(*other.tracks), our_result = context.relocated(*other.tracks);
					same = same && reference_result == our_result && !our_result;		// a DIFFERENT clip is refused (hash differs)
					same = same && context.get_compressed_tracks() == moved.tracks;
				}
				// settings that do not support the clip's formats: initialize() refuses, like the reference's static format selection would assert
				struct full_only_settings final : public acl::debug_transform_decompression_settings
				{
					static constexpr bool is_rotation_format_supported(acl::rotation_format8 format) { return format == acl::rotation_format8::quatf_full; }
				};
				acl_b200::decompression_context<full_only_settings> picky;
				const bool clip_is_full = acl::acl_impl::get_tracks_header(tracks).get_rotation_format() == acl::rotation_format8::quatf_full;
				same = same && picky.initialize(tracks) == clip_is_full;
			}
			std::printf("binding semantics: %s\n", same ? "same as the reference" : "DIFFERENT");
			ok = ok && same;
		}
	}
	catch (const acl_b200::error& e)
	{
		std::fprintf(stderr, "%s\n", e.what());
		return e.status == ACLB200_ERR_NO_DEVICE ? 3 : 1;
	}
	std::printf(ok ? "PASS\n" : "FAIL\n");
	return ok ? 0 : 1;
}
