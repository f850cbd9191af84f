// oracle/ref_tool.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin C-ABI wrapper around the UNMODIFIED reference (nfrechette/acl, header-only) so that the
// tests and the CPU-baseline leg of bench.py can
//   (1) synthesise raw clips and compress them with the reference compressor
//       (acl::compress_track_list, includes/acl/compression/compress.h:48-76),
//   (2) decompress with the reference decompression_context
//       (includes/acl/decompression/decompress.h:76-201) to produce golden outputs,
//   (3) expose the reference's seek() integers (key frames, segment bit offsets, alpha),
//   (4) time the reference's CPU path on N host threads (bench.py --impl reference),
//   (5) run the reference's calculate_compression_error (compression/impl/track_error.impl.h:400-571) and expose the
//       values it works on (raw poses, lossy poses, per bone object space error) for the SURVEY 8(f1) parity tests.
//
// It is compiled from the reference headers where they lie (/root/reference) by oracle/Makefile into
// oracle/_ref/libaclref.so (git-ignored, ships to the GPU box as a prebuilt file). No reference
// source is copied into this repository: this file only *calls* the reference's public API plus
// two acl_impl entry points (seek_v0 / initialize_v0) used for integer introspection.

#include <acl/core/ansi_allocator.h>
#include <acl/core/compressed_tracks.h>
#include <acl/compression/compress.h>
#include <acl/compression/track_array.h>
#include <acl/compression/track_error.h>
#include <acl/compression/transform_error_metrics.h>
#include <acl/decompression/decompress.h>
#include <acl/decompression/database/database.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
	acl::ansi_allocator& allocator()
	{
		static acl::ansi_allocator* s_alloc = new acl::ansi_allocator();	// leaked on purpose, process lifetime
		return *s_alloc;
	}

	// Small deterministic RNG (splitmix64) so clips depend on the seed only, never on libstdc++.
	struct rng_t
	{
		uint64_t state;
		explicit rng_t(uint64_t seed) : state(seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL) {}
		uint64_t next()
		{
			uint64_t z = (state += 0x9E3779B97F4A7C15ULL);
			z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
			z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
			return z ^ (z >> 31);
		}
		double uniform() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }			// [0,1)
		double range(double lo, double hi) { return lo + (hi - lo) * uniform(); }
		uint32_t pct() { return uint32_t(next() % 100); }
	};
}

extern "C"
{
	// Keep in sync with acl_b200/reference.py (ctypes mirror). All fields are 32 bit wide.
	struct aclref_transform_spec
	{
		uint32_t num_tracks;
		uint32_t num_samples;
		float    sample_rate;
		uint32_t seed;

		// Content, percentages in [0, 100]; the remainder of each sub-track kind is animated.
		uint32_t rot_default_pct;
		uint32_t rot_constant_pct;
		uint32_t trans_default_pct;
		uint32_t trans_constant_pct;
		uint32_t scale_default_pct;		// 100 => no scale in the clip at all
		uint32_t scale_constant_pct;
		uint32_t partial_activity_pct;	// animated sub-tracks that only move in the first third (=> constant within later segments)
		uint32_t noisy_pct;				// animated sub-tracks with white noise + tiny precision (=> highest / raw bit rates)
		uint32_t looping_content;		// 1 => last sample == first sample (lets optimize_loops kick in)
		float    translation_range;		// constant translations are uniform in [-range, range]^3
		float    precision;
		float    shell_distance;

		// Compression settings (acl::compression_settings, compression_settings.h:201-254)
		uint32_t rotation_format;		// acl::rotation_format8
		uint32_t translation_format;	// acl::vector_format8
		uint32_t scale_format;			// acl::vector_format8
		uint32_t level;					// acl::compression_level8
		uint32_t optimize_loops;
		uint32_t strip_trivial;
		float    strip_proportion;
		float    strip_threshold;
		float    rotation_offset;		// radians added to every animated rotation angle (pi => W crosses 0, the ill-conditioned end of W reconstruction)
		uint32_t negative_scale_pct;	// bones whose scale.x is mirrored (rtm::qvv_mul's matrix branch in the error metric); consumes no random numbers
	};

	struct aclref_scalar_spec
	{
		uint32_t num_tracks;
		uint32_t num_samples;
		float    sample_rate;
		uint32_t seed;
		uint32_t track_type;			// acl::track_type8: float1f=0 .. float4f=3, vector4f=4
		uint32_t constant_pct;
		uint32_t noisy_pct;
		float    precision;
	};

	// Mirrors what acl_impl::seek_v0 leaves in persistent_transform_decompression_context_v0
	// (decompression_context.transform.h:53-116).
	struct aclref_seek_info
	{
		float    sample_time;			// clamped
		float    interpolation_alpha;
		uint32_t key_frame_bit_offsets[2];
		uint32_t segment_offsets[2];	// byte offset of the segment headers relative to the blob start
		uint32_t format_offsets[2];		// byte offsets of format_per_track_data relative to the blob start
		uint32_t range_offsets[2];
		uint32_t animated_offsets[2];
		uint32_t uses_single_segment;
		float    clip_duration;
		uint32_t looping_policy;
	};
}

namespace
{
	using namespace acl;

	//////////////////////////////////////////////////////////////////////////
	// Synthetic raw clips

	rtm::quatf make_rotation(const double axis[3], double angle)
	{
		const double s = std::sin(angle * 0.5);
		const double c = std::cos(angle * 0.5);
		rtm::quatf q = rtm::quat_set(float(axis[0] * s), float(axis[1] * s), float(axis[2] * s), float(c));
		return rtm::quat_normalize(q);
	}

	// The raw clip of a spec (deterministic: the error metric harness below rebuilds the very clip that was compressed)
	void make_transform_tracks(const aclref_transform_spec& spec, track_array_qvvf& track_list)
	{
		iallocator& alloc = allocator();
		const uint32_t num_tracks = spec.num_tracks;
		const uint32_t num_samples = spec.num_samples;
		const float sample_rate = spec.sample_rate;

		rng_t rng(spec.seed);

		const double two_pi = 6.283185307179586;
		const double duration = num_samples > 1 ? double(num_samples - 1) / double(sample_rate) : 1.0;

		for (uint32_t bone = 0; bone < num_tracks; ++bone)
		{
			track_desc_transformf desc;
			desc.output_index = bone;
			desc.parent_index = bone == 0 ? k_invalid_track_index : (bone - 1) / 2;
			desc.precision = spec.precision;
			desc.shell_distance = spec.shell_distance;

			// Decide the kind of each sub-track
			const uint32_t rot_roll = rng.pct();
			const uint32_t trans_roll = rng.pct();
			const uint32_t scale_roll = rng.pct();
			const int rot_kind = rot_roll < spec.rot_default_pct ? 0 : (rot_roll < spec.rot_default_pct + spec.rot_constant_pct ? 1 : 2);
			const int trans_kind = trans_roll < spec.trans_default_pct ? 0 : (trans_roll < spec.trans_default_pct + spec.trans_constant_pct ? 1 : 2);
			const int scale_kind = scale_roll < spec.scale_default_pct ? 0 : (scale_roll < spec.scale_default_pct + spec.scale_constant_pct ? 1 : 2);

			const bool rot_partial = rng.pct() < spec.partial_activity_pct;
			const bool trans_partial = rng.pct() < spec.partial_activity_pct;
			const bool scale_partial = rng.pct() < spec.partial_activity_pct;
			const bool rot_noisy = rng.pct() < spec.noisy_pct;
			const bool trans_noisy = rng.pct() < spec.noisy_pct;
			const bool scale_noisy = rng.pct() < spec.noisy_pct;
			if (rot_noisy || trans_noisy || scale_noisy)
				desc.precision = 1.0e-7F;

			// Rotation parameters
			double axis[3] = { rng.range(-1.0, 1.0), rng.range(-1.0, 1.0), rng.range(-1.0, 1.0) };
			double axis_len = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
			if (axis_len < 1.0e-3) { axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0; axis_len = 1.0; }
			axis[0] /= axis_len; axis[1] /= axis_len; axis[2] /= axis_len;
			const double rot_base = rng.range(-1.0, 1.0);
			const double rot_phase = rng.range(0.0, two_pi);
			const double rot_phase2 = rng.range(0.0, two_pi);
			double rot_freq = rng.range(0.5, 2.5);
			double rot_freq2 = 13.0;

			// Translation / scale parameters
			const double trans_base[3] = { rng.range(-spec.translation_range, spec.translation_range), rng.range(-spec.translation_range, spec.translation_range), rng.range(-spec.translation_range, spec.translation_range) };
			const double trans_phase[3] = { rng.range(0.0, two_pi), rng.range(0.0, two_pi), rng.range(0.0, two_pi) };
			double trans_freq = rng.range(0.5, 3.0);
			const double scale_base[3] = { rng.range(0.5, 2.0), rng.range(0.5, 2.0), rng.range(0.5, 2.0) };
			const double scale_phase[3] = { rng.range(0.0, two_pi), rng.range(0.0, two_pi), rng.range(0.0, two_pi) };
			double scale_freq = rng.range(0.5, 3.0);

			if (spec.looping_content != 0)
			{
				// Snap every angular frequency onto a whole number of periods over the clip so that
				// the last sample repeats the first one.
				rot_freq = two_pi * std::floor(1.0 + rot_freq) / duration;
				rot_freq2 = two_pi * 3.0 / duration;
				trans_freq = two_pi * std::floor(1.0 + trans_freq) / duration;
				scale_freq = two_pi * std::floor(1.0 + scale_freq) / duration;
			}

			rng_t noise(uint64_t(spec.seed) * 7919ULL + bone);

			track_qvvf track = track_qvvf::make_reserve(desc, alloc, num_samples, sample_rate);
			for (uint32_t sample = 0; sample < num_samples; ++sample)
			{
				const double t = double(sample) / double(sample_rate);
				const bool looped_last = spec.looping_content != 0 && sample + 1 == num_samples && num_samples > 1;
				const double tt = looped_last ? 0.0 : t;	// exact repeat of the first sample
				const uint32_t sample_for_noise = looped_last ? 0 : sample;
				(void)sample_for_noise;

				rtm::qvvf transform = rtm::qvv_identity();

				// Rotation
				if (rot_kind == 1)
					transform.rotation = make_rotation(axis, rot_base * 2.0);
				else if (rot_kind == 2)
				{
					const double activity = (rot_partial && sample * 3 >= num_samples) ? 0.0 : 1.0;
					double angle = double(spec.rotation_offset) + rot_base + activity * (0.6 * std::sin(rot_freq * tt + rot_phase) + 0.05 * std::sin(rot_freq2 * tt + rot_phase2));
					if (rot_noisy && !looped_last)
						angle += noise.range(-0.3, 0.3);
					transform.rotation = make_rotation(axis, angle);
				}

				// Translation
				if (trans_kind == 1)
					transform.translation = rtm::vector_set(float(trans_base[0]), float(trans_base[1]), float(trans_base[2]), 0.0F);
				else if (trans_kind == 2)
				{
					const double activity = (trans_partial && sample * 3 >= num_samples) ? 0.0 : 1.0;
					double v[3];
					for (int c = 0; c < 3; ++c)
					{
						v[c] = trans_base[c] + activity * std::sin(trans_freq * tt + trans_phase[c]);
						if (trans_noisy && !looped_last)
							v[c] += noise.range(-0.5, 0.5);
					}
					transform.translation = rtm::vector_set(float(v[0]), float(v[1]), float(v[2]), 0.0F);
				}

				// Scale
				if (scale_kind == 1)
					transform.scale = rtm::vector_set(float(scale_base[0]), float(scale_base[1]), float(scale_base[2]), 0.0F);
				else if (scale_kind == 2)
				{
					const double activity = (scale_partial && sample * 3 >= num_samples) ? 0.0 : 1.0;
					double v[3];
					for (int c = 0; c < 3; ++c)
					{
						v[c] = scale_base[c] + activity * 0.2 * std::sin(scale_freq * tt + scale_phase[c]);
						if (scale_noisy && !looped_last)
							v[c] += noise.range(-0.05, 0.05);
					}
					transform.scale = rtm::vector_set(float(v[0]), float(v[1]), float(v[2]), 0.0F);
				}

				if ((bone * 7u + 3u) % 100u < spec.negative_scale_pct)
					transform.scale = rtm::vector_mul(transform.scale, rtm::vector_set(-1.0F, 1.0F, 1.0F, 0.0F));

				track[sample] = transform;
			}

			// A noisy looped clip must still repeat exactly
			if (spec.looping_content != 0 && num_samples > 1)
				track[num_samples - 1] = track[0];

			track_list[bone] = std::move(track);
		}
	}

	error_result build_transform_clip(const aclref_transform_spec& spec, compressed_tracks*& out_tracks)
	{
		iallocator& alloc = allocator();
		track_array_qvvf track_list(alloc, spec.num_tracks);
		make_transform_tracks(spec, track_list);

		qvvf_transform_error_metric error_metric;

		compression_settings settings;
		settings.level = static_cast<compression_level8>(spec.level);
		settings.rotation_format = static_cast<rotation_format8>(spec.rotation_format);
		settings.translation_format = static_cast<vector_format8>(spec.translation_format);
		settings.scale_format = static_cast<vector_format8>(spec.scale_format);
		settings.error_metric = &error_metric;
		settings.optimize_loops = spec.optimize_loops != 0;
		settings.keyframe_stripping.strip_trivial = spec.strip_trivial != 0;
		settings.keyframe_stripping.proportion = spec.strip_proportion;
		settings.keyframe_stripping.threshold = spec.strip_threshold;

		output_stats stats;
		return compress_track_list(alloc, track_list, settings, out_tracks, stats);
	}

	template<class track_type, class sample_type, uint32_t num_components>
	void fill_scalar_tracks(const aclref_scalar_spec& spec, track_array& track_list)
	{
		iallocator& alloc = allocator();
		rng_t rng(spec.seed);
		for (uint32_t index = 0; index < spec.num_tracks; ++index)
		{
			track_desc_scalarf desc;
			desc.output_index = index;
			desc.precision = spec.precision;

			const bool is_constant = rng.pct() < spec.constant_pct;
			const bool is_noisy = rng.pct() < spec.noisy_pct;
			if (is_noisy)
				desc.precision = 1.0e-8F;

			double base[4], amp[4], freq[4], phase[4];
			for (uint32_t c = 0; c < 4; ++c)
			{
				base[c] = rng.range(-5.0, 5.0);
				amp[c] = rng.range(0.1, 2.0);
				freq[c] = rng.range(0.2, 4.0);
				phase[c] = rng.range(0.0, 6.283185307179586);
			}

			rng_t noise(uint64_t(spec.seed) * 104729ULL + index);
			track_type track = track_type::make_reserve(desc, alloc, spec.num_samples, spec.sample_rate);
			for (uint32_t sample = 0; sample < spec.num_samples; ++sample)
			{
				const double t = double(sample) / double(spec.sample_rate);
				float v[4];
				for (uint32_t c = 0; c < 4; ++c)
				{
					double value = base[c];
					if (!is_constant)
						value += amp[c] * std::sin(freq[c] * t + phase[c]);
					if (is_noisy && !is_constant)
						value += noise.range(-1.0, 1.0);
					v[c] = float(value);
				}
				std::memcpy(&track[sample], v, sizeof(float) * num_components);
			}
			track_list[index] = std::move(track);
		}
	}

	error_result build_scalar_clip(const aclref_scalar_spec& spec, compressed_tracks*& out_tracks)
	{
		iallocator& alloc = allocator();
		compression_settings settings;	// unused by the scalar path besides validation, see compress.scalar.impl.h:65-269
		output_stats stats;

		switch (static_cast<track_type8>(spec.track_type))
		{
		case track_type8::float1f: { track_array_float1f list(alloc, spec.num_tracks); fill_scalar_tracks<track_float1f, float, 1>(spec, list); return compress_track_list(alloc, list, settings, out_tracks, stats); }
		case track_type8::float2f: { track_array_float2f list(alloc, spec.num_tracks); fill_scalar_tracks<track_float2f, rtm::float2f, 2>(spec, list); return compress_track_list(alloc, list, settings, out_tracks, stats); }
		case track_type8::float3f: { track_array_float3f list(alloc, spec.num_tracks); fill_scalar_tracks<track_float3f, rtm::float3f, 3>(spec, list); return compress_track_list(alloc, list, settings, out_tracks, stats); }
		case track_type8::float4f: { track_array_float4f list(alloc, spec.num_tracks); fill_scalar_tracks<track_float4f, rtm::float4f, 4>(spec, list); return compress_track_list(alloc, list, settings, out_tracks, stats); }
		case track_type8::vector4f: { track_array_vector4f list(alloc, spec.num_tracks); fill_scalar_tracks<track_vector4f, rtm::vector4f, 4>(spec, list); return compress_track_list(alloc, list, settings, out_tracks, stats); }
		default: return error_result("unsupported track type");
		}
	}

	//////////////////////////////////////////////////////////////////////////
	// Decompression settings variants (decompression_settings.h:74-232)

	// kind 0: the library default for transforms
	using settings_default = default_transform_decompression_settings;

	// kind 1: everything enabled (all formats, always normalize, per track rounding)
	using settings_debug = debug_transform_decompression_settings;
	// debug settings + database support (decompression_settings.h:163-165): what a context needs to accept a clip bound to a database
	struct settings_database final : public debug_transform_decompression_settings
	{
		using database_settings_type = default_database_settings;
	};

	// kind 2: what tools/acl_decompressor/sources/benchmark.cpp:94-101 times
	struct settings_benchmark final : public default_transform_decompression_settings
	{
		static constexpr compressed_tracks_version16 version_supported() { return compressed_tracks_version16::latest; }
		static constexpr bool skip_initialize_safety_checks() { return true; }
	};

	// kind 3: all formats, never normalize
	struct settings_never final : public debug_transform_decompression_settings
	{
		static constexpr rotation_normalization_policy_t get_rotation_normalization_policy() { return rotation_normalization_policy_t::never; }
		static constexpr bool is_per_track_rounding_supported() { return false; }
	};

	// kind 4: all formats, lerp_only, no per track rounding
	struct settings_all_lerp final : public debug_transform_decompression_settings
	{
		static constexpr rotation_normalization_policy_t get_rotation_normalization_policy() { return rotation_normalization_policy_t::lerp_only; }
		static constexpr bool is_per_track_rounding_supported() { return false; }
	};

	// kind 5: full precision rotations only (exercises should_interpolate_samples(), decompression_context.transform.h:191-200)
	struct settings_raw_only final : public decompression_settings
	{
		static constexpr bool is_track_type_supported(track_type8 type) { return type == track_type8::qvvf; }
		static constexpr bool is_rotation_format_supported(rotation_format8 format) { return format == rotation_format8::quatf_full; }
		static constexpr bool is_translation_format_supported(vector_format8 format) { return format == vector_format8::vector3f_full; }
		static constexpr bool is_scale_format_supported(vector_format8 format) { return format == vector_format8::vector3f_full; }
		static constexpr bool is_per_track_rounding_supported() { return false; }
	};

	//////////////////////////////////////////////////////////////////////////
	// Track writers (core/track_writer.h:82-216)

	struct pose_writer_base : public track_writer
	{
		float* out = nullptr;						// [num_tracks][12] : rotation xyzw, translation xyz 0, scale xyz 0
		const uint8_t* per_track_rounding = nullptr;	// optional, one sample_rounding_policy per track
		const float* variable_defaults = nullptr;	// optional, [num_tracks][12]
		float constant_defaults[12] = { 0, 0, 0, 1,  0, 0, 0, 0,  1, 1, 1, 0 };

		sample_rounding_policy get_rounding_policy(sample_rounding_policy seek_policy, uint32_t track_index) const
		{
			if (seek_policy != sample_rounding_policy::per_track || per_track_rounding == nullptr)
				return seek_policy;
			return static_cast<sample_rounding_policy>(per_track_rounding[track_index]);
		}

		rtm::quatf RTM_SIMD_CALL get_constant_default_rotation() const { return rtm::quat_load(&constant_defaults[0]); }
		rtm::vector4f RTM_SIMD_CALL get_constant_default_translation() const { return rtm::vector_load(&constant_defaults[4]); }
		rtm::vector4f RTM_SIMD_CALL get_constant_default_scale() const { return rtm::vector_load(&constant_defaults[8]); }

		rtm::quatf RTM_SIMD_CALL get_variable_default_rotation(uint32_t track_index) const { return rtm::quat_load(&variable_defaults[track_index * 12 + 0]); }
		rtm::vector4f RTM_SIMD_CALL get_variable_default_translation(uint32_t track_index) const { return rtm::vector_load(&variable_defaults[track_index * 12 + 4]); }
		rtm::vector4f RTM_SIMD_CALL get_variable_default_scale(uint32_t track_index) const { return rtm::vector_load(&variable_defaults[track_index * 12 + 8]); }

		void RTM_SIMD_CALL write_rotation(uint32_t track_index, rtm::quatf_arg0 rotation) { rtm::quat_store(rotation, &out[track_index * 12 + 0]); }
		void RTM_SIMD_CALL write_translation(uint32_t track_index, rtm::vector4f_arg0 translation) { rtm::vector_store3(translation, &out[track_index * 12 + 4]); }
		void RTM_SIMD_CALL write_scale(uint32_t track_index, rtm::vector4f_arg0 scale) { rtm::vector_store3(scale, &out[track_index * 12 + 8]); }
	};

	// mode 0: library defaults (rotation/translation constant, scale legacy)
	struct pose_writer_legacy final : public pose_writer_base {};

	// mode 1: default sub-tracks are skipped, the output keeps what the caller put there
	struct pose_writer_skipped final : public pose_writer_base
	{
		static constexpr default_sub_track_mode get_default_rotation_mode() { return default_sub_track_mode::skipped; }
		static constexpr default_sub_track_mode get_default_translation_mode() { return default_sub_track_mode::skipped; }
		static constexpr default_sub_track_mode get_default_scale_mode() { return default_sub_track_mode::skipped; }
	};

	// mode 2: constant defaults provided by the writer
	struct pose_writer_constant final : public pose_writer_base
	{
		static constexpr default_sub_track_mode get_default_rotation_mode() { return default_sub_track_mode::constant; }
		static constexpr default_sub_track_mode get_default_translation_mode() { return default_sub_track_mode::constant; }
		static constexpr default_sub_track_mode get_default_scale_mode() { return default_sub_track_mode::constant; }
	};

	// mode 3: per track defaults provided by the writer
	struct pose_writer_variable final : public pose_writer_base
	{
		static constexpr default_sub_track_mode get_default_rotation_mode() { return default_sub_track_mode::variable; }
		static constexpr default_sub_track_mode get_default_translation_mode() { return default_sub_track_mode::variable; }
		static constexpr default_sub_track_mode get_default_scale_mode() { return default_sub_track_mode::variable; }
	};

	struct decode_args
	{
		const compressed_tracks* tracks;
		float sample_time;
		sample_rounding_policy rounding;
		sample_looping_policy looping;
		int32_t track_index;				// < 0 => decompress_tracks, else decompress_track
		float* out;
		const uint8_t* per_track_rounding;
		const float* constant_defaults;		// 12 floats or null
		const float* variable_defaults;		// [num_tracks][12] or null
	};

	template<class settings_type, class writer_type>
	int decode_impl(const decode_args& args)
	{
		decompression_context<settings_type> context;
		if (!context.initialize(*args.tracks))
			return -1;

		context.set_looping_policy(args.looping);

		writer_type writer;
		writer.out = args.out;
		writer.per_track_rounding = args.per_track_rounding;
		writer.variable_defaults = args.variable_defaults;
		if (args.constant_defaults != nullptr)
			std::memcpy(writer.constant_defaults, args.constant_defaults, sizeof(writer.constant_defaults));

		context.seek(args.sample_time, args.rounding);

		if (args.track_index < 0)
			context.decompress_tracks(writer);
		else
			context.decompress_track(uint32_t(args.track_index), writer);
		return 0;
	}

	template<class settings_type>
	int decode_dispatch_writer(uint32_t writer_mode, const decode_args& args)
	{
		switch (writer_mode)
		{
		case 0: return decode_impl<settings_type, pose_writer_legacy>(args);
		case 1: return decode_impl<settings_type, pose_writer_skipped>(args);
		case 2: return decode_impl<settings_type, pose_writer_constant>(args);
		case 3: return decode_impl<settings_type, pose_writer_variable>(args);
		default: return -2;
		}
	}

	int decode_dispatch(uint32_t settings_kind, uint32_t writer_mode, const decode_args& args)
	{
		switch (settings_kind)
		{
		case 0: return decode_dispatch_writer<settings_default>(writer_mode, args);
		case 1: return decode_dispatch_writer<settings_debug>(writer_mode, args);
		case 2: return decode_dispatch_writer<settings_benchmark>(writer_mode, args);
		case 3: return decode_dispatch_writer<settings_never>(writer_mode, args);
		case 4: return decode_dispatch_writer<settings_all_lerp>(writer_mode, args);
		case 5: return decode_dispatch_writer<settings_raw_only>(writer_mode, args);
		default: return -3;
		}
	}

	//////////////////////////////////////////////////////////////////////////
	// Scalar writers

	struct scalar_writer final : public track_writer
	{
		float* out = nullptr;		// [num_tracks][4]
		const uint8_t* per_track_rounding = nullptr;

		sample_rounding_policy get_rounding_policy(sample_rounding_policy seek_policy, uint32_t track_index) const
		{
			if (seek_policy != sample_rounding_policy::per_track || per_track_rounding == nullptr)
				return seek_policy;
			return static_cast<sample_rounding_policy>(per_track_rounding[track_index]);
		}

		void RTM_SIMD_CALL write_float1(uint32_t track_index, rtm::scalarf_arg0 value) { rtm::scalar_store(value, &out[track_index * 4]); }
		void RTM_SIMD_CALL write_float2(uint32_t track_index, rtm::vector4f_arg0 value) { rtm::vector_store2(value, &out[track_index * 4]); }
		void RTM_SIMD_CALL write_float3(uint32_t track_index, rtm::vector4f_arg0 value) { rtm::vector_store3(value, &out[track_index * 4]); }
		void RTM_SIMD_CALL write_float4(uint32_t track_index, rtm::vector4f_arg0 value) { rtm::vector_store(value, &out[track_index * 4]); }
		void RTM_SIMD_CALL write_vector4(uint32_t track_index, rtm::vector4f_arg0 value) { rtm::vector_store(value, &out[track_index * 4]); }
	};

	template<class settings_type>
	int scalar_decode_impl(const compressed_tracks& tracks, float sample_time, sample_rounding_policy rounding, sample_looping_policy looping,
		int32_t track_index, float* out, const uint8_t* per_track_rounding)
	{
		decompression_context<settings_type> context;
		if (!context.initialize(tracks))
			return -1;
		context.set_looping_policy(looping);

		scalar_writer writer;
		writer.out = out;
		writer.per_track_rounding = per_track_rounding;

		context.seek(sample_time, rounding);
		if (track_index < 0)
			context.decompress_tracks(writer);
		else
			context.decompress_track(uint32_t(track_index), writer);
		return 0;
	}
}

extern "C"
{
	#define ACLREF_STR2(x) #x
#define ACLREF_STR(x) ACLREF_STR2(x)
	const char* aclref_version() { return "acl-ref " ACLREF_STR(ACL_VERSION_MAJOR) "." ACLREF_STR(ACL_VERSION_MINOR) "." ACLREF_STR(ACL_VERSION_PATCH); }

	uint32_t aclref_hardware_threads() { return std::thread::hardware_concurrency(); }

	// Returns 0 on success. The blob is owned by the caller and released with aclref_free().
	int aclref_compress_transform(const aclref_transform_spec* spec, void** out_blob, uint32_t* out_size)
	{
		compressed_tracks* tracks = nullptr;
		const error_result result = build_transform_clip(*spec, tracks);
		if (result.any() || tracks == nullptr)
		{
			fprintf(stderr, "aclref_compress_transform: %s\n", result.any() ? result.c_str() : "no output");
			return -1;
		}

		const uint32_t size = tracks->get_size();
		// Keep the 16 byte alignment the format mandates (compressed_tracks.h:53) and 64 bytes of slack
		void* copy = nullptr;
		if (posix_memalign(&copy, 64, size + 64) != 0)
			return -2;
		std::memcpy(copy, tracks, size);
		std::memset(static_cast<uint8_t*>(copy) + size, 0, 64);
		allocator().deallocate(tracks, size);

		*out_blob = copy;
		*out_size = size;
		return 0;
	}

	int aclref_compress_scalar(const aclref_scalar_spec* spec, void** out_blob, uint32_t* out_size)
	{
		compressed_tracks* tracks = nullptr;
		const error_result result = build_scalar_clip(*spec, tracks);
		if (result.any() || tracks == nullptr)
		{
			fprintf(stderr, "aclref_compress_scalar: %s\n", result.any() ? result.c_str() : "no output");
			return -1;
		}

		const uint32_t size = tracks->get_size();
		void* copy = nullptr;
		if (posix_memalign(&copy, 64, size + 64) != 0)
			return -2;
		std::memcpy(copy, tracks, size);
		std::memset(static_cast<uint8_t*>(copy) + size, 0, 64);
		allocator().deallocate(tracks, size);

		*out_blob = copy;
		*out_size = size;
		return 0;
	}

	void aclref_free(void* blob) { free(blob); }

	// 0 when the reference accepts the buffer (compressed_tracks::is_valid(check_hash), compressed_tracks.impl.h:278-301)
	int aclref_is_valid(const void* blob, uint32_t check_hash)
	{
		const compressed_tracks* tracks = static_cast<const compressed_tracks*>(blob);
		return tracks->is_valid(check_hash != 0).empty() ? 0 : -1;
	}

	// looping: acl::sample_looping_policy (0 non_looping/clamp ... see sample_looping_policy.h), pass the
	// as_compressed value to keep the clip's own policy.
	int aclref_decompress_tracks(const void* blob, float sample_time, uint32_t rounding, uint32_t looping,
		uint32_t settings_kind, uint32_t writer_mode,
		const uint8_t* per_track_rounding, const float* constant_defaults, const float* variable_defaults,
		float* out)
	{
		decode_args args;
		args.tracks = static_cast<const compressed_tracks*>(blob);
		args.sample_time = sample_time;
		args.rounding = static_cast<sample_rounding_policy>(rounding);
		args.looping = static_cast<sample_looping_policy>(looping);
		args.track_index = -1;
		args.out = out;
		args.per_track_rounding = per_track_rounding;
		args.constant_defaults = constant_defaults;
		args.variable_defaults = variable_defaults;
		return decode_dispatch(settings_kind, writer_mode, args);
	}

	int aclref_decompress_track(const void* blob, float sample_time, uint32_t rounding, uint32_t looping,
		uint32_t settings_kind, uint32_t writer_mode, uint32_t track_index,
		const uint8_t* per_track_rounding, const float* constant_defaults, const float* variable_defaults,
		float* out)
	{
		decode_args args;
		args.tracks = static_cast<const compressed_tracks*>(blob);
		args.sample_time = sample_time;
		args.rounding = static_cast<sample_rounding_policy>(rounding);
		args.looping = static_cast<sample_looping_policy>(looping);
		args.track_index = int32_t(track_index);
		args.out = out;
		args.per_track_rounding = per_track_rounding;
		args.constant_defaults = constant_defaults;
		args.variable_defaults = variable_defaults;
		return decode_dispatch(settings_kind, writer_mode, args);
	}

	// settings_kind: 0 default_scalar (no per track rounding), 1 debug_scalar
	int aclref_scalar_decompress(const void* blob, float sample_time, uint32_t rounding, uint32_t looping,
		uint32_t settings_kind, int32_t track_index, const uint8_t* per_track_rounding, float* out)
	{
		const compressed_tracks& tracks = *static_cast<const compressed_tracks*>(blob);
		const sample_rounding_policy rounding_ = static_cast<sample_rounding_policy>(rounding);
		const sample_looping_policy looping_ = static_cast<sample_looping_policy>(looping);
		if (settings_kind == 0)
			return scalar_decode_impl<default_scalar_decompression_settings>(tracks, sample_time, rounding_, looping_, track_index, out, per_track_rounding);
		return scalar_decode_impl<debug_scalar_decompression_settings>(tracks, sample_time, rounding_, looping_, track_index, out, per_track_rounding);
	}

	// Integer introspection of seek (transform clips only). Uses the reference's own seek_v0 on its own
	// context structure (decompression.transform.h:84-132,206-563).
	int aclref_seek_info_transform(const void* blob, float sample_time, uint32_t rounding, uint32_t looping, aclref_seek_info* out_info)
	{
		const compressed_tracks* tracks = static_cast<const compressed_tracks*>(blob);
		if (tracks->get_track_type() != track_type8::qvvf)
			return -1;

		acl_impl::persistent_transform_decompression_context_v0 context;
		const database_context<null_database_settings>* db = nullptr;
		acl_impl::initialize_v0<debug_transform_decompression_settings>(context, *tracks, db);
		acl_impl::set_looping_policy_v0<debug_transform_decompression_settings>(context, static_cast<sample_looping_policy>(looping));
		acl_impl::seek_v0<debug_transform_decompression_settings>(context, sample_time, static_cast<sample_rounding_policy>(rounding));

		const uint8_t* base = static_cast<const uint8_t*>(blob);
		out_info->sample_time = context.sample_time;
		out_info->interpolation_alpha = context.interpolation_alpha;
		out_info->clip_duration = context.clip_duration;
		out_info->looping_policy = context.looping_policy;
		out_info->uses_single_segment = context.uses_single_segment;
		for (int i = 0; i < 2; ++i)
		{
			out_info->key_frame_bit_offsets[i] = context.key_frame_bit_offsets[i];
			out_info->segment_offsets[i] = uint32_t(context.segment_offsets[i]);
			out_info->format_offsets[i] = uint32_t(context.format_per_track_data[i] - base);
			out_info->range_offsets[i] = uint32_t(context.segment_range_data[i] - base);
			out_info->animated_offsets[i] = uint32_t(context.animated_track_data[i] - base);
		}
		return 0;
	}

	// CPU baseline: the reference's own decompression_context<benchmark settings>, requests split statically over
	// `num_threads` threads. Like the reference benchmark (tools/acl_decompressor/sources/benchmark.cpp:246-258) a context
	// stays bound to its clip: a thread re-initialises its context only when the request's clip changes, then seek +
	// decompress_tracks per request. `out` may be null (then every thread writes into a private scratch pose, like the
	// reference benchmark does) or point at [num_requests][max_tracks][12] floats (the parity tests). Returns elapsed seconds
	// of the fastest of `repeats` passes.
	double aclref_bench_transform(const void* const* blobs, const uint32_t* request_clip, const float* request_time,
		uint32_t num_requests, uint32_t max_tracks, uint32_t num_threads, uint32_t repeats, float* out)
	{
		if (num_threads == 0)
			num_threads = 1;

		double best = 1.0e30;
		for (uint32_t repeat = 0; repeat < repeats; ++repeat)
		{
			std::atomic<uint32_t> ready(0);
			std::atomic<bool> go(false);
			std::vector<std::thread> threads;
			threads.reserve(num_threads);

			for (uint32_t thread_index = 0; thread_index < num_threads; ++thread_index)
			{
				threads.emplace_back([=, &ready, &go]()
				{
					const uint64_t begin = uint64_t(num_requests) * thread_index / num_threads;
					const uint64_t end = uint64_t(num_requests) * (thread_index + 1) / num_threads;
					std::vector<float> scratch(size_t(max_tracks) * 12 + 16);

					ready.fetch_add(1);
					while (!go.load(std::memory_order_acquire)) {}

					decompression_context<settings_benchmark> context;
					pose_writer_legacy writer;
					for (uint64_t request = begin; request < end; ++request)
					{
						const compressed_tracks* tracks = static_cast<const compressed_tracks*>(blobs[request_clip[request]]);
						if (context.get_compressed_tracks() != tracks)
							context.initialize(*tracks);
						context.seek(request_time[request], sample_rounding_policy::none);
						writer.out = out != nullptr ? out + request * size_t(max_tracks) * 12 : scratch.data();
						context.decompress_tracks(writer);
					}
				});
			}

			while (ready.load() != num_threads) {}
			const auto start = std::chrono::steady_clock::now();
			go.store(true, std::memory_order_release);
			for (std::thread& thread : threads)
				thread.join();
			const auto stop = std::chrono::steady_clock::now();
			const double seconds = std::chrono::duration<double>(stop - start).count();
			if (seconds < best)
				best = seconds;
		}
		return best;
	}

	double aclref_bench_scalar(const void* const* blobs, const uint32_t* request_clip, const float* request_time,
		uint32_t num_requests, uint32_t max_tracks, uint32_t num_threads, uint32_t repeats, float* out)
	{
		if (num_threads == 0)
			num_threads = 1;

		double best = 1.0e30;
		for (uint32_t repeat = 0; repeat < repeats; ++repeat)
		{
			std::atomic<uint32_t> ready(0);
			std::atomic<bool> go(false);
			std::vector<std::thread> threads;
			for (uint32_t thread_index = 0; thread_index < num_threads; ++thread_index)
			{
				threads.emplace_back([=, &ready, &go]()
				{
					const uint64_t begin = uint64_t(num_requests) * thread_index / num_threads;
					const uint64_t end = uint64_t(num_requests) * (thread_index + 1) / num_threads;
					std::vector<float> scratch(size_t(max_tracks) * 4 + 16);

					ready.fetch_add(1);
					while (!go.load(std::memory_order_acquire)) {}

					decompression_context<default_scalar_decompression_settings> context;
					scalar_writer writer;
					for (uint64_t request = begin; request < end; ++request)
					{
						const compressed_tracks* tracks = static_cast<const compressed_tracks*>(blobs[request_clip[request]]);
						if (context.get_compressed_tracks() != tracks)
							context.initialize(*tracks);
						context.seek(request_time[request], sample_rounding_policy::none);
						writer.out = out != nullptr ? out + request * size_t(max_tracks) * 4 : scratch.data();
						context.decompress_tracks(writer);
					}
				});
			}

			while (ready.load() != num_threads) {}
			const auto start = std::chrono::steady_clock::now();
			go.store(true, std::memory_order_release);
			for (std::thread& thread : threads)
				thread.join();
			const auto stop = std::chrono::steady_clock::now();
			const double seconds = std::chrono::duration<double>(stop - start).count();
			if (seconds < best)
				best = seconds;
		}
		return best;
	}

	// Compress many transform clips in parallel (seed = spec.seed + clip index). Blobs are written
	// back to back into `out_buffer` at 64 byte aligned offsets; returns the number of bytes used, or 0 on
	// failure / overflow. `out_offsets` and `out_sizes` have `num_clips` entries.
	uint64_t aclref_compress_transform_batch(const aclref_transform_spec* spec, uint32_t num_clips, uint32_t num_threads,
		uint8_t* out_buffer, uint64_t buffer_size, uint64_t* out_offsets, uint32_t* out_sizes)
	{
		if (num_threads == 0)
			num_threads = std::max(1u, std::thread::hardware_concurrency());

		std::vector<void*> blobs(num_clips, nullptr);
		std::atomic<uint32_t> next(0);
		std::atomic<bool> failed(false);
		std::vector<std::thread> threads;
		for (uint32_t thread_index = 0; thread_index < num_threads; ++thread_index)
		{
			threads.emplace_back([&]()
			{
				for (;;)
				{
					const uint32_t clip = next.fetch_add(1);
					if (clip >= num_clips)
						break;
					aclref_transform_spec clip_spec = *spec;
					clip_spec.seed = spec->seed + clip;
					void* blob = nullptr;
					uint32_t size = 0;
					if (aclref_compress_transform(&clip_spec, &blob, &size) != 0)
					{
						failed.store(true);
						break;
					}
					blobs[clip] = blob;
					out_sizes[clip] = size;
				}
			});
		}
		for (std::thread& thread : threads)
			thread.join();

		uint64_t offset = 0;
		bool ok = !failed.load();
		for (uint32_t clip = 0; clip < num_clips; ++clip)
		{
			if (ok && blobs[clip] != nullptr)
			{
				const uint64_t aligned_size = (uint64_t(out_sizes[clip]) + 63) & ~uint64_t(63);
				if (offset + aligned_size + 64 > buffer_size)
					ok = false;
				else
				{
					std::memcpy(out_buffer + offset, blobs[clip], out_sizes[clip]);
					std::memset(out_buffer + offset + out_sizes[clip], 0, size_t(aligned_size - out_sizes[clip]));
					out_offsets[clip] = offset;
					offset += aligned_size;
				}
			}
			free(blobs[clip]);
		}
		return ok ? offset : 0;
	}
	//////////////////////////////////////////////////////////////////////////
	// calculate_compression_error

	// acl::track_error (compression/track_error.h:48-66)
	struct aclref_track_error
	{
		uint32_t index;
		float    error;
		float    sample_time;
	};

	// The unmodified calculate_compression_error(allocator, raw_tracks, context, qvvf_transform_error_metric)
	// (track_error.impl.h:465-571 -> calculate_transform_track_error :225-392) on the raw clip of `spec` against `blob` (what
	// aclref_compress_transform(spec) returned), with decompression_context<settings_kind 0 default / 1 debug>. Every other
	// output is optional and exposes what that function works on, recomputed with the reference's own classes the way its loop does:
	//   out_rounding    the rounding policy it seeks with (nearest, or none for stripped / database clips, :556-559)
	//   out_raw_poses   [num_samples][num_tracks][12]  raw_tracks.sample_tracks(sample_time, rounding, writer)
	//   out_lossy_poses [num_samples][num_tracks][12]  seek + decompress_tracks into a debug_track_writer initialised with the bind pose
	//   out_object_poses[2][num_samples][num_tracks][12] error_metric.local_to_object_space of both
	//   out_errors      [num_samples][num_tracks]      error_metric.calculate_error per bone
	//   out_parents / out_shell_distances [num_tracks] track_desc_transformf::parent_index / shell_distance
}	// extern "C"

namespace
{
	template<class settings_type>
	int transform_error_impl(const aclref_transform_spec& spec, const compressed_tracks& tracks, aclref_track_error* out_error, uint32_t* out_rounding,
		float* out_raw_poses, float* out_lossy_poses, float* out_object_poses, float* out_errors, uint32_t* out_parents, float* out_shell_distances)
	{
		iallocator& alloc = allocator();
		track_array_qvvf raw_tracks(alloc, spec.num_tracks);
		make_transform_tracks(spec, raw_tracks);

		decompression_context<settings_type> context;
		if (!context.initialize(tracks))
			return -1;

		const qvvf_transform_error_metric error_metric;
		const track_error result = calculate_compression_error(alloc, raw_tracks, context, error_metric);
		out_error->index = result.index;
		out_error->error = result.error;
		out_error->sample_time = result.sample_time;

		const sample_rounding_policy rounding = (tracks.has_database() || tracks.has_stripped_keyframes()) ? sample_rounding_policy::none : sample_rounding_policy::nearest;
		if (out_rounding != nullptr)
			*out_rounding = uint32_t(rounding);

		const uint32_t num_tracks = raw_tracks.get_num_tracks();
		const uint32_t num_samples = raw_tracks.get_num_samples_per_track();
		const float sample_rate = raw_tracks.get_sample_rate();
		const float duration = raw_tracks.get_finite_duration();

		std::vector<uint32_t> parents(num_tracks), self(num_tracks);
		for (uint32_t bone = 0; bone < num_tracks; ++bone)
		{
			const track_desc_transformf& desc = raw_tracks[bone].get_description();
			parents[bone] = desc.parent_index;
			self[bone] = bone;
			if (out_parents != nullptr) out_parents[bone] = desc.parent_index;
			if (out_shell_distances != nullptr) out_shell_distances[bone] = desc.shell_distance;
		}

		acl_impl::debug_track_writer raw_writer(alloc, track_type8::qvvf, num_tracks);
		acl_impl::debug_track_writer lossy_writer(alloc, track_type8::qvvf, num_tracks);
		lossy_writer.initialize_with_defaults(raw_tracks);
		std::vector<rtm::qvvf> raw_object(num_tracks), lossy_object(num_tracks);

		itransform_error_metric::local_to_object_space_args object_args;
		object_args.dirty_transform_indices = self.data();
		object_args.num_dirty_transforms = num_tracks;
		object_args.parent_transform_indices = parents.data();
		object_args.num_transforms = num_tracks;

		for (uint32_t sample = 0; sample < num_samples; ++sample)
		{
			const float sample_time = rtm::scalar_min(float(sample) / sample_rate, duration);
			raw_tracks.sample_tracks(sample_time, rounding, raw_writer);
			context.seek(sample_time, rounding);
			context.decompress_tracks(lossy_writer);

			object_args.local_transforms = raw_writer.tracks_typed.qvvf;
			error_metric.local_to_object_space(object_args, raw_object.data());
			object_args.local_transforms = lossy_writer.tracks_typed.qvvf;
			error_metric.local_to_object_space(object_args, lossy_object.data());

			const size_t pose_floats = size_t(num_tracks) * 12;
			if (out_raw_poses != nullptr) std::memcpy(out_raw_poses + sample * pose_floats, raw_writer.tracks_typed.qvvf, pose_floats * sizeof(float));
			if (out_lossy_poses != nullptr) std::memcpy(out_lossy_poses + sample * pose_floats, lossy_writer.tracks_typed.qvvf, pose_floats * sizeof(float));
			if (out_object_poses != nullptr)
			{
				std::memcpy(out_object_poses + sample * pose_floats, raw_object.data(), pose_floats * sizeof(float));
				std::memcpy(out_object_poses + (size_t(num_samples) + sample) * pose_floats, lossy_object.data(), pose_floats * sizeof(float));
			}
			if (out_errors != nullptr)
				for (uint32_t bone = 0; bone < num_tracks; ++bone)
				{
					itransform_error_metric::calculate_error_args error_args;
					error_args.transform0 = &raw_object[bone];
					error_args.transform1 = &lossy_object[bone];
					error_args.construct_sphere_shell(raw_tracks[bone].get_description().shell_distance);
					out_errors[size_t(sample) * num_tracks + bone] = rtm::scalar_cast(error_metric.calculate_error(error_args));
				}
		}
		return 0;
	}

}

extern "C"
{
	int aclref_transform_error(const aclref_transform_spec* spec, const void* blob, uint32_t settings_kind, aclref_track_error* out_error, uint32_t* out_rounding,
		float* out_raw_poses, float* out_lossy_poses, float* out_object_poses, float* out_errors, uint32_t* out_parents, float* out_shell_distances)
	{
		const compressed_tracks& tracks = *static_cast<const compressed_tracks*>(blob);
		if (tracks.get_track_type() != track_type8::qvvf || tracks.get_num_tracks() != spec->num_tracks)
			return -2;
		if (settings_kind == 0)
			return transform_error_impl<settings_default>(*spec, tracks, out_error, out_rounding, out_raw_poses, out_lossy_poses, out_object_poses, out_errors, out_parents, out_shell_distances);
		return transform_error_impl<settings_debug>(*spec, tracks, out_error, out_rounding, out_raw_poses, out_lossy_poses, out_object_poses, out_errors, out_parents, out_shell_distances);
	}

	// The additive flavour: calculate_compression_error(allocator, raw_tracks, context, error_metric, additive_base_tracks)
	// (track_error.impl.h:573-680) with additive_qvvf_transform_error_metric<format> (transform_error_metrics.h:470-526). The raw clip of `spec`
	// plays the additive clip, the raw clip of `base_spec` (same track count, any sample count) the base it applies to; `blob` is what
	// aclref_compress_transform(spec) returned. additive_format: acl::additive_clip_format8 (1 relative, 2 additive0, 3 additive1).
	// Outputs as aclref_transform_error, plus out_base_poses [num_samples][num_tracks][12] = additive_base_tracks.sample_tracks at the
	// base clip's matching time (track_error.impl.h:352-356); out_errors are measured after apply_additive_to_base on both poses.
	int aclref_transform_error_additive(const aclref_transform_spec* spec, const void* blob, const aclref_transform_spec* base_spec, uint32_t additive_format,
		aclref_track_error* out_error, uint32_t* out_rounding, float* out_raw_poses, float* out_lossy_poses, float* out_base_poses, float* out_errors,
		uint32_t* out_parents, float* out_shell_distances)
	{
		const compressed_tracks& tracks = *static_cast<const compressed_tracks*>(blob);
		if (tracks.get_track_type() != track_type8::qvvf || tracks.get_num_tracks() != spec->num_tracks || base_spec->num_tracks != spec->num_tracks)
			return -2;
		iallocator& alloc = allocator();
		track_array_qvvf raw_tracks(alloc, spec->num_tracks);
		make_transform_tracks(*spec, raw_tracks);
		track_array_qvvf base_tracks(alloc, base_spec->num_tracks);
		make_transform_tracks(*base_spec, base_tracks);

		decompression_context<settings_debug> context;
		if (!context.initialize(tracks))
			return -1;

		const additive_qvvf_transform_error_metric<additive_clip_format8::relative> metric_relative;
		const additive_qvvf_transform_error_metric<additive_clip_format8::additive0> metric_additive0;
		const additive_qvvf_transform_error_metric<additive_clip_format8::additive1> metric_additive1;
		const itransform_error_metric* metric = nullptr;
		switch (static_cast<additive_clip_format8>(additive_format))
		{
		case additive_clip_format8::relative: metric = &metric_relative; break;
		case additive_clip_format8::additive0: metric = &metric_additive0; break;
		case additive_clip_format8::additive1: metric = &metric_additive1; break;
		default: return -3;
		}
		const itransform_error_metric& error_metric = *metric;

		const track_error result = calculate_compression_error(alloc, raw_tracks, context, error_metric, base_tracks);
		out_error->index = result.index;
		out_error->error = result.error;
		out_error->sample_time = result.sample_time;

		const sample_rounding_policy rounding = (tracks.has_database() || tracks.has_stripped_keyframes()) ? sample_rounding_policy::none : sample_rounding_policy::nearest;
		if (out_rounding != nullptr)
			*out_rounding = uint32_t(rounding);

		const uint32_t num_tracks = raw_tracks.get_num_tracks();
		const uint32_t num_samples = raw_tracks.get_num_samples_per_track();
		const float sample_rate = raw_tracks.get_sample_rate();
		const float duration = raw_tracks.get_finite_duration();
		const uint32_t base_num_samples = base_tracks.get_num_samples_per_track();
		const float base_duration = base_tracks.get_finite_duration();

		std::vector<uint32_t> parents(num_tracks), self(num_tracks);
		for (uint32_t bone = 0; bone < num_tracks; ++bone)
		{
			const track_desc_transformf& desc = raw_tracks[bone].get_description();
			parents[bone] = desc.parent_index;
			self[bone] = bone;
			if (out_parents != nullptr) out_parents[bone] = desc.parent_index;
			if (out_shell_distances != nullptr) out_shell_distances[bone] = desc.shell_distance;
		}

		acl_impl::debug_track_writer raw_writer(alloc, track_type8::qvvf, num_tracks);
		acl_impl::debug_track_writer lossy_writer(alloc, track_type8::qvvf, num_tracks);
		acl_impl::debug_track_writer base_writer(alloc, track_type8::qvvf, num_tracks);
		lossy_writer.initialize_with_defaults(raw_tracks);
		std::vector<rtm::qvvf> raw_object(num_tracks), lossy_object(num_tracks);

		itransform_error_metric::apply_additive_to_base_args additive_args;
		additive_args.dirty_transform_indices = self.data();
		additive_args.num_dirty_transforms = num_tracks;
		additive_args.base_transforms = base_writer.tracks_typed.qvvf;
		additive_args.num_transforms = num_tracks;
		itransform_error_metric::local_to_object_space_args object_args;
		object_args.dirty_transform_indices = self.data();
		object_args.num_dirty_transforms = num_tracks;
		object_args.parent_transform_indices = parents.data();
		object_args.num_transforms = num_tracks;

		for (uint32_t sample = 0; sample < num_samples; ++sample)
		{
			const float sample_time = rtm::scalar_min(float(sample) / sample_rate, duration);
			raw_tracks.sample_tracks(sample_time, rounding, raw_writer);
			context.seek(sample_time, rounding);
			context.decompress_tracks(lossy_writer);
			const float normalized_sample_time = base_num_samples > 1 ? (sample_time / duration) : 0.0F;		// track_error.impl.h:352-353
			const float additive_sample_time = base_num_samples > 1 ? (normalized_sample_time * base_duration) : 0.0F;
			base_tracks.sample_tracks(additive_sample_time, rounding, base_writer);

			const size_t pose_floats = size_t(num_tracks) * 12;
			if (out_raw_poses != nullptr) std::memcpy(out_raw_poses + sample * pose_floats, raw_writer.tracks_typed.qvvf, pose_floats * sizeof(float));
			if (out_lossy_poses != nullptr) std::memcpy(out_lossy_poses + sample * pose_floats, lossy_writer.tracks_typed.qvvf, pose_floats * sizeof(float));
			if (out_base_poses != nullptr) std::memcpy(out_base_poses + sample * pose_floats, base_writer.tracks_typed.qvvf, pose_floats * sizeof(float));

			std::vector<rtm::qvvf> raw_local(raw_writer.tracks_typed.qvvf, raw_writer.tracks_typed.qvvf + num_tracks);
			std::vector<rtm::qvvf> lossy_local(lossy_writer.tracks_typed.qvvf, lossy_writer.tracks_typed.qvvf + num_tracks);
			additive_args.local_transforms = raw_local.data();
			error_metric.apply_additive_to_base(additive_args, raw_local.data());
			additive_args.local_transforms = lossy_local.data();
			error_metric.apply_additive_to_base(additive_args, lossy_local.data());
			object_args.local_transforms = raw_local.data();
			error_metric.local_to_object_space(object_args, raw_object.data());
			object_args.local_transforms = lossy_local.data();
			error_metric.local_to_object_space(object_args, lossy_object.data());
			if (out_errors != nullptr)
				for (uint32_t bone = 0; bone < num_tracks; ++bone)
				{
					itransform_error_metric::calculate_error_args error_args;
					error_args.transform0 = &raw_object[bone];
					error_args.transform1 = &lossy_object[bone];
					error_args.construct_sphere_shell(raw_tracks[bone].get_description().shell_distance);
					out_errors[size_t(sample) * num_tracks + bone] = rtm::scalar_cast(error_metric.calculate_error(error_args));
				}
		}
		return 0;
	}

	// calculate_compression_error with the reference's other metric, qvvf_matrix3x4f_transform_error_metric (transform_error_metrics.h:389-464:
	// transforms converted to 3x4 matrices, object space by matrix_mul, shell points through matrix_mul_point3). Same raw / decoded poses as
	// aclref_transform_error (debug settings); out_errors [num_samples][num_tracks] replayed with the metric's own functions.
	int aclref_transform_error_matrix(const aclref_transform_spec* spec, const void* blob, aclref_track_error* out_error, float* out_errors)
	{
		const compressed_tracks& tracks = *static_cast<const compressed_tracks*>(blob);
		if (tracks.get_track_type() != track_type8::qvvf || tracks.get_num_tracks() != spec->num_tracks)
			return -2;
		iallocator& alloc = allocator();
		track_array_qvvf raw_tracks(alloc, spec->num_tracks);
		make_transform_tracks(*spec, raw_tracks);
		decompression_context<settings_debug> context;
		if (!context.initialize(tracks))
			return -1;

		const qvvf_matrix3x4f_transform_error_metric error_metric;
		const track_error result = calculate_compression_error(alloc, raw_tracks, context, error_metric);
		out_error->index = result.index;
		out_error->error = result.error;
		out_error->sample_time = result.sample_time;
		if (out_errors == nullptr)
			return 0;

		const sample_rounding_policy rounding = (tracks.has_database() || tracks.has_stripped_keyframes()) ? sample_rounding_policy::none : sample_rounding_policy::nearest;
		const uint32_t num_tracks = raw_tracks.get_num_tracks();
		const uint32_t num_samples = raw_tracks.get_num_samples_per_track();
		std::vector<uint32_t> parents(num_tracks), self(num_tracks);
		for (uint32_t bone = 0; bone < num_tracks; ++bone)
		{
			parents[bone] = raw_tracks[bone].get_description().parent_index;
			self[bone] = bone;
		}
		acl_impl::debug_track_writer raw_writer(alloc, track_type8::qvvf, num_tracks);
		acl_impl::debug_track_writer lossy_writer(alloc, track_type8::qvvf, num_tracks);
		lossy_writer.initialize_with_defaults(raw_tracks);
		std::vector<rtm::matrix3x4f> raw_local(num_tracks), lossy_local(num_tracks), raw_object(num_tracks), lossy_object(num_tracks);

		itransform_error_metric::convert_transforms_args convert_args;
		convert_args.dirty_transform_indices = self.data();
		convert_args.num_dirty_transforms = num_tracks;
		convert_args.num_transforms = num_tracks;
		convert_args.is_additive_base = false;
		itransform_error_metric::local_to_object_space_args object_args;
		object_args.dirty_transform_indices = self.data();
		object_args.num_dirty_transforms = num_tracks;
		object_args.parent_transform_indices = parents.data();
		object_args.num_transforms = num_tracks;
		for (uint32_t sample = 0; sample < num_samples; ++sample)
		{
			const float sample_time = rtm::scalar_min(float(sample) / raw_tracks.get_sample_rate(), raw_tracks.get_finite_duration());
			raw_tracks.sample_tracks(sample_time, rounding, raw_writer);
			context.seek(sample_time, rounding);
			context.decompress_tracks(lossy_writer);
			convert_args.sample_index = sample;
			convert_args.transforms = raw_writer.tracks_typed.qvvf;
			convert_args.is_lossy = false;
			error_metric.convert_transforms(convert_args, raw_local.data());
			convert_args.transforms = lossy_writer.tracks_typed.qvvf;
			convert_args.is_lossy = true;
			error_metric.convert_transforms(convert_args, lossy_local.data());
			object_args.local_transforms = raw_local.data();
			error_metric.local_to_object_space(object_args, raw_object.data());
			object_args.local_transforms = lossy_local.data();
			error_metric.local_to_object_space(object_args, lossy_object.data());
			for (uint32_t bone = 0; bone < num_tracks; ++bone)
			{
				itransform_error_metric::calculate_error_args error_args;
				error_args.transform0 = &raw_object[bone];
				error_args.transform1 = &lossy_object[bone];
				error_args.construct_sphere_shell(raw_tracks[bone].get_description().shell_distance);
				out_errors[size_t(sample) * num_tracks + bone] = rtm::scalar_cast(error_metric.calculate_error(error_args));
			}
		}
		return 0;
	}

	//////////////////////////////////////////////////////////////////////////
	// SURVEY 8(f2), first step: clips bound to a streaming database, decoded from the key frames that stay resident in the clip

	// Compresses the raw clip of `spec` with database support (compress_track_list with enable_database_support, compress.transform.impl.h:158-166)
	// and splits it with acl::build_database (compress.h:86-100; `medium` / `low` = compression_database_settings proportions): the returned
	// blob is the compressed_tracks instance BOUND TO THE DATABASE (has_database, the movable key frames moved out). The database itself is
	// dropped: these clips are decoded without it.
	int aclref_compress_transform_database(const aclref_transform_spec* spec, float medium_proportion, float low_proportion, void** out_blob, uint32_t* out_size)
	{
		iallocator& alloc = allocator();
		track_array_qvvf track_list(alloc, spec->num_tracks);
		make_transform_tracks(*spec, track_list);

		qvvf_transform_error_metric error_metric;
		compression_settings settings;
		settings.level = static_cast<compression_level8>(spec->level);
		settings.rotation_format = static_cast<rotation_format8>(spec->rotation_format);
		settings.translation_format = static_cast<vector_format8>(spec->translation_format);
		settings.scale_format = static_cast<vector_format8>(spec->scale_format);
		settings.error_metric = &error_metric;
		settings.optimize_loops = spec->optimize_loops != 0;
		settings.enable_database_support = true;

		compressed_tracks* tracks = nullptr;
		output_stats stats;
		error_result result = compress_track_list(alloc, track_list, settings, tracks, stats);
		if (result.any() || tracks == nullptr)
		{
			fprintf(stderr, "aclref_compress_transform_database: %s\n", result.any() ? result.c_str() : "no output");
			return -1;
		}

		compression_database_settings database_settings;
		database_settings.medium_importance_tier_proportion = medium_proportion;
		database_settings.low_importance_tier_proportion = low_proportion;
		const compressed_tracks* input_list[1] = { tracks };
		compressed_tracks* bound_list[1] = { nullptr };
		compressed_database* database = nullptr;
		result = build_database(alloc, database_settings, input_list, 1, bound_list, database);
		alloc.deallocate(tracks, tracks->get_size());
		if (result.any() || bound_list[0] == nullptr)
		{
			fprintf(stderr, "aclref_compress_transform_database: build_database: %s\n", result.any() ? result.c_str() : "no output");
			return -2;
		}
		if (database != nullptr)
			alloc.deallocate(database, database->get_size());

		const uint32_t size = bound_list[0]->get_size();
		void* copy = nullptr;
		if (posix_memalign(&copy, 64, size + 64) != 0)
			return -3;
		std::memcpy(copy, bound_list[0], size);
		std::memset(static_cast<uint8_t*>(copy) + size, 0, 64);
		alloc.deallocate(bound_list[0], size);
		*out_blob = copy;
		*out_size = size;
		return 0;
	}

	// decompression_context<settings with database support>::initialize(tracks) -- NO database bound (decompress.impl.h:67-83: database =
	// nullptr) -- then seek + decompress_tracks: what a database clip gives before any tier is streamed in. writer_mode as aclref_decompress_tracks.
	int aclref_decompress_tracks_without_database(const void* blob, float sample_time, uint32_t rounding, uint32_t looping, uint32_t writer_mode, float* out)
	{
		decode_args args;
		args.tracks = static_cast<const compressed_tracks*>(blob);
		args.sample_time = sample_time;
		args.rounding = static_cast<sample_rounding_policy>(rounding);
		args.looping = static_cast<sample_looping_policy>(looping);
		args.track_index = -1;
		args.out = out;
		args.per_track_rounding = nullptr;
		args.constant_defaults = nullptr;
		args.variable_defaults = nullptr;
		return decode_dispatch_writer<settings_database>(writer_mode, args);
	}

	// The scalar flavour: calculate_compression_error(allocator, raw_tracks, context) (track_error.impl.h:400-463 -> calculate_scalar_track_error
	// :166-223). out_raw_values [num_samples][num_tracks][4] = raw_tracks.sample_tracks(...) (first N components of each row).
	int aclref_scalar_error(const aclref_scalar_spec* spec, const void* blob, aclref_track_error* out_error, uint32_t* out_rounding, float* out_raw_values)
	{
		iallocator& alloc = allocator();
		const compressed_tracks& tracks = *static_cast<const compressed_tracks*>(blob);
		if (tracks.get_track_type() == track_type8::qvvf || tracks.get_num_tracks() != spec->num_tracks)
			return -2;

		track_array raw_tracks;
		switch (static_cast<track_type8>(spec->track_type))
		{
		case track_type8::float1f: { track_array_float1f list(alloc, spec->num_tracks); fill_scalar_tracks<track_float1f, float, 1>(*spec, list); raw_tracks = std::move(list); break; }
		case track_type8::float2f: { track_array_float2f list(alloc, spec->num_tracks); fill_scalar_tracks<track_float2f, rtm::float2f, 2>(*spec, list); raw_tracks = std::move(list); break; }
		case track_type8::float3f: { track_array_float3f list(alloc, spec->num_tracks); fill_scalar_tracks<track_float3f, rtm::float3f, 3>(*spec, list); raw_tracks = std::move(list); break; }
		case track_type8::float4f: { track_array_float4f list(alloc, spec->num_tracks); fill_scalar_tracks<track_float4f, rtm::float4f, 4>(*spec, list); raw_tracks = std::move(list); break; }
		case track_type8::vector4f: { track_array_vector4f list(alloc, spec->num_tracks); fill_scalar_tracks<track_vector4f, rtm::vector4f, 4>(*spec, list); raw_tracks = std::move(list); break; }
		default: return -3;
		}

		decompression_context<default_scalar_decompression_settings> context;
		if (!context.initialize(tracks))
			return -1;
		const track_error result = calculate_compression_error(alloc, raw_tracks, context);
		out_error->index = result.index;
		out_error->error = result.error;
		out_error->sample_time = result.sample_time;

		const sample_rounding_policy rounding = (tracks.has_database() || tracks.has_stripped_keyframes()) ? sample_rounding_policy::none : sample_rounding_policy::nearest;
		if (out_rounding != nullptr)
			*out_rounding = uint32_t(rounding);
		if (out_raw_values != nullptr)
		{
			const uint32_t num_tracks = raw_tracks.get_num_tracks();
			const uint32_t num_samples = raw_tracks.get_num_samples_per_track();
			scalar_writer writer;
			for (uint32_t sample = 0; sample < num_samples; ++sample)
			{
				const float sample_time = rtm::scalar_min(float(sample) / raw_tracks.get_sample_rate(), raw_tracks.get_finite_duration());
				writer.out = out_raw_values + size_t(sample) * num_tracks * 4;
				raw_tracks.sample_tracks(sample_time, rounding, writer);
			}
		}
		return 0;
	}

	// The raw poses calculate_compression_error would sample for clips spec.seed .. spec.seed + num_clips - 1 (raw_tracks.sample_tracks at
	// min(i / sample_rate, duration) with the `nearest` policy, track_error.impl.h:337-338; `none` gives the same poses at those times up to
	// the interpolation's rounding, callers with stripped clips use aclref_transform_error instead), clips dealt to `num_threads` threads:
	// out_raw_poses [num_clips][num_samples][num_tracks][12]. Also the skeleton of the (shared) rig: out_parents / out_shell_distances [num_tracks].
	int aclref_sample_raw_transform_batch(const aclref_transform_spec* spec, uint32_t num_clips, uint32_t num_threads, float* out_raw_poses,
		uint32_t* out_parents, float* out_shell_distances)
	{
		if (num_threads == 0)
			num_threads = 1;
		std::atomic<uint32_t> next(0);
		std::vector<std::thread> threads;
		for (uint32_t thread_index = 0; thread_index < num_threads; ++thread_index)
		{
			threads.emplace_back([&]()
			{
				iallocator& alloc = allocator();
				acl_impl::debug_track_writer writer(alloc, track_type8::qvvf, spec->num_tracks);
				for (;;)
				{
					const uint32_t clip = next.fetch_add(1);
					if (clip >= num_clips)
						break;
					aclref_transform_spec clip_spec = *spec;
					clip_spec.seed = spec->seed + clip;
					track_array_qvvf raw_tracks(alloc, clip_spec.num_tracks);
					make_transform_tracks(clip_spec, raw_tracks);
					const uint32_t num_samples = raw_tracks.get_num_samples_per_track();
					const size_t pose_floats = size_t(clip_spec.num_tracks) * 12;
					for (uint32_t sample = 0; sample < num_samples; ++sample)
					{
						const float sample_time = rtm::scalar_min(float(sample) / raw_tracks.get_sample_rate(), raw_tracks.get_finite_duration());
						raw_tracks.sample_tracks(sample_time, sample_rounding_policy::nearest, writer);
						std::memcpy(out_raw_poses + (size_t(clip) * num_samples + sample) * pose_floats, writer.tracks_typed.qvvf, pose_floats * sizeof(float));
					}
					if (clip == 0)
						for (uint32_t bone = 0; bone < clip_spec.num_tracks; ++bone)
						{
							const track_desc_transformf& desc = raw_tracks[bone].get_description();
							if (out_parents != nullptr) out_parents[bone] = desc.parent_index;
							if (out_shell_distances != nullptr) out_shell_distances[bone] = desc.shell_distance;
						}
				}
			});
		}
		for (std::thread& thread : threads)
			thread.join();
		return 0;
	}

	// CPU baseline of the 8(f1) workload: calculate_compression_error of clips spec.seed .. spec.seed + num_clips - 1 (their blobs in `blobs`),
	// clips dealt to `num_threads` threads. The raw clips are rebuilt before the clock starts; only the error measurement is timed.
	// Returns elapsed seconds; out_errors (optional) receives one aclref_track_error per clip.
	double aclref_bench_transform_error(const aclref_transform_spec* spec, const void* const* blobs, uint32_t num_clips, uint32_t num_threads, aclref_track_error* out_errors)
	{
		if (num_threads == 0)
			num_threads = 1;
		iallocator& alloc = allocator();
		std::vector<track_array_qvvf> raw_clips;
		raw_clips.reserve(num_clips);
		for (uint32_t clip = 0; clip < num_clips; ++clip)
		{
			aclref_transform_spec clip_spec = *spec;
			clip_spec.seed = spec->seed + clip;
			raw_clips.emplace_back(alloc, clip_spec.num_tracks);
			make_transform_tracks(clip_spec, raw_clips.back());
		}

		std::atomic<uint32_t> next(0);
		std::vector<std::thread> threads;
		const auto start = std::chrono::steady_clock::now();
		for (uint32_t thread_index = 0; thread_index < num_threads; ++thread_index)
		{
			threads.emplace_back([&]()
			{
				const qvvf_transform_error_metric error_metric;
				for (;;)
				{
					const uint32_t clip = next.fetch_add(1);
					if (clip >= num_clips)
						break;
					decompression_context<settings_debug> context;
					if (!context.initialize(*static_cast<const compressed_tracks*>(blobs[clip])))
						continue;
					const track_error result = calculate_compression_error(alloc, raw_clips[clip], context, error_metric);
					if (out_errors != nullptr)
					{
						out_errors[clip].index = result.index;
						out_errors[clip].error = result.error;
						out_errors[clip].sample_time = result.sample_time;
					}
				}
			});
		}
		for (std::thread& thread : threads)
			thread.join();
		return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
	}
}
