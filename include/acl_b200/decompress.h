// include/acl_b200/decompress.h -- C++ header shim over the C ABI of libaclb200 (include/aclb200.h).
//
// It keeps the names and the call sequence of the reference's decompression front end so that call sites read the same:
//
//   reference (includes/acl/decompression/decompress.h:90-172)          this header
//   ------------------------------------------------------------------  -------------------------------------------
//   acl::decompression_context<settings> context;                       acl_b200::decompression_context<settings> context(device);
//   context.initialize(*tracks)                                         context.initialize(blob, size)
//   context.is_bound_to(*tracks) / is_initialized()                     same
//   context.set_looping_policy(policy)                                  same
//   context.seek(sample_time, rounding_policy)                          same
//   context.decompress_tracks(writer)                                   same (writer: the track_writer concept below)
//   context.decompress_track(track_index, writer)                       same
//
// `settings` is any type with the static constexpr members of acl::decompression_settings
// (decompression_settings.h:74-166); `writer` any type with the members of acl::track_writer (core/track_writer.h:82-216).
// Both are duck-typed, so the reference's own settings / writer classes work once their rtm argument types are constructible
// from acl_b200::float4 (see INTEGRATION.md for the two-line adapter).
//
// A decompression_context decodes ONE pose per call through the GPU, which costs a launch and a PCIe round trip: it exists for
// drop-in compatibility and for tests. Throughput comes from acl_b200::batch_decompressor below: upload the clips once, then decode
// thousands of (clip, sample_time) requests per launch into device memory.
//
// Nothing here decodes on the CPU: without the library or without a B200 every call fails with a status, never silently.
#pragma once

#include "../aclb200.h"

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace acl_b200
{
	// acl::sample_rounding_policy (core/interpolation_utils.h:40-70)
	enum class sample_rounding_policy : uint32_t { none = ACLB200_ROUND_NONE, floor = ACLB200_ROUND_FLOOR, ceil = ACLB200_ROUND_CEIL, nearest = ACLB200_ROUND_NEAREST, per_track = ACLB200_ROUND_PER_TRACK };
	// acl::sample_looping_policy (core/sample_looping_policy.h:40-66)
	enum class sample_looping_policy : uint32_t { clamp = ACLB200_LOOP_CLAMP, wrap = ACLB200_LOOP_WRAP, as_compressed = ACLB200_LOOP_AS_COMPRESSED };
	// acl::rotation_normalization_policy_t (decompression/decompression_settings.h:48-70)
	enum class rotation_normalization_policy_t : uint32_t { never = ACLB200_NORMALIZE_NEVER, lerp_only = ACLB200_NORMALIZE_LERP_ONLY, always = ACLB200_NORMALIZE_ALWAYS };
	// acl::default_sub_track_mode (core/track_writer.h:49-80)
	enum class default_sub_track_mode : uint32_t { skipped = ACLB200_DEFAULT_SKIPPED, constant = ACLB200_DEFAULT_CONSTANT, variable = ACLB200_DEFAULT_VARIABLE, legacy = ACLB200_DEFAULT_LEGACY };

	// What a writer receives: four floats (rotations xyzw; translations / scales xyz, w unspecified like in the reference)
	struct float4
	{
		float x, y, z, w;
	};

	// acl::decompression_settings (decompression_settings.h:74-166), same member names and defaults
	struct decompression_settings
	{
		static constexpr bool clamp_sample_time() { return true; }
		static constexpr rotation_normalization_policy_t get_rotation_normalization_policy() { return rotation_normalization_policy_t::always; }
		static constexpr bool is_wrapping_supported() { return true; }
		static constexpr bool is_per_track_rounding_supported() { return true; }
		// more than one rotation format compiled in (debug settings): see aclb200_options::multiple_rotation_formats
		static constexpr bool supports_multiple_rotation_formats() { return true; }
	};
	// acl::default_transform_decompression_settings (decompression_settings.h:138-159)
	struct default_transform_decompression_settings : decompression_settings
	{
		static constexpr rotation_normalization_policy_t get_rotation_normalization_policy() { return rotation_normalization_policy_t::lerp_only; }
		static constexpr bool is_per_track_rounding_supported() { return false; }
		static constexpr bool supports_multiple_rotation_formats() { return false; }
	};
	using debug_transform_decompression_settings = decompression_settings;		// decompression_settings.h:110-118

	// acl::track_writer (core/track_writer.h:82-216), same member names and defaults
	struct track_writer
	{
		static constexpr default_sub_track_mode get_default_rotation_mode() { return default_sub_track_mode::constant; }
		static constexpr default_sub_track_mode get_default_translation_mode() { return default_sub_track_mode::constant; }
		static constexpr default_sub_track_mode get_default_scale_mode() { return default_sub_track_mode::legacy; }
		float4 get_constant_default_rotation() const { return float4{ 0.0F, 0.0F, 0.0F, 1.0F }; }
		float4 get_constant_default_translation() const { return float4{ 0.0F, 0.0F, 0.0F, 0.0F }; }
		float4 get_constant_default_scale() const { return float4{ 1.0F, 1.0F, 1.0F, 1.0F }; }
		float4 get_variable_default_rotation(uint32_t) const { return float4{ 0.0F, 0.0F, 0.0F, 1.0F }; }
		float4 get_variable_default_translation(uint32_t) const { return float4{ 0.0F, 0.0F, 0.0F, 0.0F }; }
		float4 get_variable_default_scale(uint32_t) const { return float4{ 1.0F, 1.0F, 1.0F, 1.0F }; }
		sample_rounding_policy get_rounding_policy(sample_rounding_policy policy, uint32_t) const { return policy; }	// per track rounding hook, track_writer.h:100-106
		static constexpr bool skip_all_rotations() { return false; }
		static constexpr bool skip_all_translations() { return false; }
		static constexpr bool skip_all_scales() { return false; }
		static constexpr bool skip_track_rotation(uint32_t) { return false; }
		static constexpr bool skip_track_translation(uint32_t) { return false; }
		static constexpr bool skip_track_scale(uint32_t) { return false; }
		void write_rotation(uint32_t, float4) {}
		void write_translation(uint32_t, float4) {}
		void write_scale(uint32_t, float4) {}
	};

	class error : public std::runtime_error
	{
	public:
		error(aclb200_status status_, const std::string& what_) : std::runtime_error(what_), status(status_) {}
		aclb200_status status;
	};

	// One CUDA device. Owns the library context.
	class device_context
	{
	public:
		explicit device_context(int device = 0)
		{
			const aclb200_status status = aclb200_create(device, &m_context);
			if (status != ACLB200_OK)
				throw error(status, std::string("aclb200_create: ") + aclb200_status_string(status));
		}
		~device_context() { aclb200_destroy(m_context); }
		device_context(const device_context&) = delete;
		device_context& operator=(const device_context&) = delete;

		aclb200_context* get() const { return m_context; }
		void check(aclb200_status status, const char* what) const
		{
			if (status != ACLB200_OK)
				throw error(status, std::string(what) + ": " + aclb200_last_error(m_context));
		}

	private:
		aclb200_context* m_context = nullptr;
	};

	template<class settings_type, class writer_type>
	inline aclb200_options make_options(const writer_type& writer, sample_rounding_policy rounding, sample_looping_policy looping)
	{
		aclb200_options options;
		std::memset(&options, 0, sizeof(options));
		aclb200_default_options(&options);
		options.rounding_policy = static_cast<uint32_t>(rounding);
		options.looping_policy = static_cast<uint32_t>(looping);
		options.normalization = static_cast<uint32_t>(settings_type::get_rotation_normalization_policy());
		options.per_track_rounding = settings_type::is_per_track_rounding_supported() ? 1u : 0u;
		options.wrapping = settings_type::is_wrapping_supported() ? 1u : 0u;
		options.clamp_sample_time = settings_type::clamp_sample_time() ? 1u : 0u;
		options.multiple_rotation_formats = settings_type::supports_multiple_rotation_formats() ? 1u : 0u;
		// `variable` defaults come from writer callbacks: the device leaves those sub-tracks alone (skipped) and the replay asks the
		// writer, which is exactly what the reference does (decompression.transform.h:1566-1650)
		const auto device_mode = [](default_sub_track_mode mode) { return static_cast<uint32_t>(mode == default_sub_track_mode::variable ? default_sub_track_mode::skipped : mode); };
		options.default_rotation_mode = device_mode(writer_type::get_default_rotation_mode());
		options.default_translation_mode = device_mode(writer_type::get_default_translation_mode());
		options.default_scale_mode = device_mode(writer_type::get_default_scale_mode());
		const float4 rotation = writer.get_constant_default_rotation(), translation = writer.get_constant_default_translation(), scale = writer.get_constant_default_scale();
		const float defaults[12] = { rotation.x, rotation.y, rotation.z, rotation.w, translation.x, translation.y, translation.z, 0.0F, scale.x, scale.y, scale.z, 0.0F };
		std::memcpy(options.constant_defaults, defaults, sizeof(defaults));
		options.output_layout = ACLB200_LAYOUT_QVV48;
		return options;
	}

	// The throughput interface: a set of clips resident in HBM + batched decodes into device memory.
	// Replaces N x { context.initialize(clip); context.seek(t, policy); context.decompress_tracks(writer); } by ONE launch.
	class batch_decompressor
	{
	public:
		explicit batch_decompressor(device_context& device) : m_device(device) {}
		~batch_decompressor() { release(); }
		batch_decompressor(const batch_decompressor&) = delete;
		batch_decompressor& operator=(const batch_decompressor&) = delete;

		// compressed_tracks buffers as the reference's compressor wrote them (16 byte alignment not required here).
		// Returns false and reports the offending clip when one is not a valid / supported compressed_tracks instance,
		// which is what decompression_context::initialize() reports by returning false.
		bool upload(const void* const* blobs, const uint32_t* sizes, uint32_t num_clips, bool check_hash = true, uint32_t* out_failed_clip = nullptr)
		{
			release();
			const aclb200_status status = aclb200_upload_clips(m_device.get(), blobs, sizes, num_clips, check_hash ? 1u : 0u, &m_clipset, out_failed_clip);
			if (status == ACLB200_ERR_INVALID_CLIP || status == ACLB200_ERR_UNSUPPORTED)
				return false;
			m_device.check(status, "aclb200_upload_clips");
			m_device.check(aclb200_clipset_get_info(m_clipset, &m_info), "aclb200_clipset_get_info");
			return true;
		}
		void release()
		{
			if (m_clipset != nullptr)
				aclb200_release_clipset(m_device.get(), m_clipset);
			m_clipset = nullptr;
		}

		const aclb200_clipset_info& info() const { return m_info; }
		const aclb200_clipset* clipset() const { return m_clipset; }

		// d_requests / d_out are DEVICE pointers; `stream` a cudaStream_t. Asynchronous like any kernel launch.
		void decompress_tracks(const aclb200_request* d_requests, uint32_t num_requests, const aclb200_options& options, void* d_out, void* stream = nullptr)
		{
			m_device.check(aclb200_decompress_tracks(m_device.get(), m_clipset, d_requests, num_requests, &options, d_out, stream), "aclb200_decompress_tracks");
		}
		void decompress_track(const aclb200_request* d_requests, const uint32_t* d_track_indices, uint32_t num_requests, const aclb200_options& options, void* d_out, void* stream = nullptr)
		{
			m_device.check(aclb200_decompress_track(m_device.get(), m_clipset, d_requests, d_track_indices, num_requests, &options, d_out, stream), "aclb200_decompress_track");
		}
		// host buffers in, host buffers out, synchronous
		void decompress_tracks_host(const aclb200_request* requests, uint32_t num_requests, const aclb200_options& options, void* out, size_t out_bytes)
		{
			m_device.check(aclb200_decompress_tracks_host(m_device.get(), m_clipset, requests, num_requests, &options, out, out_bytes), "aclb200_decompress_tracks_host");
		}

	private:
		device_context& m_device;
		aclb200_clipset* m_clipset = nullptr;
		aclb200_clipset_info m_info = {};
	};

	// Where a batch writes its poses: device memory, request r at d_poses + r * pose_stride_bytes (0 = packed)
	struct device_pose_writer : track_writer
	{
		void* d_poses = nullptr;
		uint32_t output_layout = ACLB200_LAYOUT_QVV48;
		uint64_t pose_stride_bytes = 0;
		void* stream = nullptr;		// cudaStream_t
	};

	// The batched form of decompression_context<settings>: bind N clips, seek N' (clip, time) requests, decode them in one launch.
	// The settings type plays the role it plays in the reference (decompress.h:76-88): it fixes the normalisation policy, wrapping,
	// per track rounding and sample time clamping at compile time.
	template<class settings_type = default_transform_decompression_settings>
	class batch_context
	{
		static_assert(std::is_base_of<decompression_settings, settings_type>::value, "settings_type must derive from decompression_settings");

	public:
		explicit batch_context(device_context& device) : m_batch(device) {}

		// initialize() for every clip of the batch; false (and the offending index) when one is not a valid compressed_tracks
		bool bind(const void* const* compressed_tracks, const uint32_t* sizes, uint32_t num_clips, uint32_t* out_failed_clip = nullptr)
		{
			m_num_requests = 0;
			return m_batch.upload(compressed_tracks, sizes, num_clips, true, out_failed_clip);
		}
		void reset() { m_batch.release(); m_num_requests = 0; }
		bool is_initialized() const { return m_batch.clipset() != nullptr; }
		uint32_t get_num_clips() const { return m_batch.info().num_clips; }
		uint32_t get_max_num_tracks() const { return m_batch.info().max_tracks; }
		void set_looping_policy(sample_looping_policy policy) { m_looping = policy; }
		sample_looping_policy get_looping_policy() const { return m_looping; }

		// seek() for a batch: d_requests is a DEVICE array of { clip index, sample_time }; evaluated on the GPU with the decode
		void seek(const aclb200_request* d_requests, uint32_t num_requests, sample_rounding_policy rounding_policy)
		{
			m_requests = d_requests;
			m_num_requests = num_requests;
			m_rounding = rounding_policy;
		}

		// decompress_tracks() for the batch, asynchronous on writer.stream
		template<class writer_type>
		void decompress_tracks(const writer_type& writer, const uint8_t* d_per_track_rounding = nullptr, const float* d_variable_defaults = nullptr)
		{
			if (!is_initialized() || m_num_requests == 0)
				return;
			aclb200_options options = make_options<settings_type>(writer, m_rounding, m_looping);
			// unlike the one pose shim, variable defaults live on the device here
			options.default_rotation_mode = static_cast<uint32_t>(writer_type::get_default_rotation_mode());
			options.default_translation_mode = static_cast<uint32_t>(writer_type::get_default_translation_mode());
			options.default_scale_mode = static_cast<uint32_t>(writer_type::get_default_scale_mode());
			options.d_variable_defaults = d_variable_defaults;
			options.d_per_track_rounding = d_per_track_rounding;
			options.output_layout = writer.output_layout;
			options.pose_stride_bytes = writer.pose_stride_bytes;
			m_batch.decompress_tracks(m_requests, m_num_requests, options, writer.d_poses, writer.stream);
		}

		batch_decompressor& decompressor() { return m_batch; }

	private:
		batch_decompressor m_batch;
		const aclb200_request* m_requests = nullptr;
		uint32_t m_num_requests = 0;
		sample_looping_policy m_looping = sample_looping_policy::as_compressed;
		sample_rounding_policy m_rounding = sample_rounding_policy::none;
	};

	// Drop-in for acl::decompression_context<settings> on transform clips: one clip bound, one pose per decompress_tracks().
	template<class settings_type = default_transform_decompression_settings>
	class decompression_context
	{
		static_assert(std::is_base_of<decompression_settings, settings_type>::value, "settings_type must derive from decompression_settings");		// decompress.h:197

	public:
		explicit decompression_context(device_context& device) : m_batch(device) {}

		// initialize(const compressed_tracks&), decompress.impl.h:66-129: false when the buffer is not a valid compressed_tracks
		bool initialize(const void* compressed_tracks, uint32_t size)
		{
			m_bound = nullptr;
			const void* blobs[1] = { compressed_tracks };
			const uint32_t sizes[1] = { size };
			if (!m_batch.upload(blobs, sizes, 1))
				return false;
			if (m_batch.info().track_type != ACLB200_TRACK_QVVF)
				return false;		// this shim covers transform clips; scalar clips go through aclb200_scalar_decompress_tracks
			m_bound = compressed_tracks;
			m_looping = sample_looping_policy::as_compressed;
			m_has_seeked = false;
			m_pose.assign(size_t(m_batch.info().max_tracks) * 12, 0.0F);
			return true;
		}
		// relocated(const compressed_tracks&), decompress.impl.h:131-156: the clip moved in host memory; the device copy is unaffected
		bool relocated(const void* compressed_tracks)
		{
			if (m_bound == nullptr || compressed_tracks == nullptr)
				return false;
			m_bound = compressed_tracks;
			return true;
		}
		void reset()		// decompress.h:120-124
		{
			m_batch.release();
			m_bound = nullptr;
			m_has_seeked = false;
		}
		const void* get_compressed_tracks() const { return m_bound; }
		bool is_initialized() const { return m_bound != nullptr; }
		bool is_bound_to(const void* compressed_tracks) const { return m_bound != nullptr && m_bound == compressed_tracks; }	// decompress.impl.h:158-177 compares pointer + hash; the hash was verified at upload
		void set_looping_policy(sample_looping_policy policy) { m_looping = policy; }
		sample_looping_policy get_looping_policy() const { return m_looping; }

		// seek(), decompress.impl.h:207-222: remembered here, evaluated on the GPU together with the decode
		void seek(float sample_time, sample_rounding_policy rounding_policy)
		{
			// per track rounding needs a device array of policies (aclb200_options::d_per_track_rounding): use the C ABI for it
			if (rounding_policy == sample_rounding_policy::per_track)
				throw error(ACLB200_ERR_UNSUPPORTED, "decompression_context::seek: per_track rounding is available through aclb200_options::d_per_track_rounding");
			m_sample_time = sample_time;
			m_rounding = rounding_policy;
			m_has_seeked = true;
		}

		template<class writer_type>
		void decompress_tracks(writer_type& writer)
		{
			if (!is_initialized() || !m_has_seeked)
				return;		// the reference asserts; like it, nothing is written
			aclb200_options options = make_options<settings_type>(writer, m_rounding, m_looping);
			const uint32_t num_tracks = m_batch.info().max_tracks;
			prime_skipped(writer, options, num_tracks);
			const aclb200_request request = { 0u, m_sample_time };
			m_batch.decompress_tracks_host(&request, 1, options, m_pose.data(), m_pose.size() * sizeof(float));
			for (uint32_t track = 0; track < num_tracks; ++track)
				replay(writer, options, track, m_pose.data() + size_t(track) * 12);
		}

		template<class writer_type>
		void decompress_track(uint32_t track_index, writer_type& writer)
		{
			if (!is_initialized() || !m_has_seeked || track_index >= m_batch.info().max_tracks)
				return;
			// a one pose decode costs the same as a one bone decode here: reuse the full path and replay one bone.
			// (The batched aclb200_decompress_track entry point is the one that follows decompress_track_v0 operation for operation.)
			aclb200_options options = make_options<settings_type>(writer, m_rounding, m_looping);
			const uint32_t num_tracks = m_batch.info().max_tracks;
			prime_skipped(writer, options, num_tracks);
			const aclb200_request request = { 0u, m_sample_time };
			m_batch.decompress_tracks_host(&request, 1, options, m_pose.data(), m_pose.size() * sizeof(float));
			replay(writer, options, track_index, m_pose.data() + size_t(track_index) * 12);
		}

	private:
		// `skipped` default sub-tracks are left untouched by the device: mark them so the replay can tell them from written ones
		template<class writer_type>
		void prime_skipped(const writer_type&, const aclb200_options& options, uint32_t num_tracks)
		{
			const bool any_skipped = options.default_rotation_mode == ACLB200_DEFAULT_SKIPPED || options.default_translation_mode == ACLB200_DEFAULT_SKIPPED
				|| options.default_scale_mode == ACLB200_DEFAULT_SKIPPED;
			if (!any_skipped)
				return;
			const uint32_t marker = 0x7FC0ACB2u;		// a NaN payload no decode produces
			for (uint32_t i = 0; i < num_tracks * 12; ++i)
				std::memcpy(&m_pose[i], &marker, sizeof(marker));
		}

		static bool is_marker(const float* v)
		{
			uint32_t bits;
			std::memcpy(&bits, v, sizeof(bits));
			return bits == 0x7FC0ACB2u;
		}

		template<class writer_type>
		static void replay(writer_type& writer, const aclb200_options& options, uint32_t track, const float* bone)
		{
			if (!writer_type::skip_all_rotations() && !writer.skip_track_rotation(track))
			{
				if (!(options.default_rotation_mode == ACLB200_DEFAULT_SKIPPED && is_marker(bone)))
					writer.write_rotation(track, float4{ bone[0], bone[1], bone[2], bone[3] });
				else if (writer_type::get_default_rotation_mode() == default_sub_track_mode::variable)
					writer.write_rotation(track, writer.get_variable_default_rotation(track));
			}
			if (!writer_type::skip_all_translations() && !writer.skip_track_translation(track))
			{
				if (!(options.default_translation_mode == ACLB200_DEFAULT_SKIPPED && is_marker(bone + 4)))
					writer.write_translation(track, float4{ bone[4], bone[5], bone[6], 0.0F });
				else if (writer_type::get_default_translation_mode() == default_sub_track_mode::variable)
					writer.write_translation(track, writer.get_variable_default_translation(track));
			}
			if (!writer_type::skip_all_scales() && !writer.skip_track_scale(track))
			{
				if (!(options.default_scale_mode == ACLB200_DEFAULT_SKIPPED && is_marker(bone + 8)))
					writer.write_scale(track, float4{ bone[8], bone[9], bone[10], 0.0F });
				else if (writer_type::get_default_scale_mode() == default_sub_track_mode::variable)
					writer.write_scale(track, writer.get_variable_default_scale(track));
			}
		}

		batch_decompressor m_batch;
		const void* m_bound = nullptr;
		sample_looping_policy m_looping = sample_looping_policy::as_compressed;
		sample_rounding_policy m_rounding = sample_rounding_policy::none;
		float m_sample_time = 0.0F;
		bool m_has_seeked = false;
		std::vector<float> m_pose;
	};
}
