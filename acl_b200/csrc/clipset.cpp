// acl_b200/csrc/clipset.cpp -- host side of aclb200_upload_clips: validate the caller's compressed_tracks blobs,
// lay them out for HBM and build the acceleration index (see layout.h).
//
// This is the batched counterpart of decompression_context::initialize()
// (includes/acl/decompression/impl/decompress.impl.h:66-83 -> initialize_v0,
// includes/acl/decompression/impl/decompression.transform.h:84-132 / decompression.scalar.h:99-123): the
// reference caches a handful of header fields per context; we resolve every offset the decoder needs once per
// clip and, because a GPU thread cannot run the reference's serial cursors, also tabulate the per sub-track bit
// offsets of every segment.
#include "context.h"

#include <cstring>
#include <functional>
#include <limits>

namespace aclb200
{
	namespace
	{
		inline uint32_t rd_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
		inline uint16_t rd_u16(const uint8_t* p) { uint16_t v; std::memcpy(&v, p, 2); return v; }
		inline float rd_f32(const uint8_t* p) { float v; std::memcpy(&v, p, 4); return v; }
		inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) & ~(a - 1); }
		inline uint64_t align_up64(uint64_t v, uint64_t a) { return (v + a - 1) & ~(a - 1); }

		// core/hash.h:44-84 (FNV-1a 32), the hash compressed_tracks::is_valid(true) verifies
		uint32_t hash32(const uint8_t* data, size_t size)
		{
			uint32_t acc = 2166136261u;
			for (size_t i = 0; i < size; ++i)
				acc = (acc ^ data[i]) * 16777619u;
			return acc;
		}

		// core/impl/variable_bit_rates.h:39-43
		const uint8_t k_bit_rate_num_bits_v0[] = { 0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 32 };
		const uint8_t k_bit_rate_num_bits[] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 32 };

		// calculate_finite_duration, core/impl/time_utils.impl.h:105-114
		float finite_duration(uint32_t num_samples, float sample_rate)
		{
			if (num_samples <= 1)
				return 0.0F;
			return float(num_samples - 1) / sample_rate;
		}

		struct parse_result
		{
			ClipDesc desc;
			std::vector<uint8_t> index;		// this clip's index block
			uint32_t looping_policy;
			uint32_t track_type;
		};

		// Returns an empty string on success, else why the clip is rejected. `unsupported` tells apart valid ACL data we
		// refuse (database clips) from invalid buffers.
		std::string parse_clip(const uint8_t* blob, uint32_t size, bool check_hash, parse_result& out, bool& unsupported)
		{
			unsupported = false;

			// compressed_tracks::is_valid, core/impl/compressed_tracks.impl.h:278-301
			if (blob == nullptr || size < k_type_header_offset + 20)
				return "buffer too small";
			const uint32_t stored_size = rd_u32(blob + 0);
			if (stored_size > size || stored_size < k_type_header_offset + 20)
				return "stored size does not fit the buffer";
			if (rd_u32(blob + 8) != k_tag)
				return "invalid tag";
			const uint32_t version = rd_u16(blob + 12);
			if (blob[14] != 0)
				return "invalid algorithm type";
			if (version < k_version_first || version > k_version_latest)
				return "invalid algorithm version";
			if (check_hash && hash32(blob + 8, stored_size - 8) != rd_u32(blob + 4))
				return "invalid hash";

			const uint32_t track_type = blob[15];
			const uint32_t num_tracks = rd_u32(blob + 16);
			const uint32_t num_samples = rd_u32(blob + 20);
			const float sample_rate = rd_f32(blob + 24);
			const uint32_t misc = rd_u32(blob + 28);
			const bool is_wrap = version > k_version_first && ((misc >> 30) & 1) != 0;	// compressed_tracks.impl.h:127-134

			if (track_type != k_track_qvvf && track_type > 4)
				return "unsupported track type";
			if (num_tracks > k_max_tracks)
			{
				unsupported = true;
				return "too many tracks for the bone index table";
			}

			ClipDesc& d = out.desc;
			std::memset(&d, 0, sizeof(d));
			d.num_tracks = num_tracks;
			d.num_samples = num_samples;
			d.sample_rate = sample_rate;
			d.duration_clamp = finite_duration(num_samples, sample_rate);
			d.duration_wrap = finite_duration(num_samples != 0 ? num_samples + 1 : 0, sample_rate);
			d.hash = rd_u32(blob + 4);
			d.size = stored_size;
			out.looping_policy = is_wrap ? ACLB200_LOOP_WRAP : ACLB200_LOOP_CLAMP;
			out.track_type = track_type;
			if (is_wrap)
				d.flags |= k_clip_wrap;

			const uint8_t* th = blob + k_type_header_offset;
			auto in_bounds = [&](uint64_t offset_from_blob, uint64_t bytes) { return offset_from_blob + bytes <= stored_size; };

			if (track_type != k_track_qvvf)
			{
				// scalar_tracks_header, core/impl/compressed_headers.h:140-165; stream writers compress.scalar.impl.h:93-175
				const uint32_t num_bits_per_frame = rd_u32(th + 0);
				const uint32_t metadata_offset = k_type_header_offset + rd_u32(th + 4);
				const uint32_t constant_offset = k_type_header_offset + rd_u32(th + 8);
				const uint32_t range_offset = k_type_header_offset + rd_u32(th + 12);
				const uint32_t animated_offset = k_type_header_offset + rd_u32(th + 16);
				if (num_tracks != 0 && !in_bounds(metadata_offset, num_tracks))
					return "track metadata out of bounds";

				const uint32_t num_components = track_type <= 3 ? track_type + 1 : 4;
				const uint8_t* table = version == k_version_first ? k_bit_rate_num_bits_v0 : k_bit_rate_num_bits;
				const uint32_t table_size = version == k_version_first ? sizeof(k_bit_rate_num_bits_v0) : sizeof(k_bit_rate_num_bits);

				out.index.resize(size_t(num_tracks) * sizeof(ScalarTrackDesc));
				ScalarTrackDesc* tracks = reinterpret_cast<ScalarTrackDesc*>(out.index.data());
				uint32_t bit_offset = 0, constant_index = 0, range_index = 0;
				for (uint32_t track = 0; track < num_tracks; ++track)
				{
					const uint32_t bit_rate = blob[metadata_offset + track];
					if (bit_rate >= table_size)
						return "invalid scalar bit rate";
					const uint32_t num_bits = table[bit_rate];
					tracks[track].bit_offset = bit_offset;
					if (num_bits == 0)
					{
						tracks[track].value_index_and_bits = (constant_index << 8) | 0u;
						constant_index += num_components;
					}
					else if (num_bits == 32)
						tracks[track].value_index_and_bits = 32u;
					else
					{
						tracks[track].value_index_and_bits = (range_index << 8) | num_bits;
						range_index += num_components * 2;
					}
					bit_offset += num_bits * num_components;
				}
				if (bit_offset != num_bits_per_frame)
					return "scalar bit rates do not add up to num_bits_per_frame";
				if (!in_bounds(constant_offset, uint64_t(constant_index) * 4) || !in_bounds(range_offset, uint64_t(range_index) * 4))
					return "scalar constant / range values out of bounds";
				if (num_samples != 0 && !in_bounds(animated_offset, (uint64_t(num_bits_per_frame) * num_samples + 7) / 8))
					return "scalar animated values out of bounds";
				if ((constant_offset & 3) != 0 || (range_offset & 3) != 0)
					return "scalar constant / range values are not 4 byte aligned";

				d.num_segments = 1;
				d.samples_per_segment = num_samples;
				d.num_constant[0] = num_bits_per_frame;
				d.constant_offset[0] = constant_offset;
				d.constant_offset[1] = range_offset;
				d.constant_offset[2] = animated_offset;
				d.bone_table_offset = 0;
				return std::string();
			}

			// ---- transform clips: tracks_header::misc_packed accessors, compressed_headers.h:95-124 ----
			if (!in_bounds(k_type_header_offset, k_transform_header_size))
				return "transform header out of bounds";
			const bool has_scale = (misc & 1) != 0;
			const uint32_t scale_format = (misc >> 2) & 1;
			const uint32_t translation_format = (misc >> 3) & 1;
			const uint32_t rotation_format = (misc >> 4) & 15;
			const bool has_database = ((misc >> 8) & 1) != 0;
			const bool has_stripped = ((misc >> 10) & 1) != 0;
			if (has_database)
			{
				unsupported = true;
				return "clips bound to a streaming database are not supported";
			}
			if (rotation_format != k_rot_full && rotation_format != k_rot_drop_w_full && rotation_format != k_rot_drop_w_variable)
				return "invalid rotation format";

			// transform_tracks_header, compressed_headers.h:227-262
			const uint32_t num_segments = rd_u32(th + 0);
			const uint32_t num_variable = rd_u32(th + 4);
			const uint32_t num_animated[3] = { rd_u32(th + 8), rd_u32(th + 12), has_scale ? rd_u32(th + 16) : 0u };
			const uint32_t num_constant[3] = { rd_u32(th + 20), rd_u32(th + 24), has_scale ? rd_u32(th + 28) : 0u };
			const uint32_t segment_headers_offset = k_type_header_offset + rd_u32(th + 36);
			const uint32_t sub_track_types_offset = k_type_header_offset + rd_u32(th + 40);
			const uint32_t constant_data_offset = k_type_header_offset + rd_u32(th + 44);
			const uint32_t clip_range_data_offset = k_type_header_offset + rd_u32(th + 48);

			if (num_tracks == 0)
			{
				d.num_segments = 0;
				return std::string();	// empty track list: seek/decompress are no-ops (decompression.transform.h:211-212)
			}
			if (num_segments == 0)
				return "clip without segments";
			for (int k = 0; k < 3; ++k)
				if (num_animated[k] > num_tracks || num_constant[k] > num_tracks)
					return "sub-track counts exceed the track count";

			const bool rot_variable = rotation_format == k_rot_drop_w_variable;
			const bool rot_full = rotation_format == k_rot_full;
			const bool trans_variable = translation_format == 1;
			const bool scale_variable = scale_format == 1;
			const uint32_t raw_marker = version >= k_version_raw31 ? 31u : 32u;		// animated_track_cache.transform.h:523
			const uint32_t segment_header_size = has_stripped ? 20u : 16u;				// compressed_headers.h:171-197
			const uint32_t num_entries = (num_tracks + 15) / 16;
			const uint32_t padded_rotations = align_up(num_animated[0], 4);
			const uint32_t num_animated_total = num_animated[0] + num_animated[1] + num_animated[2];

			if (!in_bounds(segment_headers_offset, uint64_t(segment_header_size) * num_segments))
				return "segment headers out of bounds";
			if (!in_bounds(sub_track_types_offset, uint64_t(num_entries) * 4 * (has_scale ? 3 : 2)))
				return "sub-track types out of bounds";
			if (num_segments > 1 && !in_bounds(84, uint64_t(num_segments + 1) * 4))
				return "segment start indices out of bounds";

			d.flags |= has_scale ? k_clip_has_scale : 0u;
			d.flags |= ((misc >> 1) & 1) ? k_clip_default_scale_one : 0u;
			d.flags |= rot_variable ? k_clip_rot_variable : 0u;
			d.flags |= trans_variable ? k_clip_trans_variable : 0u;
			d.flags |= (has_scale && scale_variable) ? k_clip_scale_variable : 0u;
			d.flags |= rot_full ? k_clip_rot_full : 0u;
			d.flags |= has_stripped ? k_clip_stripped : 0u;
			d.flags |= num_segments > 1 ? k_clip_has_segments : 0u;
			d.num_segments = num_segments;
			d.samples_per_segment = num_samples / num_segments;
			for (int k = 0; k < 3; ++k)
			{
				d.num_animated[k] = num_animated[k];
				d.num_constant[k] = num_constant[k];
			}
			d.num_animated_total = num_animated_total;
			d.start_indices_offset = 84;		// transform_tracks_header::get_segment_start_indices, compressed_headers.h:271-272

			// constant_track_cache_v0::initialize, constant_track_cache.transform.h:96-110
			d.constant_offset[0] = constant_data_offset;
			d.constant_offset[1] = d.constant_offset[0] + (rot_full ? 16u : 12u) * num_constant[0];
			d.constant_offset[2] = d.constant_offset[1] + 12u * num_constant[1];
			if (!in_bounds(d.constant_offset[2], 12ull * num_constant[2]))
				return "constant track data out of bounds";

			// animated_track_cache_v0::initialize, animated_track_cache.transform.h:1259-1262,1293-1294
			d.clip_range_offset[0] = clip_range_data_offset;
			d.clip_range_offset[1] = d.clip_range_offset[0] + (rot_variable ? 24u * num_animated[0] : 0u);
			d.clip_range_offset[2] = d.clip_range_offset[1] + (trans_variable ? 24u * num_animated[1] : 0u);
			{
				const uint64_t clip_range_end = uint64_t(d.clip_range_offset[2]) + ((has_scale && scale_variable) ? 24ull * num_animated[2] : 0ull);
				if ((rot_variable || trans_variable || (has_scale && scale_variable)) && clip_range_end > stored_size)
					return "clip range data out of bounds";
			}

			// Sections the kernels read as aligned words; the reference writer guarantees this
			// (compression/impl/compress.transform.impl.h:316-327), anything else is not an ACL buffer.
			if ((d.constant_offset[0] & 3) != 0 || (d.clip_range_offset[0] & 3) != 0)
				return "constant / clip range data is not 4 byte aligned";

			// ---- index block: BoneDesc[num_tracks] | SegDesc[num_segments] | entries[num_segments][num_animated_total] ----
			const size_t bone_table_bytes = align_up64(uint64_t(num_tracks) * 8, 16);
			const size_t seg_table_bytes = size_t(num_segments) * sizeof(SegDesc);
			const size_t entries_bytes_per_segment = align_up64(uint64_t(num_animated_total) * 4, 16);
			out.index.assign(bone_table_bytes + seg_table_bytes + entries_bytes_per_segment * num_segments, 0);
			d.bone_table_offset = 0;
			d.seg_table_offset = uint32_t(bone_table_bytes);

			// Bone table: rank of each bone among the constant / animated sub-tracks of its kind, i.e. the prefix popcounts of
			// decompress_track_v0 (decompression.transform.h:1873-1891) evaluated once for every bone.
			{
				uint64_t* bones = reinterpret_cast<uint64_t*>(out.index.data());
				uint32_t constant_rank[3] = { 0, 0, 0 };
				uint32_t animated_rank[3] = { 0, 0, 0 };
				const uint32_t num_kinds = has_scale ? 3u : 2u;
				for (uint32_t track = 0; track < num_tracks; ++track)
				{
					uint64_t desc = 0;
					for (uint32_t kind = 0; kind < num_kinds; ++kind)
					{
						// packed_sub_track_types: 16 sub-tracks per u32, 2 bits each, MSB first (compressed_headers.h:214-224)
						const uint32_t packed = rd_u32(blob + sub_track_types_offset + 4 * (kind * num_entries + track / 16));
						const uint32_t type = (packed >> ((15 - track % 16) * 2)) & 3;
						if (type == 3)
							return "invalid sub-track type";
						uint32_t rank = 0;
						if (type == 1)
							rank = constant_rank[kind]++;
						else if (type == 2)
							rank = animated_rank[kind]++;
						desc |= (uint64_t(type) | (uint64_t(rank) << 2)) << (k_bone_kind_shift * kind);
					}
					bones[track] = desc;
				}
				for (uint32_t kind = 0; kind < num_kinds; ++kind)
					if (constant_rank[kind] != num_constant[kind] || animated_rank[kind] != num_animated[kind])
						return "sub-track types disagree with the header counts";
			}

			// Segment tables
			SegDesc* segs = reinterpret_cast<SegDesc*>(out.index.data() + bone_table_bytes);
			for (uint32_t segment = 0; segment < num_segments; ++segment)
			{
				const uint8_t* header = blob + segment_headers_offset + segment_header_size * segment;
				const uint32_t pose_bit_size = rd_u32(header + 0);
				const uint32_t rotation_bit_size = rd_u32(header + 4);
				const uint32_t translation_bit_size = rd_u32(header + 8);
				const uint32_t segment_data = k_type_header_offset + rd_u32(header + 12);

				// transform_tracks_header::get_segment_data, compressed_headers.h:310-324 (the blob is 16 byte aligned, so aligning
				// offsets equals aligning addresses)
				const uint32_t format_offset = segment_data;
				const uint32_t range_offset = align_up(format_offset + num_variable, 2);
				const uint32_t range_size = num_segments > 1 ? 6u * num_variable : 0u;
				const uint32_t animated_offset = align_up(range_offset + range_size, 4);
				if (!in_bounds(format_offset, num_variable) || !in_bounds(range_offset, range_size) || !in_bounds(animated_offset, 0))
					return "segment data out of bounds";

				SegDesc& seg = segs[segment];
				seg.animated_offset = animated_offset;
				seg.pose_bit_size = pose_bit_size;
				seg.sample_indices = has_stripped ? rd_u32(header + 16) : 0xFFFFFFFFu;
				seg.entries_offset = uint32_t(bone_table_bytes + seg_table_bytes + entries_bytes_per_segment * segment);
				seg.format_offset = format_offset;
				// rotation range data is padded to groups of 4, translations / scales are not (animated_track_cache.transform.h:1271-1276,1300-1302)
				seg.range_offset[0] = range_offset;
				seg.range_offset[1] = seg.range_offset[0] + (rot_variable ? 6u * padded_rotations : 0u);
				seg.range_offset[2] = seg.range_offset[1] + (trans_variable ? 6u * num_animated[1] : 0u);

				// Number of stored key frames, to bound the animated data (stripped segments store popcount(sample_indices) frames)
				uint32_t stored_key_frames;
				if (has_stripped)
					stored_key_frames = uint32_t(__builtin_popcount(seg.sample_indices));
				else if (num_segments == 1)
					stored_key_frames = num_samples;
				else
				{
					const uint32_t start = rd_u32(blob + 84 + 4 * segment);
					const uint32_t next = segment + 1 < num_segments ? rd_u32(blob + 84 + 4 * (segment + 1)) : num_samples;
					if (next < start)
						return "segment start indices are not sorted";
					stored_key_frames = next - start;
				}
				if (!in_bounds(animated_offset, (uint64_t(pose_bit_size) * stored_key_frames + 7) / 8))
					return "animated data out of bounds";

				// Per sub-track bit offsets: the running sum of segment_animated_sampling_context_v0::animated_track_data_bit_offset
				// (animated_track_cache.transform.h:598-599,653, cursors set up at :1240-1308)
				uint32_t* entries = reinterpret_cast<uint32_t*>(out.index.data() + seg.entries_offset);
				const uint8_t* format = blob + format_offset;
				const bool variable[3] = { rot_variable, trans_variable, scale_variable };
				const uint32_t kind_bit_offset[3] = { 0u, rotation_bit_size, rotation_bit_size + translation_bit_size };
				const uint32_t kind_format_offset[3] = { 0u, rot_variable ? padded_rotations : 0u,
					(rot_variable ? padded_rotations : 0u) + (trans_variable ? num_animated[1] : 0u) };
				uint32_t entry_index = 0;
				for (uint32_t kind = 0; kind < 3; ++kind)
				{
					uint32_t bit_offset = kind_bit_offset[kind];
					for (uint32_t j = 0; j < num_animated[kind]; ++j, ++entry_index)
					{
						uint32_t code, stream_bits;
						if (variable[kind])
						{
							const uint32_t stored = format[kind_format_offset[kind] + j];
							if (stored == 0)
							{
								code = 0;
								stream_bits = 0;
								if (num_segments == 1)
									return "constant bit rate inside a single segment clip";
							}
							else if (stored == raw_marker)
							{
								code = 32u | k_entry_raw;
								stream_bits = 96;
							}
							else if (stored <= 23)
							{
								code = stored;
								stream_bits = stored * 3;
							}
							else
								return "invalid per track bit count";
						}
						else
						{
							code = 32u | k_entry_raw;
							stream_bits = (kind == 0 && rot_full) ? 128u : 96u;
						}
						if (bit_offset >= (1u << 24))
							return "key frame too large for the sub-track entry table";
						entries[entry_index] = (bit_offset << 8) | code;
						bit_offset += stream_bits;
					}
					if (bit_offset > pose_bit_size && num_animated[kind] != 0)
						return "sub-track bit widths exceed the animated pose size";
				}
			}
			return std::string();
		}
	}

	aclb200_status set_error(aclb200_context* context, aclb200_status status, const std::string& message)
	{
		if (context != nullptr)
			context->last_error = message;
		return status;
	}

	aclb200_status check_cuda(aclb200_context* context, cudaError_t error, const char* what)
	{
		if (error == cudaSuccess)
			return ACLB200_OK;
		const aclb200_status status = error == cudaErrorMemoryAllocation ? ACLB200_ERR_OUT_OF_MEMORY : ACLB200_ERR_CUDA;
		return set_error(context, status, std::string(what) + ": " + cudaGetErrorString(error));
	}

	aclb200_status build_clipset(aclb200_context* context, const std::function<const uint8_t*(uint32_t)>& get_blob, const uint32_t* sizes,
		uint32_t num_clips, bool check_hash, aclb200_clipset** out_clipset, uint32_t* out_failed_clip)
	{
		if (context == nullptr || out_clipset == nullptr || sizes == nullptr || num_clips == 0)
			return set_error(context, ACLB200_ERR_INVALID_ARGUMENT, "upload_clips: null argument or empty clip list");
		*out_clipset = nullptr;

		aclb200_clipset* set = new aclb200_clipset();
		set->device = context->device;
		set->host_clips.resize(num_clips);
		set->host_looping.resize(num_clips);

		// Pass 1: parse, validate and size
		std::vector<uint8_t> index_host;
		uint64_t blob_bytes = 0;
		uint32_t set_track_type = 0xFFFFFFFFu;
		uint32_t max_tracks = 0, min_tracks = std::numeric_limits<uint32_t>::max(), max_animated_total = 0;
		parse_result parsed;
		for (uint32_t clip = 0; clip < num_clips; ++clip)
		{
			bool unsupported = false;
			const std::string error = parse_clip(get_blob(clip), sizes[clip], check_hash, parsed, unsupported);
			if (!error.empty())
			{
				if (out_failed_clip != nullptr)
					*out_failed_clip = clip;
				delete set;
				return set_error(context, unsupported ? ACLB200_ERR_UNSUPPORTED : ACLB200_ERR_INVALID_CLIP,
					"clip " + std::to_string(clip) + ": " + error);
			}
			if (set_track_type == 0xFFFFFFFFu)
				set_track_type = parsed.track_type;
			else if (set_track_type != parsed.track_type)
			{
				if (out_failed_clip != nullptr)
					*out_failed_clip = clip;
				delete set;
				return set_error(context, ACLB200_ERR_UNSUPPORTED, "clip " + std::to_string(clip) + ": a clip set holds a single track type");
			}

			ClipDesc& desc = set->host_clips[clip];
			desc = parsed.desc;
			desc.blob_offset = blob_bytes;
			desc.index_offset = index_host.size();
			blob_bytes += align_up64(desc.size, k_blob_alignment);
			index_host.insert(index_host.end(), parsed.index.begin(), parsed.index.end());
			index_host.resize(align_up64(index_host.size(), 16), 0);
			set->host_looping[clip] = parsed.looping_policy;
			max_tracks = desc.num_tracks > max_tracks ? desc.num_tracks : max_tracks;
			min_tracks = desc.num_tracks < min_tracks ? desc.num_tracks : min_tracks;
			max_animated_total = desc.num_animated_total > max_animated_total ? desc.num_animated_total : max_animated_total;
		}
		blob_bytes += k_tail_slack;
		if (index_host.empty())
			index_host.resize(16, 0);

		set->info.num_clips = num_clips;
		set->info.track_type = set_track_type;
		set->info.max_tracks = max_tracks;
		set->info.min_tracks = min_tracks;
		set->info.blob_bytes = blob_bytes;
		set->info.index_bytes = index_host.size();
		set->max_animated_total = max_animated_total;

		// Pass 2: device allocation + copies (clips go through a bounded staging buffer so that a multi-GB clip set does
		// not need a second full host copy)
		cudaError_t error = cudaSetDevice(context->device);
		if (error == cudaSuccess) error = cudaMalloc(reinterpret_cast<void**>(&set->d_blobs), blob_bytes);
		if (error == cudaSuccess) error = cudaMalloc(reinterpret_cast<void**>(&set->d_index), index_host.size());
		if (error == cudaSuccess) error = cudaMalloc(reinterpret_cast<void**>(&set->d_clips), sizeof(ClipDesc) * size_t(num_clips));
		if (error == cudaSuccess) error = cudaMemset(set->d_blobs + (blob_bytes - k_tail_slack), 0, k_tail_slack);
		if (error == cudaSuccess)
		{
			const size_t staging_capacity = size_t(64) << 20;
			std::vector<uint8_t> staging;
			staging.reserve(staging_capacity);
			uint64_t staging_base = 0;
			auto flush = [&]() -> cudaError_t
			{
				if (staging.empty())
					return cudaSuccess;
				const cudaError_t e = cudaMemcpy(set->d_blobs + staging_base, staging.data(), staging.size(), cudaMemcpyHostToDevice);
				staging_base += staging.size();
				staging.clear();
				return e;
			};
			for (uint32_t clip = 0; clip < num_clips && error == cudaSuccess; ++clip)
			{
				const ClipDesc& desc = set->host_clips[clip];
				const size_t padded = size_t(align_up64(desc.size, k_blob_alignment));
				if (staging.size() + padded > staging_capacity)
					error = flush();
				if (padded > staging_capacity)
				{
					// a single huge clip: copy it directly
					if (error == cudaSuccess) error = cudaMemcpy(set->d_blobs + desc.blob_offset, get_blob(clip), desc.size, cudaMemcpyHostToDevice);
					if (error == cudaSuccess && padded > desc.size) error = cudaMemset(set->d_blobs + desc.blob_offset + desc.size, 0, padded - desc.size);
					staging_base += padded;
					continue;
				}
				const uint8_t* blob = get_blob(clip);
				staging.insert(staging.end(), blob, blob + desc.size);
				staging.resize(staging.size() + (padded - desc.size), 0);
			}
			if (error == cudaSuccess) error = flush();
		}
		if (error == cudaSuccess) error = cudaMemcpy(set->d_index, index_host.data(), index_host.size(), cudaMemcpyHostToDevice);
		if (error == cudaSuccess) error = cudaMemcpy(set->d_clips, set->host_clips.data(), sizeof(ClipDesc) * size_t(num_clips), cudaMemcpyHostToDevice);
		if (error != cudaSuccess)
		{
			const aclb200_status status = check_cuda(context, error, "upload_clips");
			cudaFree(set->d_blobs);
			cudaFree(set->d_index);
			cudaFree(set->d_clips);
			delete set;
			return status;
		}

		*out_clipset = set;
		return ACLB200_OK;
	}
}
