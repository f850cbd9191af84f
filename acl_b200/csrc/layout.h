// acl_b200/csrc/layout.h -- what lives in HBM for a clip set, shared by the host builder and the kernels.
//
// A clip set is three device buffers:
//   blobs : the caller's compressed_tracks buffers, byte for byte, each starting on a 16 byte boundary
//           (the format's own alignment, includes/acl/core/compressed_tracks.h:53) plus 64 bytes of tail
//           slack so the big-endian unaligned reads the format relies on never leave the allocation.
//   clips : one ClipDesc per clip: everything decompression_context::initialize() caches
//           (initialize_v0, decompression/impl/decompression.transform.h:84-132) plus offsets the reference
//           re-derives on every seek (transform_tracks_header::get_segment_data, core/impl/compressed_headers.h:310-324).
//   index : per clip acceleration tables that replace the two serial scans of the CPU decoder:
//           - BoneDesc[num_tracks]   : rank of every bone among the constant / animated sub-tracks of its kind
//                                      (the popcount walk of decompress_track_v0, decompression.transform.h:1873-1891)
//           - SegDesc[num_segments]  : per segment offsets
//           - u32 entries[segment][animated sub-track] : bit offset inside a key frame + bit width, i.e. the
//                                      running sum kept in animated_track_data_bit_offset
//                                      (animated_track_cache.transform.h:598-599,653)
//           The index costs ~6 % of the blob bytes and makes every (request, bone) independent.
#pragma once

#include <stdint.h>

namespace aclb200
{
	// ---- reference binary format constants (core/impl/compressed_headers.h) -------------------------------
	constexpr uint32_t k_tag = 0xac11ac11u;						// core/buffer_tag.h:49
	constexpr uint32_t k_version_first = 7;						// v02_00_00, core/compressed_tracks_version.h:75
	constexpr uint32_t k_version_raw31 = 9;						// v02_01_99_1: raw bit rate stored as 31 in the per-track format
	constexpr uint32_t k_version_latest = 10;					// v02_01_00
	constexpr uint32_t k_track_qvvf = 12;						// core/track_types.h:68
	constexpr uint32_t k_tracks_header_offset = 8;				// after raw_buffer_header {size, hash}
	constexpr uint32_t k_type_header_offset = 32;				// transform_tracks_header / scalar_tracks_header
	constexpr uint32_t k_transform_header_size = 52;
	constexpr uint32_t k_blob_alignment = 16;
	constexpr uint32_t k_tail_slack = 64;

	constexpr uint32_t k_rot_full = 0;							// rotation_format8, core/track_formats.h:48-53
	constexpr uint32_t k_rot_drop_w_full = 2;
	constexpr uint32_t k_rot_drop_w_variable = 3;

	// ---- ClipDesc::flags ---------------------------------------------------------------------------------
	constexpr uint32_t k_clip_has_scale = 1u << 0;
	constexpr uint32_t k_clip_default_scale_one = 1u << 1;
	constexpr uint32_t k_clip_rot_variable = 1u << 2;
	constexpr uint32_t k_clip_trans_variable = 1u << 3;
	constexpr uint32_t k_clip_scale_variable = 1u << 4;
	constexpr uint32_t k_clip_rot_full = 1u << 5;				// quatf_full: 4 stored components, no W reconstruction
	constexpr uint32_t k_clip_stripped = 1u << 6;				// has_stripped_keyframes
	constexpr uint32_t k_clip_wrap = 1u << 7;					// compressed_tracks::get_looping_policy() == wrap
	constexpr uint32_t k_clip_has_segments = 1u << 8;			// more than one segment => segment range data exists

	struct alignas(16) ClipDesc
	{
		uint64_t blob_offset;				// byte offset of the clip inside the blobs buffer
		uint64_t index_offset;				// byte offset of the clip's tables inside the index buffer
		uint32_t num_tracks;
		uint32_t num_samples;
		float    sample_rate;
		uint32_t flags;
		uint32_t num_segments;
		uint32_t samples_per_segment;		// num_samples / num_segments (the reference's segment guess, decompression.transform.h:377)
		float    duration_clamp;			// get_finite_duration(clamp), core/impl/compressed_tracks.impl.h:113-134
		float    duration_wrap;				// get_finite_duration(wrap)
		uint32_t num_animated[3];			// rotation, translation, scale
		uint32_t num_constant[3];
		uint32_t constant_offset[3];		// blob relative: constant rotations / translations / scales
		uint32_t clip_range_offset[3];		// blob relative: clip range of animated rotations / translations / scales
		uint32_t bone_table_offset;			// index relative: BoneDesc[num_tracks]
		uint32_t seg_table_offset;			// index relative: SegDesc[num_segments]
		uint32_t start_indices_offset;		// blob relative: segment_start_indices (only when num_segments > 1)
		uint32_t num_animated_total;		// rotations + translations + scales (entries per segment)
		// scalar clips (decompression.scalar.h) reuse: num_constant[0] = num_bits_per_frame,
		// constant_offset[0..2] = constant values / range values / animated values, bone_table_offset = ScalarTrackDesc[]
		uint32_t hash;
		uint32_t size;
	};
	static_assert(sizeof(ClipDesc) % 16 == 0, "ClipDesc must stay 16 byte sized");

	// BoneDesc: one u64 per bone. kind k in {0 rotation, 1 translation, 2 scale}:
	//   type  = (desc >> (22 * k)) & 3          0 default, 1 constant, 2 animated (packed_sub_track_types, compressed_headers.h:214-224)
	//   index = (desc >> (22 * k + 2)) & 0xFFFFF rank among the constant or animated sub-tracks of that kind
	constexpr uint32_t k_bone_kind_shift = 22;
	constexpr uint32_t k_bone_index_mask = 0xFFFFFu;
	constexpr uint32_t k_max_tracks = 1u << 18;					// scale rank has 18 bits left

	struct alignas(16) SegDesc
	{
		uint32_t animated_offset;			// blob relative byte offset of the segment's animated bit stream
		uint32_t pose_bit_size;				// segment_header::animated_pose_bit_size
		uint32_t sample_indices;			// stripped_segment_header_t::sample_indices (0xFFFFFFFF when nothing is stripped)
		uint32_t entries_offset;			// index relative: u32 entries[num_animated_total]
		uint32_t range_offset[3];			// blob relative: segment range data of rotations (SOA groups of 4) / translations / scales (AOS 6 B)
		uint32_t format_offset;				// blob relative: format_per_track_data (kept for the parity hooks)
	};
	static_assert(sizeof(SegDesc) == 32, "SegDesc is 32 bytes");

	// Sub-track entry: (bit offset inside the key frame << 8) | code, code = number of bits per component in the
	// stream (1..23), 0 = constant inside the segment (sample lives in the segment range bytes), 32 | k_entry_raw = raw floats.
	constexpr uint32_t k_entry_raw = 0x80u;
	constexpr uint32_t k_entry_bits_mask = 0x3Fu;

	// ScalarTrackDesc: { bit offset inside a frame, (value index << 8) | num_bits } ; value index counts floats in
	// constant_values (num_bits == 0) or range_values (0 < num_bits < 32).
	struct ScalarTrackDesc
	{
		uint32_t bit_offset;
		uint32_t value_index_and_bits;
	};
}
