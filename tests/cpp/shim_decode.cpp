// tests/cpp/shim_decode.cpp -- exercises include/acl_b200/decompress.h the way a reference call site would be written.
// usage: shim_decode <clip.acl.bin> <default|debug> <t0> [t1 ...]   prints one line per (time, track): 12 floats as hex words.
// (settings: acl's default_transform_decompression_settings only support the variable formats, like the reference's; clips in
// full precision formats need the debug settings)
// Exit code 3 when no usable GPU exists (the library has no CPU fallback), so the CPU-side test can check exactly that.
#include "../../include/acl_b200/decompress.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

namespace
{
	struct pose_writer : acl_b200::track_writer
	{
		explicit pose_writer(uint32_t num_tracks) : values(size_t(num_tracks) * 12, 0.0F) {}
		void write_rotation(uint32_t track, acl_b200::float4 v) { float* d = &values[size_t(track) * 12]; d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
		void write_translation(uint32_t track, acl_b200::float4 v) { float* d = &values[size_t(track) * 12 + 4]; d[0] = v.x; d[1] = v.y; d[2] = v.z; }
		void write_scale(uint32_t track, acl_b200::float4 v) { float* d = &values[size_t(track) * 12 + 8]; d[0] = v.x; d[1] = v.y; d[2] = v.z; }
		std::vector<float> values;
	};
}

template<class settings_type>
int run(const std::vector<char>& blob, int argc, char** argv);

int main(int argc, char** argv)
{
	if (argc < 4)
		return 2;
	std::ifstream file(argv[1], std::ios::binary);
	std::vector<char> blob((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
	if (blob.empty())
		return 2;
	if (std::string(argv[2]) == "debug")
		return run<acl_b200::debug_transform_decompression_settings>(blob, argc, argv);
	return run<acl_b200::default_transform_decompression_settings>(blob, argc, argv);
}

template<class settings_type>
int run(const std::vector<char>& blob, int argc, char** argv)
{
	try
	{
		acl_b200::device_context device(0);
		acl_b200::decompression_context<settings_type> context(device);
		if (!context.initialize(blob.data(), uint32_t(blob.size())))
		{
			std::printf("initialize failed\n");
			return 1;
		}
		uint32_t num_tracks = 0;
		std::memcpy(&num_tracks, blob.data() + 16, 4);		// tracks_header::num_tracks
		for (int i = 3; i < argc; ++i)
		{
			pose_writer writer(num_tracks);
			context.seek(float(std::atof(argv[i])), acl_b200::sample_rounding_policy::none);
			context.decompress_tracks(writer);
			for (uint32_t track = 0; track < num_tracks; ++track)
			{
				std::printf("%d %u", i - 3, track);
				for (int c = 0; c < 12; ++c)
				{
					uint32_t bits;
					std::memcpy(&bits, &writer.values[size_t(track) * 12 + c], 4);
					std::printf(" %08x", bits);
				}
				std::printf("\n");
			}
		}
	}
	catch (const acl_b200::error& e)
	{
		std::fprintf(stderr, "%s\n", e.what());
		return e.status == ACLB200_ERR_NO_DEVICE ? 3 : 1;
	}
	return 0;
}
