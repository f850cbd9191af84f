// tests/cpp/shim_binding.cpp -- binding semantics of acl_b200::decompression_context with the stand-alone types (no reference
// headers): initialize / relocated / is_bound_to / reset follow decompress.impl.h:66-205 and decompression.transform.h:134-176.
// usage: shim_binding <clip.acl.bin> <other_clip.acl.bin>     exit 0 = all checks hold, 3 = no usable GPU, 1 = a check failed
#define ACLB200_WITH_ACL_HEADERS 0
#include "../../include/acl_b200/decompress.h"

#include <cstdio>
#include <fstream>
#include <iterator>

namespace
{
	std::vector<char> read_file(const char* path)
	{
		std::ifstream file(path, std::ios::binary);
		return std::vector<char>((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
	}

	struct counting_writer : acl_b200::track_writer
	{
		uint32_t writes = 0;
		void write_rotation(uint32_t, acl_b200::float4) { ++writes; }
		void write_translation(uint32_t, acl_b200::float4) { ++writes; }
		void write_scale(uint32_t, acl_b200::float4) { ++writes; }
	};

	int failures = 0;
	void expect(bool condition, const char* what)
	{
		if (!condition)
		{
			std::fprintf(stderr, "FAILED: %s\n", what);
			++failures;
		}
	}
}

int main(int argc, char** argv)
{
	if (argc < 3)
		return 2;
	const std::vector<char> clip = read_file(argv[1]), other = read_file(argv[2]);
	std::vector<char> moved(clip);
	if (clip.empty() || other.empty())
		return 2;
	try
	{
		using context_type = acl_b200::decompression_context<acl_b200::debug_transform_decompression_settings>;
		const acl_b200::compressed_tracks& tracks = *acl_b200::make_compressed_tracks(clip.data());
		const acl_b200::compressed_tracks& moved_tracks = *acl_b200::make_compressed_tracks(moved.data());
		const acl_b200::compressed_tracks& other_tracks = *acl_b200::make_compressed_tracks(other.data());

		context_type context;		// default constructible, the device is attached lazily
		counting_writer writer;
		expect(!context.is_initialized(), "a new context is not initialized");
		expect(!context.relocated(tracks), "relocated() on an unbound context returns false");
		context.seek(0.1F, acl_b200::sample_rounding_policy::none);
		context.decompress_tracks(writer);
		expect(writer.writes == 0, "decompress_tracks on an unbound context writes nothing");

		expect(context.initialize(tracks), "initialize a valid clip");
		expect(context.is_initialized() && context.get_compressed_tracks() == &tracks, "bound to the clip");
		expect(context.is_bound_to(tracks), "is_bound_to(the clip)");
		expect(!context.is_bound_to(moved_tracks), "is_bound_to(a copy at another address) is false");
		context.decompress_tracks(writer);
		expect(writer.writes == 0, "decompress_tracks before a seek writes nothing");
		context.seek(0.1F, acl_b200::sample_rounding_policy::none);
		context.decompress_tracks(writer);
		expect(writer.writes == tracks.get_num_tracks() * 3, "one rotation, translation and scale per track");

		expect(context.relocated(moved_tracks), "relocated(the same clip elsewhere) is accepted");
		expect(context.get_compressed_tracks() == &moved_tracks && context.is_bound_to(moved_tracks), "rebound to the new address");
		expect(!context.relocated(other_tracks), "relocated(a different clip) is refused: the hash differs");
		expect(context.get_compressed_tracks() == &moved_tracks, "a refused relocation leaves the binding alone");

		std::vector<char> corrupt(clip);
		corrupt[8] ^= 0x5A;		// buffer tag
		context_type second;
		expect(!second.initialize(*acl_b200::make_compressed_tracks(corrupt.data())), "initialize refuses a corrupt buffer");
		expect(!second.is_initialized(), "a failed initialize leaves the context unbound");

		context.reset();
		expect(!context.is_initialized() && context.get_compressed_tracks() == nullptr, "reset() unbinds");
	}
	catch (const acl_b200::error& e)
	{
		std::fprintf(stderr, "%s\n", e.what());
		return e.status == ACLB200_ERR_NO_DEVICE ? 3 : 1;
	}
	std::printf(failures == 0 ? "PASS\n" : "FAIL\n");
	return failures == 0 ? 0 : 1;
}
