// tests/cpp/shim_batch.cpp -- acl_b200::batch_context<settings>: bind N clips, seek a batch of (clip, time) requests that live in
// device memory, decode them with one launch into device memory (what SURVEY 8b calls the batched form of the context).
// usage: shim_batch <num_clips> <clip.acl.bin>... <clip index> <time> [<clip index> <time> ...]
// prints one line per (request, track): 12 floats as hex words.
#include "../../include/acl_b200/decompress.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

int main(int argc, char** argv)
{
	if (argc < 2)
		return 2;
	const int num_clips = std::atoi(argv[1]);
	if (num_clips <= 0 || argc < 2 + num_clips + 2 || ((argc - 2 - num_clips) % 2) != 0)
		return 2;
	std::vector<std::vector<char>> blobs;
	for (int i = 0; i < num_clips; ++i)
	{
		std::ifstream file(argv[2 + i], std::ios::binary);
		blobs.emplace_back((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
	}
	std::vector<aclb200_request> requests;
	for (int i = 2 + num_clips; i + 1 < argc; i += 2)
		requests.push_back(aclb200_request{ uint32_t(std::atoi(argv[i])), float(std::atof(argv[i + 1])) });

	try
	{
		acl_b200::device_context device(0);
		acl_b200::batch_context<acl_b200::default_transform_decompression_settings> context(device);
		std::vector<const void*> pointers;
		std::vector<uint32_t> sizes;
		for (const std::vector<char>& blob : blobs)
		{
			pointers.push_back(blob.data());
			sizes.push_back(uint32_t(blob.size()));
		}
		if (!context.bind(pointers.data(), sizes.data(), uint32_t(num_clips)))
			return 1;

		const uint32_t max_tracks = context.get_max_num_tracks();
		const size_t pose_floats = size_t(max_tracks) * 12;
		aclb200_request* d_requests = nullptr;
		float* d_poses = nullptr;
		if (cudaMalloc(&d_requests, requests.size() * sizeof(aclb200_request)) != cudaSuccess
			|| cudaMalloc(&d_poses, requests.size() * pose_floats * sizeof(float)) != cudaSuccess)
			return 1;
		cudaMemcpy(d_requests, requests.data(), requests.size() * sizeof(aclb200_request), cudaMemcpyHostToDevice);
		cudaMemset(d_poses, 0, requests.size() * pose_floats * sizeof(float));

		acl_b200::device_pose_writer writer;
		writer.d_poses = d_poses;
		writer.output_layout = ACLB200_LAYOUT_QVV48;
		context.seek(d_requests, uint32_t(requests.size()), acl_b200::sample_rounding_policy::none);
		context.decompress_tracks(writer);

		std::vector<float> poses(requests.size() * pose_floats);
		if (cudaMemcpy(poses.data(), d_poses, poses.size() * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
			return 1;
		for (size_t r = 0; r < requests.size(); ++r)
			for (uint32_t track = 0; track < max_tracks; ++track)
			{
				std::printf("%zu %u", r, track);
				for (int c = 0; c < 12; ++c)
				{
					uint32_t bits;
					std::memcpy(&bits, &poses[r * pose_floats + size_t(track) * 12 + c], 4);
					std::printf(" %08x", bits);
				}
				std::printf("\n");
			}
		cudaFree(d_requests);
		cudaFree(d_poses);
	}
	catch (const acl_b200::error& e)
	{
		std::fprintf(stderr, "%s\n", e.what());
		return e.status == ACLB200_ERR_NO_DEVICE ? 3 : 1;
	}
	return 0;
}
