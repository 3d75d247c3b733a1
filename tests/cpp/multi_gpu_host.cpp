// tests/cpp/multi_gpu_host.cpp -- the multi-GPU RX path driven from C++ host code only (north_star: "host code stays C/C++"; no Python,
// no torch): ONE process, one handle + one host thread per GPU, the table blob built on rank 0 and broadcast over NCCL
// (mercury_b200_broadcast_tables), then every rank decodes its contiguous shard of the batch with no further collective.
//
//   g++ -std=c++14 -O2 -I include tests/cpp/multi_gpu_host.cpp -o multi_gpu_host -L mercury_b200 -lmercury_b200 -lnccl -lcudart -pthread
//   ./multi_gpu_host <ldpc_tables.bin> <gpus> <frames per gpu> <steps> [config=8] [format: c64 | i16]
//
// Prints one JSON line: whole-job frames/s through the host-buffer batch call (pinned memory; H2D + kernels + D2H timed, max over
// ranks, `steps` passes after one warm-up), frames decoded, payload mismatches against the transmitted payloads -- summed over ALL ranks.
// Exit code 0 only if every decoded payload of every rank is the transmitted one.
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "mercury_b200.h"

struct Rank {
	int dev = 0, rc = 0;
	mercury_b200_t *h = nullptr;
	ncclComm_t comm = nullptr;
	size_t lo = 0, hi = 0;            // shard of the batch: frames [lo, hi)
	double seconds = 0;
	long decoded = 0, mismatches = 0;
	std::string err;
};

// contiguous shards whose sizes differ by at most one frame (mercury_b200/dist.py shard_range)
static void shard_range(size_t n, int rank, int world, size_t *lo, size_t *hi)
{
	const size_t base = n / world, rem = n % world;
	*lo = rank * base + std::min<size_t>(rank, rem);
	*hi = *lo + base + (rank < (int)rem ? 1 : 0);
}

int main(int argc, char **argv)
{
	if (argc < 5) return fprintf(stderr, "usage: %s ldpc_tables.bin gpus frames_per_gpu steps [config] [c64|i16]\n", argv[0]), 2;
	const char *tables = argv[1];
	int world = atoi(argv[2]);
	const size_t per_gpu = (size_t)atoll(argv[3]);
	const int steps = atoi(argv[4]);
	const int config = argc > 5 ? atoi(argv[5]) : 8;
	const bool i16 = argc > 6 && !strcmp(argv[6], "i16");
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) return fprintf(stderr, "no usable CUDA device (no CPU fallback exists)\n"), 3;
	if (world < 1 || world > n_dev) world = n_dev;
	const size_t n_total = per_gpu * (size_t)world;

	// input synthesis on the host (outside every timed region): `unique` distinct noisy frames, tiled over the batch
	mercury_b200_t *probe = nullptr;
	if (mercury_b200_create(0, &probe) != MERCURY_B200_OK || mercury_b200_load_tables(probe, tables) != MERCURY_B200_OK ||
	    mercury_b200_load_configuration(probe, config, 50) != MERCURY_B200_OK)
		return fprintf(stderr, "cannot set up the library on device 0\n"), 3;
	mercury_b200_geometry g;
	mercury_b200_get_geometry(probe, &g);
	mercury_b200_destroy(probe);
	const size_t unique = std::min<size_t>(4096, per_gpu), frame_floats = (size_t)g.Nsymb * g.Nofdm * 2;
	std::vector<float> clean(unique * frame_floats);
	std::vector<uint8_t> pay_u(unique * (size_t)g.frame_bytes);
	const double esn0 = config == 8 ? 2.5 : 30.0;  // mode 8: threshold + 2 dB like bench.py; other modes: light noise
	if (mercury_b200_synth_frames(tables, config, unique, 0x4D455243ull, esn0, nullptr, clean.data(), pay_u.data(), 8) != MERCURY_B200_OK)
		return fprintf(stderr, "synth_frames failed\n"), 1;
	float peak = 0.f;
	for (float v : clean) peak = std::max(peak, std::fabs(v));
	const float scale = peak / 32000.0f;

	std::vector<ncclComm_t> comms(world);
	std::vector<int> devs(world);
	for (int i = 0; i < world; i++) devs[i] = i;
	if (ncclCommInitAll(comms.data(), world, devs.data()) != ncclSuccess) return fprintf(stderr, "ncclCommInitAll failed\n"), 1;

	std::vector<Rank> ranks(world);
	std::vector<std::thread> threads;
	for (int r = 0; r < world; r++) {
		ranks[r].dev = r, ranks[r].comm = comms[r];
		shard_range(n_total, r, world, &ranks[r].lo, &ranks[r].hi);
		threads.emplace_back([&, r]() {
			Rank &k = ranks[r];
			auto fail = [&](const char *what) { k.rc = 1, k.err = std::string(what) + ": " + (k.h ? mercury_b200_last_error(k.h) : "?"); };
			cudaSetDevice(k.dev);
			if (mercury_b200_create(k.dev, &k.h) != MERCURY_B200_OK) return fail("create");
			if (r == 0 && mercury_b200_load_tables(k.h, tables) != MERCURY_B200_OK) return fail("load_tables");  // ONLY rank 0 reads the table file
			if (mercury_b200_broadcast_tables(k.h, k.comm, 0, nullptr) != MERCURY_B200_OK) return fail("broadcast_tables");
			if (mercury_b200_load_configuration(k.h, config, 50) != MERCURY_B200_OK) return fail("load_configuration");
			const size_t n = k.hi - k.lo, esz = i16 ? sizeof(int16_t) : sizeof(float);
			void *x = mercury_b200_host_alloc(n * frame_floats * esz);
			uint8_t *pay = static_cast<uint8_t *>(mercury_b200_host_alloc(n * (size_t)g.frame_bytes));
			mercury_b200_rx_stats *st = static_cast<mercury_b200_rx_stats *>(mercury_b200_host_alloc(n * sizeof(mercury_b200_rx_stats)));
			if (!x || !pay || !st) return fail("host_alloc");
			for (size_t f = 0; f < n; f++) {  // this rank's frames of the global batch: frame index (lo + f) -> unique frame (lo + f) % unique
				const float *src = clean.data() + ((k.lo + f) % unique) * frame_floats;
				if (i16) {
					int16_t *dst = static_cast<int16_t *>(x) + f * frame_floats;
					for (size_t i = 0; i < frame_floats; i++) dst[i] = (int16_t)std::lrintf(src[i] / scale);
				} else
					memcpy(static_cast<float *>(x) + f * frame_floats, src, frame_floats * sizeof(float));
			}
			auto run = [&]() {
				return mercury_b200_demod_decode_batch_fmt(k.h, x, i16 ? MERCURY_B200_BASEBAND_CI16 : MERCURY_B200_BASEBAND_C64, scale, n, pay, st, nullptr);
			};
			if (run() != MERCURY_B200_OK) return fail("demod_decode_batch (warm-up)");
			const auto t0 = std::chrono::steady_clock::now();
			for (int s = 0; s < steps; s++)
				if (run() != MERCURY_B200_OK) return fail("demod_decode_batch");
			k.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			for (size_t f = 0; f < n; f++)
				if (st[f].message_decoded == 1) {
					k.decoded++;
					if (memcmp(pay + f * g.frame_bytes, pay_u.data() + ((k.lo + f) % unique) * g.frame_bytes, g.frame_bytes)) k.mismatches++;
				}
			mercury_b200_host_free(x), mercury_b200_host_free(pay), mercury_b200_host_free(st);
			mercury_b200_destroy(k.h);
		});
	}
	for (auto &t : threads) t.join();
	for (int r = 0; r < world; r++) ncclCommDestroy(comms[r]);
	double worst = 0;
	long decoded = 0, mism = 0;
	for (const Rank &k : ranks) {
		if (k.rc) return fprintf(stderr, "rank %d: %s\n", k.dev, k.err.c_str()), 1;
		worst = std::max(worst, k.seconds), decoded += k.decoded, mism += k.mismatches;
	}
	printf("{\"program\": \"tests/cpp/multi_gpu_host.cpp\", \"n_gpus\": %d, \"config\": %d, \"frames_total\": %zu, \"frames_per_gpu\": %zu, \"steps\": %d, "
	       "\"sample_format\": \"%s\", \"e2e_frames_per_s\": %.1f, \"frames_decoded\": %ld, \"payload_mismatches\": %ld, "
	       "\"tables\": \"built on rank 0, mercury_b200_broadcast_tables over NCCL to the other handles\"}\n",
	       world, config, n_total, per_gpu, steps, i16 ? "complex int16" : "complex64", (double)n_total * steps / worst, decoded, mism);
	return mism == 0 && decoded > 0 ? 0 : 1;
}
