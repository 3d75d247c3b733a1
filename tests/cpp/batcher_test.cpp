// Stress test of the multi-link batcher (include/mercury_b200.h: mercury_b200_batcher_*): L link threads each push K frames
// through the synchronous per-frame call, concurrently.
//   batcher_test mock <links> <frames per link> <max_batch> <max_wait_us>
//       CPU only: the batch function is a test double (payload byte i = checksum of the frame's samples + i, after a short
//       sleep), so every caller can verify it got ITS frame's result.  Build with -fsanitize=thread to race-check the machinery.
//   batcher_test gpu <ldpc_tables.bin> <config> <links> <frames per link> <max_batch> <max_wait_us>
//       real decode: frames synthesised by the library, results compared with one direct mercury_b200_demod_decode_batch call.
// Prints one summary line; exit code 0 = every result correct.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "mercury_b200.h"

static const size_t kMockFloats = 64, kMockBytes = 8;

static int mock_run(void *ctx, const float *x, size_t n, uint8_t *payload, mercury_b200_rx_stats *stats)
{
	std::atomic<long> *calls = static_cast<std::atomic<long> *>(ctx);
	calls->fetch_add(1);
	std::this_thread::sleep_for(std::chrono::microseconds(300));  // a "kernel"
	for (size_t f = 0; f < n; f++) {
		unsigned sum = 0;
		for (size_t i = 0; i < kMockFloats; i++) sum += (unsigned)x[f * kMockFloats + i];
		for (size_t i = 0; i < kMockBytes; i++) payload[f * kMockBytes + i] = (uint8_t)(sum + i);
		memset(&stats[f], 0, sizeof(stats[f]));
		stats[f].iterations_done = (int)(sum & 0xFFFF);
		stats[f].message_decoded = 1;
	}
	return MERCURY_B200_OK;
}

int main(int argc, char **argv)
{
	if (argc < 2) return 2;
	const bool mock = strcmp(argv[1], "mock") == 0;
	if ((mock && argc != 6) || (!mock && argc != 8)) {
		fprintf(stderr, "usage: %s mock L K max_batch max_wait_us | gpu tables config L K max_batch max_wait_us\n", argv[0]);
		return 2;
	}
	const int a0 = mock ? 2 : 4;
	const int L = atoi(argv[a0]), K = atoi(argv[a0 + 1]);
	const size_t max_batch = (size_t)atol(argv[a0 + 2]);
	const unsigned wait_us = (unsigned)atoi(argv[a0 + 3]);
	const size_t total = (size_t)L * K;

	mercury_b200_t *h = nullptr;
	mercury_b200_batcher_t *b = nullptr;
	std::atomic<long> mock_calls(0);
	size_t ffl = kMockFloats, fby = kMockBytes;
	std::vector<float> x;
	std::vector<uint8_t> want_pay;
	std::vector<mercury_b200_rx_stats> want_st;
	if (mock) {
		x.resize(total * ffl);
		for (size_t f = 0; f < total; f++)
			for (size_t i = 0; i < ffl; i++) x[f * ffl + i] = (float)((f * 7 + i * 3) % 251);
		if (mercury_b200_batcher_create_with_backend(ffl, fby, max_batch, wait_us, mock_run, &mock_calls, &b) != MERCURY_B200_OK) return 2;
	} else {
		int rc = mercury_b200_create(0, &h);
		if (rc != MERCURY_B200_OK) {
			fprintf(stderr, "mercury_b200_create: %s\n", mercury_b200_strerror(rc));
			return 3;
		}
		const int cfg = atoi(argv[3]);
		if (mercury_b200_load_tables(h, argv[2]) != MERCURY_B200_OK || mercury_b200_load_configuration(h, cfg, 50) != MERCURY_B200_OK) return 2;
		mercury_b200_geometry g;
		mercury_b200_get_geometry(h, &g);
		ffl = (size_t)g.Nsymb * MERCURY_B200_NOFDM * 2, fby = (size_t)g.frame_bytes;
		x.resize(total * ffl);
		std::vector<uint8_t> sent(total * fby);
		if (mercury_b200_synth_frames(argv[2], cfg, total, 77, 4.0, nullptr, x.data(), sent.data(), 8) != MERCURY_B200_OK) return 2;
		want_pay.resize(total * fby);
		want_st.resize(total);
		if (mercury_b200_demod_decode_batch(h, x.data(), total, want_pay.data(), want_st.data(), nullptr) != MERCURY_B200_OK) return 2;
		if (mercury_b200_batcher_create(h, max_batch, wait_us, &b) != MERCURY_B200_OK) return 2;
	}

	std::atomic<size_t> bad(0);
	const auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> links;
	for (int l = 0; l < L; l++)
		links.emplace_back([&, l]() {
			std::vector<uint8_t> pay(fby);
			mercury_b200_rx_stats st;
			for (int k = 0; k < K; k++) {
				const size_t f = (size_t)l * K + k;
				if (mercury_b200_batcher_receive_baseband(b, x.data() + f * ffl, pay.data(), &st) != MERCURY_B200_OK) {
					bad++;
					continue;
				}
				if (mock) {
					unsigned sum = 0;
					for (size_t i = 0; i < ffl; i++) sum += (unsigned)x[f * ffl + i];
					for (size_t i = 0; i < fby; i++)
						if (pay[i] != (uint8_t)(sum + i)) { bad++; break; }
					if (st.iterations_done != (int)(sum & 0xFFFF)) bad++;
				} else {
					if (memcmp(pay.data(), want_pay.data() + f * fby, fby) != 0 || memcmp(&st, &want_st[f], sizeof(st)) != 0) bad++;
				}
			}
		});
	for (auto &t : links) t.join();
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	uint64_t batches = 0, frames = 0, full = 0;
	mercury_b200_batcher_get_counters(b, &batches, &frames, &full);
	mercury_b200_batcher_destroy(b);
	if (h) mercury_b200_destroy(h);
	printf("links %d frames %zu batches %llu full %llu mean_batch %.1f bad %zu frames_per_s %.0f seconds %.3f\n", L, total,
	       (unsigned long long)batches, (unsigned long long)full, batches ? (double)frames / (double)batches : 0.0, (size_t)bad, total / secs, secs);
	return (bad == 0 && frames == total) ? 0 : 1;
}
