// The pass-band flavour of the multi-link batcher (mercury_b200_batcher_create_passband / _receive_byte): L link threads each make K
// synchronous receive_byte()-shaped calls with a whole capture buffer; every result must equal a direct
// mercury_b200_receive_byte_batch call over the same captures.
//   batcher_passband_test <ldpc_tables.bin> <config> <capture.f32: one capture holding a frame> <links> <calls per link> <max_batch> <max_wait_us>
// Link l's captures are the file's capture rotated by 37 * (l * K + k) samples (different delays, same frame).
// Prints one summary line; exit code 0 = every result correct and at least half the calls decoded.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "mercury_b200.h"

int main(int argc, char **argv)
{
	if (argc != 8) {
		fprintf(stderr, "usage: %s tables config capture.f32 L K max_batch max_wait_us\n", argv[0]);
		return 2;
	}
	const int cfg = atoi(argv[2]), L = atoi(argv[4]), K = atoi(argv[5]);
	const size_t max_batch = (size_t)atol(argv[6]);
	const unsigned wait_us = (unsigned)atoi(argv[7]);
	const size_t total = (size_t)L * K;
	mercury_b200_t *h = nullptr;
	int rc = mercury_b200_create(0, &h);
	if (rc != MERCURY_B200_OK) {
		fprintf(stderr, "mercury_b200_create: %s\n", mercury_b200_strerror(rc));
		return 3;
	}
	if (mercury_b200_load_tables(h, argv[1]) != MERCURY_B200_OK || mercury_b200_load_configuration(h, cfg, 50) != MERCURY_B200_OK) return 2;
	const size_t cs = (size_t)mercury_b200_get_capture_samples(h), fby = (size_t)mercury_b200_get_frame_size_bytes(h);
	std::vector<float> base(cs);
	FILE *f = fopen(argv[3], "rb");
	if (!f || fread(base.data(), sizeof(float), cs, f) != cs) return 2;
	fclose(f);
	std::vector<float> x(total * cs);
	for (size_t i = 0; i < total; i++) {
		const size_t shift = (37 * i) % 4000;
		for (size_t n = 0; n < cs; n++) x[i * cs + n] = base[(n + cs - shift) % cs];
	}
	std::vector<uint8_t> want_pay(total * fby);
	std::vector<mercury_b200_receive_stats> want_st(total);
	for (auto &s : want_st) memset(&s, 0, sizeof(s)), s.delay_of_last_decoded_message = -1;
	if (mercury_b200_receive_byte_batch(h, x.data(), MERCURY_B200_SAMPLES_F32, total, want_pay.data(), want_st.data(), nullptr) != MERCURY_B200_OK) return 2;
	mercury_b200_batcher_t *b = nullptr;
	if (mercury_b200_batcher_create_passband(h, MERCURY_B200_SAMPLES_F32, max_batch, wait_us, &b) != MERCURY_B200_OK) return 2;

	std::atomic<size_t> bad(0), decoded(0);
	const auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> links;
	for (int l = 0; l < L; l++)
		links.emplace_back([&, l]() {
			std::vector<uint8_t> pay(fby);
			for (int k = 0; k < K; k++) {
				const size_t i = (size_t)l * K + k;
				mercury_b200_receive_stats st;
				memset(&st, 0, sizeof(st));
				st.delay_of_last_decoded_message = -1;
				if (mercury_b200_batcher_receive_byte(b, x.data() + i * cs, pay.data(), &st) != MERCURY_B200_OK) {
					bad++;
					continue;
				}
				if (memcmp(pay.data(), want_pay.data() + i * fby, fby) != 0 || memcmp(&st, &want_st[i], sizeof(st)) != 0) bad++;
				if (st.message_decoded) decoded++;
			}
		});
	for (auto &t : links) t.join();
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	uint64_t batches = 0, frames = 0, full = 0;
	mercury_b200_batcher_get_counters(b, &batches, &frames, &full);
	mercury_b200_batcher_destroy(b);
	mercury_b200_destroy(h);
	printf("links %d calls %zu batches %llu mean_batch %.1f bad %zu decoded %zu calls_per_s %.0f seconds %.3f\n", L, total, (unsigned long long)batches,
	       batches ? (double)frames / (double)batches : 0.0, (size_t)bad, (size_t)decoded, total / secs, secs);
	return (bad == 0 && frames == total && decoded * 2 >= total) ? 0 : 1;
}
