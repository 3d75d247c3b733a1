// Exercises include/mercury_b200.hpp (the C++ mirror of cl_telecom_system for the RX tail) the way reference code would:
// construct, load_configuration, receive_byte / receive_bit on one synchronised frame, print what the reference prints.
//   usage: host_mirror_test <ldpc_tables.bin> <config> <ldpc iterations> <frame.bin: Nsymb*272 complex<double>> [capture.bin: pass-band doubles]
// With a capture file it also makes the reference's own call, receive_byte(double* data, int* out) on a whole pass-band capture, twice
// (the second call sees the link state the first one left, like consecutive calls on the reference object).
// Exit codes: 0 ok, 2 usage / IO, 3 no usable device (the library has no CPU fallback and says so).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mercury_b200.hpp"

int main(int argc, char **argv)
{
	if (argc != 5 && argc != 6) {
		fprintf(stderr, "usage: %s ldpc_tables.bin config iterations frame.bin\n", argv[0]);
		return 2;
	}
	try {
		mb200::cl_telecom_system telecom_system(0, argv[1]);
		telecom_system.default_configurations_telecom_system.ldpc_nIteration_max = atoi(argv[3]);
		telecom_system.load_configuration(atoi(argv[2]));
		const size_t n = (size_t)telecom_system.data_container.Nsymb * telecom_system.data_container.Nofdm;
		std::vector<std::complex<double>> baseband_data(n);
		FILE *f = fopen(argv[4], "rb");
		if (!f || fread(baseband_data.data(), sizeof(std::complex<double>), n, f) != n) {
			fprintf(stderr, "cannot read %zu samples from %s\n", n, argv[4]);
			return 2;
		}
		fclose(f);
		std::vector<int> out((size_t)telecom_system.get_frame_size_bytes());
		mb200::st_receive_stats st = telecom_system.receive_byte(baseband_data.data(), out.data());
		printf("decoded %d iterations %d crc %d all_zeros %d snr %.6f frame_bytes %d frame_bits %d\n", st.message_decoded, st.iterations_done,
		       st.crc, st.all_zeros, st.SNR, telecom_system.get_frame_size_bytes(), telecom_system.get_frame_size_bits());
		printf("bytes");
		for (int v : out) printf(" %d", v);
		printf("\n");
		std::vector<int> bits((size_t)(telecom_system.data_container.nBits - telecom_system.ldpc.P) / 8 * 8);
		telecom_system.receive_bit(baseband_data.data(), bits.data());
		printf("bits");
		for (int v : bits) printf(" %d", v);
		printf("\n");
		if (argc == 6) {
			const size_t cs = (size_t)telecom_system.data_container.Nofdm * telecom_system.data_container.buffer_Nsymb *
					  telecom_system.data_container.interpolation_rate;
			std::vector<double> data(cs);
			FILE *c = fopen(argv[5], "rb");
			if (!c || fread(data.data(), sizeof(double), cs, c) != cs) {
				fprintf(stderr, "cannot read %zu pass-band samples from %s\n", cs, argv[5]);
				return 2;
			}
			fclose(c);
			for (int call = 0; call < 2; call++) {
				std::vector<int> o2((size_t)telecom_system.get_frame_size_bytes());
				mb200::st_receive_stats r = telecom_system.receive_byte(data.data(), o2.data());
				printf("capture%d %d delay %d trials %d iterations %d crc %d freq %.9f metric %.17g last_delay %d\n", call, r.message_decoded, r.delay,
				       r.sync_trials, r.iterations_done, r.crc, r.freq_offset, r.coarse_metric, r.delay_of_last_decoded_message);
				printf("capture%d_bytes", call);
				for (int v : o2) printf(" %d", v);
				printf("\n");
			}
		}
		{  // the TX-side and ARQ-facing members compile and run with the reference's signatures
			std::vector<double> tx((size_t)mercury_b200_get_total_frame_size(telecom_system.handle()));
			std::vector<int> msg((size_t)telecom_system.get_frame_size_bytes(), 7);
			telecom_system.transmit_byte(msg.data(), (int)msg.size(), tx.data(), MERCURY_B200_NO_FILTER_MESSAGE);
			std::vector<double> pat(16 * 1088), filtered(16 * 1088);
			const int np = telecom_system.generate_ack_pattern_passband(pat.data());
			telecom_system.fir_tx_apply(pat.data(), filtered.data(), np);
			int matched = 0;
			std::vector<double> rxbuf(40 * 1088, 0.0);
			for (int i = 0; i < np; i++) rxbuf[5 * 1088 + i] = pat[(size_t)i];
			const double metric = telecom_system.detect_ack_pattern_from_passband(rxbuf.data(), (int)rxbuf.size(), &matched);
			printf("tx_side samples %zu pattern %d ack_metric %.3f matched %d break_metric %.3f config_for_5dB %d active_nsymb %d\n", tx.size(), np, metric, matched,
			       telecom_system.detect_break_pattern_from_passband(rxbuf.data(), (int)rxbuf.size()), (int)telecom_system.get_configuration(5.0),
			       telecom_system.get_active_nsymb());
		}
		telecom_system.load_configuration(99);  // ignored, like the reference (telecom_system.cc:2494-2497)
		printf("after_bad_config frame_bytes %d\n", telecom_system.get_frame_size_bytes());
	} catch (const std::exception &e) {
		fprintf(stderr, "mercury_b200: %s\n", e.what());
		return 3;
	}
	return 0;
}
