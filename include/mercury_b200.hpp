// mercury_b200.hpp -- C++ host-side mirror of the reference's physical-layer object for the RX tail, over the C ABI.
//
// The reference is C++ (g++ -std=c++14) and reaches its physical layer through members of one long-lived
// cl_telecom_system object (reference include/physical_layer/telecom_system.h:85-199).  This header-only class keeps the
// member names, argument meaning and error behaviour of that object for the path this library accelerates, so that
// code written against the reference reads the same:
//
//   reference member                                   here
//   -------------------------------------------------  ------------------------------------------------------------
//   void load_configuration(int)        (.h:176)        load_configuration(int)          -- O(1), all 17 modes resident
//   default_configurations_telecom_system
//       .ldpc_nIteration_max            (main.cc:550)   default_configurations_telecom_system.ldpc_nIteration_max
//   int get_frame_size_bytes()/bits()   (.h:180-181)    get_frame_size_bytes() / get_frame_size_bits()
//   st_receive_stats receive_byte(double*, int*)        receive_byte(double* data, int* out)  -- THE reference signature, whole: `data` is the
//                                       (.h:142)            pass-band capture (data_container.Nofdm * buffer_Nsymb * interpolation_rate
//                                                           doubles); front-end + tail on the GPU (.cc:646-1518, OFDM branch)
//                                                       receive_byte(const std::complex<double>* baseband_data, int* out)
//                                       (.h:142)            = the tail of receive_byte (.cc:1132-1429) on the
//                                                             post-synchronisation data_container.baseband_data
//   st_receive_stats receive_bit(double*, int*) (.h:139) receive_bit(const std::complex<double>*, int* out)
//   st_receive_stats receive_stats      (.h:114)        receive_stats (the fields the tail writes)
//   data_container.{Nsymb,Nofdm,nBits,preamble_nSymb}   data_container.{...} (read-only copies), ldpc.{N,K,P}
//   void transmit_byte(int*, int, double*, int) (.h:138) transmit_byte(...) SINGLE_MESSAGE / NO_FILTER_MESSAGE; fir_tx_apply(); ofdm.passband_start_sample
//   set_mfsk_ctrl_mode / get_active_nsymb / get_configuration / measure_signal_only / generate_* / detect_*_pattern_*_passband  -- same names
//   -                                                   receive_byte_batch(): many synchronised frames per call
//
// Error behaviour follows the reference: a frame that does not decode is RETURNED (message_decoded == NO, SNR == -99.9),
// never thrown.  What the reference handles with exit(1) (no usable device here, missing table file) throws
// std::runtime_error from the constructor instead; there is no CPU fallback behind any call.
//
// Pure host C++ (>= C++11): needs only mercury_b200.h and libmercury_b200.so, no CUDA headers.
#ifndef MERCURY_B200_HPP
#define MERCURY_B200_HPP

#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mercury_b200.h"

namespace mb200 {

enum { NO = 0, YES = 1 };  // reference: include/physical_layer/physical_defines.h

// The st_receive_stats fields written by telecom_system.cc:1310-1429 (same names; front-end fields are not produced here).
struct st_receive_stats {
	int iterations_done = 0;   // 0..I, I+1 = not converged, -1 = decode skipped by the mean|H| < 0.3 gate (:1271)
	int delay = 0;             // front-end fields (telecom_system.h:65-82), written by receive_byte(double*, int*)
	int delay_of_last_decoded_message = -1;
	double freq_offset = 0;
	double freq_offset_of_last_decoded_message = 0;
	double signal_stregth_dbm = 0;
	double coarse_metric = 0;
	int sync_trials = 0;       // incremented on a failed decode like :1358
	int message_decoded = NO;
	double SNR = -99.9;
	int crc = 0;
	int all_zeros = NO;
	int mfsk_search_raw = 0;         // MFSK anti-re-decode: first symbol of the tone-preamble search before the nUnder adjustment (telecom_system.h:79, .cc:684)
	int frame_overflow_symbols = 0;  // > 0: the MFSK frame runs past the end of the capture by this many symbols (telecom_system.h:80, .cc:702-715)
	float variance = 0.f;      // pilot noise variance behind the LLRs (:1291)
	float mean_H = 0.f;        // mean |H| over pilots after AGC (:1225-1244)
};

class cl_telecom_system {
public:
	struct {
		int ldpc_nIteration_max = 50;  // "-I n" (main.cc:303-311,547-575); applied by load_configuration()
	} default_configurations_telecom_system;
	struct {
		int Nsymb = 0, Nofdm = MERCURY_B200_NOFDM, Nc = 0, nBits = 0, nData = 0, preamble_nSymb = 0;
		int buffer_Nsymb = 0, interpolation_rate = 4;  // capture = Nofdm * buffer_Nsymb * interpolation_rate samples (data_container.cc:133-153)
		int nUnder_processing_events = 0;              // written by the caller's audio side; receive_byte subtracts it from mfsk_search_raw (.cc:684)
	} data_container;
	struct {
		int N = MERCURY_B200_N, K = 0, P = 0;
	} ldpc;
	st_receive_stats receive_stats;
	int M = 0;
	int mfsk_fixed_delay = -1;  // >= 0: the next receive_byte() skips the tone-preamble search and uses this delay, once (telecom_system.h:110, .cc:663-673;
	                            // the ARQ layer's overflow recapture sets it, arq_common.cc:2830-2833)

	explicit cl_telecom_system(int device = 0, const std::string &ldpc_table_path = "")
	{
		int rc = mercury_b200_create(device, &h_);
		if (rc != MERCURY_B200_OK) throw std::runtime_error(std::string("mercury_b200_create: ") + mercury_b200_strerror(rc));
		rc = mercury_b200_load_tables(h_, ldpc_table_path.c_str());
		if (rc != MERCURY_B200_OK) {
			const std::string msg = std::string("mercury_b200_load_tables: ") + mercury_b200_last_error(h_);
			mercury_b200_destroy(h_);
			h_ = nullptr;
			throw std::runtime_error(msg);
		}
	}
	~cl_telecom_system() { mercury_b200_destroy(h_); }
	cl_telecom_system(const cl_telecom_system &) = delete;
	cl_telecom_system &operator=(const cl_telecom_system &) = delete;

	// reference: out-of-range configurations are silently ignored (telecom_system.cc:2494-2497); so they are here
	void load_configuration(int configuration)
	{
		if (!(configuration >= 100 && configuration <= 102) && (configuration < 0 || configuration >= MERCURY_B200_NUM_CONFIGS)) return;
		if (mercury_b200_load_configuration(h_, configuration, default_configurations_telecom_system.ldpc_nIteration_max) != MERCURY_B200_OK) return;
		mercury_b200_geometry g;
		mercury_b200_get_geometry(h_, &g);
		data_container.Nsymb = g.Nsymb, data_container.Nc = g.Nc, data_container.nBits = g.nBits, data_container.nData = g.nData;
		data_container.preamble_nSymb = g.preamble_nSymb;
		data_container.buffer_Nsymb = mercury_b200_get_capture_samples(h_) / (data_container.Nofdm * data_container.interpolation_rate);
		ldpc.K = g.K, ldpc.P = g.P;
		M = g.M;
		nReal_ = g.nReal;
		if (M == 200 && ofdm.passband_start_sample == 1088) ofdm.passband_start_sample = 0;  // MFSK: no pre-equalisation pass moved the counter
	}
	int get_frame_size_bytes() const { return mercury_b200_get_frame_size_bytes(h_); }
	int get_frame_size_bits() const { return mercury_b200_get_frame_size_bits(h_); }

	// The reference's own call (telecom_system.h:142): data = one pass-band capture buffer, out = one int per payload byte.  The link
	// state the reference keeps in receive_stats between calls (delay / frequency offset of the last decoded message) lives in this
	// object's receive_stats the same way.
	st_receive_stats receive_byte(double *data, int *out)
	{
		mercury_b200_receive_stats rs = mercury_b200_receive_stats();
		rs.delay_of_last_decoded_message = receive_stats.delay_of_last_decoded_message;
		rs.freq_offset_of_last_decoded_message = receive_stats.freq_offset_of_last_decoded_message;
		rs.mfsk_search_or_overflow = mfsk_fixed_delay >= 0 ? MERCURY_B200_MFSK_FIXED_DELAY(mfsk_fixed_delay)
		                                                   : receive_stats.mfsk_search_raw - data_container.nUnder_processing_events;
		mfsk_fixed_delay = -1;  // consumed (.cc:669)
		const int rc = mercury_b200_receive_byte(h_, data, out, &rs);
		if (rc != MERCURY_B200_OK) throw std::runtime_error(std::string("mercury_b200_receive_byte: ") + mercury_b200_last_error(h_));
		receive_stats.iterations_done = rs.iterations_done, receive_stats.delay = rs.delay, receive_stats.sync_trials = rs.sync_trials;
		receive_stats.delay_of_last_decoded_message = rs.delay_of_last_decoded_message;
		receive_stats.freq_offset = rs.freq_offset, receive_stats.freq_offset_of_last_decoded_message = rs.freq_offset_of_last_decoded_message;
		receive_stats.message_decoded = rs.message_decoded ? YES : NO, receive_stats.SNR = rs.SNR;
		receive_stats.crc = rs.crc, receive_stats.all_zeros = rs.all_zeros;
		receive_stats.signal_stregth_dbm = rs.signal_stregth_dbm, receive_stats.coarse_metric = rs.coarse_metric;
		receive_stats.frame_overflow_symbols = rs.mfsk_search_or_overflow;  // 0 in OFDM configurations (.cc:654)
		return receive_stats;
	}

	// ---- the rest of what the datalink layer calls on the object (source/datalink_layer/arq_*.cc) -------------------------------------
	// void transmit_byte(int* data, int nBytes, double* out, int message_location) (telecom_system.h:138): SINGLE_MESSAGE (3) or
	// NO_FILTER_MESSAGE (4, the ARQ layer's choice, arq_common.cc:2224); the running carrier counter lives in ofdm.passband_start_sample
	struct {
		uint64_t passband_start_sample = 1088;  // where a freshly initialised reference object stands in OFDM modes (0 in MFSK modes)
	} ofdm;
	void transmit_byte(int *data, int nBytes, double *out, int message_location = MERCURY_B200_SINGLE_MESSAGE)
	{
		std::vector<uint8_t> pl((size_t)get_frame_size_bytes(), 0);
		for (int i = 0; i < nBytes && i < (int)pl.size(); i++) pl[(size_t)i] = (uint8_t)data[i];
		const uint64_t start = ofdm.passband_start_sample;
		const int rc = mercury_b200_transmit_byte_batch_ex(h_, pl.data(), &start, 1, out, MERCURY_B200_SAMPLES_F64, message_location, nullptr);
		if (rc != MERCURY_B200_OK) throw std::runtime_error(std::string("mercury_b200_transmit_byte: ") + mercury_b200_last_error(h_));
		ofdm.passband_start_sample = start + (uint64_t)mercury_b200_get_total_frame_size(h_);
	}
	// ofdm.FIR_tx1.apply + ofdm.FIR_tx2.apply over a padded batch of NO_FILTER frames (arq_common.cc:2243-2246)
	void fir_tx_apply(const double *in, double *out, int nItems)
	{
		if (mercury_b200_fir_tx_apply(h_, in, (size_t)nItems, out) != MERCURY_B200_OK) throw std::runtime_error(mercury_b200_last_error(h_));
	}
	void set_mfsk_ctrl_mode(bool enable) { mercury_b200_set_mfsk_ctrl_mode(h_, enable ? 1 : 0); }
	// g_gui_state.coarse_freq_sync_enabled (a global of the reference's GUI state, read by receive_byte(): telecom_system.cc:949)
	void set_coarse_freq_sync(bool enable) { mercury_b200_set_coarse_freq_sync(h_, enable ? 1 : 0); }
	int get_active_nsymb() const { return mercury_b200_get_active_nsymb(h_); }
	char get_configuration(double SNR) const { return (char)mercury_b200_get_configuration(SNR); }
	double measure_signal_only(double *data)
	{
		double dbm = 0;
		if (mercury_b200_measure_signal_only_batch(h_, data, MERCURY_B200_SAMPLES_F64, 1, &dbm) != MERCURY_B200_OK) throw std::runtime_error(mercury_b200_last_error(h_));
		receive_stats.signal_stregth_dbm = dbm;
		return dbm;
	}
	int generate_ack_pattern_passband(double *out) { return pattern_tx(0, out); }
	int generate_break_pattern_passband(double *out) { return pattern_tx(1, out); }
	double detect_ack_pattern_from_passband(double *data, int size, int *out_matched = nullptr) { return pattern_rx(0, data, size, out_matched); }
	double detect_break_pattern_from_passband(double *data, int size, int *out_matched = nullptr) { return pattern_rx(1, data, size, out_matched); }

	// baseband_data: Nsymb * Nofdm samples, the frame's data symbols after time/frequency synchronisation (the reference
	// indexes data_container.baseband_data at (preamble_nSymb + i) * Nofdm, telecom_system.cc:1137).  out: one int per byte.
	st_receive_stats receive_byte(const std::complex<double> *baseband_data, int *out)
	{
		mercury_b200_rx_stats rs;
		const int rc = mercury_b200_receive_baseband(h_, reinterpret_cast<const double *>(baseband_data), out, &rs);
		if (rc != MERCURY_B200_OK) throw std::runtime_error(std::string("mercury_b200_receive_baseband: ") + mercury_b200_last_error(h_));
		receive_stats.iterations_done = rs.iterations_done;
		receive_stats.crc = rs.crc;
		receive_stats.all_zeros = rs.all_zeros;
		receive_stats.message_decoded = rs.message_decoded ? YES : NO;
		receive_stats.SNR = rs.message_decoded ? (double)rs.SNR : -99.9;  // :1347
		receive_stats.variance = rs.variance;
		receive_stats.mean_H = rs.mean_H;
		if (!rs.message_decoded) receive_stats.sync_trials++;  // :1358
		return receive_stats;
	}
	// reference receive_bit (telecom_system.cc:636-644): the bytes of the frame INCLUDING its two CRC bytes, LSB first.
	// The CRC bytes are recomputed from the payload (identical to the received ones whenever the self check passed).
	st_receive_stats receive_bit(const std::complex<double> *baseband_data, int *out)
	{
		std::vector<int> bytes((size_t)nReal_ / 8, 0);
		const st_receive_stats st = receive_byte(baseband_data, bytes.data());
		const int fb = get_frame_size_bytes();
		unsigned crc = 0xFFFF;  // CRC16_MODBUS_RTU_calc, crc16_modbus_rtu.cc:25-45
		for (int i = 0; i < fb; i++) {
			crc ^= (unsigned)bytes[(size_t)i] & 0xFF;
			for (int b = 0; b < 8; b++) crc = (crc & 1) ? (crc >> 1) ^ 0xA001 : crc >> 1;
		}
		if (st.message_decoded == YES && fb + 1 < (int)bytes.size()) bytes[(size_t)fb] = crc & 0xFF, bytes[(size_t)fb + 1] = crc >> 8;
		for (size_t i = 0; i < bytes.size(); i++)  // byte_to_bit, misc.cc:93-105
			for (int j = 0; j < 8; j++) out[i * 8 + j] = (bytes[i] >> j) & 1;
		return st;
	}
	// Many independent, already synchronised frames in one call (north_star path). baseband: n x Nsymb x Nofdm complex<float>,
	// payload: n x get_frame_size_bytes(), stats: n records of the C ABI.
	void receive_byte_batch(const std::complex<float> *baseband, size_t n_frames, uint8_t *payload, mercury_b200_rx_stats *stats)
	{
		const int rc = mercury_b200_demod_decode_batch(h_, reinterpret_cast<const float *>(baseband), n_frames, payload, stats, nullptr);
		if (rc != MERCURY_B200_OK) throw std::runtime_error(std::string("mercury_b200_demod_decode_batch: ") + mercury_b200_last_error(h_));
	}
	mercury_b200_t *handle() { return h_; }

private:
	int pattern_tx(int brk, double *out)
	{
		uint64_t c = ofdm.passband_start_sample;
		const int n = mercury_b200_generate_pattern_passband(h_, brk, out, &c);
		if (n < 0) throw std::runtime_error(mercury_b200_last_error(h_));
		ofdm.passband_start_sample = c;
		return n;
	}
	double pattern_rx(int brk, double *data, int size, int *out_matched)
	{
		mercury_b200_mfsk_pattern_result r;
		if (mercury_b200_detect_patterns_from_passband_batch(h_, data, MERCURY_B200_SAMPLES_F64, 1, size, &r) != MERCURY_B200_OK)
			throw std::runtime_error(mercury_b200_last_error(h_));
		if (out_matched) *out_matched = brk ? r.break_matched : r.ack_matched;
		return brk ? r.break_metric : r.ack_metric;
	}
	mercury_b200_t *h_ = nullptr;
	int nReal_ = 0;
};

}  // namespace mb200
#endif  // MERCURY_B200_HPP
