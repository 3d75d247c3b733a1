// mb_kernels.cuh -- launch interfaces of the two sm_100a kernels of the RX path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "mb_tables.h"

#define MB_HANDOFF_STRIDE 2400  // floats per frame of the stage hand-off buffer (= MERCURY_B200_HANDOFF_FLOATS): 1600 LLRs + ZF-mode data symbols

// Mirrors mercury_b200_rx_stats (include/mercury_b200.h); 32 bytes.
struct MbRxStats {
	int32_t iterations_done, crc, all_zeros, message_decoded;
	float SNR, variance, mean_H;
	int32_t reserved;
};

struct MbDemodArgs {
	const float2 *x;       // [B][Nsymb][sym_stride] complex64 baseband, preamble stripped; the 256 useful samples start at sym_skip
	int32_t sym_stride;    // 272 = symbols as delivered (guard interval present, skipped: sym_skip 16); 256 = guard interval already removed
	int32_t sym_skip;
	int32_t x_format;      // 0: complex64; 1: complex int16 (re, im; value = int x x_scale); 2: complex fp16 -- MERCURY_B200_BASEBAND_*
	float x_scale;
	float *llr;            // [B][1600] LLRs in the hand-off layout (internal variable order, 32-float rows rotated: MB_HANDOFF)
	float *llr_cw;         // optional [B][1600] LLRs in codeword order (parity output)
	MbRxStats *stats;      // [B]
	float2 *dbg_Y, *dbg_H, *dbg_Z;  // optional [B][Nsymb*50] stage captures
	const uint8_t *blob;   // device copy of the table blob
	uint32_t off_twiddle;
	uint32_t off_var_of_cw;
	unsigned long long n_frames;  // filled by mb_launch_demod (the kernel is persistent: grid < frames)
	MbMode mode;
};

struct MbLdpcArgs {
	const float *llr;      // [B][1600] hand-off layout (MB_HANDOFF)
	uint8_t *payload;      // [B][frame_bytes]
	MbRxStats *stats;      // [B]  (SNR/variance/mean_H already filled by the demod kernel)
	const uint8_t *blob;
	MbMode mode;
	MbRate rate;
	int32_t max_iters;
	int32_t check_gate;    // 1: honour the mean|H| < 0.3 gate recorded by the demod kernel
	int32_t cheap_test_threads;  // run the syndrome-only test after an iteration that started with <= this many unhappy threads
	unsigned *queue;       // [2] device words, zero between launches: next frame of the batch, CTAs finished (the kernel re-arms them)
	float *lch_scratch;    // [mb_ldpc_max_ctas()][2 * 1600] per-CTA channel-LLR scratch (same stream as the queue)
	unsigned long long n_frames;  // filled by mb_launch_ldpc (the kernel is persistent: resident CTAs pull frames from the queue)
	// filled by mb_launch_ldpc: blob + rate.off_edge_varb / off_vedgeb / off_vtail (the kernel re-derives a pointer per task otherwise)
	const uint16_t *edge_var;
	const uint16_t *vedge;
	const uint32_t *vtail;
};

size_t mb_ldpc_smem_bytes(int c_slots);

cudaError_t mb_launch_demod(const MbDemodArgs &a, size_t n_frames, cudaStream_t stream);
cudaError_t mb_launch_ldpc(const MbLdpcArgs &a, size_t n_frames, int algo, cudaStream_t stream);
cudaError_t mb_demod_init();  // opt-in shared memory attributes, occupancy of every instantiation
int mb_demod_ctas_per_sm(int Nsymb, int M, int estimator, int phase_only);  // resident CTAs per SM (0 = no such instantiation)
cudaError_t mb_ldpc_init();
size_t mb_ldpc_max_ctas();  // largest grid mb_launch_ldpc uses on this device (sizes lch_scratch)
int mb_ldpc_ctas_per_sm(int algo, int rate_idx, int rate_num, int c_slots);  // resident decoder CTAs (frame pairs) per SM

// ---------------------------------------------------------------------------------------------------------------------
// RX front-end (mb_frontend.cu): pass-band capture buffers -> synchronised base-band frames for the tail.
// ---------------------------------------------------------------------------------------------------------------------
#define MB_FE_SYM 1088   // pass-band samples per OFDM symbol: Nofdm 272 x frequency_interpolation_rate 4
#define MB_FE_TAPS 33    // both receive FIRs: (int)(4 / (3000 / 24000)) = 32 -> odd 33 (fir_filter.cc:56-61, physical_config.cc:93-101)

enum MbFePhase {
	MB_FE_COARSE_WAIT = 0,  // waiting for the full-buffer coarse Schmidl-Cox run
	MB_FE_REC_BOUNDS,       // waiting for the re-run of the bounds recovery   (telecom_system.cc:734-798)
	MB_FE_GATE,             // energy / metric gates                           (:800-851)
	MB_FE_REC_SILENCE,      // waiting for the re-run of the silence skip      (:861-924)
	MB_FE_TRIAL,            // head of the trial loop                          (:931-1019)
	MB_FE_FINE_WAIT,        // waiting for the fine Schmidl-Cox run
	MB_FE_POSTDELAY,        // clamp + post-fine-sync energy fix               (:1020-1069)
	MB_FE_EXTRACT,          // waiting for k_fe_extract
	MB_FE_TAIL_WAIT,        // waiting for the tail's verdict
	MB_FE_TRIALS_END,       // SKIP-H recovery check                           (:1436-1504)
	MB_FE_REC_SKIPH,
	MB_FE_CFS_WAIT,         // waiting for one of the three coarse-frequency runs of trial 1 (:949-1013, optional)
	MB_FE_DONE
};

struct MbFeConst {
	double c_ts[MB_FE_TAPS + 1], c_data[MB_FE_TAPS + 1];  // FIR_rx_time_sync, FIR_rx_data
	double fs, fc, amp, bandwidth, Ts, ignore_limit;
	int32_t trials_max, use_last_time, use_last_freq, pad;
};

// Per-capture state: st_receive_stats fields (telecom_system.h:63-82) + the locals of receive_byte() + what the capture waits for.
struct MbFeState {
	double last_freq, SNR, freq_offset, coarse_metric, signal_dbm, cur_f, freq_offset_measured;
	int32_t last_delay, delay, sync_trials, message_decoded, iterations_done, crc, all_zeros;
	int32_t phase, pream_symb_loc, skip_h_count, skip_h_recovery_attempted;
	int32_t cur_kind;    // what baseband_data_interpolated holds: 0 = time-sync filter at fc (materialised), 1 = data filter at cur_f (on demand), 2 = time-sync filter at cur_f (on demand: after the coarse frequency search)
	int32_t sc_pending, sc_src, sc_start, sc_size, sc_step, sc_npos;  // the Schmidl-Cox run this capture waits for (sc_src: 0 the time-sync base-band, 1 a window of the data-filter base-band at cur_f, 2 a window of the time-sync filter at win_f)
	int32_t slot;        // tail slot of the running trial
	int32_t extract_pending;  // k_fe_moose has chosen the carrier, k_fe_extract_tiles still has to write the frame
	int32_t sc_from;     // the pending run's location_to_return: only positions >= it can be selected (ofdm.cc:1821-1823, 1946-1958)
	int32_t sc_pad;
	unsigned long long sc_max_key;  // approximate maximum of the pending run (pass A of the two-pass Schmidl-Cox), order-preserving key
	// the optional coarse frequency search of trial 1 (telecom_system.cc:949-1013; g_gui_state.coarse_freq_sync_enabled)
	double coarse_off;   // coarse_freq_offset: 0 or +-30 Hz once the search has applied one; effective carrier = fc + coarse_off
	double win_f;        // carrier of the time-sync-filter window a pending run of source 2 is computed at
	double cfs_best, cfs_zero, cfs_best_off;  // best_correlation, zero_hz_correlation, best_offset
	int32_t cfs_i, cfs_best_delay;            // which of {-30, 0, +30} Hz is running; best_delay
};

struct MbFeArgs {
	const void *x;       // [n][buf] pass-band samples
	int32_t x_format /* MERCURY_B200_SAMPLES_*: 0 f64, 1 f32, 2 i16, 3 i32 */, n, buf, pre, S, buffer_Nsymb, frame_bytes;
	int32_t coarse_freq_sync;  // g_gui_state.coarse_freq_sync_enabled: the +-30 Hz search before trial 1 (win / pref_win then hold (2 pre + S) symbols)
	const double2 *carrier;  // [>= buf] (cos, sin)(2 pi fc i Ts), host libm
	MbFeState *st;       // [n]
	double2 *bbi;        // [n][buf]  time-sync base-band
	double *energy_part; // [n][ceil(buf/1024)]
	double2 *win;        // [n][win_stride] fine-sync window of the data-filter base-band
	int32_t win_stride;
	double *vals;        // [n][vals_stride] correlation metrics of the pending run
	int32_t vals_stride;
	uint8_t *flags;      // [n][vals_stride] positions whose norms sit on the 0.001 threshold (forced into the exact pass)
	double *pref_ts;     // [n][3][buf / 4 + 1]    exclusive prefix sums over the time-sync base-band, one entry per 4 samples: |w|^2, lag-1024 and lag-512 dot products
	double *tile_base;   // [n][ceil(buf / 1024) + 1][3] per-tile bases of pref_ts (its entries are tile-local)
	double *pref_win;    // [n][3][win_stride + 1] the same at full resolution over the window of the pending fine run
	float2 *frames;      // [n][S][272] tail input, by slot
	double2 *dbg_bb;     // optional [n][(pre+S)*272] fp64 copy of baseband_data (by capture)
	const MbRxStats *tail_stats;   // [n] by slot
	const uint8_t *tail_payload;   // [n][tail_payload_stride] by slot
	int32_t tail_payload_stride;
	uint8_t *payload_out;          // [n][frame_bytes] by capture
	int32_t *counters;   // [4]: tail slots handed out, captures not done, captures waiting for a Schmidl-Cox run, exact (pass B) evaluations
};

// Mirrors mercury_b200_receive_stats (include/mercury_b200.h); 72 bytes.
constexpr int32_t MB_MFSK_FIXED_DELAY_FLAG = 0x40000000;  // == MERCURY_B200_MFSK_FIXED_DELAY_FLAG (static_assert in mb_api.cu)
struct MbReceiveStats {
	int32_t iterations_done, delay, delay_of_last_decoded_message, sync_trials;
	int32_t message_decoded, crc, all_zeros, mfsk_search_or_overflow;
	double freq_offset, freq_offset_of_last_decoded_message, SNR, signal_stregth_dbm, coarse_metric;
};

void mb_fe_host_const(MbFeConst *k);
cudaError_t mb_fe_begin(const MbFeArgs &a, const MbReceiveStats *d_stats_in, cudaStream_t s);  // states <- link state, first Schmidl-Cox request
cudaError_t mb_fe_finish(const MbFeArgs &a, MbReceiveStats *d_stats_out, cudaStream_t s);
int mb_fe_buffer_nsymb(int Nsymb, int pre);
void mb_fe_host_carrier(const MbFeConst &k, double *cs, int n);
cudaError_t mb_fe_init(const MbFeConst &k);
cudaError_t mb_fe_p2b_full(const MbFeArgs &a, cudaStream_t s);
cudaError_t mb_fe_step(const MbFeArgs &a, bool run_sc, cudaStream_t s);  // [k_fe_window, k_fe_sc,] k_fe_decide
cudaError_t mb_fe_extract(const MbFeArgs &a, cudaStream_t s);       // k_fe_moose + k_fe_extract_tiles
cudaError_t mb_fe_extract_data(const MbFeArgs &a, cudaStream_t s);  // k_fe_extract_tiles only
cudaError_t mb_fe_p2b_data(const MbFeArgs &a, cudaStream_t s);      // whole-buffer mix + FIR_rx_data (uses x, x_format, n, buf, carrier, bbi, energy_part)

// ---------------------------------------------------------------------------------------------------------------------
// TX chain (mb_tx.cu): payload bytes -> pass-band frames, transmit_byte(SINGLE_MESSAGE).
// ---------------------------------------------------------------------------------------------------------------------
struct MbTxMode {
	int32_t S, pre, nData, nPilots, nBits, nReal, nVirtual, K, P, bps, M, frame_bytes;
	double fc, Ts, amp, scale_data, scale_pre, papr_pre_lin, papr_data_lin;
	unsigned long long start_after_init;  // ofdm.passband_start_sample right after init (one symbol: telecom_system.cc:3125-3126)
	// byte offsets into the mode's TX table buffer
	uint32_t off_bit_src;   // u16[nBits]   codeword position of mapped bit i (bit interleave o parity compaction, inverted)
	uint32_t off_sym_cell;  // u16[nData]   grid cell of data symbol q (T/F interleave o framer)
	uint32_t off_scr;       // u8 [1600]    bit_energy_dispersal sequence
	uint32_t off_row_off;   // u16[P + 1]   CSR of the data variables of each check row (reference check order)
	uint32_t off_row_var;   // u16[...]
	uint32_t off_pilot;     // f64[S * 50]  pilot value at pilot cells, 0 at data cells
	uint32_t off_cons;      // c128[M]
	uint32_t off_preamble;  // c128[pre * 50]
	uint32_t off_pre_eq;    // c128[50]     pre_equalization_channel
	uint32_t off_c1, off_c2;  // f64[97]    FIR_tx1, FIR_tx2
	uint32_t pad;
};

struct MbTxArgs {
	const MbTxMode *tm;        // device copy
	const MbTxMode *tm_host;
	const uint8_t *tables;     // device: the mode's TX table buffer
	const uint8_t *payload;    // [n][frame_bytes] zero-padded payloads
	const unsigned long long *start_sample;  // [n] running carrier sample counter per frame, or NULL (= start_after_init)
	int32_t n, out_f32, no_filter;  // no_filter: NO_FILTER_MESSAGE, stop after the PAPR clip
	int32_t S_active;          // data symbols actually modulated (MFSK control frames: < S; what follows is silence)
	double2 *bb;               // [n][(pre + S) * 272] scaled base-band symbols
	double *pb, *p1;           // [n][total] pass-band before / after FIR_tx1
	double *power_part;        // [n][ceil(total / 256)][2]
	void *out;                 // [n][total] double or float
	uint8_t *dbg_cw;           // optional [n][1600] codewords (parity tests)
	const MbMfsk *tone;        // ROBUST (MFSK) modes: the tone plan (host pointer, passed to the kernel by value)
};

std::string mb_tx_build(const std::vector<uint8_t> &blob, int config, const MbFeConst &fe, MbTxMode *tm, std::vector<uint8_t> *bytes);
std::string mb_tx_build_mfsk(const std::vector<uint8_t> &blob, const MbMode &m, const MbMfsk &t, const MbFeConst &fe, MbTxMode *tm, std::vector<uint8_t> *bytes);
cudaError_t mb_tx_init(const MbFeConst &fe);
cudaError_t mb_tx_launch(const MbTxArgs &a, cudaStream_t s);
// generate_ack / generate_break_pattern_passband (telecom_system.cc:1589-1631,1657-1689): 16 hopped tones of the dedicated 16-MFSK plan
cudaError_t mb_tx_pattern(const MbMfsk &plan, int use_break_tones, double fc, double Ts, double amp, unsigned long long start_sample, double2 *d_bb, double *d_pb,
			  double *d_power_part, double *d_out, cudaStream_t s);
cudaError_t mb_tx_fir_apply(const uint8_t *tables, const MbTxMode &tm_host, const double *d_in, int n, double *d_tmp, double *d_out, cudaStream_t s);

// ---------------------------------------------------------------------------------------------------------------------
// MFSK row (mb_mfsk.cu): ROBUST_0..2 demodulator + tone-pattern detectors.
// ---------------------------------------------------------------------------------------------------------------------
struct MbMfskArgs {
	const float2 *x;       // [B][Nsymb][sym_stride] complex64 base-band, preamble stripped; useful samples start at sym_skip
	int32_t sym_stride, sym_skip;
	float *llr;            // [B][MB_HANDOFF_STRIDE] hand-off records for the LDPC kernel
	float *llr_cw;         // optional [B][1600] LLRs in codeword order
	MbRxStats *stats;      // [B]
	const uint8_t *blob;   // device blob + MFSK extension
	MbMode mode;
	MbMfsk tone;
	int32_t active_nsymb;  // get_active_nsymb(): < Nsymb for control frames (set_mfsk_ctrl_mode), the rest of the codeword is erased
};

// Mirrors mercury_b200_mfsk_pattern_result (include/mercury_b200.h); 32 bytes.
struct MbMfskPatternResult {
	int32_t time_sync_delay, ack_matched, break_matched, reserved;
	double ack_metric, break_metric;
};

cudaError_t mb_launch_mfsk_demod(const MbMfskArgs &a, size_t n_frames, cudaStream_t s);
// search start: one value for all buffers, or (d_search_start_each != NULL) one int32 per buffer every each_stride int32s
cudaError_t mb_launch_mfsk_patterns(const void *d_bbi, int is_f32, size_t n_buffers, int n_samples, int search_start_symb, const int32_t *d_search_start_each,
				    int each_stride, const MbMfsk &t, int pre, double *d_energies, MbMfskPatternResult *d_out, cudaStream_t s);
cudaError_t mb_launch_mfsk_rx_decide(const MbMfskPatternResult *pat, const double *energy_part, int nblk, int buf, int pre, int S, int S_active, int buffer_Nsymb,
				     double fc, MbFeState *st, MbReceiveStats *stats, int n, int *counters, cudaStream_t s);
cudaError_t mb_launch_mfsk_rx_finish(const MbFeState *st, const MbRxStats *tail_stats, const uint8_t *tail_payload, int frame_bytes, uint8_t *payload_out,
				     MbReceiveStats *stats, int n, cudaStream_t s);
