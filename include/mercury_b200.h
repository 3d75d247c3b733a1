/*
 * mercury_b200.h -- C ABI of the B200-native Mercury physical-layer RX hot path.
 *
 * Drop-in boundary.  The reference has no FFI layer: the boundary is the C++ member API of
 * one long-lived object, cl_telecom_system (reference include/physical_layer/telecom_system.h:85-199).  The
 * entry points below are what a binding for THAT path needs; each cites the reference interface it replaces
 * (paths relative to the reference repository, Rhizomatica/mercury @ c91aa4b):
 *
 *   mercury_b200_create / destroy        cl_telecom_system::cl_telecom_system / ~  (telecom_system.cc:38-94)
 *   mercury_b200_load_tables             the table-building half of cl_telecom_system::init()
 *                                        (telecom_system.cc:1804-1982: ofdm.init, ldpc.init, scrambler)
 *   mercury_b200_export/import_tables    (new) the blob rank 0 broadcasts over NCCL (north_star)
 *   mercury_b200_load_configuration      void load_configuration(int)          (telecom_system.h:176, .cc:2487)
 *                                        + "-I n" = ldpc_nIteration_max        (main.cc:303-311,547-575)
 *   mercury_b200_get_frame_size_bytes    int get_frame_size_bytes()            (telecom_system.h:180, .cc:332)
 *   mercury_b200_get_frame_size_bits     int get_frame_size_bits()             (telecom_system.h:181, .cc:337)
 *   mercury_b200_get_geometry            public members data_container.{Nsymb,Nofdm,nBits,...}, ldpc.{N,K,P}
 *   mercury_b200_receive_baseband        the tail of st_receive_stats receive_byte(double*,int*)
 *                                        (telecom_system.h:142, .cc:1132-1429) on its post-synchronisation
 *                                        data_container.baseband_data
 *   mercury_b200_demod_decode_batch      the same tail, batched over independent frames (host buffers)
 *   mercury_b200_demod_decode_batch_device   "  (device-resident buffers, caller's stream)
 *   mercury_b200_demod_batch_device      symbol_demod .. psk.demod .. LLR expand   (telecom_system.cc:1135-1308)
 *   mercury_b200_ldpc_decode_batch_device    ldpc.decode .. CRC16                  (telecom_system.cc:1310-1349)
 *   mercury_b200_receive_byte(_batch)    st_receive_stats receive_byte(double*,int*) WHOLE (telecom_system.h:142, .cc:646-1518):
 *                                        pass-band capture in, front-end on the GPU
 *   mercury_b200_transmit_byte(_batch)   void transmit_byte(int*, int, double*, SINGLE_MESSAGE) (telecom_system.h:138, .cc:342-553): the TX chain
 *                                        to pass-band on the GPU
 *   mercury_b200_rx_stats                st_receive_stats (telecom_system.h:63-82), the fields the tail writes
 *   mercury_b200_batcher_*               (new) many concurrent links' receive calls -> one GPU batch; replaces the one-frame-per-call
 *                                        pattern of arq_common.cc:2619-2668 / audioio.c:999-1069 for a multi-link gateway
 *   mercury_b200_synth_frames            (test/bench input synthesis) transmit_byte bit chain + the baseband
 *                                        modulation chain of baseband_test_EsN0 (telecom_system.cc:342-416,129-153)
 *
 * Conventions: every function returns 0 on success or a negative MERCURY_B200_E* code, never exits, never
 * throws.  No CUDA or torch types appear in the signatures: device buffers and streams are passed as void*
 * (a cudaStream_t is a pointer; NULL = the legacy default stream).  There is NO CPU fallback: compute entry
 * points fail with MERCURY_B200_ENODEV when no sm_100 device is usable.
 */
#ifndef MERCURY_B200_H
#define MERCURY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MERCURY_B200_OK 0
#define MERCURY_B200_EINVAL (-1)   /* bad argument / unknown configuration */
#define MERCURY_B200_ENODEV (-2)   /* no usable CUDA device (this library never falls back to the CPU) */
#define MERCURY_B200_ECUDA (-3)    /* a CUDA runtime call or kernel failed; see mercury_b200_last_error() */
#define MERCURY_B200_ESTATE (-4)   /* tables not loaded / configuration not selected */
#define MERCURY_B200_EIO (-5)      /* LDPC table file missing or malformed */
#define MERCURY_B200_ENOMEM (-6)

#define MERCURY_B200_DECODER_SPA 0     /* sum-product, the reference's default (physical_config.cc:72) */
#define MERCURY_B200_DECODER_MINSUM 1  /* normalised min-sum (north_star fast path) */

#define MERCURY_B200_NUM_CONFIGS 17    /* CONFIG_0 .. CONFIG_16 */
#define MERCURY_B200_N 1600            /* LDPC block length, one codeword per OFDM frame */
#define MERCURY_B200_NOFDM 272         /* samples per OFDM symbol at the decimated rate (Nfft 256 + GI 16) */
#define MERCURY_B200_HANDOFF_FLOATS 2400 /* float32 per frame of the stage hand-off buffer between the two kernels:
                                            1600 LLRs + (ZF modes) up to 400 equalised data symbols for the SNR report */

typedef struct mercury_b200 mercury_b200_t;

/* Frame geometry of the selected configuration (reference: members of data_container / ofdm / ldpc). */
typedef struct mercury_b200_geometry {
	int32_t config, M, bits_per_symbol, ldpc_rate_num /* rate = n/16 */;
	int32_t Nsymb, Nc, Nfft, Ngi, Nofdm;
	int32_t nData, nPilots, nBits;
	int32_t N, K, P, nReal /* nBits-P */, nVirtual /* N-nBits */;
	int32_t preamble_nSymb, frame_bytes;
	int32_t estimator /* 0 ZERO_FORCE, 1 LEAST_SQUARE */, phase_only /* amplitude restoration */;
	int32_t ldpc_iters, ldpc_edges, decoder;
} mercury_b200_geometry;

/* Per-frame result record: the st_receive_stats fields written by telecom_system.cc:1310-1375. */
typedef struct mercury_b200_rx_stats {
	int32_t iterations_done; /* 0..I, I+1 = not converged (ldpc_decoder_SPA.cc:217), -1 = decode skipped (mean|H| gate, :1271) */
	int32_t crc;             /* CRC16 over payload+crc bytes, 0 = good (telecom_system.cc:1337-1341) */
	int32_t all_zeros;       /* telecom_system.cc:1319-1327 */
	int32_t message_decoded; /* 1 = YES */
	float SNR;               /* dB, -99.9 when not decoded (:1347,1368-1375); ZF modes: re-encode path (:1376-1400) */
	float variance;          /* pilot noise variance used for the LLRs (:1291) */
	float mean_H;            /* mean |H| over pilots after AGC (:1225-1244) */
	int32_t reserved;
} mercury_b200_rx_stats;

const char *mercury_b200_version(void);
const char *mercury_b200_strerror(int code);

int mercury_b200_create(int device, mercury_b200_t **out);
void mercury_b200_destroy(mercury_b200_t *h);
const char *mercury_b200_last_error(const mercury_b200_t *h);

/* Build all 17 mode tables + 8 LDPC graphs from the LDPC table file and upload them (one copy to HBM). */
int mercury_b200_load_tables(mercury_b200_t *h, const char *ldpc_table_path);
/* Table blob exchange for multi-GPU start-up: rank 0 exports, broadcasts (NCCL), the others import. */
int mercury_b200_export_tables(const mercury_b200_t *h, void *buf, size_t *size /* in: capacity, out: bytes */);
int mercury_b200_import_tables(mercury_b200_t *h, const void *buf, size_t size);
/*
 * The same exchange as ONE call from C / C++ host code (north_star: "partitioned across GPUs only as independent frame shards via NCCL
 * broadcast of the codeword tables"): every rank of `nccl_comm` (an ncclComm_t the host program created, one rank per
 * handle / GPU -- several handles of one process with ncclCommInitAll, or one per process) calls this once; the root's handle must have
 * its tables loaded (mercury_b200_load_tables), the others receive the blob over NCCL (NVLink / NVSwitch between the GPUs of a node),
 * validate and import it.  `stream` is a cudaStream_t of the handle's device (NULL: the default stream).  After it, ranks process disjoint
 * frame ranges with no further collective.  libnccl.so.2 is loaded on first use (dlopen): the library has no link-time NCCL dependency.
 */
int mercury_b200_broadcast_tables(mercury_b200_t *h, void *nccl_comm, int root, void *stream);
/* Host-only table construction (no device needed): same blob as export_tables(). buf may be NULL to query size. */
int mercury_b200_build_tables_host(const char *ldpc_table_path, void *buf, size_t *size);

/* O(1): all configurations stay resident. config 0..16, ldpc_iters clamped to 5..50 like the CLI's -I. */
int mercury_b200_load_configuration(mercury_b200_t *h, int config, int ldpc_iters);
int mercury_b200_set_decoder(mercury_b200_t *h, int decoder);
int mercury_b200_get_geometry(const mercury_b200_t *h, mercury_b200_geometry *out);
int mercury_b200_get_frame_size_bytes(const mercury_b200_t *h);
int mercury_b200_get_frame_size_bits(const mercury_b200_t *h);

/*
 * Batched RX tail on frames that are already time/frequency synchronised (what receive_byte() holds in
 * data_container.baseband_data after telecom_system.cc:1131), preamble stripped:
 *   baseband : n_frames x Nsymb x 272 complex samples, interleaved (re,im) float32
 *   payload  : n_frames x frame_bytes bytes (CRC bytes stripped, telecom_system.cc:1329-1332)
 *   stats    : n_frames records
 *   llr_cw   : optional n_frames x 1600 float32, LLRs in codeword order (the input of ldpc.decode), or NULL
 * Host variant: pointers are host memory (pinned memory from mercury_b200_host_alloc gives full PCIe rate);
 * copies are pipelined with the kernels in chunks.  Device variant: pointers are device memory on h's device;
 * work is enqueued on `stream` and NOT synchronised.
 */
int mercury_b200_demod_decode_batch(mercury_b200_t *h, const float *baseband, size_t n_frames, uint8_t *payload,
				    mercury_b200_rx_stats *stats, float *llr_cw);
int mercury_b200_demod_decode_batch_device(mercury_b200_t *h, const void *d_baseband, size_t n_frames, void *d_payload,
					   void *d_stats, void *d_llr_cw, void *stream);
/*
 * The same two calls for narrower base-band samples.  A host batch is bound by the PCIe link (H2D of the samples), so the
 * bytes per sample are the throughput of the call: 4-byte complex samples double it against complex64.
 *   MERCURY_B200_BASEBAND_C64   interleaved (re, im) float32 -- what the calls above take
 *   MERCURY_B200_BASEBAND_CI16  interleaved (re, im) int16; sample = integer x scale (what a fixed-point front-end / ADC delivers)
 *   MERCURY_B200_BASEBAND_CF16  interleaved (re, im) IEEE binary16 (scale ignored)
 * The kernel widens the samples to float32 in its load; everything after that is the arithmetic of the complex64 call, i.e. the
 * result is that of the reference on the same quantised values (parity tests feed the reference exactly those, as doubles).
 * OFDM configurations only (the ROBUST / MFSK tail takes complex64).
 */
#define MERCURY_B200_BASEBAND_C64 0
#define MERCURY_B200_BASEBAND_CI16 1
#define MERCURY_B200_BASEBAND_CF16 2
int mercury_b200_demod_decode_batch_fmt(mercury_b200_t *h, const void *baseband, int sample_format, float scale, size_t n_frames,
					uint8_t *payload, mercury_b200_rx_stats *stats, float *llr_cw);
int mercury_b200_demod_decode_batch_device_fmt(mercury_b200_t *h, const void *d_baseband, int sample_format, float scale, size_t n_frames,
					       void *d_payload, void *d_stats, void *d_llr_cw, void *stream);
/* The two stages separately (device buffers). d_llr is the stage hand-off: n_frames x MERCURY_B200_HANDOFF_FLOATS float32
 * (LLRs in the decoder's internal layout, then -- ZF modes only -- the equalised data symbols the decoder's SNR report needs). */
int mercury_b200_demod_batch_device(mercury_b200_t *h, const void *d_baseband, size_t n_frames, void *d_llr, void *d_stats,
				    void *d_llr_cw, void *stream);
int mercury_b200_ldpc_decode_batch_device(mercury_b200_t *h, const void *d_llr, size_t n_frames, void *d_payload, void *d_stats,
					  void *stream);
/* Optional per-stage capture for parity tests (device, n_frames x Nsymb x 50 complex64 each; any may be NULL). */
int mercury_b200_set_debug_capture(mercury_b200_t *h, void *d_Y, void *d_H, void *d_Z);
/* Bytes of scratch (internal-order LLRs) demod_decode_batch_device needs per frame; owned by the handle and grown on demand. */

/*
 * One frame, the reference's own types: baseband = Nsymb*272 std::complex<double> (re,im doubles), out = one int
 * per payload byte like receive_byte()'s `int* out`.  Returns the stats record by pointer.
 */
int mercury_b200_receive_baseband(mercury_b200_t *h, const double *baseband, int *out, mercury_b200_rx_stats *stats);

/*
 * The WHOLE receive_byte(): pass-band capture in, payload out -- the reference's front-end
 * (st_receive_stats cl_telecom_system::receive_byte(double* data, int* out), telecom_system.h:142, .cc:646-1518, OFDM branch:
 * passband_to_baseband ofdm.cc:2316-2339, time_sync_preamble(_with_metric) ofdm.cc:1735-1967, the gates / recoveries / trial
 * loop of telecom_system.cc:700-1131, carrier_sampling_frequency_sync ofdm.cc:540-595, SKIP-H recovery :1436-1504) runs on the
 * GPU in front of the same tail.  A capture is mercury_b200_get_capture_samples() = Nofdm * buffer_Nsymb * 4 real samples at
 * 48 kHz (data_container.cc:133-153), the frame anywhere inside it.
 *
 * mercury_b200_receive_stats = st_receive_stats (telecom_system.h:63-82), the fields the OFDM branch writes.  The two
 * *_of_last_decoded_message fields are the reference's cross-call link state (telecom_system.cc:945-947,1108-1110,1423-1427):
 * they are INPUTS as well as outputs (initialise to -1 and 0.0 for a new link, then pass the record back in on the next call).
 * Sync decisions (delay, sync_trials) are bit-identical to the reference; see mb_frontend.cu for how.
 */
typedef struct mercury_b200_receive_stats {
	int32_t iterations_done, delay, delay_of_last_decoded_message, sync_trials;
	int32_t message_decoded, crc, all_zeros;
	int32_t mfsk_search_or_overflow; /* ROBUST (MFSK) configurations only.  IN: first symbol of the tone-preamble search, the reference's
	                                    receive_stats.mfsk_search_raw - nUnder_processing_events (telecom_system.cc:684).  OUT:
	                                    frame_overflow_symbols (:702-715), > 0 when the frame runs past the end of the capture.
	                                    IN, alternatively, MERCURY_B200_MFSK_FIXED_DELAY(d): the reference's one-shot member mfsk_fixed_delay = d
	                                    (telecom_system.h:110, .cc:663-673; set by the ARQ layer's overflow recapture, arq_common.cc:2830-2833):
	                                    no search, delay = d samples, signal_stregth_dbm = 0.  Consumed like the reference's: OUT never carries it. */
	double freq_offset, freq_offset_of_last_decoded_message, SNR, signal_stregth_dbm, coarse_metric;
} mercury_b200_receive_stats;

#define MERCURY_B200_MFSK_FIXED_DELAY_FLAG 0x40000000
#define MERCURY_B200_MFSK_FIXED_DELAY(d) ((int32_t)(MERCURY_B200_MFSK_FIXED_DELAY_FLAG | ((d) < 0 ? 0 : (d))))

#define MERCURY_B200_SAMPLES_F64 0  /* double, the reference's own type */
#define MERCURY_B200_SAMPLES_F32 1  /* float: half the bytes; every float is exactly representable as the double the reference would see */
/* The other capture formats of the reference's audio layer, converted on the device exactly as source/audioio/audioio.c:893-940 does: */
#define MERCURY_B200_SAMPLES_I16 2  /* int16 PCM, x / 32768.0 (a quarter of the bytes of double) */
#define MERCURY_B200_SAMPLES_I32 3  /* int32 PCM, x / (double)INT_MAX (the reference's default capture format, audioio.c:744) */

int mercury_b200_get_capture_samples(const mercury_b200_t *h);
/* g_gui_state.coarse_freq_sync_enabled (include/common/gui_state.h:143; read by receive_byte() at telecom_system.cc:949): the optional coarse
 * frequency search of trial 1 -- when trial 0 fails, Schmidl-Cox over the head of the buffer with the time-sync filter at fc - 30, fc and
 * fc + 30 Hz (:949-983), the winning carrier kept for the rest of the call if it beats 0 Hz by 0.1 (:985-993). Off by default, like the
 * reference's flag. Affects mercury_b200_receive_byte* in the OFDM configurations. */
int mercury_b200_set_coarse_freq_sync(mercury_b200_t *h, int enable);

/* void set_mfsk_ctrl_mode(bool) / int get_active_nsymb() (telecom_system.cc:1572-1580): shortened control frames in ROBUST_0 (240 of 320 symbols)
 * and ROBUST_1 (175 of 200): transmit_byte modulates only the active symbols (silence follows), receive_byte / the tail demodulate only those and
 * erase the rest of the codeword.  Both return the active symbol count; load_configuration switches the mode off like the reference. */
int mercury_b200_set_mfsk_ctrl_mode(mercury_b200_t *h, int enable);
int mercury_b200_get_active_nsymb(const mercury_b200_t *h);
/* char get_configuration(double SNR) (telecom_system.h:178, .cc:3036-3106): the gear-shift ladder; pure host function. */
int mercury_b200_get_configuration(double SNR);
/* double measure_signal_only(double* data) (telecom_system.h, .cc:1520-1541): signal strength in dBm of n captures, no search, no decode. */
int mercury_b200_measure_signal_only_batch(mercury_b200_t *h, const void *passband, int sample_format, size_t n_captures, double *signal_dbm);
/* Host-only (no device needed): the two receive FIR designs (33 taps each), {fs, fc, carrier amplitude, bandwidth, time_sync_trials_max,
 * use_last_good_time_sync, use_last_good_freq_offset, freq_offset_ignore_limit} and the first n_carrier (cos, sin) pairs of the carrier table. */
int mercury_b200_build_frontend_tables_host(double *ts_coef, double *data_coef, double *consts, double *carrier, int n_carrier);
/* One capture, the reference's own types (out = one int per payload byte). */
int mercury_b200_receive_byte(mercury_b200_t *h, const double *passband, int *out, mercury_b200_receive_stats *stats);
/* n captures of independent links: passband n x capture_samples (host), payload n x frame_bytes, stats n records (in/out).
 * baseband_dbg (optional, host): n x (preamble_nSymb + Nsymb) x 272 complex128 = data_container.baseband_data of the last trial. */
int mercury_b200_receive_byte_batch(mercury_b200_t *h, const void *passband, int sample_format, size_t n_captures, uint8_t *payload,
				    mercury_b200_receive_stats *stats, double *baseband_dbg);
/* Same on device-resident buffers.  The per-capture control flow is a device-side state machine, but the loop around it reads
 * three counters back per round, so this call synchronises `stream` before it returns. */
int mercury_b200_receive_byte_batch_device(mercury_b200_t *h, const void *d_passband, int sample_format, size_t n_captures, void *d_payload,
					   void *d_stats, void *stream);

/*
 * TX chain on the GPU: void cl_telecom_system::transmit_byte(int* data, int nBytes, double* out, int message_location)
 * with message_location == SINGLE_MESSAGE (telecom_system.h:138, .cc:342-553, OFDM branch): zero pad + CRC16, scrambler, LDPC encode,
 * interleavers, mapper, framer, preamble, pre-equalisation (telecom_system.cc:3108-3146), IFFT + guard interval, x4 interpolation and mixing
 * (ofdm.cc:2279-2315), PAPR clip (ofdm.cc:1565-1592), FIR_tx1, FIR_tx2 -> mercury_b200_get_total_frame_size() pass-band samples per frame.
 * The reference's running carrier sample counter (ofdm.passband_start_sample, advanced by every baseband_to_passband call) is explicit:
 * start_sample[i] for frame i (NULL = the value a freshly initialised reference object has, one symbol = 1088 samples).
 */
int mercury_b200_get_total_frame_size(const mercury_b200_t *h);
/* Host-only table construction for one configuration (no device needed): pre_eq 50 complex, preamble preamble_nSymb x 50 complex, 97 + 97 taps. */
int mercury_b200_build_tx_tables_host(const char *ldpc_table_path, int config, double *pre_eq, double *preamble, double *tx1, double *tx2);
/* One frame, the reference's own types; *passband_start_sample is advanced like the reference's counter (may be NULL). */
int mercury_b200_transmit_byte(mercury_b200_t *h, const int *data, int nBytes, double *out, uint64_t *passband_start_sample);
/* n frames: payload n x frame_bytes (zero-padded), passband n x total_frame_size of out_format (MERCURY_B200_SAMPLES_F64 or _F32);
 * codeword_dbg (optional, host variant): n x 1600 LDPC codeword bits for parity tests. */
int mercury_b200_transmit_byte_batch(mercury_b200_t *h, const uint8_t *payload, const uint64_t *start_sample, size_t n_frames, void *passband,
				     int out_format, uint8_t *codeword_dbg);
/* message_location (common_defines.h:197-201): SINGLE_MESSAGE = the filtered frame; NO_FILTER_MESSAGE = the clipped frame before the
 * transmit FIRs, which is what the ARQ layer requests per frame (arq_common.cc:2224) before it filters the whole padded batch with
 * ofdm.FIR_tx1.apply / ofdm.FIR_tx2.apply (arq_common.cc:2243-2246) = mercury_b200_fir_tx_apply.  The streaming FIRST / MIDDLE / FLUSH
 * locations (TX_TEST's three-frame filter buffer) are single-stream: mercury_b200_transmit_byte_loc below. */
#define MERCURY_B200_SINGLE_MESSAGE 3
#define MERCURY_B200_NO_FILTER_MESSAGE 4
int mercury_b200_transmit_byte_batch_ex(mercury_b200_t *h, const uint8_t *payload, const uint64_t *start_sample, size_t n_frames, void *passband,
					int out_format, int message_location, uint8_t *codeword_dbg);
int mercury_b200_fir_tx_apply(mercury_b200_t *h, const double *in, size_t n_samples, double *out);
/* One frame with any message_location, the reference's own types: FIRST_MESSAGE 0 / MIDDLE_MESSAGE 1 / FLUSH_MESSAGE 2 stream through a
 * three-frame filter buffer kept on the device (telecom_system.cc:559-594; TX_TEST, :2033-2038) and return the PREVIOUS frame filtered with both
 * neighbours as context; 3 and 4 as above.  One stream per handle. */
#define MERCURY_B200_FIRST_MESSAGE 0
#define MERCURY_B200_MIDDLE_MESSAGE 1
#define MERCURY_B200_FLUSH_MESSAGE 2
int mercury_b200_transmit_byte_loc(mercury_b200_t *h, const int *data, int nBytes, double *out, uint64_t *passband_start_sample, int message_location);
int mercury_b200_reset_tx_stream(mercury_b200_t *h);
int mercury_b200_transmit_byte_batch_device(mercury_b200_t *h, const void *d_payload, const void *d_start_sample, size_t n_frames, void *d_passband,
					    int out_format, void *stream);

/*
 * MFSK row.  mercury_b200_load_configuration also accepts the ROBUST configurations 100..102 (ROBUST_0..2,
 * common_defines.h:63-65: 32-MFSK x1 / 16-MFSK x2, LDPC 1/16, 1/16, 4/16): the batch and single-frame TAIL entry points above then run the
 * MFSK branch of the tail (telecom_system.cc:1132-1198: symbol_demod + cl_mfsk::demod, mfsk.cc:305-390) in front of the same decoder, on
 * Nsymb = 320 / 200 / 200 synchronised symbols per frame; SNR reads 0 for a decoded frame like the reference's.  The pass-band entry
 * points (receive_byte / transmit_byte and their batch forms) run the MFSK branches of the reference in these configurations as well.
 * The tone-pattern detectors work on base-band buffers at the pass-band rate (baseband_data_interpolated), n_samples complex samples each:
 *   time_sync_delay            int cl_ofdm::time_sync_mfsk(...)            ofdm.cc:1969-2065 (preamble tones of the loaded configuration)
 *   ack_metric / ack_matched   double cl_ofdm::detect_ack_pattern(...)     ofdm.cc:2067-2186 with mfsk.ack_tones   (mfsk.cc:113-136)
 *   break_metric / _matched    the same with mfsk.break_tones              (mfsk.cc:138-160)
 */
typedef struct mercury_b200_mfsk_pattern_result {
	int32_t time_sync_delay, ack_matched, break_matched, reserved;
	double ack_metric, break_metric;
} mercury_b200_mfsk_pattern_result;
int mercury_b200_mfsk_patterns_batch(mercury_b200_t *h, const void *bbi /* n_buffers x n_samples x (re, im) */, int complex_format /* _F64 | _F32 */,
				     size_t n_buffers, int n_samples, int search_start_symb, mercury_b200_mfsk_pattern_result *out);

/*
 * The ARQ-facing tone-pattern calls (telecom_system.h:122-130), valid in EVERY configuration: they use the reference's dedicated
 * ack_mfsk plan (16-MFSK, one stream, telecom_system.cc:3003-3008).
 *   mercury_b200_generate_pattern_passband           int generate_ack_pattern_passband(double* out) / generate_break_pattern_passband
 *                                                    (telecom_system.cc:1589-1631,1657-1689): 16 x 1088 clipped pass-band samples, no FIRs;
 *                                                    returns the sample count; *passband_start_sample = ofdm.passband_start_sample, advanced
 *   mercury_b200_detect_patterns_from_passband_batch double detect_ack_pattern_from_passband(double* data, int size, int* matched) and
 *                                                    detect_break_pattern_from_passband (telecom_system.cc:1633-1655,1691-1710): mix + FIR_rx_data
 *                                                    over the whole buffer, then the matched tone detector, ACK and BREAK in one pass
 *                                                    (time_sync_delay in the result is the 16-MFSK preamble search and has no reference caller)
 */
int mercury_b200_generate_pattern_passband(mercury_b200_t *h, int use_break_tones, double *out, uint64_t *passband_start_sample);
int mercury_b200_detect_patterns_from_passband_batch(mercury_b200_t *h, const void *passband, int sample_format, size_t n_buffers, int n_samples,
						     mercury_b200_mfsk_pattern_result *out);

/* Pinned host memory and plain device memory helpers for callers that do not link the CUDA runtime. */
void *mercury_b200_host_alloc(size_t bytes);
void mercury_b200_host_free(void *p);
void *mercury_b200_device_alloc(mercury_b200_t *h, size_t bytes);
void mercury_b200_device_free(mercury_b200_t *h, void *p);
int mercury_b200_memcpy_h2d(mercury_b200_t *h, void *dst, const void *src, size_t bytes);
int mercury_b200_memcpy_d2h(mercury_b200_t *h, void *dst, const void *src, size_t bytes);
int mercury_b200_synchronize(mercury_b200_t *h);

/*
 * Multi-link batcher.  In the reference every link decodes one frame per receive_byte() call on its own
 * thread (arq_common.cc:2619-2668; the capture thread feeds it under capture_prep_mutex, audioio.c:999-1069).  A gateway that
 * terminates many links has thousands of such calls in flight: the batcher keeps each call synchronous and per frame (same
 * contract as mercury_b200_receive_baseband, float samples) and turns the concurrency into GPU batch size.  A batch is closed
 * when max_batch frames are waiting or when its oldest frame has waited max_wait_us.  Thread safe; bound to the configuration
 * selected on `h` when it is created (do not call load_configuration on `h` while a batcher is alive).
 */
typedef struct mercury_b200_batcher mercury_b200_batcher_t;
int mercury_b200_batcher_create(mercury_b200_t *h, size_t max_batch, unsigned max_wait_us, mercury_b200_batcher_t **out);
int mercury_b200_batcher_receive_baseband(mercury_b200_batcher_t *b, const float *baseband /* Nsymb x 272 x (re,im) */, uint8_t *payload,
					  mercury_b200_rx_stats *stats);
/* Pass-band flavour: every link's synchronous call hands over its whole capture buffer (capture_samples samples of `sample_format`)
 * and its link-state record, like cl_telecom_system::receive_byte(double* data, int* out) (arq_common.cc:2668); one batch = one
 * mercury_b200_receive_byte_batch over all waiting links. */
int mercury_b200_batcher_create_passband(mercury_b200_t *h, int sample_format, size_t max_batch, unsigned max_wait_us, mercury_b200_batcher_t **out);
int mercury_b200_batcher_receive_byte(mercury_b200_batcher_t *b, const void *passband, uint8_t *payload, mercury_b200_receive_stats *stats /* in/out */);
int mercury_b200_batcher_get_counters(mercury_b200_batcher_t *b, uint64_t *batches, uint64_t *frames, uint64_t *full_batches);
void mercury_b200_batcher_destroy(mercury_b200_batcher_t *b);
/* Test hook: the same batching machinery in front of a caller-supplied batch function (plain host memory), so that the
 * concurrency logic can be exercised on a machine without a GPU.  Not a CPU decode path: the function is the test's. */
int mercury_b200_batcher_create_with_backend(size_t frame_floats, size_t frame_bytes, size_t max_batch, unsigned max_wait_us,
					     int (*run)(void *ctx, const float *x, size_t n, uint8_t *payload, mercury_b200_rx_stats *stats),
					     void *ctx, mercury_b200_batcher_t **out);

/* Number of kernels this library has launched through this handle (bench.py's gpu_launches). */
uint64_t mercury_b200_kernel_launches(const mercury_b200_t *h);

/*
 * Input synthesis on the host (tests / bench only; not part of the RX path): n_frames frames of configuration
 * `config`, payload bytes from splitmix64(seed, frame index) (or `payload_in` if non-NULL), through the TX bit
 * chain and baseband OFDM modulator, plus complex AWGN at Es/N0 = esn0_db (baseband_test_EsN0 normalisation,
 * telecom_system.cc:141-153; pass esn0_db >= 200 for no noise).  Needs the LDPC table file, not a device.
 *   baseband_out : n_frames x Nsymb x 272 x 2 float32      payload_out : n_frames x frame_bytes (may be NULL)
 */
int mercury_b200_synth_frames(const char *ldpc_table_path, int config, size_t n_frames, uint64_t seed, double esn0_db,
			      const uint8_t *payload_in, float *baseband_out, uint8_t *payload_out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* MERCURY_B200_H */
