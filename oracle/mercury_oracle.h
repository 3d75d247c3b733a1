/*
 * oracle/mercury_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product).
 *
 * Plain-C, double-precision CPU restatement of the Mercury physical-layer RX hot path
 * (reference: source/physical_layer/telecom_system.cc:1132-1341 and its callees) plus the TX bit /
 * modulation chain needed to synthesise inputs.  Every function in mercury_oracle.c cites the
 * reference file:line it follows.  Pinned against the unmodified reference (oracle/_ref) and the
 * golden fixtures in tests/golden/ by tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may use it.
 */
#ifndef MERCURY_ORACLE_H
#define MERCURY_ORACLE_H

#include <complex.h>
#include <stdint.h>

#ifdef __cplusplus
#error "plain C only"
#endif

#define MO_N 1600
#define MO_NC 50
#define MO_NFFT 256
#define MO_NGI 16
#define MO_NOFDM 272
#define MO_MAX_CELLS (48 * MO_NC)

/* RX front-end constants (physical_config.cc:59,79-101) and the two receive FIR designs (fir_filter.cc:45-163). */
typedef struct mo_frontend {
	int interp, buffer_Nsymb, trials_max, use_last_time, use_last_freq, ntaps_ts, ntaps_data, coarse_freq_sync;
	double fs, fc, amp, bandwidth, ignore_limit;
	double c_ts[64], c_data[64];
} mo_frontend;

/* MFSK tone plan of the ROBUST modes (cl_mfsk, include/physical_layer/mfsk.h). */
typedef struct mo_mfsk {
	int M, nBits, Nc, nStreams, tone_hop_step, preamble_nSymb;
	int stream_offsets[4], preamble_tones[8], ack_tones[8], break_tones[8];
} mo_mfsk;

/* TX-side tables, built lazily by mo_tx_init(). */
typedef struct mo_tx {
	int ready, ntaps1, ntaps2;
	int preamble_type[4 * MO_NC];		 /* 1 = PREAMBLE carrier, 0 = ZERO */
	double complex preamble[4 * MO_NC];
	double complex pre_eq[MO_NC];
	double c1[128], c2[128];
	double output_power, preamble_boost, preamble_papr, data_papr;
	unsigned long start_sample_after_init;
} mo_tx;

typedef struct mo_mode {
	int config, M, bits_per_symbol, rate_num;
	int Nsymb, Nc, Nfft, Ngi, Nofdm, nData, nPilots, nBits;
	int N, K, P, nReal, nVirtual, preamble_nSymb, frame_bytes;
	int estimator;	 /* 0 = ZERO_FORCE, 1 = LEAST_SQUARE */
	int phase_only;	 /* channel_estimator_amplitude_restoration == YES */
	int ldpc_iters, bit_il_block, tf_il_block, ls_window;
	double boost;
	unsigned char is_pilot[MO_MAX_CELLS];
	double pilot_val[MO_MAX_CELLS];
	double complex constellation[64];
	int scrambler[MO_N];
	int Cwidth, Vwidth, n_edges;
	int *C;	   /* [P][Cwidth], -1 padded   (QCmatrixC)   */
	int *V;	   /* [N][Vwidth], -1 padded   (QCmatrixV)   */
	int *Vpos; /* [P][Cwidth]              (V_pos)       */
	int *vdeg; /* [N]                      (from QCmatrixd) */
	double complex twiddle[MO_NFFT / 2];
	int bitrev[MO_NFFT];
	mo_frontend fe;
	mo_tx tx;
	mo_mfsk mfsk; /* ROBUST_0..2 (config 100..102, M == 200) only */
	int ctrl_nBits, ctrl_nsymb, ctrl_mode; /* MFSK control frames (telecom_system.cc:1572-1585, 2966-2995) */
	double *tx_stream;		       /* passband_data_tx_buffer: three frames, streaming message locations */
} mo_mode;

typedef struct mo_rx_out {
	double complex *Y;     /* [Nsymb*Nc] after FFT + AGC       */
	double complex *H;     /* [Nsymb*Nc] channel used by the equaliser */
	double complex *Z;     /* [Nsymb*Nc] equalised grid        */
	float *llr_demod;      /* [nBits]                          */
	float *llr_cw;	       /* [N]  codeword order              */
	int *bits;	       /* [K]  hard decisions (scrambled)  */
	int *bytes;	       /* [nReal/8] incl. CRC              */
	int *payload;	       /* [frame_bytes]                    */
	double *stats;	       /* [8] iterations, crc, all_zeros, decoded, SNR, variance, 0, mean_H */
} mo_rx_out;

int mo_mode_init(mo_mode *m, int config, int ldpc_iters, const char *ldpc_blob_path);
void mo_mode_free(mo_mode *m);
mo_mode *mo_mode_new(int config, int ldpc_iters, const char *ldpc_blob_path);
void mo_mode_delete(mo_mode *m);
void mo_geometry(const mo_mode *m, int *g /*[28], same order as ref_driver.cc*/);
void mo_tables(const mo_mode *m, int *carrier_type, double *pilot_seq, int *scrambler, double *constellation, double *boost);
void mo_ldpc_tables(const mo_mode *m, int *dims, int *C, int *V, int *d, int *Enc);

void mo_srandom(unsigned seed);
int mo_random(void);
void mo_random_seq(unsigned seed, int n, int *out);
int mo_crc16(const int *bytes, int n);

void mo_tx_baseband(const mo_mode *m, const int *payload, int nBytes, double complex *out, int *info_bits, int *codeword, double complex *framed);
void mo_rx_tail(const mo_mode *m, const double complex *baseband, mo_rx_out *o);
double mo_rx_tail_timed(const mo_mode *m, const double complex *baseband, int n_frames, int *payloads, int *decoded, int *iterations);
int mo_ldpc_decode(const mo_mode *m, const float *llr_cw, int *bits_out);
void mo_ldpc_encode(const mo_mode *m, const int *data, int *encoded);

/* RX front-end: the whole receive_byte(), pass-band capture in, payload out. */
void mo_frontend_init(mo_mode *m);
void mo_frontend_tables(const mo_mode *m, int *ntaps /*[2]*/, double *ts_coef, double *data_coef, double *consts /*[8]*/);
void mo_receive_byte(const mo_mode *m, const double *passband, int *out, double *stats /*[12]*/, double *state /*[5]*/,
		     double complex *baseband_out);
/* TX chain to pass-band: transmit_byte(SINGLE_MESSAGE). */
void mo_tx_init(mo_mode *m);
void mo_tx_tables(mo_mode *m, double complex *preamble, int *preamble_type, double complex *pre_eq, int *ntaps, double *c1, double *c2, double *consts);
int mo_transmit_byte(mo_mode *m, const int *payload, int nBytes, double *out, double *start_sample_inout);
int mo_transmit_byte_nofilter(mo_mode *m, const int *payload, int nBytes, double *out, double *start_sample_inout);
void mo_fir_tx_apply(mo_mode *m, const double *in, int n, double *out);
/* MFSK pattern functions: bbi = n complex samples at the pass-band rate. */
int mo_time_sync_mfsk(const mo_mode *m, const double complex *bbi, int n, int search_start_symb);
double mo_detect_ack_pattern(const mo_mode *m, const double complex *bbi, int n, int use_break_tones, int *matched_out);
void mo_ack_pattern_baseband(const mo_mode *m, int use_break_tones, double complex *out /*[16 * 272]*/);
void mo_mfsk_tables(const mo_mode *m, int *out /*[32]*/);
int mo_generate_pattern_passband(mo_mode *m, int use_break_tones, double *out /*[16 * 272 * 4]*/, double *start_sample_inout);
double mo_detect_pattern_from_passband(const mo_mode *m, const double *data, int size, int use_break_tones, int *matched_out);
int mo_set_mfsk_ctrl_mode(mo_mode *m, int enable);
void mo_set_coarse_freq_sync(mo_mode *m, int enable);
void mo_reset_tx_stream(mo_mode *m);
int mo_transmit_byte_loc(mo_mode *m, const int *payload, int nBytes, double *out, double *start_sample_inout, int message_location);
double mo_receive_byte_timed(const mo_mode *m, const double *passband, int n_calls, int *decoded_flags);

#endif
