/*
 * oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin extern "C" driver around the UNMODIFIED reference physical layer, compiled
 * from the sources where they lie under /root/reference into oracle/_ref/libmercury_ref.so
 * (recipe: oracle/Makefile).  It only calls PUBLIC members of cl_telecom_system /
 * cl_ofdm / cl_psk / cl_ldpc in the order the reference's own receive_byte() does
 * (telecom_system.cc:1132-1341) so that every intermediate tensor of the RX hot path can
 * be exported as a golden vector, and it exposes the reference's own transmit_byte() /
 * receive_byte() unchanged for the passband loop-back case (BASELINE config #1).
 *
 * All reference calls run with fd 1 temporarily pointed at /dev/null: the reference prints
 * unconditional diagnostics from inside the hot path (ofdm.cc:1486, 1339; telecom_system.cc:1199).
 */
#include <fcntl.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <complex>

#include "physical_layer/telecom_system.h"
#include "gui/gui_state.h"

namespace {

struct QuietStdout {
	int saved;
	QuietStdout()
	{
		fflush(stdout);
		std::cout.flush();
		saved = dup(1);
		int devnull = open("/dev/null", O_WRONLY);
		dup2(devnull, 1);
		close(devnull);
	}
	~QuietStdout()
	{
		fflush(stdout);
		std::cout.flush();
		dup2(saved, 1);
		close(saved);
	}
};

struct Ref {
	cl_telecom_system *ts;
	int config;
};

inline cl_telecom_system &T(void *h) { return *static_cast<Ref *>(h)->ts; }

}  // namespace

extern "C" {

/* Geometry record filled by mref_geometry(); index names are mirrored in tests/refdrv.py. */
enum {
	G_NSYMB = 0, G_NC, G_NFFT, G_NGI, G_NOFDM, G_NDATA, G_NPILOTS, G_NBITS, G_N, G_K, G_P, G_M,
	G_PREAMBLE_NSYMB, G_FRAME_BYTES, G_ESTIMATOR, G_AMP_RESTORE, G_BIT_IL_BLOCK, G_TF_IL_BLOCK,
	G_INTERP_RATE, G_BUFFER_NSYMB, G_TOTAL_FRAME_SIZE, G_CWIDTH, G_VWIDTH, G_DWIDTH, G_LDPC_ITERS,
	G_OUTER_RESERVED, G_LS_WIN_H, G_LS_WIN_W, G_COUNT
};

void *mref_create(int config, int ldpc_iters)
{
	QuietStdout q;
	Ref *r = new Ref;
	r->ts = new cl_telecom_system;
	r->config = config;
	r->ts->operation_mode = BER_PLOT_baseband;
	r->ts->default_configurations_telecom_system.ldpc_nIteration_max = ldpc_iters;
	r->ts->load_configuration(config);
	return r;
}

/* cl_telecom_system::load_configuration(int) on the SAME object, as the ARQ layer does between data and ack configurations (arq_commander.cc:431,574,661) */
void mref_load_configuration(void *h, int config)
{
	QuietStdout q;
	Ref *r = static_cast<Ref *>(h);
	r->config = config;
	r->ts->load_configuration(config);
}

void mref_destroy(void *h)
{
	QuietStdout q;
	Ref *r = static_cast<Ref *>(h);
	delete r->ts;
	delete r;
}

void mref_geometry(void *h, int *g)
{
	cl_telecom_system &ts = T(h);
	g[G_NSYMB] = ts.ofdm.Nsymb;
	g[G_NC] = ts.ofdm.Nc;
	g[G_NFFT] = ts.ofdm.Nfft;
	g[G_NGI] = ts.data_container.Nofdm - ts.ofdm.Nfft; /* cl_ofdm::Ngi is private */
	g[G_NOFDM] = ts.data_container.Nofdm;
	g[G_NDATA] = ts.data_container.nData;
	g[G_NPILOTS] = ts.ofdm.pilot_configurator.nPilots;
	g[G_NBITS] = ts.data_container.nBits;
	g[G_N] = ts.ldpc.N;
	g[G_K] = ts.ldpc.K;
	g[G_P] = ts.ldpc.P;
	g[G_M] = (int)ts.M;
	g[G_PREAMBLE_NSYMB] = ts.data_container.preamble_nSymb;
	g[G_FRAME_BYTES] = ts.get_frame_size_bytes();
	g[G_ESTIMATOR] = ts.ofdm.channel_estimator;
	g[G_AMP_RESTORE] = ts.ofdm.channel_estimator_amplitude_restoration;
	g[G_BIT_IL_BLOCK] = ts.bit_interleaver_block_size;
	g[G_TF_IL_BLOCK] = ts.time_freq_interleaver_block_size;
	g[G_INTERP_RATE] = ts.data_container.interpolation_rate;
	g[G_BUFFER_NSYMB] = ts.data_container.buffer_Nsymb;
	g[G_TOTAL_FRAME_SIZE] = ts.data_container.total_frame_size;
	g[G_CWIDTH] = 0; /* private in cl_ldpc: see mref_ldpc_tables() */
	g[G_VWIDTH] = 0;
	g[G_DWIDTH] = 0;
	g[G_LDPC_ITERS] = ts.ldpc.nIteration_max;
	g[G_OUTER_RESERVED] = ts.outer_code_reserved_bits;
	g[G_LS_WIN_H] = ts.ofdm.LS_window_hight;
	g[G_LS_WIN_W] = ts.ofdm.LS_window_width;
}

/* Init-time tables. Any pointer may be NULL. */
void mref_tables(void *h, int *carrier_type /*[Nsymb*Nc]*/, double *pilot_seq /*[nPilots] (real)*/,
		 int *scrambler /*[N]*/, double *constellation /*[2*M]*/, double *pilot_boost /*[1]*/)
{
	cl_telecom_system &ts = T(h);
	int cells = ts.ofdm.Nsymb * ts.ofdm.Nc;
	if (carrier_type)
		for (int i = 0; i < cells; i++) carrier_type[i] = ts.ofdm.ofdm_frame[i].type;
	if (pilot_seq)
		for (int i = 0; i < ts.ofdm.pilot_configurator.nPilots; i++)
			pilot_seq[i] = ts.ofdm.pilot_configurator.sequence[i].real();
	if (scrambler)
		for (int i = 0; i < ts.ldpc.N; i++) scrambler[i] = ts.data_container.bit_energy_dispersal_sequence[i];
	if (constellation)
		for (int i = 0; i < (int)ts.M; i++) {
			/* cl_psk::constellation is private: map index i (MSB-first bits) through psk.mod() */
			int nb = 0, bits[8];
			while ((1 << nb) < (int)ts.M) nb++;
			for (int b = 0; b < nb; b++) bits[b] = (i >> (nb - 1 - b)) & 1;
			std::complex<double> pt;
			ts.psk.mod(bits, nb, &pt);
			constellation[2 * i] = pt.real();
			constellation[2 * i + 1] = pt.imag();
		}
	if (pilot_boost) *pilot_boost = ts.ofdm.pilot_configurator.boost;
}

/*
 * LDPC tables exactly as the reference holds them (mercury_normal_*_16.cc, bound in ldpc.cc:135-263).
 * cl_ldpc keeps its table pointers private, so the public globals are read directly, selected by the
 * rate numerator (1,2,3,4,5,6,8,14).  Returns 0 on success; widths come back in dims[3] = {Cwidth,Vwidth,dwidth}.
 */
#define MREF_RATE_CASE(n)                                                                     \
	case n:                                                                               \
		Cw = mercury_normal_Cwidth_##n##_16;                                          \
		Vw = mercury_normal_Vwidth_##n##_16;                                          \
		dw = mercury_normal_dwidth_##n##_16;                                          \
		pC = &mercury_normal_QCmatrixC_##n##_16[0][0];                                \
		pV = &mercury_normal_QCmatrixV_##n##_16[0][0];                                \
		pd = &mercury_normal_QCmatrixd_##n##_16[0];                                   \
		pE = &mercury_normal_QCmatrixEnc_##n##_16[0][0];                              \
		break;

int mref_ldpc_tables(int rate_num, int *dims, int *C /*[P*Cwidth]*/, int *V /*[N*Vwidth]*/, int *d /*[dwidth]*/, int *Enc /*[P*(Cwidth-1)]*/)
{
	int Cw = 0, Vw = 0, dw = 0;
	int *pC = nullptr, *pV = nullptr, *pd = nullptr, *pE = nullptr;
	switch (rate_num) {
		MREF_RATE_CASE(1)
		MREF_RATE_CASE(2)
		MREF_RATE_CASE(3)
		MREF_RATE_CASE(4)
		MREF_RATE_CASE(5)
		MREF_RATE_CASE(6)
		MREF_RATE_CASE(8)
		MREF_RATE_CASE(14)
	default:
		return -1;
	}
	int N = N_MAX, K = N * rate_num / 16, P = N - K;
	dims[0] = Cw;
	dims[1] = Vw;
	dims[2] = dw;
	if (C) memcpy(C, pC, sizeof(int) * P * Cw);
	if (V) memcpy(V, pV, sizeof(int) * N * Vw);
	if (d) memcpy(d, pd, sizeof(int) * dw);
	if (Enc) memcpy(Enc, pE, sizeof(int) * P * (Cw - 1));
	return 0;
}

/* The reference's vendored glibc PRNG (os_interop.cc:235-283): n draws after __srandom(seed). */
void mref_random(unsigned seed, int n, int *out)
{
	__srandom(seed);
	for (int i = 0; i < n; i++) out[i] = (int)__random();
}

int mref_crc16(const int *bytes, int n)
{
	return CRC16_MODBUS_RTU_calc(const_cast<int *>(bytes), n);
}

/*
 * TX bit chain of transmit_byte()/transmit_bit() (telecom_system.cc:342-416) followed by the
 * baseband modulation chain of baseband_test_EsN0() (telecom_system.cc:129-142): payload bytes ->
 * zero pad -> CRC16 -> scramble -> LDPC encode -> compaction -> bit interleave -> PSK map ->
 * T/F interleave -> framer -> IFFT + GI per symbol.  No pre-equalisation, no preamble, no
 * pass-band stage: the output is Nsymb*Nofdm complex baseband samples with exactly the scaling
 * symbol_demod() expects (i.e. after the "/sqrt(Nfft) ... *sqrt(Nfft)" pair of :139-153 cancels).
 * Optional outputs: the info bits after CRC (nReal), the codeword (N), the framed grid (Nsymb*Nc).
 */
void mref_tx_baseband(void *h, const int *payload, int nBytes, double *out_cplx, int *info_bits, int *codeword, double *framed)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	cl_data_container &dc = ts.data_container;
	int nReal = dc.nBits - ts.ldpc.P;
	int nVirtual = ts.ldpc.N - dc.nBits;
	int frame_size = (nReal - ts.outer_code_reserved_bits) / 8;
	int bytes[N_MAX / 8 + 8];
	for (int i = 0; i < frame_size; i++) bytes[i] = (i < nBytes) ? payload[i] : 0;

	byte_to_bit(bytes, dc.data_bit, frame_size);
	if (ts.outer_code == CRC16_MODBUS_RTU) {
		int crc = CRC16_MODBUS_RTU_calc(bytes, frame_size);
		int msB = (crc & 0xff00) >> 8, lsB = crc & 0x00ff;
		byte_to_bit(&lsB, &dc.data_bit[frame_size * 8], 1);
		byte_to_bit(&msB, &dc.data_bit[(frame_size + 1) * 8], 1);
	}
	for (int i = frame_size * 8 + ts.outer_code_reserved_bits; i < nReal; i++) dc.data_bit[i] = 0;
	if (info_bits) memcpy(info_bits, dc.data_bit, sizeof(int) * nReal);

	bit_energy_dispersal(dc.data_bit, dc.bit_energy_dispersal_sequence, dc.data_bit_energy_dispersal, nReal);
	for (int i = 0; i < nVirtual; i++) dc.data_bit_energy_dispersal[nReal + i] = dc.data_bit_energy_dispersal[i];
	ts.ldpc.encode(dc.data_bit_energy_dispersal, dc.encoded_data);
	if (codeword) memcpy(codeword, dc.encoded_data, sizeof(int) * ts.ldpc.N);
	for (int i = 0; i < ts.ldpc.P; i++) dc.encoded_data[nReal + i] = dc.encoded_data[i + ts.ldpc.K];
	interleaver(dc.encoded_data, dc.bit_interleaved_data, dc.nBits, ts.bit_interleaver_block_size);
	if (ts.M == MOD_MFSK) {
		for (int i = 0; i < dc.Nsymb * dc.Nc; i++) dc.ofdm_framed_data[i] = 0;
		ts.mfsk.mod(dc.bit_interleaved_data, ts.get_active_nbits(), dc.ofdm_framed_data); /* telecom_system.cc:411-416: one-hot tones, no framer */
	} else {
		ts.psk.mod(dc.bit_interleaved_data, dc.nBits, dc.modulated_data);
		interleaver(dc.modulated_data, dc.ofdm_time_freq_interleaved_data, dc.nData, ts.time_freq_interleaver_block_size);
		ts.ofdm.framer(dc.ofdm_time_freq_interleaved_data, dc.ofdm_framed_data);
	}
	if (framed) memcpy(framed, dc.ofdm_framed_data, sizeof(double) * 2 * dc.Nsymb * dc.Nc);
	for (int i = 0; i < dc.Nsymb; i++)
		ts.ofdm.symbol_mod(&dc.ofdm_framed_data[i * dc.Nc], &dc.ofdm_symbol_modulated_data[i * dc.Nofdm]);
	memcpy(out_cplx, dc.ofdm_symbol_modulated_data, sizeof(double) * 2 * dc.Nsymb * dc.Nofdm);
}

/* Per-stage outputs of the RX hot path; any pointer may be NULL. */
struct mref_rx_out {
	double *Y;	   /* [Nsymb*Nc*2]  after symbol_demod + AGC          (a2-a4)  */
	double *H;	   /* [Nsymb*Nc*2]  channel used by the equaliser     (a5-a8)  */
	double *Z;	   /* [Nsymb*Nc*2]  equalised grid                    (a9)     */
	float *llr_demod;  /* [nBits]       psk.demod output                  (a13)    */
	float *llr_cw;	   /* [N]           LLRs in codeword order            (a14)    */
	int *bits;	   /* [K]           decoder hard decisions (scrambled)(a15)    */
	int *bytes;	   /* [nReal/8]     de-scrambled, packed (incl. CRC)  (a16)    */
	int *payload;	   /* [frame_bytes] what receive_byte() writes to out (a16)    */
	double *stats;	   /* [8] iterations, crc, all_zeros, decoded, SNR, variance(float), agc_unused, mean_H */
};

/*
 * The RX tail exactly as receive_byte() runs it (telecom_system.cc:1135-1341 and the success
 * bookkeeping :1343-1375), driven through the public members, on Nsymb*Nofdm complex samples
 * (preamble already stripped: the reference indexes baseband_data at (pre+i)*Nofdm, :1137).
 */
void mref_rx_tail(void *h, const double *baseband, mref_rx_out *o)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	cl_data_container &dc = ts.data_container;
	cl_ofdm &ofdm = ts.ofdm;
	int nReal = dc.nBits - ts.ldpc.P;
	int nVirtual = ts.ldpc.N - dc.nBits;
	int cells = dc.Nsymb * dc.Nc;
	const std::complex<double> *bb = reinterpret_cast<const std::complex<double> *>(baseband);
	std::complex<double> *sym = new std::complex<double>[dc.Nofdm];

	for (int i = 0; i < dc.Nsymb; i++) {
		memcpy(sym, bb + (size_t)i * dc.Nofdm, sizeof(std::complex<double>) * dc.Nofdm);
		ofdm.symbol_demod(sym, &dc.ofdm_symbol_demodulated_data[i * dc.Nc]);
	}
	delete[] sym;
	if (ts.M == MOD_MFSK) {
		/* MFSK branch of the tail (telecom_system.cc:1142-1198): non-coherent energy detection, no AGC / channel estimate / gate */
		if (o->Y) memcpy(o->Y, dc.ofdm_symbol_demodulated_data, sizeof(double) * 2 * cells);
		ts.mfsk.demod(dc.ofdm_symbol_demodulated_data, ts.get_active_nbits(), dc.demodulated_data);
		for (int i = ts.get_active_nbits(); i < dc.nBits; i++) dc.demodulated_data[i] = 0.0f; /* :1188-1197: punctured positions = erasures */
		if (o->llr_demod) memcpy(o->llr_demod, dc.demodulated_data, sizeof(float) * dc.nBits);
		deinterleaver(dc.demodulated_data, dc.deinterleaved_data, dc.nBits, ts.bit_interleaver_block_size);
		for (int i = ts.ldpc.P - 1; i >= 0; i--) dc.deinterleaved_data[i + nReal + nVirtual] = dc.deinterleaved_data[i + nReal];
		for (int i = 0; i < nVirtual; i++) dc.deinterleaved_data[nReal + i] = dc.deinterleaved_data[i];
		if (o->llr_cw) memcpy(o->llr_cw, dc.deinterleaved_data, sizeof(float) * ts.ldpc.N);
		int iterations = ts.ldpc.decode(dc.deinterleaved_data, dc.hd_decoded_data_bit);
		if (o->bits) memcpy(o->bits, dc.hd_decoded_data_bit, sizeof(int) * ts.ldpc.K);
		bit_energy_dispersal(dc.hd_decoded_data_bit, dc.bit_energy_dispersal_sequence, dc.hd_decoded_data_bit, nReal);
		bit_to_byte(dc.hd_decoded_data_bit, dc.hd_decoded_data_byte, nReal);
		int all_zeros = YES;
		for (int i = 0; i < nReal / 8; i++)
			if (dc.hd_decoded_data_byte[i] != 0) {
				all_zeros = NO;
				break;
			}
		if (o->bytes) memcpy(o->bytes, dc.hd_decoded_data_byte, sizeof(int) * (nReal / 8));
		if (o->payload)
			for (int i = 0; i < (nReal - ts.outer_code_reserved_bits) / 8; i++) o->payload[i] = dc.hd_decoded_data_byte[i];
		int crc = 0;
		if (ts.outer_code == CRC16_MODBUS_RTU && all_zeros == NO) crc = CRC16_MODBUS_RTU_calc(dc.hd_decoded_data_byte, nReal / 8);
		int decoded = !(all_zeros == YES || crc != 0);
		if (o->stats) {
			o->stats[0] = iterations, o->stats[1] = crc, o->stats[2] = all_zeros, o->stats[3] = decoded;
			o->stats[4] = decoded ? 0.0 : -99.9; /* :1362-1367: MFSK reports SNR 0 */
			o->stats[5] = 0, o->stats[6] = 0, o->stats[7] = 1.0;
		}
		return;
	}
	ofdm.automatic_gain_control(dc.ofdm_symbol_demodulated_data);
	if (o->Y) memcpy(o->Y, dc.ofdm_symbol_demodulated_data, sizeof(double) * 2 * cells);

	if (ofdm.channel_estimator == ZERO_FORCE)
		ofdm.ZF_channel_estimator(dc.ofdm_symbol_demodulated_data);
	else
		ofdm.LS_channel_estimator(dc.ofdm_symbol_demodulated_data);

	double h_sum = 0;
	int h_measured = 0;
	for (int ci = 0; ci < cells; ci++)
		if (ofdm.estimated_channel[ci].status == MEASURED) {
			h_sum += std::abs(ofdm.estimated_channel[ci].value);
			h_measured++;
		}
	double mean_H = h_measured ? h_sum / h_measured : -1.0;

	if (ofdm.channel_estimator_amplitude_restoration == YES) {
		ofdm.restore_channel_amplitude();
		ofdm.channel_equalizer_without_amplitude_restoration(dc.ofdm_symbol_demodulated_data, dc.equalized_data_without_amplitude_restoration);
		ofdm.deframer(dc.equalized_data_without_amplitude_restoration, dc.ofdm_deframed_data_without_amplitude_restoration);
	}
	if (o->H)
		for (int ci = 0; ci < cells; ci++) {
			o->H[2 * ci] = ofdm.estimated_channel[ci].value.real();
			o->H[2 * ci + 1] = ofdm.estimated_channel[ci].value.imag();
		}
	ofdm.channel_equalizer(dc.ofdm_symbol_demodulated_data, dc.equalized_data);
	if (o->Z) memcpy(o->Z, dc.equalized_data, sizeof(double) * 2 * cells);

	float variance = ofdm.measure_variance(dc.equalized_data);

	ofdm.deframer(dc.equalized_data, dc.ofdm_deframed_data);
	deinterleaver(dc.ofdm_deframed_data, dc.ofdm_time_freq_deinterleaved_data, dc.nData, ts.time_freq_interleaver_block_size);
	ts.psk.demod(dc.ofdm_time_freq_deinterleaved_data, dc.nBits, dc.demodulated_data, variance);
	if (o->llr_demod) memcpy(o->llr_demod, dc.demodulated_data, sizeof(float) * dc.nBits);

	deinterleaver(dc.demodulated_data, dc.deinterleaved_data, dc.nBits, ts.bit_interleaver_block_size);
	for (int i = ts.ldpc.P - 1; i >= 0; i--) dc.deinterleaved_data[i + nReal + nVirtual] = dc.deinterleaved_data[i + nReal];
	for (int i = 0; i < nVirtual; i++) dc.deinterleaved_data[nReal + i] = dc.deinterleaved_data[i];
	if (o->llr_cw) memcpy(o->llr_cw, dc.deinterleaved_data, sizeof(float) * ts.ldpc.N);

	int iterations = ts.ldpc.decode(dc.deinterleaved_data, dc.hd_decoded_data_bit);
	if (o->bits) memcpy(o->bits, dc.hd_decoded_data_bit, sizeof(int) * ts.ldpc.K);

	bit_energy_dispersal(dc.hd_decoded_data_bit, dc.bit_energy_dispersal_sequence, dc.hd_decoded_data_bit, nReal);
	bit_to_byte(dc.hd_decoded_data_bit, dc.hd_decoded_data_byte, nReal);
	int all_zeros = YES;
	for (int i = 0; i < nReal / 8; i++)
		if (dc.hd_decoded_data_byte[i] != 0) {
			all_zeros = NO;
			break;
		}
	if (o->bytes) memcpy(o->bytes, dc.hd_decoded_data_byte, sizeof(int) * (nReal / 8));
	if (o->payload)
		for (int i = 0; i < (nReal - ts.outer_code_reserved_bits) / 8; i++) o->payload[i] = dc.hd_decoded_data_byte[i];
	int crc = 0;
	if (ts.outer_code == CRC16_MODBUS_RTU && all_zeros == NO) crc = CRC16_MODBUS_RTU_calc(dc.hd_decoded_data_byte, nReal / 8);

	int decoded = YES;
	double snr = -99.9;
	if (all_zeros == YES || (ts.outer_code == CRC16_MODBUS_RTU && crc != 0) ||
	    (ts.outer_code != CRC16_MODBUS_RTU && iterations > (ts.ldpc.nIteration_max - 1))) {
		decoded = NO;
	} else if (ofdm.channel_estimator == LEAST_SQUARE) {
		float v = variance;
		if (ofdm.channel_estimator_amplitude_restoration == YES) v = ofdm.measure_variance(dc.equalized_data_without_amplitude_restoration);
		snr = 10.0 * log10(1.0 / v);
	} else {
		/* ZF: the reference re-encodes the decoded frame and measures the distance of the equalised data symbols to it
		 * (telecom_system.cc:1376-1400), with the reference's own functions in the reference's order */
		bit_energy_dispersal(dc.hd_decoded_data_bit, dc.bit_energy_dispersal_sequence, dc.hd_decoded_data_bit, nReal);
		for (int i = 0; i < nVirtual; i++) dc.hd_decoded_data_bit[nReal + i] = dc.hd_decoded_data_bit[i];
		ts.ldpc.encode(dc.hd_decoded_data_bit, dc.encoded_data);
		for (int i = 0; i < ts.ldpc.P; i++) dc.encoded_data[nReal + i] = dc.encoded_data[i + ts.ldpc.K];
		interleaver(dc.encoded_data, dc.bit_interleaved_data, dc.nBits, ts.bit_interleaver_block_size);
		ts.psk.mod(dc.bit_interleaved_data, dc.nBits, dc.modulated_data);
		interleaver(dc.modulated_data, dc.ofdm_time_freq_interleaved_data, dc.nData, ts.time_freq_interleaver_block_size);
		if (ofdm.channel_estimator_amplitude_restoration == YES)
			snr = ofdm.measure_SNR(dc.ofdm_deframed_data_without_amplitude_restoration, dc.ofdm_time_freq_interleaved_data, dc.nData);
		else
			snr = ofdm.measure_SNR(dc.ofdm_deframed_data, dc.ofdm_time_freq_interleaved_data, dc.nData);
	}
	if (o->stats) {
		o->stats[0] = iterations;
		o->stats[1] = crc;
		o->stats[2] = all_zeros;
		o->stats[3] = decoded;
		o->stats[4] = snr;
		o->stats[5] = variance;
		o->stats[6] = 0;
		o->stats[7] = mean_H;
	}
}

/* The same, repeated over a batch, timed: returns seconds of wall time for n frames (CPU baseline leg). */
double mref_rx_tail_timed(void *h, const double *baseband, int n_frames, int *payloads, int *decoded_flags, int *iterations)
{
	cl_telecom_system &ts = T(h);
	cl_data_container &dc = ts.data_container;
	size_t stride = (size_t)dc.Nsymb * dc.Nofdm * 2;
	int fb = ts.get_frame_size_bytes();
	double st[8];
	auto t0 = std::chrono::steady_clock::now();
	for (int f = 0; f < n_frames; f++) {
		mref_rx_out o;
		memset(&o, 0, sizeof(o));
		o.stats = st;
		o.payload = payloads ? payloads + (size_t)f * fb : nullptr;
		mref_rx_tail(h, baseband + f * stride, &o);
		if (decoded_flags) decoded_flags[f] = (int)st[3];
		if (iterations) iterations[f] = (int)st[0];
	}
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

/* LDPC decoder alone on codeword-order LLRs (ldpc.cc:266-278 -> ldpc_decoder_SPA.cc:25-218). */
int mref_ldpc_decode(void *h, const float *llr_cw, int *bits_out /*[K]*/)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	return ts.ldpc.decode(llr_cw, bits_out);
}

/* Reference pass-band TX (telecom_system.cc:342-634), SINGLE_MESSAGE; returns total_frame_size samples written. */
int mref_transmit_byte(void *h, const int *payload, int nBytes, double *passband_out)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	int buf[N_MAX];
	memset(buf, 0, sizeof(buf));
	for (int i = 0; i < nBytes; i++) buf[i] = payload[i];
	ts.transmit_byte(buf, nBytes, passband_out, SINGLE_MESSAGE);
	return ts.data_container.total_frame_size;
}

/*
 * Reference receive_byte() unchanged (telecom_system.cc:646-1518) on a full capture buffer of
 * Nofdm*buffer_Nsymb*interpolation_rate doubles.  Also copies out the post-synchronisation
 * baseband_data ((pre+Nsymb)*Nofdm complex) that its own hot path consumed, which is the input of
 * the batch entry point of the B200 path.
 */
void mref_receive_byte(void *h, const double *passband, int *out, double *stats /*[8]*/, double *baseband_out)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	cl_data_container &dc = ts.data_container;
	size_t n = (size_t)dc.Nofdm * dc.buffer_Nsymb * dc.interpolation_rate;
	double *copy = new double[n];
	memcpy(copy, passband, sizeof(double) * n);
	int obuf[N_MAX];
	memset(obuf, 0, sizeof(obuf));
	st_receive_stats rs = ts.receive_byte(copy, obuf);
	delete[] copy;
	for (int i = 0; i < ts.get_frame_size_bytes(); i++) out[i] = obuf[i];
	stats[0] = rs.iterations_done;
	stats[1] = rs.crc;
	stats[2] = rs.all_zeros;
	stats[3] = rs.message_decoded;
	stats[4] = rs.SNR;
	stats[5] = rs.delay;
	stats[6] = rs.sync_trials;
	stats[7] = rs.freq_offset;
	if (baseband_out) memcpy(baseband_out, dc.baseband_data, sizeof(double) * 2 * (dc.Nsymb + dc.preamble_nSymb) * dc.Nofdm);
}

/*
 * Reference receive_byte() unchanged, with the cross-call link state made explicit: state[0] =
 * receive_stats.delay_of_last_decoded_message, state[1] = receive_stats.freq_offset_of_last_decoded_message are
 * written into the object before the call and read back after it (telecom_system.cc:945-947,1108-1110,1423-1427).
 * stats[12]: iterations, crc, all_zeros, decoded, SNR, delay, sync_trials, freq_offset, coarse_metric,
 * signal_stregth_dbm, buffer samples, frame_bytes.
 */
void mref_receive_byte2(void *h, const double *passband, int *out, double *stats /*[12]*/, double *state /*[5]*/, double *baseband_out)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	cl_data_container &dc = ts.data_container;
	size_t n = (size_t)dc.Nofdm * dc.buffer_Nsymb * dc.interpolation_rate;
	double *copy = new double[n];
	memcpy(copy, passband, sizeof(double) * n);
	int obuf[N_MAX];
	memset(obuf, 0, sizeof(obuf));
	ts.receive_stats.delay_of_last_decoded_message = (int)state[0];
	ts.receive_stats.freq_offset_of_last_decoded_message = state[1];
	ts.receive_stats.mfsk_search_raw = (int)state[2]; /* MFSK: first symbol of the preamble search (telecom_system.cc:684-686) */
	ts.mfsk_fixed_delay = (int)state[4]; /* >= 0: the ARQ layer's overflow recapture bypasses the search once (arq_common.cc:2830-2833, telecom_system.cc:663-673) */
	dc.nUnder_processing_events = 0;
	ts.receive_stats.freq_offset = 0;
	ts.receive_stats.coarse_metric = 0;
	ts.receive_stats.iterations_done = 0;
	ts.receive_stats.crc = 0;
	ts.receive_stats.all_zeros = 0;
	ts.receive_stats.SNR = 0;
	st_receive_stats rs = ts.receive_byte(copy, obuf);
	delete[] copy;
	for (int i = 0; i < ts.get_frame_size_bytes(); i++) out[i] = obuf[i];
	stats[0] = rs.iterations_done;
	stats[1] = rs.crc;
	stats[2] = rs.all_zeros;
	stats[3] = rs.message_decoded;
	stats[4] = rs.SNR;
	stats[5] = rs.delay;
	stats[6] = rs.sync_trials;
	stats[7] = rs.freq_offset;
	stats[8] = rs.coarse_metric;
	stats[9] = rs.signal_stregth_dbm;
	stats[10] = (double)n;
	stats[11] = ts.get_frame_size_bytes();
	state[0] = rs.delay_of_last_decoded_message;
	state[1] = rs.freq_offset_of_last_decoded_message;
	state[3] = rs.frame_overflow_symbols;
	state[4] = ts.mfsk_fixed_delay; /* consumed: -1 after the call (:669) */
	if (baseband_out) memcpy(baseband_out, dc.baseband_data, sizeof(double) * 2 * (dc.Nsymb + dc.preamble_nSymb) * dc.Nofdm);
}

/* Wall time (seconds) of n_calls unchanged receive_byte() calls on consecutive capture buffers (CPU baseline of the front-end row). */
double mref_receive_byte_timed(void *h, const double *passband, int n_calls, int *decoded_flags)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	cl_data_container &dc = ts.data_container;
	size_t n = (size_t)dc.Nofdm * dc.buffer_Nsymb * dc.interpolation_rate;
	double *copy = new double[n];
	int obuf[N_MAX];
	auto t0 = std::chrono::steady_clock::now();
	for (int c = 0; c < n_calls; c++) {
		memcpy(copy, passband + (size_t)c * n, sizeof(double) * n);
		ts.receive_stats.delay_of_last_decoded_message = -1;
		ts.receive_stats.freq_offset_of_last_decoded_message = 0;
		st_receive_stats rs = ts.receive_byte(copy, obuf);
		if (decoded_flags) decoded_flags[c] = rs.message_decoded == YES;
	}
	auto t1 = std::chrono::steady_clock::now();
	delete[] copy;
	return std::chrono::duration<double>(t1 - t0).count();
}

/* The reference's own FIR designs (fir_filter.cc:45-163) and front-end constants, for pinning the restatement. */
void mref_frontend_tables(void *h, int *ntaps /*[2]*/, double *ts_coef, double *data_coef, double *consts /*[8]*/)
{
	cl_telecom_system &ts = T(h);
	ntaps[0] = ts.ofdm.FIR_rx_time_sync.filter_nTaps;
	ntaps[1] = ts.ofdm.FIR_rx_data.filter_nTaps;
	/* the coefficient array is private: read it as the response to a centred unit impulse (apply(), fir_filter.cc:189-210;
	 * every other product is an exact zero, so the response equals the coefficients bit for bit) */
	for (int f = 0; f < 2; f++) {
		cl_FIR &fir = f == 0 ? ts.ofdm.FIR_rx_time_sync : ts.ofdm.FIR_rx_data;
		int nt = ntaps[f];
		double *imp = new double[nt](), *resp = new double[nt]();
		imp[(nt - 1) / 2] = 1.0;
		fir.apply(imp, resp, nt);
		for (int i = 0; i < nt; i++) (f == 0 ? ts_coef : data_coef)[i] = resp[i];
		delete[] imp;
		delete[] resp;
	}
	consts[0] = ts.sampling_frequency;
	consts[1] = ts.carrier_frequency;
	consts[2] = ts.carrier_amplitude;
	consts[3] = ts.bandwidth;
	consts[4] = ts.time_sync_trials_max;
	consts[5] = ts.use_last_good_time_sync;
	consts[6] = ts.use_last_good_freq_offset;
	consts[7] = ts.ofdm.freq_offset_ignore_limit;
}

/* TX-side tables of the reference for pinning the restatement: preamble carriers (value, type), pre-equalisation channel
 * (telecom_system.cc:3108-3146), the two transmit FIRs read as impulse responses, and the running carrier sample counter. */
void mref_tx_tables(void *h, double *preamble /*[pre*Nc*2]*/, int *preamble_type /*[pre*Nc]*/, double *pre_eq /*[Nc*2]*/, int *ntaps /*[2]*/,
		    double *tx1_coef, double *tx2_coef, double *consts /*[8]*/)
{
	cl_telecom_system &ts = T(h);
	int n = ts.data_container.preamble_nSymb * ts.data_container.Nc;
	for (int i = 0; i < n; i++) {
		preamble[2 * i] = ts.ofdm.ofdm_preamble[i].value.real();
		preamble[2 * i + 1] = ts.ofdm.ofdm_preamble[i].value.imag();
		preamble_type[i] = ts.ofdm.ofdm_preamble[i].type;
	}
	for (int j = 0; j < ts.data_container.Nc; j++) {
		pre_eq[2 * j] = ts.pre_equalization_channel[j].value.real();
		pre_eq[2 * j + 1] = ts.pre_equalization_channel[j].value.imag();
	}
	ntaps[0] = ts.ofdm.FIR_tx1.filter_nTaps;
	ntaps[1] = ts.ofdm.FIR_tx2.filter_nTaps;
	for (int f = 0; f < 2; f++) {
		cl_FIR &fir = f == 0 ? ts.ofdm.FIR_tx1 : ts.ofdm.FIR_tx2;
		int nt = ntaps[f];
		double *imp = new double[nt](), *resp = new double[nt]();
		imp[(nt - 1) / 2] = 1.0;
		fir.apply(imp, resp, nt);
		for (int i = 0; i < nt; i++) (f == 0 ? tx1_coef : tx2_coef)[i] = resp[i];
		delete[] imp;
		delete[] resp;
	}
	consts[0] = ts.output_power_Watt;
	consts[1] = ts.ofdm.preamble_configurator.boost;
	consts[2] = ts.ofdm.preamble_papr_cut;
	consts[3] = ts.ofdm.data_papr_cut;
	consts[4] = (double)ts.ofdm.passband_start_sample;
	consts[5] = ts.data_container.total_frame_size;
	consts[6] = PREAMBLE;
	consts[7] = ZERO;
}

/* transmit_byte(SINGLE_MESSAGE) from a chosen carrier phase: sets ofdm.passband_start_sample first, returns it afterwards. */
int mref_transmit_byte2(void *h, const int *payload, int nBytes, double *passband_out, double *start_sample_inout)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	int buf[N_MAX];
	memset(buf, 0, sizeof(buf));
	for (int i = 0; i < nBytes; i++) buf[i] = payload[i];
	ts.ofdm.passband_start_sample = (long unsigned)*start_sample_inout;
	ts.transmit_byte(buf, nBytes, passband_out, SINGLE_MESSAGE);
	*start_sample_inout = (double)ts.ofdm.passband_start_sample;
	return ts.data_container.total_frame_size;
}

/* MFSK pattern functions on a pass-band-rate base-band buffer: time_sync_mfsk (ofdm.cc:1969-2065) and
 * detect_ack_pattern (ofdm.cc:2067-2186) with the ACK or BREAK tones of the loaded ROBUST configuration; bbi = n complex doubles. */
int mref_time_sync_mfsk(void *h, const double *bbi, int n, int search_start_symb)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	std::complex<double> *b = const_cast<std::complex<double> *>(reinterpret_cast<const std::complex<double> *>(bbi));
	return ts.ofdm.time_sync_mfsk(b, n, ts.data_container.interpolation_rate, ts.data_container.preamble_nSymb, ts.mfsk.preamble_tones, ts.mfsk.M,
				      ts.mfsk.nStreams, ts.mfsk.stream_offsets, search_start_symb);
}

double mref_detect_ack_pattern(void *h, const double *bbi, int n, int use_break_tones, int *matched)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	std::complex<double> *b = const_cast<std::complex<double> *>(reinterpret_cast<const std::complex<double> *>(bbi));
	return ts.ofdm.detect_ack_pattern(b, n, ts.data_container.interpolation_rate, cl_mfsk::ACK_PATTERN_NSYMB,
					  use_break_tones ? ts.mfsk.break_tones : ts.mfsk.ack_tones, cl_mfsk::ACK_PATTERN_LEN, ts.mfsk.tone_hop_step,
					  ts.mfsk.M, ts.mfsk.nStreams, ts.mfsk.stream_offsets, matched);
}

/* The ACK / BREAK pattern as carriers (mfsk.cc:197-252) and symbol-modulated base-band (16 symbols x Nofdm), for building test buffers. */
void mref_ack_pattern_baseband(void *h, int use_break_tones, double *out_cplx)
{
	cl_telecom_system &ts = T(h);
	int Nc = ts.data_container.Nc, No = ts.data_container.Nofdm;
	std::complex<double> *pat = new std::complex<double>[cl_mfsk::ACK_PATTERN_NSYMB * Nc];
	if (use_break_tones) ts.mfsk.generate_break_pattern(pat);
	else ts.mfsk.generate_ack_pattern(pat);
	std::complex<double> *o = reinterpret_cast<std::complex<double> *>(out_cplx);
	for (int s = 0; s < cl_mfsk::ACK_PATTERN_NSYMB; s++) ts.ofdm.symbol_mod(pat + s * Nc, o + s * No);
	delete[] pat;
}

void mref_mfsk_tables(void *h, int *out /*[32]: M, nBits, nStreams, tone_hop_step, stream_offsets[4], preamble_tones[4], ack_tones[8], break_tones[8]*/)
{
	cl_telecom_system &ts = T(h);
	int k = 0;
	out[k++] = ts.mfsk.M, out[k++] = ts.mfsk.nBits, out[k++] = ts.mfsk.nStreams, out[k++] = ts.mfsk.tone_hop_step;
	for (int i = 0; i < 4; i++) out[k++] = ts.mfsk.stream_offsets[i];
	for (int i = 0; i < 4; i++) out[k++] = ts.mfsk.preamble_tones[i];
	for (int i = 0; i < 8; i++) out[k++] = ts.mfsk.ack_tones[i];
	for (int i = 0; i < 8; i++) out[k++] = ts.mfsk.break_tones[i];
}

/* The ARQ layer's TX path (arq_common.cc:2224-2247): transmit_byte(..., NO_FILTER_MESSAGE) per frame, then the two transmit FIRs over the
 * whole padded batch buffer through the public ofdm.FIR_tx1 / FIR_tx2. */
int mref_transmit_byte_nofilter(void *h, const int *payload, int nBytes, double *passband_out, double *start_sample_inout)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	int buf[N_MAX];
	memset(buf, 0, sizeof(buf));
	for (int i = 0; i < nBytes; i++) buf[i] = payload[i];
	ts.ofdm.passband_start_sample = (long unsigned)*start_sample_inout;
	ts.transmit_byte(buf, nBytes, passband_out, NO_FILTER_MESSAGE);
	*start_sample_inout = (double)ts.ofdm.passband_start_sample;
	return ts.data_container.total_frame_size;
}

void mref_fir_tx_apply(void *h, const double *in, int n, double *out)
{
	cl_telecom_system &ts = T(h);
	double *tmp = new double[n]();
	ts.ofdm.FIR_tx1.apply(const_cast<double *>(in), tmp, n);
	ts.ofdm.FIR_tx2.apply(tmp, out, n);
	delete[] tmp;
}

/* The ARQ-facing tone-pattern calls (telecom_system.h:122-130, .cc:1589-1710): config-independent, dedicated ack_mfsk (M = 16, 1 stream). */
int mref_generate_pattern_passband(void *h, int use_break_tones, double *out, double *start_sample_inout)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	ts.ofdm.passband_start_sample = (long unsigned)*start_sample_inout;
	int n = use_break_tones ? ts.generate_break_pattern_passband(out) : ts.generate_ack_pattern_passband(out);
	*start_sample_inout = (double)ts.ofdm.passband_start_sample;
	return n;
}

double mref_detect_pattern_from_passband(void *h, const double *data, int size, int use_break_tones, int *matched)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	double *copy = new double[size];
	memcpy(copy, data, sizeof(double) * size);
	double m = use_break_tones ? ts.detect_break_pattern_from_passband(copy, size, matched) : ts.detect_ack_pattern_from_passband(copy, size, matched);
	delete[] copy;
	return m;
}

/* MFSK control frames (telecom_system.cc:1572-1585, 2966-2995): shortened frames in ROBUST_0 / ROBUST_1. Returns get_active_nsymb(). */
int mref_set_mfsk_ctrl_mode(void *h, int enable)
{
	cl_telecom_system &ts = T(h);
	ts.set_mfsk_ctrl_mode(enable != 0);
	return ts.get_active_nsymb();
}

/* The optional coarse frequency search of trial 1 (telecom_system.cc:949-1013) is switched by g_gui_state.coarse_freq_sync_enabled
 * (gui_state.h:143, off by default; the GUI / ini file turn it on for HF radio use). */
void mref_set_coarse_freq_sync(int enable) { g_gui_state.coarse_freq_sync_enabled.store(enable != 0); }

/* transmit_byte with any message_location (FIRST 0, MIDDLE 1, FLUSH 2, SINGLE 3, NO_FILTER 4): the streaming locations keep the reference's
 * own three-frame filter buffer between calls (telecom_system.cc:559-594). */
int mref_transmit_byte_loc(void *h, const int *payload, int nBytes, double *passband_out, double *start_sample_inout, int message_location)
{
	QuietStdout q;
	cl_telecom_system &ts = T(h);
	int buf[N_MAX];
	memset(buf, 0, sizeof(buf));
	for (int i = 0; i < nBytes; i++) buf[i] = payload[i];
	ts.ofdm.passband_start_sample = (long unsigned)*start_sample_inout;
	ts.transmit_byte(buf, nBytes, passband_out, message_location);
	*start_sample_inout = (double)ts.ofdm.passband_start_sample;
	return ts.data_container.total_frame_size;
}

/* zero the reference's streaming filter buffer (it is allocated uninitialised) so that a sequence starts from a defined state */
void mref_reset_tx_stream(void *h)
{
	cl_telecom_system &ts = T(h);
	for (int i = 0; i < 3 * ts.data_container.total_frame_size; i++) ts.data_container.passband_data_tx_buffer[i] = 0;
}

}  // extern "C"
