/*
 * oracle/mercury_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product).
 *
 * Plain-C (C11 + C99 complex), double-precision CPU restatement of the Mercury physical-layer RX hot
 * path and of the TX chain that synthesises its inputs.  All paths below are relative to
 * /root/reference/ (Rhizomatica/mercury @ c91aa4b).  Operation order follows the reference so that,
 * compiled by the same gcc without FMA contraction, results are bit-identical to oracle/_ref
 * (tests/test_oracle_vs_ref.py); complex products / quotients use C99 `double complex`, which gcc
 * lowers to the same libgcc helpers as std::complex<double>.
 *
 * Parity status: PINNED -- against the unmodified reference (oracle/_ref, this container) and against
 * the committed fixtures in tests/golden/ that were generated from it (tests/golden/make_golden.py).
 */
#define _GNU_SOURCE
#include "mercury_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------
 * PRNG: glibc random() TYPE_3, as vendored in source/common/os_interop.cc:100-283 (__srandom/__random).
 * Restated in its textbook form: additive lagged Fibonacci r[i] = r[i-3] + r[i-31] (mod 2^32), seeded by
 * the Park-Miller LCG (16807, Schrage's method), first 310 outputs discarded, output = r >> 1.
 * seed 0 is replaced by 1 (os_interop.cc:252-253).
 * ---------------------------------------------------------------------------------------------- */
static uint32_t g_rng[34];
static int g_rng_i;

void mo_srandom(unsigned seed)
{
	int32_t r[34];
	if (seed == 0) seed = 1;
	r[0] = (int32_t)seed;
	for (int i = 1; i < 31; i++) {
		long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
		long w = 16807 * lo - 2836 * hi;
		if (w < 0) w += 2147483647;
		r[i] = (int32_t)w;
	}
	for (int i = 31; i < 34; i++) r[i] = r[i - 31];
	for (int i = 0; i < 34; i++) g_rng[i] = (uint32_t)r[i];
	g_rng_i = 0;
	for (int i = 34; i < 344; i++) (void)mo_random();
}

int mo_random(void)
{
	/* ring of 34: slot i holds r[n], r[n-31] is slot (i+3)%34, r[n-3] is slot (i+31)%34 */
	uint32_t v = g_rng[(g_rng_i + 3) % 34] + g_rng[(g_rng_i + 31) % 34];
	g_rng[g_rng_i] = v;
	g_rng_i = (g_rng_i + 1) % 34;
	return (int)(v >> 1);
}

void mo_random_seq(unsigned seed, int n, int *out)
{
	mo_srandom(seed);
	for (int i = 0; i < n; i++) out[i] = mo_random();
}

/* CRC-16/MODBUS: source/physical_layer/crc16_modbus_rtu.cc:25-45 (init 0xFFFF, reflected poly 0xA001). */
int mo_crc16(const int *bytes, int n)
{
	uint16_t crc = 0xffff;
	for (int j = 0; j < n; j++) {
		crc ^= (uint16_t)(bytes[j] & 0xFF);
		for (int i = 0; i < 8; i++) crc = (crc & 1) ? (uint16_t)((crc >> 1) ^ 0xA001) : (uint16_t)(crc >> 1);
	}
	return crc;
}

/* ------------------------------------------------------------------------------------------------
 * Mode table: source/physical_layer/telecom_system.cc:2506-2624 (CONFIG_0..16), Nsymb by modulation
 * :1818-1826 (HIGH_DENSITY, the compiled default physical_config.cc:48), phase-only flag :2647-2654.
 * ---------------------------------------------------------------------------------------------- */
static const struct { int M, rate, pre, est; } k_modes[17] = {
	{2, 1, 4, 1},  {2, 2, 4, 1},  {2, 3, 4, 1},  {2, 4, 4, 1},  {2, 5, 4, 1},  {2, 6, 4, 1},
	{2, 8, 4, 1},  {4, 5, 4, 1},  {4, 6, 4, 1},  {4, 8, 4, 1},  {8, 6, 3, 1},  {8, 8, 3, 1},
	{4, 14, 3, 1}, {16, 8, 2, 1}, {8, 14, 2, 1}, {16, 14, 2, 0}, {32, 14, 1, 0},
};

static int nsymb_of(int M)
{
	switch (M) {
	case 2: return 48;
	case 4: return 24;
	case 8: return 16;
	case 16: return 12;
	case 32: return 9;
	case 64: return 8;
	}
	return 0;
}

/* Constellations: source/physical_layer/psk.cc:65-226, unit mean power normalisation :229-256
 * (the normaliser is accumulated and kept in *float*, :231,245-248). */
static void build_constellation(mo_mode *m)
{
	static const signed char q16[16][2] = {{-3, 3}, {-3, 1}, {-3, -3}, {-3, -1}, {-1, 3}, {-1, 1}, {-1, -3}, {-1, -1},
					       {3, 3},	{3, 1},	 {3, -3},  {3, -1},  {1, 3},  {1, 1},  {1, -3},	 {1, -1}};
	static const signed char q32[32][2] = {{-3, 5}, {-1, 5}, {-3, -5}, {-1, -5}, {-5, 3}, {-5, 1}, {-5, -3}, {-5, -1},
					       {-1, 3}, {-1, 1}, {-1, -3}, {-1, -1}, {-3, 3}, {-3, 1}, {-3, -3}, {-3, -1},
					       {3, 5},	{1, 5},	 {3, -5},  {1, -5},  {5, 3},  {5, 1},  {5, -3},	 {5, -1},
					       {1, 3},	{1, 1},	 {1, -3},  {1, -1},  {3, 3},  {3, 1},  {3, -3},	 {3, -1}};
	double complex *c = m->constellation;
	double h = sqrt(2.0) / 2.0;
	if (m->M == 2) {
		c[0] = 1;
		c[1] = -1;
	} else if (m->M == 4) {
		c[0] = -1 + 1 * I;
		c[1] = -1 - 1 * I;
		c[2] = 1 + 1 * I;
		c[3] = 1 - 1 * I;
	} else if (m->M == 8) {
		c[0] = (-1 - 1 * I) * h;
		c[1] = -1;
		c[2] = 1 * I;
		c[3] = (-1 + 1 * I) * h;
		c[4] = -1 * I;
		c[5] = (1 - 1 * I) * h;
		c[6] = (1 + 1 * I) * h;
		c[7] = 1;
	} else if (m->M == 16) {
		for (int i = 0; i < 16; i++) c[i] = q16[i][0] + q16[i][1] * I;
	} else if (m->M == 32) {
		for (int i = 0; i < 32; i++) c[i] = q32[i][0] + q32[i][1] * I;
	}
	float pn = 0;
	for (int i = 0; i < m->M; i++) pn += creal(c[i]) * creal(c[i]) + cimag(c[i]) * cimag(c[i]);
	pn = 1 / (sqrt(pn / m->M));
	for (int i = 0; i < m->M; i++) c[i] *= pn;
}

/* Pilot lattice: cl_pilot_configurator::configure, source/physical_layer/ofdm.cc:976-1064 with the
 * defaults Dx=1 (telecom_system.cc:1838-1846), Dy=3 (:1848-1858), all edge options DATA, last_col AUTO
 * (physical_config.cc:40-46).  Pilot DBPSK sequence: ofdm.cc:940-951, seed 0 (physical_config.cc:47). */
static void build_pilots(mo_mode *m)
{
	enum { DATA = 0, PILOT = 1 };
	int Dx = 1, Dy = 3;
	int ncm = m->Nc > m->Nsymb ? m->Nc : m->Nsymb;
	unsigned char *v = calloc((size_t)ncm * ncm, 1);
	int x = 0, y = 0;
	while (x < ncm && y < ncm) {
		for (int j = y; j < ncm; j += Dy) v[j * ncm + x] = PILOT;
		for (int j = y; j >= 0; j -= Dy) v[j * ncm + x] = PILOT;
		y++;
		x += Dx;
	}
	int cnt = 0;
	for (int j = 0; j < m->Nsymb; j++) cnt += v[j * ncm + m->Nc - 1] == PILOT;
	if (cnt < 2) /* last_col AUTO_SELLECT -> COPY_FIRST_COL (ofdm.cc:1004-1007,1031-1034) */
		for (int j = 0; j < ncm; j++) v[j * ncm + m->Nc - 1] = v[j * ncm + 0];
	m->nPilots = 0;
	m->nData = m->Nc * m->Nsymb;
	for (int j = 0; j < m->Nsymb; j++)
		for (int i = 0; i < m->Nc; i++) {
			m->is_pilot[j * m->Nc + i] = v[j * ncm + i] == PILOT;
			if (v[j * ncm + i] == PILOT) {
				m->nPilots++;
				m->nData--;
			}
		}
	free(v);
	mo_srandom(0);
	int last = 0, k = 0;
	for (int c = 0; c < m->Nsymb * m->Nc; c++) {
		m->pilot_val[c] = 0;
		if (!m->is_pilot[c]) continue;
		int pv = (mo_random() % 2) ^ last;
		m->pilot_val[c] = (double)(2 * pv - 1) * m->boost;
		last = pv;
		k++;
	}
}

/* FFT tables: cl_ofdm::init_fft_tables, source/physical_layer/ofdm.cc:256-290. */
static void build_fft_tables(mo_mode *m)
{
	int n = MO_NFFT;
	for (int k = 0; k < n / 2; k++) {
		double angle = -2.0 * M_PI * k / n;
		m->twiddle[k] = cos(angle) + sin(angle) * I;
	}
	for (int i = 0; i < n; i++) {
		int rev = 0;
		for (int j = 0; j < 8; j++)
			if (i & (1 << j)) rev |= 1 << (7 - j);
		m->bitrev[i] = rev;
	}
}

/* LDPC tables from mercury_b200/data/ldpc_tables.bin (format: tools/extract_ldpc_tables.py), rebuilt into
 * the reference's QCmatrixC / QCmatrixV layout (mercury_normal_*_16.cc) and V_pos (ldpc_decoder_SPA.cc:81-104). */
static int load_ldpc(mo_mode *m, const char *path)
{
	FILE *f = fopen(path, "rb");
	if (!f) return -1;
	unsigned char hdr[12];
	if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "MLDP", 4) != 0) {
		fclose(f);
		return -2;
	}
	uint32_t nrates;
	memcpy(&nrates, hdr + 8, 4);
	int rc = -3;
	for (uint32_t r = 0; r < nrates; r++) {
		uint16_t h[6];
		uint32_t ne;
		if (fread(h, 2, 6, f) != 6 || fread(&ne, 4, 1, f) != 1) break;
		int N = h[1], P = h[3], Cw = h[4], Vw = h[5];
		uint16_t *cdeg = malloc(2 * P), *ev = malloc(2 * ne), *vdeg = malloc(2 * N), *vc = malloc(2 * ne);
		int ok = fread(cdeg, 2, P, f) == (size_t)P && fread(ev, 2, ne, f) == ne && fread(vdeg, 2, N, f) == (size_t)N &&
			 fread(vc, 2, ne, f) == ne;
		if (ok && h[0] == m->rate_num) {
			m->Cwidth = Cw;
			m->Vwidth = Vw;
			m->n_edges = (int)ne;
			m->C = malloc(sizeof(int) * P * Cw);
			m->V = malloc(sizeof(int) * N * Vw);
			m->Vpos = malloc(sizeof(int) * P * Cw);
			m->vdeg = malloc(sizeof(int) * N);
			for (int i = 0; i < P * Cw; i++) m->C[i] = -1;
			for (int i = 0; i < N * Vw; i++) m->V[i] = -1;
			int e = 0;
			for (int c = 0; c < P; c++)
				for (int j = 0; j < cdeg[c]; j++) m->C[c * Cw + j] = ev[e++];
			e = 0;
			for (int v = 0; v < N; v++) {
				m->vdeg[v] = vdeg[v];
				for (int j = 0; j < vdeg[v]; j++) m->V[v * Vw + j] = vc[e++];
			}
			for (int c = 0; c < P; c++)
				for (int j = 0; j < Cw; j++) {
					int v = m->C[c * Cw + j], pos = -1;
					if (v != -1)
						for (int k = 0; k < Vw; k++)
							if (m->V[v * Vw + k] == c) {
								pos = k;
								break;
							}
					m->Vpos[c * Cw + j] = pos;
				}
			rc = 0;
		}
		free(cdeg);
		free(ev);
		free(vdeg);
		free(vc);
		if (!ok || rc == 0) break;
	}
	fclose(f);
	return rc;
}

/* cl_telecom_system::load_configuration(int) + init(): source/physical_layer/telecom_system.cc:2487-3025, 1804-1982. */
/* cl_mfsk::init: source/physical_layer/mfsk.cc:49-160 (tone plan, hop step, preamble / ACK / BREAK tone sequences). */
static void mfsk_init(mo_mfsk *f, int M, int Nc, int nStreams)
{
	static const int pre32[4] = {4, 20, 12, 28}, pre16[4] = {2, 10, 6, 14};
	static const int ack32[8] = {8, 14, 10, 24, 26, 2, 18, 30}, ack16[8] = {4, 7, 5, 12, 13, 1, 9, 15};
	static const int brk32[8] = {12, 28, 4, 6, 20, 16, 22, 30}, brk16[8] = {6, 14, 2, 3, 10, 8, 11, 15};
	f->M = M, f->Nc = Nc, f->nStreams = nStreams;
	f->nBits = 0;
	for (int t = M; t > 1; t >>= 1) f->nBits++;
	f->tone_hop_step = M == 32 ? 13 : (M == 16 ? 7 : 1);
	int global_offset = (Nc - nStreams * M) / 2;
	if (global_offset < 0) global_offset = 0;
	for (int k = 0; k < nStreams; k++) f->stream_offsets[k] = global_offset + k * M;
	f->preamble_nSymb = 4;
	for (int i = 0; i < 4; i++) f->preamble_tones[i] = M == 32 ? pre32[i] : pre16[i];
	for (int i = 0; i < 8; i++) f->ack_tones[i] = M == 32 ? ack32[i] : ack16[i];
	for (int i = 0; i < 8; i++) f->break_tones[i] = M == 32 ? brk32[i] : brk16[i];
}

/* get_active_nsymb / get_active_nbits: telecom_system.cc:1577-1585. */
static int active_nsymb(const mo_mode *m) { return (m->ctrl_mode && m->ctrl_nsymb > 0) ? m->ctrl_nsymb : m->Nsymb; }
static int active_nbits(const mo_mode *m) { return (m->ctrl_mode && m->ctrl_nBits > 0) ? m->ctrl_nBits : m->nBits; }

/* set_mfsk_ctrl_mode: telecom_system.cc:1572-1575.  Returns the active symbol count. */
int mo_set_mfsk_ctrl_mode(mo_mode *m, int enable)
{
	m->ctrl_mode = enable && m->M == 200 && m->ctrl_nBits > 0 && m->ctrl_nBits < m->nBits;
	return active_nsymb(m);
}

/* ROBUST_0..2 (common_defines.h:63-65, telecom_system.cc:2625-2645,2694-2700,1812-1817): MFSK modes, Nsymb = N / (nBits * nStreams). */
static int mo_mode_init_mfsk(mo_mode *m, int config, int ldpc_iters, const char *ldpc_blob_path)
{
	memset(m, 0, sizeof(*m));
	m->config = config;
	m->M = 200; /* MOD_MFSK */
	m->rate_num = config == 102 ? 4 : 1;
	m->preamble_nSymb = 4;
	m->estimator = 1;
	mfsk_init(&m->mfsk, config == 100 ? 32 : 16, MO_NC, config == 100 ? 1 : 2);
	m->bits_per_symbol = m->mfsk.nBits * m->mfsk.nStreams;
	m->Nsymb = MO_N / m->bits_per_symbol;
	m->ctrl_nBits = config == 100 ? 1200 : (config == 101 ? 1400 : 0); /* telecom_system.cc:2966-2990 */
	m->ctrl_nsymb = m->ctrl_nBits / m->bits_per_symbol;
	m->ctrl_mode = 0;
	m->Nc = MO_NC, m->Nfft = MO_NFFT, m->Ngi = MO_NGI, m->Nofdm = MO_NOFDM;
	m->boost = (double)1.33f;
	m->ls_window = 21;
	m->ldpc_iters = ldpc_iters;
	m->N = MO_N;
	m->K = (int)((float)m->N * ((float)m->rate_num / 16.0f));
	m->P = m->N - m->K;
	m->nData = m->Nsymb, m->nPilots = 0;
	m->nBits = m->nData * m->bits_per_symbol;
	m->nReal = m->nBits - m->P, m->nVirtual = m->N - m->nBits;
	m->frame_bytes = (m->nReal - 16) / 8;
	m->bit_il_block = m->nBits / 10, m->tf_il_block = m->nData / 10;
	build_fft_tables(m);
	mo_srandom(0);
	for (int i = 0; i < m->N; i++) m->scrambler[i] = mo_random() % 2;
	mo_frontend_init(m);
	return load_ldpc(m, ldpc_blob_path);
}

int mo_mode_init(mo_mode *m, int config, int ldpc_iters, const char *ldpc_blob_path)
{
	if (config >= 100 && config <= 102) return mo_mode_init_mfsk(m, config, ldpc_iters, ldpc_blob_path);
	if (config < 0 || config > 16) return -1;
	memset(m, 0, sizeof(*m));
	m->config = config;
	m->M = k_modes[config].M;
	m->rate_num = k_modes[config].rate;
	m->preamble_nSymb = k_modes[config].pre;
	m->estimator = k_modes[config].est;
	m->phase_only = (m->M == 2 || m->M == 4 || m->M == 8);
	m->bits_per_symbol = (int)log2(m->M);
	m->Nsymb = nsymb_of(m->M);
	m->Nc = MO_NC;
	m->Nfft = MO_NFFT;
	m->Ngi = MO_NGI;
	m->Nofdm = MO_NOFDM;
	m->boost = (double)1.33f;	       /* physical_config.cc:46; stored in a float (ofdm.h) */
	m->ls_window = 21;		       /* 20 -> odd 21, telecom_system.cc:2799-2809 */
	m->ldpc_iters = ldpc_iters;	       /* physical_config.cc:74 / main.cc:547-575 */
	m->N = MO_N;
	m->K = (int)((float)m->N * ((float)m->rate_num / 16.0f)); /* ldpc.cc:66 */
	m->P = m->N - m->K;
	build_pilots(m);
	m->nBits = m->nData * m->bits_per_symbol;  /* data_container.cc:90-172 */
	m->nReal = m->nBits - m->P;
	m->nVirtual = m->N - m->nBits;
	m->frame_bytes = (m->nReal - 16) / 8;	   /* telecom_system.cc:332-335 (outer_code_reserved_bits = 16) */
	m->bit_il_block = m->nBits / 10;	   /* telecom_system.cc:2910 */
	m->tf_il_block = m->nData / 10;		   /* telecom_system.cc:2911 */
	build_constellation(m);
	build_fft_tables(m);
	mo_srandom(0);				   /* telecom_system.cc:1961-1966 */
	for (int i = 0; i < m->N; i++) m->scrambler[i] = mo_random() % 2;
	mo_frontend_init(m);
	return load_ldpc(m, ldpc_blob_path);
}

void mo_mode_free(mo_mode *m)
{
	free(m->C);
	free(m->V);
	free(m->Vpos);
	free(m->vdeg);
	free(m->tx_stream);
	m->C = m->V = m->Vpos = m->vdeg = NULL;
	m->tx_stream = NULL;
}

mo_mode *mo_mode_new(int config, int ldpc_iters, const char *ldpc_blob_path)
{
	mo_mode *m = malloc(sizeof(mo_mode));
	if (mo_mode_init(m, config, ldpc_iters, ldpc_blob_path) != 0) {
		free(m);
		return NULL;
	}
	return m;
}

void mo_mode_delete(mo_mode *m)
{
	if (m) {
		mo_mode_free(m);
		free(m);
	}
}

void mo_geometry(const mo_mode *m, int *g)
{
	int v[28] = {m->Nsymb, m->Nc, m->Nfft, m->Ngi, m->Nofdm, m->nData, m->nPilots, m->nBits, m->N, m->K, m->P, m->M,
		     m->preamble_nSymb, m->frame_bytes, m->estimator, m->phase_only, m->bit_il_block, m->tf_il_block,
		     4, m->fe.buffer_Nsymb, (m->Nsymb + m->preamble_nSymb) * m->Nofdm * 4, m->Cwidth, m->Vwidth, 0, m->ldpc_iters, 16, m->ls_window, m->ls_window};
	memcpy(g, v, sizeof(v));
}

void mo_tables(const mo_mode *m, int *carrier_type, double *pilot_seq, int *scrambler, double *constellation, double *boost)
{
	int k = 0;
	for (int c = 0; c < m->Nsymb * m->Nc; c++) {
		if (carrier_type) carrier_type[c] = m->is_pilot[c] ? 1 /*PILOT*/ : 0 /*DATA*/;
		if (m->is_pilot[c]) {
			if (pilot_seq) pilot_seq[k] = m->pilot_val[c];
			k++;
		}
	}
	if (scrambler) memcpy(scrambler, m->scrambler, sizeof(int) * m->N);
	if (constellation)
		for (int i = 0; i < m->M; i++) {
			constellation[2 * i] = creal(m->constellation[i]);
			constellation[2 * i + 1] = cimag(m->constellation[i]);
		}
	if (boost) *boost = m->boost;
}

void mo_ldpc_tables(const mo_mode *m, int *dims, int *C, int *V, int *d, int *Enc)
{
	dims[0] = m->Cwidth;
	dims[1] = m->Vwidth;
	if (C) memcpy(C, m->C, sizeof(int) * m->P * m->Cwidth);
	if (V) memcpy(V, m->V, sizeof(int) * m->N * m->Vwidth);
	int nd = 0;
	for (int i = 0; i < m->N;) { /* QCmatrixd = run-length of variable degrees */
		int j = i;
		while (j < m->N && m->vdeg[j] == m->vdeg[i]) j++;
		if (d) {
			d[nd] = j - i;
			d[nd + 1] = m->vdeg[i];
		}
		nd += 2;
		i = j;
	}
	dims[2] = nd;
	if (Enc) /* QCmatrixEnc[i] = QCmatrixC[i] without the check's own parity bit K+i */
		for (int c = 0; c < m->P; c++) {
			int k = 0;
			for (int j = 0; j < m->Cwidth - 1; j++) Enc[c * (m->Cwidth - 1) + j] = -1;
			for (int j = 0; j < m->Cwidth; j++) {
				int v = m->C[c * m->Cwidth + j];
				if (v != -1 && v != m->K + c) Enc[c * (m->Cwidth - 1) + k++] = v;
			}
		}
}

/* ------------------------------------------------------------------------------------------------
 * FFT: cl_ofdm::_fft_fast / _ifft_fast, source/physical_layer/ofdm.cc:310-377 (iterative radix-2 DIT,
 * bit-reversal table, twiddle table); fft() scales by 1/N (:431-444), ifft() does not (:489-496).
 * ---------------------------------------------------------------------------------------------- */
static void fft_core(const mo_mode *m, double complex *v, int inverse)
{
	int n = MO_NFFT;
	for (int i = 0; i < n; i++)
		if (i < m->bitrev[i]) {
			double complex t = v[i];
			v[i] = v[m->bitrev[i]];
			v[m->bitrev[i]] = t;
		}
	for (int size = 2; size <= n; size *= 2) {
		int half = size / 2, step = n / size;
		for (int i = 0; i < n; i += size)
			for (int j = 0; j < half; j++) {
				double complex w = inverse ? conj(m->twiddle[j * step]) : m->twiddle[j * step];
				double complex t = w * v[i + j + half];
				v[i + j + half] = v[i + j] - t;
				v[i + j] = v[i + j] + t;
			}
	}
}

/* cl_ofdm::symbol_demod = gi_remover + fft + zero_depadder: ofdm.cc:862-867, 423-429, 431-444, 401-411 (start_shift=1). */
static void symbol_demod(const mo_mode *m, const double complex *in, double complex *out)
{
	double complex v[MO_NFFT];
	for (int j = 0; j < MO_NFFT; j++) v[j] = in[j + MO_NGI];
	fft_core(m, v, 0);
	for (int i = 0; i < MO_NFFT; i++) v[i] = v[i] / (double)MO_NFFT;
	for (int j = 0; j < MO_NC / 2; j++) out[j] = v[j + MO_NFFT - MO_NC / 2];
	for (int j = MO_NC / 2; j < MO_NC; j++) out[j] = v[j - MO_NC / 2 + 1];
}

/* cl_ofdm::symbol_mod = zero_padder + ifft + gi_adder: ofdm.cc:855-860, 379-400, 489-496, 412-422. */
static void symbol_mod(const mo_mode *m, const double complex *in, double complex *out)
{
	double complex v[MO_NFFT];
	for (int j = 0; j < MO_NFFT; j++) v[j] = 0;
	for (int j = 0; j < MO_NC / 2; j++) v[j + MO_NFFT - MO_NC / 2] = in[j];
	for (int j = MO_NC / 2; j < MO_NC; j++) v[j - MO_NC / 2 + 1] = in[j];
	fft_core(m, v, 1);
	for (int j = 0; j < MO_NFFT; j++) out[j + MO_NGI] = v[j];
	for (int j = 0; j < MO_NGI; j++) out[j] = v[j + MO_NFFT - MO_NGI];
}

/* interleaver / deinterleaver: source/physical_layer/interleaver.cc:26-109. Index maps only. */
static int il_src(int i, int n, int bs) /* interleaver: out[dst] = in[i]  -> returns dst */
{
	int nb = n / bs;
	if (i >= nb * bs) return i;
	return (i % bs) * nb + i / bs;
}

/* cl_ldpc::encode: source/physical_layer/ldpc.cc:111-132 (IRA accumulate through QCmatrixEnc). */
void mo_ldpc_encode(const mo_mode *m, const int *data, int *enc)
{
	for (int i = 0; i < m->K; i++) enc[i] = data[i];
	for (int i = 0; i < m->P; i++) {
		int b = 0;
		for (int j = 0; j < m->Cwidth; j++) {
			int v = m->C[i * m->Cwidth + j];
			if (v != -1 && v != m->K + i) b ^= enc[v];
		}
		enc[i + m->K] = b;
	}
}

/* TX: transmit_byte/transmit_bit bit chain (telecom_system.cc:342-416) + baseband modulation chain of
 * baseband_test_EsN0 (telecom_system.cc:129-137); psk.mod psk.cc:259-272; framer ofdm.cc:814-835. */
void mo_tx_baseband(const mo_mode *m, const int *payload, int nBytes, double complex *out, int *info_bits, int *codeword,
		    double complex *framed_out)
{
	int bytes[MO_N / 8 + 8], bit[MO_N], scr[MO_N], enc[MO_N], il[MO_N];
	double complex mod[MO_N], tf[MO_N], framed[MO_MAX_CELLS];
	int fs = m->frame_bytes;
	for (int i = 0; i < fs; i++) bytes[i] = i < nBytes ? (payload[i] & 0xFF) : 0;
	for (int i = 0; i < fs; i++) /* byte_to_bit, misc.cc:93-105: LSB first */
		for (int j = 0; j < 8; j++) bit[i * 8 + j] = (bytes[i] >> j) & 1;
	int crc = mo_crc16(bytes, fs);
	for (int j = 0; j < 8; j++) bit[fs * 8 + j] = ((crc & 0xff) >> j) & 1;
	for (int j = 0; j < 8; j++) bit[(fs + 1) * 8 + j] = ((crc >> 8) >> j) & 1;
	for (int i = fs * 8 + 16; i < m->nReal; i++) bit[i] = 0;
	if (info_bits) memcpy(info_bits, bit, sizeof(int) * m->nReal);
	for (int i = 0; i < m->nReal; i++) scr[i] = bit[i] ^ m->scrambler[i]; /* interleaver.cc:111-117 */
	for (int i = 0; i < m->nVirtual; i++) scr[m->nReal + i] = scr[i];
	mo_ldpc_encode(m, scr, enc);
	if (codeword) memcpy(codeword, enc, sizeof(int) * m->N);
	for (int i = 0; i < m->P; i++) enc[m->nReal + i] = enc[i + m->K];
	for (int i = 0; i < m->nBits; i++) il[il_src(i, m->nBits, m->bit_il_block)] = enc[i];
	if (m->M == 200) { /* cl_mfsk::mod, mfsk.cc:254-303: Gray-mapped one-hot tones with hopping, no pilots, no T/F interleaver */
		const mo_mfsk *f = &m->mfsk;
		double complex *fr = calloc((size_t)m->Nsymb * MO_NC, sizeof(double complex));
		double amp = sqrt((double)MO_NC / f->nStreams);
		for (int s = 0; s < active_nsymb(m); s++) /* ctrl mode: only the first ctrl_nBits interleaved bits are sent (:414-416) */
			for (int st = 0; st < f->nStreams; st++) {
				int off = s * m->bits_per_symbol + st * f->nBits, tone = 0;
				for (int bb = 0; bb < f->nBits; bb++)
					if (il[off + bb]) tone |= (1 << (f->nBits - 1 - bb));
				int bin = tone;
				for (int sh = 1; sh < f->nBits; sh++) bin ^= (tone >> sh);
				if (bin >= f->M) bin = f->M - 1;
				fr[s * MO_NC + f->stream_offsets[st] + (bin + s * f->tone_hop_step) % f->M] = amp;
			}
		if (framed_out) memcpy(framed_out, fr, sizeof(double complex) * m->Nsymb * MO_NC);
		for (int s = 0; s < m->Nsymb; s++) symbol_mod(m, fr + s * MO_NC, out + s * m->Nofdm);
		free(fr);
		return;
	}
	int b = m->bits_per_symbol;
	for (int i = 0; i < m->nBits; i += b) {
		unsigned loc = 0;
		for (int j = 0; j < b; j++) loc = (loc << 1) | (unsigned)il[i + j];
		mod[i / b] = m->constellation[loc];
	}
	for (int i = 0; i < m->nData; i++) tf[il_src(i, m->nData, m->tf_il_block)] = mod[i];
	int di = 0;
	for (int c = 0; c < m->Nsymb * m->Nc; c++) framed[c] = m->is_pilot[c] ? m->pilot_val[c] : tf[di++];
	if (framed_out) memcpy(framed_out, framed, sizeof(double complex) * m->Nsymb * m->Nc);
	for (int s = 0; s < m->Nsymb; s++) symbol_mod(m, framed + s * m->Nc, out + s * m->Nofdm);
}

/* interpolate_linear(complex): source/physical_layer/interpolator.cc:34-41. */
static double complex lerp(double complex a, double ax, double complex b, double bx, double x)
{
	return a + (b - a) * (x - ax) / (bx - ax);
}

/* interpolate_linear_col(st_channel_complex*): source/physical_layer/interpolator.cc:163-254. */
static void interp_col(double complex *H, unsigned char *st, int ncol, int nrow, int col)
{
	int ls = 0, le = nrow - 1, nloc = nrow - 1;
	while (nloc > 0) {
		for (int i = ls; i < nrow; i++)
			if (st[i * ncol + col] == 1) {
				ls = i;
				break;
			}
		for (int i = ls + 1; i < nrow; i++)
			if (st[i * ncol + col] == 1) {
				le = i;
				break;
			}
		nloc = le - ls;
		for (int i = ls + 1; i < le; i++) {
			H[i * ncol + col] = lerp(H[ls * ncol + col], ls, H[le * ncol + col], le, i);
			st[i * ncol + col] = 2;
		}
		ls = le;
	}
	ls = 0;
	le = nrow - 1;
	for (int i = 0; i < nrow; i++)
		if (st[i * ncol + col] == 1) {
			ls = i;
			break;
		}
	for (int i = ls + 1; i < nrow; i++)
		if (st[i * ncol + col] == 1) {
			le = i;
			break;
		}
	if (ls != 0)
		for (int i = 0; i < ls; i++) {
			H[i * ncol + col] = lerp(H[ls * ncol + col], ls, H[le * ncol + col], le, i);
			st[i * ncol + col] = 2;
		}
	le = 0;
	ls = nrow - 1;
	for (int i = nrow - 1; i >= 0; i--)
		if (st[i * ncol + col] == 1) {
			le = i;
			break;
		}
	for (int i = le - 1; i >= 0; i--)
		if (st[i * ncol + col] == 1) {
			ls = i;
			break;
		}
	if (le != nrow - 1)
		for (int i = nrow - 1; i > le; i--) {
			H[i * ncol + col] = lerp(H[ls * ncol + col], ls, H[le * ncol + col], le, i);
			st[i * ncol + col] = 2;
		}
}

/* get_angle: source/physical_layer/misc.cc:34-56 (pi/2 whenever the real part is exactly 0). */
static double get_angle(double complex v)
{
	double re = creal(v), im = cimag(v);
	if (re == 0) return M_PI / 2;
	if (re > 0) return atan(im / re);
	if (im >= 0) return atan(im / re) + M_PI;
	return atan(im / re) - M_PI;
}

/* decode_SPA: source/physical_layer/ldpc_decoder_SPA.cc:25-218 (flooding tanh/atanh BP in double). */
int mo_ldpc_decode(const mo_mode *m, const float *LLRi, int *LLRo)
{
	int N = m->N, P = m->P, K = m->K, Cw = m->Cwidth, Vw = m->Vwidth;
	static _Thread_local double R[MO_N * 16], Q[MO_N * 16], LLRtmp[MO_N];
	static _Thread_local int LLRbin[MO_N];
	int iteration = 0, nOnes = 0;
	for (int i = 0; i < N; i++) {
		for (int j = 0; j < Vw; j++) R[i * Vw + j] = Q[i * Vw + j] = 0;
		LLRbin[i] = LLRi[i] < 0;
		LLRtmp[i] = LLRi[i];
	}
	for (int i = 0; i < P; i++) { /* :62-76 initial syndrome */
		int c = LLRbin[m->C[i * Cw]];
		for (int j = 1; j < Cw; j++)
			if (m->C[i * Cw + j] != -1) c ^= LLRbin[m->C[i * Cw + j]];
		nOnes += c;
	}
	if (nOnes != 0) {
		for (int i = 0; i < N; i++) /* :106-122 Q init over the first deg(v) slots */
			for (int j = 0; j < m->vdeg[i]; j++) Q[i * Vw + j] = LLRi[i];
		for (iteration = 1; iteration <= m->ldpc_iters; iteration++) {
			for (int ci = 0; ci < P; ci++) /* :129-160 check update */
				for (int cj = 0; cj < Cw; cj++) {
					int j = m->C[ci * Cw + cj];
					if (j == -1) continue;
					double temp = 1;
					for (int k = 0; k < Cw; k++) {
						int i1 = m->C[ci * Cw + k];
						if (i1 != j && i1 != -1) temp *= tanh(0.5 * Q[i1 * Vw + m->Vpos[ci * Cw + k]]);
					}
					if (temp == 1) temp = 0.9999999;
					if (temp == -1) temp = -0.9999999;
					R[j * Vw + m->Vpos[ci * Cw + cj]] = 2 * atanh(temp);
				}
			for (int i = 0; i < N; i++) { /* :162-170 posterior */
				LLRtmp[i] = LLRi[i];
				for (int j = 0; j < Vw; j++) LLRtmp[i] += R[i * Vw + j];
				LLRbin[i] = LLRtmp[i] < 0;
			}
			nOnes = 0; /* :173-190 syndrome + early exit */
			for (int i = 0; i < P; i++) {
				int c = LLRbin[m->C[i * Cw]];
				for (int j = 1; j < Cw; j++)
					if (m->C[i * Cw + j] != -1) c ^= LLRbin[m->C[i * Cw + j]];
				nOnes += c;
			}
			if (nOnes == 0) break;
			for (int i = 0; i < N; i++) /* :193-209 */
				for (int j = 0; j < m->vdeg[i]; j++) Q[i * Vw + j] = LLRtmp[i] - R[i * Vw + j];
		}
	}
	for (int i = 0; i < K; i++) LLRo[i] = LLRtmp[i] < 0;
	return iteration;
}

/* The RX tail: source/physical_layer/telecom_system.cc:1132-1341 (+ success bookkeeping :1343-1375). */
/* cl_mfsk::demod: mfsk.cc:305-390 (per symbol: noise variance from the carriers outside the tone bands, per stream the hop-reversed
 * tone energies, max-log LLR over the Gray-mapped tones scaled by 1/(2 sigma^2), clamped to +-5). */
static void mfsk_demod(const mo_mfsk *f, const double complex *fft_in, int total_bits, float *llr_out)
{
	int bps = f->nBits * f->nStreams, nSymbols = total_bits / bps;
	for (int s = 0; s < nSymbols; s++) {
		int band_start = f->stream_offsets[0], band_end = f->stream_offsets[f->nStreams - 1] + f->M;
		double noise_sum = 0.0;
		int noise_bins = 0;
		for (int k = 0; k < f->Nc; k++)
			if (k < band_start || k >= band_end) {
				double complex v = fft_in[s * f->Nc + k];
				double e = creal(v) * creal(v) + cimag(v) * cimag(v);
				if (isfinite(e)) {
					noise_sum += e;
					noise_bins++;
				}
			}
		double noise_var = (noise_bins > 0) ? noise_sum / noise_bins : 1e-30;
		if (noise_var < 1e-30) noise_var = 1e-30;
		double llr_scale = 1.0 / (2.0 * noise_var);
		for (int st = 0; st < f->nStreams; st++) {
			double E_raw[64], E[64];
			for (int q = 0; q < f->M; q++) {
				double complex v = fft_in[s * f->Nc + f->stream_offsets[st] + q];
				E_raw[q] = creal(v) * creal(v) + cimag(v) * cimag(v);
				if (!isfinite(E_raw[q])) E_raw[q] = 0.0;
			}
			int hop = (s * f->tone_hop_step) % f->M;
			for (int q = 0; q < f->M; q++) E[q] = E_raw[(q + hop) % f->M];
			int off = s * bps + st * f->nBits;
			for (int k = 0; k < f->nBits; k++) {
				int mask = 1 << (f->nBits - 1 - k);
				double max_E1 = -1e30, max_E0 = -1e30;
				for (int q = 0; q < f->M; q++) {
					int gray = q ^ (q >> 1);
					if (gray & mask) {
						if (E[q] > max_E1) max_E1 = E[q];
					} else if (E[q] > max_E0)
						max_E0 = E[q];
				}
				double llr = (max_E0 - max_E1) * llr_scale;
				if (!isfinite(llr)) llr = 0.0;
				else if (llr > 5.0) llr = 5.0;
				else if (llr < -5.0) llr = -5.0;
				llr_out[off + k] = (float)llr;
			}
		}
	}
}

/* The MFSK branch of the RX tail: telecom_system.cc:1132-1198 then the common :1296-1367 (no AGC, no channel estimate, no gate; SNR 0). */
static void mo_rx_tail_mfsk(const mo_mode *m, const double complex *bb, mo_rx_out *o)
{
	int S = m->Nsymb, cells = S * MO_NC;
	double complex *Y = malloc(sizeof(double complex) * cells);
	float llr[MO_N], llr_cw[MO_N + 8];
	int bits[MO_N], bytes[MO_N / 8 + 1];
	for (int c = 0; c < cells; c++) Y[c] = 0;
	for (int s = 0; s < active_nsymb(m); s++) symbol_demod(m, bb + (size_t)s * m->Nofdm, Y + s * MO_NC);
	if (o->Y) memcpy(o->Y, Y, sizeof(double complex) * cells);
	mfsk_demod(&m->mfsk, Y, active_nbits(m), llr);
	for (int i = active_nbits(m); i < m->nBits; i++) llr[i] = 0.0f; /* :1188-1197: punctured positions are erasures */
	free(Y);
	if (o->llr_demod) memcpy(o->llr_demod, llr, sizeof(float) * m->nBits);
	int bs = m->bit_il_block, nb = m->nBits / bs;
	for (int i = 0; i < nb; i++)
		for (int j = 0; j < bs; j++) llr_cw[i * bs + j] = llr[j * nb + i];
	for (int i = nb * bs; i < m->nBits; i++) llr_cw[i] = llr[i];
	for (int i = m->P - 1; i >= 0; i--) llr_cw[i + m->nReal + m->nVirtual] = llr_cw[i + m->nReal];
	for (int i = 0; i < m->nVirtual; i++) llr_cw[m->nReal + i] = llr_cw[i];
	if (o->llr_cw) memcpy(o->llr_cw, llr_cw, sizeof(float) * m->N);
	int iterations = mo_ldpc_decode(m, llr_cw, bits);
	if (o->bits) memcpy(o->bits, bits, sizeof(int) * m->K);
	for (int i = 0; i < m->nReal; i++) bits[i] ^= m->scrambler[i];
	for (int i = 0; i < m->nReal / 8; i++) {
		bytes[i] = 0;
		for (int j = 0; j < 8; j++) bytes[i] |= bits[i * 8 + j] << j;
	}
	int all_zeros = 1;
	for (int i = 0; i < m->nReal / 8; i++)
		if (bytes[i] != 0) {
			all_zeros = 0;
			break;
		}
	if (o->bytes) memcpy(o->bytes, bytes, sizeof(int) * (m->nReal / 8));
	if (o->payload) memcpy(o->payload, bytes, sizeof(int) * m->frame_bytes);
	int crc = all_zeros ? 0 : mo_crc16(bytes, m->nReal / 8);
	int decoded = !(all_zeros || crc != 0);
	if (o->stats) {
		o->stats[0] = iterations, o->stats[1] = crc, o->stats[2] = all_zeros, o->stats[3] = decoded;
		o->stats[4] = decoded ? 0.0 : -99.9, o->stats[5] = 0, o->stats[6] = 0, o->stats[7] = 1.0;
	}
}

void mo_rx_tail(const mo_mode *m, const double complex *bb, mo_rx_out *o)
{
	if (m->M == 200) {
		mo_rx_tail_mfsk(m, bb, o);
		return;
	}
	int S = m->Nsymb, C = m->Nc, cells = S * C;
	double complex Y[MO_MAX_CELLS], H[MO_MAX_CELLS], Hna[MO_MAX_CELLS], Z[MO_MAX_CELLS], Zna[MO_MAX_CELLS];
	unsigned char st[MO_MAX_CELLS];
	double complex defr[MO_N], tfd[MO_N];
	float llr[MO_N], llr_cw[MO_N + 8];
	int bits[MO_N], bytes[MO_N / 8 + 1];

	for (int s = 0; s < S; s++) symbol_demod(m, bb + (size_t)s * m->Nofdm, Y + s * C); /* :1135-1138 */

	/* automatic_gain_control: ofdm.cc:1467-1498; get_amplitude misc.cc:58-63 */
	double amp = 0;
	int np = 0;
	for (int c = 0; c < cells; c++)
		if (m->is_pilot[c]) {
			amp += sqrt(creal(Y[c]) * creal(Y[c]) + cimag(Y[c]) * cimag(Y[c]));
			np++;
		}
	amp /= np;
	double agc = m->boost / amp;
	for (int c = 0; c < cells; c++) Y[c] *= agc;
	if (o->Y) memcpy(o->Y, Y, sizeof(double complex) * cells);

	for (int c = 0; c < cells; c++) {
		H[c] = 0;
		st[c] = 0;
	}
	if (m->estimator == 0) { /* ZF_channel_estimator: ofdm.cc:1266-1285 */
		for (int c = 0; c < cells; c++)
			if (m->is_pilot[c]) {
				H[c] = Y[c] / (double complex)(m->pilot_val[c] + 0.0 * I);
				st[c] = 1;
			}
	} else { /* LS_channel_estimator: ofdm.cc:1315-1422; matrix_multiplication misc.cc:73-91 */
		int hw = m->ls_window / 2;
		for (int j = 0; j < C; j++)
			for (int i = 0; i < S; i++) {
				if (!m->is_pilot[i * C + j]) continue;
				double complex x[512], y[512];
				int n = 0;
				for (int k = i - hw; k <= i + hw; k++) {
					if (k < 0 || k >= S) continue;
					for (int l = j - hw; l <= j + hw; l++) {
						if (l < 0 || l >= C) continue;
						if (m->is_pilot[k * C + l]) {
							x[n] = m->pilot_val[k * C + l];
							y[n] = Y[k * C + l];
							n++;
						}
					}
				}
				double complex ch = 0;
				for (int k = 0; k < n; k++) ch += x[k] * x[k];
				ch = 1.0 / ch;
				for (int k = 0; k < n; k++) x[k] *= ch;
				double complex acc = 0;
				for (int k = 0; k < n; k++) acc += x[k] * y[k];
				H[i * C + j] = acc;
				st[i * C + j] = 1;
			}
	}
	for (int j = 0; j < C; j++) interp_col(H, st, C, S, j); /* Dx = 1: every column (ofdm.cc:1287-1297,1425-1435) */

	double hsum = 0;
	int hm = 0; /* mean |H| at pilots, telecom_system.cc:1225-1244 */
	for (int c = 0; c < cells; c++)
		if (st[c] == 1) {
			hsum += cabs(H[c]);
			hm++;
		}
	double mean_H = hm ? hsum / hm : -1.0;

	if (m->phase_only) { /* restore_channel_amplitude: ofdm.cc:1453-1466; set_complex misc.cc:65-71 */
		for (int c = 0; c < cells; c++) {
			Hna[c] = H[c];
			double th = get_angle(H[c]);
			H[c] = 1 * cos(th) + (1 * sin(th)) * I;
		}
		for (int c = 0; c < cells; c++) Zna[c] = Y[c] / Hna[c]; /* ofdm.cc:1648-1657 */
	}
	if (o->H) memcpy(o->H, H, sizeof(double complex) * cells);
	for (int c = 0; c < cells; c++) Z[c] = Y[c] / H[c]; /* channel_equalizer: ofdm.cc:1637-1647 */
	if (o->Z) memcpy(o->Z, Z, sizeof(double complex) * cells);

	/* measure_variance: ofdm.cc:1500-1521 -> float (telecom_system.cc:649,1291) */
	double var = 0;
	np = 0;
	for (int c = 0; c < cells; c++)
		if (m->is_pilot[c]) {
			double complex d = Z[c] - m->pilot_val[c];
			var += creal(d) * creal(d) + cimag(d) * cimag(d);
			np++;
		}
	var /= (double)np;
	float variance = (float)var;

	int di = 0; /* deframer: ofdm.cc:837-852 */
	for (int c = 0; c < cells; c++)
		if (!m->is_pilot[c]) defr[di++] = Z[c];
	{ /* deinterleaver(complex): interleaver.cc:94-109 */
		int bs = m->tf_il_block, nb = m->nData / bs;
		for (int i = 0; i < nb; i++)
			for (int j = 0; j < bs; j++) tfd[i * bs + j] = defr[j * nb + i];
		for (int i = nb * bs; i < m->nData; i++) tfd[i] = defr[i];
	}
	{ /* cl_psk::demod: psk.cc:278-326 */
		int b = m->bits_per_symbol;
		float D[64], L[8];
		for (int i = 0; i < m->nBits; i += b) {
			double complex z = tfd[i / b];
			for (int j = 0; j < m->M; j++) {
				double dr = creal(z) - creal(m->constellation[j]), dq = cimag(z) - cimag(m->constellation[j]);
				D[j] = dr * dr + dq * dq;
			}
			unsigned mask = 1;
			for (int k = 0; k < b; k++) {
				float d0 = D[0], d1 = D[mask];
				for (int j = 0; j < m->M; j++) {
					if ((j & mask) == 0) {
						if (D[j] < d0) d0 = D[j];
					} else if (D[j] < d1)
						d1 = D[j];
				}
				L[k] = ((1 / variance) * (d1 - d0));
				mask <<= 1;
			}
			for (int j = 0; j < b; j++) llr[i + j] = L[b - j - 1];
		}
	}
	if (o->llr_demod) memcpy(o->llr_demod, llr, sizeof(float) * m->nBits);
	{ /* deinterleaver(float): interleaver.cc:77-92, then LLR expand telecom_system.cc:1300-1308 */
		int bs = m->bit_il_block, nb = m->nBits / bs;
		for (int i = 0; i < nb; i++)
			for (int j = 0; j < bs; j++) llr_cw[i * bs + j] = llr[j * nb + i];
		for (int i = nb * bs; i < m->nBits; i++) llr_cw[i] = llr[i];
		for (int i = m->P - 1; i >= 0; i--) llr_cw[i + m->nReal + m->nVirtual] = llr_cw[i + m->nReal];
		for (int i = 0; i < m->nVirtual; i++) llr_cw[m->nReal + i] = llr_cw[i];
	}
	if (o->llr_cw) memcpy(o->llr_cw, llr_cw, sizeof(float) * m->N);

	int iterations = mo_ldpc_decode(m, llr_cw, bits); /* :1310 */
	if (o->bits) memcpy(o->bits, bits, sizeof(int) * m->K);
	for (int i = 0; i < m->nReal; i++) bits[i] ^= m->scrambler[i]; /* :1313 */
	for (int i = 0; i < m->nReal / 8; i++) {		       /* bit_to_byte misc.cc:107-130 */
		bytes[i] = 0;
		for (int j = 0; j < 8; j++) bytes[i] |= bits[i * 8 + j] << j;
	}
	int all_zeros = 1; /* :1319-1327 */
	for (int i = 0; i < m->nReal / 8; i++)
		if (bytes[i] != 0) {
			all_zeros = 0;
			break;
		}
	if (o->bytes) memcpy(o->bytes, bytes, sizeof(int) * (m->nReal / 8));
	if (o->payload) memcpy(o->payload, bytes, sizeof(int) * m->frame_bytes);
	int crc = 0;
	if (!all_zeros) crc = mo_crc16(bytes, m->nReal / 8); /* :1337-1341 */
	int decoded = !(all_zeros || crc != 0);		     /* :1343-1349 */
	double snr = -99.9;
	if (decoded) {
		if (m->estimator == 1) { /* :1368-1375 */
			float v = variance;
			if (m->phase_only) {
				double vv = 0;
				int n2 = 0;
				for (int c = 0; c < cells; c++)
					if (m->is_pilot[c]) {
						double complex d = Zna[c] - m->pilot_val[c];
						vv += creal(d) * creal(d) + cimag(d) * cimag(d);
						n2++;
					}
				v = (float)(vv / (double)n2);
			}
			snr = 10.0 * log10(1.0 / v);
		} else {
			/* ZF: re-encode the decoded frame and measure the distance of the equalised data symbols to it
			 * (telecom_system.cc:1376-1400; measure_SNR ofdm.cc:1622-1635) */
			int hd[MO_N], enc[MO_N], il[MO_N];
			double complex mod[MO_N], tf[MO_N];
			for (int i = 0; i < m->nReal; i++) hd[i] = bits[i] ^ m->scrambler[i]; /* scramble back: bit_energy_dispersal is an XOR */
			for (int i = 0; i < m->nVirtual; i++) hd[m->nReal + i] = hd[i];
			mo_ldpc_encode(m, hd, enc);
			for (int i = 0; i < m->P; i++) enc[m->nReal + i] = enc[i + m->K];
			for (int i = 0; i < m->nBits; i++) il[il_src(i, m->nBits, m->bit_il_block)] = enc[i];
			int b = m->bits_per_symbol;
			for (int i = 0; i < m->nBits; i += b) {
				unsigned loc = 0;
				for (int j = 0; j < b; j++) loc = (loc << 1) | (unsigned)il[i + j];
				mod[i / b] = m->constellation[loc];
			}
			for (int i = 0; i < m->nData; i++) tf[il_src(i, m->nData, m->tf_il_block)] = mod[i];
			double complex *eq = Z;
			if (m->phase_only) eq = Zna;
			double vv = 0;
			int di2 = 0;
			for (int c = 0; c < cells; c++)
				if (!m->is_pilot[c]) {
					double complex d = tf[di2++] - eq[c];
					vv += creal(d) * creal(d) + cimag(d) * cimag(d);
				}
			vv /= (double)m->nData;
			snr = -10.0 * log10(vv);
		}
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

double mo_rx_tail_timed(const mo_mode *m, const double complex *bb, int n_frames, int *payloads, int *decoded, int *iterations)
{
	struct timespec t0, t1;
	double st[8];
	size_t stride = (size_t)m->Nsymb * m->Nofdm;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int f = 0; f < n_frames; f++) {
		mo_rx_out o;
		memset(&o, 0, sizeof(o));
		o.stats = st;
		o.payload = payloads ? payloads + (size_t)f * m->frame_bytes : NULL;
		mo_rx_tail(m, bb + f * stride, &o);
		if (decoded) decoded[f] = (int)st[3];
		if (iterations) iterations[f] = (int)st[0];
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ================================================================================================
 * RX FRONT-END: everything receive_byte() does before the tail, and the trial loop
 * around it.  source/physical_layer/telecom_system.cc:646-1131 and 1343-1518, OFDM branch only
 * (M != MOD_MFSK, mfsk_fixed_delay < 0, g_gui_state.coarse_freq_sync_enabled == false: the compiled
 * defaults, gui_state.h:143).  Operation order follows the reference so the results are bit-identical
 * to oracle/_ref compiled by the same gcc (tests/test_oracle_frontend.py).
 * ============================================================================================== */

/* cl_FIR::design, LPF + HAMMING: source/physical_layer/fir_filter.cc:45-163. */
static void fir_design_lpf_hamming(double fcut, double tbw, double fs, int *ntaps_out, double *c)
{
	int n = (int)(4.0 / (tbw / (fs / 2.0)));
	if (n % 2 == 0) n++;
	double Ts = 1.0 / (fs);
	double temp;
	c[n / 2] = 1;
	for (int i = 0; i < n / 2; i++) {
		temp = 2 * M_PI * fcut * (double)(n / 2 - i) * Ts;
		c[i] = sin(temp) / temp;
		c[n - i - 1] = c[i];
	}
	temp = 0;
	for (int i = 0; i < n; i++) temp += c[i];
	for (int i = 0; i < n; i++) c[i] /= temp;
	for (int i = 0; i < n; i++) c[i] *= 0.54 - 0.46 * cos(2.0 * M_PI * (double)i / (n - 1));
	*ntaps_out = n;
}

/* Front-end constants: physical_config.cc:59,79-101; telecom_system.cc:69 (carrier_amplitude), data_container.cc:133-143 (buffer_Nsymb). */
void mo_frontend_init(mo_mode *m)
{
	mo_frontend *f = &m->fe;
	f->interp = 4;
	f->fs = 48000.0;
	f->bandwidth = 48000.0 * 50.0 / 256 / 4;
	f->fc = 0.0 + (f->bandwidth / 2 + 300);
	f->amp = sqrt(2.0);
	f->trials_max = 2;
	f->use_last_time = 1;
	f->use_last_freq = 1;
	f->ignore_limit = (double)0.1f; /* stored in a float (ofdm.h) */
	fir_design_lpf_hamming(0.9 * f->bandwidth / 2, 3000, f->fs, &f->ntaps_ts, f->c_ts);
	fir_design_lpf_hamming(1.0 * f->bandwidth / 2, 3000, f->fs, &f->ntaps_data, f->c_data);
	double sym_time_ms = 1000.0 * m->Nofdm * f->interp / 48000.0;
	int turnaround_symb = (int)ceil(1200.0 / sym_time_ms) + 4;
	int frame_symb = m->preamble_nSymb + m->Nsymb;
	int min_buf = frame_symb * 2;
	if (frame_symb + turnaround_symb > min_buf) min_buf = frame_symb + turnaround_symb;
	if (min_buf < 32) min_buf = 32;
	f->buffer_Nsymb = min_buf;
}

void mo_frontend_tables(const mo_mode *m, int *ntaps, double *ts_coef, double *data_coef, double *consts)
{
	const mo_frontend *f = &m->fe;
	ntaps[0] = f->ntaps_ts;
	ntaps[1] = f->ntaps_data;
	memcpy(ts_coef, f->c_ts, sizeof(double) * f->ntaps_ts);
	memcpy(data_coef, f->c_data, sizeof(double) * f->ntaps_data);
	double v[8] = {f->fs, f->fc, f->amp, f->bandwidth, f->trials_max, f->use_last_time, f->use_last_freq, f->ignore_limit};
	memcpy(consts, v, sizeof(v));
}

/* cl_FIR::apply(complex): fir_filter.cc:164-187 (zero-phase: output delayed by (nTaps-1)/2 is dropped). */
static void fir_apply(const double *c, int nt, const double complex *in, double complex *out, int n)
{
	for (int i = 0; i < n + nt - 1; i++) {
		double ar = 0, ai = 0;
		for (int j = 0; j < nt; j++)
			if ((i - j) >= 0 && (i - j) < n) {
				ar += creal(in[i - j]) * c[j];
				ai += cimag(in[i - j]) * c[j];
			}
		if (i >= (nt - 1) / 2 && i < n + (nt - 1) / 2) out[i - (nt - 1) / 2] = ar + ai * I;
	}
}

/* cl_ofdm::passband_to_baseband with decimation_rate 1: ofdm.cc:2316-2339. */
static void p2b(const mo_mode *m, const double *in, int n, double complex *out, double fc, int data_filter, double complex *scratch)
{
	const mo_frontend *f = &m->fe;
	double Ts = 1.0 / f->fs;
	for (int i = 0; i < n; i++) {
		double re = in[i] * f->amp * cos(2 * M_PI * fc * (double)i * Ts);
		double im = in[i] * f->amp * sin(2 * M_PI * fc * (double)i * Ts);
		scratch[i] = re + im * I;
	}
	if (data_filter) fir_apply(f->c_data, f->ntaps_data, scratch, out, n);
	else fir_apply(f->c_ts, f->ntaps_ts, scratch, out, n);
}

/* cl_ofdm::time_sync_preamble_with_metric / time_sync_preamble: ofdm.cc:1846-1967 / 1735-1845 (Schmidl-Cox self-correlation of
 * GI vs symbol tail and of the two symbol halves over the preamble, then the reference's partial selection "sort"). */
static int time_sync(const mo_mode *m, const double complex *in, int size, int rate, int location_to_return, int step,
		     int nTrials_max, double *corr_out, int *loc, double *vals)
{
	int pre = m->preamble_nSymb, S = (MO_NGI + MO_NFFT) * rate;
	for (int i = 0; i < size; i++) {
		loc[i] = -1;
		vals[i] = 0;
	}
	for (int i = 0; i < size - pre * S; i += step) {
		const double complex *data = in + i;
		double cc = 0, na = 0, nb = 0;
		for (int l = 0; l < pre; l++) {
			const double complex *a = data + l * S, *b = data + l * S + MO_NFFT * rate;
			for (int q = 0; q < MO_NGI * rate; q++) {
				cc += creal(a[q]) * creal(b[q]);
				na += creal(a[q]) * creal(a[q]);
				nb += creal(b[q]) * creal(b[q]);
				cc += cimag(a[q]) * cimag(b[q]);
				na += cimag(a[q]) * cimag(a[q]);
				nb += cimag(b[q]) * cimag(b[q]);
			}
			a = data + l * S + MO_NGI * rate;
			b = data + l * S + (MO_NGI + MO_NFFT / 2) * rate;
			for (int q = 0; q < (MO_NFFT / 2) * rate; q++) {
				cc += creal(a[q]) * creal(b[q]);
				na += creal(a[q]) * creal(a[q]);
				nb += creal(b[q]) * creal(b[q]);
				cc += cimag(a[q]) * cimag(b[q]);
				na += cimag(a[q]) * cimag(a[q]);
				nb += cimag(b[q]) * cimag(b[q]);
			}
		}
		if (na < 0.001 || nb < 0.001) cc = 0.0;
		else cc = cc / sqrt(na * nb);
		vals[i] = cc;
		loc[i] = i;
	}
	if (location_to_return >= nTrials_max) location_to_return = nTrials_max - 1;
	for (int j = 0; j < nTrials_max; j++) {
		loc[j] = j;
		for (int i = j + 1; i < size; i++)
			if (vals[i] > vals[j]) {
				vals[j] = vals[i];
				loc[j] = i;
			}
	}
	if (corr_out) *corr_out = vals[location_to_return];
	return loc[location_to_return];
}

/* mean energy of one passband-rate symbol starting at `pos` (the gates of telecom_system.cc:741-753,768-776,818-826,...) */
static double sym_energy(const double complex *bbi, int pos, int sym_samples, int buf_samples)
{
	double e = 0;
	int cnt = 0;
	for (int i = 0; i < sym_samples && (pos + i) < buf_samples; i++) {
		double re = creal(bbi[pos + i]), im = cimag(bbi[pos + i]);
		e += re * re + im * im;
		cnt++;
	}
	return (cnt > 0) ? e / cnt : 0.0;
}

/* cl_ofdm::carrier_sampling_frequency_sync (Moose): ofdm.cc:540-595. */
static double moose(const mo_mode *m, const double complex *in, double carrier_width, int pre)
{
	double complex frame[MO_NFFT], d1[MO_NC], d2[MO_NC], mul = 0;
	pre = (pre / 2 == 0) ? 1 : pre / 2;
	for (int j = 0; j < pre; j++) {
		for (int h = 0; h < 2; h++) {
			for (int i = 0; i < MO_NFFT / 2; i++) {
				frame[i] = in[j * MO_NOFDM + i + h * MO_NFFT / 2];
				frame[i + MO_NFFT / 2] = in[j * MO_NOFDM + i + h * MO_NFFT / 2];
			}
			fft_core(m, frame, 0);
			for (int i = 0; i < MO_NFFT; i++) frame[i] = frame[i] / (double)MO_NFFT;
			double complex *d = h ? d2 : d1;
			for (int i = 0; i < MO_NC / 2; i++) d[i] = frame[i + MO_NFFT - MO_NC / 2];
			for (int i = MO_NC / 2; i < MO_NC; i++) d[i] = frame[i - MO_NC / 2 + 1];
		}
		for (int i = 0; i < MO_NC; i++) mul += conj(d2[i]) * d1[i];
	}
	return (get_angle(mul) / M_PI) * carrier_width;
}

/*
 * cl_telecom_system::receive_byte, OFDM branch: telecom_system.cc:646-1518.
 *   passband : Nofdm*buffer_Nsymb*4 doubles;  out : frame_bytes ints
 *   stats[12]: iterations, crc, all_zeros, decoded, SNR, delay, sync_trials, freq_offset, coarse_metric, signal_stregth_dbm,
 *              buffer samples, frame_bytes;  state[2] in/out: delay_of_last_decoded_message, freq_offset_of_last_decoded_message
 *   baseband_out (optional): the (pre+Nsymb)*272 post-synchronisation samples the last trial's tail consumed.
 */
int mo_time_sync_mfsk(const mo_mode *m, const double complex *bbi, int n, int search_start_symb);

/* The MFSK branch of receive_byte() (ctrl mode off): telecom_system.cc:646-716, 928-943, 1020-1031, 1081-1198,
 * 1296-1367.  One trial: tone-preamble sync on the time-sync base-band (symbol grid), frame-completeness check, data-filter mix +
 * decimation at that delay, no frequency correction, the MFSK tail.  state[2] in: first symbol of the preamble search
 * (receive_stats.mfsk_search_raw - nUnder_processing_events); state[3] out: frame_overflow_symbols; state[4] in/out: mfsk_fixed_delay
 * (>= 0 bypasses the search once, -1 afterwards; only this branch honours it: the reference sets it in MFSK configurations only). */
static void mo_receive_byte_mfsk(const mo_mode *m, const double *passband, int *out, double *stats, double *state, double complex *baseband_out)
{
	const mo_frontend *f = &m->fe;
	int rate = f->interp, sym = m->Nofdm * rate, buf = m->Nofdm * f->buffer_Nsymb * rate, pre = m->preamble_nSymb, S = m->Nsymb;
	int frame_dec = m->Nofdm * (S + pre);
	double complex *bbi = malloc(sizeof(double complex) * buf), *scratch = malloc(sizeof(double complex) * buf);
	double complex *bb = calloc(frame_dec, sizeof(double complex));
	double tail_stats[8] = {0};
	int payload[MO_N / 8];
	mo_rx_out ro;
	memset(&ro, 0, sizeof(ro));
	ro.payload = payload, ro.stats = tail_stats;
	int last_delay = (int)state[0], message_decoded = 0, sync_trials = 0, iterations = 0, crc = 0, all_zeros = 0, overflow = 0;
	double SNR = 0;
	double signal_dbm;
	int delay, fixed_delay = (int)state[4];
	if (fixed_delay >= 0) { /* :663-673: known delay (the ARQ layer's overflow recapture, arq_common.cc:2830-2833) -- no mix, no search, used once */
		delay = fixed_delay;
		state[4] = -1;
		signal_dbm = 0;
	} else {
		p2b(m, passband, buf, bbi, f->fc, 0, scratch);
		double ss = 0;
		for (int i = 0; i < buf; i++) ss += pow(creal(bbi[i]), 2) + pow(cimag(bbi[i]), 2);
		ss /= buf;
		signal_dbm = 10.0 * log10(ss / 0.001);
		int search_start = (int)state[2];
		if (search_start < 0) search_start = 0;
		delay = mo_time_sync_mfsk(m, bbi, buf, search_start); /* :686 */
	}
	int pream_symb_loc = delay / sym;
	if (pream_symb_loc < 1) pream_symb_loc = 1;
	int frame_end = delay + (pre + active_nsymb(m)) * sym; /* :702-715 */
	if (frame_end > buf) {
		overflow = (frame_end - buf + sym - 1) / sym;
	} else {
		int lower = pre, upper = f->buffer_Nsymb - (S + pre);
		if (pream_symb_loc > lower && pream_symb_loc < upper) {
			if (delay < 0) delay = 0;
			int max_delay = buf - frame_dec * rate;
			if (delay > max_delay) delay = max_delay;
			p2b(m, passband, buf, bbi, f->fc, 1, scratch);
			for (int i = 0, k = 0; i < frame_dec * rate; i += rate) bb[k++] = bbi[delay + i];
			mo_rx_tail(m, bb + pre * m->Nofdm, &ro);
			iterations = (int)tail_stats[0], crc = (int)tail_stats[1], all_zeros = (int)tail_stats[2];
			for (int i = 0; i < m->frame_bytes; i++) out[i] = payload[i];
			if (!(int)tail_stats[3]) {
				SNR = -99.9;
				sync_trials++;
			} else {
				SNR = 0.0;
				message_decoded = 1;
				last_delay = delay;
			}
		}
	}
	double v[12] = {iterations, crc, all_zeros, message_decoded, SNR, delay, sync_trials, 0, 0, signal_dbm, buf, m->frame_bytes};
	memcpy(stats, v, sizeof(v));
	state[0] = last_delay;
	state[3] = overflow;
	if (baseband_out) memcpy(baseband_out, bb, sizeof(double complex) * frame_dec);
	free(bbi), free(scratch), free(bb);
}

void mo_receive_byte(const mo_mode *m, const double *passband, int *out, double *stats, double *state, double complex *baseband_out)
{
	if (m->M == 200) {
		mo_receive_byte_mfsk(m, passband, out, stats, state, baseband_out);
		return;
	}
	const mo_frontend *f = &m->fe;
	int rate = f->interp, sym = m->Nofdm * rate, buf = m->Nofdm * f->buffer_Nsymb * rate;
	int pre = m->preamble_nSymb, S = m->Nsymb;
	int frame_dec = m->Nofdm * (S + pre);
	double complex *bbi = malloc(sizeof(double complex) * buf), *scratch = malloc(sizeof(double complex) * buf);
	double complex *bb = malloc(sizeof(double complex) * frame_dec);
	int *loc = malloc(sizeof(int) * buf);
	double *vals = malloc(sizeof(double) * buf);
	double tail_stats[8] = {0};
	int payload[MO_N / 8];
	mo_rx_out ro;
	memset(&ro, 0, sizeof(ro));
	ro.payload = payload;
	ro.stats = tail_stats;

	int last_delay = (int)state[0];
	double last_freq = state[1];
	int message_decoded = 0, sync_trials = 0, iterations = 0, crc = 0, all_zeros = 0;
	double SNR = 0, freq_offset = 0, freq_offset_measured = 0, coarse_metric = 0;
	int step = 100, delay, pream_symb_loc;
	double coarse_freq_offset = 0.0;

	p2b(m, passband, buf, bbi, f->fc, 0, scratch); /* :676 */
	double ss = 0;				       /* measure_signal_stregth: ofdm.cc:1523-1539 */
	for (int i = 0; i < buf; i++) ss += pow(creal(bbi[i]), 2) + pow(cimag(bbi[i]), 2);
	ss /= buf;
	double signal_dbm = 10.0 * log10(ss / 0.001);

	delay = time_sync(m, bbi, buf, rate, 0, step, 1, &coarse_metric, loc, vals); /* :691-693 */
	pream_symb_loc = delay / sym;
	if (pream_symb_loc < 1) pream_symb_loc = 1;
	int lower_bound = pre, upper_bound = f->buffer_Nsymb - (S + pre);

	if (!(pream_symb_loc > lower_bound && pream_symb_loc < upper_bound)) { /* bounds recovery :734-798 */
		int signal_start_symb = -1;
		for (int s = lower_bound + 1; s < upper_bound; s++)
			if (sym_energy(bbi, s * sym, sym, buf) > 0.001) {
				signal_start_symb = s;
				break;
			}
		if (signal_start_symb >= 0) {
			int search_start = signal_start_symb * sym, available = buf - search_start;
			if (available > pre * sym) {
				double rc;
				int rd = time_sync(m, bbi + search_start, available, rate, 0, step, 1, &rc, loc, vals) + search_start;
				int retry_symb = rd / sym;
				if (retry_symb < 1) retry_symb = 1;
				double re = sym_energy(bbi, rd, sym, buf);
				if (re >= 0.001 && rc >= 0.5 && retry_symb > lower_bound && retry_symb < upper_bound) {
					delay = rd;
					coarse_metric = rc;
					pream_symb_loc = retry_symb;
				}
			}
		}
	}

	if (pream_symb_loc > lower_bound && pream_symb_loc < upper_bound) {
		int energy_ok = 1;
		double mean_energy = sym_energy(bbi, delay, sym, buf); /* :811-838 */
		if (mean_energy < 0.001) energy_ok = 0;
		if (energy_ok && coarse_metric < 0.5) energy_ok = 0; /* :844-851 */
		if (!energy_ok) {				     /* silence skip :861-924 */
			int signal_start_symb = -1;
			for (int s = pream_symb_loc + 1; s < upper_bound; s++)
				if (sym_energy(bbi, s * sym, sym, buf) > 0.001) {
					signal_start_symb = s;
					break;
				}
			if (signal_start_symb >= 0) {
				int search_start = signal_start_symb * sym, available = buf - search_start;
				if (available > pre * sym) {
					double rc;
					int rd = time_sync(m, bbi + search_start, available, rate, 0, step, 1, &rc, loc, vals) + search_start;
					int retry_symb = rd / sym;
					if (retry_symb < 1) retry_symb = 1;
					double re = sym_energy(bbi, rd, sym, buf);
					if (re >= 0.001 && rc >= 0.5 && retry_symb > lower_bound && retry_symb < upper_bound) {
						delay = rd;
						coarse_metric = rc;
						pream_symb_loc = retry_symb;
						energy_ok = 1;
					}
				}
			}
		}
		if (energy_ok) {
			int skip_h_count = 0, skip_h_recovery_attempted = 0;
		skip_h_retry_point:
			while (sync_trials <= f->trials_max) {
				if (sync_trials == f->trials_max && f->use_last_time && last_delay != -1) delay = last_delay; /* :945-948 */
				else if (sync_trials == 1 && f->coarse_freq_sync) { /* :949-1013: +-30 Hz search (g_gui_state.coarse_freq_sync_enabled) */
					const double freq_search[3] = {-30.0, 0.0, 30.0};
					double best_correlation = 0.0, best_offset = 0.0, zero_hz_correlation = 0.0;
					int best_delay = delay;
					for (int i = 0; i < 3; i++) {
						p2b(m, passband, buf, bbi, f->fc + freq_search[i], 0, scratch);
						double corr;
						int d = time_sync(m, bbi, m->Nofdm * (2 * pre + S) * rate, rate, 0, step, 1, &corr, loc, vals);
						if (fabs(freq_search[i]) < 0.1) zero_hz_correlation = corr;
						if (corr > best_correlation) {
							best_correlation = corr;
							best_offset = freq_search[i];
							best_delay = d;
						}
					}
					if (fabs(best_offset) > 1.0 && best_correlation > 0.5 && best_correlation > zero_hz_correlation + 0.1) {
						coarse_freq_offset = best_offset;
						delay = best_delay;
						pream_symb_loc = delay / sym;
						if (pream_symb_loc < 1) pream_symb_loc = 1;
					}
					p2b(m, passband, buf, bbi, f->fc + coarse_freq_offset, 0, scratch);
					delay = (pream_symb_loc - 1) * sym +
						time_sync(m, bbi + (pream_symb_loc - 1) * sym, (pre + 4) * sym, rate, sync_trials, 1, f->trials_max, NULL, loc, vals);
				} else /* :1017 */
					delay = (pream_symb_loc - 1) * sym +
						time_sync(m, bbi + (pream_symb_loc - 1) * sym, (pre + 4) * sym, rate, sync_trials, 1, f->trials_max, NULL, loc, vals);
				if (delay < 0) delay = 0;
				int max_delay = buf - frame_dec * rate; /* :1022-1031 */
				if (delay > max_delay) delay = max_delay;
				{ /* post-fine-sync energy gate :1040-1069 */
					double fine_energy = 0.0;
					for (int i = 0; i < sym && (delay + i) < buf; i++)
						fine_energy += creal(bbi[delay + i]) * creal(bbi[delay + i]) + cimag(bbi[delay + i]) * cimag(bbi[delay + i]);
					fine_energy /= sym;
					if (fine_energy < 0.001) {
						int orig = delay;
						for (int fwd = sym; fwd <= 3 * sym; fwd += sym) {
							int cand = orig + fwd;
							if (cand + sym > buf) break;
							double e = 0.0;
							for (int i = 0; i < sym; i++)
								e += creal(bbi[cand + i]) * creal(bbi[cand + i]) + cimag(bbi[cand + i]) * cimag(bbi[cand + i]);
							e /= sym;
							if (e >= 0.001) {
								delay = cand;
								break;
							}
						}
					}
				}
				double effective_fc = f->fc + coarse_freq_offset;
				p2b(m, passband, buf, bbi, effective_fc, 1, scratch); /* :1081 */
				for (int i = 0, k = 0; i < frame_dec * rate; i += rate) bb[k++] = bbi[delay + i]; /* :1103 */
				if (sync_trials == f->trials_max && f->use_last_freq && last_freq != 0) freq_offset_measured = last_freq;
				else freq_offset_measured = moose(m, bb + MO_NGI, f->bandwidth / (double)m->Nc, pre); /* :1117 */
				if (fabs(freq_offset_measured) > f->ignore_limit) {				      /* :1126-1131 */
					p2b(m, passband, buf, bbi, effective_fc + freq_offset_measured, 1, scratch);
					for (int i = 0, k = 0; i < frame_dec * rate; i += rate) bb[k++] = bbi[delay + i];
				}
				mo_rx_tail(m, bb + pre * m->Nofdm, &ro); /* :1132-1341 */
				if (tail_stats[7] < 0.3) {		   /* mean_H gate :1271-1281 */
					skip_h_count++;
					sync_trials++;
					continue;
				}
				iterations = (int)tail_stats[0];
				crc = (int)tail_stats[1];
				all_zeros = (int)tail_stats[2];
				for (int i = 0; i < m->frame_bytes; i++) out[i] = payload[i]; /* :1329-1332: written before the CRC verdict */
				if (!(int)tail_stats[3]) { /* :1343-1359 */
					SNR = -99.9;
					message_decoded = 0;
					sync_trials++;
				} else {
					SNR = tail_stats[4];
					message_decoded = 1;
					last_freq = freq_offset_measured; /* :1421-1427 */
					freq_offset = freq_offset_measured;
					last_delay = delay;
					break;
				}
			}
			if (!message_decoded && skip_h_count >= f->trials_max + 1 && !skip_h_recovery_attempted) { /* :1436-1504 */
				skip_h_recovery_attempted = 1;
				int search_start_symb = pream_symb_loc + 2, search_start = search_start_symb * sym;
				int search_size = m->Nofdm * (2 * pre + S) * rate;
				int available = buf - search_start;
				if (available > search_size) available = search_size;
				if (search_start_symb < upper_bound && available > pre * sym) {
					p2b(m, passband, buf, bbi, f->fc, 0, scratch);
					double rc;
					int rd = time_sync(m, bbi + search_start, available, rate, 0, step, 1, &rc, loc, vals) + search_start;
					int retry_symb = rd / sym;
					if (retry_symb < 1) retry_symb = 1;
					double re = sym_energy(bbi, rd, sym, buf);
					if (re >= 0.001 && retry_symb > lower_bound && retry_symb < upper_bound) {
						delay = rd;
						coarse_metric = rc;
						pream_symb_loc = retry_symb;
						sync_trials = 0;
						skip_h_count = 0;
						coarse_freq_offset = 0.0;
						goto skip_h_retry_point;
					}
				}
			}
		}
	}
	stats[0] = iterations;
	stats[1] = crc;
	stats[2] = all_zeros;
	stats[3] = message_decoded;
	stats[4] = SNR;
	stats[5] = delay;
	stats[6] = sync_trials;
	stats[7] = freq_offset;
	stats[8] = coarse_metric;
	stats[9] = signal_dbm;
	stats[10] = buf;
	stats[11] = m->frame_bytes;
	state[0] = last_delay;
	state[1] = last_freq;
	state[3] = 0;
	if (baseband_out) memcpy(baseband_out, bb, sizeof(double complex) * frame_dec);
	free(bbi);
	free(scratch);
	free(bb);
	free(loc);
	free(vals);
}

/* Wall time (seconds) of n_calls mo_receive_byte() calls on consecutive capture buffers. */
double mo_receive_byte_timed(const mo_mode *m, const double *passband, int n_calls, int *decoded_flags)
{
	int buf = m->Nofdm * m->fe.buffer_Nsymb * m->fe.interp, out[MO_N / 8];
	double stats[12], state[5] = {0};
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int c = 0; c < n_calls; c++) {
		state[0] = -1;
		state[1] = 0;
		state[4] = -1;
		mo_receive_byte(m, passband + (size_t)c * buf, out, stats, state, NULL);
		if (decoded_flags) decoded_flags[c] = (int)stats[3];
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ================================================================================================
 * TX CHAIN to pass-band: cl_telecom_system::transmit_byte / transmit_bit with
 * message_location == SINGLE_MESSAGE, source/physical_layer/telecom_system.cc:342-553, OFDM branch.
 * ============================================================================================== */

/* cl_FIR::design for the two transmit filters (fir_filter.cc:45-163): tx1 = HPF (spectral inversion of the LPF at the HPF cut) +
 * HAMMING, tx2 = LPF + BLACKMAN; 1000 Hz transition -> 97 taps (physical_config.cc:103-113). */
static void fir_design_tx(int type_hpf, int blackman, double fcut, double tbw, double fs, int *ntaps_out, double *c)
{
	int n = (int)(4.0 / (tbw / (fs / 2.0)));
	if (n % 2 == 0) n++;
	double Ts = 1.0 / (fs), temp;
	c[n / 2] = 1;
	for (int i = 0; i < n / 2; i++) {
		temp = 2 * M_PI * fcut * (double)(n / 2 - i) * Ts;
		c[i] = sin(temp) / temp;
		c[n - i - 1] = c[i];
	}
	temp = 0;
	for (int i = 0; i < n; i++) temp += c[i];
	for (int i = 0; i < n; i++) c[i] /= temp;
	if (type_hpf) {
		for (int i = 0; i < n; i++) c[i] *= -1;
		c[(int)(n - 1) / 2] += 1;
	}
	if (!blackman)
		for (int i = 0; i < n; i++) c[i] *= 0.54 - 0.46 * cos(2.0 * M_PI * (double)i / (n - 1));
	else
		for (int i = 0; i < n; i++) c[i] *= 0.42 - 0.5 * cos(2.0 * M_PI * (double)i / n) + 0.08 * cos(4.0 * M_PI * (double)i / n);
	*ntaps_out = n;
}

/* cl_FIR::apply(double*): fir_filter.cc:189-210. */
static void fir_apply_real(const double *c, int nt, const double *in, double *out, int n)
{
	for (int i = 0; i < n + nt - 1; i++) {
		double acc = 0;
		for (int j = 0; j < nt; j++)
			if ((i - j) >= 0 && (i - j) < n) acc += in[i - j] * c[j];
		if (i >= (nt - 1) / 2 && i < n + (nt - 1) / 2) out[i - (nt - 1) / 2] = acc;
	}
}

/* cl_ofdm::baseband_to_passband: ofdm.cc:2294-2315 (rational_resampler INTERPOLATION :2279-2292, then the mix with the RUNNING
 * carrier sample counter passband_start_sample). */
static void b2p(const mo_mode *m, const double complex *in, int n, double *out, double fc, unsigned long *start_sample)
{
	const mo_frontend *f = &m->fe;
	int rate = f->interp;
	double complex *di = malloc(sizeof(double complex) * n * rate);
	double Ts = 1.0 / f->fs;
	for (int i = 0; i < n - 1; i++)
		for (int j = 0; j < rate; j++) di[i * rate + j] = lerp(in[i], 0, in[i + 1], rate, j);
	for (int j = 0; j < rate; j++) di[(n - 1) * rate + j] = lerp(in[n - 2], 0, in[n - 1], rate, rate + j);
	for (int i = 0; i < n * rate; i++) {
		out[i] = creal(di[i]) * f->amp * cos(2 * M_PI * fc * (double)(*start_sample) * Ts);
		out[i] += cimag(di[i]) * f->amp * sin(2 * M_PI * fc * (double)(*start_sample) * Ts);
		(*start_sample)++;
	}
	free(di);
}

/* cl_ofdm::peak_clip(double*): ofdm.cc:1565-1592. */
static void peak_clip(double *x, int n, double papr)
{
	double avg = 0;
	for (int i = 0; i < n; i++) avg += pow(x[i], 2);
	avg /= n;
	double peak = sqrt(avg * pow(10, papr / 10.0));
	for (int i = 0; i < n; i++) {
		if (x[i] > 0 && x[i] > peak) x[i] = peak;
		if (x[i] < 0 && x[i] < -peak) x[i] = -peak;
	}
}

/* TX-side tables: preamble (cl_preamble_configurator::init/configure, ofdm.cc:1128-1239; QPSK, boost sqrt(2), seed 1,
 * physical_config.cc:50-54), the pilot PRNG stream that precedes it, the transmit FIRs and the pre-equalisation channel
 * (get_pre_equalization_channel, telecom_system.cc:3108-3146: 1000 random symbols through TX filters -> RX filter -> FFT). */
void mo_tx_init(mo_mode *m)
{
	mo_tx *t = &m->tx;
	mo_frontend *f = &m->fe;
	int pre = m->preamble_nSymb;
	if (m->M == 200) { /* MFSK: tone preamble (mfsk.cc:162-195), no pre-equalisation (telecom_system.cc:475-493) */
		t->preamble_boost = sqrt(2);
		t->output_power = 0.1;
		t->preamble_papr = 7, t->data_papr = 10;
		fir_design_tx(1, 0, f->fc - f->bandwidth / 2, 1000, f->fs, &t->ntaps1, t->c1);
		fir_design_tx(0, 1, f->fc + f->bandwidth / 2, 1000, f->fs, &t->ntaps2, t->c2);
		t->start_sample_after_init = 0; /* get_pre_equalization_channel is skipped (:1954): the counter stays where ofdm.init() left it */
		t->ready = 1;
		return;
	}
	/* the preamble is configured BEFORE the pilots (ofdm.cc:112-113): srandom(seed 1), two draws per carrier slot of the sequence */
	mo_srandom(1);
	double complex seq[4 * MO_NC];
	for (int i = 0; i < pre * MO_NC; i++) {
		/* std::complex<double>(2*(__random()%2)-1, 2*(__random()%2)-1): g++ evaluates the constructor arguments right to left */
		int second = mo_random() % 2, first = mo_random() % 2;
		seq[i] = ((double)(2 * first - 1) + (double)(2 * second - 1) * I) / sqrt(2);
	}
	int k = 0;
	for (int s = 0; s < pre; s++)
		for (int c = 0; c < MO_NC; c++) {
			int bin = c < MO_NC / 2 ? c + MO_NFFT - MO_NC / 2 : c - MO_NC / 2 + 1; /* start_shift = 1 */
			int is_pre = (bin % 2 == 0);
			t->preamble_type[s * MO_NC + c] = is_pre;
			t->preamble[s * MO_NC + c] = is_pre ? seq[k++] : 0;
		}
	t->preamble_boost = sqrt(2);
	t->output_power = 0.1;
	t->preamble_papr = 7, t->data_papr = 10;
	fir_design_tx(1, 0, f->fc - f->bandwidth / 2, 1000, f->fs, &t->ntaps1, t->c1);
	fir_design_tx(0, 1, f->fc + f->bandwidth / 2, 1000, f->fs, &t->ntaps2, t->c2);
	/* pre-equalisation: the PRNG continues from where the pilot sequence left it (srandom(0), one draw per pilot, ofdm.cc:940-951) */
	mo_srandom(0);
	for (int i = 0; i < m->nPilots; i++) (void)mo_random();
	int nb = (int)(MO_NC * log2(m->M)), b = m->bits_per_symbol, sym = m->Nofdm * f->interp;
	double complex acc[MO_NC], mod[MO_NC], smod[MO_NOFDM], bb[MO_NOFDM], dem[MO_NC];
	double complex *l = malloc(sizeof(double complex) * sym), *lf = malloc(sizeof(double complex) * sym);
	double *pb = malloc(sizeof(double) * sym), *p1 = malloc(sizeof(double) * sym), *p2 = malloc(sizeof(double) * sym);
	int bits[MO_NC * 8];
	for (int i = 0; i < MO_NC; i++) acc[i] = 0;
	for (int trial = 0; trial < 1000; trial++) {
		for (int i = 0; i < nb; i++) bits[i] = mo_random() % 2;
		for (int i = 0; i < nb; i += b) {
			unsigned loc = 0;
			for (int j = 0; j < b; j++) loc = (loc << 1) | (unsigned)bits[i + j];
			mod[i / b] = m->constellation[loc];
		}
		symbol_mod(m, mod, smod);
		unsigned long start = 0;
		b2p(m, smod, m->Nofdm, pb, f->fc, &start);
		fir_apply_real(t->c1, t->ntaps1, pb, p1, sym);
		fir_apply_real(t->c2, t->ntaps2, p1, p2, sym);
		p2b(m, p2, sym, lf, f->fc, 1, l);
		for (int i = 0, q = 0; i < sym; i += f->interp) bb[q++] = lf[i];
		symbol_demod(m, bb, dem);
		for (int i = 0; i < MO_NC; i++) acc[i] += mod[i] / dem[i];
	}
	for (int i = 0; i < MO_NC; i++) t->pre_eq[i] = acc[i] / (double)1000;
	t->start_sample_after_init = (unsigned long)sym;
	free(l), free(lf), free(pb), free(p1), free(p2);
	t->ready = 1;
}

void mo_tx_tables(mo_mode *m, double complex *preamble, int *preamble_type, double complex *pre_eq, int *ntaps, double *c1, double *c2, double *consts)
{
	if (!m->tx.ready) mo_tx_init(m);
	const mo_tx *t = &m->tx;
	memcpy(preamble, t->preamble, sizeof(double complex) * m->preamble_nSymb * MO_NC);
	memcpy(preamble_type, t->preamble_type, sizeof(int) * m->preamble_nSymb * MO_NC);
	memcpy(pre_eq, t->pre_eq, sizeof(double complex) * MO_NC);
	ntaps[0] = t->ntaps1, ntaps[1] = t->ntaps2;
	memcpy(c1, t->c1, sizeof(double) * t->ntaps1);
	memcpy(c2, t->c2, sizeof(double) * t->ntaps2);
	double v[8] = {t->output_power, t->preamble_boost, t->preamble_papr, t->data_papr, (double)t->start_sample_after_init,
		       (double)((m->Nsymb + m->preamble_nSymb) * m->Nofdm * m->fe.interp), 1, 0};
	memcpy(consts, v, sizeof(v));
}

/* transmit_byte + transmit_bit, SINGLE_MESSAGE: telecom_system.cc:342-553.  Returns total_frame_size; *start_sample is the running
 * carrier sample counter (ofdm.passband_start_sample), in and out. */
static int transmit_byte_impl(mo_mode *m, const int *payload, int nBytes, double *out, double *start_sample_inout, int no_filter);

int mo_transmit_byte(mo_mode *m, const int *payload, int nBytes, double *out, double *start_sample_inout)
{
	return transmit_byte_impl(m, payload, nBytes, out, start_sample_inout, 0);
}

/* message_location == NO_FILTER_MESSAGE (telecom_system.cc:537-544): the clipped pass-band frame before the transmit FIRs, what the
 * ARQ layer asks for (arq_common.cc:2224). */
int mo_transmit_byte_nofilter(mo_mode *m, const int *payload, int nBytes, double *out, double *start_sample_inout)
{
	return transmit_byte_impl(m, payload, nBytes, out, start_sample_inout, 1);
}

/* ofdm.FIR_tx1.apply + ofdm.FIR_tx2.apply over a buffer of any length (arq_common.cc:2243-2246). */
void mo_fir_tx_apply(mo_mode *m, const double *in, int n, double *out)
{
	if (!m->tx.ready) mo_tx_init(m);
	double *tmp = calloc(n, sizeof(double));
	fir_apply_real(m->tx.c1, m->tx.ntaps1, in, tmp, n);
	fir_apply_real(m->tx.c2, m->tx.ntaps2, tmp, out, n);
	free(tmp);
}

static int transmit_byte_impl(mo_mode *m, const int *payload, int nBytes, double *out, double *start_sample_inout, int no_filter)
{
	if (!m->tx.ready) mo_tx_init(m);
	const mo_tx *t = &m->tx;
	const mo_frontend *f = &m->fe;
	int pre = m->preamble_nSymb, S = m->Nsymb, rate = f->interp, No = m->Nofdm;
	int total = (S + pre) * No * rate;
	double complex framed[MO_MAX_CELLS], pdata[4 * MO_NC];
	double complex *bbp = malloc(sizeof(double complex) * pre * No), *bbd = malloc(sizeof(double complex) * S * No);
	double mfsk_boost = 1.0;
	if (m->M == 200) { /* MFSK: tones, no pre-equalisation, drive-level boost (:411-416,461-465,507-515) */
		const mo_mfsk *mf = &m->mfsk;
		mo_tx_baseband(m, payload, nBytes, bbd, NULL, NULL, NULL); /* bit chain + cl_mfsk::mod + symbol_mod */
		double amp = sqrt((double)MO_NC / mf->nStreams);
		for (int s = 0; s < pre; s++) { /* cl_mfsk::generate_preamble: mfsk.cc:162-195 */
			for (int k = 0; k < MO_NC; k++) pdata[k] = 0;
			for (int st = 0; st < mf->nStreams; st++) pdata[mf->stream_offsets[st] + mf->preamble_tones[s % mf->preamble_nSymb]] = amp;
			symbol_mod(m, pdata, bbp + s * No);
		}
		mfsk_boost = sqrt((double)MO_NC / mf->nStreams) * pow(10.0, -2.0 / 20.0);
	} else {
		double complex *scratch = malloc(sizeof(double complex) * S * No);
		mo_tx_baseband(m, payload, nBytes, scratch, NULL, NULL, framed); /* bit chain + framer (:384-462) */
		free(scratch);
		for (int i = 0; i < pre * MO_NC; i++) pdata[i] = t->preamble[i]; /* :466-473 */
		for (int i = 0; i < pre; i++)
			for (int j = 0; j < MO_NC; j++) pdata[i * MO_NC + j] *= t->pre_eq[j]; /* :477-492 */
		for (int i = 0; i < S; i++)
			for (int j = 0; j < MO_NC; j++) framed[i * MO_NC + j] *= t->pre_eq[j];
		for (int i = 0; i < pre; i++) symbol_mod(m, pdata + i * MO_NC, bbp + i * No); /* :495-505 */
		for (int i = 0; i < S; i++) symbol_mod(m, framed + i * MO_NC, bbd + i * No);
	}
	float power_normalization = sqrt((double)(MO_NFFT * rate)); /* :388 */
	for (int j = 0; j < No * pre; j++) {			     /* :517-527 */
		bbp[j] /= power_normalization;
		bbp[j] *= sqrt(t->output_power) * t->preamble_boost * mfsk_boost;
	}
	for (int j = 0; j < No * S; j++) {
		bbd[j] /= power_normalization;
		bbd[j] *= sqrt(t->output_power) * mfsk_boost;
	}
	unsigned long start = (unsigned long)*start_sample_inout;
	double *pb = calloc(total, sizeof(double)), *p1 = malloc(sizeof(double) * total);
	int Sa = active_nsymb(m); /* ctrl frames: only the active symbols are modulated; what follows in the reference's buffer is stale, zero here */
	b2p(m, bbp, No * pre, pb, f->fc, &start); /* :531-532 */
	b2p(m, bbd, No * Sa, pb + No * pre * rate, f->fc, &start);
	peak_clip(pb, No * pre * rate, t->preamble_papr); /* :534-535 */
	peak_clip(pb + No * pre * rate, No * Sa * rate, t->data_papr);
	if (no_filter) {
		memcpy(out, pb, sizeof(double) * total);
	} else {
		fir_apply_real(t->c1, t->ntaps1, pb, p1, total); /* :546-553 */
		fir_apply_real(t->c2, t->ntaps2, p1, out, total);
	}
	*start_sample_inout = (double)start;
	free(bbp), free(bbd), free(pb), free(p1);
	return total;
}


/* ================================================================================================
 * MFSK pattern functions on a pass-band-rate base-band buffer.
 * ============================================================================================== */
static int carrier_to_bin(int sub) { return sub < MO_NC / 2 ? MO_NFFT - MO_NC / 2 + sub : 1 + (sub - MO_NC / 2); }

/* energies of the 50 active carriers of the symbol starting at sample `offset` (every 4th sample, 256-point FFT scaled 1/N:
 * the loop bodies of ofdm.cc:2013-2021 / 2104-2112) */
static void symbol_energies(const mo_mode *m, const double complex *bbi, int offset, int rate, double complex *fft_out)
{
	for (int i = 0; i < MO_NFFT; i++) fft_out[i] = bbi[offset + i * rate];
	fft_core(m, fft_out, 0);
	for (int i = 0; i < MO_NFFT; i++) fft_out[i] = fft_out[i] / (double)MO_NFFT;
}

/* cl_ofdm::time_sync_mfsk: ofdm.cc:1969-2065. */
int mo_time_sync_mfsk(const mo_mode *m, const double complex *bbi, int n, int search_start_symb)
{
	const mo_mfsk *f = &m->mfsk;
	int rate = m->fe.interp, sym = MO_NOFDM * rate, buffer_nsymb = n / sym, pre = m->preamble_nSymb;
	int bins[8][4];
	for (int p = 0; p < pre; p++)
		for (int st = 0; st < f->nStreams; st++) bins[p][st] = carrier_to_bin(f->stream_offsets[st] + f->preamble_tones[p % pre]);
	double best_metric = -1;
	int best = 0;
	double complex fo[MO_NFFT];
	for (int s = search_start_symb > 0 ? search_start_symb : 0; s <= buffer_nsymb - pre; s++) {
		double metric = 0;
		for (int p = 0; p < pre; p++) {
			int offset = (s + p) * sym + MO_NGI * rate;
			if (offset + MO_NFFT * rate > n) break;
			symbol_energies(m, bbi, offset, rate, fo);
			double e_target = 0, e_total = 0;
			for (int st = 0; st < f->nStreams; st++) {
				int b = bins[p][st];
				e_target += creal(fo[b]) * creal(fo[b]) + cimag(fo[b]) * cimag(fo[b]);
			}
			for (int k = 0; k < MO_NC; k++) {
				int b = carrier_to_bin(k);
				double e = creal(fo[b]) * creal(fo[b]) + cimag(fo[b]) * cimag(fo[b]);
				e_total += e;
			}
			if (e_total > 0) metric += e_target / e_total;
		}
		if (metric > best_metric) {
			best_metric = metric;
			best = s;
		}
	}
	return best * sym;
}

/* cl_ofdm::detect_ack_pattern: ofdm.cc:2067-2186, with the ACK or the BREAK tone sequence (mfsk.cc:113-160). */
double mo_detect_ack_pattern(const mo_mode *m, const double complex *bbi, int n, int use_break_tones, int *matched_out)
{
	const mo_mfsk *f = &m->mfsk;
	const int *tones = use_break_tones ? f->break_tones : f->ack_tones;
	int rate = m->fe.interp, sym = MO_NOFDM * rate, buffer_nsymb = n / sym, ack_nsymb = 16;
	if (matched_out) *matched_out = 0;
	if (buffer_nsymb < ack_nsymb) return 0.0;
	double best_metric = 0.0;
	int best_matched = 0;
	double complex fo[MO_NFFT];
	for (int s = 0; s <= buffer_nsymb - ack_nsymb; s++) {
		double metric = 0;
		int matched = 0;
		for (int p = 0; p < ack_nsymb; p++) {
			int offset = (s + p) * sym + MO_NGI * rate;
			if (offset + MO_NFFT * rate > n) break;
			symbol_energies(m, bbi, offset, rate, fo);
			int actual_tone = (tones[p % 8] + p * f->tone_hop_step) % f->M;
			int any_peak = 0;
			double e_target = 0;
			for (int st = 0; st < f->nStreams; st++) {
				int eb = carrier_to_bin(f->stream_offsets[st] + actual_tone);
				double e_exp = creal(fo[eb]) * creal(fo[eb]) + cimag(fo[eb]) * cimag(fo[eb]);
				e_target += e_exp;
				double peak = -1.0;
				for (int t = 0; t < f->M; t++) {
					int b = carrier_to_bin(f->stream_offsets[st] + t);
					double e = creal(fo[b]) * creal(fo[b]) + cimag(fo[b]) * cimag(fo[b]);
					if (e > peak) peak = e;
				}
				if (e_exp >= peak) any_peak = 1;
			}
			if (!any_peak) continue;
			matched++;
			double e_total = 0;
			for (int k = 0; k < MO_NC; k++) {
				int b = carrier_to_bin(k);
				double e = creal(fo[b]) * creal(fo[b]) + cimag(fo[b]) * cimag(fo[b]);
				e_total += e;
			}
			if (e_total > 0) metric += e_target / e_total;
		}
		if (metric > best_metric) {
			best_metric = metric;
			best_matched = matched;
		}
	}
	if (matched_out) *matched_out = best_matched;
	return best_metric;
}

/* cl_mfsk::generate_ack_pattern / generate_break_pattern (mfsk.cc:197-252) + symbol_mod: 16 symbols of base-band. */
void mo_ack_pattern_baseband(const mo_mode *m, int use_break_tones, double complex *out)
{
	const mo_mfsk *f = &m->mfsk;
	const int *tones = use_break_tones ? f->break_tones : f->ack_tones;
	double amp = sqrt((double)MO_NC / f->nStreams);
	for (int s = 0; s < 16; s++) {
		double complex pat[MO_NC];
		for (int k = 0; k < MO_NC; k++) pat[k] = 0;
		int actual = (tones[s % 8] + s * f->tone_hop_step) % f->M;
		for (int st = 0; st < f->nStreams; st++) pat[f->stream_offsets[st] + actual] = amp;
		symbol_mod(m, pat, out + s * MO_NOFDM);
	}
}

void mo_mfsk_tables(const mo_mode *m, int *out)
{
	const mo_mfsk *f = &m->mfsk;
	int k = 0;
	out[k++] = f->M, out[k++] = f->nBits, out[k++] = f->nStreams, out[k++] = f->tone_hop_step;
	for (int i = 0; i < 4; i++) out[k++] = f->stream_offsets[i];
	for (int i = 0; i < 4; i++) out[k++] = f->preamble_tones[i];
	for (int i = 0; i < 8; i++) out[k++] = f->ack_tones[i];
	for (int i = 0; i < 8; i++) out[k++] = f->break_tones[i];
}


/* ================================================================================================
 * The ARQ-facing tone-pattern calls (telecom_system.h:122-130): generate_ack/break_pattern_passband
 * (telecom_system.cc:1589-1631,1657-1689) and detect_ack/break_pattern_from_passband (:1633-1655,
 * 1691-1710).  Config independent: a dedicated cl_mfsk with M = 16, one stream (:3003-3008).
 * ============================================================================================== */
static void ack_mfsk_plan(mo_mfsk *f) { mfsk_init(f, 16, MO_NC, 1); }

int mo_generate_pattern_passband(mo_mode *m, int use_break_tones, double *out, double *start_sample_inout)
{
	mo_mfsk f;
	ack_mfsk_plan(&f);
	const int *tones = use_break_tones ? f.break_tones : f.ack_tones;
	int nsymb = 16, rate = m->fe.interp, n = nsymb * MO_NOFDM * rate;
	double complex *bb = malloc(sizeof(double complex) * nsymb * MO_NOFDM);
	double amp = sqrt((double)MO_NC / f.nStreams);
	for (int s = 0; s < nsymb; s++) {
		double complex pat[MO_NC];
		for (int k = 0; k < MO_NC; k++) pat[k] = 0;
		pat[f.stream_offsets[0] + (tones[s % 8] + s * f.tone_hop_step) % f.M] = amp;
		symbol_mod(m, pat, bb + s * MO_NOFDM);
	}
	float power_normalization = sqrt((double)(MO_NFFT * rate));
	double ack_boost = sqrt((double)MO_NC / f.nStreams) * pow(10.0, -2.0 / 20.0);
	for (int j = 0; j < MO_NOFDM * nsymb; j++) {
		bb[j] /= power_normalization;
		bb[j] *= sqrt(0.1) * ack_boost;
	}
	unsigned long start = (unsigned long)*start_sample_inout;
	b2p(m, bb, MO_NOFDM * nsymb, out, m->fe.fc, &start);
	peak_clip(out, n, 10);
	*start_sample_inout = (double)start;
	free(bb);
	return n;
}

double mo_detect_pattern_from_passband(const mo_mode *m, const double *data, int size, int use_break_tones, int *matched_out)
{
	mo_mode tmp = *m; /* shallow copy: only the tone plan differs (the detector reads m->mfsk, m->fe and the FFT tables) */
	ack_mfsk_plan(&tmp.mfsk);
	double complex *bbi = malloc(sizeof(double complex) * size), *scratch = malloc(sizeof(double complex) * size);
	p2b(m, data, size, bbi, m->fe.fc, 1, scratch);
	double v = mo_detect_ack_pattern(&tmp, bbi, size, use_break_tones, matched_out);
	free(bbi), free(scratch);
	return v;
}


/* g_gui_state.coarse_freq_sync_enabled (gui_state.h:143): the optional +-30 Hz search of trial 1. */
void mo_set_coarse_freq_sync(mo_mode *m, int enable) { m->fe.coarse_freq_sync = enable != 0; }


/* transmit_bit's streaming locations FIRST_MESSAGE / MIDDLE_MESSAGE / FLUSH_MESSAGE (telecom_system.cc:559-594): a three-frame buffer of
 * clipped frames, both FIRs over the two frames centred on the middle one, which is what comes out (one frame of latency). */
void mo_reset_tx_stream(mo_mode *m)
{
	int total = (m->Nsymb + m->preamble_nSymb) * m->Nofdm * m->fe.interp;
	free(m->tx_stream);
	m->tx_stream = calloc(3 * (size_t)total, sizeof(double));
}

int mo_transmit_byte_loc(mo_mode *m, const int *payload, int nBytes, double *out, double *start_sample_inout, int message_location)
{
	if (message_location == 3) return transmit_byte_impl(m, payload, nBytes, out, start_sample_inout, 0);
	if (message_location == 4) return transmit_byte_impl(m, payload, nBytes, out, start_sample_inout, 1);
	if (!m->tx.ready) mo_tx_init(m);
	int T = (m->Nsymb + m->preamble_nSymb) * m->Nofdm * m->fe.interp;
	if (!m->tx_stream) mo_reset_tx_stream(m);
	double *frame = malloc(sizeof(double) * T), *f1 = calloc(2 * (size_t)T, sizeof(double)), *f2 = calloc(2 * (size_t)T, sizeof(double));
	transmit_byte_impl(m, payload, nBytes, frame, start_sample_inout, 1);
	double *buf = m->tx_stream;
	if (message_location == 0)
		for (int i = 0; i < T; i++) buf[T + i] = frame[i], buf[2 * T + i] = frame[i];
	else
		for (int i = 0; i < T; i++) buf[2 * T + i] = frame[i];
	fir_apply_real(m->tx.c1, m->tx.ntaps1, buf + T / 2, f1, 2 * T);
	fir_apply_real(m->tx.c2, m->tx.ntaps2, f1, f2, 2 * T);
	for (int i = 0; i < T; i++) out[i] = f2[T / 2 + i];
	for (int j = 0; j < 2 * T; j++) buf[j] = buf[j + T]; /* shift_left(buffer, 3T, T): misc.cc:24-32 */
	free(frame), free(f1), free(f2);
	return T;
}
