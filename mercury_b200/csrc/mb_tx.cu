// mb_tx.cu -- the TX chain on the GPU: cl_telecom_system::transmit_byte / transmit_bit with
// message_location == SINGLE_MESSAGE (reference source/physical_layer/telecom_system.cc:342-553, OFDM branch), batched over frames:
//   byte_to_bit + CRC16 + zero pad                      telecom_system.cc:342-382, misc.cc:93-105, crc16_modbus_rtu.cc:25-45
//   bit_energy_dispersal, virtual bits, ldpc.encode     telecom_system.cc:394-409, interleaver.cc:111-117, ldpc.cc:111-132
//   interleaver, psk.mod, interleaver, framer           interleaver.cc:26-75, psk.cc:259-272, ofdm.cc:814-835
//   preamble, pre-equalisation, symbol_mod, power scale telecom_system.cc:466-527, ofdm.cc:855-860,379-422
//   baseband_to_passband (x4 linear interpolation, mix) ofdm.cc:2279-2315
//   peak_clip (preamble / data part), FIR_tx1, FIR_tx2  ofdm.cc:1565-1592, fir_filter.cc:189-210, telecom_system.cc:534-553
// Host side: the TX tables of a mode (preamble sequence, pre-equalisation channel = get_pre_equalization_channel
// telecom_system.cc:3108-3146, transmit FIR designs fir_filter.cc:45-163) are built lazily on the first transmit of that mode.
//
// Numerics: the bit chain is exact; the sample chain is fp64 (FMA allowed, device sincos): pass-band samples agree with the
// reference's doubles to ~1e-12 relative (tests assert 1e-9).  The reference's running carrier sample counter
// (ofdm.passband_start_sample) is an explicit per-frame input.
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstring>
#include <string>
#include <vector>

#include "mb_kernels.cuh"

namespace {

constexpr int kTxTaps = 97;   // (int)(4 / (1000 / 24000)) = 96 -> odd 97 (fir_filter.cc:56-61, physical_config.cc:103-113)
constexpr int kTxHalf = 48;
constexpr int kTxTile = 1024;

// FIR_tx1 / FIR_tx2 (mode independent): in constant memory so that every tap's coefficient is an operand of the DFMA, not a load
__constant__ double tx_c[2][kTxTaps + 1];

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// ---- bits -> framed grid -> IDFT + guard interval -> scaled base-band symbols (preamble first) ----
__global__ void __launch_bounds__(256) k_tx_baseband(const MbTxMode *__restrict__ tm_p, const uint8_t *__restrict__ tb, const uint8_t *__restrict__ payload_all,
						       double2 *__restrict__ bb_all, uint8_t *__restrict__ dbg_cw)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const MbTxMode &tm = *tm_p;
	double2 *X = reinterpret_cast<double2 *>(smem);              // [(pre + S) * 50] carriers, pre-equalised
	double2 *W = X + (tm.pre + tm.S) * MB_NC;                     // [256] exp(+2 pi i k / 256)
	uint8_t *cw = reinterpret_cast<uint8_t *>(W + 256);           // [1600]
	uint8_t *dpar = cw + MB_N;                                    // [P] parity of the data variables of each check row
	__shared__ uint8_t bytes[MB_N / 8 + 2];
	__shared__ uint32_t wtot[8];
	const int b = blockIdx.x, tid = threadIdx.x;
	const int fs = tm.frame_bytes, nR = tm.nReal, K = tm.K, P = tm.P;
	const uint8_t *scr = tb + tm.off_scr;
	const uint16_t *row_off = reinterpret_cast<const uint16_t *>(tb + tm.off_row_off), *row_var = reinterpret_cast<const uint16_t *>(tb + tm.off_row_var);
	for (int i = tid; i < fs; i += 256) bytes[i] = payload_all[(size_t)b * fs + i];
	{
		double s, c;
		sincospi(2.0 * tid / 256.0, &s, &c);
		W[tid] = make_double2(c, s);
	}
	__syncthreads();
	if (tid == 0) {  // CRC16-MODBUS over the zero-padded frame, appended LSB first (telecom_system.cc:363-373)
		unsigned crc = 0xFFFF;
		for (int j = 0; j < fs; j++) {
			crc ^= bytes[j];
			for (int i = 0; i < 8; i++) crc = (crc & 1) ? (crc >> 1) ^ 0xA001 : crc >> 1;
		}
		bytes[fs] = crc & 0xFF, bytes[fs + 1] = crc >> 8;
	}
	__syncthreads();
	for (int i = tid; i < nR; i += 256) {
		const int bit = i < (fs + 2) * 8 ? (bytes[i >> 3] >> (i & 7)) & 1 : 0;
		cw[i] = (uint8_t)(bit ^ scr[i]);
	}
	__syncthreads();
	for (int i = tid; i < tm.nVirtual; i += 256) cw[nR + i] = cw[i];
	__syncthreads();
	// IRA encoder (ldpc.cc:111-132): every check row is {data variables, parity c-1, parity c}, so parity c is the running XOR of
	// the rows' data parities: per-row parity, then a block-wide prefix XOR.
	for (int c = tid; c < P; c += 256) {
		unsigned x = 0;
		for (int e = row_off[c]; e < row_off[c + 1]; e++) x ^= cw[row_var[e]];
		dpar[c] = (uint8_t)x;
	}
	__syncthreads();
	{
		const int per = (P + 255) / 256, c0 = tid * per, c1 = min(P, c0 + per);
		unsigned loc = 0;
		for (int c = c0; c < c1; c++) loc ^= dpar[c];
		unsigned inc = loc;  // inclusive prefix XOR of the chunk totals
		const int lane = tid & 31, warp = tid >> 5;
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc ^= v;
		}
		if (lane == 31) wtot[warp] = inc;
		__syncthreads();
		unsigned run = inc ^ loc;
		for (int w = 0; w < warp; w++) run ^= wtot[w];
		for (int c = c0; c < c1; c++) {
			run ^= dpar[c];
			cw[K + c] = (uint8_t)run;
		}
	}
	__syncthreads();
	if (dbg_cw)
		for (int i = tid; i < MB_N; i += 256) dbg_cw[(size_t)b * MB_N + i] = cw[i];
	// preamble carriers, then the framed data symbols (pilots + mapped data), all times the pre-equalisation channel (:466-492)
	const double2 *pre_eq = reinterpret_cast<const double2 *>(tb + tm.off_pre_eq), *preamble = reinterpret_cast<const double2 *>(tb + tm.off_preamble);
	const double2 *cons = reinterpret_cast<const double2 *>(tb + tm.off_cons);
	const double *pilot = reinterpret_cast<const double *>(tb + tm.off_pilot);
	const uint16_t *sym_cell = reinterpret_cast<const uint16_t *>(tb + tm.off_sym_cell), *bit_src = reinterpret_cast<const uint16_t *>(tb + tm.off_bit_src);
	for (int i = tid; i < tm.pre * MB_NC; i += 256) X[i] = cmul(preamble[i], pre_eq[i % MB_NC]);
	double2 *Xd = X + tm.pre * MB_NC;
	for (int c = tid; c < tm.S * MB_NC; c += 256)
		if (pilot[c] != 0.0) Xd[c] = make_double2(pilot[c] * pre_eq[c % MB_NC].x, pilot[c] * pre_eq[c % MB_NC].y);
	for (int q = tid; q < tm.nData; q += 256) {
		unsigned loc = 0;
		for (int t = 0; t < tm.bps; t++) loc = (loc << 1) | cw[bit_src[q * tm.bps + t]];
		const int cell = sym_cell[q];
		Xd[cell] = cmul(cons[loc], pre_eq[cell % MB_NC]);
	}
	__syncthreads();
	// symbol_mod: zero_padder (carriers c < 25 -> bins 231 + c, c >= 25 -> bins c - 24), unscaled IFFT, guard interval = last 16 samples.
	// Thread n evaluates x[n] = w^231 * P_lo(w) + w * P_hi(w) with w = exp(2 pi i n / 256) by Horner's rule: two 24-step chains of complex
	// FMAs on broadcast carrier reads, no twiddle table (the per-carrier gather W[(bin * n) & 255] was this kernel's bottleneck).
	double2 *bb = bb_all + (size_t)b * (tm.pre + tm.S) * MB_NOFDM;
	const double2 w = W[tid], w231 = W[(231 * tid) & 255];
	for (int s = 0; s < tm.pre + tm.S; s++) {
		const double2 *Xs = X + s * MB_NC;
		double2 lo = Xs[MB_NC / 2 - 1], hi = Xs[MB_NC - 1];
#pragma unroll 8
		for (int c = MB_NC / 2 - 2; c >= 0; c--) {
			const double2 a = Xs[c], h2 = Xs[MB_NC / 2 + c];
			lo = make_double2(fma(lo.x, w.x, fma(-lo.y, w.y, a.x)), fma(lo.x, w.y, fma(lo.y, w.x, a.y)));
			hi = make_double2(fma(hi.x, w.x, fma(-hi.y, w.y, h2.x)), fma(hi.x, w.y, fma(hi.y, w.x, h2.y)));
		}
		const double2 tl = cmul(lo, w231), th = cmul(hi, w);
		const double ar = tl.x + th.x, ai = tl.y + th.y;
		const double sc = s < tm.pre ? tm.scale_pre : tm.scale_data;  // / power_normalization * sqrt(output_power_Watt) [* preamble boost] (:517-527)
		const double2 v = make_double2(ar * sc, ai * sc);
		bb[s * MB_NOFDM + MB_NGI + tid] = v;
		if (tid >= MB_NFFT - MB_NGI) bb[s * MB_NOFDM + tid - (MB_NFFT - MB_NGI)] = v;
	}
}

// ---- ROBUST (MFSK) modes: bits -> tones -> base-band symbols (telecom_system.cc:384-416,461-465,495-527; cl_mfsk::mod mfsk.cc:254-303,
// generate_preamble :162-195).  Every symbol holds one tone per stream, so its IDFT is a sum of nStreams rotating phasors. ----
__global__ void __launch_bounds__(256) k_tx_baseband_mfsk(const MbTxMode *__restrict__ tm_p, const uint8_t *__restrict__ tb, const MbMfsk t, int S_active,
							    const uint8_t *__restrict__ payload_all, double2 *__restrict__ bb_all, uint8_t *__restrict__ dbg_cw)
{
	__shared__ uint8_t cw[MB_N];
	__shared__ uint8_t dpar[MB_N];
	__shared__ uint8_t bytes[MB_N / 8 + 2];
	__shared__ uint32_t wtot[8];
	__shared__ double2 W[MB_NFFT];
	__shared__ uint8_t tone[MB_N / 4][2];  // actual tone of (symbol, stream)
	const MbTxMode &tm = *tm_p;
	const int b = blockIdx.x, tid = threadIdx.x;
	const int fs = tm.frame_bytes, nR = tm.nReal, K = tm.K, P = tm.P;
	const uint8_t *scr = tb + tm.off_scr;
	const uint16_t *row_off = reinterpret_cast<const uint16_t *>(tb + tm.off_row_off), *row_var = reinterpret_cast<const uint16_t *>(tb + tm.off_row_var);
	const uint16_t *bit_src = reinterpret_cast<const uint16_t *>(tb + tm.off_bit_src);
	for (int i = tid; i < fs; i += 256) bytes[i] = payload_all[(size_t)b * fs + i];
	{
		double s, c;
		sincospi(2.0 * tid / 256.0, &s, &c);
		W[tid] = make_double2(c, s);
	}
	__syncthreads();
	if (tid == 0) {
		unsigned crc = 0xFFFF;
		for (int j = 0; j < fs; j++) {
			crc ^= bytes[j];
			for (int i = 0; i < 8; i++) crc = (crc & 1) ? (crc >> 1) ^ 0xA001 : crc >> 1;
		}
		bytes[fs] = crc & 0xFF, bytes[fs + 1] = crc >> 8;
	}
	__syncthreads();
	for (int i = tid; i < nR; i += 256) {
		const int bit = i < (fs + 2) * 8 ? (bytes[i >> 3] >> (i & 7)) & 1 : 0;
		cw[i] = (uint8_t)(bit ^ scr[i]);
	}
	__syncthreads();
	for (int c = tid; c < P; c += 256) {
		unsigned x = 0;
		for (int e = row_off[c]; e < row_off[c + 1]; e++) x ^= cw[row_var[e]];
		dpar[c] = (uint8_t)x;
	}
	__syncthreads();
	{
		const int per = (P + 255) / 256, c0 = tid * per, c1 = min(P, c0 + per);
		unsigned loc = 0;
		for (int c = c0; c < c1; c++) loc ^= dpar[c];
		unsigned inc = loc;
		const int lane = tid & 31, warp = tid >> 5;
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc ^= v;
		}
		if (lane == 31) wtot[warp] = inc;
		__syncthreads();
		unsigned run = inc ^ loc;
		for (int w = 0; w < warp; w++) run ^= wtot[w];
		for (int c = c0; c < c1; c++) {
			run ^= dpar[c];
			cw[K + c] = (uint8_t)run;
		}
	}
	__syncthreads();
	if (dbg_cw)
		for (int i = tid; i < MB_N; i += 256) dbg_cw[(size_t)b * MB_N + i] = cw[i];
	// cl_mfsk::mod: nBits interleaved bits (MSB first) -> Gray -> binary tone index, hopped by s * tone_hop_step
	for (int q = tid; q < tm.S * t.nStreams; q += 256) {
		const int s = q / t.nStreams, st = q % t.nStreams, off = s * tm.bps + st * t.nBits;
		int g = 0;
		for (int k = 0; k < t.nBits; k++) g |= cw[bit_src[off + k]] << (t.nBits - 1 - k);
		int bin = g;
		for (int sh = 1; sh < t.nBits; sh++) bin ^= (g >> sh);
		if (bin >= t.M) bin = t.M - 1;
		tone[s][st] = (uint8_t)((bin + s * t.tone_hop_step) % t.M);
	}
	__syncthreads();
	double2 *bb = bb_all + (size_t)b * (tm.pre + tm.S) * MB_NOFDM;
	for (int s = 0; s < tm.pre + tm.S; s++) {
		double ar = 0, ai = 0;
		for (int st = 0; st < t.nStreams; st++) {
			const int tn = s < tm.pre ? t.preamble_tones[s % 4] : tone[s - tm.pre][st];
			const int c = t.stream_offsets[st] + tn;
			const int bin = c < MB_NC / 2 ? c + MB_NFFT - MB_NC / 2 : c - MB_NC / 2 + 1;
			const double2 w = W[(bin * tid) & 255];
			ar += w.x, ai += w.y;
		}
		const double sc = s < tm.pre ? tm.scale_pre : (s < tm.pre + S_active ? tm.scale_data : 0.0);  // tone amplitude, power normalisation, output power, boosts;
		const double2 v = make_double2(ar * sc, ai * sc);                                            // control frames: nothing after the active symbols
		bb[s * MB_NOFDM + MB_NGI + tid] = v;
		if (tid >= MB_NFFT - MB_NGI) bb[s * MB_NOFDM + tid - (MB_NFFT - MB_NGI)] = v;
	}
}

// ---- baseband_to_passband: x4 linear interpolation inside each part (preamble / data), mix with the running carrier ----
__global__ void __launch_bounds__(256) k_tx_mix(const MbTxMode *__restrict__ tm_p, const double2 *__restrict__ bb_all, const unsigned long long *__restrict__ start_all,
						  double *__restrict__ pb_all, double *__restrict__ power_part, int nblk)
{
	__shared__ double red[2][8];
	const MbTxMode &tm = *tm_p;
	const int b = blockIdx.y, total = (tm.pre + tm.S) * MB_FE_SYM, npre = tm.pre * MB_FE_SYM;
	const int i = blockIdx.x * 256 + threadIdx.x;
	double p0 = 0, p1 = 0;
	if (i < total) {
		const bool in_pre = i < npre;
		const int ip = in_pre ? i : i - npre, n = (in_pre ? tm.pre : tm.S) * MB_NOFDM;
		const double2 *in = bb_all + (size_t)b * (tm.pre + tm.S) * MB_NOFDM + (in_pre ? 0 : tm.pre * MB_NOFDM);
		int k = ip >> 2, j = ip & 3;
		if (k == n - 1) k = n - 2, j += 4;  // the last input sample extrapolates the last segment (ofdm.cc:2288-2291)
		const double2 a = in[k], c = in[k + 1];
		const double t = (double)j / 4.0;  // interpolate_linear: a + (b - a) * (x - 0) / (rate - 0)
		const double re = a.x + (c.x - a.x) * t, im = a.y + (c.y - a.y) * t;
		const unsigned long long n0 = (start_all ? start_all[b] : tm.start_after_init) + (unsigned long long)i;
		double s, co;
		sincos(2 * M_PI * tm.fc * (double)n0 * tm.Ts, &s, &co);
		const double v = re * tm.amp * co + im * tm.amp * s;
		pb_all[(size_t)b * total + i] = v;
		if (in_pre) p0 = v * v;
		else p1 = v * v;
	}
	for (int o = 16; o; o >>= 1) p0 += __shfl_xor_sync(0xffffffffu, p0, o), p1 += __shfl_xor_sync(0xffffffffu, p1, o);
	if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = p0, red[1][threadIdx.x >> 5] = p1;
	__syncthreads();
	if (threadIdx.x < 2) {
		double t = 0;
		for (int w = 0; w < 8; w++) t += red[threadIdx.x][w];
		power_part[((size_t)b * nblk + blockIdx.x) * 2 + threadIdx.x] = t;
	}
}

// ---- one 97-tap real FIR pass (zero-phase, fir_filter.cc:189-210); pass 1 clips its input on the fly (peak_clip) ----
template <bool CLIP, typename OUT, int WHICH>
__global__ void __launch_bounds__(256) k_tx_fir(const double *__restrict__ in_all, int total, int npre, int ndata, double papr_pre_lin, double papr_data_lin,
						  const double *__restrict__ power_part, int nblk_mix, OUT *__restrict__ out_all)
{
	__shared__ double l[(kTxTile + kTxTaps - 1) * 5 / 4 + 2];
	__shared__ double peak[2];
	const int b = blockIdx.y, t0 = blockIdx.x * kTxTile;
	const double *in = in_all + (size_t)b * total;
	if (CLIP) {
		if (threadIdx.x < 2) {  // peak_allowed = sqrt(mean power of the part * 10^(papr/10))  (ofdm.cc:1570-1578)
			double t = 0;
			for (int k = 0; k < nblk_mix; k++) t += power_part[((size_t)b * nblk_mix + k) * 2 + threadIdx.x];
			const int n = threadIdx.x == 0 ? npre : ndata;
			peak[threadIdx.x] = sqrt(t / n * (threadIdx.x == 0 ? papr_pre_lin : papr_data_lin));
		}
		__syncthreads();
	}
	for (int i = threadIdx.x; i < kTxTile + kTxTaps - 1; i += 256) {
		const int n = t0 - kTxHalf + i;
		double v = (n >= 0 && n < total) ? in[n] : 0.0;
		if (CLIP) {
			const double pk = n < npre ? peak[0] : peak[1];
			v = v > pk ? pk : (v < -pk ? -pk : v);
		}
		l[i + (i >> 2)] = v;
	}
	__syncthreads();
	const int o = t0 + 4 * threadIdx.x;
	if (o >= total) return;
	// output o + r, tap j uses in[o + r + 48 - j] = staged index 4 t + r + 96 - j; 4 outputs share a sliding window (see k_fe_p2b_full)
	double acc[4] = {0, 0, 0, 0}, w[4];
	const int base = 5 * threadIdx.x;
#pragma unroll
	for (int r = 1; r < 4; r++) w[r] = l[base + (96 + r) + ((96 + r) >> 2)];
#pragma unroll
	for (int j = 0; j < kTxTaps; j++) {
		w[0] = l[base + (96 - j) + ((96 - j) >> 2)];
		const double cj = tx_c[WHICH][j];
#pragma unroll
		for (int r = 0; r < 4; r++) acc[r] += w[r] * cj;
		w[3] = w[2], w[2] = w[1], w[1] = w[0];
	}
	OUT *out = out_all + (size_t)b * total + o;
#pragma unroll
	for (int r = 0; r < 4; r++)
		if (o + r < total) out[r] = (OUT)acc[r];
}

// NO_FILTER_MESSAGE (telecom_system.cc:537-544): the clipped pass-band frame, before the transmit FIRs
template <typename OUT>
__global__ void __launch_bounds__(256) k_tx_clip(const double *__restrict__ in_all, int total, int npre, int ndata, double papr_pre_lin, double papr_data_lin,
						   const double *__restrict__ power_part, int nblk_mix, OUT *__restrict__ out_all)
{
	__shared__ double peak[2];
	const int b = blockIdx.y;
	if (threadIdx.x < 2) {
		double t = 0;
		for (int k = 0; k < nblk_mix; k++) t += power_part[((size_t)b * nblk_mix + k) * 2 + threadIdx.x];
		const int n = threadIdx.x == 0 ? npre : ndata;
		peak[threadIdx.x] = sqrt(t / n * (threadIdx.x == 0 ? papr_pre_lin : papr_data_lin));
	}
	__syncthreads();
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= total) return;
	const double pk = i < npre ? peak[0] : peak[1];
	double v = in_all[(size_t)b * total + i];
	v = v > pk ? pk : (v < -pk ? -pk : v);
	out_all[(size_t)b * total + i] = (OUT)v;
}

// ------------------------------------------------------------------------------------------------------------------
// host: TX tables of one mode
// ------------------------------------------------------------------------------------------------------------------
typedef std::complex<double> cd;

void fir_design_tx(bool hpf, bool blackman, double fcut, double tbw, double fs, double *c)
{  // fir_filter.cc:45-163
	int n = (int)(4.0 / (tbw / (fs / 2.0)));
	if (n % 2 == 0) n++;
	const double Ts = 1.0 / (fs);
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
	if (hpf) {
		for (int i = 0; i < n; i++) c[i] *= -1;
		c[(n - 1) / 2] += 1;
	}
	for (int i = 0; i < n; i++)
		c[i] *= blackman ? 0.42 - 0.5 * cos(2.0 * M_PI * (double)i / n) + 0.08 * cos(4.0 * M_PI * (double)i / n)
				 : 0.54 - 0.46 * cos(2.0 * M_PI * (double)i / (n - 1));
}

void build_constellation_d(int M, std::vector<cd> &out)
{  // psk.cc:65-256 (unit mean power, the normaliser accumulated and kept in float)
	static const signed char q16[16][2] = {{-3, 3}, {-3, 1}, {-3, -3}, {-3, -1}, {-1, 3}, {-1, 1}, {-1, -3}, {-1, -1},
					       {3, 3},	{3, 1},	 {3, -3},  {3, -1},  {1, 3},  {1, 1},  {1, -3},	 {1, -1}};
	static const signed char q32[32][2] = {{-3, 5}, {-1, 5}, {-3, -5}, {-1, -5}, {-5, 3}, {-5, 1}, {-5, -3}, {-5, -1},
					       {-1, 3}, {-1, 1}, {-1, -3}, {-1, -1}, {-3, 3}, {-3, 1}, {-3, -3}, {-3, -1},
					       {3, 5},	{1, 5},	 {3, -5},  {1, -5},  {5, 3},  {5, 1},  {5, -3},	 {5, -1},
					       {1, 3},	{1, 1},	 {1, -3},  {1, -1},  {3, 3},  {3, 1},  {3, -3},	 {3, -1}};
	out.assign(M, cd(0, 0));
	const double h = std::sqrt(2.0) / 2.0;
	if (M == 2) out = {cd(1, 0), cd(-1, 0)};
	else if (M == 4) out = {cd(-1, 1), cd(-1, -1), cd(1, 1), cd(1, -1)};
	else if (M == 8) out = {cd(-1, -1) * h, cd(-1, 0), cd(0, 1), cd(-1, 1) * h, cd(0, -1), cd(1, -1) * h, cd(1, 1) * h, cd(1, 0)};
	else if (M == 16)
		for (int i = 0; i < 16; i++) out[i] = cd(q16[i][0], q16[i][1]);
	else
		for (int i = 0; i < 32; i++) out[i] = cd(q32[i][0], q32[i][1]);
	float pn = 0;
	for (int i = 0; i < M; i++) pn += out[i].real() * out[i].real() + out[i].imag() * out[i].imag();
	pn = 1 / (std::sqrt(pn / M));
	for (int i = 0; i < M; i++) out[i] *= (double)pn;
}

inline int carrier_bin(int c) { return c < MB_NC / 2 ? c + MB_NFFT - MB_NC / 2 : c - MB_NC / 2 + 1; }

template <typename T>
uint32_t put(std::vector<uint8_t> &buf, const T *p, size_t n)
{
	while (buf.size() % 16) buf.push_back(0);
	const uint32_t off = (uint32_t)buf.size();
	buf.insert(buf.end(), reinterpret_cast<const uint8_t *>(p), reinterpret_cast<const uint8_t *>(p) + n * sizeof(T));
	return off;
}

}  // namespace

// Builds MbTxMode + its table bytes for `config` from the RX blob (index tables, scrambler, check rows) and the front-end constants.
std::string mb_tx_build(const std::vector<uint8_t> &blob, int config, const MbFeConst &fe, MbTxMode *tm, std::vector<uint8_t> *bytes)
{
	MbBlobHeader h;
	memcpy(&h, blob.data(), sizeof(h));
	const MbMode &m = h.modes[config];
	const MbRate &r = h.rates[m.rate_idx];
	const uint8_t *b = blob.data();
	memset(tm, 0, sizeof(*tm));
	tm->S = m.Nsymb, tm->pre = m.preamble_nSymb, tm->nData = m.nData, tm->nPilots = m.nPilots, tm->nBits = m.nBits, tm->nReal = m.nReal;
	tm->nVirtual = m.nVirtual, tm->K = m.K, tm->P = m.P, tm->bps = m.bps, tm->M = m.M, tm->frame_bytes = m.frame_bytes;
	tm->fc = fe.fc, tm->Ts = fe.Ts, tm->amp = fe.amp;
	const double power_normalization = (double)(float)std::sqrt((double)(MB_NFFT * 4));  // telecom_system.cc:388 (a float)
	tm->scale_data = 1.0 / power_normalization * (std::sqrt(0.1) * 1.0);               // output_power_Watt 0.1 (physical_config.cc:89)
	tm->scale_pre = 1.0 / power_normalization * (std::sqrt(0.1) * std::sqrt(2) * 1.0);  // preamble boost sqrt(2) (:53)
	tm->papr_pre_lin = std::pow(10, 7 / 10.0), tm->papr_data_lin = std::pow(10, 10 / 10.0);  // physical_config.cc:115-116
	tm->start_after_init = MB_FE_SYM;  // get_pre_equalization_channel leaves the carrier counter one symbol in (telecom_system.cc:3125-3126)
	bytes->clear();
	// index tables from the RX blob: the TX maps are the inverses the host synthesiser already uses (mb_synth.cpp)
	const uint16_t *var_of_cw = reinterpret_cast<const uint16_t *>(b + r.off_var_of_cw);
	std::vector<uint16_t> cw_of_var(MB_N);
	for (int i = 0; i < MB_N; i++) cw_of_var[var_of_cw[i]] = (uint16_t)i;
	const uint16_t *llr_dst = reinterpret_cast<const uint16_t *>(b + m.off_llr_dst);
	std::vector<uint16_t> bit_src(m.nBits);
	for (int i = 0; i < m.nBits; i++) bit_src[i] = cw_of_var[llr_dst[i]];
	tm->off_bit_src = put(*bytes, bit_src.data(), bit_src.size());
	tm->off_sym_cell = put(*bytes, reinterpret_cast<const uint16_t *>(b + m.off_sym_cell), (size_t)m.nData);
	tm->off_scr = put(*bytes, b + m.off_scr, (size_t)MB_N);
	{  // check rows in the reference's check order, data variables only (codeword positions < K)
		const uint8_t *cdeg = b + r.off_cdeg;
		const uint32_t *cgbase = reinterpret_cast<const uint32_t *>(b + r.off_cgbase);
		const uint16_t *ev = reinterpret_cast<const uint16_t *>(b + r.off_edge_var), *cos_ = reinterpret_cast<const uint16_t *>(b + r.off_check_of_sorted);
		std::vector<std::vector<uint16_t>> rows(r.P);
		for (int cs = 0; cs < r.P; cs++)
			for (int k = 0; k < cdeg[cs]; k++) {
				const uint16_t v = cw_of_var[ev[mb_ldpc_cslot(cgbase, cdeg[cs & ~31], cs, k)]];
				const int c = cos_[cs];
				if (v < r.K) rows[c].push_back(v);
				else if (v != r.K + c && v != r.K + c - 1) return "check row is not {data, parity c-1, parity c}: the prefix-XOR encoder does not apply";
			}
		std::vector<uint16_t> off(r.P + 1, 0), var;
		for (int c = 0; c < r.P; c++) {
			for (uint16_t v : rows[c]) var.push_back(v);
			off[c + 1] = (uint16_t)var.size();
		}
		tm->off_row_off = put(*bytes, off.data(), off.size());
		tm->off_row_var = put(*bytes, var.data(), var.size());
	}
	std::vector<double> pilot((size_t)m.Nsymb * MB_NC);
	const float *pval = reinterpret_cast<const float *>(b + m.off_pval);
	for (size_t i = 0; i < pilot.size(); i++) pilot[i] = (double)pval[i];
	tm->off_pilot = put(*bytes, pilot.data(), pilot.size());
	std::vector<cd> cons;
	build_constellation_d(m.M, cons);
	tm->off_cons = put(*bytes, cons.data(), cons.size());
	// preamble: configured before the pilots (ofdm.cc:112-113): srandom(1), QPSK / sqrt(2), two draws per sequence slot, the imaginary
	// part drawn first (g++ evaluates the constructor arguments right to left); even FFT bins only (ofdm.cc:1191-1239)
	uint32_t st[35];
	mb_srandom(st, 1);
	std::vector<cd> seq((size_t)m.preamble_nSymb * MB_NC), preamble((size_t)m.preamble_nSymb * MB_NC);
	for (auto &s : seq) {
		const int second = mb_random(st) % 2, first = mb_random(st) % 2;
		s = cd(2 * first - 1, 2 * second - 1) / std::sqrt(2);
	}
	size_t k = 0;
	for (int s = 0; s < m.preamble_nSymb; s++)
		for (int c = 0; c < MB_NC; c++) preamble[(size_t)s * MB_NC + c] = (carrier_bin(c) % 2 == 0) ? seq[k++] : cd(0, 0);
	tm->off_preamble = put(*bytes, preamble.data(), preamble.size());
	std::vector<double> c1(kTxTaps), c2(kTxTaps);
	fir_design_tx(true, false, fe.fc - fe.bandwidth / 2, 1000, fe.fs, c1.data());
	fir_design_tx(false, true, fe.fc + fe.bandwidth / 2, 1000, fe.fs, c2.data());
	tm->off_c1 = put(*bytes, c1.data(), c1.size());
	tm->off_c2 = put(*bytes, c2.data(), c2.size());
	// pre-equalisation channel (telecom_system.cc:3108-3146): 1000 random symbols of this mode's constellation through symbol_mod,
	// baseband_to_passband (counter reset to 0), FIR_tx1, FIR_tx2, passband_to_baseband (FIR_rx_data, decimation 4), symbol_demod; the
	// PRNG continues from where the pilot sequence left it (srandom(0), one draw per pilot)
	mb_srandom(st, 0);
	for (int i = 0; i < m.nPilots; i++) (void)mb_random(st);
	const int nb = (int)(MB_NC * std::log2((double)m.M)), sym = MB_FE_SYM;
	std::vector<cd> acc(MB_NC, cd(0, 0)), mod(MB_NC), tdom(MB_NOFDM), W(256);
	for (int i = 0; i < 256; i++) W[i] = std::polar(1.0, 2.0 * M_PI * i / 256.0);
	std::vector<double> pb(sym), p1(sym), p2(sym);
	std::vector<cd> mixed(sym);
	auto fir_real = [&](const std::vector<double> &c, const std::vector<double> &in, std::vector<double> &out) {
		for (int o = 0; o < sym; o++) {
			double a = 0;
			for (int j = 0; j < kTxTaps; j++) {
				const int n = o + kTxHalf - j;
				if (n >= 0 && n < sym) a += in[n] * c[j];
			}
			out[o] = a;
		}
	};
	std::vector<int> bits(nb);
	for (int trial = 0; trial < 1000; trial++) {
		for (int i = 0; i < nb; i++) bits[i] = mb_random(st) % 2;
		for (int i = 0; i < nb; i += m.bps) {
			unsigned loc = 0;
			for (int j = 0; j < m.bps; j++) loc = (loc << 1) | (unsigned)bits[i + j];
			mod[i / m.bps] = cons[loc];
		}
		for (int n = 0; n < MB_NFFT; n++) {  // symbol_mod
			cd a(0, 0);
			for (int c = 0; c < MB_NC; c++) a += mod[c] * W[(carrier_bin(c) * n) & 255];
			tdom[MB_NGI + n] = a;
		}
		for (int n = 0; n < MB_NGI; n++) tdom[n] = tdom[MB_NFFT + n];
		for (int i = 0; i < sym; i++) {  // baseband_to_passband from counter 0
			int kk = i >> 2, j = i & 3;
			if (kk == MB_NOFDM - 1) kk = MB_NOFDM - 2, j += 4;
			const cd v = tdom[kk] + (tdom[kk + 1] - tdom[kk]) * ((double)j / 4.0);
			pb[i] = v.real() * fe.amp * cos(2 * M_PI * fe.fc * (double)i * fe.Ts) + v.imag() * fe.amp * sin(2 * M_PI * fe.fc * (double)i * fe.Ts);
		}
		fir_real(c1, pb, p1);
		fir_real(c2, p1, p2);
		for (int i = 0; i < sym; i++)
			mixed[i] = cd(p2[i] * fe.amp * cos(2 * M_PI * fe.fc * (double)i * fe.Ts), p2[i] * fe.amp * sin(2 * M_PI * fe.fc * (double)i * fe.Ts));
		std::vector<cd> bb(MB_NOFDM);
		for (int q = 0; q < MB_NOFDM; q++) {  // FIR_rx_data at the decimated positions
			const int o = 4 * q;
			cd a(0, 0);
			for (int j = 0; j < MB_FE_TAPS; j++) {
				const int n = o + MB_FE_TAPS / 2 - j;
				if (n >= 0 && n < sym) a += mixed[n] * fe.c_data[j];
			}
			bb[q] = a;
		}
		for (int c = 0; c < MB_NC; c++) {  // symbol_demod: 256-point DFT / 256 at the active bins
			cd a(0, 0);
			const int bin = carrier_bin(c);
			for (int n = 0; n < MB_NFFT; n++) a += bb[MB_NGI + n] * std::conj(W[(bin * n) & 255]);
			acc[c] += mod[c] / (a / 256.0);
		}
	}
	std::vector<cd> pre_eq(MB_NC);
	for (int c = 0; c < MB_NC; c++) pre_eq[c] = acc[c] / 1000.0;
	tm->off_pre_eq = put(*bytes, pre_eq.data(), pre_eq.size());
	return "";
}

// TX tables of a ROBUST (MFSK) mode: no pilots, constellation, preamble table or pre-equalisation; the interleaver map, the
// scrambler, the check rows and the transmit FIRs remain.
std::string mb_tx_build_mfsk(const std::vector<uint8_t> &blob, const MbMode &m, const MbMfsk &t, const MbFeConst &fe, MbTxMode *tm, std::vector<uint8_t> *bytes)
{
	MbBlobHeader h;
	memcpy(&h, blob.data(), sizeof(h));
	const MbRate &r = h.rates[m.rate_idx];
	const uint8_t *b = blob.data();
	memset(tm, 0, sizeof(*tm));
	tm->S = m.Nsymb, tm->pre = m.preamble_nSymb, tm->nData = m.nData, tm->nPilots = 0, tm->nBits = m.nBits, tm->nReal = m.nReal;
	tm->nVirtual = m.nVirtual, tm->K = m.K, tm->P = m.P, tm->bps = m.bps, tm->M = m.M, tm->frame_bytes = m.frame_bytes;
	tm->fc = fe.fc, tm->Ts = fe.Ts, tm->amp = fe.amp;
	const double power_normalization = (double)(float)std::sqrt((double)(MB_NFFT * 4));
	const double amp = std::sqrt((double)MB_NC / t.nStreams);                                        // mfsk.cc:166,262
	const double mfsk_boost = std::sqrt((double)MB_NC / t.nStreams) * std::pow(10.0, -2.0 / 20.0);  // telecom_system.cc:511-515
	tm->scale_data = amp / power_normalization * (std::sqrt(0.1) * mfsk_boost);
	tm->scale_pre = amp / power_normalization * (std::sqrt(0.1) * std::sqrt(2) * mfsk_boost);
	tm->papr_pre_lin = std::pow(10, 7 / 10.0), tm->papr_data_lin = std::pow(10, 10 / 10.0);
	tm->start_after_init = 0;  // get_pre_equalization_channel is skipped for MFSK (telecom_system.cc:1954): the counter stays at 0
	bytes->clear();
	// interleaved bit j <- compacted codeword bit i with il_dst(i) = j (interleaver.cc:26-52); no virtual bits, so compacted index = codeword position
	const int bs = m.nBits / 10, nb = m.nBits / bs;
	std::vector<uint16_t> bit_src(m.nBits);
	for (int i = 0; i < m.nBits; i++) bit_src[i < nb * bs ? (i % bs) * nb + i / bs : i] = (uint16_t)i;
	tm->off_bit_src = put(*bytes, bit_src.data(), bit_src.size());
	std::vector<uint8_t> scr(MB_N);
	uint32_t st[35];
	mb_srandom(st, 0);
	for (int i = 0; i < MB_N; i++) scr[i] = (uint8_t)(mb_random(st) % 2);
	tm->off_scr = put(*bytes, scr.data(), scr.size());
	{
		const uint16_t *var_of_cw = reinterpret_cast<const uint16_t *>(b + r.off_var_of_cw);
		std::vector<uint16_t> cw_of_var(MB_N);
		for (int i = 0; i < MB_N; i++) cw_of_var[var_of_cw[i]] = (uint16_t)i;
		const uint8_t *cdeg = b + r.off_cdeg;
		const uint32_t *cgbase = reinterpret_cast<const uint32_t *>(b + r.off_cgbase);
		const uint16_t *ev = reinterpret_cast<const uint16_t *>(b + r.off_edge_var), *cos_ = reinterpret_cast<const uint16_t *>(b + r.off_check_of_sorted);
		std::vector<std::vector<uint16_t>> rows(r.P);
		for (int cs = 0; cs < r.P; cs++)
			for (int k = 0; k < cdeg[cs]; k++) {
				const uint16_t v = cw_of_var[ev[mb_ldpc_cslot(cgbase, cdeg[cs & ~31], cs, k)]];
				const int c = cos_[cs];
				if (v < r.K) rows[c].push_back(v);
				else if (v != r.K + c && v != r.K + c - 1) return "check row is not {data, parity c-1, parity c}";
			}
		std::vector<uint16_t> off(r.P + 1, 0), var;
		for (int c = 0; c < r.P; c++) {
			for (uint16_t v : rows[c]) var.push_back(v);
			off[c + 1] = (uint16_t)var.size();
		}
		tm->off_row_off = put(*bytes, off.data(), off.size());
		tm->off_row_var = put(*bytes, var.data(), var.size());
	}
	std::vector<double> c1(kTxTaps), c2(kTxTaps);
	fir_design_tx(true, false, fe.fc - fe.bandwidth / 2, 1000, fe.fs, c1.data());
	fir_design_tx(false, true, fe.fc + fe.bandwidth / 2, 1000, fe.fs, c2.data());
	tm->off_c1 = put(*bytes, c1.data(), c1.size());
	tm->off_c2 = put(*bytes, c2.data(), c2.size());
	return "";
}

size_t mb_tx_smem_bytes(const MbTxMode &tm) { return (size_t)((tm.pre + tm.S) * MB_NC + 256) * sizeof(double2) + MB_N + (size_t)tm.P + 16; }

cudaError_t mb_tx_init(const MbFeConst &fe)
{
	double c[2][kTxTaps + 1] = {};
	fir_design_tx(true, false, fe.fc - fe.bandwidth / 2, 1000, fe.fs, c[0]);   // FIR_tx1: HPF, Hamming (physical_config.cc:103-107)
	fir_design_tx(false, true, fe.fc + fe.bandwidth / 2, 1000, fe.fs, c[1]);   // FIR_tx2: LPF, Blackman (:109-113)
	cudaError_t e = cudaMemcpyToSymbol(tx_c, c, sizeof(c));
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute(k_tx_baseband, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

cudaError_t mb_tx_launch(const MbTxArgs &a, cudaStream_t s)
{
	const int total = (a.tm_host->pre + a.tm_host->S) * MB_FE_SYM;
	const int nblk_mix = (total + 255) / 256;
	const int S_active = (a.S_active > 0 && a.S_active < a.tm_host->S) ? a.S_active : a.tm_host->S;
	const int ndata = S_active * MB_FE_SYM;
	if (a.tm_host->M == 200) k_tx_baseband_mfsk<<<a.n, 256, 0, s>>>(a.tm, a.tables, *a.tone, S_active, a.payload, a.bb, a.dbg_cw);
	else k_tx_baseband<<<a.n, 256, mb_tx_smem_bytes(*a.tm_host), s>>>(a.tm, a.tables, a.payload, a.bb, a.dbg_cw);
	k_tx_mix<<<dim3(nblk_mix, a.n), 256, 0, s>>>(a.tm, a.bb, a.start_sample, a.pb, a.power_part, nblk_mix);
	const int npre = a.tm_host->pre * MB_FE_SYM;
	const double pp = a.tm_host->papr_pre_lin, pd = a.tm_host->papr_data_lin;
	if (a.no_filter) {
		const dim3 g2(nblk_mix, a.n);
		if (a.out_f32) k_tx_clip<float><<<g2, 256, 0, s>>>(a.pb, total, npre, ndata, pp, pd, a.power_part, nblk_mix, static_cast<float *>(a.out));
		else k_tx_clip<double><<<g2, 256, 0, s>>>(a.pb, total, npre, ndata, pp, pd, a.power_part, nblk_mix, static_cast<double *>(a.out));
		return cudaGetLastError();
	}
	const dim3 grid((total + kTxTile - 1) / kTxTile, a.n);
	k_tx_fir<true, double, 0><<<grid, 256, 0, s>>>(a.pb, total, npre, ndata, pp, pd, a.power_part, nblk_mix, a.p1);
	if (a.out_f32) k_tx_fir<false, float, 1><<<grid, 256, 0, s>>>(a.p1, total, 0, 0, 0.0, 0.0, nullptr, 0, static_cast<float *>(a.out));
	else k_tx_fir<false, double, 1><<<grid, 256, 0, s>>>(a.p1, total, 0, 0, 0.0, 0.0, nullptr, 0, static_cast<double *>(a.out));
	return cudaGetLastError();
}

// ofdm.FIR_tx1.apply + ofdm.FIR_tx2.apply over one device buffer of n doubles (the ARQ layer filters a whole padded batch of
// NO_FILTER frames at once, arq_common.cc:2243-2246); tmp: n doubles of scratch.
cudaError_t mb_tx_fir_apply(const uint8_t *tables, const MbTxMode &tm_host, const double *d_in, int n, double *d_tmp, double *d_out, cudaStream_t s)
{
	const dim3 grid((n + kTxTile - 1) / kTxTile, 1);
	(void)tables, (void)tm_host;
	k_tx_fir<false, double, 0><<<grid, 256, 0, s>>>(d_in, n, 0, 0, 0.0, 0.0, nullptr, 0, d_tmp);
	k_tx_fir<false, double, 1><<<grid, 256, 0, s>>>(d_tmp, n, 0, 0, 0.0, 0.0, nullptr, 0, d_out);
	return cudaGetLastError();
}

namespace {
// 16 symbols, one hopped tone each (cl_mfsk::generate_ack_pattern / generate_break_pattern, mfsk.cc:197-252), scaled like :1611-1617
__global__ void __launch_bounds__(256) k_tx_pattern_baseband(const MbMfsk t, int use_break, double scale, double2 *__restrict__ bb)
{
	const int tid = threadIdx.x;
	const int *tones = use_break ? t.break_tones : t.ack_tones;
	for (int s = 0; s < 16; s++) {
		const int c = t.stream_offsets[0] + (tones[s % 8] + s * t.tone_hop_step) % t.M;
		const int bin = c < MB_NC / 2 ? c + MB_NFFT - MB_NC / 2 : c - MB_NC / 2 + 1;
		double sn, cs;
		sincospi(2.0 * ((bin * tid) & 255) / 256.0, &sn, &cs);
		const double2 v = make_double2(cs * scale, sn * scale);
		bb[s * MB_NOFDM + MB_NGI + tid] = v;
		if (tid >= MB_NFFT - MB_NGI) bb[s * MB_NOFDM + tid - (MB_NFFT - MB_NGI)] = v;
	}
}
}  // namespace

cudaError_t mb_tx_pattern(const MbMfsk &plan, int use_break_tones, double fc, double Ts, double amp, unsigned long long start_sample, double2 *d_bb, double *d_pb,
			  double *d_power_part, double *d_out, cudaStream_t s)
{
	// one "frame" without a preamble part: 16 symbols, clipped at the data PAPR (ofdm.peak_clip(out, samples, data_papr_cut), :1629)
	MbTxMode tm;
	memset(&tm, 0, sizeof(tm));
	tm.S = 16, tm.pre = 0, tm.fc = fc, tm.Ts = Ts, tm.amp = amp, tm.start_after_init = start_sample;
	tm.papr_pre_lin = tm.papr_data_lin = std::pow(10, 10 / 10.0);
	MbTxMode *d_tm = nullptr;
	cudaError_t e = cudaMalloc(&d_tm, sizeof(tm));
	if (e != cudaSuccess) return e;
	e = cudaMemcpyAsync(d_tm, &tm, sizeof(tm), cudaMemcpyHostToDevice, s);
	const double power_normalization = (double)(float)std::sqrt((double)(MB_NFFT * 4));
	const double tone_amp = std::sqrt((double)MB_NC / plan.nStreams), boost = std::sqrt((double)MB_NC / plan.nStreams) * std::pow(10.0, -2.0 / 20.0);
	const int total = 16 * MB_FE_SYM, nblk = (total + 255) / 256;
	if (e == cudaSuccess) {
		k_tx_pattern_baseband<<<1, 256, 0, s>>>(plan, use_break_tones, tone_amp / power_normalization * (std::sqrt(0.1) * boost), d_bb);
		k_tx_mix<<<dim3(nblk, 1), 256, 0, s>>>(d_tm, d_bb, nullptr, d_pb, d_power_part, nblk);
		k_tx_clip<double><<<dim3(nblk, 1), 256, 0, s>>>(d_pb, total, 0, total, tm.papr_pre_lin, tm.papr_data_lin, d_power_part, nblk, d_out);
		e = cudaGetLastError();
	}
	cudaError_t e2 = cudaStreamSynchronize(s);
	cudaFree(d_tm);
	return e != cudaSuccess ? e : e2;
}
