// mb_synth.cpp -- host-side synthesis of OFDM baseband test frames (tests / bench input only; NOT on the RX path).
//
// What: payload bytes -> zero pad -> CRC16 -> scramble -> LDPC (IRA) encode -> compaction -> bit interleave ->
// PSK/QAM map -> T/F interleave -> framer (pilots) -> IFFT-256 + guard interval per symbol, then complex AWGN.
// Reference: transmit_byte/transmit_bit bit chain (telecom_system.cc:342-416), ldpc.encode (ldpc.cc:111-132),
// psk.mod (psk.cc:259-272), framer (ofdm.cc:814-835), symbol_mod (ofdm.cc:855-860), and the noise normalisation of
// baseband_test_EsN0 (telecom_system.cc:139-153: x/sqrt(Nfft) + sigma*CN(0,1), then *sqrt(Nfft)).
//
// It is the exact inverse of the RX index tables in the blob (sym_cell, llr_dst), so the same tables drive both
// directions; tests/test_synth.py checks it against the oracle's TX chain.
#include <atomic>
#include <cmath>
#include <complex>
#include <cstring>
#include <thread>
#include <vector>

#include "mb_tables.h"

namespace {

inline uint64_t splitmix64(uint64_t &x)
{
	uint64_t z = (x += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

uint16_t crc16_modbus(const uint8_t *d, int n)  // crc16_modbus_rtu.cc:25-45
{
	uint16_t crc = 0xFFFF;
	for (int j = 0; j < n; j++) {
		crc ^= d[j];
		for (int i = 0; i < 8; i++) crc = (crc & 1) ? (uint16_t)((crc >> 1) ^ 0xA001) : (uint16_t)(crc >> 1);
	}
	return crc;
}

struct TxTables {
	MbMode m;
	MbRate r;
	std::vector<uint16_t> cw_of_var, sym_cell, llr_dst, pilot_cell;
	std::vector<float> pval, cons;
	std::vector<uint8_t> scr;
	std::vector<std::vector<uint16_t>> check_rows;  // reference check index -> codeword positions
	std::vector<std::complex<double>> tw;           // exp(+2 pi i k / 256)
};

void ifft256(std::complex<double> *v, const std::vector<std::complex<double>> &tw)
{
	const int n = 256;
	for (int i = 0, j = 0; i < n; i++) {
		if (i < j) std::swap(v[i], v[j]);
		int bit = n >> 1;
		for (; j & bit; bit >>= 1) j ^= bit;
		j ^= bit;
	}
	for (int size = 2; size <= n; size *= 2) {
		const int half = size / 2, step = n / size;
		for (int i = 0; i < n; i += size)
			for (int j = 0; j < half; j++) {
				const std::complex<double> t = tw[j * step] * v[i + j + half];
				v[i + j + half] = v[i + j] - t;
				v[i + j] += t;
			}
	}
}

void synth_one(const TxTables &T, const uint8_t *payload, double sigma, uint64_t noise_seed, float *out)
{
	const MbMode &m = T.m;
	const int fs = m.frame_bytes, nR = m.nReal, K = m.K, P = m.P;
	std::vector<uint8_t> bytes(fs + 2), cw(MB_N, 0);
	memcpy(bytes.data(), payload, fs);
	const uint16_t crc = crc16_modbus(bytes.data(), fs);
	bytes[fs] = (uint8_t)(crc & 0xFF);  // LSB first, telecom_system.cc:366-372
	bytes[fs + 1] = (uint8_t)(crc >> 8);
	for (int i = 0; i < nR; i++) {
		const int bit = i < (fs + 2) * 8 ? (bytes[i >> 3] >> (i & 7)) & 1 : 0;
		cw[i] = (uint8_t)(bit ^ T.scr[i]);
	}
	for (int i = 0; i < m.nVirtual; i++) cw[nR + i] = cw[i];
	for (int c = 0; c < P; c++) {  // IRA accumulate: p_c = xor of every other variable of check c
		uint8_t b = 0;
		for (uint16_t v : T.check_rows[c])
			if (v != K + c) b ^= cw[v];
		cw[K + c] = b;
	}
	std::vector<std::complex<double>> grid((size_t)m.Nsymb * MB_NC);
	for (int p = 0; p < m.nPilots; p++) grid[T.pilot_cell[p]] = T.pval[T.pilot_cell[p]];
	for (int q = 0; q < m.nData; q++) {
		unsigned loc = 0;
		for (int t = 0; t < m.bps; t++) loc = (loc << 1) | cw[T.cw_of_var[T.llr_dst[q * m.bps + t]]];
		grid[T.sym_cell[q]] = std::complex<double>(T.cons[2 * loc], T.cons[2 * loc + 1]);
	}
	uint64_t rs = noise_seed;
	for (int s = 0; s < m.Nsymb; s++) {
		std::complex<double> v[256];
		for (auto &x : v) x = 0;
		for (int j = 0; j < 25; j++) v[j + 256 - 25] = grid[(size_t)s * MB_NC + j];   // zero_padder, ofdm.cc:379-400
		for (int j = 25; j < 50; j++) v[j - 25 + 1] = grid[(size_t)s * MB_NC + j];
		ifft256(v, T.tw);
		float *o = out + (size_t)s * MB_NOFDM * 2;
		for (int n = 0; n < MB_NOFDM; n++) {
			std::complex<double> x = n < MB_NGI ? v[n + 256 - MB_NGI] : v[n - MB_NGI];  // gi_adder, ofdm.cc:412-422
			if (sigma > 0) {
				const double u1 = ((double)(splitmix64(rs) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
				const double u2 = (double)(splitmix64(rs) >> 11) * (1.0 / 9007199254740992.0);
				const double rad = std::sqrt(-std::log(u1)) * sigma;  // CN(0, sigma^2): each part N(0, sigma^2/2)
				x += std::complex<double>(rad * std::cos(2.0 * M_PI * u2), rad * std::sin(2.0 * M_PI * u2));
			}
			o[2 * n] = (float)x.real();
			o[2 * n + 1] = (float)x.imag();
		}
	}
}

}  // namespace

std::string mb_synth_frames(const std::vector<uint8_t> &blob, int config, size_t n_frames, uint64_t seed, double esn0_db,
			    const uint8_t *payload_in, float *baseband_out, uint8_t *payload_out, int n_threads)
{
	if (config < 0 || config >= MB_NMODES) return "unknown configuration";
	MbBlobHeader h;
	memcpy(&h, blob.data(), sizeof(h));
	TxTables T;
	T.m = h.modes[config];
	T.r = h.rates[T.m.rate_idx];
	const MbMode &m = T.m;
	const uint8_t *b = blob.data();
	auto u16 = [&](uint32_t off, size_t n) { return std::vector<uint16_t>((const uint16_t *)(b + off), (const uint16_t *)(b + off) + n); };
	std::vector<uint16_t> var_of_cw = u16(T.r.off_var_of_cw, MB_N);
	T.cw_of_var.resize(MB_N);
	for (int i = 0; i < MB_N; i++) T.cw_of_var[var_of_cw[i]] = (uint16_t)i;
	T.sym_cell = u16(m.off_sym_cell, m.nData);
	T.llr_dst = u16(m.off_llr_dst, m.nBits);
	T.pilot_cell = u16(m.off_pilot_cell, m.nPilots);
	T.pval.assign((const float *)(b + m.off_pval), (const float *)(b + m.off_pval) + (size_t)m.Nsymb * MB_NC);
	T.cons.assign((const float *)(b + m.off_const), (const float *)(b + m.off_const) + 2 * m.M);
	T.scr.assign(b + m.off_scr, b + m.off_scr + MB_N);
	{
		const uint8_t *cdeg = b + T.r.off_cdeg;
		const uint32_t *cgbase = (const uint32_t *)(b + T.r.off_cgbase);
		const uint16_t *ev = (const uint16_t *)(b + T.r.off_edge_var);
		const uint16_t *cos_ = (const uint16_t *)(b + T.r.off_check_of_sorted);
		T.check_rows.resize(T.r.P);
		for (int cs = 0; cs < T.r.P; cs++)
			for (int k = 0; k < cdeg[cs]; k++) T.check_rows[cos_[cs]].push_back(T.cw_of_var[ev[mb_ldpc_cslot(cgbase, cdeg[cs & ~31], cs, k)]]);
	}
	T.tw.resize(128);
	for (int k = 0; k < 128; k++) T.tw[k] = std::polar(1.0, 2.0 * M_PI * k / 256.0);
	// baseband_test_EsN0 (telecom_system.cc:98-99,139-153): sigma = 10^(-EsN0/20) on the /sqrt(Nfft) scale
	const double sigma = esn0_db >= 200.0 ? 0.0 : std::pow(10.0, -esn0_db / 20.0) * 16.0;
	const size_t fstride = (size_t)m.Nsymb * MB_NOFDM * 2;
	if (n_threads < 1) n_threads = 1;
	std::atomic<size_t> next(0);
	auto worker = [&]() {
		std::vector<uint8_t> pl(m.frame_bytes);
		for (;;) {
			const size_t f0 = next.fetch_add(64);
			if (f0 >= n_frames) break;
			for (size_t f = f0; f < std::min(n_frames, f0 + 64); f++) {
				if (payload_in) {
					memcpy(pl.data(), payload_in + f * m.frame_bytes, m.frame_bytes);
				} else {
					uint64_t s = seed * 0x9E3779B97F4A7C15ull + f;
					for (int i = 0; i < m.frame_bytes; i++) pl[i] = (uint8_t)(splitmix64(s) >> 56);
				}
				if (payload_out) memcpy(payload_out + f * m.frame_bytes, pl.data(), m.frame_bytes);
				uint64_t ns = (seed ^ 0xA5A5A5A5DEADBEEFull) + 0x632BE59BD9B4E019ull * (f + 1);
				synth_one(T, pl.data(), sigma, ns, baseband_out + f * fstride);
			}
		}
	};
	std::vector<std::thread> th;
	for (int i = 1; i < n_threads; i++) th.emplace_back(worker);
	worker();
	for (auto &t : th) t.join();
	return "";
}
