// mb_mfsk.cu -- the MFSK row: the ROBUST_0..2 modes' demodulator and the tone-pattern detectors.
//   k_mfsk_demod     symbol_demod (ofdm.cc:862-867) + cl_mfsk::demod (mfsk.cc:305-390) + bit de-interleaver (interleaver.cc:77-92)
//                    on synchronised base-band frames -> LLRs in the LDPC kernel's hand-off layout ("same FFT, different demapper":
//                    the MFSK branch of the RX tail, telecom_system.cc:1132-1198; the decoder kernel is the OFDM modes' own)
//   k_mfsk_energies  per-symbol energies of the 50 active carriers of a pass-band-rate base-band buffer (every 4th sample,
//   k_mfsk_patterns  256-point DFT / 256), then cl_ofdm::time_sync_mfsk (ofdm.cc:1969-2065) and cl_ofdm::detect_ack_pattern
//                    (ofdm.cc:2067-2186) for the ACK and the BREAK tone sequences (mfsk.cc:113-160) in one pass
// The demodulator is fp32 like the OFDM one (LLRs within 1e-4 of the reference's); the pattern detectors are fp64 because their
// outputs are an ARGMAX (the sync delay) and a thresholded metric.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "mb_fft.cuh"
#include "mb_kernels.cuh"

namespace {

__device__ __forceinline__ int carrier_bin(int c) { return c < MB_NC / 2 ? c + MB_NFFT - MB_NC / 2 : c - MB_NC / 2 + 1; }

// 16 threads per symbol (the OFDM demodulator's FFT-256: two radix-4x4 16-point DFTs in registers around a padded shared-memory
// transpose, twiddles by running product, second DFT pruned to the 4 outputs per thread that reach the 50 active carriers);
// a 128-thread CTA handles 8 symbols per round (measured on B200, ms per 8,192 ROBUST_0 / ROBUST_2 frames: 16 per round 1.58 / 0.87,
// 8 per round 1.46 / 0.76, 4 per round 1.49 / 0.77 -- like the OFDM demodulator, smaller CTAs overlap loads and arithmetic at a finer grain).
#ifndef MB_MFSK_SYM_PER_ROUND
#define MB_MFSK_SYM_PER_ROUND 8
#endif
constexpr int kSymPerRound = MB_MFSK_SYM_PER_ROUND;

__global__ void __launch_bounds__(16 * kSymPerRound) k_mfsk_demod(const MbMfskArgs a)
{
	__shared__ float2 scr[kSymPerRound][16 * 17];
	__shared__ float E[kSymPerRound][MB_NC + 2];
	const MbMode &m = a.mode;
	const MbMfsk &t = a.tone;
	const int tid = threadIdx.x, grp = tid >> 4, q = tid & 15;
	const size_t frame = blockIdx.x;
	const uint16_t *__restrict__ dst = reinterpret_cast<const uint16_t *>(a.blob + m.off_llr_dst);
	float *llr = a.llr + frame * MB_HANDOFF_STRIDE;
	float2 w1;
	{
		float s, c;
		sincospif(-2.0f * q / 256.0f, &s, &c);
		w1 = make_float2(c, s);  // W256^q
	}
	const int band_start = t.stream_offsets[0], band_end = t.stream_offsets[t.nStreams - 1] + t.M;
	const int bs = m.nBits / 10, nb = 10;
	float2 *buf = scr[grp];
	for (int s0 = 0; s0 < m.Nsymb; s0 += kSymPerRound) {
		const int s = s0 + grp;
		if (s0 >= a.active_nsymb) {  // control frame (set_mfsk_ctrl_mode): the punctured tail of the codeword is erased (telecom_system.cc:1188-1197)
			if (s < m.Nsymb && q < m.bps) {
				const int i = s * m.bps + q;
				llr[dst[i]] = 0.f;
				if (a.llr_cw) a.llr_cw[frame * MB_N + (i < nb * bs ? (i % nb) * bs + i / nb : i)] = 0.f;
			}
			continue;
		}
		const bool active = s < a.active_nsymb;
		float2 v[16];
		if (active) {
			const float2 *xs = a.x + (frame * m.Nsymb + (size_t)s) * a.sym_stride + a.sym_skip + q;
#pragma unroll
			for (int n1 = 0; n1 < 16; n1++) v[n1] = mbfft::ld_stream(xs + 16 * n1);  // x[16 n1 + q], guard interval skipped
			float2 A[16];
			mbfft::fft16(v, A);
			buf[q * 17] = A[0];
			float2 wk = w1;
#pragma unroll
			for (int k1 = 1; k1 < 16; k1++) {
				buf[q * 17 + k1] = mbfft::cmul(A[k1], wk);
				if (k1 < 15) wk = mbfft::cmul(wk, w1);
			}
		}
		__syncwarp();
		if (active) {
			float2 X0, X1, X14, X15;
#pragma unroll
			for (int n2 = 0; n2 < 16; n2++) v[n2] = buf[n2 * 17 + q];
			mbfft::fft16_pruned(v, X0, X1, X14, X15);  // bins q, 16 + q, 224 + q, 240 + q
			// zero_depadder (ofdm.cc:401-411): bins 231..255 -> carriers 0..24, bins 1..25 -> carriers 25..49; fft() scales by 1/N
			auto emit = [&](const float2 X, const int c) { E[grp][c] = mbfft::cnorm2(X) * (1.0f / 65536.0f); };
			if (q >= 1) emit(X0, 24 + q);
			if (q <= 9) emit(X1, 40 + q);
			if (q >= 7) emit(X14, q - 7);
			emit(X15, 9 + q);
		}
		__syncwarp();
		// noise variance from the carriers outside the tone bands (mfsk.cc:318-338): 16 lanes per symbol
		float ns = 0.f;
		int nbins = 0;
		if (active)
			for (int c = q; c < MB_NC; c += 16)
				if (c < band_start || c >= band_end) {
					const float e = E[grp][c];
					if (isfinite(e)) ns += e, nbins++;
				}
		for (int o = 8; o; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o), nbins += __shfl_xor_sync(0xffffffffu, nbins, o);
		float nv = nbins > 0 ? ns / nbins : 1e-30f;
		if (nv < 1e-30f) nv = 1e-30f;
		const float scale = 1.0f / (2.0f * nv);
		if (!active && s < m.Nsymb && q < m.bps) {  // erased symbols inside the last active round
			const int i = s * m.bps + q;
			llr[dst[i]] = 0.f;
			if (a.llr_cw) a.llr_cw[frame * MB_N + (i < nb * bs ? (i % nb) * bs + i / nb : i)] = 0.f;
		}
		if (active && q < m.bps) {  // one LLR per lane: stream st, bit k (mfsk.cc:341-387)
			const int st = q / t.nBits, k = q % t.nBits, mask = 1 << (t.nBits - 1 - k);
			const int hop = (s * t.tone_hop_step) % t.M;
			float m1 = -1e30f, m0 = -1e30f;
			for (int j = 0; j < t.M; j++) {
				float e = E[grp][t.stream_offsets[st] + ((j + hop) & (t.M - 1))];
				if (!isfinite(e)) e = 0.f;
				if ((j ^ (j >> 1)) & mask) m1 = fmaxf(m1, e);
				else m0 = fmaxf(m0, e);
			}
			float l = (m0 - m1) * scale;
			if (!isfinite(l)) l = 0.f;
			else l = fminf(5.f, fmaxf(-5.f, l));
			const int i = s * m.bps + q;
			llr[dst[i]] = l;
			if (a.llr_cw) a.llr_cw[frame * MB_N + (i < nb * bs ? (i % nb) * bs + i / nb : i)] = l;
		}
		__syncwarp();
	}
	if (tid == 0) {
		MbRxStats st;
		st.iterations_done = 0, st.crc = 0, st.all_zeros = 0, st.message_decoded = 0;
		st.SNR = 0.0f;       // telecom_system.cc:1362-1367: MFSK reports SNR 0 for a decoded frame
		st.variance = 0.0f;
		st.mean_H = 1.0f;    // no channel estimate, no mean|H| gate in the MFSK branch
		st.reserved = 0;
		a.stats[frame] = st;
	}
}

// energies[b][s][c] = |DFT256(bbi[s * 1088 + 64 + 4 i])[bin(c)] / 256|^2, fp64
template <typename T2>
__global__ void __launch_bounds__(256) k_mfsk_energies(const T2 *__restrict__ bbi_all, int n_samples, int nsymb, double *__restrict__ en_all)
{
	__shared__ double2 xs[MB_NFFT];
	__shared__ double2 W[MB_NFFT];
	const int b = blockIdx.y, s = blockIdx.x, tid = threadIdx.x;
	const T2 *bbi = bbi_all + (size_t)b * n_samples + (size_t)s * MB_FE_SYM + MB_NGI * 4;
	{
		double sn, cs;
		sincospi(-2.0 * tid / 256.0, &sn, &cs);
		W[tid] = make_double2(cs, sn);
		const T2 v = bbi[4 * tid];
		xs[tid] = make_double2((double)v.x, (double)v.y);
	}
	__syncthreads();
	if (tid < MB_NC) {
		const int bin = carrier_bin(tid);
		double ar = 0, ai = 0;
		for (int n = 0; n < MB_NFFT; n++) {
			const double2 w = W[(bin * n) & 255], v = xs[n];
			ar += v.x * w.x - v.y * w.y;
			ai += v.x * w.y + v.y * w.x;
		}
		ar /= 256.0, ai /= 256.0;
		en_all[((size_t)b * nsymb + s) * MB_NC + tid] = ar * ar + ai * ai;
	}
}

// one CTA per buffer: time_sync_mfsk + detect_ack_pattern (ACK and BREAK) over the per-symbol carrier energies
__global__ void __launch_bounds__(256) k_mfsk_patterns(const double *__restrict__ en_all, int nsymb, int n_samples, int search_start_common,
							 const int32_t *__restrict__ search_start_each, int each_stride, const MbMfsk t, int pre,
							 MbMfskPatternResult *__restrict__ out)
{
	__shared__ double red_v[256];
	__shared__ int red_i[256], red_m[256];
	const int b = blockIdx.x, tid = threadIdx.x;
	const int search_start_symb = search_start_each ? search_start_each[(size_t)b * each_stride] : search_start_common;
	const double *en = en_all + (size_t)b * nsymb * MB_NC;
	auto e_total = [&](int s) {
		double tot = 0;
		for (int k = 0; k < MB_NC; k++) tot += en[(size_t)s * MB_NC + k];
		return tot;
	};
	// ---- time_sync_mfsk: metric(s) = sum over the preamble symbols of (energy at the expected tones) / (energy of all carriers) ----
	{
		double best = -1;
		int bi = 0x7fffffff;
		for (int s = (search_start_symb > 0 ? search_start_symb : 0) + tid; s <= nsymb - pre; s += 256) {
			double metric = 0;
			for (int p = 0; p < pre; p++) {
				double et = 0;
				for (int st = 0; st < t.nStreams; st++) et += en[(size_t)(s + p) * MB_NC + t.stream_offsets[st] + t.preamble_tones[p % pre]];
				const double tot = e_total(s + p);
				if (tot > 0) metric += et / tot;
			}
			if (metric > best) best = metric, bi = s;
		}
		red_v[tid] = best, red_i[tid] = bi;
		__syncthreads();
		if (tid == 0) {
			double bv = -1;
			int bs_ = 0;
			bool any = false;
			for (int i = 0; i < 256; i++)
				if (red_i[i] != 0x7fffffff && (red_v[i] > bv || (red_v[i] == bv && any && red_i[i] < bs_))) bv = red_v[i], bs_ = red_i[i], any = true;
			out[b].time_sync_delay = bs_ * MB_FE_SYM;
		}
		__syncthreads();
	}
	// ---- detect_ack_pattern for the ACK tones (which = 0) and the BREAK tones (which = 1) ----
	for (int which = 0; which < 2; which++) {
		const int *tones = which ? t.break_tones : t.ack_tones;
		double best = 0.0;
		int bi = 0x7fffffff, bm = 0;
		if (nsymb >= 16)
			for (int s = tid; s <= nsymb - 16; s += 256) {
				double metric = 0;
				int matched = 0;
				for (int p = 0; p < 16; p++) {
					const double *e = en + (size_t)(s + p) * MB_NC;
					const int actual = (tones[p % 8] + p * t.tone_hop_step) % t.M;
					bool any_peak = false;
					double et = 0;
					for (int st = 0; st < t.nStreams; st++) {
						const double ex = e[t.stream_offsets[st] + actual];
						et += ex;
						double peak = -1.0;
						for (int q = 0; q < t.M; q++) peak = fmax(peak, e[t.stream_offsets[st] + q]);
						if (ex >= peak) any_peak = true;
					}
					if (!any_peak) continue;
					matched++;
					const double tot = e_total(s + p);
					if (tot > 0) metric += et / tot;
				}
				if (metric > best) best = metric, bi = s, bm = matched;
			}
		red_v[tid] = best, red_i[tid] = bi, red_m[tid] = bm;
		__syncthreads();
		if (tid == 0) {
			double bv = 0.0;
			int bs_ = 0x7fffffff, bmm = 0;
			for (int i = 0; i < 256; i++)
				if (red_i[i] != 0x7fffffff && (red_v[i] > bv || (red_v[i] == bv && red_i[i] < bs_))) bv = red_v[i], bs_ = red_i[i], bmm = red_m[i];
			if (which == 0) out[b].ack_metric = bv, out[b].ack_matched = bmm;
			else out[b].break_metric = bv, out[b].break_matched = bmm;
		}
		__syncthreads();
	}
	(void)n_samples;
}

}  // namespace

cudaError_t mb_launch_mfsk_demod(const MbMfskArgs &a, size_t n_frames, cudaStream_t s)
{
	k_mfsk_demod<<<(unsigned)n_frames, 16 * kSymPerRound, 0, s>>>(a);
	return cudaGetLastError();
}

cudaError_t mb_launch_mfsk_patterns(const void *d_bbi, int is_f32, size_t n_buffers, int n_samples, int search_start_symb, const int32_t *d_search_start_each,
				    int each_stride, const MbMfsk &t, int pre, double *d_energies, MbMfskPatternResult *d_out, cudaStream_t s)
{
	const int nsymb = n_samples / MB_FE_SYM;
	if (nsymb <= 0) return cudaErrorInvalidValue;
	const dim3 grid(nsymb, (unsigned)n_buffers);
	if (is_f32) k_mfsk_energies<float2><<<grid, 256, 0, s>>>(static_cast<const float2 *>(d_bbi), n_samples, nsymb, d_energies);
	else k_mfsk_energies<double2><<<grid, 256, 0, s>>>(static_cast<const double2 *>(d_bbi), n_samples, nsymb, d_energies);
	k_mfsk_patterns<<<(unsigned)n_buffers, 256, 0, s>>>(d_energies, nsymb, n_samples, search_start_symb, d_search_start_each, each_stride, t, pre, d_out);
	return cudaGetLastError();
}

namespace {

// ---- the MFSK branch of receive_byte() around the kernels above (telecom_system.cc:646-716, 928-943, 1020-1031, 1343-1367) ----
// after the tone-preamble sync: frame-completeness check, bounds, clamp; hands the capture to k_fe_extract_tiles + the MFSK tail
__global__ void k_mfsk_rx_decide(const MbMfskPatternResult *__restrict__ pat, const double *__restrict__ energy_part, int nblk, int buf, int pre, int S,
				 int S_active, int buffer_Nsymb, double fc, MbFeState *__restrict__ st_all, MbReceiveStats *__restrict__ stats, int n, int *__restrict__ counters)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n) return;
	MbFeState st;
	memset(&st, 0, sizeof(st));
	MbReceiveStats r = stats[b];
	double e = 0;
	for (int i = 0; i < nblk; i++) e += energy_part[(size_t)b * nblk + i];
	r.signal_stregth_dbm = 10.0 * log10((e / buf) / 0.001);
	r.iterations_done = 0, r.sync_trials = 0, r.message_decoded = 0, r.crc = 0, r.all_zeros = 0, r.SNR = 0, r.freq_offset = 0, r.coarse_metric = 0;
	int delay = pat[b].time_sync_delay;
	if (r.mfsk_search_or_overflow >= MB_MFSK_FIXED_DELAY_FLAG) {
		// mfsk_fixed_delay >= 0 (:663-673): the caller knows where the preamble is (overflow recapture); the search result is not used
		delay = r.mfsk_search_or_overflow - MB_MFSK_FIXED_DELAY_FLAG;
		r.signal_stregth_dbm = 0;
	}
	int pream = delay / MB_FE_SYM;
	if (pream < 1) pream = 1;
	r.mfsk_search_or_overflow = 0;
	st.slot = -1;
	const int frame_end = delay + (pre + S_active) * MB_FE_SYM;  // get_active_nsymb(): control frames are shorter
	if (frame_end > buf) {
		r.mfsk_search_or_overflow = (frame_end - buf + MB_FE_SYM - 1) / MB_FE_SYM;  // frame_overflow_symbols (:702-715)
	} else if (pream > pre && pream < buffer_Nsymb - (S + pre)) {
		if (delay < 0) delay = 0;
		const int max_delay = buf - (MB_NOFDM * (S + pre)) * 4;
		if (delay > max_delay) delay = max_delay;
		st.extract_pending = 1, st.cur_kind = 1, st.cur_f = fc;
		st.slot = atomicAdd(&counters[0], 1);
	}
	st.delay = delay;
	r.delay = delay;
	st_all[b] = st;
	stats[b] = r;
}

__global__ void k_mfsk_rx_finish(const MbFeState *__restrict__ st_all, const MbRxStats *__restrict__ tail_stats, const uint8_t *__restrict__ tail_payload,
				 int frame_bytes, uint8_t *__restrict__ payload_out, MbReceiveStats *__restrict__ stats, int n)
{
	const int b = blockIdx.x;
	if (b >= n) return;
	const MbFeState &st = st_all[b];
	if (st.slot < 0) return;
	const MbRxStats ts = tail_stats[st.slot];
	for (int i = threadIdx.x; i < frame_bytes; i += blockDim.x) payload_out[(size_t)b * frame_bytes + i] = tail_payload[(size_t)st.slot * frame_bytes + i];
	if (threadIdx.x == 0) {
		MbReceiveStats r = stats[b];
		r.iterations_done = ts.iterations_done, r.crc = ts.crc, r.all_zeros = ts.all_zeros;
		if (ts.message_decoded) {
			r.SNR = 0.0, r.message_decoded = 1;  // :1362-1367
			r.delay_of_last_decoded_message = st.delay;  // :1427 (the frequency-offset fields are not touched in MFSK modes, :1421)
		} else {
			r.SNR = -99.9, r.message_decoded = 0, r.sync_trials = 1;  // :1343-1359; the trial loop ends after one trial (:938-943)
		}
		stats[b] = r;
	}
}

}  // namespace

cudaError_t mb_launch_mfsk_rx_decide(const MbMfskPatternResult *pat, const double *energy_part, int nblk, int buf, int pre, int S, int S_active, int buffer_Nsymb,
				     double fc, MbFeState *st, MbReceiveStats *stats, int n, int *counters, cudaStream_t s)
{
	k_mfsk_rx_decide<<<(n + 127) / 128, 128, 0, s>>>(pat, energy_part, nblk, buf, pre, S, S_active, buffer_Nsymb, fc, st, stats, n, counters);
	return cudaGetLastError();
}

cudaError_t mb_launch_mfsk_rx_finish(const MbFeState *st, const MbRxStats *tail_stats, const uint8_t *tail_payload, int frame_bytes, uint8_t *payload_out,
				     MbReceiveStats *stats, int n, cudaStream_t s)
{
	k_mfsk_rx_finish<<<n, 64, 0, s>>>(st, tail_stats, tail_payload, frame_bytes, payload_out, stats, n);
	return cudaGetLastError();
}
