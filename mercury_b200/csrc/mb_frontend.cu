// mb_frontend.cu -- the RX front-end on the GPU: everything cl_telecom_system::receive_byte() does to a
// pass-band capture buffer before and around the RX tail (reference source/physical_layer/telecom_system.cc:646-1131 and
// 1343-1518, OFDM branch):
//   passband_to_baseband        ofdm.cc:2316-2339   mix with the carrier, 33-tap zero-phase FIR (fir_filter.cc:164-187)
//   measure_signal_stregth      ofdm.cc:1523-1539
//   time_sync_preamble(_with_metric)  ofdm.cc:1735-1967   Schmidl-Cox self-correlation, coarse (step 100) and fine (step 1)
//   energy / metric gates, bounds recovery, silence skip   telecom_system.cc:734-924
//   trial loop, last-good fallbacks, post-fine-sync energy fix   telecom_system.cc:928-1069
//   data-filter mix + decimation by 4   telecom_system.cc:1081-1103, ofdm.cc:2267-2277
//   carrier_sampling_frequency_sync (Moose)   ofdm.cc:540-595, re-mix at the corrected carrier   telecom_system.cc:1126-1131
//   verdict bookkeeping, SKIP-H recovery   telecom_system.cc:1343-1504
//
// Numerics.  Every value that feeds an INTEGER decision of the reference (the sync delay) is computed in fp64 with the
// reference's own operation order and without FMA contraction (__dmul_rn/__dadd_rn), from a carrier table and FIR designs
// computed on the host by the same libm calls the reference makes: the correlation metrics, hence the chosen delays, are
// bit-identical to the reference (also in the tie cases a silent capture produces).  The data path behind the decision
// (mix at the Moose-corrected carrier, FIR, decimation) is fp64 with device sincos and is rounded to the tail's complex64.
//
// Control flow is a per-capture state machine in global memory advanced by k_fe_decide; the heavy stages are separate
// kernels that act on whatever each capture is waiting for (MbFeState::phase), so a batch needs no host-side branching:
//   k_fe_p2b_full, k_fe_prefix -> loop { k_fe_window, k_fe_prefix, k_fe_sc_approx, k_fe_sc_exact, k_fe_decide, k_fe_moose, k_fe_extract_tiles, tail (demod + LDPC
//   kernels) } until all captures are done.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cmath>
#include <cstring>

#include "mb_kernels.cuh"

namespace {

__constant__ MbFeConst fe_c;

constexpr int kDecideThreads = 256;
constexpr int kScThreads = 128;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// The capture formats of the reference's audio layer and their conversion to its double samples (source/audioio/audioio.c:893-940):
// FLOAT32 (double)x, INT16 x / 32768.0, INT32 x / (double)INT_MAX -- all reproduced exactly (the last by an IEEE division).
__device__ __forceinline__ double sample_to_double(double x) { return x; }
__device__ __forceinline__ double sample_to_double(float x) { return (double)x; }
__device__ __forceinline__ double sample_to_double(int16_t x) { return (double)x * (1.0 / 32768.0); }
__device__ __forceinline__ double sample_to_double(int32_t x) { return __ddiv_rn((double)x, 2147483647.0); }

// One mixed sample l[i] = (x[i] * amp) * (cos, sin)(2 pi f i Ts)   (ofdm.cc:2330-2334)
template <typename T>
__device__ __forceinline__ double2 mixed_sample(const T *x, int i, int buf, const double2 *carrier, bool table, double f)
{
	if (i < 0 || i >= buf) return make_double2(0.0, 0.0);
	const double v = dmul(sample_to_double(x[i]), fe_c.amp);
	double c, s;
	if (table) {
		const double2 cs = carrier[i];
		c = cs.x, s = cs.y;
	} else {
		const double ph = dmul(dmul(dmul(2 * M_PI, f), (double)i), fe_c.Ts);
		sincos(ph, &s, &c);
	}
	return make_double2(dmul(v, c), dmul(v, s));
}

// FIR output sample o of the zero-phase filter (fir_filter.cc:164-187): sum over j ascending of l[o + 16 - j] * c[j]; samples
// outside the buffer contribute an exact zero, which leaves the accumulator unchanged like the reference's skipped terms.
template <typename T>
__device__ __forceinline__ double2 fir_on_demand(const T *x, int o, int buf, const double2 *carrier, bool table, double f, const double *coef)
{
	double ar = 0, ai = 0;
#pragma unroll 1
	for (int j = 0; j < MB_FE_TAPS; j++) {
		const double2 l = mixed_sample(x, o + MB_FE_TAPS / 2 - j, buf, carrier, table, f);
		ar = dadd(ar, dmul(l.x, coef[j]));
		ai = dadd(ai, dmul(l.y, coef[j]));
	}
	return make_double2(ar, ai);
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	double t = 0;
	for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
	return t;
}

// ---- full-buffer mix + time-sync FIR (telecom_system.cc:676) and the energy partials of measure_signal_stregth ----
// A CTA owns 1024 consecutive outputs; each thread runs the 33 taps of FOUR neighbouring outputs off a 4-deep sliding register
// window (9 shared-memory reads per output instead of 33: the kernel is bound by the fp64 pipe, not by shared memory).  The
// staging array is skewed by one element per four so that the stride-4 accesses of a quarter warp fall in distinct banks.
constexpr int kP2bTile = 1024;
__device__ __forceinline__ int p2b_skew(int i) { return i + (i >> 2); }

template <typename T, bool DATA_FILTER = false>
__global__ void __launch_bounds__(256, 2) k_fe_p2b_full(const T *__restrict__ x_all, int buf, const double2 *__restrict__ carrier, double2 *__restrict__ bbi_all,
						       double *__restrict__ energy_part, int nblk)
{
	__shared__ double2 l[(kP2bTile + MB_FE_TAPS - 1) * 5 / 4 + 2];
	__shared__ double red[8];
	const int b = blockIdx.y, t0 = blockIdx.x * kP2bTile;
	const T *x = x_all + (size_t)b * buf;
	for (int i = threadIdx.x; i < kP2bTile + MB_FE_TAPS - 1; i += 256) l[p2b_skew(i)] = mixed_sample(x, t0 - MB_FE_TAPS / 2 + i, buf, carrier, true, 0.0);
	__syncthreads();
	const int o = t0 + 4 * threadIdx.x;
	double e = 0;
	if (o < buf) {  // buf is a multiple of 4
		// output o + r, tap j uses l[o + r + 16 - j] = staged index 4 t + r + 32 - j
		double ar[4] = {0, 0, 0, 0}, ai[4] = {0, 0, 0, 0};
		const int base = 5 * threadIdx.x;  // p2b_skew(4 t + c) = 5 t + c + (c >> 2)
		double2 w[4];
#pragma unroll
		for (int r = 1; r < 4; r++) w[r] = l[base + (32 + r) + ((32 + r) >> 2)];
#pragma unroll
		for (int j = 0; j < MB_FE_TAPS; j++) {
			w[0] = l[base + (32 - j) + ((32 - j) >> 2)];
			const double cj = DATA_FILTER ? fe_c.c_data[j] : fe_c.c_ts[j];
#pragma unroll
			for (int r = 0; r < 4; r++) {
				ar[r] = dadd(ar[r], dmul(w[r].x, cj));
				ai[r] = dadd(ai[r], dmul(w[r].y, cj));
			}
			w[3] = w[2], w[2] = w[1], w[1] = w[0];
		}
		double2 *out = bbi_all + (size_t)b * buf + o;
#pragma unroll
		for (int r = 0; r < 4; r++) {
			out[r] = make_double2(ar[r], ai[r]);
			e += ar[r] * ar[r] + ai[r] * ai[r];
		}
	}
	e = block_sum(e, red);
	if (threadIdx.x == 0) energy_part[(size_t)b * nblk + blockIdx.x] = e;
}

// ---- fine-sync window of the data-filter base-band (what baseband_data_interpolated holds after trial 0: telecom_system.cc:1081,1129) ----
template <typename T>
__global__ void __launch_bounds__(256) k_fe_window(const T *__restrict__ x_all, int buf, const double2 *__restrict__ carrier, const MbFeState *__restrict__ st_all,
						     double2 *__restrict__ win_all, int win_stride)
{
	const int b = blockIdx.y;
	const MbFeState &st = st_all[b];
	if (!st.sc_pending || st.sc_src == 0) return;
	const T *x = x_all + (size_t)b * buf;
	const double f = st.sc_src == 1 ? st.cur_f : st.win_f;           // 2: the time-sync filter at a carrier of the coarse frequency search
	const double *coef = st.sc_src == 1 ? fe_c.c_data : fe_c.c_ts;
	const bool table = f == fe_c.fc;
	for (int i = blockIdx.x * 256 + threadIdx.x; i < st.sc_size; i += gridDim.x * 256)
		win_all[(size_t)b * win_stride + i] = fir_on_demand(x, st.sc_start + i, buf, carrier, table, f, coef);
}

// ---- Schmidl-Cox metric (ofdm.cc:1891-1940) in two passes -------------------------------------------------------------
// The reference's metric at a position is three running sums over 576 * pre sample pairs in a fixed order; only the ARGMAX over
// the positions reaches the output (the delay), and it has to be the reference's own, ties included.  So:
//   pass A  every position from exclusive prefix sums of |w|^2 and of the two lagged dot products (any summation order,
//           absolute error ~1e-11): an approximate metric, and the capture's approximate maximum;
//   pass B  the positions within kScTol of that maximum, and those whose norm sits on the 0.001 threshold, are re-evaluated
//           with the reference's operation order in fp64 without FMA -> bit-identical values where it matters.
// A position outside that band cannot hold the exact maximum (error bound << kScTol), and one whose norms are clearly below
// the threshold is exactly 0 in the reference too, so the selection over {exact band values, approximate rest} returns the
// reference's index.  Pass B is typically 1-3 positions per run (hundreds when a clean frame sits in exact silence).
constexpr double kScTol = 1e-6;
constexpr double kScThrBand = 1e-8;
constexpr int kPrefThreads = 1024;

struct Pref3 {
	double e, p1, p2;
};

__device__ __forceinline__ Pref3 pref_terms(const double2 *w, int m, int len)
{
	Pref3 t;
	const double2 u = w[m];
	t.e = u.x * u.x + u.y * u.y;
	t.p1 = t.p2 = 0.0;
	if (m + MB_NFFT * 4 < len) {
		const double2 v = w[m + MB_NFFT * 4];
		t.p1 = u.x * v.x + u.y * v.y;
	}
	if (m + (MB_NFFT / 2) * 4 < len) {
		const double2 v = w[m + (MB_NFFT / 2) * 4];
		t.p2 = u.x * v.x + u.y * v.y;
	}
	return t;
}

// Exclusive prefix sums CE, C1, C2 of one source per capture, one entry every RES samples (entry r = sum over samples < RES * r):
//   RES 4, once per capture: the whole time-sync base-band, for the coarse runs (their positions and every segment edge are
//          multiples of 4: step 100, symbol 1088, GI 64, half symbol 512);
//   RES 1, per fine run: the (pre + 4)-symbol search window, of the time-sync base-band (trial 0) or of the data-filter window.
// A CTA walks its capture in coalesced tiles of 1024 * RES samples with a running carry.
template <int RES>
__global__ void __launch_bounds__(kPrefThreads) k_fe_prefix(const MbFeState *__restrict__ st_all, const double2 *__restrict__ bbi_all, int buf,
							      const double2 *__restrict__ win_all, int win_stride, double *__restrict__ pref_all, size_t pstride)
{
	__shared__ Pref3 wsum[kPrefThreads / 32];
	const int b = blockIdx.x;
	const double2 *w = bbi_all + (size_t)b * buf;
	int len = buf;
	if (RES == 1) {
		const MbFeState &st = st_all[b];
		if (!st.sc_pending || (st.sc_step != 1 && st.sc_src == 0)) return;  // fine runs, and every run over a window
		len = st.sc_size;
		w = st.sc_src == 0 ? w + st.sc_start : win_all + (size_t)b * win_stride;
	}
	double *CE = pref_all + (size_t)b * 3 * pstride, *C1 = CE + pstride, *C2 = C1 + pstride;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	Pref3 carry = {0.0, 0.0, 0.0};
	for (int t0 = 0; t0 < len; t0 += kPrefThreads * RES) {
		const int m = t0 + threadIdx.x * RES;
		Pref3 acc = {0.0, 0.0, 0.0};
#pragma unroll
		for (int r = 0; r < RES; r++)
			if (m + r < len) {
				const Pref3 t = pref_terms(w, m + r, len);
				acc.e += t.e, acc.p1 += t.p1, acc.p2 += t.p2;
			}
		Pref3 inc = acc;
		for (int o = 1; o < 32; o <<= 1) {
			const double e = __shfl_up_sync(0xffffffffu, inc.e, o), p1 = __shfl_up_sync(0xffffffffu, inc.p1, o), p2 = __shfl_up_sync(0xffffffffu, inc.p2, o);
			if (lane >= o) inc.e += e, inc.p1 += p1, inc.p2 += p2;
		}
		__syncthreads();  // wsum of the previous tile has been consumed
		if (lane == 31) wsum[warp] = inc;
		__syncthreads();
		if (warp == 0) {
			Pref3 v = wsum[lane];
			for (int o = 1; o < 32; o <<= 1) {
				const double e = __shfl_up_sync(0xffffffffu, v.e, o), p1 = __shfl_up_sync(0xffffffffu, v.p1, o), p2 = __shfl_up_sync(0xffffffffu, v.p2, o);
				if (lane >= o) v.e += e, v.p1 += p1, v.p2 += p2;
			}
			wsum[lane] = v;
		}
		__syncthreads();
		Pref3 ex = {carry.e + (inc.e - acc.e), carry.p1 + (inc.p1 - acc.p1), carry.p2 + (inc.p2 - acc.p2)};
		if (warp > 0) ex.e += wsum[warp - 1].e, ex.p1 += wsum[warp - 1].p1, ex.p2 += wsum[warp - 1].p2;
		if (m <= len) CE[m / RES] = ex.e, C1[m / RES] = ex.p1, C2[m / RES] = ex.p2;
		carry.e += wsum[31].e, carry.p1 += wsum[31].p1, carry.p2 += wsum[31].p2;
	}
	if (len % (kPrefThreads * RES) == 0 && threadIdx.x == 0) CE[len / RES] = carry.e, C1[len / RES] = carry.p1, C2[len / RES] = carry.p2;
}

// The coarse (RES 4) prefix of the whole time-sync base-band, fully parallel: every CTA scans ONE tile of 4096 samples (local
// exclusive prefix, one entry per 4 samples, + the tile's totals); k_fe_prefix4_base turns the totals into per-tile bases.
// A reader adds the two: C(r) = local[r] + base[r >> kTileShift].
constexpr int kTileEntries = 256;   // entries (= threads) per tile: small CTAs, 8 per SM, so that loads of one overlap the scans of others
constexpr int kTileShift = 8;
__global__ void __launch_bounds__(kTileEntries) k_fe_prefix4_tiles(const double2 *__restrict__ bbi_all, int buf, double *__restrict__ pref_all, size_t pstride,
								     double *__restrict__ tile_tot, int ntile)
{
	constexpr int NW = kTileEntries / 32;
	__shared__ Pref3 wsum[NW];
	const int b = blockIdx.y, tile = blockIdx.x;
	const double2 *w = bbi_all + (size_t)b * buf;
	double *CE = pref_all + (size_t)b * 3 * pstride, *C1 = CE + pstride, *C2 = C1 + pstride;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int m = tile * kTileEntries * 4 + threadIdx.x * 4;
	Pref3 acc = {0.0, 0.0, 0.0};
#pragma unroll
	for (int r = 0; r < 4; r++)
		if (m + r < buf) {
			const Pref3 t = pref_terms(w, m + r, buf);
			acc.e += t.e, acc.p1 += t.p1, acc.p2 += t.p2;
		}
	Pref3 inc = acc;
	for (int o = 1; o < 32; o <<= 1) {
		const double e = __shfl_up_sync(0xffffffffu, inc.e, o), p1 = __shfl_up_sync(0xffffffffu, inc.p1, o), p2 = __shfl_up_sync(0xffffffffu, inc.p2, o);
		if (lane >= o) inc.e += e, inc.p1 += p1, inc.p2 += p2;
	}
	if (lane == 31) wsum[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		Pref3 v = wsum[lane < NW ? lane : NW - 1];
		if (lane >= NW) v.e = v.p1 = v.p2 = 0.0;
		for (int o = 1; o < 32; o <<= 1) {
			const double e = __shfl_up_sync(0xffffffffu, v.e, o), p1 = __shfl_up_sync(0xffffffffu, v.p1, o), p2 = __shfl_up_sync(0xffffffffu, v.p2, o);
			if (lane >= o) v.e += e, v.p1 += p1, v.p2 += p2;
		}
		if (lane < NW) wsum[lane] = v;
	}
	__syncthreads();
	Pref3 ex = {inc.e - acc.e, inc.p1 - acc.p1, inc.p2 - acc.p2};
	if (warp > 0) ex.e += wsum[warp - 1].e, ex.p1 += wsum[warp - 1].p1, ex.p2 += wsum[warp - 1].p2;
	if (m <= buf) CE[m >> 2] = ex.e, C1[m >> 2] = ex.p1, C2[m >> 2] = ex.p2;
	if (threadIdx.x == 0) {
		double *tt = tile_tot + ((size_t)b * (ntile + 1) + tile) * 3;
		tt[0] = wsum[NW - 1].e, tt[1] = wsum[NW - 1].p1, tt[2] = wsum[NW - 1].p2;
	}
}

// per capture: exclusive scan of the tile totals -> bases (in place); the terminal entry when the buffer ends on a tile edge
__global__ void k_fe_prefix4_base(double *__restrict__ tile_tot, int ntile, int buf, double *__restrict__ pref_all, size_t pstride, int n)
{  // one warp per capture: every lane scans a contiguous chunk of tiles, a warp scan joins the chunks
	const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (b >= n) return;
	double *tt = tile_tot + (size_t)b * (ntile + 1) * 3;
	const int per = (ntile + 31) / 32, t0 = lane * per, t1 = min(ntile, t0 + per);
	double e = 0, p1 = 0, p2 = 0;
	for (int t = t0; t < t1; t++) e += tt[3 * t], p1 += tt[3 * t + 1], p2 += tt[3 * t + 2];
	double ie = e, i1 = p1, i2 = p2;
	for (int o = 1; o < 32; o <<= 1) {
		const double a = __shfl_up_sync(0xffffffffu, ie, o), c = __shfl_up_sync(0xffffffffu, i1, o), d = __shfl_up_sync(0xffffffffu, i2, o);
		if (lane >= o) ie += a, i1 += c, i2 += d;
	}
	const double te = __shfl_sync(0xffffffffu, ie, 31), tp1 = __shfl_sync(0xffffffffu, i1, 31), tp2 = __shfl_sync(0xffffffffu, i2, 31);
	double re = ie - e, r1 = i1 - p1, r2 = i2 - p2;  // exclusive base of this lane's chunk
	for (int t = t0; t < t1; t++) {
		const double ve = tt[3 * t], v1 = tt[3 * t + 1], v2 = tt[3 * t + 2];
		tt[3 * t] = re, tt[3 * t + 1] = r1, tt[3 * t + 2] = r2;
		re += ve, r1 += v1, r2 += v2;
	}
	if (lane == 0) {
		tt[3 * ntile] = te, tt[3 * ntile + 1] = tp1, tt[3 * ntile + 2] = tp2;
		if ((buf >> 2) % kTileEntries == 0) {
			double *CE = pref_all + (size_t)b * 3 * pstride;
			CE[buf >> 2] = 0, CE[pstride + (buf >> 2)] = 0, CE[2 * pstride + (buf >> 2)] = 0;
		}
	}
}

__device__ __forceinline__ unsigned long long metric_key(double v) { return (unsigned long long)__double_as_longlong(v + 2.0); }  // monotone for v in [-1, 1]

// pass A
__global__ void __launch_bounds__(kScThreads) k_fe_sc_approx(MbFeState *__restrict__ st_all, const double *__restrict__ pref_ts, size_t pstride_ts,
							       const double *__restrict__ pref_win, size_t pstride_win, const double *__restrict__ tile_base, int ntile,
							       double *__restrict__ vals_all, int vals_stride, uint8_t *__restrict__ flags_all, int pre)
{
	const int b = blockIdx.y;
	MbFeState &st = st_all[b];
	if (!st.sc_pending) return;
	const int k = blockIdx.x * kScThreads + threadIdx.x;
	double v = -3.0;
	if (k < st.sc_npos) {
		// coarse runs (step 100) read the tiled RES-4 prefix of the whole time-sync base-band (local + tile base), fine runs (step 1) the
		// RES-1 prefix of their window
		double cc = 0, na = 0, nb = 0;
		if (st.sc_step == 1 || st.sc_src != 0) {
			const double *CE = pref_win + (size_t)b * 3 * pstride_win + (size_t)k * st.sc_step, *C1 = CE + pstride_win, *C2 = C1 + pstride_win;
			for (int l = 0; l < pre; l++) {
				const int o = l * MB_FE_SYM;
				cc += (C1[o + 64] - C1[o]) + (C2[o + 576] - C2[o + 64]);
				na += CE[o + 576] - CE[o];
				nb += (CE[o + 1088] - CE[o + 576]) + (CE[o + 1088] - CE[o + 1024]);
			}
		} else {
			const double *L = pref_ts + (size_t)b * 3 * pstride_ts;
			const double *B = tile_base + (size_t)b * (ntile + 1) * 3;
			const int r0 = (st.sc_start + k * st.sc_step) >> 2;
			auto C = [&](int which, int off) {
				const int r = r0 + (off >> 2);
				return L[(size_t)which * pstride_ts + r] + B[3 * (r >> kTileShift) + which];
			};
			for (int l = 0; l < pre; l++) {
				const int o = l * MB_FE_SYM;
				cc += (C(1, o + 64) - C(1, o)) + (C(2, o + 576) - C(2, o + 64));
				na += C(0, o + 576) - C(0, o);
				nb += (C(0, o + 1088) - C(0, o + 576)) + (C(0, o + 1088) - C(0, o + 1024));
			}
		}
		const bool amb = fabs(na - 0.001) <= kScThrBand || fabs(nb - 0.001) <= kScThrBand;
		v = (na < 0.001 || nb < 0.001) ? 0.0 : cc / sqrt(na * nb);
		vals_all[(size_t)b * vals_stride + k] = v;
		flags_all[(size_t)b * vals_stride + k] = amb ? 1 : 0;
	}
	// the maximum that anchors pass B's band runs over the positions the selection can return (k * step >= location_to_return): with the
	// global maximum at an excluded position the survivors could otherwise fall outside the band and be compared by their approximations
	if (k < st.sc_npos && k * st.sc_step < st.sc_from) v = -3.0;
	for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	if ((threadIdx.x & 31) == 0 && v > -3.0) atomicMax(&st.sc_max_key, metric_key(v));
}

// pass B: the reference's summation order (ofdm.cc:1901-1930).  The three running sums of a position are serial chains of 4,608
// additions; to keep their latency off the memory system a whole warp serves one band position at a time: the lanes fetch 32
// sample pairs (coalesced, the next chunk already in flight) and form the six products of each, the owner lane then adds them
// in the reference's order out of shared memory.
__global__ void __launch_bounds__(kScThreads) k_fe_sc_exact(const MbFeState *__restrict__ st_all, const double2 *__restrict__ bbi_all, int buf,
							      const double2 *__restrict__ win_all, int win_stride, double *__restrict__ vals_all, int vals_stride,
							      const uint8_t *__restrict__ flags_all, int pre, int *__restrict__ counters)
{
	__shared__ double prod[kScThreads / 32][32][6];
	const int b = blockIdx.y;
	const MbFeState &st = st_all[b];
	if (!st.sc_pending) return;
	const int k = blockIdx.x * kScThreads + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	bool cand = false;
	if (k < st.sc_npos) {
		const double approx = vals_all[(size_t)b * vals_stride + k];
		cand = flags_all[(size_t)b * vals_stride + k] || metric_key(approx + kScTol) >= st.sc_max_key;
	}
	unsigned todo = __ballot_sync(0xffffffffu, cand);
	if (todo == 0) return;
	if (lane == 0) atomicAdd(&counters[3], __popc(todo));
	const double2 *src = st.sc_src == 0 ? bbi_all + (size_t)b * buf + st.sc_start : win_all + (size_t)b * win_stride;
	// chunk c of a position: symbol l = c / 18, within it chunks 0-1 = guard interval vs symbol tail (lag 1024), 2-17 = the two halves (lag 512)
	const int nchunks = pre * 18;
	while (todo) {
		const int owner = __ffs(todo) - 1;
		todo &= todo - 1;
		const double2 *in = src + (size_t)(blockIdx.x * kScThreads + warp * 32 + owner) * st.sc_step;
		double cc = 0, na = 0, nb = 0;
		auto chunk_ptr = [&](int c, const double2 *&a, const double2 *&bq) {
			const int l = c / 18, r = c % 18;
			if (r < 2) a = in + l * MB_FE_SYM + 32 * r, bq = a + MB_NFFT * 4;
			else a = in + l * MB_FE_SYM + MB_NGI * 4 + 32 * (r - 2), bq = a + (MB_NFFT / 2) * 4;
		};
		const double2 *pa, *pb;
		chunk_ptr(0, pa, pb);
		double2 u = pa[lane], v = pb[lane];
		for (int c = 0; c < nchunks; c++) {
			const double2 cu = u, cv = v;
			if (c + 1 < nchunks) {
				chunk_ptr(c + 1, pa, pb);
				u = pa[lane], v = pb[lane];
			}
			double *pr = prod[warp][lane];
			pr[0] = dmul(cu.x, cv.x), pr[1] = dmul(cu.x, cu.x), pr[2] = dmul(cv.x, cv.x);
			pr[3] = dmul(cu.y, cv.y), pr[4] = dmul(cu.y, cu.y), pr[5] = dmul(cv.y, cv.y);
			__syncwarp();
			if (lane == owner) {
#pragma unroll 8
				for (int q = 0; q < 32; q++) {
					const double *p = prod[warp][q];
					cc = dadd(cc, p[0]);
					na = dadd(na, p[1]);
					nb = dadd(nb, p[2]);
					cc = dadd(cc, p[3]);
					na = dadd(na, p[4]);
					nb = dadd(nb, p[5]);
				}
			}
			__syncwarp();
		}
		if (lane == owner) {
			if (na < 0.001 || nb < 0.001) cc = 0.0;
			else cc = __ddiv_rn(cc, __dsqrt_rn(dmul(na, nb)));
			vals_all[(size_t)b * vals_stride + k] = cc;
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// k_fe_decide: one CTA per capture advances the state machine as far as it can without another kernel.
// ------------------------------------------------------------------------------------------------------------------
struct DecideCtx {
	MbFeState st;
	int buf, pre, S, buffer_Nsymb, lower, upper;
	int tmp_i;
	double tmp_d;
};

// The reference's partial selection "sort" (ofdm.cc:1946-1958): entry `from` of the result is the FIRST index >= from holding the
// maximum of the correlation array over [from, size), where the array is the computed metrics at multiples of `step` below the
// search limit and zero everywhere else.
__device__ void sc_result(const DecideCtx &c, const double *vals, int from, int *loc_out, double *corr_out, double *red_v, int *red_k)
{
	const MbFeState &st = c.st;
	double best = -CUDART_INF;
	int bk = 0x7fffffff;
	for (int k = threadIdx.x; k < st.sc_npos; k += blockDim.x) {
		if (k * st.sc_step < from) continue;
		const double v = vals[k];
		if (v > best) best = v, bk = k;  // ascending k per thread: strict > keeps the first
	}
	for (int o = 16; o; o >>= 1) {
		const double ov = __shfl_xor_sync(0xffffffffu, best, o);
		const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
		if (ov > best || (ov == best && ok < bk)) best = ov, bk = ok;
	}
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red_v[threadIdx.x >> 5] = best, red_k[threadIdx.x >> 5] = bk;
	__syncthreads();
	best = red_v[0], bk = red_k[0];
	for (int w = 1; w < (int)(blockDim.x >> 5); w++)
		if (red_v[w] > best || (red_v[w] == best && red_k[w] < bk)) best = red_v[w], bk = red_k[w];
	// first index >= from that is NOT a computed position (its array entry is an exact zero)
	int nz;
	if (st.sc_step == 1) nz = from < st.sc_npos ? st.sc_npos : from;
	else nz = ((from % st.sc_step) == 0 && from / st.sc_step < st.sc_npos) ? from + 1 : from;
	const int p = bk == 0x7fffffff ? -1 : bk * st.sc_step;
	int loc;
	double corr;
	if (p >= 0 && best > 0) loc = p, corr = best;
	else if (p >= 0 && best == 0) loc = (nz < st.sc_size && nz < p) ? nz : p, corr = 0;
	else if (nz < st.sc_size) loc = nz, corr = 0;
	else if (p >= 0) loc = p, corr = best;
	else loc = from, corr = 0;
	*loc_out = loc, *corr_out = corr;
}

// mean energy of up to one symbol of the time-sync base-band starting at pos (the gates: telecom_system.cc:741-753 etc.)
__device__ double energy_ts(const double2 *bbi, int pos, int buf, double *red)
{
	double e = 0;
	int cnt = buf - pos;
	cnt = cnt < 0 ? 0 : (cnt > MB_FE_SYM ? MB_FE_SYM : cnt);
	for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
		const double2 v = bbi[pos + i];
		e += v.x * v.x + v.y * v.y;
	}
	e = block_sum(e, red);
	return cnt > 0 ? e / cnt : 0.0;
}

// first symbol s in [s0, s1) whose mean energy exceeds 0.001, or -1 (telecom_system.cc:741-761, 868-887)
__device__ int energy_scan(const double2 *bbi, int s0, int s1, int buf, int *sh_first)
{
	if (threadIdx.x == 0) *sh_first = 0x7fffffff;
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
	for (int s = s0 + warp; s < s1; s += nw) {
		const int pos = s * MB_FE_SYM;
		int cnt = buf - pos;
		cnt = cnt < 0 ? 0 : (cnt > MB_FE_SYM ? MB_FE_SYM : cnt);
		double e = 0;
		for (int i = lane; i < cnt; i += 32) {
			const double2 v = bbi[pos + i];
			e += v.x * v.x + v.y * v.y;
		}
		for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
		e = cnt > 0 ? e / cnt : 0.0;
		if (lane == 0 && e > 0.001) atomicMin(sh_first, s);
		if (*((volatile int *)sh_first) < s) break;  // a lower symbol already qualified
	}
	__syncthreads();
	const int r = *sh_first;
	__syncthreads();
	return r == 0x7fffffff ? -1 : r;
}

// sum of |.|^2 over one symbol at pos of whatever baseband_data_interpolated currently holds (telecom_system.cc:1040-1069)
template <typename T>
__device__ double energy_cur_sum(const DecideCtx &c, const double2 *bbi, const T *x, const double2 *carrier, int pos, int n, double *red)
{
	double e = 0;
	const bool table = c.st.cur_f == fe_c.fc;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		if (pos + i >= c.buf) break;
		const double2 v = c.st.cur_kind == 0 ? bbi[pos + i] : fir_on_demand(x, pos + i, c.buf, carrier, table, c.st.cur_f, c.st.cur_kind == 1 ? fe_c.c_data : fe_c.c_ts);
		e += v.x * v.x + v.y * v.y;
	}
	return block_sum(e, red);
}

__device__ __forceinline__ void request_sc(MbFeState &st, int src, int start, int size, int step, int pre, int from = 0)
{
	st.sc_pending = 1;
	st.sc_from = from;  // positions below it cannot be returned: they stay out of the approximate maximum that defines pass B's band
	st.sc_src = src, st.sc_start = start, st.sc_size = size, st.sc_step = step;
	const int span = size - pre * MB_FE_SYM;
	st.sc_npos = span > 0 ? (span + step - 1) / step : 0;
	st.sc_max_key = 0;
}

template <typename T>
__global__ void __launch_bounds__(kDecideThreads) k_fe_decide(MbFeState *__restrict__ st_all, const T *__restrict__ x_all, int buf, const double2 *__restrict__ carrier,
								const double2 *__restrict__ bbi_all, const double *__restrict__ vals_all, int vals_stride,
								const double *__restrict__ energy_part, int nblk, const MbRxStats *__restrict__ tail_stats,
								const uint8_t *__restrict__ tail_payload, int tail_payload_stride, uint8_t *__restrict__ payload_out,
								int frame_bytes, int pre, int S, int buffer_Nsymb, int coarse_freq_sync, int *__restrict__ counters)
{
	__shared__ DecideCtx c;
	__shared__ double red[8];
	__shared__ double red_v[8];
	__shared__ int red_k[8];
	__shared__ int sh_first;
	const int b = blockIdx.x;
	if (threadIdx.x == 0) {
		c.st = st_all[b];
		c.buf = buf, c.pre = pre, c.S = S, c.buffer_Nsymb = buffer_Nsymb;
		c.lower = pre, c.upper = buffer_Nsymb - (S + pre);
	}
	__syncthreads();
	if (c.st.phase == MB_FE_DONE) return;
	const double2 *bbi = bbi_all + (size_t)b * buf;
	const T *x = x_all + (size_t)b * buf;
	const double *vals = vals_all + (size_t)b * vals_stride;
	const int sym = MB_FE_SYM;
	MbFeState &st = c.st;
	bool wait = false;
	while (!wait) {
		__syncthreads();
		const int phase = st.phase;
		__syncthreads();  // thread 0 may rewrite st.phase as soon as it enters the case
		int loc;
		double corr;
		switch (phase) {
		case MB_FE_COARSE_WAIT: {  // telecom_system.cc:676-697
			double e = 0;
			for (int i = threadIdx.x; i < nblk; i += blockDim.x) e += energy_part[(size_t)b * nblk + i];
			e = block_sum(e, red);
			sc_result(c, vals, 0, &loc, &corr, red_v, red_k);
			__syncthreads();
			if (threadIdx.x == 0) {
				st.signal_dbm = 10.0 * log10((e / buf) / 0.001);
				st.sc_pending = 0;
				st.delay = loc, st.coarse_metric = corr;
				st.pream_symb_loc = loc / sym < 1 ? 1 : loc / sym;
				st.phase = MB_FE_GATE;
			}
			__syncthreads();
			if (!(st.pream_symb_loc > c.lower && st.pream_symb_loc < c.upper)) {  // bounds recovery, :734-798
				const int s0 = energy_scan(bbi, c.lower + 1, c.upper, buf, &sh_first);
				if (threadIdx.x == 0 && s0 >= 0 && buf - s0 * sym > pre * sym) {
					request_sc(st, 0, s0 * sym, buf - s0 * sym, 100, pre);
					st.phase = MB_FE_REC_BOUNDS;
				}
				__syncthreads();
				if (st.phase == MB_FE_REC_BOUNDS) wait = true;
			}
			break;
		}
		case MB_FE_REC_BOUNDS:
		case MB_FE_REC_SILENCE:
		case MB_FE_REC_SKIPH: {  // :763-797, :889-922, :1466-1503
			sc_result(c, vals, 0, &loc, &corr, red_v, red_k);
			const int rd = loc + st.sc_start;
			const int retry_symb = rd / sym < 1 ? 1 : rd / sym;
			const double re = energy_ts(bbi, rd, buf, red);
			__syncthreads();
			if (threadIdx.x == 0) {
				st.sc_pending = 0;
				const bool ok = re >= 0.001 && (phase == MB_FE_REC_SKIPH || corr >= 0.5) && retry_symb > c.lower && retry_symb < c.upper;
				if (ok) st.delay = rd, st.coarse_metric = corr, st.pream_symb_loc = retry_symb;
				if (phase == MB_FE_REC_BOUNDS) st.phase = MB_FE_GATE;
				else if (phase == MB_FE_REC_SILENCE) st.phase = ok ? MB_FE_TRIAL : MB_FE_DONE, st.skip_h_count = 0, st.skip_h_recovery_attempted = 0;
				else {
					if (ok) st.sync_trials = 0, st.skip_h_count = 0, st.coarse_off = 0.0, st.phase = MB_FE_TRIAL;  // (:1493: coarse_freq_offset = 0)
					else st.phase = MB_FE_DONE;
				}
			}
			break;
		}
		case MB_FE_GATE: {  // :800-926
			if (!(st.pream_symb_loc > c.lower && st.pream_symb_loc < c.upper)) {
				if (threadIdx.x == 0) st.phase = MB_FE_DONE;
				break;
			}
			const double mean_energy = energy_ts(bbi, st.delay, buf, red);
			bool energy_ok = !(mean_energy < 0.001);
			if (energy_ok && st.coarse_metric < 0.5) energy_ok = false;
			if (energy_ok) {
				if (threadIdx.x == 0) st.phase = MB_FE_TRIAL, st.skip_h_count = 0, st.skip_h_recovery_attempted = 0;
				break;
			}
			const int s0 = energy_scan(bbi, st.pream_symb_loc + 1, c.upper, buf, &sh_first);
			if (threadIdx.x == 0) {
				if (s0 >= 0 && buf - s0 * sym > pre * sym) {
					request_sc(st, 0, s0 * sym, buf - s0 * sym, 100, pre);
					st.phase = MB_FE_REC_SILENCE;
				} else st.phase = MB_FE_DONE;
			}
			__syncthreads();
			if (st.phase == MB_FE_REC_SILENCE) wait = true;
			break;
		}
		case MB_FE_TRIAL: {  // loop head :931-1019
			if (threadIdx.x == 0) {
				if (st.sync_trials > fe_c.trials_max) st.phase = MB_FE_TRIALS_END;
				else if (st.sync_trials == fe_c.trials_max && fe_c.use_last_time && st.last_delay != -1) {
					st.delay = st.last_delay;
					st.phase = MB_FE_POSTDELAY;
				} else if (st.sync_trials == 1 && coarse_freq_sync) {
					// trial 0 failed: Schmidl-Cox over the head of the buffer with the time-sync filter at fc - 30, fc, fc + 30 Hz (:949-983);
					// the windows are computed on demand (k_fe_window, source 2), the time-sync base-band at fc stays as it is
					st.cfs_i = 0, st.cfs_best = 0.0, st.cfs_best_off = 0.0, st.cfs_zero = 0.0, st.cfs_best_delay = st.delay;
					st.win_f = fe_c.fc + -30.0;
					request_sc(st, 2, 0, MB_NOFDM * (2 * pre + S) * 4, 100, pre);
					st.phase = MB_FE_CFS_WAIT;
				} else {
					request_sc(st, st.cur_kind, (st.pream_symb_loc - 1) * sym, (pre + 4) * sym, 1, pre,
						   st.sync_trials >= fe_c.trials_max ? fe_c.trials_max - 1 : st.sync_trials);  // = `want` of MB_FE_FINE_WAIT
					st.phase = MB_FE_FINE_WAIT;
				}
			}
			__syncthreads();
			if (st.phase == MB_FE_FINE_WAIT || st.phase == MB_FE_CFS_WAIT) wait = true;
			break;
		}
		case MB_FE_CFS_WAIT: {  // one of the three runs has finished (:966-983); after the last: decision (:985-993) and the fine sync (:995-1013)
			sc_result(c, vals, 0, &loc, &corr, red_v, red_k);
			__syncthreads();
			if (threadIdx.x == 0) {
				st.sc_pending = 0;
				const double off = st.cfs_i == 0 ? -30.0 : (st.cfs_i == 1 ? 0.0 : 30.0);
				if (fabs(off) < 0.1) st.cfs_zero = corr;
				if (corr > st.cfs_best) st.cfs_best = corr, st.cfs_best_off = off, st.cfs_best_delay = loc;
				st.cfs_i++;
				if (st.cfs_i < 3) {
					st.win_f = fe_c.fc + (st.cfs_i == 1 ? 0.0 : 30.0);
					request_sc(st, 2, 0, MB_NOFDM * (2 * pre + S) * 4, 100, pre);
				} else {
					if (fabs(st.cfs_best_off) > 1.0 && st.cfs_best > 0.5 && st.cfs_best > st.cfs_zero + 0.1) {
						st.coarse_off = st.cfs_best_off;
						st.delay = st.cfs_best_delay;
						st.pream_symb_loc = st.delay / sym < 1 ? 1 : st.delay / sym;
					}
					// baseband_data_interpolated = the time-sync filter at the (possibly corrected) carrier; fine sync on it
					st.cur_kind = 2, st.cur_f = fe_c.fc + st.coarse_off, st.win_f = st.cur_f;
					request_sc(st, 2, (st.pream_symb_loc - 1) * sym, (pre + 4) * sym, 1, pre,
						   st.sync_trials >= fe_c.trials_max ? fe_c.trials_max - 1 : st.sync_trials);
					st.phase = MB_FE_FINE_WAIT;
				}
			}
			wait = true;
			break;
		}
		case MB_FE_FINE_WAIT: {
			const int want = st.sync_trials >= fe_c.trials_max ? fe_c.trials_max - 1 : st.sync_trials;  // location_to_return clamp, ofdm.cc:1821-1823
			sc_result(c, vals, want, &loc, &corr, red_v, red_k);
			__syncthreads();
			if (threadIdx.x == 0) {
				st.sc_pending = 0;
				st.delay = st.sc_start + loc;
				st.phase = MB_FE_POSTDELAY;
			}
			break;
		}
		case MB_FE_POSTDELAY: {  // :1020-1069
			int delay = st.delay;
			if (delay < 0) delay = 0;
			const int max_delay = buf - (MB_NOFDM * (S + pre)) * 4;
			if (delay > max_delay) delay = max_delay;
			double fe = energy_cur_sum(c, bbi, x, carrier, delay, sym, red) / sym;
			if (fe < 0.001) {
				const int orig = delay;
				for (int fwd = sym; fwd <= 3 * sym; fwd += sym) {
					const int cand = orig + fwd;
					if (cand + sym > buf) break;
					const double e = energy_cur_sum(c, bbi, x, carrier, cand, sym, red) / sym;
					if (e >= 0.001) {
						delay = cand;
						break;
					}
				}
			}
			__syncthreads();
			if (threadIdx.x == 0) {
				st.delay = delay;
				st.slot = atomicAdd(&counters[0], 1);
				st.phase = MB_FE_EXTRACT;
			}
			wait = true;
			break;
		}
		case MB_FE_TAIL_WAIT: {  // verdict :1271-1281, :1310-1429
			const MbRxStats ts = tail_stats[st.slot];
			if (ts.iterations_done >= 0)
				for (int i = threadIdx.x; i < frame_bytes; i += blockDim.x)
					payload_out[(size_t)b * frame_bytes + i] = tail_payload[(size_t)st.slot * tail_payload_stride + i];
			__syncthreads();
			if (threadIdx.x == 0) {
				st.slot = -1;
				st.extract_pending = 0;
				if (ts.iterations_done < 0) {  // mean|H| < 0.3: LDPC skipped
					st.skip_h_count++;
					st.sync_trials++;
					st.phase = MB_FE_TRIAL;
				} else {
					st.iterations_done = ts.iterations_done, st.crc = ts.crc, st.all_zeros = ts.all_zeros;
					if (!ts.message_decoded) {
						st.SNR = -99.9;
						st.message_decoded = 0;
						st.sync_trials++;
						st.phase = MB_FE_TRIAL;
					} else {
						st.SNR = ts.SNR;
						st.message_decoded = 1;
						st.last_freq = st.freq_offset_measured;
						st.freq_offset = st.freq_offset_measured;
						st.last_delay = st.delay;
						st.phase = MB_FE_DONE;
					}
				}
			}
			break;
		}
		case MB_FE_TRIALS_END: {  // SKIP-H recovery :1436-1504
			if (threadIdx.x == 0) {
				st.phase = MB_FE_DONE;
				if (!st.message_decoded && st.skip_h_count >= fe_c.trials_max + 1 && !st.skip_h_recovery_attempted) {
					st.skip_h_recovery_attempted = 1;
					const int ss = st.pream_symb_loc + 2, search_start = ss * sym;
					const int search_size = MB_NOFDM * (2 * pre + S) * 4;
					int available = buf - search_start;
					if (available > search_size) available = search_size;
					if (ss < c.upper && available > pre * sym) {
						st.cur_kind = 0, st.cur_f = fe_c.fc;  // the time-sync base-band is restored (:1457-1461)
						request_sc(st, 0, search_start, available, 100, pre);
						st.phase = MB_FE_REC_SKIPH;
					}
				}
			}
			__syncthreads();
			if (st.phase == MB_FE_REC_SKIPH) wait = true;
			break;
		}
		default:  // MB_FE_DONE, MB_FE_EXTRACT (waiting for k_fe_extract)
			wait = true;
			break;
		}
		__syncthreads();
		if (st.phase == MB_FE_DONE) wait = true;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		st_all[b] = st;
		if (st.phase != MB_FE_DONE) atomicAdd(&counters[1], 1);
		if (st.sc_pending) atomicAdd(&counters[2], 1);
	}
}

// ---- Moose on the data-filter base-band of the preamble (telecom_system.cc:1081-1117, ofdm.cc:540-595): one CTA per capture ----
template <typename T>
__global__ void __launch_bounds__(256) k_fe_moose(MbFeState *__restrict__ st_all, const T *__restrict__ x_all, int buf, const double2 *__restrict__ carrier, int pre)
{
	__shared__ double2 pb[2 * MB_NOFDM];  // the preamble symbols Moose looks at (decimated rate)
	__shared__ double2 W128[128];
	__shared__ double2 G[2][2][24][4];
	__shared__ double2 mul_sh;
	const int b = blockIdx.x, tid = threadIdx.x;
	MbFeState &st = st_all[b];
	if (st.phase != MB_FE_EXTRACT) return;
	const T *x = x_all + (size_t)b * buf;
	const int delay = st.delay;
	const int np2 = pre / 2 == 0 ? 1 : pre / 2;  // ofdm.cc:548-555
	const bool use_last = st.sync_trials == fe_c.trials_max && fe_c.use_last_freq && st.last_freq != 0;  // :1108-1111
	double fm = st.last_freq;
	const double fbase = fe_c.fc + st.coarse_off;  // effective_fc (:1076): fc unless the coarse frequency search moved it
	if (!use_last) {
		if (tid < 128) {
			double sn, cs;
			sincospi(-2.0 * tid / 128.0, &sn, &cs);
			W128[tid] = make_double2(cs, sn);
		}
		for (int k = tid; k < np2 * MB_NOFDM; k += 256) {  // FIR_rx_data at the decimated positions, carrier from the table
			const int o = delay + 4 * k;
			double ar = 0, ai = 0;
#pragma unroll
			for (int j = 0; j < MB_FE_TAPS; j++) {
				const double2 l = mixed_sample(x, o + MB_FE_TAPS / 2 - j, buf, carrier, fbase == fe_c.fc, fbase);
				ar = dadd(ar, dmul(l.x, fe_c.c_data[j]));
				ai = dadd(ai, dmul(l.y, fe_c.c_data[j]));
			}
			pb[k] = make_double2(ar, ai);
		}
		__syncthreads();
		// each half symbol is repeated before the reference's 256-point FFT, so only the even bins carry signal and they equal the
		// 128-point DFT of the half; 24 of the 50 active carriers are even (bins 232..254 and 2..24).  192 threads: (bin, half, quarter).
		if (tid == 0) mul_sh = make_double2(0.0, 0.0);
		for (int j = 0; j < np2; j++) {
			if (tid < 192) {
				const int bin = tid % 24, half = (tid / 24) & 1, quarter = tid / 48;
				const int k128 = bin < 12 ? 116 + bin : bin - 11;
				double gr = 0, gi = 0;
				const double2 *g = pb + j * MB_NOFDM + MB_NGI + half * 128;
				for (int n = quarter * 32; n < quarter * 32 + 32; n++) {
					const double2 w = W128[(k128 * n) & 127];
					gr += g[n].x * w.x - g[n].y * w.y;
					gi += g[n].x * w.y + g[n].y * w.x;
				}
				G[j & 1][half][bin][quarter] = make_double2(gr, gi);
			}
			__syncthreads();
			if (tid == 0) {
				double2 mul = mul_sh;
				for (int q = 0; q < 24; q++) {  // mul += conj(d2) * d1
					double2 d1 = make_double2(0.0, 0.0), d2 = d1;
					for (int r = 0; r < 4; r++) {
						d1.x += G[j & 1][0][q][r].x, d1.y += G[j & 1][0][q][r].y;
						d2.x += G[j & 1][1][q][r].x, d2.y += G[j & 1][1][q][r].y;
					}
					mul.x += d2.x * d1.x + d2.y * d1.y;
					mul.y += d2.x * d1.y - d2.y * d1.x;
				}
				mul_sh = mul;
			}
			__syncthreads();
		}
		const double2 mul = mul_sh;
		double th;  // get_angle, misc.cc:34-56
		if (mul.x == 0) th = M_PI / 2;
		else if (mul.x > 0) th = atan(mul.y / mul.x);
		else if (mul.y >= 0) th = atan(mul.y / mul.x) + M_PI;
		else th = atan(mul.y / mul.x) - M_PI;
		fm = (th / M_PI) * (fe_c.bandwidth / (double)MB_NC);
	}
	__syncthreads();
	if (tid == 0) {
		const bool corrected = fabs(fm) > fe_c.ignore_limit;  // :1126
		st.freq_offset_measured = fm;
		st.cur_kind = 1, st.cur_f = corrected ? fbase + fm : fbase;
		st.extract_pending = 1;
		st.phase = MB_FE_TAIL_WAIT;
	}
}

// ---- data-filter mix at the final carrier + decimation by 4 (:1081-1103,1126-1131): one CTA per tile of 256 decimated outputs ----
template <typename T>
__global__ void __launch_bounds__(256) k_fe_extract_tiles(const MbFeState *__restrict__ st_all, const T *__restrict__ x_all, int buf, const double2 *__restrict__ carrier,
							    float2 *__restrict__ frames, double2 *__restrict__ dbg_bb, int pre, int S)
{
	__shared__ double2 lt[1024 + MB_FE_TAPS - 1];
	__shared__ double2 r256_sh;
	const int b = blockIdx.y, tid = threadIdx.x;
	const MbFeState &st = st_all[b];
	if (!st.extract_pending) return;
	const int k_begin = dbg_bb ? 0 : pre * MB_NOFDM, k_end = (S + pre) * MB_NOFDM;
	const int k0 = k_begin + blockIdx.x * 256;
	if (k0 >= k_end) return;
	const T *x = x_all + (size_t)b * buf;
	const double f = st.cur_f;
	const bool corrected = f != fe_c.fc;
	const int p0 = st.delay + 4 * k0 - MB_FE_TAPS / 2;
	// corrected carrier: one sincos per thread for its first staged sample, then rotations by 256 samples
	double2 cur = make_double2(1.0, 0.0);
	if (corrected) {
		sincos(dmul(dmul(dmul(2 * M_PI, f), (double)(p0 + tid)), fe_c.Ts), &cur.y, &cur.x);
		if (tid == 0) sincos(dmul(dmul(dmul(2 * M_PI, f), 256.0), fe_c.Ts), &r256_sh.y, &r256_sh.x);
		__syncthreads();
	}
	const double2 r256 = corrected ? r256_sh : cur;
	for (int i = tid; i < 1024 + MB_FE_TAPS - 1; i += 256) {
		const int n = p0 + i;
		double2 cs = cur;
		if (corrected) cur = make_double2(cur.x * r256.x - cur.y * r256.y, cur.x * r256.y + cur.y * r256.x);
		else if (n >= 0 && n < buf) cs = carrier[n];
		double2 m = make_double2(0.0, 0.0);
		if (n >= 0 && n < buf) {
			const double v = dmul(sample_to_double(x[n]), fe_c.amp);
			m = make_double2(dmul(v, cs.x), dmul(v, cs.y));
		}
		lt[i] = m;
	}
	__syncthreads();
	const int k = k0 + tid;
	if (k < k_end) {
		double ar = 0, ai = 0;
#pragma unroll
		for (int j = 0; j < MB_FE_TAPS; j++) {
			const double2 l = lt[4 * tid + MB_FE_TAPS - 1 - j];
			ar = dadd(ar, dmul(l.x, fe_c.c_data[j]));
			ai = dadd(ai, dmul(l.y, fe_c.c_data[j]));
		}
		if (k >= pre * MB_NOFDM) frames[(size_t)st.slot * S * MB_NOFDM + k - pre * MB_NOFDM] = make_float2((float)ar, (float)ai);
		if (dbg_bb) dbg_bb[(size_t)b * (S + pre) * MB_NOFDM + k] = make_double2(ar, ai);
	}
}

// receive_byte() entry (:646-663): reset the per-call fields, take the link state, ask for the full-buffer coarse run (:691)
__global__ void k_fe_begin(MbFeState *__restrict__ st_all, const MbReceiveStats *__restrict__ in, int n, int buf, int pre)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n) return;
	MbFeState st;
	memset(&st, 0, sizeof(st));
	st.last_delay = in[b].delay_of_last_decoded_message;
	st.last_freq = in[b].freq_offset_of_last_decoded_message;
	st.phase = MB_FE_COARSE_WAIT;
	st.cur_kind = 0, st.cur_f = fe_c.fc;
	st.slot = -1;
	request_sc(st, 0, 0, buf, 100, pre);
	st_all[b] = st;
}

__global__ void k_fe_finish(const MbFeState *__restrict__ st_all, MbReceiveStats *__restrict__ out, int n)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n) return;
	const MbFeState &st = st_all[b];
	MbReceiveStats r;
	r.iterations_done = st.iterations_done, r.delay = st.delay, r.delay_of_last_decoded_message = st.last_delay, r.sync_trials = st.sync_trials;
	r.message_decoded = st.message_decoded, r.crc = st.crc, r.all_zeros = st.all_zeros, r.mfsk_search_or_overflow = 0;
	r.freq_offset = st.freq_offset, r.freq_offset_of_last_decoded_message = st.last_freq, r.SNR = st.SNR;
	r.signal_stregth_dbm = st.signal_dbm, r.coarse_metric = st.coarse_metric;
	out[b] = r;
}

}  // namespace

// ---- host side: tables and launchers ----

// cl_FIR::design, LPF + HAMMING (fir_filter.cc:45-163) with the receive filters' parameters (physical_config.cc:93-101).
static void fir_design_lpf_hamming(double fcut, double tbw, double fs, double *c)
{
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
	for (int i = 0; i < n; i++) c[i] *= 0.54 - 0.46 * cos(2.0 * M_PI * (double)i / (n - 1));
}

void mb_fe_host_const(MbFeConst *k)
{
	memset(k, 0, sizeof(*k));
	k->fs = 48000.0;                           // physical_config.cc:78
	k->bandwidth = 48000.0 * 50.0 / 256 / 4;   // :81
	k->fc = 0.0 + (k->bandwidth / 2 + 300);    // :88 (carrier_frequency_offset = 0)
	k->amp = sqrt(2.0);                        // telecom_system.cc:69
	k->Ts = 1.0 / k->fs;
	k->ignore_limit = (double)0.1f;            // physical_config.cc:59, stored in a float
	k->trials_max = 2, k->use_last_time = 1, k->use_last_freq = 1;  // :85-87
	fir_design_lpf_hamming(0.9 * k->bandwidth / 2, 3000, k->fs, k->c_ts);
	fir_design_lpf_hamming(1.0 * k->bandwidth / 2, 3000, k->fs, k->c_data);
}

int mb_fe_buffer_nsymb(int Nsymb, int pre)
{  // data_container.cc:133-143
	const double sym_time_ms = 1000.0 * MB_NOFDM * 4 / 48000.0;
	const int turnaround_symb = (int)ceil(1200.0 / sym_time_ms) + 4;
	const int frame_symb = pre + Nsymb;
	int min_buf = frame_symb * 2;
	if (frame_symb + turnaround_symb > min_buf) min_buf = frame_symb + turnaround_symb;
	if (min_buf < 32) min_buf = 32;
	return min_buf;
}

// (cos, sin)(2 pi fc i Ts) by the host's libm, the same calls in the same expression order as ofdm.cc:2332-2333
void mb_fe_host_carrier(const MbFeConst &k, double *cs, int n)
{
	const double sampling_interval = 1.0 / k.fs;
	for (int i = 0; i < n; i++) {
		cs[2 * i] = cos(2 * M_PI * k.fc * (double)i * sampling_interval);
		cs[2 * i + 1] = sin(2 * M_PI * k.fc * (double)i * sampling_interval);
	}
}

cudaError_t mb_fe_init(const MbFeConst &k)
{
	return cudaMemcpyToSymbol(fe_c, &k, sizeof(k));
}

template <typename T>
static cudaError_t fe_p2b_full_t(const MbFeArgs &a, cudaStream_t s)
{
	const int nblk = (a.buf + kP2bTile - 1) / kP2bTile;
	k_fe_p2b_full<T><<<dim3(nblk, a.n), 256, 0, s>>>(static_cast<const T *>(a.x), a.buf, a.carrier, a.bbi, a.energy_part, nblk);
	if (!a.pref_ts) return cudaGetLastError();  // callers that only want the base-band and its energy (measure_signal_only, the MFSK branch)
	const int ntile = (a.buf / 4 + kTileEntries - 1) / kTileEntries;
	k_fe_prefix4_tiles<<<dim3(ntile, a.n), kTileEntries, 0, s>>>(a.bbi, a.buf, a.pref_ts, (size_t)a.buf / 4 + 1, a.tile_base, ntile);
	k_fe_prefix4_base<<<(a.n * 32 + 127) / 128, 128, 0, s>>>(a.tile_base, ntile, a.buf, a.pref_ts, (size_t)a.buf / 4 + 1, a.n);
	return cudaGetLastError();
}

template <typename T>
static cudaError_t fe_step_t(const MbFeArgs &a, bool run_sc, cudaStream_t s)
{
	const int nblk = (a.buf + kP2bTile - 1) / kP2bTile;
	if (run_sc) {
		k_fe_window<T><<<dim3(8, a.n), 256, 0, s>>>(static_cast<const T *>(a.x), a.buf, a.carrier, a.st, a.win, a.win_stride);
		k_fe_prefix<1><<<a.n, kPrefThreads, 0, s>>>(a.st, a.bbi, a.buf, a.win, a.win_stride, a.pref_win, (size_t)a.win_stride + 1);
		const dim3 grid((a.vals_stride + kScThreads - 1) / kScThreads, a.n);
		k_fe_sc_approx<<<grid, kScThreads, 0, s>>>(a.st, a.pref_ts, (size_t)a.buf / 4 + 1, a.pref_win, (size_t)a.win_stride + 1, a.tile_base,
							   (a.buf / 4 + kTileEntries - 1) / kTileEntries, a.vals, a.vals_stride, a.flags, a.pre);
		k_fe_sc_exact<<<grid, kScThreads, 0, s>>>(a.st, a.bbi, a.buf, a.win, a.win_stride, a.vals, a.vals_stride, a.flags, a.pre, a.counters);
	}
	k_fe_decide<T><<<a.n, kDecideThreads, 0, s>>>(a.st, static_cast<const T *>(a.x), a.buf, a.carrier, a.bbi, a.vals, a.vals_stride, a.energy_part, nblk,
							a.tail_stats, a.tail_payload, a.tail_payload_stride, a.payload_out, a.frame_bytes, a.pre, a.S,
							a.buffer_Nsymb, a.coarse_freq_sync, a.counters);
	return cudaGetLastError();
}

template <typename T>
static cudaError_t fe_extract_t(const MbFeArgs &a, cudaStream_t s)
{
	k_fe_moose<T><<<a.n, 256, 0, s>>>(a.st, static_cast<const T *>(a.x), a.buf, a.carrier, a.pre);
	const int outs = (a.dbg_bb ? a.S + a.pre : a.S) * MB_NOFDM;
	k_fe_extract_tiles<T><<<dim3((outs + 255) / 256, a.n), 256, 0, s>>>(a.st, static_cast<const T *>(a.x), a.buf, a.carrier, a.frames, a.dbg_bb, a.pre, a.S);
	return cudaGetLastError();
}

// the data kernel alone (the MFSK branch has no Moose step: the carrier stays at fc)
template <typename T>
static cudaError_t fe_extract_data_t(const MbFeArgs &a, cudaStream_t s)
{
	const int outs = (a.dbg_bb ? a.S + a.pre : a.S) * MB_NOFDM;
	k_fe_extract_tiles<T><<<dim3((outs + 255) / 256, a.n), 256, 0, s>>>(a.st, static_cast<const T *>(a.x), a.buf, a.carrier, a.frames, a.dbg_bb, a.pre, a.S);
	return cudaGetLastError();
}

// passband_to_baseband with FIR_rx_data over whole buffers (detect_ack_pattern_from_passband, telecom_system.cc:1637-1641)
template <typename T>
static cudaError_t fe_p2b_data_t(const MbFeArgs &a, cudaStream_t s)
{
	const int nblk = (a.buf + kP2bTile - 1) / kP2bTile;
	k_fe_p2b_full<T, true><<<dim3(nblk, a.n), 256, 0, s>>>(static_cast<const T *>(a.x), a.buf, a.carrier, a.bbi, a.energy_part, nblk);
	return cudaGetLastError();
}

#define MB_FE_DISPATCH(fn, ...)                                      \
	switch (a.x_format) {                                        \
	case 0: return fn<double>(__VA_ARGS__);                      \
	case 1: return fn<float>(__VA_ARGS__);                       \
	case 2: return fn<int16_t>(__VA_ARGS__);                     \
	case 3: return fn<int32_t>(__VA_ARGS__);                     \
	}                                                            \
	return cudaErrorInvalidValue

cudaError_t mb_fe_p2b_full(const MbFeArgs &a, cudaStream_t s) { MB_FE_DISPATCH(fe_p2b_full_t, a, s); }
cudaError_t mb_fe_step(const MbFeArgs &a, bool run_sc, cudaStream_t s) { MB_FE_DISPATCH(fe_step_t, a, run_sc, s); }
cudaError_t mb_fe_extract(const MbFeArgs &a, cudaStream_t s) { MB_FE_DISPATCH(fe_extract_t, a, s); }
cudaError_t mb_fe_extract_data(const MbFeArgs &a, cudaStream_t s) { MB_FE_DISPATCH(fe_extract_data_t, a, s); }
cudaError_t mb_fe_p2b_data(const MbFeArgs &a, cudaStream_t s) { MB_FE_DISPATCH(fe_p2b_data_t, a, s); }

cudaError_t mb_fe_begin(const MbFeArgs &a, const MbReceiveStats *d_stats_in, cudaStream_t s)
{
	k_fe_begin<<<(a.n + 127) / 128, 128, 0, s>>>(a.st, d_stats_in, a.n, a.buf, a.pre);
	return cudaGetLastError();
}

cudaError_t mb_fe_finish(const MbFeArgs &a, MbReceiveStats *d_stats_out, cudaStream_t s)
{
	k_fe_finish<<<(a.n + 127) / 128, 128, 0, s>>>(a.st, d_stats_out, a.n);
	return cudaGetLastError();
}
