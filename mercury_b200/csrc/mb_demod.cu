// mb_demod.cu -- K_demod: fused OFDM demodulator for sm_100a, one CTA per frame.
//
//   FFT-256 per symbol  ->  AGC  ->  LS / ZF channel estimate  ->  column interpolation  ->  (phase-only)
//   equalise  ->  pilot noise variance  ->  deframe + T/F de-interleave  ->  max-log soft de-map  ->
//   bit de-interleave + LLR expand  ->  LLR[1600]
//
// Reference (what is computed; paths relative to /root/reference/source/physical_layer):
//   symbol_demod            ofdm.cc:862-867 (gi_remover 423-429, fft 431-444, zero_depadder 401-411)
//   automatic_gain_control  ofdm.cc:1467-1498
//   LS_channel_estimator    ofdm.cc:1315-1451     ZF_channel_estimator ofdm.cc:1266-1313
//   interpolate_linear_col  interpolator.cc:163-254
//   restore_channel_amplitude ofdm.cc:1453-1466 (get_angle misc.cc:34-56)
//   channel_equalizer       ofdm.cc:1637-1657     measure_variance ofdm.cc:1500-1521
//   deframer ofdm.cc:837-852, deinterleaver interleaver.cc:77-109, cl_psk::demod psk.cc:278-326,
//   LLR expand telecom_system.cc:1300-1308
//
// How (B200-first, not the reference's loops):
//   * HBM-bound stage: every sample is read exactly once with 8-byte streaming loads (16 independent loads in
//     flight per thread, guard interval never fetched); the only other HBM traffic is the 6.4 KB LLR vector.
//   * FFT-256 = 16 x 16: 16 threads per OFDM symbol, two radix-4x4 16-point DFTs in registers around one
//     conflict-free (stride-17) shared-memory transpose that needs only __syncwarp (a symbol lives in half a warp).
//     Twiddles (with the reference's 1/N folded in) come from a 2 KB shared table laid out for broadcast reads.
//     The second DFT is pruned to the 4 of 16 outputs that land on the 50 active carriers.
//   * The LS estimator's O(pilots x window) double loop is restated as the windowed mean it is:
//     a separable clipped 21x21 box sum over the pilot lattice, row pass then column pass, in shared memory.
//   * deframe, both de-interleavers and the LLR expand are composed on the host into two gather/scatter index
//     tables (mb_tables.cpp), so no intermediate vector is ever materialised.
#include "mb_kernels.cuh"

namespace {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)
__device__ __forceinline__ float2 cdiv(float2 a, float2 b)
{
	float inv = 1.0f / (b.x * b.x + b.y * b.y);
	return make_float2((a.x * b.x + a.y * b.y) * inv, (a.y * b.x - a.x * b.y) * inv);
}

// forward 4-point DFT (W4 = -i)
__device__ __forceinline__ void dft4(float2 a, float2 b, float2 c, float2 d, float2 &y0, float2 &y1, float2 &y2, float2 &y3)
{
	float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
	y0 = cadd(t0, t2);
	y2 = csub(t0, t2);
	y1 = cadd(t1, mul_mi(t3));
	y3 = cadd(t1, mul_pi(t3));
}

#define MB_C1 0.92387953251128674f
#define MB_S1 0.38268343236508977f
#define MB_R2 0.70710678118654752f

// Stage 1+2 of the radix-4x4 16-point forward DFT: A[n2][k1] = W16^(n2 k1) * sum_n1 x[4 n1 + n2] (-i)^(n1 k1)
__device__ __forceinline__ void fft16_front(const float2 (&x)[16], float2 (&A)[4][4])
{
#pragma unroll
	for (int n2 = 0; n2 < 4; n2++) dft4(x[n2], x[4 + n2], x[8 + n2], x[12 + n2], A[n2][0], A[n2][1], A[n2][2], A[n2][3]);
	A[1][1] = cmul(A[1][1], make_float2(MB_C1, -MB_S1));   // W16^1
	A[1][2] = cmul(A[1][2], make_float2(MB_R2, -MB_R2));   // W16^2
	A[1][3] = cmul(A[1][3], make_float2(MB_S1, -MB_C1));   // W16^3
	A[2][1] = cmul(A[2][1], make_float2(MB_R2, -MB_R2));   // W16^2
	A[2][2] = mul_mi(A[2][2]);                             // W16^4
	A[2][3] = cmul(A[2][3], make_float2(-MB_R2, -MB_R2));  // W16^6
	A[3][1] = cmul(A[3][1], make_float2(MB_S1, -MB_C1));   // W16^3
	A[3][2] = cmul(A[3][2], make_float2(-MB_R2, -MB_R2));  // W16^6
	A[3][3] = cmul(A[3][3], make_float2(-MB_C1, MB_S1));   // W16^9
}

// full 16-point forward DFT, natural order out: X[k1 + 4 k2]
__device__ __forceinline__ void fft16(const float2 (&x)[16], float2 (&X)[16])
{
	float2 A[4][4];
	fft16_front(x, A);
#pragma unroll
	for (int k1 = 0; k1 < 4; k1++) dft4(A[0][k1], A[1][k1], A[2][k1], A[3][k1], X[k1], X[k1 + 4], X[k1 + 8], X[k1 + 12]);
}

// 16-point forward DFT pruned to outputs 0, 1, 14, 15 (the only ones that reach the 50 active carriers)
__device__ __forceinline__ void fft16_pruned(const float2 (&x)[16], float2 &X0, float2 &X1, float2 &X14, float2 &X15)
{
	float2 A[4][4];
	fft16_front(x, A);
	X0 = cadd(cadd(A[0][0], A[2][0]), cadd(A[1][0], A[3][0]));                   // k1=0, k2=0
	X1 = cadd(cadd(A[0][1], A[2][1]), cadd(A[1][1], A[3][1]));                   // k1=1, k2=0
	X14 = cadd(csub(A[0][2], A[2][2]), mul_pi(csub(A[1][2], A[3][2])));          // k1=2, k2=3
	X15 = cadd(csub(A[0][3], A[2][3]), mul_pi(csub(A[1][3], A[3][3])));          // k1=3, k2=3
}

// Warp-level sum of three values; lane 0 of every warp parks its partial sums in s_part[warp*3 .. +2].  The CTA-wide
// totals are formed by every thread after the next barrier the algorithm needs anyway (no reduction-only barriers).
__device__ __forceinline__ void warp_partials3(float a, float b, float c, float *s_part)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		a += __shfl_xor_sync(0xffffffffu, a, o);
		b += __shfl_xor_sync(0xffffffffu, b, o);
		c += __shfl_xor_sync(0xffffffffu, c, o);
	}
	if ((threadIdx.x & 31) == 0) {
		const int w = threadIdx.x >> 5;
		s_part[w * 3 + 0] = a;
		s_part[w * 3 + 1] = b;
		s_part[w * 3 + 2] = c;
	}
}

// Max-log LLRs of one equalised symbol (psk.cc:278-326): for bit k (mask 1<<k) (min_{bit=1} D - min_{bit=0} D) / variance,
// emitted MSB first.  M is a compile-time constant so every (j >> k) & 1 test folds away.
template <int M, int BPS>
__device__ __forceinline__ void demap_scatter(const float2 z, const float inv_var, const float2 *s_cons, const int base,
					      const uint16_t *__restrict__ g_dst, const uint16_t *__restrict__ g_dst2, float *s_L)
{
	float d0[BPS], d1[BPS];
#pragma unroll
	for (int k = 0; k < BPS; k++) d0[k] = d1[k] = 3.0e38f;
#pragma unroll
	for (int j = 0; j < M; j++) {
		const float2 cj = s_cons[j];
		const float dx = z.x - cj.x, dy = z.y - cj.y;
		const float D = dx * dx + dy * dy;
#pragma unroll
		for (int k = 0; k < BPS; k++) {
			if ((j >> k) & 1)
				d1[k] = fminf(d1[k], D);
			else
				d0[k] = fminf(d0[k], D);
		}
	}
#pragma unroll
	for (int k = 0; k < BPS; k++) {
		const float llr = inv_var * (d1[k] - d0[k]);
		const int i = base + (BPS - 1 - k);
		s_L[g_dst[i]] = llr;
		const unsigned d2 = g_dst2[i];
		if (d2 != MB_NO_DST) s_L[d2] = llr;
	}
}

constexpr int kSmemHeadFloats = 2 * 32 + 2 * 32;  // constellation, two sets of per-warp partial sums
constexpr int kZfStride = 27;                     // compact pilot row: 4 zeros | <=17 pilots | zeros
constexpr int kMaxThreads = 192;

template <bool kDebug>
__global__ void __launch_bounds__(kMaxThreads, 5) mb_demod_kernel(const MbDemodArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const MbMode &m = a.mode;
	const int T = blockDim.x, tid = threadIdx.x, nwarps = T >> 5;
	const int S = m.Nsymb, cells = S * MB_NC;
	float2 *s_cons = reinterpret_cast<float2 *>(smem_raw);
	float *s_part1 = reinterpret_cast<float *>(s_cons + 32);  // AGC partials
	float *s_part2 = s_part1 + 32;                             // pilot-statistics partials
	float2 *s_Y = reinterpret_cast<float2 *>(s_part2 + 32);
	float2 *s_buf = s_Y + cells;               // FFT transpose scratch, then reused:
	float2 *s_zf = s_buf;                      //   [S][27] compact zero-padded pilot rows of Y/p, later the channel at pilots
	const int zf_elems = (S * kZfStride + 1) & ~1;  // keep everything behind it 16-byte aligned (float4 reads of s_L)
	float2 *s_T = s_buf + zf_elems;            //   [S][50] clipped 21-column window sums of each row
	float *s_L = reinterpret_cast<float *>(s_T + cells);  // [1600] LLRs, decoder order

	const size_t frame = blockIdx.x;
	const float2 *__restrict__ xf = a.x + frame * (size_t)S * MB_NOFDM;
	const float *__restrict__ g_pinv = reinterpret_cast<const float *>(a.blob + m.off_pinv);
	const float *__restrict__ g_pval = reinterpret_cast<const float *>(a.blob + m.off_pval);
	const float *__restrict__ g_invn = reinterpret_cast<const float *>(a.blob + m.off_invn);
	const uint32_t *__restrict__ g_pilot_info = reinterpret_cast<const uint32_t *>(a.blob + m.off_pilot_info);
	const uint32_t *__restrict__ g_sym_info = reinterpret_cast<const uint32_t *>(a.blob + m.off_sym_info);
	const uint16_t *__restrict__ g_dst = reinterpret_cast<const uint16_t *>(a.blob + m.off_llr_dst);
	const uint16_t *__restrict__ g_dst2 = reinterpret_cast<const uint16_t *>(a.blob + m.off_llr_dst2);

	if (tid < m.M) s_cons[tid] = reinterpret_cast<const float2 *>(a.blob + m.off_const)[tid];  // published by the barrier after the FFT

	// ---------------- FFT-256 per symbol (a2, a3) ------------------------------------------------------------
	{
		const float2 *__restrict__ g_tw = reinterpret_cast<const float2 *>(a.blob + a.off_twiddle);  // 2 KB, L1 resident
		const int grp = tid >> 4, t = tid & 15, spr = T >> 4;
		float2 *buf = s_buf + grp * (16 * 17);
		float2 tw[16];
#pragma unroll
		for (int k1 = 1; k1 < 16; k1++) tw[k1] = __ldg(g_tw + k1 * 16 + t);  // W256^(t k1)/256, kept across rounds
		tw[0] = make_float2(1.0f / 256.0f, 0.f);
		for (int s0 = 0; s0 < S; s0 += spr) {
			const int s = s0 + grp;
			const bool active = s < S;
			if (active) {
				const float2 *__restrict__ xs = xf + (size_t)s * MB_NOFDM + MB_NGI + t;
				float2 v[16], A[16];
#pragma unroll
				for (int n1 = 0; n1 < 16; n1++) v[n1] = __ldcs(xs + 16 * n1);  // x[16 n1 + t], GI skipped
				fft16(v, A);
				buf[t * 17] = cscale(A[0], 1.0f / 256.0f);
#pragma unroll
				for (int k1 = 1; k1 < 16; k1++) buf[t * 17 + k1] = cmul(A[k1], tw[k1]);
			}
			__syncwarp();
			if (active) {
				float2 w[16], X0, X1, X14, X15;
#pragma unroll
				for (int n2 = 0; n2 < 16; n2++) w[n2] = buf[n2 * 17 + t];
				fft16_pruned(w, X0, X1, X14, X15);  // bins t, 16+t, 224+t, 240+t
				float2 *row = s_Y + s * MB_NC;
				// zero_depadder (ofdm.cc:401-411): bins 231..255 -> carriers 0..24, bins 1..25 -> carriers 25..49
				if (t >= 1) row[24 + t] = X0;
				if (t <= 9) row[40 + t] = X1;
				if (t >= 7) row[t - 7] = X14;
				row[9 + t] = X15;
			}
			__syncwarp();
		}
	}
	__syncthreads();

	// ---------------- AGC sum (a4) + zero-forced pilots into compact, zero-padded rows ---------------------------
	// Row s holds its pilots (columns s%3 + 3j) at [4 + j]; everything else in the 27-wide row is zero, so that every
	// clipped 21-column window is exactly 7 consecutive entries (21 consecutive integers hold 7 of each residue mod 3).
	for (int i = tid; i < S * kZfStride; i += T) {
		const int s = i / kZfStride, j = i - s * kZfStride;
		if (j < 4 || j > 20 || (j == 20 && s % 3 == 2)) s_zf[i] = make_float2(0.f, 0.f);  // rows with s%3==2 hold 16 pilots only
	}
	{
		float acc = 0.f;
		for (int p = tid; p < m.nPilots; p += T) {
			const unsigned info = g_pilot_info[p];
			const int cell = info & 0xFFF, s = (info >> 12) & 0x3F, j = info >> 18;
			const float2 y = s_Y[cell];
			acc += sqrtf(y.x * y.x + y.y * y.y);
			const float w = g_pinv[cell];
			s_zf[s * kZfStride + 4 + j] = make_float2(y.x * w, y.y * w);  // ZF estimate Y/p (AGC gain applied later: all linear)
		}
		warp_partials3(acc, 0.f, 0.f, s_part1);
	}
	__syncthreads();

	// ---------------- LS estimate (a5), row pass: clipped 21-column window sums over the pilot lattice -----------------
	if (m.estimator == 1 && tid < 3 * MB_NC) {
		const int r = tid / MB_NC, c = tid - r * MB_NC;  // this thread: column c of the rows with s%3 == r
		const int lo = (c + 4 - r) / 3;                  // compact index of the first pilot column >= c-10 in such a row
		for (int k = r; k < S; k += 3) {
			const float2 *row = s_zf + k * kZfStride + lo;
			float sx = 0.f, sy = 0.f;
#pragma unroll
			for (int j = 0; j < 7; j++) {
				sx += row[j].x;
				sy += row[j].y;
			}
			s_T[k * MB_NC + c] = make_float2(sx, sy);
		}
	}
	float g;
	{
		float acc = 0.f;
		for (int w = 0; w < nwarps; w++) acc += s_part1[w * 3];
		g = m.boost / (acc / (float)m.nPilots);  // automatic_gain_control, ofdm.cc:1467-1498
	}
	__syncthreads();

	// ---------------- channel at pilots (column pass), pilot-domain statistics (a5/a6, a8-a10) ------------------------
	{
		float accH = 0.f, accV = 0.f, accVn = 0.f;
		for (int p = tid; p < m.nPilots; p += T) {
			const unsigned info = g_pilot_info[p];
			const int cell = info & 0xFFF, s = (info >> 12) & 0x3F, j = info >> 18;
			const float2 yg = cscale(s_Y[cell], g);
			float2 h;
			if (m.estimator == 1) {
				const int c = cell - s * MB_NC;
				const int k0 = max(0, s - MB_LS_HALF), k1 = min(S - 1, s + MB_LS_HALF);
				float sx0 = 0.f, sy0 = 0.f, sx1 = 0.f, sy1 = 0.f;
				int k = k0;
				for (; k + 1 <= k1; k += 2) {
					const float2 ta = s_T[k * MB_NC + c], tb = s_T[(k + 1) * MB_NC + c];
					sx0 += ta.x, sy0 += ta.y, sx1 += tb.x, sy1 += tb.y;
				}
				if (k <= k1) {
					const float2 ta = s_T[k * MB_NC + c];
					sx0 += ta.x, sy0 += ta.y;
				}
				const float w = g_invn[cell] * g;
				h = make_float2((sx0 + sx1) * w, (sy0 + sy1) * w);
			} else {
				h = cscale(s_zf[s * kZfStride + 4 + j], g);  // ZF: H = Y / p
			}
			// The channel at pilots goes back into the compact rows at the pilot's own slot: in LS mode the rows are dead
			// (the row pass consumed them before the barrier), in ZF mode this thread is the slot's only user.
			s_zf[s * kZfStride + 4 + j] = h;
			const float h2 = h.x * h.x + h.y * h.y;
			accH += sqrtf(h2);
			const float pv = g_pval[cell];
			const float2 yc = make_float2(yg.x * h.x + yg.y * h.y, yg.y * h.x - yg.x * h.y);  // yg * conj(h)
			float2 z, heq = h;
			if (m.phase_only) {
				// restore_channel_amplitude (ofdm.cc:1453-1466): H <- exp(j arg H); dividing by a unit-modulus number is
				// multiplying by its conjugate.  get_angle() returns pi/2 whenever Re H == 0 (misc.cc:38-41).
				const float inv = rsqrtf(h2), inv2 = 1.0f / h2;
				if (h.x == 0.f) {
					heq = make_float2(0.f, 1.f);
					z = make_float2(yg.y, -yg.x);
				} else {
					heq = make_float2(h.x * inv, h.y * inv);
					z = make_float2(yc.x * inv, yc.y * inv);
				}
				const float2 zn = make_float2(yc.x * inv2, yc.y * inv2);  // without amplitude restoration: SNR report only
				accVn += (zn.x - pv) * (zn.x - pv) + zn.y * zn.y;
			} else {
				const float inv2 = 1.0f / h2;
				z = make_float2(yc.x * inv2, yc.y * inv2);
			}
			accV += (z.x - pv) * (z.x - pv) + z.y * z.y;
			if (kDebug) {
				const size_t o = frame * (size_t)cells + cell;
				if (a.dbg_Y) a.dbg_Y[o] = yg;
				if (a.dbg_H) a.dbg_H[o] = heq;
				if (a.dbg_Z) a.dbg_Z[o] = z;
			}
		}
		warp_partials3(accH, accV, accVn, s_part2);
	}
	__syncthreads();
	float accH = 0.f, accV = 0.f, accVn = 0.f;
	for (int w = 0; w < nwarps; w++) {
		accH += s_part2[w * 3 + 0];
		accV += s_part2[w * 3 + 1];
		accVn += s_part2[w * 3 + 2];
	}
	const float inv_np = 1.0f / (float)m.nPilots;
	// The reference has no floor here; 1e-30 only matters where it would produce inf/NaN LLRs (ZF modes)
	const float variance = fmaxf(accV * inv_np, 1e-30f);
	const float inv_var = 1.0f / variance;

	// ---------------- data cells: interpolate, equalise, de-map, scatter (a7-a9, a11-a14) --------------------
	for (int q = tid; q < m.nData; q += T) {
		const unsigned info = g_sym_info[q];
		const int cell = info & 0xFFF, r0 = (info >> 12) & 0x3F, j = info >> 21;
		const float t3 = (float)((int)((info >> 18) & 7) - 2) * (1.0f / 3.0f);
		// interpolate_linear_col (interpolator.cc:163-254): a + (b - a) * (x - xa) / (xb - xa), pilot rows 3 apart
		const float2 ha = s_zf[r0 * kZfStride + 4 + j], hb = s_zf[(r0 + 3) * kZfStride + 4 + j];
		const float2 h = make_float2(fmaf(hb.x - ha.x, t3, ha.x), fmaf(hb.y - ha.y, t3, ha.y));
		const float2 yg = cscale(s_Y[cell], g);
		const float h2 = h.x * h.x + h.y * h.y;
		const float2 yc = make_float2(yg.x * h.x + yg.y * h.y, yg.y * h.x - yg.x * h.y);
		float2 z, heq = h;
		if (m.phase_only) {
			const float inv = rsqrtf(h2);
			if (h.x == 0.f) {
				heq = make_float2(0.f, 1.f);
				z = make_float2(yg.y, -yg.x);
			} else {
				heq = make_float2(h.x * inv, h.y * inv);
				z = make_float2(yc.x * inv, yc.y * inv);
			}
		} else {
			const float inv2 = 1.0f / h2;
			z = make_float2(yc.x * inv2, yc.y * inv2);
		}
		if (kDebug) {
			const size_t o = frame * (size_t)cells + cell;
			if (a.dbg_Y) a.dbg_Y[o] = yg;
			if (a.dbg_H) a.dbg_H[o] = heq;
			if (a.dbg_Z) a.dbg_Z[o] = z;
		}
		switch (m.M) {
		case 2: demap_scatter<2, 1>(z, inv_var, s_cons, q, g_dst, g_dst2, s_L); break;
		case 4: demap_scatter<4, 2>(z, inv_var, s_cons, q * 2, g_dst, g_dst2, s_L); break;
		case 8: demap_scatter<8, 3>(z, inv_var, s_cons, q * 3, g_dst, g_dst2, s_L); break;
		case 16: demap_scatter<16, 4>(z, inv_var, s_cons, q * 4, g_dst, g_dst2, s_L); break;
		default: demap_scatter<32, 5>(z, inv_var, s_cons, q * 5, g_dst, g_dst2, s_L); break;
		}
	}
	__syncthreads();

	// ---------------- write LLRs (decoder order, coalesced) and the demod half of the stats record -----------
	{
		float4 *__restrict__ out = reinterpret_cast<float4 *>(a.llr + frame * (size_t)MB_N);
		const float4 *src = reinterpret_cast<const float4 *>(s_L);
		for (int i = tid; i < MB_N / 4; i += T) out[i] = src[i];
		if (a.llr_cw) {
			const uint16_t *__restrict__ g_voc = reinterpret_cast<const uint16_t *>(a.blob + a.off_var_of_cw);
			float *__restrict__ o2 = a.llr_cw + frame * (size_t)MB_N;
			for (int i = tid; i < MB_N; i += T) o2[i] = s_L[g_voc[i]];
		}
		if (tid == 0) {
			MbRxStats st;
			st.iterations_done = -1;
			st.crc = 0;
			st.all_zeros = 0;
			st.message_decoded = 0;
			const float v_rep = m.phase_only ? accVn * inv_np : variance;
			st.SNR = m.estimator == 1 ? 10.0f * log10f(1.0f / v_rep) : 0.0f;  // candidate; finalised by the decoder
			st.variance = variance;
			st.mean_H = accH * inv_np;
			st.reserved = 0;
			a.stats[frame] = st;
		}
	}
}

}  // namespace

size_t mb_demod_smem_bytes(int Nsymb)
{
	const int T = mb_demod_threads(Nsymb), cells = Nsymb * MB_NC;
	size_t fftbuf = (size_t)(T / 16) * 16 * 17 * sizeof(float2);
	size_t reuse = ((size_t)((Nsymb * kZfStride + 1) & ~1) + (size_t)cells) * sizeof(float2) + MB_N * sizeof(float);
	return kSmemHeadFloats * sizeof(float) + (size_t)cells * sizeof(float2) + (fftbuf > reuse ? fftbuf : reuse);
}

cudaError_t mb_demod_init()
{
	cudaError_t e = cudaFuncSetAttribute(mb_demod_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute(mb_demod_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

cudaError_t mb_launch_demod(const MbDemodArgs &a, size_t n_frames, cudaStream_t stream)
{
	if (n_frames == 0) return cudaSuccess;
	const int T = mb_demod_threads(a.mode.Nsymb);
	const size_t smem = mb_demod_smem_bytes(a.mode.Nsymb);
	const bool dbg = a.dbg_Y || a.dbg_H || a.dbg_Z;
	if (dbg)
		mb_demod_kernel<true><<<(unsigned)n_frames, T, smem, stream>>>(a);
	else
		mb_demod_kernel<false><<<(unsigned)n_frames, T, smem, stream>>>(a);
	return cudaGetLastError();
}
