// mb_demod.cu -- K_demod: fused OFDM demodulator for sm_100a; one CTA per frame, streaming 8-byte loads in, one bulk (TMA engine) store out.
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
//   * HBM-bound stage: one CTA per frame, every sample is read exactly once with 8-byte streaming loads (16 independent
//     loads in flight per thread, guard interval never fetched); the only other HBM traffic is the 6.4 KB LLR vector,
//     which leaves through one bulk (TMA engine) store from shared memory.
//   * FFT-256 = 16 x 16: 16 threads per OFDM symbol, two radix-4x4 16-point DFTs in registers around a padded (stride 17,
//     conflict-free) shared-memory transpose that needs only __syncwarp (a symbol lives in half a warp).  All complex
//     arithmetic uses the sm_100 packed fp32 instructions (FADD2 / FMUL2 / FFMA2: one instruction per complex add, two per
//     complex multiply, +-i rotations folded into operand swizzles); they issue at half the scalar rate, so this halves
//     issue slots, not pipe time (profiles/r1e_f32x2_issue_rates.txt).  The second DFT is pruned to the 4 of 16 outputs
//     that land on the 50 active carriers.
//   * The LS estimator's O(pilots x window) double loop is restated as the clipped 21x21 box mean it is:
//     compact pilot rows -> 7-entry window sums (18 distinct windows per row) -> running sums over the rows of each
//     lattice residue -> every pilot reads 3 x (upper - lower) entries through host-resolved byte offsets.
//   * deframe, both de-interleavers and the LLR expand are composed on the host into per-cell records (mb_tables.cpp), so
//     no intermediate vector is ever materialised; the kernel is instantiated per (Nsymb, M, estimator, phase-only).
//   (A persistent variant fed by a cp.async.bulk ring was measured first: 74 KB of shared memory per CTA left 2 CTAs per
//    SM and 29 % of the HBM roofline, profiles/r1e_ncu_full_demod_tma_persistent.txt; occupancy wins on this kernel.)
#include "mb_fft.cuh"
#include <cuda_fp16.h>

#include "mb_kernels.cuh"

namespace {
using namespace mbfft;

// ------------------------------------------------------------------------------------------------------------------
// bulk (TMA engine) store of the LLR vector: shared -> global (SASS: UBLKCP)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
__device__ __forceinline__ void warp_partial1(float a, float *s_part)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
	if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = a;
}

// Max-log LLRs of one equalised symbol (psk.cc:278-326): for bit k (mask 1<<k) (min_{bit=1} D - min_{bit=0} D) / variance,
// emitted MSB first; dw[] holds the byte offsets of the emitted LLRs in the internal-order vector, two per word.
template <int M, int BPS, typename ConsT>
__device__ __forceinline__ void demap_scatter(const float2 z, const float inv_var, const ConsT &s_cons, const uint32_t (&dw)[3],
					      unsigned char *s_L)
{
	float d0[BPS], d1[BPS];
#pragma unroll
	for (int k = 0; k < BPS; k++) d0[k] = d1[k] = 3.0e38f;
#pragma unroll
	for (int j = 0; j < M; j++) {
		const float2 d = csub(z, s_cons[j]);
		const float D = cnorm2(d);
#pragma unroll
		for (int k = 0; k < BPS; k++) {
			if ((j >> k) & 1)
				d1[k] = fminf(d1[k], D);
			else
				d0[k] = fminf(d0[k], D);
		}
	}
#pragma unroll
	for (int e = 0; e < BPS; e++) {
		const int k = BPS - 1 - e;
		const float llr = inv_var * (d1[k] - d0[k]);
		const uint32_t off = (e & 1) ? (dw[e >> 1] >> 16) : (dw[e >> 1] & 0xFFFFu);
		*reinterpret_cast<float *>(s_L + off) = llr;
	}
}

// 16QAM (psk.cc:117-140): the constellation is a product set -- index bits 3,2 pick the in-phase level, bits 1,0 the quadrature level --
// so the minimum of |z - c|^2 over the points with bit k = b splits into a minimum over the 2 levels of that bit's axis plus the minimum
// over the other axis, and the latter is common to both hypotheses: it cancels in Dmin1 - Dmin0.  8 squared differences and 8 minima
// per symbol instead of 16 distances and 64 minima; equal to the reference's float distances up to their own rounding (~1e-7 D / sigma^2,
// three orders inside the LLR tolerance).  lv[0..3] = in-phase level of index bits 3,2 = a, lv[4..7] = quadrature level of bits 1,0 = b.
__device__ __forceinline__ void demap_scatter_qam16(const float2 z, const float inv_var, const float (&lv)[8], const uint32_t (&dw)[3], unsigned char *s_L)
{
	float dI[4], dQ[4];
#pragma unroll
	for (int a = 0; a < 4; a++) {
		const float x = z.x - lv[a], y = z.y - lv[4 + a];
		dI[a] = x * x, dQ[a] = y * y;
	}
	const float llr[4] = {inv_var * (fminf(dI[2], dI[3]) - fminf(dI[0], dI[1])),   // index bit 3, emitted first (MSB first)
			      inv_var * (fminf(dI[1], dI[3]) - fminf(dI[0], dI[2])),   // bit 2
			      inv_var * (fminf(dQ[2], dQ[3]) - fminf(dQ[0], dQ[1])),   // bit 1
			      inv_var * (fminf(dQ[1], dQ[3]) - fminf(dQ[0], dQ[2]))};  // bit 0
#pragma unroll
	for (int e = 0; e < 4; e++) {
		const uint32_t off = (e & 1) ? (dw[e >> 1] >> 16) : (dw[e >> 1] & 0xFFFFu);
		*reinterpret_cast<float *>(s_L + off) = llr[e];
	}
}

// 32QAM (psk.cc:142-190): the cross constellation -- levels {+-1, +-3, +-5} u on both axes without the four corners -- is a union of
// rectangles: per half plane (index bit 4 = sign of the in-phase level), bits 3,2 = 00 is the cap (in-phase 3 u or 1 u by bit 0, quadrature
// +-5 u by bit 1), 01 / 10 / 11 the columns 5 u / 1 u / 3 u with quadrature 3, 1, -3, -1 u by bits 1,0.  The minimum of |z - c|^2 over a
// rectangle is the sum of the two per-axis minima, so every Dmin is the smaller of at most two such sums: 12 squared differences and ~45
// minima / sums per symbol instead of 32 distances and 160 minima (checked against the brute force over the table in tests/test_tables.py).
__device__ __forceinline__ void demap_scatter_qam32(const float2 z, const float inv_var, const float u, const uint32_t (&dw)[3], unsigned char *s_L)
{
	auto sq = [](float a) { return a * a; };
	const float X5m = sq(z.x + 5.f * u), X3m = sq(z.x + 3.f * u), X1m = sq(z.x + u), X1p = sq(z.x - u), X3p = sq(z.x - 3.f * u), X5p = sq(z.x - 5.f * u);
	const float Y5m = sq(z.y + 5.f * u), Y3m = sq(z.y + 3.f * u), Y1m = sq(z.y + u), Y1p = sq(z.y - u), Y3p = sq(z.y - 3.f * u), Y5p = sq(z.y - 5.f * u);
	const float Am = fminf(X3m, X1m), Ap = fminf(X3p, X1p), Bm = fminf(Am, X5m), Bp = fminf(Ap, X5p);
	const float A = fminf(Am, Ap), B = fminf(Bm, Bp), C = fminf(fminf(X5m, X3m), fminf(X5p, X3p));
	const float X3 = fminf(X3m, X3p), X1 = fminf(X1m, X1p), X5 = fminf(X5m, X5p);
	const float Yb1_0 = fminf(Y3p, Y1p), Yb1_1 = fminf(Y3m, Y1m), Yb0_0 = fminf(Y3p, Y3m), Yb0_1 = fminf(Y1p, Y1m);
	const float Yc = fminf(Yb1_0, Yb1_1), Ycap = fminf(Y5p, Y5m), AYcap = A + Ycap;
	const float llr[5] = {inv_var * (fminf(Ap + Ycap, Bp + Yc) - fminf(Am + Ycap, Bm + Yc)),        // index bit 4, emitted first (MSB first)
			      inv_var * ((A + Yc) - fminf(AYcap, X5 + Yc)),                                 // bit 3
			      inv_var * ((C + Yc) - fminf(AYcap, X1 + Yc)),                                 // bit 2
			      inv_var * (fminf(A + Y5m, B + Yb1_1) - fminf(A + Y5p, B + Yb1_0)),             // bit 1
			      inv_var * (fminf(X1 + Ycap, B + Yb0_1) - fminf(X3 + Ycap, B + Yb0_0))};        // bit 0
#pragma unroll
	for (int e = 0; e < 5; e++) {
		const uint32_t off = (e & 1) ? (dw[e >> 1] >> 16) : (dw[e >> 1] & 0xFFFFu);
		*reinterpret_cast<float *>(s_L + off) = llr[e];
	}
}

// ------------------------------------------------------------------------------------------------------------------
// compile-time geometry of one instantiation
// ------------------------------------------------------------------------------------------------------------------
template <int S>
struct Geo {
	// Symbols transformed per round = threads per CTA / 16.  Measured on B200 (tools/r2_tune10.sh, ms per 65,536 frames): a 24-symbol frame
	// in 4 rounds of 6 (96 threads, 6 CTAs per SM at 94 registers) 0.864 against 0.927 in 2 rounds of 12 and 0.886 / 0.967 / 1.10 in rounds
	// of 8 / 4 / 3; a 12-symbol frame in 2 rounds of 6 at 80 registers 0.461 against 0.527 in one round (0.47 / 0.51 in rounds of 4 / 3);
	// 48-symbol frames stay at 12 (6: 1.84, 8: 1.79 against 1.72), 16-symbol frames at 8 (4: 0.65, 16: 0.73 against 0.63), 9-symbol
	// frames at 9 (3: 0.42 against 0.43, 1: 0.61).  Small CTAs overlap one frame's loads with another's arithmetic at a finer grain.
	static constexpr int SC = S == 48 ? 12 : (S == 24 ? 6 : (S == 16 ? 8 : (S == 12 ? 6 : S)));
	static constexpr int NCH = S / SC;                                                // rounds per frame
	static constexpr int T = (SC * 16 + 31) / 32 * 32;                                // threads per CTA
	static constexpr int NW = T / 32;
	// resident CTAs per SM the register file is budgeted for (96 threads: 94 registers for the 24-symbol frames, 80 for the 12-symbol ones)
	static constexpr int MINB = T <= 96 ? (S == 12 ? 7 : 6) : (T <= 160 ? 6 : (T <= 192 ? 5 : 3));
	static constexpr int CELLS = S * MB_NC;
	static constexpr int NPIL = (S * MB_NC + 2) / 3;   // pilots: cells with s%3 == c%3
	static constexpr int NDATA = CELLS - NPIL;
	static constexpr int ZF = S * MB_ZF_STRIDE;
	static constexpr int PM = (S + 1) * MB_LS_COLS;
	// shared memory layout (bytes); the FFT transpose scratch is dead before the estimator's arrays are first written
	static constexpr int OFF_PART = 0;                       // 64 floats of per-warp partial sums
	static constexpr int OFF_CONS = OFF_PART + 256;          // 32 float2 constellation
	static constexpr int OFF_Y = OFF_CONS + 256;             // [S][50] float2 carriers
	static constexpr int OFF_ZF = OFF_Y + CELLS * 8;         // [S][27] float2 compact pilot rows (filled by the FFT epilogue), later the channel at pilots
	static constexpr int OFF_SCR = OFF_ZF + ((ZF * 8 + 15) & ~15);  // [SC][16][17] float2 transpose scratch, aliased with:
	static constexpr int OFF_PM = OFF_SCR;                   //   [S+1][18] float2 window sums / running sums
	static constexpr int OFF_L = OFF_PM + PM * 8;            //   [1600] float LLRs, hand-off order
	static constexpr int SCR_BYTES = SC * 16 * 17 * 8;
	static constexpr int EST_BYTES = OFF_L + MB_N * 4 - OFF_SCR;
	static constexpr int SMEM = OFF_SCR + (SCR_BYTES > EST_BYTES ? SCR_BYTES : EST_BYTES);
	static_assert(S % SC == 0, "rounds must tile the frame");
	static_assert(OFF_L % 16 == 0 && OFF_Y % 16 == 0 && OFF_ZF % 16 == 0 && OFF_PM % 16 == 0, "alignment");
};

template <int S, int M, bool LS, bool PHASE, bool NARROW>
__global__ void __launch_bounds__(Geo<S>::T, Geo<S>::MINB) mb_demod_kernel(const MbDemodArgs a)
{
	using G = Geo<S>;
	constexpr int T = G::T, NW = G::NW, SC = G::SC, NCH = G::NCH;
	constexpr int BPS = M == 2 ? 1 : (M == 4 ? 2 : (M == 8 ? 3 : (M == 16 ? 4 : 5)));
	constexpr int RECW = BPS <= 2 ? 2 : 4;
	extern __shared__ __align__(128) unsigned char smem[];
	const MbMode &m = a.mode;
	const int tid = threadIdx.x;
	float *s_part = reinterpret_cast<float *>(smem + G::OFF_PART);
	float2 *s_cons = reinterpret_cast<float2 *>(smem + G::OFF_CONS);
	float2 *s_Y = reinterpret_cast<float2 *>(smem + G::OFF_Y);
	unsigned char *s_Yb = smem + G::OFF_Y;
	float2 *s_zf = reinterpret_cast<float2 *>(smem + G::OFF_ZF);
	unsigned char *s_zfb = smem + G::OFF_ZF;
	float2 *s_pm = reinterpret_cast<float2 *>(smem + G::OFF_PM);
	unsigned char *s_pmb = smem + G::OFF_PM;
	unsigned char *s_Lb = smem + G::OFF_L;
	float *s_L = reinterpret_cast<float *>(s_Lb);

	const size_t frame = blockIdx.x;
	const uint4 *__restrict__ g_prec = reinterpret_cast<const uint4 *>(a.blob + m.off_pilot_rec);
	const float2 *__restrict__ g_pf = reinterpret_cast<const float2 *>(a.blob + m.off_pilot_f);
	const uint32_t *__restrict__ g_drec = reinterpret_cast<const uint32_t *>(a.blob + m.off_data_rec);
	const bool dbg = a.dbg_Y || a.dbg_H || a.dbg_Z;

	if (tid < M) s_cons[tid] = reinterpret_cast<const float2 *>(a.blob + m.off_const)[tid];  // published by the barrier after the FFT

	// zero padding of the compact pilot rows: row s holds its pilots (columns s%3 + 3j) at [4 + j], everything else in the
	// 27-wide row is zero, so that every clipped 21-column window is exactly 7 consecutive entries
	for (int i = tid; i < S * 11; i += T) {
		const int sr = i / 11, q = i - sr * 11;
		const int j = q < 4 ? q : (q < 10 ? 17 + q : (sr % 3 == 2 ? 20 : 26));  // rows with s%3 == 2 hold 16 pilots only
		s_zf[sr * MB_ZF_STRIDE + j] = make_float2(0.f, 0.f);
	}

	// ---------------- FFT-256 per symbol (a2, a3), SC symbols per round ---------------------------------------------------
	const uint32_t pinv_bits = __float_as_uint(m.pinv_mag);
	float agc = 0.f;
	{
		const unsigned long long *__restrict__ g_neg = reinterpret_cast<const unsigned long long *>(a.blob + m.off_pilot_neg);
		const int grp = tid >> 4, t = tid & 15;
		const bool active = grp < SC;  // Nsymb = 9: the last half warp has no symbol, but still takes part in the warp syncs
		// Transpose scratch of one symbol: 16 rows of 17 float2 (the pad makes both the row writes and the column reads
		// conflict free, and every access is base + immediate).
		float2 *buf = reinterpret_cast<float2 *>(smem + G::OFF_SCR) + (active ? grp : 0) * (16 * 17);
		// Twiddles W256^(t k1), k1 = 1..15, are powers of w1 = W256^t: generated by a running product (two packed instructions
		// each, 15 roundings at most: ~1e-6 relative) instead of 15 table reads per round -- the shared-memory / L1 data pipe
		// is this kernel's busiest unit, the FMA pipe is not.  The table's 1/256 (ofdm.cc:439-442) moves to the 4 outputs.
		const float2 w1 = cscale(__ldg(reinterpret_cast<const float2 *>(a.blob + a.off_twiddle) + 16 + t), 256.0f);
		const size_t stride = (size_t)a.sym_stride;
		// sample formats of the input (MbDemodArgs::x_format): complex64, or (NARROW instantiations) 4-byte samples -- complex int16
		// (x a.x_scale) / complex fp16 -- that halve the bytes a host batch moves over PCIe; converted here, in the load, to the float2
		// the rest of the kernel works on
		const int fmt = NARROW ? a.x_format : 0;
		constexpr unsigned esz = NARROW ? 4u : 8u;
		const char *__restrict__ xs = reinterpret_cast<const char *>(a.x) + ((frame * S + (size_t)(active ? grp : 0)) * stride + a.sym_skip + t) * esz;
		if (NCH > 1 && active) {  // later rounds: one 128-byte line per thread into L2 while round 0 is in flight
			const char *line = reinterpret_cast<const char *>(a.x) + ((frame * S + (size_t)grp) * stride + a.sym_skip) * esz + t * 128;
#pragma unroll
			for (int ch = 1; ch < NCH; ch++)
				if (!NARROW || t < 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(line + (size_t)ch * SC * stride * esz));
		}
#pragma unroll 1
		for (int ch = 0; ch < NCH; ch++, xs += (size_t)SC * stride * esz) {
			float2 v[16];
			if (active) {
				if (!NARROW) {
#pragma unroll
					for (int n1 = 0; n1 < 16; n1++) v[n1] = ld_stream(reinterpret_cast<const float2 *>(xs) + 16 * n1);  // x[16 n1 + t], GI skipped
				} else {
					uint32_t w[16];
#pragma unroll
					for (int n1 = 0; n1 < 16; n1++) w[n1] = ld_stream_b32(reinterpret_cast<const uint32_t *>(xs) + 16 * n1);
					if (fmt == 1) {
						const float sc = a.x_scale;
#pragma unroll
						for (int n1 = 0; n1 < 16; n1++) v[n1] = make_float2((float)(short)(w[n1] & 0xFFFFu) * sc, (float)(short)(w[n1] >> 16) * sc);
					} else {
#pragma unroll
						for (int n1 = 0; n1 < 16; n1++) v[n1] = __half22float2(*reinterpret_cast<const __half2 *>(&w[n1]));
					}
				}
				float2 A[16];
				fft16(v, A);
				buf[t * 17] = A[0];
				float2 wk = w1;
#pragma unroll
				for (int k1 = 1; k1 < 16; k1++) {
					buf[t * 17 + k1] = cmul(A[k1], wk);
					if (k1 < 15) wk = cmul(wk, w1);
				}
			}
			__syncwarp();
			if (active) {
				float2 X0, X1, X14, X15;
#pragma unroll
				for (int n2 = 0; n2 < 16; n2++) v[n2] = buf[n2 * 17 + t];
				fft16_pruned(v, X0, X1, X14, X15);  // bins t, 16+t, 224+t, 240+t
				// zero_depadder (ofdm.cc:401-411): bins 231..255 -> carriers 0..24, bins 1..25 -> carriers 25..49.  Carriers that are
				// pilots (s%3 == c%3) also feed the AGC sum (a4) and, zero-forced (Y/p), the compact pilot rows of the estimator.
				const int sr = ch * SC + grp;
				float2 *row = s_Y + sr * MB_NC;
				float2 *zrow = s_zf + sr * MB_ZF_STRIDE + 4;
				const int r3 = sr % 3;
				const unsigned long long neg = __ldg(g_neg + sr);
				auto emit = [&](const float2 X, const int c) {
					const float2 y = cscale(X, 1.0f / 256.0f);
					row[c] = y;
					if (c % 3 == r3) {
						agc += fast_sqrt(cnorm2(y));
						zrow[c / 3] = cscale(y, __uint_as_float(pinv_bits ^ ((unsigned)(neg >> c) << 31)));
					}
				};
				if (t >= 1) emit(X0, 24 + t);
				if (t <= 9) emit(X1, 40 + t);
				if (t >= 7) emit(X14, t - 7);
				emit(X15, 9 + t);
			}
			__syncwarp();
		}
	}
	warp_partial1(agc, s_part);
	// descriptor records are L2 hits at best: fetch the first ones before the barriers they would otherwise wait behind
	uint4 rec = __ldg(g_prec + (tid < G::NPIL ? tid : 0));  // Nsymb = 9 has fewer pilots (150) than threads (160)
	float2 pf = __ldg(g_pf + (tid < G::NPIL ? tid : 0));
	static_assert(G::NDATA >= T, "first data descriptor of every thread exists");
	__syncthreads();
	if (LS && tid >= T - MB_LS_COLS) s_pm[S * MB_LS_COLS + (tid - (T - MB_LS_COLS))] = make_float2(0.f, 0.f);  // the "no lower bound" row

	// ---------------- AGC gain (a4) -------------------------------------------------------------------------------------
	const float inv_np = 1.0f / (float)G::NPIL;
	float g;
	{
		float acc = 0.f;
#pragma unroll
		for (int w = 0; w < NW; w++) acc += s_part[w];
		g = m.boost * fast_rcp(acc * inv_np);  // automatic_gain_control, ofdm.cc:1467-1498
	}
	// ---------------- LS estimate (a5): window sums + running sums over the rows of each lattice residue --------------
	if (LS) {
		if (tid < 3 * MB_LS_COLS) {
			const int r = tid / MB_LS_COLS, jj = tid - r * MB_LS_COLS;
			float2 run = make_float2(0.f, 0.f);
#pragma unroll
			for (int k3 = 0; k3 < (S + 2) / 3; k3++) {
				const int k = r + 3 * k3;
				if (k < S) {
					const float2 *row = s_zf + k * MB_ZF_STRIDE + jj;
					const float2 s01 = cadd(row[0], row[1]), s23 = cadd(row[2], row[3]), s45 = cadd(row[4], row[5]);
					run = cadd(run, cadd(cadd(s01, s23), cadd(s45, row[6])));
					s_pm[k * MB_LS_COLS + jj] = run;
				}
			}
		}
		__syncthreads();
	}

	// ---------------- channel at pilots, pilot-domain statistics (a5/a6, a8-a10) ------------------------------------
	{
		float accH = 0.f, accV = 0.f, accVn = 0.f;
#pragma unroll 1
		for (int p = tid; p < G::NPIL; p += T) {
			uint4 rec_n = rec;
			float2 pf_n = pf;
			if (p + T < G::NPIL) {
				rec_n = __ldg(g_prec + p + T);
				pf_n = __ldg(g_pf + p + T);
			}
			const uint32_t cellb = rec.w & 0xFFFFu, zslotb = rec.w >> 16;
			const float2 yg = cscale(*reinterpret_cast<const float2 *>(s_Yb + cellb), g);
			float2 h;
			if (LS) {
				const float2 u0 = *reinterpret_cast<const float2 *>(s_pmb + (rec.x & 0xFFFFu)), l0 = *reinterpret_cast<const float2 *>(s_pmb + (rec.x >> 16));
				const float2 u1 = *reinterpret_cast<const float2 *>(s_pmb + (rec.y & 0xFFFFu)), l1 = *reinterpret_cast<const float2 *>(s_pmb + (rec.y >> 16));
				const float2 u2 = *reinterpret_cast<const float2 *>(s_pmb + (rec.z & 0xFFFFu)), l2 = *reinterpret_cast<const float2 *>(s_pmb + (rec.z >> 16));
				const float2 sum = cadd(cadd(csub(u0, l0), csub(u1, l1)), csub(u2, l2));
				h = cscale(sum, pf.x * g);
			} else {
				h = cscale(*reinterpret_cast<const float2 *>(s_zfb + zslotb), g);  // ZF: H = Y / p
			}
			const float h2 = cnorm2(h);
			accH += fast_sqrt(h2);
			const float pv = pf.y;
			const float2 yc = cmul_conj(yg, h);
			float2 z, heq = h;
			if (PHASE) {
				// restore_channel_amplitude (ofdm.cc:1453-1466): H <- exp(j arg H); dividing by a unit-modulus number is
				// multiplying by its conjugate.  get_angle() returns pi/2 whenever Re H == 0 (misc.cc:38-41).
				const float inv = fast_rsqrt(h2), inv2 = fast_rcp(h2);
				if (h.x == 0.f) {
					heq = make_float2(0.f, 1.f);
					z = make_float2(yg.y, -yg.x);
				} else {
					heq = cscale(h, inv);
					z = cscale(yc, inv);
				}
				const float2 zn = cscale(yc, inv2);  // without amplitude restoration: SNR report only
				accVn += (zn.x - pv) * (zn.x - pv) + zn.y * zn.y;
			} else {
				z = cscale(yc, fast_rcp(h2));
			}
			accV += (z.x - pv) * (z.x - pv) + z.y * z.y;
			// The channel at pilots goes back into the compact rows at the pilot's own slot: in LS mode the rows are dead
			// (consumed by the window sums before the barrier), in ZF mode this thread is the slot's only user.
			*reinterpret_cast<float2 *>(s_zfb + zslotb) = h;
			if (dbg) {
				const size_t o = frame * (size_t)G::CELLS + (cellb >> 3);
				if (a.dbg_Y) a.dbg_Y[o] = yg;
				if (a.dbg_H) a.dbg_H[o] = heq;
				if (a.dbg_Z) a.dbg_Z[o] = z;
			}
			rec = rec_n;
			pf = pf_n;
		}
		warp_partials3(accH, accV, accVn, s_part + 8);
	}
	uint32_t dr[4] = {0u, 0u, 0u, 0u};
	if (RECW == 2) {
		const uint2 r = __ldg(reinterpret_cast<const uint2 *>(g_drec) + tid);
		dr[0] = r.x, dr[1] = r.y;
	} else {
		const uint4 r = __ldg(reinterpret_cast<const uint4 *>(g_drec) + tid);
		dr[0] = r.x, dr[1] = r.y, dr[2] = r.z, dr[3] = r.w;
	}
	__syncthreads();
	float accH = 0.f, accV = 0.f, accVn = 0.f;
#pragma unroll
	for (int w = 0; w < NW; w++) {
		accH += s_part[8 + w * 3 + 0];
		accV += s_part[8 + w * 3 + 1];
		accVn += s_part[8 + w * 3 + 2];
	}
	// The reference has no floor here; 1e-30 only matters where it would produce inf/NaN LLRs (ZF modes)
	const float variance = fmaxf(accV * inv_np, 1e-30f);
	const float inv_var = fast_rcp(variance);

	// ---------------- data cells in grid order: interpolate, equalise, de-map, scatter (a7-a9, a11-a14) --------------
	float2 *__restrict__ zf_out = reinterpret_cast<float2 *>(a.llr + frame * (size_t)MB_HANDOFF_STRIDE + MB_N);
	float2 c_reg[M <= 4 ? M : 1];  // BPSK / QPSK: the constellation lives in registers for the whole loop
#pragma unroll
	for (int j = 0; j < (M <= 4 ? M : 1); j++) c_reg[j] = s_cons[j];
	const float unit32 = M == 32 ? s_cons[9].y : 0.f;  // 32QAM: index 9 is (-1, +1) u
	float lv16[8];  // 16QAM: the 4 + 4 axis levels (index = a * 4 + b: in-phase from c[4 a], quadrature from c[b])
#pragma unroll
	for (int j = 0; j < 4; j++) lv16[j] = M == 16 ? s_cons[4 * j].x : 0.f, lv16[4 + j] = M == 16 ? s_cons[j].y : 0.f;
#pragma unroll 1
	for (int d = tid; d < G::NDATA; d += T) {
		const uint32_t w0 = dr[0], dw[3] = {dr[1], dr[2], dr[3]};
		if (d + T < G::NDATA) {  // next record in flight while this cell is equalised and de-mapped
			if (RECW == 2) {
				const uint2 r = __ldg(reinterpret_cast<const uint2 *>(g_drec) + d + T);
				dr[0] = r.x, dr[1] = r.y;
			} else {
				const uint4 r = __ldg(reinterpret_cast<const uint4 *>(g_drec) + d + T);
				dr[0] = r.x, dr[1] = r.y, dr[2] = r.z, dr[3] = r.w;
			}
		}
		const uint32_t cellb = w0 & 0x7FFFu, zs = (w0 >> 15) & 0x3FFFu;
		const float t3 = (float)((int)(w0 >> 29) - 2) * (1.0f / 3.0f);
		// interpolate_linear_col (interpolator.cc:163-254): a + (b - a) * (x - xa) / (xb - xa), pilot rows 3 apart
		const float2 ha = *reinterpret_cast<const float2 *>(s_zfb + zs), hb = *reinterpret_cast<const float2 *>(s_zfb + zs + 3 * MB_ZF_STRIDE * 8);
		const float2 h = __ffma2_rn(csub(hb, ha), make_float2(t3, t3), ha);
		const float2 yg = cscale(*reinterpret_cast<const float2 *>(s_Yb + cellb), g);
		const float h2 = cnorm2(h);
		const float2 yc = cmul_conj(yg, h);
		float2 z, heq = h;
		if (PHASE) {
			const float inv = fast_rsqrt(h2);
			if (h.x == 0.f) {
				heq = make_float2(0.f, 1.f);
				z = make_float2(yg.y, -yg.x);
			} else {
				heq = cscale(h, inv);
				z = cscale(yc, inv);
			}
		} else {
			z = cscale(yc, fast_rcp(h2));
		}
		if (dbg) {
			const size_t o = frame * (size_t)G::CELLS + (cellb >> 3);
			if (a.dbg_Y) a.dbg_Y[o] = yg;
			if (a.dbg_H) a.dbg_H[o] = heq;
			if (a.dbg_Z) a.dbg_Z[o] = z;
		}
		if (!LS) zf_out[d] = z;  // ZF modes: the decoder's SNR report re-encodes the frame and needs the equalised data symbols (:1376-1400)
		if (M <= 4)
			demap_scatter<M, BPS>(z, inv_var, c_reg, dw, s_Lb);
		else if (M == 16)
			demap_scatter_qam16(z, inv_var, lv16, dw, s_Lb);
		else if (M == 32)
			demap_scatter_qam32(z, inv_var, unit32, dw, s_Lb);
		else
			demap_scatter<M, BPS>(z, inv_var, s_cons, dw, s_Lb);
	}
	if (m.nVirtual > 0) {  // virtual bits are copies of the first LLRs (telecom_system.cc:1303-1306)
		__syncthreads();
		const uint32_t *__restrict__ g_virt = reinterpret_cast<const uint32_t *>(a.blob + m.off_virt);
		for (int i = tid; i < m.nVirtual; i += T) {
			const uint32_t w = __ldg(g_virt + i);
			*reinterpret_cast<float *>(s_Lb + (w >> 16)) = *reinterpret_cast<const float *>(s_Lb + (w & 0xFFFFu));
		}
	}
	fence_proxy_async();  // s_L was written through the generic proxy; the bulk store reads it through the async proxy
	__syncthreads();

	// ---------------- LLRs out (one bulk store, decoder order) and the demod half of the stats record ----------------
	if (tid == 0) {
		bulk_s2g(a.llr + frame * (size_t)MB_HANDOFF_STRIDE, smem_u32(s_L), MB_N * 4);
		MbRxStats st;
		st.iterations_done = -1;
		st.crc = 0;
		st.all_zeros = 0;
		st.message_decoded = 0;
		const float v_rep = PHASE ? accVn * inv_np : variance;
		st.SNR = LS ? 10.0f * log10f(1.0f / v_rep) : 0.0f;  // candidate; finalised by the decoder
		st.variance = variance;
		st.mean_H = accH * inv_np;
		st.reserved = 0;
		a.stats[frame] = st;
	}
	if (a.llr_cw) {
		const uint16_t *__restrict__ g_voc = reinterpret_cast<const uint16_t *>(a.blob + a.off_var_of_cw);
		float *__restrict__ o2 = a.llr_cw + frame * (size_t)MB_N;
		for (int i = tid; i < MB_N; i += T) o2[i] = s_L[MB_HANDOFF((uint32_t)g_voc[i])];
	}
	if (tid == 0) bulk_wait_read();  // shared memory must outlive the bulk store's read
}

// ------------------------------------------------------------------------------------------------------------------
// launch plumbing: one instantiation per (Nsymb, M, estimator, phase-only) combination of the 17 modes
// ------------------------------------------------------------------------------------------------------------------
struct Variant {
	int S, M, ls, phase, narrow;
	const void *fn;
	int threads, smem, ctas_per_sm;
};

#define MB_VARIANT(S, M, LS, PH)                                                                                   \
	{S, M, LS, PH, 0, (const void *)mb_demod_kernel<S, M, LS, PH, false>, Geo<S>::T, Geo<S>::SMEM, 0},         \
	{S, M, LS, PH, 1, (const void *)mb_demod_kernel<S, M, LS, PH, true>, Geo<S>::T, Geo<S>::SMEM, 0}
Variant g_variants[] = {
	MB_VARIANT(48, 2, true, true),    // CONFIG_0..6   BPSK
	MB_VARIANT(24, 4, true, true),    // CONFIG_7..9,12 QPSK
	MB_VARIANT(16, 8, true, true),    // CONFIG_10,11,14 8PSK
	MB_VARIANT(12, 16, true, false),  // CONFIG_13     16QAM, LS
	MB_VARIANT(12, 16, false, false), // CONFIG_15     16QAM, ZF
	MB_VARIANT(9, 32, false, false),  // CONFIG_16     32QAM, ZF
};
int g_num_sms = 0;

Variant *find_variant(const MbMode &m, int narrow = 0)
{
	for (Variant &v : g_variants)
		if (v.S == m.Nsymb && v.M == m.M && v.ls == (m.estimator == 1) && v.phase == (m.phase_only != 0) && v.narrow == narrow) return &v;
	return nullptr;
}

}  // namespace

cudaError_t mb_demod_init()
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
	if (e != cudaSuccess) return e;
	for (Variant &v : g_variants) {
		e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem);
		if (e != cudaSuccess) return e;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v.ctas_per_sm, v.fn, v.threads, (size_t)v.smem);
		if (e != cudaSuccess) return e;
		if (v.ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
	}
	return cudaSuccess;
}

int mb_demod_ctas_per_sm(int Nsymb, int M, int estimator, int phase_only)
{
	MbMode m;
	m.Nsymb = Nsymb, m.M = M, m.estimator = estimator, m.phase_only = phase_only;
	const Variant *v = find_variant(m);
	return v ? v->ctas_per_sm : 0;
}

cudaError_t mb_launch_demod(const MbDemodArgs &a, size_t n_frames, cudaStream_t stream)
{
	if (n_frames == 0) return cudaSuccess;
	const Variant *v = find_variant(a.mode, a.x_format != 0 ? 1 : 0);
	if (!v || g_num_sms <= 0) return cudaErrorInvalidDeviceFunction;
	// bulk copies need 16-byte aligned global addresses; frames are multiples of 16 bytes, so only the bases matter
	if ((reinterpret_cast<uintptr_t>(a.x) & 15u) || (reinterpret_cast<uintptr_t>(a.llr) & 15u)) return cudaErrorMisalignedAddress;
	MbDemodArgs args = a;
	args.n_frames = (unsigned long long)n_frames;
	void *params[] = {&args};
	return cudaLaunchKernel(v->fn, dim3((unsigned)n_frames), dim3((unsigned)v->threads), params, (size_t)v->smem, stream);
}
