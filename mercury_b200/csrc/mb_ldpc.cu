// mb_ldpc.cu -- K_ldpc: batched LDPC belief-propagation decoder + de-scramble + pack + CRC16 for sm_100a.
//
// Reference (what is computed; paths relative to /root/reference/source/physical_layer):
//   cl_ldpc::decode -> decode_SPA   ldpc.cc:266-278, ldpc_decoder_SPA.cc:25-218  (flooding schedule:
//       initial syndrome -> 0 iterations; check update R = 2 atanh(prod tanh(Q/2)) with the +-1 -> +-0.9999999
//       clamp; posterior; syndrome + early exit; Q = posterior - R; returns I+1 when not converged)
//   bit_energy_dispersal interleaver.cc:111-117, bit_to_byte misc.cc:107-130 (LSB first),
//   all-zeros test + CRC16 self check + decision telecom_system.cc:1319-1349, SNR report :1368-1375
//
// How (B200-first):
//   * TWO frames per thread.  A CTA (8 warps) decodes a PAIR of frames in lock step; every per-edge quantity is a float2
//     (x = frame in slot A, y = frame in slot B) interleaved in shared memory, so each index load, address computation,
//     shared-memory access (LDS.64 / STS.64), loop counter and branch is shared by the two frames, and the arithmetic runs on the
//     packed fp32 pipe (FADD2 / FMUL2 / FFMA2: one issue slot for both frames).  The kernel is bound by instruction count and
//     latency (issue slots 61 % busy, XU 54 %), not by memory, so this -- with the check-node arithmetic below -- is what sets its speed.
//   * The two slots are independent decodes: each has its own iteration counter, early exit and verdict.  When a slot finishes
//     (converged, or ran out of iterations) its epilogue runs and the slot is REFILLED with the next frame of the batch from a
//     global queue (one atomic per frame), while the other slot carries on: no frame waits for its neighbour, and a batch with a
//     few non-converging frames (50 iterations against ~5) does not leave SMs idle behind them.  Grid = resident CTAs only.
//     The hand-over requests the next frame's data first and runs the old frame's epilogue while it travels (refill_slot()).
//   * The decoder state of a pair lives in shared memory for all iterations: posterior[1601] + one message per Tanner-graph edge
//     slot, x2 frames (59-77 KB per pair, 3 pairs resident per SM up to rate 8/16); the channel LLRs, read once per iteration and
//     variable, in a per-CTA global scratch that stays in the L2.  HBM is touched once per frame: 6.4 KB of LLRs in, <= 175 + 32 bytes out.
//   * Only posterior and check->variable messages are stored: the variable->check message of the reference (its Q array) is
//     recomputed as posterior - R, which is exactly the reference's Q update.
//   * The Tanner graph is a warp-blocked ELL on BOTH sides (csrc/mb_tables.cpp): checks (variables) sorted by degree, cut into
//     groups of 32 (one warp), padded to the group's largest degree with edges to a +inf posterior (the neutral element of every
//     reduction of the check node), static degree-balanced warp schedules.  Index tables hold BYTE offsets into the interleaved
//     shared arrays and are read through L1.
//   * The syndrome of iteration i is evaluated inside the check pass of iteration i+1 (it gathers the same posteriors anyway).
//   * Sum-product check node in the "e-domain", all state in base-2 units (LLR * log2 e, so ex2 / lg2 need no scaling):
//         e_k = 2^-|q_k|                                (1 MUFU; tanh(|q|/2) = (1 - e)/(1 + e))
//         (P, M) <- (P + e M, M + e P)  from (1, 0)      (the tanh addition rule: M/P = tanh(sum atanh e_k); all terms positive)
//         |R_k| = log2(P_k / M_k)  over the OTHER edges  (2 MUFU: rcp, lg2)
//     which is 2 atanh(prod_{j != k} tanh(|q_j|/2)) exactly, at 3 MUFU per edge and iteration instead of the 6 of the log-domain
//     form (phi forward + phi backward), with no cancellation anywhere: leave-one-out = prefix (+) suffix pairs in registers,
//     one fully unrolled body per edge count (3..8); checks above 7 edges are split over 2 / 4 / 8 lanes that exchange their (P, M)
//     totals by shuffles, the other lanes' total entering as one more edge e = M / P -- so split and unsplit checks run the same bodies.
//     It reproduces the reference's DOUBLE-precision clamp rule (ldpc_decoder_SPA.cc:147-155): a factor whose tanh rounds to 1.0
//     in double (e < 2^-55) contributes e = 0, and an all-saturated product (M == 0) yields 2 atanh(0.9999999).
//     tools/ldpc_numerics.py: on frames at the decoding threshold this arithmetic agrees with the double-precision reference on
//     converged / not converged and on the iteration count as the former log-domain form did (100 % of the frames sampled).
//   * MINSUM mode (north_star): normalised min-sum (alpha 1 for degree<=2 where min-sum is exact, 0.85 for 3, 0.75 above), same
//     schedule / exit / clamp, same pair structure.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "mb_kernels.cuh"

namespace {

constexpr int kThreads = 32 * MB_LDPC_WARPS;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kClampR2 = 16.811242831518264f * kLog2e;  // 2 atanh(0.9999999) (ldpc_decoder_SPA.cc:147-155) in base-2 units
constexpr float kSatQ2 = 55.0f;                           // |q| (base-2 units) above which tanh(|q|/2) == 1.0 in double: e = 2^-|q| < 2^-55

typedef float2 f2;

// shared-memory layout (bytes from the start of the dynamic segment); constant offsets keep every hot access [reg + imm].
// Every array is float2: .x = slot A, .y = slot B.
constexpr int kOffBytes = 0;                       // u8[256] packed bytes of the frame being finished; [255] = its verdict
constexpr int kOffCnt = 256;                       // u32[2][MB_LDPC_WARPS] per-warp vote counts (check pass, syndrome test): A | B << 16
constexpr int kOffNext = 384;                      // int[2] queue hand-off: tickets of the next two refills, alternating
constexpr int kOffCsched = 512;                    // u32[MB_LDPC_WARPS][MB_SCHED_LEN] check tasks per warp
constexpr int kOffVsched = kOffCsched + MB_LDPC_WARPS * MB_SCHED_LEN * 4;  // ... variable groups per warp
constexpr int kOffLam = kOffVsched + MB_LDPC_WARPS * MB_SCHED_LEN * 4;     // f2[1601] posterior; [1600] = +inf, the variable every padding slot points at
// The channel LLRs are NOT in shared memory: they are read once per iteration and variable, in order, so they live in a per-CTA global
// scratch (L2 resident, 12.8 KB per pair, written at refill) and are streamed past L1.  That keeps three pairs per SM inside 196 KB of
// shared memory and leaves the L1 60 KB instead of 28 KB: the index tables (22-28 KB per rate, gathered by every warp in every pass)
// stay L1 resident (ncu: 72 % -> L1 hit rate, long-scoreboard stalls on the index loads).
#ifndef MB_LDPC_LCH_SMEM
#define MB_LDPC_LCH_SMEM 0
#endif
constexpr int kOffLch = kOffLam + (MB_N + 2) * 8;  // f2[1600] channel LLR (only with MB_LDPC_LCH_SMEM)
constexpr int kOffR = kOffLch + (MB_LDPC_LCH_SMEM ? MB_N * 8 : 0);  // f2[c_slots + 1] check -> variable message per check-side slot; [c_slots] == 0 always
static_assert(2 * MB_LDPC_WARPS * 4 <= kOffNext - kOffCnt && kOffLam % 16 == 0, "decoder shared-memory layout");

__host__ __device__ constexpr int r_bytes(int c_slots) { return ((c_slots + 2) & ~1) * 8; }

// The decoder's shared memory is ONE dynamic window (no static __shared__ in this file) whose shared-space address is a constant of
// the toolchain: kSmemBase (the first KB of the window is the system's).  Every gather is therefore [index register + immediate]
// -- no address add per access (ncu r2p: 4.6 % of the decoder's instructions were those adds).  The kernel traps at entry if the
// constant is ever not what the compiler produced.
constexpr unsigned kSmemBase = 0x400;
extern __shared__ __align__(16) unsigned char mb_smem[];
template <int IMM>
__device__ __forceinline__ f2 lds2i(unsigned reg)
{
	f2 v;
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(reg), "n"(IMM));
	return v;
}
__device__ __forceinline__ f2 lds2(unsigned addr) { return lds2i<0>(addr); }
__device__ __forceinline__ void sts2(unsigned addr, f2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory"); }

__device__ __forceinline__ float ldg_stream(const float *p)
{
	float v;
	asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ unsigned fbits(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float ex2_negabs(float x)
{
	float r;
	asm("{\n\t.reg .f32 t;\n\tabs.f32 t, %1;\n\tneg.f32 t, t;\n\tex2.approx.ftz.f32 %0, t;\n\t}" : "=f"(r) : "f"(x));  // -|x| folds into the MUFU operand
	return r;
}
__device__ __forceinline__ float rcp_fast(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ float lg2_fast(float x)
{
	float r;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
// (a & 0x7fffffff) | (s & 0x80000000): one LOP3
__device__ __forceinline__ float copysign_bits(float mag, unsigned s) { return __uint_as_float((fbits(mag) & 0x7fffffffu) | (s & 0x80000000u)); }

// e = 2^-|q| for both frames, flushed to 0 below 2^-55 (the factor's tanh is 1.0 in double): scaling by 2^-71 under FTZ drops
// exactly those values and is exact for the others (two packed multiplies instead of two compare + select pairs).
__device__ __forceinline__ f2 edge_e(f2 q)
{
	const f2 e = make_float2(ex2_negabs(q.x), ex2_negabs(q.y));
	f2 t;
	asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\tmul.rn.ftz.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
	    : "=f"(t.x), "=f"(t.y)
	    : "f"(e.x), "f"(e.y), "f"(4.235164736e-22f /* 2^-71 */));
	return __fmul2_rn(t, make_float2(2.361183241e21f, 2.361183241e21f) /* 2^71 */);
}

// (P, M) pairs of the tanh addition rule, packed over the two frames
struct Gen {
	f2 P, M;
};
__device__ __forceinline__ Gen comb_ss(f2 a, f2 b) { return Gen{__ffma2_rn(a, b, make_float2(1.f, 1.f)), __fadd2_rn(a, b)}; }
__device__ __forceinline__ Gen comb_gs(Gen g, f2 e) { return Gen{__ffma2_rn(e, g.M, g.P), __ffma2_rn(e, g.P, g.M)}; }
__device__ __forceinline__ Gen comb_gg(Gen a, Gen b)
{
	return Gen{__ffma2_rn(a.M, b.M, __fmul2_rn(a.P, b.P)), __ffma2_rn(a.M, b.P, __fmul2_rn(a.P, b.M))};
}

// |R| = log2(P / M) of the other edges, M == 0 (every other factor saturated) -> the reference's clamp; sign = product of the other
// edges' signs = parity of all of them ^ this edge's own.  The sign goes on as one packed multiply by (+-1, +-1): sgn_k = (q_k & sign
// bit) ^ one, where `one` is 1.0f carrying the parity of the whole check in its sign bit (one LOP3 per frame, not xor + copysign).
__device__ __forceinline__ void emit(unsigned addr, Gen o, float qa, float qb, unsigned one_a, unsigned one_b)
{
	const f2 ratio = __fmul2_rn(o.P, make_float2(rcp_fast(o.M.x), rcp_fast(o.M.y)));
	float lx = lg2_fast(ratio.x), ly = lg2_fast(ratio.y);
	lx = lx < 3.0e38f ? lx : kClampR2;
	ly = ly < 3.0e38f ? ly : kClampR2;
	const f2 sgn = make_float2(__uint_as_float((fbits(qa) & 0x80000000u) ^ one_a), __uint_as_float((fbits(qb) & 0x80000000u) ^ one_b));
	sts2(addr, __fmul2_rn(make_float2(lx, ly), sgn));
}

// Shuffle helpers for checks split over S = 2, 4, 8 lanes (lanes of one check are 32 / S apart: xor masks 32 / S, .., 16).
__device__ __forceinline__ Gen shfl_gen(Gen g, int mask)
{
	Gen r;
	r.P.x = __shfl_xor_sync(0xffffffffu, g.P.x, mask), r.P.y = __shfl_xor_sync(0xffffffffu, g.P.y, mask);
	r.M.x = __shfl_xor_sync(0xffffffffu, g.M.x, mask), r.M.y = __shfl_xor_sync(0xffffffffu, g.M.y, mask);
	return r;
}
// (+) of the totals T of all OTHER lanes of this lane's check (xor butterfly: the partner's total, then the other pair's, then the other quad's)
__device__ __forceinline__ Gen lanes_outside(Gen T, int log2s)
{
	int mask = 32 >> log2s;
	Gen out = shfl_gen(T, mask);
	if (log2s > 1) {
		Gen U = comb_gg(T, out);
		mask <<= 1;
		const Gen U2 = shfl_gen(U, mask);
		out = comb_gg(out, U2);
		if (log2s > 2) {
			U = comb_gg(U, U2);
			out = comb_gg(out, shfl_gen(U, mask << 1));
		}
	}
	return out;
}
__device__ __forceinline__ unsigned lanes_xor(unsigned x, int log2s)
{
	for (int mask = 32 >> log2s; mask < 32; mask <<= 1) x ^= __shfl_xor_sync(0xffffffffu, x, mask);
	return x;
}

// Sum-product check node of one lane for a PAIR of frames, D edges (compile time, 3..MB_LDPC_DMAX + 1): reads the posteriors and
// messages, writes the new messages, XORs the hard decisions into hard_a / hard_b (sign bit).  Leave-one-out by prefix and suffix
// (P, M) pairs held in registers; the first combinations are specialised ((1, 0) and (1, e) need no multiplies).
// log2s > 0 (warp uniform): the check is spread over 2^log2s lanes (mb_tables.h: mb_ldpc_split), D - 1 real edges in this lane.  The
// lanes exchange their totals, sign parities and hard-decision parities by shuffles, and the (+) of the OTHER lanes' totals enters this
// lane's leave-one-out as one more edge, e = M / P (a (P, M) pair and the single term (1, M / P) are the same up to a factor that
// cancels in every ratio).  So split and unsplit checks run the SAME body: the decoder's hot code has to stay inside the 32 KB
// instruction cache (ncu: with separate bodies the warps stalled on instruction fetch as often as on the barriers).
template <int D>
__device__ __forceinline__ void spa_check_pair(unsigned sbase, const uint16_t *__restrict__ ve, unsigned raddr, int log2s, f2 nm, unsigned &hard_a, unsigned &hard_b)
{
	f2 q[D], e[D];
	unsigned pa = 0, pb = 0, ha = 0, hb = 0;
	const bool split = D >= 5 && log2s != 0;  // split tasks hold 4..MB_LDPC_DMAX edges per lane
#pragma unroll
	for (int k = 0; k < D; k++) {
		if (k == D - 1 && split) break;
		const unsigned off = ve[k * 32];
		const f2 lam = lds2i<kSmemBase + kOffLam>(off);
		const f2 r = lds2(raddr + k * 256);
		q[k] = __ffma2_rn(r, nm, lam);  // nm = -1, or 0 for a slot whose messages are not written yet (just refilled)
		ha ^= fbits(lam.x);
		hb ^= fbits(lam.y);
		pa ^= fbits(q[k].x);
		pb ^= fbits(q[k].y);
		e[k] = edge_e(q[k]);
	}
	if (split) {
		// sign parities and hard-decision parities of the whole check: one word per frame through the butterfly (bit 31: q, bit 30: posterior)
		pa = lanes_xor((pa & 0x80000000u) | ((ha >> 1) & 0x40000000u), log2s);
		pb = lanes_xor((pb & 0x80000000u) | ((hb >> 1) & 0x40000000u), log2s);
		ha = pa << 1, hb = pb << 1;
		Gen tot = comb_ss(e[0], e[1]);
#pragma unroll
		for (int k = 2; k < D - 1; k++) tot = comb_gs(tot, e[k]);
		const Gen out = lanes_outside(tot, log2s);
		e[D - 1] = __fmul2_rn(out.M, make_float2(rcp_fast(out.P.x), rcp_fast(out.P.y)));  // P >= 1
		q[D - 1] = make_float2(0.f, 0.f);
	}
	hard_a ^= ha, hard_b ^= hb;
	const unsigned one_a = (pa & 0x80000000u) | 0x3f800000u, one_b = (pb & 0x80000000u) | 0x3f800000u;
	Gen pre[D];  // pre[k] = e_0 (+) .. (+) e_{k-1}, k >= 2
	pre[2] = comb_ss(e[0], e[1]);
#pragma unroll
	for (int k = 3; k < D; k++) pre[k] = comb_gs(pre[k - 1], e[k - 1]);
	if (!split) emit(raddr + (D - 1) * 256, pre[D - 1], q[D - 1].x, q[D - 1].y, one_a, one_b);
	if (D >= 4)
		emit(raddr + (D - 2) * 256, comb_gs(pre[D >= 4 ? D - 2 : 2], e[D - 1]), q[D - 2].x, q[D - 2].y, one_a, one_b);
	else
		emit(raddr + 256, comb_ss(e[0], e[2]), q[1].x, q[1].y, one_a, one_b);
	Gen suf = comb_ss(e[D - 2], e[D - 1]);
#pragma unroll
	for (int k = D - 3; k >= 0; k--) {
		const Gen o = k >= 2 ? comb_gg(pre[k >= 2 ? k : 2], suf) : (k == 1 ? comb_gs(suf, e[0]) : suf);
		emit(raddr + k * 256, o, q[k].x, q[k].y, one_a, one_b);
		if (k > 0) suf = comb_gs(suf, e[k]);
	}
}

// Degree-2 check (two thirds of the checks of the low-rate codes): R_a = 2 atanh(tanh(q_b / 2)) is q_b itself until tanh rounds to 1.0
// in double, where the reference's clamp gives 2 atanh(0.9999999) (ldpc_decoder_SPA.cc:147-155).  A padding edge (q = +inf) saturates
// to the clamp: exactly the reference's empty product of a degree-1 check.
__device__ __forceinline__ float sat_q(float q) { return fabsf(q) > kSatQ2 ? copysign_bits(kClampR2, fbits(q)) : q; }
__device__ __forceinline__ void spa_check_pair_2(unsigned sbase, const uint16_t *__restrict__ ve, unsigned raddr, f2 nm, unsigned &hard_a, unsigned &hard_b)
{
	const f2 l0 = lds2i<kSmemBase + kOffLam>(ve[0]), l1 = lds2i<kSmemBase + kOffLam>(ve[32]);
	const f2 q0 = __ffma2_rn(lds2(raddr), nm, l0), q1 = __ffma2_rn(lds2(raddr + 256), nm, l1);
	hard_a ^= fbits(l0.x) ^ fbits(l1.x);
	hard_b ^= fbits(l0.y) ^ fbits(l1.y);
	sts2(raddr, make_float2(sat_q(q1.x), sat_q(q1.y)));
	sts2(raddr + 256, make_float2(sat_q(q0.x), sat_q(q0.y)));
}

// Normalised min-sum check node of one lane for a pair of frames (alpha 1 for true degree <= 2 where min-sum is exact, 0.85 for 3,
// 0.75 above), d edges per lane (warp uniform, <= MB_LDPC_DMAX); a check split over 2^log2s lanes merges the lanes' two smallest
// magnitudes and sign parities by shuffles.
struct MinSum {
	float m1 = 3.0e38f, m2 = 3.0e38f;
	int arg = -1;
	unsigned signs = 0u;
	unsigned par = 0;
	__device__ __forceinline__ void take(float q, int k)
	{
		par ^= fbits(q);
		signs |= (fbits(q) >> 31) << k;
		const float aq = fabsf(q);
		const bool lt1 = aq < m1;
		m2 = lt1 ? m1 : fminf(m2, aq);
		arg = lt1 ? k : arg;
		m1 = lt1 ? aq : m1;
	}
	// the smallest magnitude over the OTHER lanes of the check (-> o1), and the sign parity over ALL its lanes
	__device__ __forceinline__ void exchange(int log2s, float &o1)
	{
		o1 = 3.0e38f;
		float u1 = m1;
		for (int mask = 32 >> log2s; mask < 32; mask <<= 1) {
			const float t1 = __shfl_xor_sync(0xffffffffu, u1, mask);
			o1 = fminf(o1, t1), u1 = fminf(u1, t1);
			par ^= __shfl_xor_sync(0xffffffffu, par, mask);
		}
	}
	__device__ __forceinline__ float give(int k, float alpha, float o1) const
	{
		const float mag = fminf(alpha * fminf(k == arg ? m2 : m1, o1), kClampR2);
		return ((par >> 31) ^ ((signs >> k) & 1u)) ? -mag : mag;
	}
};
__device__ __forceinline__ void minsum_check_pair(unsigned sbase, const uint16_t *__restrict__ ve, unsigned raddr, int d, int log2s, int dc, f2 nm,
						  unsigned &hard_a, unsigned &hard_b)
{
	MinSum a, b;
	unsigned ha = 0, hb = 0;
#pragma unroll 4
	for (int k = 0; k < d; k++) {
		const f2 lam = lds2i<kSmemBase + kOffLam>(ve[k * 32]);
		const f2 q = __ffma2_rn(lds2(raddr + k * 256), nm, lam);
		ha ^= fbits(lam.x);
		hb ^= fbits(lam.y);
		a.take(q.x, k);
		b.take(q.y, k);
	}
	float oa = 3.0e38f, ob = 3.0e38f;
	if (log2s > 0) {
		a.exchange(log2s, oa);
		b.exchange(log2s, ob);
		ha = lanes_xor(ha, log2s), hb = lanes_xor(hb, log2s);
	}
	hard_a ^= ha, hard_b ^= hb;
	const float alpha = dc <= 2 ? 1.0f : (dc == 3 ? 0.85f : 0.75f);
#pragma unroll 4
	for (int k = 0; k < d; k++) sts2(raddr + k * 256, make_float2(a.give(k, alpha, oa), b.give(k, alpha, ob)));
}

// Variable node of the head (degree > 2, groups padded to an even degree): channel LLR + the incoming messages in the table's order.
// D > 0: fixed group degree, every index load and gather at an immediate offset.
template <int D>
__device__ __forceinline__ f2 var_node_sum(unsigned sbase, const uint16_t *__restrict__ se, f2 acc, int d_rt, f2 vm)
{
	if (D > 0) {
		unsigned idx[D > 0 ? D : 1];
#pragma unroll
		for (int k = 0; k < D; k++) idx[k] = se[k * 32];
#pragma unroll
		for (int k = 0; k < D; k++) acc = __ffma2_rn(lds2i<kSmemBase + kOffR>(idx[k]), vm, acc);  // vm = 1, or 0 for a just-refilled slot
		return acc;
	}
	int k = 0;
#pragma unroll 1
	for (; k + 4 <= d_rt; k += 4) {
		const unsigned i0 = se[0], i1 = se[32], i2 = se[64], i3 = se[96];
		se += 128;
		acc = __ffma2_rn(lds2i<kSmemBase + kOffR>(i0), vm, acc);
		acc = __ffma2_rn(lds2i<kSmemBase + kOffR>(i1), vm, acc);
		acc = __ffma2_rn(lds2i<kSmemBase + kOffR>(i2), vm, acc);
		acc = __ffma2_rn(lds2i<kSmemBase + kOffR>(i3), vm, acc);
	}
	if (k < d_rt) {
		const unsigned i0 = se[0], i1 = se[32];
		acc = __ffma2_rn(lds2i<kSmemBase + kOffR>(i0), vm, acc);
		acc = __ffma2_rn(lds2i<kSmemBase + kOffR>(i1), vm, acc);
	}
	return acc;
}

// XOR of the hard decisions of one check (syndrome-only test), both frames; D edges per lane, the check spread over 2^log2s lanes
template <int D>
__device__ __forceinline__ void check_parity(unsigned sbase, const uint16_t *__restrict__ ve, int log2s, unsigned &bad_a, unsigned &bad_b)
{
	unsigned idx[D], ha = 0, hb = 0;
#pragma unroll
	for (int k = 0; k < D; k++) idx[k] = ve[k * 32];
#pragma unroll
	for (int k = 0; k < D; k++) {
		const f2 lam = lds2i<kSmemBase + kOffLam>(idx[k]);
		ha ^= fbits(lam.x);
		hb ^= fbits(lam.y);
	}
	if (log2s > 0) ha = lanes_xor(ha, log2s), hb = lanes_xor(hb, log2s);
	bad_a |= ha >> 31;
	bad_b |= hb >> 31;
}

// Shared-memory views: every address is mb_smem + a compile-time constant (the noinline epilogue / refill used to chase these pointers
// through a struct in local memory, one dependent load per store).  Passed BY VALUE: the only datum is the scratch pointer.
struct Smem {
	static constexpr unsigned sbase = kSmemBase;
	float *lch_g;           // this CTA's channel-LLR scratch in global memory, f2[1600] (x = slot A, y = slot B)
	__device__ __forceinline__ static unsigned char *bytes() { return mb_smem + kOffBytes; }
	__device__ __forceinline__ static unsigned *cnt() { return reinterpret_cast<unsigned *>(mb_smem + kOffCnt); }
	__device__ __forceinline__ static volatile int *next() { return reinterpret_cast<volatile int *>(mb_smem + kOffNext); }
	__device__ __forceinline__ static uint32_t *csched() { return reinterpret_cast<uint32_t *>(mb_smem + kOffCsched); }
	__device__ __forceinline__ static uint32_t *vsched() { return reinterpret_cast<uint32_t *>(mb_smem + kOffVsched); }
	__device__ __forceinline__ static float &lam(int v, int X) { return reinterpret_cast<float *>(mb_smem + kOffLam)[2 * v + X]; }
	__device__ __forceinline__ static float &R(int i, int X) { return reinterpret_cast<float *>(mb_smem + kOffR)[2 * i + X]; }
	__device__ __forceinline__ void set_lch(int v, int X, float w) const
	{
		if (MB_LDPC_LCH_SMEM) reinterpret_cast<float *>(mb_smem + kOffLch)[2 * v + X] = w;
		else lch_g[2 * v + X] = w;
	}
	__device__ __forceinline__ f2 get_lch(int v) const
	{
		if (MB_LDPC_LCH_SMEM) return lds2(sbase + kOffLch + v * 8);
		f2 r;
		asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(lch_g + 2 * v));
		return r;
	}
};
// The fields of the launch arguments the epilogue / refill need right away, by value (registers): through the reference they are
// generic loads from the parameter block, each one a dependent round trip before the first useful request leaves.
struct Hot {
	const float *llr;
	MbRxStats *stats;
	unsigned *queue;
	unsigned long long n_frames;
	int check_gate;
};

// ---- the frame's record: thread 0 folds the per-warp CRC words the epilogue left in shared memory (the CRC is linear: XOR over the
// message's set bits of a per-bit table, MbMode::off_crcbit) --------------------------------------------------------------------------
__device__ __forceinline__ void crc_and_record(const MbLdpcArgs &a, const Smem s, const Hot h, size_t frame, int iterations, int nonzero)
{
	if (threadIdx.x != 0) return;
	const unsigned *crcw = reinterpret_cast<const unsigned *>(s.bytes());
	unsigned adv = 0;
#pragma unroll
	for (int w = 0; w < MB_LDPC_WARPS; w++) adv ^= crcw[w];
	// only the fields the decoder owns are written (no read-modify-write: the record's load would sit on this warp's path to the refill)
	const int all_zeros = nonzero ? 0 : 1;
	const int crc = all_zeros ? 0 : (int)((adv ^ a.mode.crc_init) & 0xFFFFu);  // telecom_system.cc:1337-1341
	const int decoded = (!all_zeros && crc == 0) ? 1 : 0;                    // telecom_system.cc:1343-1349
	MbRxStats *rec = h.stats + frame;
	*reinterpret_cast<int4 *>(rec) = make_int4(iterations, crc, all_zeros, decoded);  // iterations_done, crc, all_zeros, message_decoded
	if (!decoded) rec->SNR = -99.9f;
	s.bytes()[255] = (unsigned char)decoded;
}

// ---- hard decision -> de-scramble -> pack LSB first -> payload bytes + per-warp CRC words of slot X; returns "any bit set" ------------
// Called by the whole CTA (uniform, one barrier inside); the other slot's interleaved state is not touched.
__device__ __forceinline__ int hard_decisions(const MbLdpcArgs &a, const Smem s, int X, size_t frame)
{
	const MbMode &m = a.mode;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint16_t *__restrict__ g_bit_var = reinterpret_cast<const uint16_t *>(a.blob + m.off_bit_var);
	const uint8_t *__restrict__ g_scr = a.blob + m.off_scr;
	const uint16_t *__restrict__ g_crcbit = reinterpret_cast<const uint16_t *>(a.blob + m.off_crcbit);
	unsigned byte = 0, crcw = 0;
	if (tid < m.crc_bytes) {
		unsigned tb[8];
#pragma unroll
		for (int b = 0; b < 8; b++) tb[b] = g_crcbit[tid * 8 + b];  // requested with the bit tables, not after the decisions
#pragma unroll
		for (int b = 0; b < 8; b++) {
			const int i = tid * 8 + b;
			const unsigned bit = (fbits(s.lam(g_bit_var[i], X)) >> 31) ^ (unsigned)g_scr[i];  // hard decision = sign bit (LLR < 0)
			byte |= bit << b;
			crcw ^= (0u - bit) & tb[b];
		}
		if (tid < m.frame_bytes) a.payload[frame * (size_t)m.frame_bytes + tid] = (uint8_t)byte;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) crcw ^= __shfl_xor_sync(0xffffffffu, crcw, o);
	if (lane == 0) reinterpret_cast<unsigned *>(s.bytes())[warp] = crcw;
	return __syncthreads_or((int)byte);
}

// ---- epilogue of the ZF modes: the above, the record, and the SNR report of a decoded frame -------------------------------------------
// (LS modes: the demodulator's pilot variance is the SNR report, telecom_system.cc:1368-1375; their epilogue runs INSIDE refill_slot(),
// between the requests for the next frame's data and their use.)
__device__ __noinline__ void finish_slot_zf(const MbLdpcArgs &a, const Smem s, const Hot h, int X, size_t frame, int iterations)
{
	const MbMode &m = a.mode;
	const MbRate &rt = a.rate;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	crc_and_record(a, s, h, frame, iterations, hard_decisions(a, s, X, frame));

	// ---- ZF modes: SNR report of a decoded frame (telecom_system.cc:1376-1400) -----------------------------------------
	// Re-encode the hard decisions (scrambled info bits, virtual copies, IRA parity), re-map them onto the constellation through
	// the same composed interleaver records the demodulator scatters by, and measure the mean squared distance of the equalised
	// data symbols (kept by the demodulator behind the LLRs of the hand-off record) to the re-encoded ones (ofdm.cc:1622-1635).
	// Scratch: this slot's messages are dead: [0, P) data parity per check, [P, P + 1600) re-encoded bit per internal variable.
	__syncthreads();
	if (!s.bytes()[255]) return;
	const uint16_t *__restrict__ g_edge_var = a.edge_var;
	const uint16_t *__restrict__ g_voc = reinterpret_cast<const uint16_t *>(a.blob + rt.off_var_of_cw);
	const int K = m.K, P = rt.P, nReal = m.nReal, nVirtual = m.nVirtual;
	auto bit_at = [&](int v) -> unsigned & { return reinterpret_cast<unsigned &>(s.R(P + v, X)); };  // P + 1600 <= edges <= c_slots in all eight codes
	auto dpar_at = [&](int c) -> unsigned & { return reinterpret_cast<unsigned &>(s.R(c, X)); };
	for (int v = tid; v < MB_N; v += kThreads) bit_at(v) = 0u;  // parity variables stay 0 while the data parities are formed
	__syncthreads();
	for (int i = tid; i < nReal; i += kThreads) {  // hd_decoded_data_bit, scrambled back = the decoder's hard decisions (:1378)
		const unsigned b = fbits(s.lam(g_voc[i], X)) >> 31;
		bit_at(g_voc[i]) = b;
		if (i < nVirtual) bit_at(g_voc[nReal + i]) = b;  // virtual bits are copies of the first ones (:1380-1383)
	}
	__syncthreads();
	{  // cl_ldpc::encode (ldpc.cc:111-132): every check row is {data bits, parity i-1, parity i}, so parity = running XOR of the data parities
		const uint16_t *__restrict__ g_cos = reinterpret_cast<const uint16_t *>(a.blob + rt.off_check_of_sorted);
		const uint32_t *sched = s.csched() + warp * MB_SCHED_LEN;
		for (uint32_t desc = *sched; desc != 0u; desc = *++sched) {  // this warp's check tasks, in the check pass's layout
			const int dp = (int)MB_CDESC_DP(desc), l2 = (int)MB_CDESC_LOG2S(desc), per = 32 >> l2;
			const uint16_t *__restrict__ ve = g_edge_var + MB_CDESC_BASE(desc) + lane;
			unsigned x = 0;
			for (int k = 0; k < dp; k++) {
				const int v = (int)(ve[k * 32] >> 3);
				if (v < MB_N) x ^= bit_at(v);  // padding edges name variable N
			}
			x = lanes_xor(x, l2);
			const int c = (int)MB_CDESC_GROUP(desc) * 32 + (int)MB_CDESC_TASK(desc) * per + (lane & (per - 1));
			if (lane < per && c < P) dpar_at(g_cos[c]) = x;
		}
	}
	__syncthreads();
	if (warp == 0) {
		const int chunk = (P + 31) >> 5;
		unsigned x = 0;
		for (int i = lane * chunk; i < min(P, (lane + 1) * chunk); i++) x ^= dpar_at(i);
		unsigned carry = x;  // inclusive XOR scan of the chunk parities over the lanes
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned y = __shfl_up_sync(0xffffffffu, carry, o);
			if (lane >= o) carry ^= y;
		}
		unsigned run = carry ^ x;  // parity of everything before this lane's chunk
		for (int i = lane * chunk; i < min(P, (lane + 1) * chunk); i++) {
			run ^= dpar_at(i);
			bit_at(g_voc[K + i]) = run;
		}
	}
	__syncthreads();
	{
		const uint32_t *__restrict__ g_drec = reinterpret_cast<const uint32_t *>(a.blob + m.off_data_rec);
		const float2 *__restrict__ g_cons = reinterpret_cast<const float2 *>(a.blob + m.off_const);
		const float2 *__restrict__ zf = reinterpret_cast<const float2 *>(a.llr + frame * (size_t)MB_HANDOFF_STRIDE + MB_N);
		const int bps = m.bps, rw = m.data_rec_words;
		float acc = 0.f;
		for (int d = tid; d < m.nData; d += kThreads) {
			unsigned loc = 0;
			for (int e = 0; e < bps; e++) {  // interleaver o psk.mod o interleaver (:1389-1391): bits MSB first
				const uint32_t w = g_drec[d * rw + 1 + (e >> 1)];
				const uint32_t off = (e & 1) ? (w >> 16) : (w & 0xFFFFu);
				loc = (loc << 1) | bit_at(MB_HANDOFF_INV(off >> 2));
			}
			const float2 c = g_cons[loc], z = zf[d];
			acc += (c.x - z.x) * (c.x - z.x) + (c.y - z.y) * (c.y - z.y);
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
		float *s_acc = reinterpret_cast<float *>(s.bytes());  // the packed bytes are dead now
		__syncthreads();
		if (lane == 0) s_acc[warp] = acc;
		__syncthreads();
		if (tid == 0) {
			float tot = 0.f;
			for (int w = 0; w < kThreads / 32; w++) tot += s_acc[w];
			a.stats[frame].SNR = -10.0f * log10f(tot / (float)m.nData);  // measure_SNR, ofdm.cc:1622-1635
		}
	}
}

// ---- next frame of the batch into slot X (uniform).  Returns its index, or -1 when the queue is empty (the slot then holds an
// all-satisfied dummy and is ignored).  Frames the demodulator gated out (mean|H| < 0.3) get their record here and are skipped.
// The queue is read one grab ahead (thread 0 holds the ticket of the NEXT refill in shared memory, so the atomic's latency is never
// waited for), and every grab prefetches into L2 the LLRs of the frame that will be handed out one "wave" of resident slots later
// (frames are handed out in order), so a refill reads L2, not DRAM, while the pair's other slot waits.
// `turn`: which of the two ticket words holds this refill's ticket (uniform; the caller keeps it).  Returns {frame, next turn}.
// done_frame >= 0 (LS modes): the frame leaving the slot; its epilogue (hard decisions, CRC, record) runs here while the requests travel.
__device__ __noinline__ int2 refill_slot(const MbLdpcArgs &a, const Smem s, const Hot h, int X, int turn, long long done_frame = -1, int done_iterations = 0)
{
	const MbMode &m = a.mode;
	const int tid = threadIdx.x;
	constexpr int kPer = (MB_N + kThreads - 1) / kThreads;
	int frame;
	for (;;) {
		// The ticket was published before a barrier every thread has passed (the end of the previous refill / the kernel prologue); the
		// other word takes the ticket of the refill after this one, fetched by a thread of warp 1 (warp 0 may still be in the epilogue's
		// CRC) and published by the barrier at the end of this refill.  No barrier here: the loads below overlap the epilogue's tail.
		// Everything this refill reads from global memory -- the next ticket, the demodulator's gate value, the LLRs -- is requested
		// BEFORE anything is waited for: one L2 round trip on the path to the barrier, not three.
		const unsigned f = (unsigned)s.next()[turn];
		turn ^= 1;
		unsigned nxt = 0u;
		if (tid == 32) nxt = atomicAdd(h.queue, 1u);
		if ((unsigned long long)f >= h.n_frames) {
			if (done_frame >= 0) crc_and_record(a, s, h, (size_t)done_frame, done_iterations, hard_decisions(a, s, X, (size_t)done_frame));
			if (tid == 32) s.next()[turn] = (int)nxt;
			frame = -1;
			break;
		}
		{
			const unsigned long long ahead = (unsigned long long)f + 2ull * gridDim.x;
			if (ahead < h.n_frames && tid < (MB_N * 4 + 127) / 128)
				asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(h.llr + ahead * MB_HANDOFF_STRIDE) + tid * 128));
		}
		// hand-off layout: every 32-float row arrives rotated by its row index (MB_HANDOFF); coalesced loads, conflict-free stores.
		const float *__restrict__ src = h.llr + (size_t)f * MB_HANDOFF_STRIDE;
		float v[kPer];
#pragma unroll
		for (int j = 0; j < kPer; j++) {
			const int i = tid + j * kThreads;
			v[j] = i < MB_N ? ldg_stream(src + i) : 0.f;  // all loads in flight together; no L1 allocation (the index tables live there)
		}
		const float mean_H = h.check_gate ? ldg_stream(&h.stats[f].mean_H) : 1.0f;
		if (done_frame >= 0) crc_and_record(a, s, h, (size_t)done_frame, done_iterations, hard_decisions(a, s, X, (size_t)done_frame));  // while the loads travel
		done_frame = -1;
		if (tid == 32) s.next()[turn] = (int)nxt;
		if (!(mean_H >= 0.3f)) {
			// telecom_system.cc:1268-1280: a channel estimate this weak means a false sync; the reference skips the decode
			for (int i = tid; i < m.frame_bytes; i += kThreads) a.payload[f * (size_t)m.frame_bytes + i] = 0;
			if (tid == 0) {
				MbRxStats *rec = h.stats + f;
				*reinterpret_cast<int4 *>(rec) = make_int4(-1, 0, 0, 0);  // iterations_done, crc, all_zeros, message_decoded
				rec->SNR = -99.9f;
			}
			__syncthreads();  // the next ticket is visible
			continue;
		}
		frame = (int)f;
		// Base-2 units from here on; "+ 0" turns an LLR of -0 into +0 (hard decision 0 like the reference's `< 0` test).
#pragma unroll
		for (int j = 0; j < kPer; j++) {
			const int i = tid + j * kThreads;
			if (i < MB_N) {
				const float w = fmaf(v[j], kLog2e, 0.0f);
				const unsigned p = MB_HANDOFF_INV((unsigned)i);
				s.lam((int)p, X) = w;
				s.set_lch((int)p, X, w);
			}
		}
		break;
	}
	if (frame < 0)
		for (int i = tid; i < MB_N; i += kThreads) s.lam(i, X) = 1.0f, s.set_lch(i, X, 1.0f);
	// the slot's old messages stay where they are: until its first check pass has rewritten them they are multiplied by 0 (the kernel's nm / vm)
	__syncthreads();
	return make_int2(frame, turn);
}

// a finished frame leaves slot X, the next one of the queue enters; returns {the new frame's index (-1: queue empty), next turn}
__device__ __forceinline__ int2 retire_slot(const MbLdpcArgs &a, const Smem s, const Hot h, int X, int turn, int frame, int iterations)
{
	if (a.mode.estimator == 1) return refill_slot(a, s, h, X, turn, (long long)frame, iterations);
	finish_slot_zf(a, s, h, X, (size_t)frame, iterations);
	return refill_slot(a, s, h, X, turn);
}

// Development aid (-DMB_LDPC_TIMING): cycles per phase and warp, printed by the first CTAs (never in the shipped build).
#ifdef MB_LDPC_TIMING
#define MB_T_DECL unsigned t_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_last = (unsigned)clock();
#define MB_T(i) { const unsigned t_now = (unsigned)clock(); t_ph[i] += t_now - t_last; t_last = t_now; }
#else
#define MB_T_DECL
#define MB_T(i)
#endif
// MINCTAS: resident pairs per SM the registers are budgeted for -- MB_LDPC_MIN_CTAS (3: 80 registers) wherever shared memory lets three
// pairs in, 2 (up to 128 registers) for rate 14/16, whose 77 KB per pair leave room for two anyway (18.9 against 19.4 ms per 65,536
// mode-16 frames at 20 iterations).
template <int ALGO, int MINCTAS>
__global__ void __launch_bounds__(kThreads, MINCTAS) mb_ldpc_kernel(const __grid_constant__ MbLdpcArgs a)
{
	if ((unsigned)__cvta_generic_to_shared(mb_smem) != kSmemBase) __trap();  // the immediates of lds2i assume it
	const MbRate &rt = a.rate;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int CS = rt.c_slots;
	Smem s;
	uint32_t *const s_csched = Smem::csched();  // [warps][16] check-task descriptors per warp
	uint32_t *const s_vsched = Smem::vsched();  // [warps][16] variable-group descriptors per warp
	s.lch_g = a.lch_scratch + (size_t)blockIdx.x * (2 * MB_N);
	const Hot hot = {a.llr, a.stats, a.queue, a.n_frames, a.check_gate};
	constexpr unsigned sbase = Smem::sbase;

	const uint16_t *__restrict__ g_edge_var = a.edge_var;  // byte offsets into the posteriors
	const uint16_t *__restrict__ g_vedge = a.vedge;        // byte offsets into the messages
	const uint32_t *__restrict__ g_vtail = a.vtail;        // two byte offsets per degree-<=2 variable

	if (tid < MB_LDPC_WARPS * MB_SCHED_LEN) {
		s_csched[tid] = reinterpret_cast<const uint32_t *>(a.blob + rt.off_csched)[tid];
		s_vsched[tid] = reinterpret_cast<const uint32_t *>(a.blob + rt.off_vsched)[tid];
	}
	if (tid == 0) {
		s.lam(MB_N, 0) = __int_as_float(0x7f800000);  // +inf: a padding edge contributes e = 0, sign +, hard bit 0
		s.lam(MB_N, 1) = __int_as_float(0x7f800000);
	}
	for (int i = tid; i < 2 * (CS + 1); i += kThreads) reinterpret_cast<float *>(mb_smem + kOffR)[i] = 0.f;  // once: 0 x (stale finite message) is 0, 0 x (garbage NaN) is not
	int turn = 0;  // which ticket word the next refill reads (uniform)
	if (tid == 32) s.next()[0] = (int)atomicAdd(a.queue, 1u);
	__syncthreads();
	int frame[2];
	int pass[2] = {0, 0};  // check passes this slot's frame has been through = iterations completed
	bool virgin[2] = {true, true};  // refilled, messages not yet rewritten by a check pass
	{
		int2 r = refill_slot(a, s, hot, 0, turn);
		frame[0] = r.x;
		r = refill_slot(a, s, hot, 1, r.y);
		frame[1] = r.x, turn = r.y;
	}

	const int vtail0 = rt.vtail_start;
	MB_T_DECL
	while (frame[0] >= 0 || frame[1] >= 0) {
		MB_T(7)
		// ---- check pass: syndrome of the current posteriors + new check->variable messages, both slots ----------------
		// A warp owns a group of 32 checks padded to one degree; padding slots point at the +inf variable, so the loops are
		// warp-uniform (no per-thread degree, no divergence) and a padding edge is the neutral element of every reduction.
		unsigned hard_a = 0, hard_b = 0;
		const f2 nm = make_float2(virgin[0] ? 0.f : -1.f, virgin[1] ? 0.f : -1.f);
		virgin[0] = virgin[1] = false;
		// This warp's check tasks (static, cost-balanced schedule), sorted by body: one loop per body, no per-task dispatch.
		const uint32_t *sched = s_csched + warp * MB_SCHED_LEN;
		const uint32_t counts = sched[MB_SCHED_LEN - 1];
#define MB_TASK_PROLOGUE                                                          \
	const uint32_t desc = *sched++;                                           \
	const int l2 = (int)MB_CDESC_LOG2S(desc);                                 \
	const unsigned e0 = MB_CDESC_BASE(desc) + (unsigned)lane;                 \
	const unsigned raddr = sbase + kOffR + e0 * 8u;                           \
	const uint16_t *__restrict__ ve = g_edge_var + e0;                        \
	unsigned ha = 0, hb = 0;
#define MB_TASK_EPILOGUE                \
	hard_a |= ha & 0x80000000u;     \
	hard_b |= hb & 0x80000000u;
		if (ALGO == 0) {
			for (int n = (int)(counts & 15u); n > 0; n--) {
				MB_TASK_PROLOGUE
				(void)l2;
				spa_check_pair_2(sbase, ve, raddr, nm, ha, hb);
				MB_TASK_EPILOGUE
			}
#define MB_BODY_LOOP(B_)                                                                    \
	for (int n = (int)((counts >> (4 * (B_ - 2))) & 15u); n > 0; n--) {                 \
		MB_TASK_PROLOGUE                                                            \
		spa_check_pair<B_>(sbase, ve, raddr, l2, nm, ha, hb);                       \
		MB_TASK_EPILOGUE                                                            \
	}
			MB_BODY_LOOP(3)
			MB_BODY_LOOP(4)
			MB_BODY_LOOP(5)
			MB_BODY_LOOP(6)
			MB_BODY_LOOP(7)
			MB_BODY_LOOP(8)
#undef MB_BODY_LOOP
			static_assert(MB_LDPC_DMAX == 7, "check-node bodies are instantiated for 3..MB_LDPC_DMAX + 1 edges");
		} else {
			for (uint32_t d0 = *sched; d0 != 0u; d0 = *sched) {
				MB_TASK_PROLOGUE
				const int dp = (int)MB_CDESC_DP(desc), per = 32 >> l2;
				const int c = (int)MB_CDESC_GROUP(desc) * 32 + (int)MB_CDESC_TASK(desc) * per + (lane & (per - 1));
				const int dc = c < rt.P ? (int)(a.blob + rt.off_cdeg)[c] : 0;  // the true degree picks the normalisation
				minsum_check_pair(sbase, ve, raddr, dp, l2, dc, nm, ha, hb);
				MB_TASK_EPILOGUE
			}
		}
#undef MB_TASK_PROLOGUE
#undef MB_TASK_EPILOGUE
		// threads that saw an unsatisfied check, per slot: 0 = converged; a small count = "probably one iteration to go"
		{
			const unsigned ba = __ballot_sync(0xffffffffu, hard_a != 0u), bb = __ballot_sync(0xffffffffu, hard_b != 0u);
			if (lane == 0) s.cnt()[warp] = (unsigned)__popc(ba) | ((unsigned)__popc(bb) << 16);
		}
		MB_T(0)
		__syncthreads();
		MB_T(1)
		unsigned tot = 0;
#pragma unroll
		for (int w = 0; w < MB_LDPC_WARPS; w++) tot += s.cnt()[w];
		int n_unsat[2] = {(int)(tot & 0xFFFFu), (int)(tot >> 16)};
		bool fresh[2] = {false, false};
#pragma unroll
		for (int X = 0; X < 2; X++) {
			if (frame[X] < 0) continue;
			int iterations = -1;
			if (n_unsat[X] == 0)
				iterations = pass[X];  // converged after pass[X] iterations (0 = clean on arrival, ldpc_decoder_SPA.cc:62-77)
			else if (pass[X] == a.max_iters)
				iterations = a.max_iters + 1;  // ldpc_decoder_SPA.cc:127,217: loop ran out
			if (iterations >= 0) {
				MB_T(2)
				const int2 r = retire_slot(a, s, hot, X, turn, frame[X], iterations);
				frame[X] = r.x, turn = r.y;
				MB_T(3)
				pass[X] = 0;
				fresh[X] = true;
				virgin[X] = true;
			} else {
				pass[X]++;
			}
		}
		MB_T(2)
		if (frame[0] < 0 && frame[1] < 0) break;
		// both slots just refilled (the usual case where frames arrive clean: 0 iterations each): their posteriors are the channel LLRs
		// already, there is nothing for a variable pass to do
		if (virgin[0] && virgin[1]) continue;
		// ---- variable pass: posterior = channel + sum of incoming messages (reference V-row order), both slots ----------
		// Variables are numbered by descending degree.  The head (degree > 2) is walked in warp groups padded to one even
		// degree (padding reads the always-zero slot); the long tail of degree-<=2 variables (the accumulator chain of the IRA
		// code, ~60 % of all variables) is a flat loop with both message offsets packed in one word.
		// A slot refilled above keeps its old messages (x 0 here): its posterior is rewritten with the channel LLR, the next check pass is
		// its pass 0.  The channel LLRs come from the L2 (global scratch): all of a thread's loads are issued before the first is used.
		const f2 vm = make_float2(virgin[0] ? 0.f : 1.f, virgin[1] ? 0.f : 1.f);
		constexpr int kTailRegs = 6;  // the tail is at most 1600 - 128 variables (rate 1/16)
		f2 lt[kTailRegs];
#pragma unroll
		for (int i = 0; i < kTailRegs; i++) {
			const int v = vtail0 + tid + i * kThreads;
			lt[i] = v < MB_N ? s.get_lch(v) : make_float2(0.f, 0.f);
		}
		sched = s_vsched + warp * MB_SCHED_LEN;
		uint32_t desc = *sched;
		f2 lh = desc != 0u ? s.get_lch((int)((desc >> 24) - 1u) * 32 + lane) : make_float2(0.f, 0.f);  // first head group: in flight during the tail
		// The channel LLRs arrive from the L2 (~300 cycles): every sum below adds its messages first and the channel value LAST, and the
		// head groups (channel value fetched one group ahead) run before the tail, whose values were requested first.
		while (desc != 0u) {
			const int d = (int)((desc >> 16) & 0xFFu);
			const int v = (int)((desc >> 24) - 1u) * 32 + lane;
			const uint16_t *__restrict__ se = g_vedge + ((desc & 0xFFFFu) + lane);
			const uint32_t nd = *++sched;
			const f2 lcur = lh;
			if (nd != 0u) lh = s.get_lch((int)((nd >> 24) - 1u) * 32 + lane);  // the next group's, one ahead
			f2 acc = make_float2(0.f, 0.f);
			switch (d) {  // variable degrees are 3..9 in all eight codes (padded: 4, 6, 8, 10)
			case 4: acc = var_node_sum<4>(sbase, se, acc, d, vm); break;
			case 6: acc = var_node_sum<6>(sbase, se, acc, d, vm); break;
			case 8: acc = var_node_sum<8>(sbase, se, acc, d, vm); break;
			default: acc = var_node_sum<0>(sbase, se, acc, d, vm); break;
			}
			sts2(sbase + kOffLam + v * 8, __fadd2_rn(acc, lcur));
			desc = nd;
		}
#pragma unroll
		for (int i = 0; i < kTailRegs; i++) {
			const int v = vtail0 + tid + i * kThreads;
			if (v < MB_N) {
				const uint32_t w = __ldg(g_vtail + (v - vtail0));
				const f2 r0 = lds2i<kSmemBase + kOffR>(w & 0xFFFFu), r1 = lds2i<kSmemBase + kOffR>(w >> 16);
				sts2(sbase + kOffLam + v * 8, __ffma2_rn(__fadd2_rn(r0, r1), vm, lt[i]));
			}
		}
		for (int v = vtail0 + tid + kTailRegs * kThreads; v < MB_N; v += kThreads) {  // (not reached with the eight codes of the reference)
			const uint32_t w = __ldg(g_vtail + (v - vtail0));
			const f2 r0 = lds2i<kSmemBase + kOffR>(w & 0xFFFFu), r1 = lds2i<kSmemBase + kOffR>(w >> 16);
			sts2(sbase + kOffLam + v * 8, __ffma2_rn(__fadd2_rn(r0, r1), vm, s.get_lch(v)));
		}
		MB_T(4)
		__syncthreads();
		MB_T(5)
		// ---- cheap syndrome-only test when convergence is likely: saves the (expensive) message update of a final pass ----
		const bool try_a = frame[0] >= 0 && !fresh[0] && n_unsat[0] <= a.cheap_test_threads;
		const bool try_b = frame[1] >= 0 && !fresh[1] && n_unsat[1] <= a.cheap_test_threads;
		if (try_a || try_b) {
			unsigned bad_a = 0, bad_b = 0;
			sched = s_csched + warp * MB_SCHED_LEN;
			const uint32_t pcounts = sched[MB_SCHED_LEN - 1];
#define MB_PARITY_LOOP(B_)                                                                                   \
	for (int n = (int)((pcounts >> (4 * (B_ - 2))) & 15u); n > 0; n--) {                                 \
		const uint32_t desc = *sched++;                                                              \
		const int l2 = (int)MB_CDESC_LOG2S(desc);                                                    \
		const uint16_t *__restrict__ ve = g_edge_var + MB_CDESC_BASE(desc) + lane;                   \
		if (B_ <= MB_LDPC_DMAX && (B_ == 2 || l2 == 0)) check_parity<(B_ <= MB_LDPC_DMAX ? B_ : 2)>(sbase, ve, l2, bad_a, bad_b); \
		else check_parity<B_ - 1>(sbase, ve, l2, bad_a, bad_b);  /* a split task holds one edge less than its body */ \
	}
			MB_PARITY_LOOP(2)
			MB_PARITY_LOOP(3)
			MB_PARITY_LOOP(4)
			MB_PARITY_LOOP(5)
			MB_PARITY_LOOP(6)
			MB_PARITY_LOOP(7)
			MB_PARITY_LOOP(8)
#undef MB_PARITY_LOOP
			const unsigned ba = __ballot_sync(0xffffffffu, bad_a != 0u), bb = __ballot_sync(0xffffffffu, bad_b != 0u);
			if (lane == 0) s.cnt()[MB_LDPC_WARPS + warp] = (unsigned)__popc(ba) | ((unsigned)__popc(bb) << 16);
			__syncthreads();
			unsigned t2 = 0;
#pragma unroll
			for (int w = 0; w < MB_LDPC_WARPS; w++) t2 += s.cnt()[MB_LDPC_WARPS + w];
			const bool done_a = try_a && (t2 & 0xFFFFu) == 0u, done_b = try_b && (t2 >> 16) == 0u;
			if (done_a) {  // exactly what the next check pass would have reported
				const int2 r = retire_slot(a, s, hot, 0, turn, frame[0], pass[0]);
				frame[0] = r.x, turn = r.y;
				pass[0] = 0;
				virgin[0] = true;
			}
			if (done_b) {
				const int2 r = retire_slot(a, s, hot, 1, turn, frame[1], pass[1]);
				frame[1] = r.x, turn = r.y;
				pass[1] = 0;
				virgin[1] = true;
			}
		}
		MB_T(6)
	}
#ifdef MB_LDPC_TIMING
	if (lane == 0 && (blockIdx.x % 111) == 5)
		printf("T cta %d warp %d check %u barA %u decide %u retire %u var %u barB %u cheap %u loop %u\n", (int)blockIdx.x, warp, t_ph[0], t_ph[1],
		       t_ph[2], t_ph[3], t_ph[4], t_ph[5], t_ph[6], t_ph[7]);
#endif
	// the last CTA out re-arms the queue for the next launch on this stream
	__syncthreads();
	if (tid == 0) {
		__threadfence();
		if (atomicAdd(a.queue + 1, 1u) == gridDim.x - 1) {
			a.queue[0] = 0u;
			a.queue[1] = 0u;
			__threadfence();
		}
	}
}

}  // namespace

size_t mb_ldpc_smem_bytes(int c_slots)
{
	return (size_t)kOffR + (size_t)r_bytes(c_slots);
}

namespace {
typedef void (*LdpcKernel)(const MbLdpcArgs);
const LdpcKernel kKernels[2][2] = {{mb_ldpc_kernel<0, MB_LDPC_MIN_CTAS>, mb_ldpc_kernel<0, 2>},   // sum-product: three pairs per SM / two
				   {mb_ldpc_kernel<1, MB_LDPC_MIN_CTAS>, mb_ldpc_kernel<1, 2>}};  // min-sum
int g_ctas_per_sm[2][MB_NRATES] = {};
int g_variant[2][MB_NRATES] = {};  // which of the two register budgets a rate runs with
int g_sms = 0;
}  // namespace

// the shared-space address of the dynamic window in a kernel of this file (no static __shared__): what kSmemBase has to be
__global__ void mb_ldpc_smem_base_probe(unsigned *out) { *out = (unsigned)__cvta_generic_to_shared(mb_smem); }

cudaError_t mb_ldpc_init()
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	{  // fail HERE, with a reason, rather than by the decoder's entry trap, if a toolchain / driver ever moves the window
		unsigned *d = nullptr, hbase = 0;
		if ((e = cudaMalloc(&d, sizeof(unsigned))) != cudaSuccess) return e;
		mb_ldpc_smem_base_probe<<<1, 1, 1024>>>(d);
		e = cudaMemcpy(&hbase, d, sizeof(unsigned), cudaMemcpyDeviceToHost);
		cudaFree(d);
		if (e != cudaSuccess) return e;
		if (hbase != kSmemBase) {
			fprintf(stderr, "mercury_b200: the dynamic shared-memory window starts at 0x%x, the decoder was built for 0x%x (mb_ldpc.cu kSmemBase)\n", hbase, kSmemBase);
			return cudaErrorInvalidDeviceFunction;
		}
	}
	e = cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
	if (e != cudaSuccess) return e;
	for (LdpcKernel k : {kKernels[0][0], kKernels[0][1], kKernels[1][0], kKernels[1][1]}) {
		e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
		if (e != cudaSuccess) return e;
		int carve = MB_LDPC_LCH_SMEM ? (int)cudaSharedmemCarveoutMaxShared : 77;  // 196 KB of the 256 KB array: three pairs of up to 64 KB + 60 KB of L1
		if (const char *v = getenv("MERCURY_B200_LDPC_CARVEOUT")) carve = atoi(v);  // tuning: percent of the unified L1/shared array given to shared memory
		e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
		if (e != cudaSuccess) return e;
	}
	return cudaSuccess;
}

size_t mb_ldpc_max_ctas() { return (size_t)g_sms * 4; }

int mb_ldpc_ctas_per_sm(int algo, int rate_idx, int rate_num, int c_slots)
{
	const int ki = algo != 0 ? 1 : 0;
	(void)rate_num;
	if (g_ctas_per_sm[ki][rate_idx] == 0) {
		int n = 0, v = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kKernels[ki][0], kThreads, mb_ldpc_smem_bytes(c_slots)) != cudaSuccess || n < 1) n = 1;
		if (n <= 2 && MB_LDPC_MIN_CTAS > 2) {  // shared memory admits two pairs: take the kernel whose registers are budgeted for two
			int n2 = 0;
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n2, kKernels[ki][1], kThreads, mb_ldpc_smem_bytes(c_slots)) == cudaSuccess && n2 >= n) v = 1, n = n2;
		}
		g_ctas_per_sm[ki][rate_idx] = n;
		g_variant[ki][rate_idx] = v;
	}
	return g_ctas_per_sm[ki][rate_idx];
}

cudaError_t mb_launch_ldpc(const MbLdpcArgs &a, size_t n_frames, int algo, cudaStream_t stream)
{
	if (n_frames == 0) return cudaSuccess;
	if (!a.queue || !a.lch_scratch || n_frames > 0x7fff0000ull) return cudaErrorInvalidValue;  // frame tickets are 31-bit
	const size_t smem = mb_ldpc_smem_bytes(a.rate.c_slots);
	const int rate_idx = std::max(0, mb_rate_index(a.rate.rate_num));
	const size_t resident = (size_t)g_sms * (size_t)mb_ldpc_ctas_per_sm(algo, rate_idx, a.rate.rate_num, a.rate.c_slots);
	const LdpcKernel k = kKernels[algo != 0 ? 1 : 0][g_variant[algo != 0 ? 1 : 0][rate_idx]];
	MbLdpcArgs b = a;
	b.n_frames = n_frames;
	b.edge_var = reinterpret_cast<const uint16_t *>(a.blob + a.rate.off_edge_varb);
	b.vedge = reinterpret_cast<const uint16_t *>(a.blob + a.rate.off_vedgeb);
	b.vtail = reinterpret_cast<const uint32_t *>(a.blob + a.rate.off_vtail);
	const size_t grid = std::min(std::min(resident, mb_ldpc_max_ctas()), (n_frames + 1) / 2);
	k<<<(unsigned)grid, kThreads, smem, stream>>>(b);
	return cudaGetLastError();
}
