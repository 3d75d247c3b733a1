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
//   * One CTA per frame. The whole decoder state of a frame lives in shared memory for all iterations:
//     posterior[1600] + channel LLR[1600] + one float per Tanner-graph edge (26-39 KB per frame, so 5 frames
//     are resident per SM); HBM is touched once for the 6.4 KB of LLRs in and <=175 bytes + 32 bytes out.
//   * Only posterior and check->variable messages are stored: the variable->check message of the reference
//     (its Q array) is recomputed as posterior - R, which is exactly the reference's Q update.
//   * The Tanner graph is laid out on the host as a warp-blocked ELL on BOTH sides: checks (variables) sorted by
//     descending degree, cut into groups of 32, each group padded to its largest degree.  A warp owns a group, so
//     thread-per-check (thread-per-variable) loops walk the per-edge array with a constant +32-word stride, 32
//     consecutive words per step (conflict free, <12 % padding, uniform trip count inside a warp).  Index tables
//     are read through the read-only path and stay in L1 (shared by all CTAs of the SM).
//   * The syndrome of iteration i is evaluated inside the check pass of iteration i+1 (it gathers the same
//     posteriors anyway) and combined with a single __syncthreads_or: one barrier per half-iteration.
//   * SPA mode evaluates the check node in the log-magnitude domain, s = -log2 tanh(|q|/2), R = phi(sum of the
//     other s) (phi is its own inverse).  This keeps fp32 accurate where the product of tanh saturates, and
//     reproduces the reference's double-precision clamp rule: a factor whose tanh rounds to 1.0 in double
//     (s < 2^-54 nats) contributes 0, and an all-saturated product gives 2 atanh(0.9999999).
//     The leave-one-out sum never cancels: the largest term is kept apart (rest = sum of all others), so the edge
//     that owns it reads `rest` and every other edge reads (rest - self) + largest.  No fp64, no conversions: the
//     kernel is bound by instruction issue and by the XU pipe (3 MUFU per phi), not by memory.
//   * MINSUM mode (north_star): normalised min-sum (alpha 1 for degree<=2 where min-sum is exact, 0.85 for 3,
//     0.75 above), same schedule / exit / clamp.
#include "mb_kernels.cuh"

namespace {

constexpr int kThreads = 32 * MB_LDPC_WARPS;
constexpr float kLn2 = 0.69314718055994531f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kClampR = 16.811242831518264f;               // 2*atanh(0.9999999), ldpc_decoder_SPA.cc:147-155
constexpr float kTanhOne2 = 5.5511151231257827e-17f * kLog2e; // s (base-2 units) below which tanh(|q|/2) == 1.0 in double
constexpr float kSatQ = 38.123095f;                          // |q| above which tanh(|q|/2) == 1.0 in double (the same boundary as kTanhOne2)
constexpr float kTiny2 = 0.015625f * kLog2e;                 // below this S, 1 - 2^-S would cancel: use log2(2/(S ln2))

// Forward map, x = |q| (nats) -> s = -log2 tanh(x/2) = log2((1+e)/(1-e)), e = exp(-x).
//   e < 0.1 : odd series (2/ln2) e (1 + e^2/3 + e^4/5), rel. error < 2e-7  (the logarithm's argument would round to 1)
//   else    : log2((1+e)/(1-e)); x == 0 gives +inf ("this edge carries no information"), which the check node handles as such.
__device__ __forceinline__ float phi_fwd(float x)
{
	const float e = exp2f(-x * kLog2e);  // ex2.approx.ftz under -ftz=true
	const float e2 = e * e;
	const float t = fmaf(e2, 0.2f, 1.0f / 3.0f);
	const float e_s = e * (2.0f * kLog2e);
	const float series = fmaf(e_s, e2 * t, e_s);
	const float lg = __log2f(__fdividef(1.0f + e, 1.0f - e));
	return e < 0.1f ? series : lg;
}

// Backward map, S (base-2 units) -> R = phi(S) in nats = ln((1+e)/(1-e)), e = 2^-S.
//   S < kTiny2 : ln(2 / (S ln2))   (1 - e would cancel; error < 2e-5 absolute on a value > 4.8)
//   else       : ln2 * log2((1+e)/(1-e)); for large S the result is tiny and only its absolute error (1e-7) matters.
__device__ __forceinline__ float phi_bwd(float S)
{
	const float e = exp2f(-S);
	const bool tiny = S < kTiny2;
	const float num = tiny ? 2.0f * kLog2e : 1.0f + e;
	const float den = tiny ? S : 1.0f - e;
	return kLn2 * __log2f(__fdividef(num, den));
}

// x > 0 ? a : b as an opaque select: phi_bwd is evaluated unconditionally (its inf/NaN at x <= 0 is discarded), which is cheaper
// than the divergent branch the compiler would otherwise wrap around the three MUFU operations
__device__ __forceinline__ float sel_gt0(float x, float a, float b)
{
	float r;
	asm("{\n\t.reg .pred p;\n\tsetp.gt.ftz.f32 p, %1, 0f00000000;\n\tselp.f32 %0, %2, %3, p;\n\t}" : "=f"(r) : "f"(x), "f"(a), "f"(b));
	return r;
}

// parity of the hard decisions accumulated as XOR of raw sign bits -> 0/1
__device__ __forceinline__ unsigned lam_sign_fix(unsigned x) { return x >> 31; }

__device__ __forceinline__ uint16_t crc_step_byte(uint16_t crc, unsigned byte)
{
	crc ^= (uint16_t)byte;
#pragma unroll
	for (int i = 0; i < 8; i++) crc = (crc & 1) ? (uint16_t)((crc >> 1) ^ 0xA001) : (uint16_t)(crc >> 1);
	return crc;
}

// shared-memory layout (bytes from the start of the dynamic segment); constant offsets keep every hot access [reg + imm]
constexpr int kOffLam = 0;                        // float[1601] posterior; [1600] = +inf, the variable every padding slot points at
constexpr int kOffLch = (MB_N + 4) * 4;           // float[1600] channel LLR
constexpr int kOffR = kOffLch + MB_N * 4;         // float[c_slots + 1] check -> variable message per check-side slot; [c_slots] == 0 always

// 32-bit shared-window addressing for the gathers of the hot loops (one add per access, base kept in a register)
__device__ __forceinline__ float lds_f(unsigned sbase, unsigned off)
{
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbase + off));
	return v;
}

// Sum-product check node of one lane (one check of a warp group padded to degree d): reads the d posteriors and messages, writes the d new
// messages, returns the XOR of the hard decisions (sign bit in bit 0).
// Leave-one-out sums without cancellation: the largest term is kept apart (big) and rest = sum of all the others, by a running (min, max)
// pair -- no index tracking: the edge that owns the largest term is recognised by value in the second loop (ties are harmless: each tied
// edge's leave-one-out sum is the same `rest`).  s = +inf (q == 0) needs no clamp: it parks in `big`, every other edge then sees
// inf -> message 0, and two of them make rest = inf -> all messages 0.
// D > 0: the group degree is a compile-time constant (the common degrees are instantiated): both loops unroll completely, every
// shared-memory access is [register + immediate] and there is no loop counter -- about a sixth of the generic loop's instructions
// were counters, compares, branches and address arithmetic (ncu source page, profiles/r1n).  Same operations in the same order.
// Which group degrees get their own body was measured on the GPU (B200, ms per 65,536 frames; generic loop only: mode 8 5.06):
//   mode 8 (rate 6/16): {3..8} 4.67, {4..8} 4.47, {5..8} 4.51, {3..6} 4.39, {5..7} 4.41, {4..6} 4.48, {4..7} 4.32, {3..7} 4.24, {3..10} 4.98
//   mode 9 (rate 8/16): {3..7} 4.31, {5..7} 4.29, {6..7} 4.38, {7} 4.54, {6..9} 4.31, {5..9} 4.22
// More bodies are not better: a body is ~45 instructions per edge, the eight warps of a CTA sit in different bodies at once (instruction
// cache), and degrees that are multiples of 4 lose little in the generic loop anyway (it is unrolled by 4: 4 / 2 / 1 next to {4..7} gave
// 4.32 / 4.32 / 4.39).  Chunked unrolling of the larger fixed degrees (8 -> 2 x 4) did not change that in mode 8 ({3..8} 4.37-4.40), nor did keeping
// the parked magnitudes in registers instead of the message slots (degree <= 4 / 6 / 8: 4.69 / 4.74 / 4.73 against 4.66).
// So the kernel is instantiated per degree SET and the launch picks the set by rate (check degrees: SURVEY.md 8a graph table): rates
// 1..4/16 {3..5}, rates 5,6/16 {3..7}, rate 8/16 {5..9}; rate 14/16 (degrees 23-46) runs the generic loop whichever set is loaded.
constexpr int kGenUnroll = 4;
// unroll factor of a body: complete up to 7 edges, beyond that in equal chunks (8 -> 2 x 4, 9 -> 3 x 3): the trip count is still a constant
// (no remainder loop) and the body stays small.  Rate 8/16 (mode 9, 65,536 frames): {5..9} chunked 4.22 ms, {5..9} complete 4.62, {6..9} 4.31.
__host__ __device__ constexpr int fix_unroll(int D) { return D <= 0 ? kGenUnroll : (D <= 7 ? D : (D % 2 == 0 ? D / 2 : (D % 3 == 0 ? D / 3 : D))); }
template <int D>
__device__ __forceinline__ unsigned spa_check_node(unsigned sbase, const uint16_t *__restrict__ ve, float *__restrict__ Re, int d_rt)
{
	const int d = D > 0 ? D : d_rt;
	unsigned hard = 0, par = 0;
	float big = 0.f, rest = 0.f;
#pragma unroll(fix_unroll(D))
	for (int k = 0; k < d; k++) {
		const float lam = lds_f(sbase, kOffLam + ve[k * 32]);
		const float q = lam - Re[k * 32];
		hard ^= __float_as_uint(lam);  // sign bit only is used
		par ^= __float_as_uint(q);
		float s = phi_fwd(fabsf(q));
		s = s < kTanhOne2 ? 0.f : s;
		rest += fminf(s, big);
		big = fmaxf(s, big);
		Re[k * 32] = __uint_as_float(__float_as_uint(s) | (__float_as_uint(q) & 0x80000000u));  // signed magnitude parked in the slot
	}
	const unsigned pneg = par & 0x80000000u;
#pragma unroll(fix_unroll(D))
	for (int k = 0; k < d; k++) {
		const unsigned tb = __float_as_uint(Re[k * 32]);
		const float sk = __uint_as_float(tb & 0x7fffffffu);
		const float so = sk == big ? rest : (rest - sk) + big;
		const float mag = sel_gt0(so, phi_bwd(so), kClampR);  // all-saturated product -> 2 atanh(0.9999999)
		Re[k * 32] = __uint_as_float(__float_as_uint(mag) | ((tb ^ pneg) & 0x80000000u));
	}
	return lam_sign_fix(hard);
}

// Variable node of the head (degree > 2, groups padded to an even degree): channel LLR + the incoming messages in the table's order.
// D > 0: fixed group degree, every index load and gather at an immediate offset.
template <int D>
__device__ __forceinline__ float var_node_sum(unsigned sbase, const uint16_t *__restrict__ se, float acc, int d_rt)
{
	if (D > 0) {
		unsigned idx[D > 0 ? D : 1];
#pragma unroll
		for (int k = 0; k < D; k++) idx[k] = se[k * 32];
#pragma unroll
		for (int k = 0; k < D; k++) acc += lds_f(sbase, kOffR + idx[k]);
		return acc;
	}
	int k = 0;
#pragma unroll 1
	for (; k + 4 <= d_rt; k += 4) {
		const unsigned i0 = se[0], i1 = se[32], i2 = se[64], i3 = se[96];
		se += 128;
		acc += lds_f(sbase, kOffR + i0);
		acc += lds_f(sbase, kOffR + i1);
		acc += lds_f(sbase, kOffR + i2);
		acc += lds_f(sbase, kOffR + i3);
	}
	if (k < d_rt) {
		const unsigned i0 = se[0], i1 = se[32];
		acc += lds_f(sbase, kOffR + i0);
		acc += lds_f(sbase, kOffR + i1);
	}
	return acc;
}

// XOR of the hard decisions of one check (syndrome-only test)
template <int D>
__device__ __forceinline__ unsigned check_parity(unsigned sbase, const uint16_t *__restrict__ ve, int d_rt)
{
	const int d = D > 0 ? D : d_rt;
	unsigned hard = 0;
#pragma unroll(D > 0 ? D : 4)
	for (int k = 0; k < d; k++) hard ^= __float_as_uint(lds_f(sbase, kOffLam + ve[k * 32]));
	return hard >> 31;
}

// Normalised min-sum check node of one lane, same contract as spa_check_node (alpha 1 for true degree <= 2 where min-sum is exact, 0.85 for
// 3, 0.75 above); D > 0: fixed group degree, loops unrolled, the sign bits and the arg-min compare against constants.
template <int D>
__device__ __forceinline__ unsigned minsum_check_node(unsigned sbase, const uint16_t *__restrict__ ve, float *__restrict__ Re, int d_rt, int dc)
{
	const int d = D > 0 ? D : d_rt;
	unsigned hard = 0, par = 0;
	float m1 = 3.0e38f, m2 = 3.0e38f;
	int arg = -1;
	unsigned long long signs = 0ull;
#pragma unroll(fix_unroll(D))
	for (int k = 0; k < d; k++) {
		const float lam = lds_f(sbase, kOffLam + ve[k * 32]);
		const float q = lam - Re[k * 32];
		hard ^= __float_as_uint(lam);
		par ^= __float_as_uint(q);
		signs |= (unsigned long long)(__float_as_uint(q) >> 31) << k;
		const float aq = fabsf(q);
		const bool lt1 = aq < m1;
		m2 = lt1 ? m1 : fminf(m2, aq);
		arg = lt1 ? k : arg;
		m1 = lt1 ? aq : m1;
	}
	const float alpha = dc <= 2 ? 1.0f : (dc == 3 ? 0.85f : 0.75f);
	const unsigned pneg = par >> 31;
#pragma unroll(fix_unroll(D))
	for (int k = 0; k < d; k++) {
		const float mag = fminf(alpha * (k == arg ? m2 : m1), kClampR);
		const unsigned neg = pneg ^ (unsigned)((signs >> k) & 1ull);
		Re[k * 32] = neg ? -mag : mag;
	}
	return lam_sign_fix(hard);
}

template <int ALGO, int FMIN, int FMAX>
__global__ void __launch_bounds__(kThreads, 5) mb_ldpc_kernel(const MbLdpcArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr bool kSmallBodies = FMAX <= 7;  // per-degree bodies of the variable nodes and of the syndrome test as well
	const MbMode &m = a.mode;
	const MbRate &rt = a.rate;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int CS = rt.c_slots;
	float *s_lam = reinterpret_cast<float *>(smem_raw + kOffLam);
	float *s_lch = reinterpret_cast<float *>(smem_raw + kOffLch);
	float *s_R = reinterpret_cast<float *>(smem_raw + kOffR);
	uint32_t *s_csched = reinterpret_cast<uint32_t *>(smem_raw + kOffR + ((CS + 4) & ~3) * 4);  // [8][16] check-group descriptors per warp
	uint32_t *s_vsched = s_csched + MB_LDPC_WARPS * MB_SCHED_LEN;                              // [8][16] variable-group descriptors per warp
	unsigned char *s_bytes = reinterpret_cast<unsigned char *>(s_vsched + MB_LDPC_WARPS * MB_SCHED_LEN);

	const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);

	const size_t frame = blockIdx.x;
	const uint16_t *__restrict__ g_edge_var = reinterpret_cast<const uint16_t *>(a.blob + rt.off_edge_varb);  // byte offsets into s_lam
	const uint16_t *__restrict__ g_vedge = reinterpret_cast<const uint16_t *>(a.blob + rt.off_vedgeb);        // byte offsets into s_R
	const uint32_t *__restrict__ g_vtail = reinterpret_cast<const uint32_t *>(a.blob + rt.off_vtail);          // two byte offsets per degree-<=2 variable

	MbRxStats st = a.stats[frame];
	if (a.check_gate && !(st.mean_H >= 0.3f)) {
		// telecom_system.cc:1268-1280: a channel estimate this weak means a false sync; the reference skips the decode
		for (int i = tid; i < m.frame_bytes; i += kThreads) a.payload[frame * (size_t)m.frame_bytes + i] = 0;
		if (tid == 0) {
			st.iterations_done = -1;
			st.crc = 0;
			st.all_zeros = 0;
			st.message_decoded = 0;
			st.SNR = -99.9f;
			a.stats[frame] = st;
		}
		return;
	}

	{
		// hand-off layout: every 32-float row arrives rotated by its row index (MB_HANDOFF); coalesced loads, conflict-free stores
		const float *__restrict__ src = a.llr + frame * (size_t)MB_HANDOFF_STRIDE;
		for (int i = tid; i < MB_N; i += kThreads) {
			const float v = __ldcs(src + i);
			const unsigned p = MB_HANDOFF_INV((unsigned)i);
			s_lam[p] = v;
			s_lch[p] = v;
		}
		for (int i = tid; i <= CS; i += kThreads) s_R[i] = 0.f;
		if (tid < MB_LDPC_WARPS * MB_SCHED_LEN) {
			s_csched[tid] = reinterpret_cast<const uint32_t *>(a.blob + rt.off_csched)[tid];
			s_vsched[tid] = reinterpret_cast<const uint32_t *>(a.blob + rt.off_vsched)[tid];
		}
		if (tid == 0) s_lam[MB_N] = __int_as_float(0x7f800000);  // +inf: a padding edge contributes s = 0, sign +, hard bit 0
	}
	__syncthreads();

	const int vtail0 = rt.vtail_start;
	int iterations = 0;
	for (int pass = 0;; pass++) {
		// ---- check pass: syndrome of the current posterior + new check->variable messages ----------------
		// A warp owns a group of 32 checks padded to one degree; padding slots point at the +inf variable, so the loops are
		// warp-uniform (no per-thread degree, no divergence) and a padding edge is the neutral element of every reduction.
		unsigned unsat = 0;
		const uint32_t *sched = s_csched + warp * MB_SCHED_LEN;
		for (uint32_t desc = *sched; desc != 0u; desc = *++sched) {  // this warp's check groups (static, degree-balanced schedule)
			const int d = (int)((desc >> 16) & 0xFFu);
			const int e0 = (int)(desc & 0xFFFFu) + lane;
			float *__restrict__ Re = s_R + e0;
			const uint16_t *__restrict__ ve = g_edge_var + e0;
			unsigned hard = 0, par = 0;
			if (ALGO == 0 && d == 2) {
				// Degree-2 check (two thirds of the checks of the low-rate codes): R_a = 2 atanh(tanh(q_b / 2)) is q_b itself until tanh
				// rounds to 1.0 in double (|q| > 37.43), where the reference's clamp gives 2 atanh(0.9999999) (ldpc_decoder_SPA.cc:147-155).
				// A padding edge (q = +inf) saturates to the clamp: exactly the reference's empty product of a degree-1 check.
				const float l0 = lds_f(sbase, kOffLam + ve[0]), l1 = lds_f(sbase, kOffLam + ve[32]);
				const float q0 = l0 - Re[0], q1 = l1 - Re[32];
				hard = lam_sign_fix(__float_as_uint(l0) ^ __float_as_uint(l1));
				Re[0] = fabsf(q1) > kSatQ ? copysignf(kClampR, q1) : q1;
				Re[32] = fabsf(q0) > kSatQ ? copysignf(kClampR, q0) : q0;
			} else if (ALGO == 0) {
				switch (d) {
#define MB_FIX_CASE(D_) \
	case D_:            /* a degree outside the instantiated set falls through to the generic loop */ \
		if (D_ >= FMIN && D_ <= FMAX && d == D_) { hard = spa_check_node<(D_ >= FMIN && D_ <= FMAX) ? D_ : 0>(sbase, ve, Re, d); break; }
				MB_FIX_CASE(3)
				MB_FIX_CASE(4)
				MB_FIX_CASE(5)
				MB_FIX_CASE(6)
				MB_FIX_CASE(7)
				MB_FIX_CASE(8)
				MB_FIX_CASE(9)
				MB_FIX_CASE(10)
#undef MB_FIX_CASE
				default: hard = spa_check_node<0>(sbase, ve, Re, d); break;
				}
			} else {
				const int c = (int)((desc >> 24) - 1u) * 32 + lane;
				const int dc = c < rt.P ? (int)(a.blob + rt.off_cdeg)[c] : 0;  // the true degree picks the normalisation
				switch (d) {
#define MB_FIX_CASE(D_) \
	case D_:            /* a degree outside the instantiated set falls through to the generic loop */ \
		if (D_ >= FMIN && D_ <= FMAX && d == D_) { hard = minsum_check_node<(D_ >= FMIN && D_ <= FMAX) ? D_ : 0>(sbase, ve, Re, d, dc); break; }
				MB_FIX_CASE(3)
				MB_FIX_CASE(4)
				MB_FIX_CASE(5)
				MB_FIX_CASE(6)
				MB_FIX_CASE(7)
				MB_FIX_CASE(8)
				MB_FIX_CASE(9)
				MB_FIX_CASE(10)
#undef MB_FIX_CASE
				default: hard = minsum_check_node<0>(sbase, ve, Re, d, dc); break;
				}
			}
			unsat |= hard;
		}
		// number of threads that saw an unsatisfied check: 0 = converged; a small count = "probably one iteration to go"
		const int n_unsat = __syncthreads_count((int)unsat);
		if (n_unsat == 0) {
			iterations = pass;  // converged after `pass` iterations (0 = clean on arrival, ldpc_decoder_SPA.cc:62-77)
			break;
		}
		if (pass == a.max_iters) {
			iterations = a.max_iters + 1;  // ldpc_decoder_SPA.cc:127,217: loop ran out
			break;
		}
		// ---- variable pass: posterior = channel + sum of incoming messages (reference V-row order) ----------
		// Variables are numbered by descending degree.  The head (degree > 2) is walked in warp groups padded to one even
		// degree (padding reads the always-zero slot); the long tail of degree-<=2 variables (the accumulator chain of the IRA
		// code, ~60 % of all variables) is a flat loop with both message offsets packed in one word.
		sched = s_vsched + warp * MB_SCHED_LEN;
		for (uint32_t desc = *sched; desc != 0u; desc = *++sched) {
			const int d = (int)((desc >> 16) & 0xFFu);
			const int v = (int)((desc >> 24) - 1u) * 32 + lane;
			const uint16_t *__restrict__ se = g_vedge + ((desc & 0xFFFFu) + lane);
			float acc = s_lch[v];
			// (the {5..9} kernel of rate 8/16 is at its instruction-cache budget with the check bodies alone: mode 9 4.22 ms without these, 4.41 with)
			switch (kSmallBodies ? d : 0) {  // variable degrees are 3..9 in all eight codes (padded: 4, 6, 8, 10)
			case 4: acc = var_node_sum<4>(sbase, se, acc, d); break;
			case 6: acc = var_node_sum<6>(sbase, se, acc, d); break;
			case 8: acc = var_node_sum<8>(sbase, se, acc, d); break;
			case 10: acc = var_node_sum<10>(sbase, se, acc, d); break;
			default: acc = var_node_sum<0>(sbase, se, acc, d); break;
			}
			s_lam[v] = acc;
		}
		for (int v = vtail0 + tid; v < MB_N; v += kThreads) {
			const uint32_t w = __ldg(g_vtail + (v - vtail0));
			float acc = s_lch[v];
			acc += lds_f(sbase, kOffR + (w & 0xFFFFu));
			acc += lds_f(sbase, kOffR + (w >> 16));
			s_lam[v] = acc;
		}
		__syncthreads();
		// ---- cheap syndrome-only test when convergence is likely: saves the (expensive) message update of a final pass ----
		if (n_unsat <= a.cheap_test_threads) {
			unsigned bad = 0;
			sched = s_csched + warp * MB_SCHED_LEN;
			for (uint32_t desc = *sched; desc != 0u; desc = *++sched) {
				const int d = (int)((desc >> 16) & 0xFFu);
				const uint16_t *__restrict__ ve = g_edge_var + (desc & 0xFFFFu) + lane;
				switch (kSmallBodies ? d : 0) {
				case 2: bad |= check_parity<2>(sbase, ve, d); break;
				case 3: bad |= check_parity<3>(sbase, ve, d); break;
				case 4: bad |= check_parity<4>(sbase, ve, d); break;
				case 5: bad |= check_parity<5>(sbase, ve, d); break;
				case 6: bad |= check_parity<6>(sbase, ve, d); break;
				case 7: bad |= check_parity<7>(sbase, ve, d); break;
				case 8: bad |= check_parity<8>(sbase, ve, d); break;
				case 9: bad |= check_parity<9>(sbase, ve, d); break;
				default: bad |= check_parity<0>(sbase, ve, d); break;
				}
			}
			if (__syncthreads_or((int)bad) == 0) {
				iterations = pass + 1;  // exactly what the next check pass would have reported
				break;
			}
		}
	}

	// ---- hard decision -> de-scramble -> pack LSB first -> all-zeros / CRC16 -> record --------------------------
	const uint16_t *__restrict__ g_bit_var = reinterpret_cast<const uint16_t *>(a.blob + m.off_bit_var);
	const uint8_t *__restrict__ g_scr = a.blob + m.off_scr;
	unsigned byte = 0;
	if (tid < m.crc_bytes) {
#pragma unroll
		for (int b = 0; b < 8; b++) {
			const int i = tid * 8 + b;
			const unsigned bit = (__float_as_uint(s_lam[g_bit_var[i]]) >> 31) ^ (unsigned)g_scr[i];  // hard decision = sign bit (LLR < 0)
			byte |= bit << b;
		}
		s_bytes[tid] = (unsigned char)byte;
		if (tid < m.frame_bytes) a.payload[frame * (size_t)m.frame_bytes + tid] = (uint8_t)byte;
	}
	const int nonzero = __syncthreads_or((int)byte);
	if (tid < 32) {
		const uint16_t *__restrict__ g_mat = reinterpret_cast<const uint16_t *>(a.blob + m.off_crcmat);
		uint16_t part = 0;
		const int b0 = tid * m.crc_chunk;
		for (int i = 0; i < m.crc_chunk; i++)
			if (b0 + i < m.crc_bytes) part = crc_step_byte(part, s_bytes[b0 + i]);
		unsigned adv = 0;
#pragma unroll
		for (int b = 0; b < 16; b++)
			if ((part >> b) & 1) adv ^= g_mat[tid * 16 + b];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) adv ^= __shfl_xor_sync(0xffffffffu, adv, o);
		if (tid == 0) {
			const int all_zeros = nonzero ? 0 : 1;
			const int crc = all_zeros ? 0 : (int)((adv ^ m.crc_init) & 0xFFFFu);  // telecom_system.cc:1337-1341
			const int decoded = (!all_zeros && crc == 0) ? 1 : 0;               // telecom_system.cc:1343-1349
			st.iterations_done = iterations;
			st.crc = crc;
			st.all_zeros = all_zeros;
			st.message_decoded = decoded;
			st.SNR = decoded ? st.SNR : -99.9f;
			if (m.estimator == 1 || !decoded) a.stats[frame] = st;
			s_bytes[255] = (unsigned char)decoded;
		}
	}
	if (m.estimator == 1) return;  // LS modes: the demodulator's pilot variance is the SNR report (:1368-1375)

	// ---- ZF modes: SNR report of a decoded frame (telecom_system.cc:1376-1400) -----------------------------------------
	// Re-encode the hard decisions (scrambled info bits, virtual copies, IRA parity), re-map them onto the constellation through
	// the same composed interleaver records the demodulator scatters by, and measure the mean squared distance of the equalised
	// data symbols (kept by the demodulator behind the LLRs of the hand-off record) to the re-encoded ones (ofdm.cc:1622-1635).
	__syncthreads();
	if (!s_bytes[255]) return;
	uint32_t *s_bit = reinterpret_cast<uint32_t *>(smem_raw + kOffLch);          // [1600] re-encoded bit per internal variable (the channel LLRs are dead)
	unsigned char *s_d = reinterpret_cast<unsigned char *>(smem_raw + kOffR);    // [P] data parity of every check, reference check order (the messages are dead)
	const uint16_t *__restrict__ g_voc = reinterpret_cast<const uint16_t *>(a.blob + rt.off_var_of_cw);
	const int K = m.K, P = rt.P, nReal = m.nReal, nVirtual = m.nVirtual;
	for (int v = tid; v < MB_N; v += kThreads) s_bit[v] = 0u;  // parity variables stay 0 while the data parities are formed
	__syncthreads();
	for (int i = tid; i < nReal; i += kThreads) {  // hd_decoded_data_bit, scrambled back = the decoder's hard decisions (:1378)
		const unsigned b = __float_as_uint(s_lam[g_voc[i]]) >> 31;
		s_bit[g_voc[i]] = b;
		if (i < nVirtual) s_bit[g_voc[nReal + i]] = b;  // virtual bits are copies of the first ones (:1380-1383)
	}
	__syncthreads();
	{  // cl_ldpc::encode (ldpc.cc:111-132): every check row is {data bits, parity i-1, parity i}, so parity = running XOR of the data parities
		const uint8_t *__restrict__ g_cdeg = a.blob + rt.off_cdeg;
		const uint32_t *__restrict__ g_cgbase = reinterpret_cast<const uint32_t *>(a.blob + rt.off_cgbase);
		const uint16_t *__restrict__ g_cos = reinterpret_cast<const uint16_t *>(a.blob + rt.off_check_of_sorted);
		for (int c = tid; c < P; c += kThreads) {
			const uint16_t *__restrict__ ve = g_edge_var + g_cgbase[c >> 5] + (c & 31);
			unsigned x = 0;
			for (int k = 0; k < (int)g_cdeg[c]; k++) x ^= s_bit[ve[k * 32] >> 2];
			s_d[g_cos[c]] = (unsigned char)x;
		}
	}
	__syncthreads();
	if (warp == 0) {
		const int chunk = (P + 31) >> 5;
		unsigned x = 0;
		for (int i = lane * chunk; i < min(P, (lane + 1) * chunk); i++) x ^= s_d[i];
		unsigned carry = x;  // inclusive XOR scan of the chunk parities over the lanes
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned y = __shfl_up_sync(0xffffffffu, carry, o);
			if (lane >= o) carry ^= y;
		}
		unsigned run = carry ^ x;  // parity of everything before this lane's chunk
		for (int i = lane * chunk; i < min(P, (lane + 1) * chunk); i++) {
			run ^= s_d[i];
			s_bit[g_voc[K + i]] = run;
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
				loc = (loc << 1) | s_bit[MB_HANDOFF_INV(off >> 2)];
			}
			const float2 c = g_cons[loc], z = zf[d];
			acc += (c.x - z.x) * (c.x - z.x) + (c.y - z.y) * (c.y - z.y);
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
		float *s_acc = reinterpret_cast<float *>(smem_raw + kOffLam);  // the posteriors are dead now
		__syncthreads();
		if (lane == 0) s_acc[warp] = acc;
		__syncthreads();
		if (tid == 0) {
			float tot = 0.f;
			for (int w = 0; w < kThreads / 32; w++) tot += s_acc[w];
			st.SNR = -10.0f * log10f(tot / (float)m.nData);  // measure_SNR, ofdm.cc:1622-1635
			a.stats[frame] = st;
		}
	}
}

}  // namespace

size_t mb_ldpc_smem_bytes(int c_slots)
{
	return (size_t)kOffR + (size_t)((c_slots + 4) & ~3) * sizeof(float) + 2 * MB_LDPC_WARPS * MB_SCHED_LEN * sizeof(uint32_t) + 256;
}

namespace {
typedef void (*LdpcKernel)(const MbLdpcArgs);
// [algo][set]: sum-product / min-sum, each with the degree sets {3..5}, {3..7}, {5..9}
const LdpcKernel kKernels[6] = {mb_ldpc_kernel<0, 3, 5>, mb_ldpc_kernel<0, 3, 7>, mb_ldpc_kernel<0, 5, 9>,
				mb_ldpc_kernel<1, 3, 5>, mb_ldpc_kernel<1, 3, 7>, mb_ldpc_kernel<1, 5, 9>};
}  // namespace

cudaError_t mb_ldpc_init()
{
	for (LdpcKernel k : kKernels) {
		cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
		if (e != cudaSuccess) return e;
	}
	return cudaSuccess;
}

cudaError_t mb_launch_ldpc(const MbLdpcArgs &a, size_t n_frames, int algo, cudaStream_t stream)
{
	if (n_frames == 0) return cudaSuccess;
	const size_t smem = mb_ldpc_smem_bytes(a.rate.c_slots);
	const LdpcKernel k = kKernels[(algo != 0 ? 3 : 0) + mb_ldpc_degree_set(a.rate.rate_num, nullptr, nullptr)];
	k<<<(unsigned)n_frames, kThreads, smem, stream>>>(a);
	return cudaGetLastError();
}
