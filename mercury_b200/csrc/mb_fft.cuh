// mb_fft.cuh -- packed complex arithmetic and the 16-point DFT building blocks of the 256-point FFT (16 x 16, radix 4 x 4 in
// registers), shared by the OFDM demodulator (mb_demod.cu) and the MFSK demodulator (mb_mfsk.cu).  Device-only, include inside a
// translation unit compiled WITHOUT -ftz (see the Makefile: under FTZ ptxas no longer folds negations into the packed operands).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mbfft {

// ------------------------------------------------------------------------------------------------------------------
// packed complex arithmetic (float2 = re, im)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }  // a + (-i) b
__device__ __forceinline__ float2 cadd_pi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }  // a + (+i) b
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// Complex products as FMUL2 + FFMA2.  The sign pattern sits on the broadcast scalar (not on the swapped pair): that is the
// form ptxas folds into the FFMA2 operand modifiers (.LO_HI swap, .NP half negation) instead of materialising a negation.
__device__ __forceinline__ float2 cmul(float2 a, float2 w)  // (a.x w.x - a.y w.y, a.x w.y + a.y w.x)
{
	return __ffma2_rn(make_float2(w.y, w.x), make_float2(-a.y, a.y), __fmul2_rn(make_float2(a.x, a.x), w));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 h)  // a * conj(h) = h.x (a.x, a.y) + h.y (a.y, -a.x)
{
	return __ffma2_rn(make_float2(a.y, a.x), make_float2(h.y, -h.y), __fmul2_rn(make_float2(h.x, h.x), a));
}
// Streaming 8-byte load of a sample: no L1 allocation, so the descriptor tables and twiddles stay L1 resident.
__device__ __forceinline__ float2 ld_stream(const float2 *p)
{
	float2 r;
	asm("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
	return r;
}
__device__ __forceinline__ uint32_t ld_stream_b32(const uint32_t *p)
{
	uint32_t r;
	asm("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}
__device__ __forceinline__ float cnorm2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }
__device__ __forceinline__ float fast_rcp(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ float fast_rsqrt(float x)
{
	float r;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ float fast_sqrt(float x)
{
	float r;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// forward 4-point DFT (W4 = -i)
__device__ __forceinline__ void dft4(float2 a, float2 b, float2 c, float2 d, float2 &y0, float2 &y1, float2 &y2, float2 &y3)
{
	const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
	y0 = cadd(t0, t2);
	y2 = csub(t0, t2);
	y1 = cadd_mi(t1, t3);
	y3 = cadd_pi(t1, t3);
}

#define MB_C1 0.92387953251128674f
#define MB_S1 0.38268343236508977f
#define MB_R2 0.70710678118654752f

// Stage 1 of the radix-4x4 16-point forward DFT with the W16 twiddles applied, EXCEPT the -i of A[2][2] (W16^4), which
// the second stage folds into its additions:  A[n2][k1] = W16^(n2 k1) * sum_n1 x[4 n1 + n2] (-i)^(n1 k1)
__device__ __forceinline__ void fft16_front(const float2 (&x)[16], float2 (&A)[4][4])
{
#pragma unroll
	for (int n2 = 0; n2 < 4; n2++) dft4(x[n2], x[4 + n2], x[8 + n2], x[12 + n2], A[n2][0], A[n2][1], A[n2][2], A[n2][3]);
	A[1][1] = cmul(A[1][1], make_float2(MB_C1, -MB_S1));    // W16^1
	A[1][2] = cscale(cadd_mi(A[1][2], A[1][2]), MB_R2);     // W16^2 = (1 - i)/sqrt2
	A[1][3] = cmul(A[1][3], make_float2(MB_S1, -MB_C1));    // W16^3
	A[2][1] = cscale(cadd_mi(A[2][1], A[2][1]), MB_R2);     // W16^2
	A[2][3] = cscale(cadd_pi(A[2][3], A[2][3]), -MB_R2);    // W16^6 = -(1 + i)/sqrt2
	A[3][1] = cmul(A[3][1], make_float2(MB_S1, -MB_C1));    // W16^3
	A[3][2] = cscale(cadd_pi(A[3][2], A[3][2]), -MB_R2);    // W16^6
	A[3][3] = cmul(A[3][3], make_float2(-MB_C1, MB_S1));    // W16^9
}

// full 16-point forward DFT, natural order out: X[k1 + 4 k2]
__device__ __forceinline__ void fft16(const float2 (&x)[16], float2 (&X)[16])
{
	float2 A[4][4];
	fft16_front(x, A);
#pragma unroll
	for (int k1 = 0; k1 < 4; k1++) {
		if (k1 != 2) {
			dft4(A[0][k1], A[1][k1], A[2][k1], A[3][k1], X[k1], X[k1 + 4], X[k1 + 8], X[k1 + 12]);
		} else {  // the third input still lacks its -i
			const float2 t0 = cadd_mi(A[0][2], A[2][2]), t1 = cadd_pi(A[0][2], A[2][2]);
			const float2 t2 = cadd(A[1][2], A[3][2]), t3 = csub(A[1][2], A[3][2]);
			X[2] = cadd(t0, t2);
			X[10] = csub(t0, t2);
			X[6] = cadd_mi(t1, t3);
			X[14] = cadd_pi(t1, t3);
		}
	}
}

// 16-point forward DFT pruned to outputs 0, 1, 14, 15 (the only ones that reach the 50 active carriers)
__device__ __forceinline__ void fft16_pruned(const float2 (&x)[16], float2 &X0, float2 &X1, float2 &X14, float2 &X15)
{
	float2 A[4][4];
	fft16_front(x, A);
	X0 = cadd(cadd(A[0][0], A[2][0]), cadd(A[1][0], A[3][0]));          // k1=0, k2=0
	X1 = cadd(cadd(A[0][1], A[2][1]), cadd(A[1][1], A[3][1]));          // k1=1, k2=0
	X14 = cadd_pi(cadd_pi(A[0][2], A[2][2]), csub(A[1][2], A[3][2]));   // k1=2, k2=3: (a0 - (-i a2)) + i (a1 - a3)
	X15 = cadd_pi(csub(A[0][3], A[2][3]), csub(A[1][3], A[3][3]));      // k1=3, k2=3
}

}  // namespace mbfft
