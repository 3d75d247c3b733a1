// Issue-rate probe for the sm_100a packed fp32 instructions (FADD2 / FMUL2 / FFMA2) against their scalar forms.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_throughput f32x2_throughput.cu ; run on a B200.
// Prints warp-instructions per cycle per SM for 8 independent dependency chains per thread, 32 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(float2 *out, int iters, long long *cycles)
{
	float2 a[8];
	for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
	const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			if (MODE == 0) {  // scalar FFMA x2 (same flops as one FFMA2)
				a[i].x = fmaf(a[i].x, b.x, c.x);
				a[i].y = fmaf(a[i].y, b.y, c.y);
			} else if (MODE == 1) {
				a[i] = __ffma2_rn(a[i], b, c);
			} else if (MODE == 2) {
				a[i].x = a[i].x + c.x;
				a[i].y = a[i].y + c.y;
			} else if (MODE == 3) {
				a[i] = __fadd2_rn(a[i], c);
			} else if (MODE == 4) {  // complex multiply, scalar
				float2 v = a[i];
				a[i] = make_float2(v.x * b.x - v.y * b.y, v.x * b.y + v.y * b.x);
			} else {  // complex multiply, packed: FMUL2 + FFMA2
				float2 v = a[i];
				float2 t = __fmul2_rn(make_float2(v.y, v.y), make_float2(-b.y, b.x));
				a[i] = __ffma2_rn(make_float2(v.x, v.x), b, t);
			}
		}
	}
	long long t1 = clock64();
	float2 s = make_float2(0, 0);
	for (int i = 0; i < 8; i++) s.x += a[i].x, s.y += a[i].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_iter_instr)
{
	float2 *out;
	long long *cyc;
	cudaMalloc(&out, 148 * 1024 * sizeof(float2));
	cudaMalloc(&cyc, 148 * sizeof(long long));
	const int iters = 4096;
	probe<MODE><<<148, 1024>>>(out, iters, cyc);
	probe<MODE><<<148, 1024>>>(out, iters, cyc);
	cudaDeviceSynchronize();
	long long h[148];
	cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < 148; i++) avg += (double)h[i];
	avg /= 148;
	double winst = 32.0 * iters * 8 * per_iter_instr;  // warp instructions per SM
	printf("%-28s %8.0f cycles  %.3f warp-instr/cycle/SM  (%.1f lane-flop/cycle/SM)\n", name, avg, winst / avg,
	       winst / avg * 32 * (MODE <= 1 ? (MODE == 1 ? 4 : 2) : (MODE <= 3 ? (MODE == 3 ? 2 : 1) : (MODE == 5 ? 4 : 2))));
	cudaFree(out);
	cudaFree(cyc);
}

int main()
{
	run<0>("scalar FFMA x2", 2);
	run<1>("FFMA2", 1);
	run<2>("scalar FADD x2", 2);
	run<3>("FADD2", 1);
	run<4>("complex mul scalar (4 instr)", 4);
	run<5>("complex mul packed (2 instr)", 2);
	return 0;
}
