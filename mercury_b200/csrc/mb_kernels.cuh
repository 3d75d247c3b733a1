// mb_kernels.cuh -- launch interfaces of the two sm_100a kernels of the RX path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "mb_tables.h"

#define MB_HANDOFF_STRIDE 2400  // floats per frame of the stage hand-off buffer (= MERCURY_B200_HANDOFF_FLOATS): 1600 LLRs + ZF-mode data symbols

// Mirrors mercury_b200_rx_stats (include/mercury_b200.h); 32 bytes.
struct MbRxStats {
	int32_t iterations_done, crc, all_zeros, message_decoded;
	float SNR, variance, mean_H;
	int32_t reserved;
};

struct MbDemodArgs {
	const float2 *x;       // [B][Nsymb][sym_stride] complex64 baseband, preamble stripped; the 256 useful samples start at sym_skip
	int32_t sym_stride;    // 272 = symbols as delivered (guard interval present, skipped: sym_skip 16); 256 = guard interval already removed
	int32_t sym_skip;
	float *llr;            // [B][1600] LLRs in the hand-off layout (internal variable order, 32-float rows rotated: MB_HANDOFF)
	float *llr_cw;         // optional [B][1600] LLRs in codeword order (parity output)
	MbRxStats *stats;      // [B]
	float2 *dbg_Y, *dbg_H, *dbg_Z;  // optional [B][Nsymb*50] stage captures
	const uint8_t *blob;   // device copy of the table blob
	uint32_t off_twiddle;
	uint32_t off_var_of_cw;
	unsigned long long n_frames;  // filled by mb_launch_demod (the kernel is persistent: grid < frames)
	MbMode mode;
};

struct MbLdpcArgs {
	const float *llr;      // [B][1600] hand-off layout (MB_HANDOFF)
	uint8_t *payload;      // [B][frame_bytes]
	MbRxStats *stats;      // [B]  (SNR/variance/mean_H already filled by the demod kernel)
	const uint8_t *blob;
	MbMode mode;
	MbRate rate;
	int32_t max_iters;
	int32_t check_gate;    // 1: honour the mean|H| < 0.3 gate recorded by the demod kernel
	int32_t cheap_test_threads;  // run the syndrome-only test after an iteration that started with <= this many unhappy threads
};

size_t mb_ldpc_smem_bytes(int c_slots);

cudaError_t mb_launch_demod(const MbDemodArgs &a, size_t n_frames, cudaStream_t stream);
cudaError_t mb_launch_ldpc(const MbLdpcArgs &a, size_t n_frames, int algo, cudaStream_t stream);
cudaError_t mb_demod_init();  // opt-in shared memory attributes, occupancy of every instantiation
int mb_demod_ctas_per_sm(int Nsymb, int M, int estimator, int phase_only);  // resident CTAs per SM (0 = no such instantiation)
cudaError_t mb_ldpc_init();
