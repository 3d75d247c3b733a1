// mb_tables.h -- mode / code tables of the B200 RX path, as one relocatable blob.
//
// Everything the kernels need besides the samples lives in ONE contiguous byte blob built on the host at
// mercury_b200_load_tables() time: the per-mode (CONFIG_0..16) demodulator tables and the per-rate LDPC
// graphs.  The blob is position independent (tables are addressed by byte offsets from its start), so it is
// (a) uploaded to HBM with a single copy, (b) kept resident for ALL 17 modes so that load_configuration()
// is O(1) (ARQ flips between data and ack configurations every batch), and (c) the thing rank 0
// broadcasts over NCCL to the other GPUs (north_star: "NCCL broadcast of the codeword tables").
//
// Reference anchors (paths relative to /root/reference):
//   mode table            source/physical_layer/telecom_system.cc:2506-2654
//   pilot lattice/sequence source/physical_layer/ofdm.cc:904-1064
//   constellations        source/physical_layer/psk.cc:65-256
//   (de)interleavers      source/physical_layer/interleaver.cc:26-117
//   LLR expand            source/physical_layer/telecom_system.cc:1300-1308
//   scrambler             source/physical_layer/telecom_system.cc:1961-1966
//   LDPC tables           source/physical_layer/mercury_normal_*_16.cc via mercury_b200/data/ldpc_tables.bin
#pragma once
#include <cstdint>

#define MB_N 1600
#define MB_NC 50
#define MB_NFFT 256
#define MB_NGI 16
#define MB_NOFDM 272
#define MB_NMODES 17
#define MB_NRATES 8
#define MB_MAX_SYMB 48
#define MB_MAX_CELLS (MB_MAX_SYMB * MB_NC)
#define MB_LS_HALF 10        // LS window 21x21 (20 -> odd 21, telecom_system.cc:2799-2809)
#define MB_BLOB_MAGIC 0x42324d42u /* "BM2B" */
#define MB_BLOB_VERSION 15u
#define MB_NO_DST 0xFFFFu
#define MB_MAX_CDEG 48
#define MB_MAX_VDEG 16
#define MB_MAX_GROUPS 64  // warp-sized groups of checks / variables (<= 50 used)
#ifndef MB_LDPC_WARPS
#define MB_LDPC_WARPS 8   // warps per decoder CTA (the task schedules below are built for this many)
#endif
#ifndef MB_LDPC_MIN_CTAS
#define MB_LDPC_MIN_CTAS 3  // resident decoder CTAs per SM the kernel's register budget is set for
#endif
#define MB_SCHED_LEN 16   // tasks per warp in a schedule, 0 terminated
#define MB_ZF_STRIDE 27   // compact pilot row: 4 zeros | <= 17 pilots (columns s%3 + 3j at [4 + j]) | zeros
#define MB_LS_COLS 18     // distinct clipped 21-column windows per row: start index (c + 4 - s%3) / 3 = 0..17
// LLR hand-off layout between the two kernels: internal variable v sits at MB_HANDOFF(v): every 32-float row is rotated by
// its row index.  The de-interleavers send the LLRs of neighbouring cells to variables a multiple of 160 apart (the same
// shared-memory bank); the rotation spreads them over the banks (10-way -> 2.8-way conflicts on the demodulator's scatter),
// and the decoder undoes it for free while it stages the vector in shared memory.
#define MB_HANDOFF(v) (((v) & ~31u) | (((v) + ((v) >> 5)) & 31u))
#define MB_HANDOFF_INV(i) (((i) & ~31u) | (((i) - ((i) >> 5)) & 31u))

struct MbRate {
	int32_t rate_num, N, K, P, n_edges;
	int32_t max_cdeg, max_vdeg;
	int32_t c_slots, v_slots;  // padded slot counts of the two warp-blocked ELL layouts below
	int32_t reserved;
	// Check side. Checks are sorted by degree (descending) and cut into groups of 32; group g is padded to the degree of its
	// first (= largest) check and laid out as S warp tasks (mb_ldpc_split / mb_ldpc_cslot below; S = 1: one lane per check,
	// slot(k, c') = cgbase[c' >> 5] + 32 k + (c' & 31)).  A warp that owns a task walks its edges with a constant +32 stride and
	// touches 32 consecutive words per step.
	uint32_t off_cdeg;      // u8 [P]             degree of sorted check c'
	uint32_t off_cgbase;    // u32[MB_MAX_GROUPS] first slot of each group of 32 checks
	uint32_t off_edge_var;  // u16[c_slots]       internal variable index of each check-side slot (0xFFFF = padding)
	// Variable side, same layout over variables renumbered by degree (descending): vslot(k, v') = vgbase[v' >> 5] + 32 k + (v' & 31)
	uint32_t off_vdeg;      // u8 [N]
	uint32_t off_vgbase;    // u32[MB_MAX_GROUPS]
	uint32_t off_vedge;     // u16[v_slots]       check-side slot id held by each variable-side slot (padding = c_slots: an always-zero message)
	uint32_t off_var_of_cw; // u16[N]             codeword position -> internal variable index
	uint32_t off_check_of_sorted; // u16[P]       sorted check c' -> reference check index (diagnostics, TX encoder)
	uint32_t off_vgdeg;     // u8 [MB_MAX_GROUPS] padded degree (largest in the group, rounded up to even) of each group of 32 variables
	// The decoder kernel's own tables: BYTE offsets into its shared-memory arrays (no index scaling per edge), padding that is
	// the neutral element of every reduction (so the loops need no per-thread degree), and static schedules that balance the
	// padded degrees of the groups over the CTA's warps (longest-processing-time first), so no warp idles at the barriers.
	uint32_t off_edge_varb; // u16[c_slots]  8 * internal variable index of each check-side slot (padding: 8 * N, the +inf variable): byte offset of the float2 (frame pair)
	uint32_t off_vedgeb;    // u16[v_slots]  8 * check-side slot id held by each variable-side slot (padding: 8 * c_slots, an always-zero message)
	uint32_t off_csched;    // u32[MB_LDPC_WARPS][MB_SCHED_LEN] check tasks of each warp (MB_CDESC_*), sorted by body, 0 ends; last word: tasks per body, 4 bits each from body 2
	uint32_t off_vsched;    // u32[MB_LDPC_WARPS][MB_SCHED_LEN] variable groups (degree > 2 part) of each warp: first slot | padded degree << 16 | (group + 1) << 24
	uint32_t off_vtail;     // u32[N - vtail_start] variables vtail_start.. (degree <= 2): byte offsets of their two messages, low | high << 16
	int32_t vtail_start;    // multiple of 32
};

struct MbMode {
	int32_t config, M, bps, rate_idx, rate_num;
	int32_t Nsymb, nData, nPilots, nBits, nReal, nVirtual, K, P;
	int32_t frame_bytes, estimator /*0 ZF, 1 LS*/, phase_only, preamble_nSymb;
	int32_t crc_bytes;      // nReal/8: bytes covered by the CRC self check
	int32_t crc_reserved;   // (was: bytes per lane of the chunked CRC)
	uint32_t crc_init;      // contribution of the 0xFFFF preset after crc_bytes bytes
	float boost;
	uint32_t off_pinv;       // f32[cells]  1/p at pilot cells, 0 at data cells
	uint32_t off_pval;       // f32[cells]  p at pilot cells, 0 at data cells
	uint32_t off_invn;       // f32[cells]  1/(pilots inside the clipped LS window) at pilot cells
	uint32_t off_pilot_cell; // u16[nPilots] row-major pilot cell indices
	uint32_t off_sym_cell;   // u16[nData]  grid cell feeding demapped symbol q (deframe o T/F de-interleave)
	uint32_t off_llr_dst;    // u16[nBits]  internal variable receiving LLR i (bit de-interleave o expand o renumber)
	uint32_t off_llr_dst2;   // u16[nBits]  second destination for the virtual-bit copies, MB_NO_DST if none
	uint32_t off_const;      // f32[2*M]    constellation (re,im), unit mean power
	uint32_t off_bit_var;    // u16[8*crc_bytes] internal variable of info bit i
	uint32_t off_scr;        // u8 [N]      scrambler bit i (bit_energy_dispersal sequence)
	uint32_t off_crcbit;     // u16[8*crc_bytes] CRC register (preset 0) after all crc_bytes bytes of a message with only bit i set: the CRC is their XOR over the set bits
	// ---- descriptors of the persistent demodulator kernel (mb_demod.cu): all lattice / window / interleaver arithmetic is
	// resolved here, offsets are BYTE offsets into the kernel's shared-memory arrays so that no scaling is left per frame.
	uint32_t off_zf_src;     // u32[Nsymb*27] compact pilot-row slot -> (cell*8) | valid << 30 | (pilot negative) << 31
	uint32_t off_pilot_rec;  // u32[4*nPilots] row-major pilots: {hi0 | lo0 << 16, hi1 | lo1 << 16, hi2 | lo2 << 16, cell*8 | zslot*8 << 16}
	                         //   LS estimate = sum over row residues r of (PM[hi_r] - PM[lo_r]); PM = per-residue running sums over rows of
	                         //   the clipped 21-column window sums, [Nsymb+1][18] float2 (row Nsymb is all zero = "no lower bound")
	uint32_t off_pilot_f;    // f32[2*nPilots] {1 / (pilots inside the clipped 21x21 window), pilot value}
	uint32_t off_data_rec;   // u32[data_rec_words*nData] data cells in GRID (deframer) order: word 0 = cell*8 | zslot(r0)*8 << 15 | (t+2) << 29
	                         //   with channel = H[r0] + (H[r0+3] - H[r0]) * t / 3 (interpolator.cc:163-254 resolved on the host), then
	                         //   one u16 per emitted LLR (MSB first): byte offset of its destination in the hand-off LLR vector (MB_HANDOFF)
	uint32_t off_virt;       // u16[2*nVirtual] (source byte offset, destination byte offset) of the virtual-bit copies (telecom_system.cc:1303-1306)
	int32_t data_rec_words;  // 2 for bps <= 2, 4 above
	float pinv_mag;          // |1/p| as float (pilot boost 1.33)
	uint32_t off_pilot_neg;  // u64[Nsymb]   bit c of word s set <=> the pilot at (s, c) is negative (-boost)
};

struct MbBlobHeader {
	uint32_t magic, version, total_bytes, reserved;
	uint32_t off_twiddle;    // f32[2*256]  tw[k1*16+n2] = exp(-2 pi i n2 k1/256)/256
	uint32_t pad[3];
	MbMode modes[MB_NMODES];
	MbRate rates[MB_NRATES];
};

#include <string>
#include <vector>
// Builds the blob from the LDPC table file. Returns an empty string on success, else an error message.
std::string mb_build_blob(const char *ldpc_blob_path, std::vector<uint8_t> &out);
// Sanity-check an imported blob (magic / version / size / offset bounds).
std::string mb_validate_blob(const uint8_t *blob, size_t size);
// glibc TYPE_3 random() as vendored by the reference (os_interop.cc:100-283); state[34] is the ring cursor.
void mb_srandom(uint32_t state[35], unsigned seed);
int mb_random(uint32_t state[35]);
int mb_rate_index(int rate_num);
// Check-side layout of the decoder (mb_ldpc.cu).  A group of 32 sorted checks whose largest degree d exceeds MB_LDPC_DMAX is SPLIT over
// S = 2, 4 or 8 lanes per check (the smallest S with ceil(d / S) <= MB_LDPC_DMAX): the group becomes S warp tasks of 32 / S checks, each lane
// holds Dp = ceil(d / S) edges of its check in registers and the lanes of a check combine their partial results by warp shuffles.
// So every lane of every task runs one of the fully unrolled bodies (3..MB_LDPC_DMAX edges), whatever the check degree (up to 46), and the
// tasks are small enough to balance over the warps.  Groups up to MB_LDPC_DMAX are one task, one lane per check (S = 1, Dp = d).
//   task t of group g: checks g * 32 + t * (32 / S) + cl, cl < 32 / S; lane l = j * (32 / S) + cl holds edge positions p = j * Dp + k, k < Dp
//   slot(p, c') = cgbase[g] + (t * Dp + k) * 32 + l
#define MB_LDPC_DMAX 7
inline void mb_ldpc_split(int d, int *S, int *Dp)
{
	int s = 1;
	while ((d + s - 1) / s > MB_LDPC_DMAX) s *= 2;
	*S = s, *Dp = (d + s - 1) / s;
}
// slot of edge position p of sorted check cs; d_first = degree of the first (largest) check of its group
inline uint32_t mb_ldpc_cslot(const uint32_t *cgbase, int d_first, int cs, int p)
{
	int S, Dp;
	mb_ldpc_split(d_first, &S, &Dp);
	const int per = 32 / S, ci = cs & 31, t = ci / per, cl = ci % per, j = p / Dp, k = p % Dp;
	return cgbase[cs >> 5] + (uint32_t)((t * Dp + k) * 32 + j * per + cl);
}
// check-schedule descriptor: first slot of the task | Dp << 16 | log2(S) << 20 | t << 22 | (group + 1) << 25
#define MB_CDESC_BASE(x) ((x) & 0xFFFFu)
#define MB_CDESC_DP(x) (((x) >> 16) & 0xFu)
#define MB_CDESC_LOG2S(x) (((x) >> 20) & 0x3u)
#define MB_CDESC_TASK(x) (((x) >> 22) & 0x7u)
#define MB_CDESC_GROUP(x) (((x) >> 25) - 1u)
#define MB_CDESC_BODY(x) (MB_CDESC_DP(x) + (MB_CDESC_LOG2S(x) != 0u ? 1u : 0u))  // which unrolled body runs the task (mb_ldpc.cu)

// Tone plan of a ROBUST (MFSK) mode: cl_mfsk (include/physical_layer/mfsk.h, mfsk.cc:49-160).
struct MbMfsk {
	int32_t M, nBits, nStreams, tone_hop_step;
	int32_t stream_offsets[4], preamble_tones[4], ack_tones[8], break_tones[8];
};
// MbMode records + tables of ROBUST_0..2, as an extension region that starts `base` bytes after the start of the blob.
std::string mb_build_mfsk_ext(const std::vector<uint8_t> &blob, uint32_t base, MbMode modes[3], MbMfsk tones[3], std::vector<uint8_t> &ext);
