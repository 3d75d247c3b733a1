// mb_api.cu -- the C ABI (include/mercury_b200.h) over the two kernels: handle, resident tables, O(1) mode
// switch, device-resident batch entry points and the pipelined host-buffer batch path.
//
// Mirrors the reference's cl_telecom_system surface for the RX tail (see the header for file:line anchors).
// There is deliberately no CPU implementation behind any compute entry point.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mercury_b200.h"
#include "mb_kernels.cuh"
#include "mb_tables.h"

std::string mb_synth_frames(const std::vector<uint8_t> &blob, int config, size_t n_frames, uint64_t seed, double esn0_db,
			    const uint8_t *payload_in, float *baseband_out, uint8_t *payload_out, int n_threads);

static_assert(sizeof(MbRxStats) == sizeof(mercury_b200_rx_stats), "stats record layout");
static_assert(sizeof(MbRxStats) == 32, "stats record size");
static_assert(MB_HANDOFF_STRIDE == MERCURY_B200_HANDOFF_FLOATS, "hand-off stride");
static_assert(sizeof(MbMfskPatternResult) == sizeof(mercury_b200_mfsk_pattern_result) && sizeof(MbMfskPatternResult) == 32, "pattern result layout");
static_assert(sizeof(MbReceiveStats) == sizeof(mercury_b200_receive_stats) && sizeof(MbReceiveStats) == 72, "receive stats record layout");
static_assert(MB_MFSK_FIXED_DELAY_FLAG == MERCURY_B200_MFSK_FIXED_DELAY_FLAG, "mfsk_fixed_delay encoding");

extern "C" int mercury_b200_reset_tx_stream(mercury_b200_t *h);

namespace {
constexpr int kSlots = 3;

struct Slot {
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	void *d_x = nullptr, *d_llr = nullptr, *d_payload = nullptr, *d_stats = nullptr, *d_llr_cw = nullptr;
	size_t cap_frames = 0, cap_x = 0;
	bool has_llr_cw = false;
};
}  // namespace

namespace {
// Device workspace of the RX front-end (mb_frontend.cu), sized for cap captures of buf pass-band samples.
struct FeWork {
	size_t cap = 0;
	int buf = 0;
	int symb = 0;  // symbols per frame the per-capture frame / debug buffers were sized for
	int win_n = 0; // samples per capture of the on-demand window and its prefix sums (kFeWin; (2 pre + S) symbols with the coarse frequency search)
	void *d_x[2] = {nullptr, nullptr};  // staged pass-band samples of the host entry point, double buffered
	size_t x_bytes = 0;
	size_t out_slot_bytes = 0;
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
	size_t host_chunk = 384;  // captures per H2D chunk of the host entry point (copy of chunk i+1 overlaps the kernels of chunk i)
	MbFeState *st = nullptr;
	double2 *bbi = nullptr, *win = nullptr, *dbg_bb = nullptr;
	double *energy_part = nullptr, *vals = nullptr, *pref_ts = nullptr, *pref_win = nullptr, *tile_base = nullptr;
	uint8_t *flags = nullptr;
	float2 *frames = nullptr;
	float *llr = nullptr;
	MbRxStats *tail_stats = nullptr;
	uint8_t *tail_payload = nullptr, *payload = nullptr;
	MbReceiveStats *stats = nullptr;
	int32_t *counters = nullptr, *h_counters = nullptr;
	double2 *carrier = nullptr;  // (cos, sin)(2 pi fc i Ts), host libm
	int carrier_n = 0;
	bool want_dbg = false;
	cudaStream_t stream = nullptr;
};
constexpr int kFeVals = 4 * MB_FE_SYM;       // fine search: (pre + 4) - pre symbols of candidate positions, step 1
constexpr int kFeWin = 8 * MB_FE_SYM;        // fine-sync window, (pre + 4) symbols, pre <= 4
void fe_free(FeWork &w);
}  // namespace

namespace {
// TX chain (mb_tx.cu): per-mode tables built lazily on the first transmit of the mode, plus the batch workspace.
struct TxWork {
	bool built[MB_NMODES + 3] = {};  // + ROBUST_0..2
	MbTxMode mode_host[MB_NMODES + 3];
	MbTxMode *mode_dev[MB_NMODES + 3] = {};
	uint8_t *tables[MB_NMODES + 3] = {};
	size_t cap = 0, cap_total = 0;
	uint8_t *payload = nullptr, *dbg_cw = nullptr;
	unsigned long long *start = nullptr;
	double2 *bb = nullptr;
	double *pb = nullptr, *p1 = nullptr, *power_part = nullptr;
	void *out = nullptr;  // two output slots of cap * cap_total doubles: the D2H copy of one chunk overlaps the kernels of the next (host batch path)
	size_t out_slot_bytes = 0;
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
	double *stream_buf = nullptr;  // passband_data_tx_buffer: three frames of the streaming message locations, + 4 frames of scratch
	int stream_total = 0, stream_slot = -1;
	cudaStream_t stream = nullptr;
	bool init_done = false;
};
void tx_free(TxWork &w);
}  // namespace

struct mercury_b200 {
	bool coarse_freq_sync = false;  // g_gui_state.coarse_freq_sync_enabled (gui_state.h:143): the optional +-30 Hz search of trial 1 (telecom_system.cc:949-1013)
	bool mfsk_ctrl = false;  // set_mfsk_ctrl_mode: shortened control frames in ROBUST_0 / ROBUST_1 (telecom_system.cc:1572-1585, 2966-2995)
	MbMode mfsk_modes[3];  // ROBUST_0..2 (config 100..102): tables in the extension region behind the device blob
	MbMfsk mfsk_tones[3];
	double *d_mfsk_energies = nullptr;
	MbMfskPatternResult *d_mfsk_out = nullptr;
	void *d_mfsk_in = nullptr;
	size_t mfsk_cap_bytes = 0, mfsk_cap_buffers = 0, mfsk_cap_energies = 0;
	TxWork tx;
	FeWork fe;
	MbFeConst fe_const;
	bool fe_ready = false;
	size_t fe_chunk = 1024;
	uint64_t fe_rounds = 0, fe_exact = 0;
	int device = -1;
	std::vector<uint8_t> blob;
	MbBlobHeader hdr;
	uint8_t *d_blob = nullptr;
	int config = -1, ldpc_iters = 50, decoder = MERCURY_B200_DECODER_SPA;
	int cheap_test_threads = -1;  // tuning knob of the decoder's early syndrome test (MERCURY_B200_CHEAP_TEST in the environment); -1: by rate
	std::string err;
	uint64_t launches = 0;
	Slot slots[kSlots];
	void *d_scratch_llr = nullptr;  // internal-order LLRs of demod_decode_batch_device
	size_t scratch_frames = 0;
	float2 *dbg_Y = nullptr, *dbg_H = nullptr, *dbg_Z = nullptr;
	float *h_stage = nullptr;  // pinned staging for receive_baseband
	size_t h_stage_bytes = 0;
	// work queues of the persistent decoder kernel: two device words per stream (launches on one stream serialise; the kernel re-arms them)
	std::vector<std::pair<cudaStream_t, unsigned *>> ldpc_queues;
};

namespace {

inline bool is_mfsk_config(int c) { return c >= 100 && c <= 102; }
inline const MbMode &cur_mode(const mercury_b200_t *h) { return is_mfsk_config(h->config) ? h->mfsk_modes[h->config - 100] : h->hdr.modes[h->config]; }
// get_active_nsymb() (telecom_system.cc:1577-1580): ctrl_nBits = 1200 (ROBUST_0), 1400 (ROBUST_1), none otherwise (:2966-2990)
inline int active_nsymb(const mercury_b200_t *h)
{
	const MbMode &m = cur_mode(h);
	if (!h->mfsk_ctrl || m.M != 200) return m.Nsymb;
	const int ctrl_bits = h->config == 100 ? 1200 : (h->config == 101 ? 1400 : 0);
	return ctrl_bits > 0 ? ctrl_bits / m.bps : m.Nsymb;
}

int fail(mercury_b200_t *h, int code, const std::string &msg)
{
	if (h) h->err = msg;
	return code;
}

int cuda_fail(mercury_b200_t *h, cudaError_t e, const char *what)
{
	return fail(h, MERCURY_B200_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define MB_CUDA(h, call)                                                  \
	do {                                                              \
		cudaError_t e__ = (call);                                 \
		if (e__ != cudaSuccess) return cuda_fail(h, e__, #call);  \
	} while (0)

int upload_blob(mercury_b200_t *h)
{
	memcpy(&h->hdr, h->blob.data(), sizeof(MbBlobHeader));
	if (h->d_blob) cudaFree(h->d_blob);
	h->d_blob = nullptr;
	// the MFSK modes' tables are derived from the blob and appended behind its device copy (mb_build_mfsk_ext)
	const uint32_t base = (uint32_t)((h->blob.size() + 255) / 256 * 256);
	std::vector<uint8_t> ext;
	const std::string e = mb_build_mfsk_ext(h->blob, base, h->mfsk_modes, h->mfsk_tones, ext);
	if (!e.empty()) return fail(h, MERCURY_B200_EINVAL, e);
	MB_CUDA(h, cudaMalloc(&h->d_blob, base + ext.size()));
	MB_CUDA(h, cudaMemcpy(h->d_blob, h->blob.data(), h->blob.size(), cudaMemcpyHostToDevice));
	MB_CUDA(h, cudaMemcpy(h->d_blob + base, ext.data(), ext.size(), cudaMemcpyHostToDevice));
	return MERCURY_B200_OK;
}

int check_ready(mercury_b200_t *h)
{
	if (!h) return MERCURY_B200_EINVAL;
	if (!h->d_blob) return fail(h, MERCURY_B200_ESTATE, "tables not loaded (mercury_b200_load_tables / import_tables)");
	if (h->config < 0) return fail(h, MERCURY_B200_ESTATE, "no configuration selected (mercury_b200_load_configuration)");
	cudaError_t e = cudaSetDevice(h->device);
	if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
	return MERCURY_B200_OK;
}

int launch_demod(mercury_b200_t *h, const void *d_x, size_t n, void *d_llr, void *d_stats, void *d_llr_cw, size_t dbg_frame_off, cudaStream_t s,
		 bool gi_removed = false, int fmt = MERCURY_B200_BASEBAND_C64, float scale = 1.0f)
{
	const MbMode &m = cur_mode(h);
	if (fmt < MERCURY_B200_BASEBAND_C64 || fmt > MERCURY_B200_BASEBAND_CF16) return fail(h, MERCURY_B200_EINVAL, "unknown base-band sample format");
	if (m.M == 200 && fmt != MERCURY_B200_BASEBAND_C64) return fail(h, MERCURY_B200_EINVAL, "the ROBUST (MFSK) tail takes complex64 samples");
	if (m.M == 200) {  // ROBUST modes: FFT + non-coherent tone detection instead of the coherent OFDM demodulator
		MbMfskArgs f;
		memset(&f, 0, sizeof(f));
		f.x = static_cast<const float2 *>(d_x);
		f.sym_stride = gi_removed ? MB_NFFT : MB_NOFDM, f.sym_skip = gi_removed ? 0 : MB_NGI;
		f.llr = static_cast<float *>(d_llr), f.llr_cw = static_cast<float *>(d_llr_cw), f.stats = static_cast<MbRxStats *>(d_stats);
		f.blob = h->d_blob, f.mode = m, f.tone = h->mfsk_tones[h->config - 100], f.active_nsymb = active_nsymb(h);
		cudaError_t e = mb_launch_mfsk_demod(f, n, s);
		if (e != cudaSuccess) return cuda_fail(h, e, "mfsk demod kernel launch");
		h->launches++;
		return MERCURY_B200_OK;
	}
	MbDemodArgs a;
	memset(&a, 0, sizeof(a));
	a.x = static_cast<const float2 *>(d_x);
	a.sym_stride = gi_removed ? MB_NFFT : MB_NOFDM;
	a.sym_skip = gi_removed ? 0 : MB_NGI;
	a.x_format = fmt;
	a.x_scale = scale;
	a.llr = static_cast<float *>(d_llr);
	a.llr_cw = static_cast<float *>(d_llr_cw);
	a.stats = static_cast<MbRxStats *>(d_stats);
	const size_t cells = (size_t)m.Nsymb * MB_NC;
	a.dbg_Y = h->dbg_Y ? h->dbg_Y + dbg_frame_off * cells : nullptr;
	a.dbg_H = h->dbg_H ? h->dbg_H + dbg_frame_off * cells : nullptr;
	a.dbg_Z = h->dbg_Z ? h->dbg_Z + dbg_frame_off * cells : nullptr;
	a.blob = h->d_blob;
	a.off_twiddle = h->hdr.off_twiddle;
	a.off_var_of_cw = h->hdr.rates[m.rate_idx].off_var_of_cw;
	a.mode = m;
	cudaError_t e = mb_launch_demod(a, n, s);
	if (e != cudaSuccess) return cuda_fail(h, e, "demod kernel launch");
	h->launches++;
	return MERCURY_B200_OK;
}

int launch_ldpc(mercury_b200_t *h, const void *d_llr, size_t n, void *d_payload, void *d_stats, cudaStream_t s)
{
	MbLdpcArgs a;
	memset(&a, 0, sizeof(a));
	const MbMode &m = cur_mode(h);
	a.llr = static_cast<const float *>(d_llr);
	a.payload = static_cast<uint8_t *>(d_payload);
	a.stats = static_cast<MbRxStats *>(d_stats);
	a.blob = h->d_blob;
	a.mode = m;
	a.rate = h->hdr.rates[m.rate_idx];
	a.max_iters = h->ldpc_iters;
	a.check_gate = 1;
	// measured on B200 (two frames per CTA share the instruction stream, so the test only buys an earlier refill): it pays from rate 8/16 up (+5 %), costs 1-13 % below
	a.cheap_test_threads = h->cheap_test_threads >= 0 ? h->cheap_test_threads : (a.rate.rate_num >= 8 ? 16 : 0);
	for (const auto &q : h->ldpc_queues)
		if (q.first == s) a.queue = q.second;
	if (!a.queue) {
		MB_CUDA(h, cudaMalloc(&a.queue, 256 + mb_ldpc_max_ctas() * 2 * MB_N * sizeof(float)));  // queue words + per-CTA channel-LLR scratch
		MB_CUDA(h, cudaMemset(a.queue, 0, 256));
		h->ldpc_queues.emplace_back(s, a.queue);
	}
	a.lch_scratch = reinterpret_cast<float *>(reinterpret_cast<char *>(a.queue) + 256);
	cudaError_t e = mb_launch_ldpc(a, n, h->decoder, s);
	if (e != cudaSuccess) return cuda_fail(h, e, "ldpc kernel launch");
	h->launches++;
	return MERCURY_B200_OK;
}

void free_slot(Slot &s)
{
	if (s.d_x) cudaFree(s.d_x);
	if (s.d_llr) cudaFree(s.d_llr);
	if (s.d_payload) cudaFree(s.d_payload);
	if (s.d_stats) cudaFree(s.d_stats);
	if (s.d_llr_cw) cudaFree(s.d_llr_cw);
	s.d_x = s.d_llr = s.d_payload = s.d_stats = s.d_llr_cw = nullptr;
	s.cap_frames = s.cap_x = 0;
	s.has_llr_cw = false;
}

int ensure_slot(mercury_b200_t *h, Slot &s, size_t frames, size_t x_bytes, bool want_llr_cw)
{
	if (!s.stream) {
		MB_CUDA(h, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
		MB_CUDA(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
	}
	if (s.cap_frames >= frames && s.cap_x >= x_bytes && (!want_llr_cw || s.has_llr_cw)) return MERCURY_B200_OK;
	cudaStreamSynchronize(s.stream);
	free_slot(s);
	MB_CUDA(h, cudaMalloc(&s.d_x, x_bytes));
	MB_CUDA(h, cudaMalloc(&s.d_llr, frames * MB_HANDOFF_STRIDE * sizeof(float)));
	MB_CUDA(h, cudaMalloc(&s.d_payload, frames * 256));
	MB_CUDA(h, cudaMalloc(&s.d_stats, frames * sizeof(MbRxStats)));
	if (want_llr_cw) MB_CUDA(h, cudaMalloc(&s.d_llr_cw, frames * MB_N * sizeof(float)));
	s.cap_frames = frames;
	s.cap_x = x_bytes;
	s.has_llr_cw = want_llr_cw;
	return MERCURY_B200_OK;
}

}  // namespace

extern "C" {

const char *mercury_b200_version(void) { return "mercury_b200 0.1 (sm_100a)"; }

const char *mercury_b200_strerror(int code)
{
	switch (code) {
	case MERCURY_B200_OK: return "ok";
	case MERCURY_B200_EINVAL: return "invalid argument";
	case MERCURY_B200_ENODEV: return "no usable CUDA device (no CPU fallback exists)";
	case MERCURY_B200_ECUDA: return "CUDA error";
	case MERCURY_B200_ESTATE: return "tables or configuration not loaded";
	case MERCURY_B200_EIO: return "LDPC table file missing or malformed";
	case MERCURY_B200_ENOMEM: return "out of memory";
	}
	return "unknown error";
}

int mercury_b200_create(int device, mercury_b200_t **out)
{
	if (!out) return MERCURY_B200_EINVAL;
	*out = nullptr;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
		cudaGetLastError();
		return MERCURY_B200_ENODEV;
	}
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MERCURY_B200_ENODEV;
	if (prop.major != 10) return MERCURY_B200_ENODEV;  // the kernels are compiled for sm_100a only
	if (cudaSetDevice(device) != cudaSuccess) return MERCURY_B200_ENODEV;
	if (mb_demod_init() != cudaSuccess || mb_ldpc_init() != cudaSuccess) {
		cudaGetLastError();
		return MERCURY_B200_ENODEV;
	}
	mercury_b200_t *h = new (std::nothrow) mercury_b200;
	if (!h) return MERCURY_B200_ENOMEM;
	h->device = device;
	if (const char *e = getenv("MERCURY_B200_CHEAP_TEST")) h->cheap_test_threads = std::max(0, std::min(256, atoi(e)));
	*out = h;
	return MERCURY_B200_OK;
}

void mercury_b200_destroy(mercury_b200_t *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	for (Slot &s : h->slots) {
		if (s.stream) cudaStreamSynchronize(s.stream);
		free_slot(s);
		if (s.done) cudaEventDestroy(s.done);
		if (s.stream) cudaStreamDestroy(s.stream);
	}
	if (h->d_scratch_llr) cudaFree(h->d_scratch_llr);
	for (const auto &q : h->ldpc_queues) cudaFree(q.second);
	if (h->d_blob) cudaFree(h->d_blob);
	if (h->d_mfsk_energies) cudaFree(h->d_mfsk_energies);
	if (h->d_mfsk_out) cudaFree(h->d_mfsk_out);
	if (h->d_mfsk_in) cudaFree(h->d_mfsk_in);
	if (h->h_stage) cudaFreeHost(h->h_stage);
	fe_free(h->fe);
	tx_free(h->tx);
	delete h;
}

const char *mercury_b200_last_error(const mercury_b200_t *h) { return h ? h->err.c_str() : "null handle"; }

int mercury_b200_build_tables_host(const char *path, void *buf, size_t *size)
{
	if (!path || !size) return MERCURY_B200_EINVAL;
	std::vector<uint8_t> blob;
	std::string e = mb_build_blob(path, blob);
	if (!e.empty()) return MERCURY_B200_EIO;
	if (buf) {
		if (*size < blob.size()) return MERCURY_B200_EINVAL;
		memcpy(buf, blob.data(), blob.size());
	}
	*size = blob.size();
	return MERCURY_B200_OK;
}

int mercury_b200_load_tables(mercury_b200_t *h, const char *path)
{
	if (!h || !path) return MERCURY_B200_EINVAL;
	std::string e = mb_build_blob(path, h->blob);
	if (!e.empty()) return fail(h, MERCURY_B200_EIO, e);
	MB_CUDA(h, cudaSetDevice(h->device));
	return upload_blob(h);
}

int mercury_b200_export_tables(const mercury_b200_t *h, void *buf, size_t *size)
{
	if (!h || !size) return MERCURY_B200_EINVAL;
	if (h->blob.empty()) return MERCURY_B200_ESTATE;
	if (buf) {
		if (*size < h->blob.size()) return MERCURY_B200_EINVAL;
		memcpy(buf, h->blob.data(), h->blob.size());
	}
	*size = h->blob.size();
	return MERCURY_B200_OK;
}

int mercury_b200_import_tables(mercury_b200_t *h, const void *buf, size_t size)
{
	if (!h || !buf) return MERCURY_B200_EINVAL;
	std::string e = mb_validate_blob(static_cast<const uint8_t *>(buf), size);
	if (!e.empty()) return fail(h, MERCURY_B200_EINVAL, e);
	h->blob.assign(static_cast<const uint8_t *>(buf), static_cast<const uint8_t *>(buf) + size);
	MB_CUDA(h, cudaSetDevice(h->device));
	return upload_blob(h);
}

// NCCL through dlopen: ncclBroadcast(sendbuff, recvbuff, count, ncclChar = 0, root, comm, stream), ncclCommUserRank(comm, int *)
int mercury_b200_broadcast_tables(mercury_b200_t *h, void *comm, int root, void *stream)
{
	if (!h || !comm) return MERCURY_B200_EINVAL;
	typedef int (*bcast_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
	typedef int (*rank_fn)(void *, int *);
	static bcast_fn p_bcast = nullptr;
	static rank_fn p_rank = nullptr;
	if (!p_bcast) {
		void *so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!so) return fail(h, MERCURY_B200_EIO, std::string("libnccl.so.2 not found: ") + dlerror());
		p_rank = reinterpret_cast<rank_fn>(dlsym(so, "ncclCommUserRank"));
		p_bcast = reinterpret_cast<bcast_fn>(dlsym(so, "ncclBroadcast"));
		if (!p_bcast || !p_rank) return fail(h, MERCURY_B200_EIO, "ncclBroadcast / ncclCommUserRank not found in libnccl");
	}
	MB_CUDA(h, cudaSetDevice(h->device));
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	int rank = -1;
	if (p_rank(comm, &rank) != 0) return fail(h, MERCURY_B200_ECUDA, "ncclCommUserRank failed");
	if (rank == root && h->blob.empty()) return fail(h, MERCURY_B200_ESTATE, "the root's tables are not loaded (mercury_b200_load_tables)");
	unsigned long long *d_n = nullptr;
	uint8_t *d_buf = nullptr;
	unsigned long long n = rank == root ? (unsigned long long)h->blob.size() : 0ull;
	MB_CUDA(h, cudaMalloc(&d_n, sizeof(n)));
	MB_CUDA(h, cudaMemcpyAsync(d_n, &n, sizeof(n), cudaMemcpyHostToDevice, s));
	if (p_bcast(d_n, d_n, sizeof(n), /*ncclChar*/ 0, root, comm, s) != 0) return fail(h, MERCURY_B200_ECUDA, "ncclBroadcast (size) failed");
	MB_CUDA(h, cudaMemcpyAsync(&n, d_n, sizeof(n), cudaMemcpyDeviceToHost, s));
	MB_CUDA(h, cudaStreamSynchronize(s));
	cudaFree(d_n);
	if (n < sizeof(MbBlobHeader) || n > (64ull << 20)) return fail(h, MERCURY_B200_EINVAL, "broadcast table blob has an implausible size");
	MB_CUDA(h, cudaMalloc(&d_buf, n));
	if (rank == root) MB_CUDA(h, cudaMemcpyAsync(d_buf, h->blob.data(), n, cudaMemcpyHostToDevice, s));
	if (p_bcast(d_buf, d_buf, n, 0, root, comm, s) != 0) return fail(h, MERCURY_B200_ECUDA, "ncclBroadcast (blob) failed");
	int rc = MERCURY_B200_OK;
	if (rank != root) {
		std::vector<uint8_t> host(n);
		MB_CUDA(h, cudaMemcpyAsync(host.data(), d_buf, n, cudaMemcpyDeviceToHost, s));
		MB_CUDA(h, cudaStreamSynchronize(s));
		rc = mercury_b200_import_tables(h, host.data(), host.size());  // validates the blob before anything dereferences it
	} else {
		MB_CUDA(h, cudaStreamSynchronize(s));
	}
	cudaFree(d_buf);
	return rc;
}

int mercury_b200_load_configuration(mercury_b200_t *h, int config, int ldpc_iters)
{
	if (!h) return MERCURY_B200_EINVAL;
	if (h->blob.empty()) return fail(h, MERCURY_B200_ESTATE, "tables not loaded");
	// telecom_system.cc:2494-2497: out-of-range configurations are ignored by the reference; here they are an error
	if (!is_mfsk_config(config) && (config < 0 || config >= MB_NMODES))
		return fail(h, MERCURY_B200_EINVAL, "configuration must be 0..16 (CONFIG_0..CONFIG_16) or 100..102 (ROBUST_0..ROBUST_2)");
	if (h->config != config) mercury_b200_reset_tx_stream(h);
	h->config = config;
	h->mfsk_ctrl = false;  // telecom_system.cc:2989,3000
	h->ldpc_iters = std::min(50, std::max(5, ldpc_iters));  // main.cc:303-311
	return MERCURY_B200_OK;
}

int mercury_b200_set_decoder(mercury_b200_t *h, int decoder)
{
	if (!h || (decoder != MERCURY_B200_DECODER_SPA && decoder != MERCURY_B200_DECODER_MINSUM)) return MERCURY_B200_EINVAL;
	h->decoder = decoder;
	return MERCURY_B200_OK;
}

int mercury_b200_get_geometry(const mercury_b200_t *h, mercury_b200_geometry *g)
{
	if (!h || !g) return MERCURY_B200_EINVAL;
	if (h->blob.empty() || h->config < 0) return MERCURY_B200_ESTATE;
	const MbMode &m = cur_mode(h);
	const MbRate &r = h->hdr.rates[m.rate_idx];
	g->config = m.config, g->M = m.M, g->bits_per_symbol = m.bps, g->ldpc_rate_num = m.rate_num;
	g->Nsymb = m.Nsymb, g->Nc = MB_NC, g->Nfft = MB_NFFT, g->Ngi = MB_NGI, g->Nofdm = MB_NOFDM;
	g->nData = m.nData, g->nPilots = m.nPilots, g->nBits = m.nBits;
	g->N = MB_N, g->K = m.K, g->P = m.P, g->nReal = m.nReal, g->nVirtual = m.nVirtual;
	g->preamble_nSymb = m.preamble_nSymb, g->frame_bytes = m.frame_bytes;
	g->estimator = m.estimator, g->phase_only = m.phase_only;
	g->ldpc_iters = h->ldpc_iters, g->ldpc_edges = r.n_edges, g->decoder = h->decoder;
	return MERCURY_B200_OK;
}

int mercury_b200_get_frame_size_bytes(const mercury_b200_t *h)
{
	if (!h || h->blob.empty() || h->config < 0) return MERCURY_B200_ESTATE;
	return cur_mode(h).frame_bytes;
}

int mercury_b200_get_frame_size_bits(const mercury_b200_t *h)
{
	if (!h || h->blob.empty() || h->config < 0) return MERCURY_B200_ESTATE;
	return cur_mode(h).nReal - 16;
}

int mercury_b200_set_debug_capture(mercury_b200_t *h, void *d_Y, void *d_H, void *d_Z)
{
	if (!h) return MERCURY_B200_EINVAL;
	h->dbg_Y = static_cast<float2 *>(d_Y);
	h->dbg_H = static_cast<float2 *>(d_H);
	h->dbg_Z = static_cast<float2 *>(d_Z);
	return MERCURY_B200_OK;
}

int mercury_b200_demod_batch_device(mercury_b200_t *h, const void *d_x, size_t n, void *d_llr, void *d_stats, void *d_llr_cw, void *stream)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!d_x || !d_llr || !d_stats) return fail(h, MERCURY_B200_EINVAL, "null device buffer");
	return launch_demod(h, d_x, n, d_llr, d_stats, d_llr_cw, 0, static_cast<cudaStream_t>(stream));
}

int mercury_b200_ldpc_decode_batch_device(mercury_b200_t *h, const void *d_llr, size_t n, void *d_payload, void *d_stats, void *stream)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!d_llr || !d_payload || !d_stats) return fail(h, MERCURY_B200_EINVAL, "null device buffer");
	return launch_ldpc(h, d_llr, n, d_payload, d_stats, static_cast<cudaStream_t>(stream));
}

int mercury_b200_demod_decode_batch_device(mercury_b200_t *h, const void *d_x, size_t n, void *d_payload, void *d_stats, void *d_llr_cw,
					   void *stream)
{
	return mercury_b200_demod_decode_batch_device_fmt(h, d_x, MERCURY_B200_BASEBAND_C64, 1.0f, n, d_payload, d_stats, d_llr_cw, stream);
}

int mercury_b200_demod_decode_batch_device_fmt(mercury_b200_t *h, const void *d_x, int fmt, float scale, size_t n, void *d_payload, void *d_stats,
					       void *d_llr_cw, void *stream)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!d_x || !d_payload || !d_stats) return fail(h, MERCURY_B200_EINVAL, "null device buffer");
	if (h->scratch_frames < n) {
		MB_CUDA(h, cudaDeviceSynchronize());
		if (h->d_scratch_llr) cudaFree(h->d_scratch_llr);
		h->d_scratch_llr = nullptr;
		h->scratch_frames = 0;
		MB_CUDA(h, cudaMalloc(&h->d_scratch_llr, n * MB_HANDOFF_STRIDE * sizeof(float)));
		h->scratch_frames = n;
	}
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	rc = launch_demod(h, d_x, n, h->d_scratch_llr, d_stats, d_llr_cw, 0, s, false, fmt, scale);
	if (rc) return rc;
	return launch_ldpc(h, h->d_scratch_llr, n, d_payload, d_stats, s);
}

int mercury_b200_demod_decode_batch(mercury_b200_t *h, const float *x, size_t n, uint8_t *payload, mercury_b200_rx_stats *stats, float *llr_cw)
{
	return mercury_b200_demod_decode_batch_fmt(h, x, MERCURY_B200_BASEBAND_C64, 1.0f, n, payload, stats, llr_cw);
}

int mercury_b200_demod_decode_batch_fmt(mercury_b200_t *h, const void *x, int fmt, float scale, size_t n, uint8_t *payload, mercury_b200_rx_stats *stats,
					float *llr_cw)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!x || !payload || !stats) return fail(h, MERCURY_B200_EINVAL, "null host buffer");
	if (fmt < MERCURY_B200_BASEBAND_C64 || fmt > MERCURY_B200_BASEBAND_CF16) return fail(h, MERCURY_B200_EINVAL, "unknown base-band sample format");
	const MbMode &m = cur_mode(h);
	const size_t esz = fmt == MERCURY_B200_BASEBAND_C64 ? sizeof(float2) : 4;  // bytes per complex sample
	const size_t frame_x = (size_t)m.Nsymb * MB_NOFDM * esz;
	// The guard interval never crosses PCIe: a strided (2-D) copy moves the 2,048 useful bytes of every 2,176-byte symbol, so the
	// device copy of a chunk is [frames][Nsymb][256] and the demodulator is told that the GI is already gone (-5.9 % of the bytes
	// of the link that bounds this path).
	const size_t sym_in = MB_NOFDM * esz, sym_dev = MB_NFFT * esz, gi = MB_NGI * esz;
	// chunks of ~64 MB of samples: large enough to fill the GPU (>= 8 CTAs per SM), small enough to pipeline
	size_t chunk = std::max<size_t>(1184, (64u << 20) / frame_x);
	chunk = std::min(chunk, n);
	for (Slot &s : h->slots) {
		rc = ensure_slot(h, s, chunk, chunk * frame_x, llr_cw != nullptr);
		if (rc) return rc;
	}
	size_t done = 0;
	int i = 0;
	while (done < n) {
		Slot &s = h->slots[i % kSlots];
		const size_t c = std::min(chunk, n - done);
		MB_CUDA(h, cudaEventSynchronize(s.done));
		MB_CUDA(h, cudaMemcpy2DAsync(s.d_x, sym_dev, reinterpret_cast<const uint8_t *>(x) + done * frame_x + gi, sym_in, sym_dev, c * (size_t)m.Nsymb,
					     cudaMemcpyHostToDevice, s.stream));
		rc = launch_demod(h, s.d_x, c, s.d_llr, s.d_stats, llr_cw ? s.d_llr_cw : nullptr, done, s.stream, /*gi_removed=*/true, fmt, scale);
		if (rc) return rc;
		rc = launch_ldpc(h, s.d_llr, c, s.d_payload, s.d_stats, s.stream);
		if (rc) return rc;
		MB_CUDA(h, cudaMemcpyAsync(payload + done * m.frame_bytes, s.d_payload, c * m.frame_bytes, cudaMemcpyDeviceToHost, s.stream));
		MB_CUDA(h, cudaMemcpyAsync(stats + done, s.d_stats, c * sizeof(MbRxStats), cudaMemcpyDeviceToHost, s.stream));
		if (llr_cw)
			MB_CUDA(h, cudaMemcpyAsync(llr_cw + done * MB_N, s.d_llr_cw, c * MB_N * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
		MB_CUDA(h, cudaEventRecord(s.done, s.stream));
		done += c;
		i++;
	}
	for (Slot &s : h->slots) MB_CUDA(h, cudaStreamSynchronize(s.stream));
	return MERCURY_B200_OK;
}

int mercury_b200_receive_baseband(mercury_b200_t *h, const double *baseband, int *out, mercury_b200_rx_stats *stats)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (!baseband || !out || !stats) return fail(h, MERCURY_B200_EINVAL, "null buffer");
	const MbMode &m = cur_mode(h);
	const size_t n = (size_t)m.Nsymb * MB_NOFDM * 2;
	const size_t need = n * sizeof(float) + 256 + sizeof(mercury_b200_rx_stats);
	if (h->h_stage_bytes < need) {
		if (h->h_stage) cudaFreeHost(h->h_stage);
		h->h_stage = nullptr;
		h->h_stage_bytes = 0;
		MB_CUDA(h, cudaMallocHost(&h->h_stage, need));
		h->h_stage_bytes = need;
	}
	for (size_t i = 0; i < n; i++) h->h_stage[i] = (float)baseband[i];
	uint8_t *pl = reinterpret_cast<uint8_t *>(h->h_stage + n);
	mercury_b200_rx_stats *st = reinterpret_cast<mercury_b200_rx_stats *>(pl + 256);
	rc = mercury_b200_demod_decode_batch(h, h->h_stage, 1, pl, st, nullptr);
	if (rc) return rc;
	for (int i = 0; i < m.frame_bytes; i++) out[i] = pl[i];  // one int per byte, like receive_byte() (telecom_system.cc:1329-1332)
	*stats = *st;
	return MERCURY_B200_OK;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * The whole receive_byte(): RX front-end (mb_frontend.cu) + tail.
 * ------------------------------------------------------------------------------------------------------------------- */
}  // extern "C"

namespace {

void fe_free(FeWork &w)
{
	void *ptrs[] = {w.d_x[0], w.d_x[1], w.st, w.bbi, w.win, w.dbg_bb, w.energy_part, w.vals, w.pref_ts, w.pref_win, w.flags, w.tile_base, w.frames, w.llr, w.tail_stats, w.tail_payload, w.payload, w.stats, w.counters};
	for (void *p : ptrs)
		if (p) cudaFree(p);
	if (w.h_counters) cudaFreeHost(w.h_counters);
	if (w.carrier) cudaFree(w.carrier);
	if (w.stream) cudaStreamDestroy(w.stream);
	if (w.copy_stream) cudaStreamDestroy(w.copy_stream);
	for (int i = 0; i < 2; i++) {
		if (w.copied[i]) cudaEventDestroy(w.copied[i]);
		if (w.consumed[i]) cudaEventDestroy(w.consumed[i]);
	}
	w = FeWork();
}

size_t fe_sample_bytes(int fmt)
{
	switch (fmt) {
	case MERCURY_B200_SAMPLES_F64: return 8;
	case MERCURY_B200_SAMPLES_F32: return 4;
	case MERCURY_B200_SAMPLES_I16: return 2;
	case MERCURY_B200_SAMPLES_I32: return 4;
	}
	return 0;
}

int fe_capture_samples(const MbMode &m) { return MB_NOFDM * mb_fe_buffer_nsymb(m.Nsymb, m.preamble_nSymb) * 4; }

int fe_ensure(mercury_b200_t *h, size_t n, int buf, const MbMode &m, bool want_dbg, size_t stage_bytes)
{
	FeWork &w = h->fe;
	if (!h->fe_ready) {
		mb_fe_host_const(&h->fe_const);
		cudaError_t e = mb_fe_init(h->fe_const);
		if (e != cudaSuccess) return cuda_fail(h, e, "front-end constants");
		if (const char *c = getenv("MERCURY_B200_FE_CHUNK")) h->fe_chunk = std::max(1, atoi(c));
		h->fe_ready = true;
	}
	if (!w.stream) {
		MB_CUDA(h, cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
		MB_CUDA(h, cudaStreamCreateWithFlags(&w.copy_stream, cudaStreamNonBlocking));
		for (int i = 0; i < 2; i++) {
			MB_CUDA(h, cudaEventCreateWithFlags(&w.copied[i], cudaEventDisableTiming));
			MB_CUDA(h, cudaEventCreateWithFlags(&w.consumed[i], cudaEventDisableTiming));
		}
		if (const char *c = getenv("MERCURY_B200_FE_HOST_CHUNK")) w.host_chunk = std::max(1, atoi(c));
	}
	if (w.carrier_n < buf) {
		if (w.carrier) cudaFree(w.carrier);
		w.carrier = nullptr;
		std::vector<double> cs((size_t)2 * buf);
		mb_fe_host_carrier(h->fe_const, cs.data(), buf);
		MB_CUDA(h, cudaMalloc(&w.carrier, cs.size() * sizeof(double)));
		MB_CUDA(h, cudaMemcpy(w.carrier, cs.data(), cs.size() * sizeof(double), cudaMemcpyHostToDevice));
		w.carrier_n = buf;
	}
	if (!w.h_counters) {
		MB_CUDA(h, cudaMallocHost(&w.h_counters, 4 * sizeof(int32_t)));
		MB_CUDA(h, cudaMalloc(&w.counters, 4 * sizeof(int32_t)));
	}
	if (w.x_bytes < stage_bytes) {
		MB_CUDA(h, cudaDeviceSynchronize());
		for (int i = 0; i < 2; i++) {
			if (w.d_x[i]) cudaFree(w.d_x[i]);
			w.d_x[i] = nullptr;
		}
		w.x_bytes = 0;
		for (int i = 0; i < 2; i++) MB_CUDA(h, cudaMalloc(&w.d_x[i], stage_bytes));
		w.x_bytes = stage_bytes;
	}
	const size_t bb_n = (size_t)(m.Nsymb + m.preamble_nSymb) * MB_NOFDM;
	const int symb = std::max(MB_MAX_SYMB, m.Nsymb);  // ROBUST_0 frames are 320 symbols: a workspace sized in an OFDM mode must not be reused for them
	const int win_need = h->coarse_freq_sync && m.M != 200 ? std::max(kFeWin, (2 * m.preamble_nSymb + m.Nsymb) * MB_FE_SYM) : kFeWin;  // :970-973 searches the head of the buffer
	if (w.cap >= n && w.buf >= buf && w.symb >= symb && w.win_n >= win_need && (!want_dbg || w.want_dbg)) return MERCURY_B200_OK;
	MB_CUDA(h, cudaDeviceSynchronize());
	void **ptrs[] = {(void **)&w.st, (void **)&w.bbi, (void **)&w.win, (void **)&w.dbg_bb, (void **)&w.energy_part, (void **)&w.vals, (void **)&w.frames,
			 (void **)&w.pref_ts, (void **)&w.pref_win, (void **)&w.flags, (void **)&w.tile_base,
			 (void **)&w.llr, (void **)&w.tail_stats, (void **)&w.tail_payload, (void **)&w.payload, (void **)&w.stats};
	for (void **p : ptrs) {
		if (*p) cudaFree(*p);
		*p = nullptr;
	}
	const int bufmax = std::max(buf, w.buf), symbmax = std::max(symb, w.symb), winmax = std::max(win_need, w.win_n);
	const size_t cap = std::max(n, w.cap);
	w.cap = 0;
	MB_CUDA(h, cudaMalloc(&w.st, cap * sizeof(MbFeState)));
	MB_CUDA(h, cudaMalloc(&w.bbi, cap * bufmax * sizeof(double2)));
	MB_CUDA(h, cudaMalloc(&w.win, cap * (size_t)winmax * sizeof(double2)));
	MB_CUDA(h, cudaMalloc(&w.energy_part, cap * ((bufmax + 255) / 256) * sizeof(double)));  // >= one partial per 1024-sample tile
	MB_CUDA(h, cudaMalloc(&w.vals, cap * kFeVals * sizeof(double)));
	MB_CUDA(h, cudaMalloc(&w.flags, cap * kFeVals));
	MB_CUDA(h, cudaMalloc(&w.pref_ts, cap * 3 * ((size_t)bufmax / 4 + 1) * sizeof(double)));
	MB_CUDA(h, cudaMalloc(&w.tile_base, cap * ((size_t)bufmax / 1024 + 3) * 3 * sizeof(double)));
	MB_CUDA(h, cudaMalloc(&w.pref_win, cap * 3 * ((size_t)winmax + 1) * sizeof(double)));
	MB_CUDA(h, cudaMalloc(&w.frames, cap * (size_t)symbmax * MB_NOFDM * sizeof(float2)));
	MB_CUDA(h, cudaMalloc(&w.llr, cap * MB_HANDOFF_STRIDE * sizeof(float)));
	MB_CUDA(h, cudaMalloc(&w.tail_stats, cap * sizeof(MbRxStats)));
	MB_CUDA(h, cudaMalloc(&w.tail_payload, cap * 256));
	MB_CUDA(h, cudaMalloc(&w.payload, cap * 256));
	MB_CUDA(h, cudaMalloc(&w.stats, cap * sizeof(MbReceiveStats)));
	if (want_dbg) MB_CUDA(h, cudaMalloc(&w.dbg_bb, cap * (size_t)(symbmax + 4) * MB_NOFDM * sizeof(double2)));
	(void)bb_n;
	w.want_dbg = want_dbg;
	w.cap = cap;
	w.buf = bufmax;
	w.symb = symbmax;
	w.win_n = winmax;
	return MERCURY_B200_OK;
}

// One chunk of captures already on the device: d_x [n][buf], d_stats [n] in/out, d_payload [n][frame_bytes].
int fe_run_mfsk(mercury_b200_t *h, const MbFeArgs &a, MbReceiveStats *d_stats, cudaStream_t s);

int fe_run(mercury_b200_t *h, const void *d_x, int fmt, size_t n, uint8_t *d_payload, MbReceiveStats *d_stats, bool dbg, cudaStream_t s)
{
	const MbMode &m = cur_mode(h);
	FeWork &w = h->fe;
	MbFeArgs a;
	memset(&a, 0, sizeof(a));
	a.x = d_x, a.x_format = fmt, a.n = (int)n;
	a.buffer_Nsymb = mb_fe_buffer_nsymb(m.Nsymb, m.preamble_nSymb);
	a.buf = MB_NOFDM * a.buffer_Nsymb * 4, a.pre = m.preamble_nSymb, a.S = m.Nsymb, a.frame_bytes = m.frame_bytes;
	a.carrier = w.carrier, a.st = w.st, a.bbi = w.bbi, a.energy_part = w.energy_part;
	a.win = w.win, a.win_stride = w.win_n, a.vals = w.vals, a.vals_stride = kFeVals;
	a.coarse_freq_sync = h->coarse_freq_sync && m.M != 200 ? 1 : 0;
	a.flags = w.flags, a.pref_ts = w.pref_ts, a.pref_win = w.pref_win, a.tile_base = w.tile_base;
	a.frames = w.frames, a.dbg_bb = dbg ? w.dbg_bb : nullptr;
	a.tail_stats = w.tail_stats, a.tail_payload = w.tail_payload, a.tail_payload_stride = m.frame_bytes;
	a.payload_out = d_payload, a.counters = w.counters;
	MB_CUDA(h, cudaMemsetAsync(d_payload, 0, n * m.frame_bytes, s));
	if (m.M == 200) return fe_run_mfsk(h, a, d_stats, s);
	if ((a.buf - m.preamble_nSymb * MB_FE_SYM + 99) / 100 > kFeVals) return fail(h, MERCURY_B200_EINVAL, "capture too long for the correlation buffer");
	MB_CUDA(h, mb_fe_begin(a, d_stats, s));
	MB_CUDA(h, mb_fe_p2b_full(a, s));
	h->launches += 4;
	bool run_sc = true;
	// every round each capture either finishes or passes one of: coarse run, <= 2 recovery runs, 3 fine runs + 3 tails, SKIP-H
	// recovery and 3 more trials -- 32 rounds is far above the longest path through receive_byte()
	for (int round = 0; round < (a.coarse_freq_sync ? 48 : 32); round++) {  // (+ 4 rounds per pass through trial 1 with the coarse frequency search)
		MB_CUDA(h, cudaMemsetAsync(w.counters, 0, 4 * sizeof(int32_t), s));
		MB_CUDA(h, mb_fe_step(a, run_sc, s));
		h->launches += run_sc ? 5 : 1;
		MB_CUDA(h, cudaMemcpyAsync(w.h_counters, w.counters, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
		MB_CUDA(h, cudaStreamSynchronize(s));
		h->fe_rounds++;
		const int n_slots = w.h_counters[0], not_done = w.h_counters[1];
		h->fe_exact += (uint64_t)w.h_counters[3];
		run_sc = w.h_counters[2] > 0;
		if (n_slots > 0) {
			MB_CUDA(h, mb_fe_extract(a, s));
			h->launches += 2;
			int rc = launch_demod(h, w.frames, (size_t)n_slots, w.llr, w.tail_stats, nullptr, 0, s);
			if (rc) return rc;
			rc = launch_ldpc(h, w.llr, (size_t)n_slots, w.tail_payload, w.tail_stats, s);
			if (rc) return rc;
		}
		if (not_done == 0) break;
		if (round == 31) return fail(h, MERCURY_B200_ECUDA, "front-end state machine did not terminate");
	}
	MB_CUDA(h, mb_fe_finish(a, d_stats, s));
	h->launches++;
	MB_CUDA(h, cudaStreamSynchronize(s));
	return MERCURY_B200_OK;
}

// The MFSK branch of receive_byte() (ROBUST configurations; telecom_system.cc:646-716, 928-943, 1020-1031, 1081-1198, 1343-1367): full-buffer
// mix + time-sync FIR, tone-preamble sync on the symbol grid, frame-completeness check, one trial at that delay with the data filter and
// no frequency correction, the MFSK tail.  One counter read-back.
int fe_run_mfsk(mercury_b200_t *h, const MbFeArgs &a, MbReceiveStats *d_stats, cudaStream_t s)
{
	FeWork &w = h->fe;
	const MbMode &m = cur_mode(h);
	const int nsymb = a.buf / MB_FE_SYM, nblk = (a.buf + 1023) / 1024;
	if (h->mfsk_cap_buffers < (size_t)a.n || h->mfsk_cap_energies < (size_t)a.n * nsymb * MB_NC) {
		MB_CUDA(h, cudaDeviceSynchronize());
		if (h->d_mfsk_out) cudaFree(h->d_mfsk_out);
		if (h->d_mfsk_energies) cudaFree(h->d_mfsk_energies);
		h->d_mfsk_out = nullptr, h->d_mfsk_energies = nullptr;
		h->mfsk_cap_buffers = h->mfsk_cap_energies = 0;
		MB_CUDA(h, cudaMalloc(&h->d_mfsk_out, (size_t)a.n * sizeof(MbMfskPatternResult)));
		MB_CUDA(h, cudaMalloc(&h->d_mfsk_energies, (size_t)a.n * nsymb * MB_NC * sizeof(double)));
		h->mfsk_cap_buffers = a.n, h->mfsk_cap_energies = (size_t)a.n * nsymb * MB_NC;
	}
	{
		MbFeArgs a2 = a;
		a2.pref_ts = nullptr;  // the Schmidl-Cox prefix sums are not needed in this branch
		MB_CUDA(h, mb_fe_p2b_full(a2, s));
	}
	// per-capture search start = the record's mfsk_search_or_overflow (int32 #7 of every 18-int32 record)
	MB_CUDA(h, mb_launch_mfsk_patterns(a.bbi, 0, (size_t)a.n, a.buf, 0, reinterpret_cast<const int32_t *>(d_stats) + 7, (int)(sizeof(MbReceiveStats) / 4),
					   h->mfsk_tones[h->config - 100], m.preamble_nSymb, h->d_mfsk_energies, h->d_mfsk_out, s));
	MB_CUDA(h, cudaMemsetAsync(w.counters, 0, 4 * sizeof(int32_t), s));
	MB_CUDA(h, mb_launch_mfsk_rx_decide(h->d_mfsk_out, a.energy_part, nblk, a.buf, a.pre, a.S, active_nsymb(h), a.buffer_Nsymb, h->fe_const.fc, a.st, d_stats, a.n, w.counters, s));
	MB_CUDA(h, cudaMemcpyAsync(w.h_counters, w.counters, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
	MB_CUDA(h, cudaStreamSynchronize(s));
	h->launches += 4;
	h->fe_rounds++;
	const int n_slots = w.h_counters[0];
	if (n_slots > 0) {
		MB_CUDA(h, mb_fe_extract_data(a, s));
		h->launches++;
		int rc = launch_demod(h, w.frames, (size_t)n_slots, w.llr, w.tail_stats, nullptr, 0, s);
		if (rc) return rc;
		rc = launch_ldpc(h, w.llr, (size_t)n_slots, w.tail_payload, w.tail_stats, s);
		if (rc) return rc;
		MB_CUDA(h, mb_launch_mfsk_rx_finish(a.st, w.tail_stats, w.tail_payload, m.frame_bytes, a.payload_out, d_stats, a.n, s));
		h->launches++;
	}
	MB_CUDA(h, cudaStreamSynchronize(s));
	return MERCURY_B200_OK;
}

}  // namespace

extern "C" {

// Host-only (no device needed): the front-end's host-derived constants, for pinning them against the oracle on a CPU-only machine.
int mercury_b200_build_frontend_tables_host(double *ts_coef /*[33]*/, double *data_coef /*[33]*/, double *consts /*[8]*/, double *carrier /*[2 * n_carrier]*/,
					    int n_carrier)
{
	MbFeConst k;
	mb_fe_host_const(&k);
	if (ts_coef) memcpy(ts_coef, k.c_ts, sizeof(double) * MB_FE_TAPS);
	if (data_coef) memcpy(data_coef, k.c_data, sizeof(double) * MB_FE_TAPS);
	if (consts) {
		const double v[8] = {k.fs, k.fc, k.amp, k.bandwidth, (double)k.trials_max, (double)k.use_last_time, (double)k.use_last_freq, k.ignore_limit};
		memcpy(consts, v, sizeof(v));
	}
	if (carrier && n_carrier > 0) mb_fe_host_carrier(k, carrier, n_carrier);
	return MERCURY_B200_OK;
}

/* char cl_telecom_system::get_configuration(double SNR) (telecom_system.cc:3036-3106): the gear-shift ladder over the FER < 0.1 thresholds
 * of include/common/common_defines.h:130-147; CONFIG_16 is never chosen (its threshold has no rung), like the reference. */
int mercury_b200_get_configuration(double SNR)
{
	static const double rung[15] = {12.5, 9, 7.5, 6.5, 4, 3, 1.5, 0.5, -0.5, -1.5, -2.5, -3.5, -4.5, -6, -7.5};
	for (int i = 0; i < 15; i++)
		if (SNR > rung[i]) return 15 - i;
	return 0;
}

/* double cl_telecom_system::measure_signal_only(double* data) (telecom_system.cc:1520-1541): mix + FIR_rx_time_sync over the capture, mean power
 * in dBm -- no preamble search, no decoding.  n captures of capture_samples samples each (host). */
int mercury_b200_measure_signal_only_batch(mercury_b200_t *h, const void *passband, int sample_format, size_t n, double *signal_dbm)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!passband || !signal_dbm || fe_sample_bytes(sample_format) == 0) return fail(h, MERCURY_B200_EINVAL, "bad argument");
	const MbMode &m = cur_mode(h);
	const int buf = fe_capture_samples(m);
	const size_t ss = fe_sample_bytes(sample_format);
	const size_t chunk = std::min<size_t>(n, m.M == 200 ? 64 : 512);
	rc = fe_ensure(h, chunk, buf, m, false, chunk * buf * ss);
	if (rc) return rc;
	FeWork &w = h->fe;
	const int nblk = (buf + 1023) / 1024;
	std::vector<double> part(chunk * nblk);
	for (size_t done = 0; done < n; done += chunk) {
		const size_t c = std::min(chunk, n - done);
		MbFeArgs a;
		memset(&a, 0, sizeof(a));
		a.x = w.d_x[0], a.x_format = sample_format, a.n = (int)c, a.buf = buf, a.carrier = w.carrier, a.bbi = w.bbi, a.energy_part = w.energy_part;
		MB_CUDA(h, cudaMemcpyAsync(w.d_x[0], static_cast<const uint8_t *>(passband) + done * buf * ss, c * buf * ss, cudaMemcpyHostToDevice, w.stream));
		MB_CUDA(h, mb_fe_p2b_full(a, w.stream));  // pref_ts == NULL: base-band + energy partials only
		MB_CUDA(h, cudaMemcpyAsync(part.data(), w.energy_part, c * nblk * sizeof(double), cudaMemcpyDeviceToHost, w.stream));
		MB_CUDA(h, cudaStreamSynchronize(w.stream));
		h->launches += 1;
		for (size_t i = 0; i < c; i++) {
			double e = 0;
			for (int k = 0; k < nblk; k++) e += part[i * nblk + k];
			signal_dbm[done + i] = 10.0 * log10((e / buf) / 0.001);  // measure_signal_stregth, ofdm.cc:1523-1539
		}
	}
	return MERCURY_B200_OK;
}

/* g_gui_state.coarse_freq_sync_enabled (gui_state.h:143, read at telecom_system.cc:949): when trial 0 of receive_byte() fails, search fc - 30,
 * fc, fc + 30 Hz with the time-sync filter before trial 1 and keep the winning carrier for the rest of the call. Off by default, like the
 * reference's. */
int mercury_b200_set_coarse_freq_sync(mercury_b200_t *h, int enable)
{
	if (!h) return MERCURY_B200_EINVAL;
	h->coarse_freq_sync = enable != 0;
	return MERCURY_B200_OK;
}

/* void set_mfsk_ctrl_mode(bool) + int get_active_nsymb() (telecom_system.h, .cc:1572-1580): shortened control frames in ROBUST_0 (240 of 320
 * symbols) and ROBUST_1 (175 of 200); affects transmit_byte, receive_byte and the tail entry points.  Returns the active symbol count. */
int mercury_b200_set_mfsk_ctrl_mode(mercury_b200_t *h, int enable)
{
	if (!h || h->blob.empty() || h->config < 0) return MERCURY_B200_ESTATE;
	h->mfsk_ctrl = enable != 0;
	h->mfsk_ctrl = active_nsymb(h) != cur_mode(h).Nsymb;  // only where a shorter control frame exists (:1574)
	return active_nsymb(h);
}

int mercury_b200_get_active_nsymb(const mercury_b200_t *h)
{
	if (!h || h->blob.empty() || h->config < 0) return MERCURY_B200_ESTATE;
	return active_nsymb(h);
}

int mercury_b200_get_capture_samples(const mercury_b200_t *h)
{
	if (!h || h->blob.empty() || h->config < 0) return MERCURY_B200_ESTATE;
	return fe_capture_samples(cur_mode(h));
}

int mercury_b200_receive_byte_batch_device(mercury_b200_t *h, const void *d_x, int fmt, size_t n, void *d_payload, void *d_stats, void *stream)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!d_x || !d_payload || !d_stats || fe_sample_bytes(fmt) == 0) return fail(h, MERCURY_B200_EINVAL, "bad argument");
	const MbMode &m = cur_mode(h);
	const int buf = fe_capture_samples(m);
	const size_t ss = fe_sample_bytes(fmt);
	const size_t dchunk = m.M == 200 ? std::min<size_t>(h->fe_chunk, 128) : h->fe_chunk;  // ROBUST captures are 4-8x longer (11 MB of fp64 base-band each)
	rc = fe_ensure(h, std::min(n, dchunk), buf, m, false, 0);
	if (rc) return rc;
	for (size_t done = 0; done < n; done += dchunk) {
		const size_t c = std::min(dchunk, n - done);
		rc = fe_run(h, static_cast<const uint8_t *>(d_x) + done * buf * ss, fmt, c, static_cast<uint8_t *>(d_payload) + done * m.frame_bytes,
			    static_cast<MbReceiveStats *>(d_stats) + done, false, static_cast<cudaStream_t>(stream));
		if (rc) return rc;
	}
	return MERCURY_B200_OK;
}

int mercury_b200_receive_byte_batch(mercury_b200_t *h, const void *x, int fmt, size_t n, uint8_t *payload, mercury_b200_receive_stats *stats,
				    double *baseband_dbg)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!x || !payload || !stats || fe_sample_bytes(fmt) == 0) return fail(h, MERCURY_B200_EINVAL, "bad argument");
	const MbMode &m = cur_mode(h);
	const int buf = fe_capture_samples(m);
	const size_t ss = fe_sample_bytes(fmt);
	// chunks of host_chunk captures: the H2D copy of chunk i+1 (copy stream, second staging buffer) runs while the kernels of chunk i
	// do -- this path is bound by PCIe (one capture is 0.37-0.95 MB), so hiding the compute behind the copies is what matters
	// chunk schedule: a small first chunk (its copy is the only one nothing hides), then host_chunk captures (scaled up for the
	// narrower sample formats so that a chunk stays around 100-200 MB)
	size_t big = h->fe.host_chunk * (ss <= 2 ? 2 : 1);
	if (m.M == 200) big = 64;  // ROBUST captures are 4-8x longer
	big = std::min(std::min(n, h->fe_chunk), big);
	const size_t first = std::min(big, std::max<size_t>(64, big / 4));
	rc = fe_ensure(h, big, buf, m, baseband_dbg != nullptr, big * buf * ss);
	if (rc) return rc;
	FeWork &w = h->fe;
	const size_t bb_n = (size_t)(m.Nsymb + m.preamble_nSymb) * MB_NOFDM;
	const uint8_t *xb = static_cast<const uint8_t *>(x);
	MB_CUDA(h, cudaMemcpyAsync(w.d_x[0], xb, std::min(first, n) * buf * ss, cudaMemcpyHostToDevice, w.copy_stream));
	MB_CUDA(h, cudaEventRecord(w.copied[0], w.copy_stream));
	int i = 0;
	size_t c = std::min(first, n);
	for (size_t done = 0; done < n; i++) {
		const int cur = i & 1, nxt = cur ^ 1;
		const size_t c2 = std::min(big, n - done - c);  // the chunk after this one
		if (c2 > 0) {  // prefetch it once its staging buffer has been consumed
			if (i >= 1) MB_CUDA(h, cudaStreamWaitEvent(w.copy_stream, w.consumed[nxt], 0));
			MB_CUDA(h, cudaMemcpyAsync(w.d_x[nxt], xb + (done + c) * buf * ss, c2 * buf * ss, cudaMemcpyHostToDevice, w.copy_stream));
			MB_CUDA(h, cudaEventRecord(w.copied[nxt], w.copy_stream));
		}
		MB_CUDA(h, cudaStreamWaitEvent(w.stream, w.copied[cur], 0));
		MB_CUDA(h, cudaMemcpyAsync(w.stats, stats + done, c * sizeof(MbReceiveStats), cudaMemcpyHostToDevice, w.stream));
		if (baseband_dbg) MB_CUDA(h, cudaMemsetAsync(w.dbg_bb, 0, c * bb_n * sizeof(double2), w.stream));
		rc = fe_run(h, w.d_x[cur], fmt, c, w.payload, w.stats, baseband_dbg != nullptr, w.stream);
		if (rc) return rc;
		MB_CUDA(h, cudaEventRecord(w.consumed[cur], w.stream));
		MB_CUDA(h, cudaMemcpyAsync(payload + done * m.frame_bytes, w.payload, c * m.frame_bytes, cudaMemcpyDeviceToHost, w.stream));
		MB_CUDA(h, cudaMemcpyAsync(stats + done, w.stats, c * sizeof(MbReceiveStats), cudaMemcpyDeviceToHost, w.stream));
		if (baseband_dbg)
			MB_CUDA(h, cudaMemcpyAsync(baseband_dbg + done * bb_n * 2, w.dbg_bb, c * bb_n * sizeof(double2), cudaMemcpyDeviceToHost, w.stream));
		MB_CUDA(h, cudaStreamSynchronize(w.stream));
		done += c;
		c = c2;
	}
	return MERCURY_B200_OK;
}

int mercury_b200_receive_byte(mercury_b200_t *h, const double *passband, int *out, mercury_b200_receive_stats *stats)
{
	if (!h || !passband || !out || !stats) return MERCURY_B200_EINVAL;
	uint8_t pl[256];
	int rc = mercury_b200_receive_byte_batch(h, passband, MERCURY_B200_SAMPLES_F64, 1, pl, stats, nullptr);
	if (rc) return rc;
	const int fb = cur_mode(h).frame_bytes;
	for (int i = 0; i < fb; i++) out[i] = pl[i];  // one int per byte (telecom_system.cc:1329-1332)
	return MERCURY_B200_OK;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * TX chain (mb_tx.cu): transmit_byte(SINGLE_MESSAGE), batched.
 * ------------------------------------------------------------------------------------------------------------------- */
}  // extern "C"

namespace {

void tx_free(TxWork &w)
{
	for (int i = 0; i < MB_NMODES + 3; i++) {
		if (w.mode_dev[i]) cudaFree(w.mode_dev[i]);
		if (w.tables[i]) cudaFree(w.tables[i]);
	}
	void *ptrs[] = {w.payload, w.dbg_cw, w.start, w.bb, w.pb, w.p1, w.power_part, w.out, w.stream_buf};
	for (void *p : ptrs)
		if (p) cudaFree(p);
	if (w.stream) cudaStreamDestroy(w.stream);
	if (w.copy_stream) cudaStreamDestroy(w.copy_stream);
	for (int i = 0; i < 2; i++) {
		if (w.ev_done[i]) cudaEventDestroy(w.ev_done[i]);
		if (w.ev_copied[i]) cudaEventDestroy(w.ev_copied[i]);
	}
	w = TxWork();
}

int tx_total(const MbMode &m) { return (m.Nsymb + m.preamble_nSymb) * MB_FE_SYM; }
int tx_slot(const mercury_b200_t *h) { return is_mfsk_config(h->config) ? MB_NMODES + h->config - 100 : h->config; }

int tx_ensure_mode(mercury_b200_t *h)
{
	TxWork &w = h->tx;
	if (!w.init_done) {
		if (!h->fe_ready) {
			mb_fe_host_const(&h->fe_const);
			cudaError_t e = mb_fe_init(h->fe_const);
			if (e != cudaSuccess) return cuda_fail(h, e, "front-end constants");
			h->fe_ready = true;
		}
		MB_CUDA(h, mb_tx_init(h->fe_const));
		MB_CUDA(h, cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
		w.init_done = true;
	}
	const int c = tx_slot(h);
	if (w.built[c]) return MERCURY_B200_OK;
	std::vector<uint8_t> bytes;
	const std::string e = is_mfsk_config(h->config)
				      ? mb_tx_build_mfsk(h->blob, cur_mode(h), h->mfsk_tones[h->config - 100], h->fe_const, &w.mode_host[c], &bytes)
				      : mb_tx_build(h->blob, c, h->fe_const, &w.mode_host[c], &bytes);
	if (!e.empty()) return fail(h, MERCURY_B200_EINVAL, e);
	MB_CUDA(h, cudaMalloc(&w.tables[c], bytes.size()));
	MB_CUDA(h, cudaMemcpy(w.tables[c], bytes.data(), bytes.size(), cudaMemcpyHostToDevice));
	MB_CUDA(h, cudaMalloc(&w.mode_dev[c], sizeof(MbTxMode)));
	MB_CUDA(h, cudaMemcpy(w.mode_dev[c], &w.mode_host[c], sizeof(MbTxMode), cudaMemcpyHostToDevice));
	w.built[c] = true;
	return MERCURY_B200_OK;
}

int tx_ensure_work(mercury_b200_t *h, size_t n, int total, bool want_cw)
{
	TxWork &w = h->tx;
	if (w.cap >= n && w.cap_total >= (size_t)total && (!want_cw || w.dbg_cw)) return MERCURY_B200_OK;
	MB_CUDA(h, cudaDeviceSynchronize());
	void **ptrs[] = {(void **)&w.payload, (void **)&w.dbg_cw, (void **)&w.start, (void **)&w.bb, (void **)&w.pb, (void **)&w.p1, (void **)&w.power_part};
	for (void **p : ptrs) {
		if (*p) cudaFree(*p);
		*p = nullptr;
	}
	const size_t cap = std::max(n, w.cap), tot = std::max((size_t)total, w.cap_total);
	w.cap = w.cap_total = 0;
	MB_CUDA(h, cudaMalloc(&w.payload, cap * 256));
	MB_CUDA(h, cudaMalloc(&w.start, cap * sizeof(unsigned long long)));
	MB_CUDA(h, cudaMalloc(&w.bb, cap * (tot / 4) * sizeof(double2)));
	MB_CUDA(h, cudaMalloc(&w.pb, cap * tot * sizeof(double)));
	MB_CUDA(h, cudaMalloc(&w.p1, cap * tot * sizeof(double)));
	MB_CUDA(h, cudaMalloc(&w.power_part, cap * ((tot + 255) / 256) * 2 * sizeof(double)));
	if (want_cw) MB_CUDA(h, cudaMalloc(&w.dbg_cw, cap * MB_N));
	w.cap = cap, w.cap_total = tot;
	return MERCURY_B200_OK;
}

// the host batch path's two output slots (the device entry points write straight into the caller's buffer)
int tx_ensure_out(mercury_b200_t *h, size_t n, int total)
{
	TxWork &w = h->tx;
	const size_t need = n * (size_t)total * sizeof(double);
	if (w.out_slot_bytes >= need && w.copy_stream) return MERCURY_B200_OK;
	MB_CUDA(h, cudaDeviceSynchronize());
	if (w.out) cudaFree(w.out);
	w.out = nullptr, w.out_slot_bytes = 0;
	MB_CUDA(h, cudaMalloc(&w.out, 2 * need));
	w.out_slot_bytes = need;
	if (!w.copy_stream) {
		MB_CUDA(h, cudaStreamCreateWithFlags(&w.copy_stream, cudaStreamNonBlocking));
		for (int i = 0; i < 2; i++) {
			MB_CUDA(h, cudaEventCreateWithFlags(&w.ev_done[i], cudaEventDisableTiming));
			MB_CUDA(h, cudaEventCreateWithFlags(&w.ev_copied[i], cudaEventDisableTiming));
		}
	}
	return MERCURY_B200_OK;
}

// d_payload [n][frame_bytes], d_start [n] or NULL, d_out [n][total] (double, or float when out_f32), d_cw optional [n][1600]
int tx_run(mercury_b200_t *h, const uint8_t *d_payload, const unsigned long long *d_start, size_t n, void *d_out, bool out_f32, uint8_t *d_cw, cudaStream_t s,
	   bool no_filter = false)
{
	TxWork &w = h->tx;
	const int c = tx_slot(h);
	MbTxArgs a;
	memset(&a, 0, sizeof(a));
	a.tone = is_mfsk_config(h->config) ? &h->mfsk_tones[h->config - 100] : nullptr;
	a.S_active = active_nsymb(h);
	a.tm = w.mode_dev[c], a.tm_host = &w.mode_host[c], a.tables = w.tables[c];
	a.payload = d_payload, a.start_sample = d_start, a.n = (int)n, a.out_f32 = out_f32, a.no_filter = no_filter;
	a.bb = w.bb, a.pb = w.pb, a.p1 = w.p1, a.power_part = w.power_part, a.out = d_out, a.dbg_cw = d_cw;
	MB_CUDA(h, mb_tx_launch(a, s));
	h->launches += 4;
	return MERCURY_B200_OK;
}

}  // namespace

extern "C" {

// Host-only (no device needed): the TX tables of one configuration, for pinning them against the oracle on a CPU-only machine.
int mercury_b200_build_tx_tables_host(const char *ldpc_table_path, int config, double *pre_eq /*[100]*/, double *preamble /*[<=400]*/, double *tx1 /*[97]*/,
				      double *tx2 /*[97]*/)
{
	if (!ldpc_table_path || config < 0 || config >= MB_NMODES) return MERCURY_B200_EINVAL;
	std::vector<uint8_t> blob;
	if (!mb_build_blob(ldpc_table_path, blob).empty()) return MERCURY_B200_EIO;
	MbFeConst fe;
	mb_fe_host_const(&fe);
	MbTxMode tm;
	std::vector<uint8_t> bytes;
	if (!mb_tx_build(blob, config, fe, &tm, &bytes).empty()) return MERCURY_B200_EINVAL;
	if (pre_eq) memcpy(pre_eq, bytes.data() + tm.off_pre_eq, sizeof(double) * 2 * MB_NC);
	if (preamble) memcpy(preamble, bytes.data() + tm.off_preamble, sizeof(double) * 2 * MB_NC * tm.pre);
	if (tx1) memcpy(tx1, bytes.data() + tm.off_c1, sizeof(double) * 97);
	if (tx2) memcpy(tx2, bytes.data() + tm.off_c2, sizeof(double) * 97);
	return MERCURY_B200_OK;
}

int mercury_b200_get_total_frame_size(const mercury_b200_t *h)
{
	if (!h || h->blob.empty() || h->config < 0) return MERCURY_B200_ESTATE;
	return tx_total(cur_mode(h));
}

int mercury_b200_transmit_byte_batch_device(mercury_b200_t *h, const void *d_payload, const void *d_start_sample, size_t n, void *d_passband, int out_format,
					    void *stream)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!d_payload || !d_passband || (out_format != MERCURY_B200_SAMPLES_F64 && out_format != MERCURY_B200_SAMPLES_F32))
		return fail(h, MERCURY_B200_EINVAL, "bad argument");
	rc = tx_ensure_mode(h);
	if (rc) return rc;
	const MbMode &m = cur_mode(h);
	const int total = tx_total(m);
	const size_t chunk = std::min<size_t>(n, m.M == 200 ? 512 : 8192), ob = out_format == MERCURY_B200_SAMPLES_F32 ? 4 : 8;
	rc = tx_ensure_work(h, chunk, total, false);
	if (rc) return rc;
	for (size_t done = 0; done < n; done += chunk) {
		const size_t c = std::min(chunk, n - done);
		rc = tx_run(h, static_cast<const uint8_t *>(d_payload) + done * m.frame_bytes,
			    d_start_sample ? static_cast<const unsigned long long *>(d_start_sample) + done : nullptr, c,
			    static_cast<uint8_t *>(d_passband) + done * total * ob, out_format == MERCURY_B200_SAMPLES_F32, nullptr, static_cast<cudaStream_t>(stream));
		if (rc) return rc;
	}
	return MERCURY_B200_OK;
}

int mercury_b200_transmit_byte_batch_ex(mercury_b200_t *h, const uint8_t *payload, const uint64_t *start_sample, size_t n, void *passband, int out_format,
					int message_location, uint8_t *codeword_dbg)
{
	if (message_location != MERCURY_B200_SINGLE_MESSAGE && message_location != MERCURY_B200_NO_FILTER_MESSAGE)
		return fail(h, MERCURY_B200_EINVAL, "message_location must be SINGLE_MESSAGE (3) or NO_FILTER_MESSAGE (4)");
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!payload || !passband || (out_format != MERCURY_B200_SAMPLES_F64 && out_format != MERCURY_B200_SAMPLES_F32))
		return fail(h, MERCURY_B200_EINVAL, "bad argument");
	rc = tx_ensure_mode(h);
	if (rc) return rc;
	const MbMode &m = cur_mode(h);
	const int total = tx_total(m);
	// chunks of 512 frames (15-60 MB of samples each): the output is ~400x the input, so the call is bound by the D2H copy; two output slots
	// let the copy of chunk i (copy stream) run under the kernels of chunk i + 1.  With pageable host memory the copy still serialises on the
	// host side; with pinned memory (mercury_b200_host_alloc) the link stays busy.
	const size_t chunk = std::min<size_t>(n, m.M == 200 ? 256 : 512), ob = out_format == MERCURY_B200_SAMPLES_F32 ? 4 : 8;
	rc = tx_ensure_work(h, chunk, total, codeword_dbg != nullptr);
	if (rc) return rc;
	rc = tx_ensure_out(h, chunk, total);
	if (rc) return rc;
	TxWork &w = h->tx;
	const size_t slot_bytes = w.out_slot_bytes;
	size_t i = 0;
	for (size_t done = 0; done < n; done += chunk, i++) {
		const size_t c = std::min(chunk, n - done);
		const int slot = (int)(i & 1);
		uint8_t *d_out = static_cast<uint8_t *>(w.out) + (size_t)slot * slot_bytes;
		if (i >= 2) MB_CUDA(h, cudaStreamWaitEvent(w.stream, w.ev_copied[slot], 0));  // the copy of chunk i - 2 has left this slot
		MB_CUDA(h, cudaMemcpyAsync(w.payload, payload + done * m.frame_bytes, c * m.frame_bytes, cudaMemcpyHostToDevice, w.stream));
		if (start_sample) MB_CUDA(h, cudaMemcpyAsync(w.start, start_sample + done, c * sizeof(uint64_t), cudaMemcpyHostToDevice, w.stream));
		rc = tx_run(h, w.payload, start_sample ? w.start : nullptr, c, d_out, out_format == MERCURY_B200_SAMPLES_F32, codeword_dbg ? w.dbg_cw : nullptr, w.stream,
			    message_location == MERCURY_B200_NO_FILTER_MESSAGE);
		if (rc) return rc;
		if (codeword_dbg) MB_CUDA(h, cudaMemcpyAsync(codeword_dbg + done * MB_N, w.dbg_cw, c * MB_N, cudaMemcpyDeviceToHost, w.stream));
		MB_CUDA(h, cudaEventRecord(w.ev_done[slot], w.stream));
		MB_CUDA(h, cudaStreamWaitEvent(w.copy_stream, w.ev_done[slot], 0));
		MB_CUDA(h, cudaMemcpyAsync(static_cast<uint8_t *>(passband) + done * total * ob, d_out, c * total * ob, cudaMemcpyDeviceToHost, w.copy_stream));
		MB_CUDA(h, cudaEventRecord(w.ev_copied[slot], w.copy_stream));
	}
	MB_CUDA(h, cudaStreamSynchronize(w.stream));
	MB_CUDA(h, cudaStreamSynchronize(w.copy_stream));
	return MERCURY_B200_OK;
}

int mercury_b200_transmit_byte_batch(mercury_b200_t *h, const uint8_t *payload, const uint64_t *start_sample, size_t n, void *passband, int out_format,
				     uint8_t *codeword_dbg)
{
	return mercury_b200_transmit_byte_batch_ex(h, payload, start_sample, n, passband, out_format, MERCURY_B200_SINGLE_MESSAGE, codeword_dbg);
}

/* transmit_byte with the streaming message locations FIRST_MESSAGE (0) / MIDDLE_MESSAGE (1) / FLUSH_MESSAGE (2) (telecom_system.cc:559-594, the
 * TX_TEST path :2033-2038): a three-frame buffer of clipped frames lives on the device between calls, both FIRs run over the two frames centred
 * on the middle one, which is what comes out (one frame of latency).  SINGLE (3) and NO_FILTER (4) are accepted too.  One stream per handle;
 * mercury_b200_load_configuration and mercury_b200_reset_tx_stream clear the buffer. */
int mercury_b200_transmit_byte_loc(mercury_b200_t *h, const int *data, int nBytes, double *out, uint64_t *passband_start_sample, int message_location)
{
	if (!h || !data || !out) return MERCURY_B200_EINVAL;
	int rc = check_ready(h);
	if (rc) return rc;
	const MbMode &m = cur_mode(h);
	if (nBytes < 0 || nBytes > m.frame_bytes) return fail(h, MERCURY_B200_EINVAL, "message too long.. not sent.");
	if (message_location < 0 || message_location > 4) return fail(h, MERCURY_B200_EINVAL, "unknown message_location");
	uint8_t pl[256] = {0};
	for (int i = 0; i < nBytes; i++) pl[i] = (uint8_t)data[i];
	uint64_t start = passband_start_sample ? *passband_start_sample : (uint64_t)(m.M == 200 ? 0 : MB_FE_SYM);
	const int T = tx_total(m);
	if (message_location >= 3) {
		rc = mercury_b200_transmit_byte_batch_ex(h, pl, &start, 1, out, MERCURY_B200_SAMPLES_F64, message_location, nullptr);
	} else {
		rc = tx_ensure_mode(h);
		if (rc) return rc;
		rc = tx_ensure_work(h, 1, T, false);
		if (rc) return rc;
		TxWork &w = h->tx;
		if (!w.stream_buf || w.stream_total != T || w.stream_slot != tx_slot(h)) {
			if (w.stream_buf) cudaFree(w.stream_buf);
			w.stream_buf = nullptr;
			MB_CUDA(h, cudaMalloc(&w.stream_buf, (size_t)7 * T * sizeof(double)));
			MB_CUDA(h, cudaMemsetAsync(w.stream_buf, 0, (size_t)7 * T * sizeof(double), w.stream));
			w.stream_total = T, w.stream_slot = tx_slot(h);
		}
		double *buf = w.stream_buf, *f1 = buf + 3 * (size_t)T, *f2 = f1 + 2 * (size_t)T;
		MB_CUDA(h, cudaMemcpyAsync(w.payload, pl, m.frame_bytes, cudaMemcpyHostToDevice, w.stream));
		MB_CUDA(h, cudaMemcpyAsync(w.start, &start, sizeof(uint64_t), cudaMemcpyHostToDevice, w.stream));
		rc = tx_run(h, w.payload, w.start, 1, buf + 2 * (size_t)T, false, nullptr, w.stream, /*no_filter=*/true);  // the clipped frame -> third slot
		if (rc) return rc;
		if (message_location == 0)  // FIRST_MESSAGE: the frame also fills the middle slot (:559-566)
			MB_CUDA(h, cudaMemcpyAsync(buf + T, buf + 2 * (size_t)T, (size_t)T * sizeof(double), cudaMemcpyDeviceToDevice, w.stream));
		MB_CUDA(h, mb_tx_fir_apply(w.tables[tx_slot(h)], w.mode_host[tx_slot(h)], buf + T / 2, 2 * T, f1, f2, w.stream));
		h->launches += 2;
		MB_CUDA(h, cudaMemcpyAsync(out, f2 + T / 2, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, w.stream));
		// shift_left(buffer, 3T, T): slot 0 <- slot 1, slot 1 <- slot 2 (slot 2 keeps its content, like the reference)
		MB_CUDA(h, cudaMemcpyAsync(buf, buf + T, (size_t)T * sizeof(double), cudaMemcpyDeviceToDevice, w.stream));
		MB_CUDA(h, cudaMemcpyAsync(buf + T, buf + 2 * (size_t)T, (size_t)T * sizeof(double), cudaMemcpyDeviceToDevice, w.stream));
		MB_CUDA(h, cudaStreamSynchronize(w.stream));
	}
	if (rc) return rc;
	if (passband_start_sample) *passband_start_sample = start + (uint64_t)(active_nsymb(h) + m.preamble_nSymb) * MB_FE_SYM;
	return MERCURY_B200_OK;
}

int mercury_b200_reset_tx_stream(mercury_b200_t *h)
{
	if (!h) return MERCURY_B200_EINVAL;
	TxWork &w = h->tx;
	if (w.stream_buf) {
		cudaSetDevice(h->device);
		cudaFree(w.stream_buf);
		w.stream_buf = nullptr;
	}
	w.stream_total = 0, w.stream_slot = -1;
	return MERCURY_B200_OK;
}

/* ofdm.FIR_tx1.apply + ofdm.FIR_tx2.apply over one host buffer (the ARQ layer's batch filtering, arq_common.cc:2243-2246). */
int mercury_b200_fir_tx_apply(mercury_b200_t *h, const double *in, size_t n, double *out)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (!in || !out || n == 0 || n > (size_t)1 << 30) return fail(h, MERCURY_B200_EINVAL, "bad argument");
	rc = tx_ensure_mode(h);
	if (rc) return rc;
	TxWork &w = h->tx;
	double *d = nullptr;
	MB_CUDA(h, cudaMalloc(&d, 3 * n * sizeof(double)));
	cudaError_t e = cudaMemcpyAsync(d, in, n * sizeof(double), cudaMemcpyHostToDevice, w.stream);
	if (e == cudaSuccess) e = mb_tx_fir_apply(w.tables[tx_slot(h)], w.mode_host[tx_slot(h)], d, (int)n, d + n, d + 2 * n, w.stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost, w.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(w.stream);
	cudaFree(d);
	if (e != cudaSuccess) return cuda_fail(h, e, "fir_tx_apply");
	h->launches += 2;
	return MERCURY_B200_OK;
}

int mercury_b200_transmit_byte(mercury_b200_t *h, const int *data, int nBytes, double *out, uint64_t *passband_start_sample)
{
	if (!h || !data || !out) return MERCURY_B200_EINVAL;
	int rc = check_ready(h);
	if (rc) return rc;
	const MbMode &m = cur_mode(h);
	if (nBytes < 0 || nBytes > m.frame_bytes) return fail(h, MERCURY_B200_EINVAL, "message too long.. not sent.");  // telecom_system.cc:348-352
	uint8_t pl[256] = {0};
	for (int i = 0; i < nBytes; i++) pl[i] = (uint8_t)data[i];
	uint64_t start = passband_start_sample ? *passband_start_sample : (uint64_t)(m.M == 200 ? 0 : MB_FE_SYM);
	rc = mercury_b200_transmit_byte_batch(h, pl, &start, 1, out, MERCURY_B200_SAMPLES_F64, nullptr);
	if (rc) return rc;
	// the carrier counter advances by what baseband_to_passband was given: (preamble + ACTIVE symbols) * Nofdm * 4 -- fewer than the frame
	// in ROBUST control-frame mode (telecom_system.cc:531-532, ofdm.cc:2313), like mercury_b200_transmit_byte_loc
	if (passband_start_sample) *passband_start_sample = start + (uint64_t)(active_nsymb(h) + m.preamble_nSymb) * MB_FE_SYM;
	return MERCURY_B200_OK;
}

/* MFSK tone-pattern detectors (mb_mfsk.cu): time_sync_mfsk + detect_ack_pattern (ACK and BREAK) over n_buffers base-band buffers. */
int mercury_b200_mfsk_patterns_batch(mercury_b200_t *h, const void *bbi, int complex_format, size_t n_buffers, int n_samples, int search_start_symb,
				     mercury_b200_mfsk_pattern_result *out)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n_buffers == 0) return MERCURY_B200_OK;
	if (!bbi || !out || n_samples < MB_FE_SYM || (complex_format != MERCURY_B200_SAMPLES_F64 && complex_format != MERCURY_B200_SAMPLES_F32))
		return fail(h, MERCURY_B200_EINVAL, "bad argument");
	if (!is_mfsk_config(h->config)) return fail(h, MERCURY_B200_EINVAL, "the tone-pattern detectors need a ROBUST configuration (100..102)");
	const size_t es = complex_format == MERCURY_B200_SAMPLES_F32 ? 8 : 16;
	const int nsymb = n_samples / MB_FE_SYM;
	const size_t chunk = std::min<size_t>(n_buffers, std::max<size_t>(1, (512u << 20) / ((size_t)n_samples * es)));
	if (h->mfsk_cap_bytes < chunk * n_samples * es || h->mfsk_cap_buffers < chunk || h->mfsk_cap_energies < chunk * nsymb * MB_NC) {
		MB_CUDA(h, cudaDeviceSynchronize());
		if (h->d_mfsk_in) cudaFree(h->d_mfsk_in);
		if (h->d_mfsk_out) cudaFree(h->d_mfsk_out);
		if (h->d_mfsk_energies) cudaFree(h->d_mfsk_energies);
		h->d_mfsk_in = nullptr, h->d_mfsk_out = nullptr, h->d_mfsk_energies = nullptr;
		h->mfsk_cap_bytes = h->mfsk_cap_buffers = h->mfsk_cap_energies = 0;
		MB_CUDA(h, cudaMalloc(&h->d_mfsk_in, chunk * n_samples * es));
		MB_CUDA(h, cudaMalloc(&h->d_mfsk_out, chunk * sizeof(MbMfskPatternResult)));
		MB_CUDA(h, cudaMalloc(&h->d_mfsk_energies, chunk * nsymb * MB_NC * sizeof(double)));
		h->mfsk_cap_bytes = chunk * n_samples * es, h->mfsk_cap_buffers = chunk, h->mfsk_cap_energies = chunk * nsymb * MB_NC;
	}
	const MbMode &m = cur_mode(h);
	for (size_t done = 0; done < n_buffers; done += chunk) {
		const size_t c = std::min(chunk, n_buffers - done);
		MB_CUDA(h, cudaMemcpy(h->d_mfsk_in, static_cast<const uint8_t *>(bbi) + done * n_samples * es, c * n_samples * es, cudaMemcpyHostToDevice));
		MB_CUDA(h, mb_launch_mfsk_patterns(h->d_mfsk_in, complex_format == MERCURY_B200_SAMPLES_F32, c, n_samples, search_start_symb, nullptr, 0,
						   h->mfsk_tones[h->config - 100], m.preamble_nSymb, h->d_mfsk_energies, h->d_mfsk_out, nullptr));
		h->launches += 2;
		MB_CUDA(h, cudaMemcpy(out + done, h->d_mfsk_out, c * sizeof(MbMfskPatternResult), cudaMemcpyDeviceToHost));
	}
	return MERCURY_B200_OK;
}

/* The ARQ-facing tone-pattern calls (telecom_system.h:122-130): any configuration, dedicated 16-MFSK x 1 plan (telecom_system.cc:3003-3008). */
}  // extern "C"

namespace {
MbMfsk ack_mfsk_plan()
{  // cl_mfsk::init(16, 50, 1): mfsk.cc:49-160
	MbMfsk t;
	memset(&t, 0, sizeof(t));
	static const int pre16[4] = {2, 10, 6, 14}, ack16[8] = {4, 7, 5, 12, 13, 1, 9, 15}, brk16[8] = {6, 14, 2, 3, 10, 8, 11, 15};
	t.M = 16, t.nBits = 4, t.nStreams = 1, t.tone_hop_step = 7, t.stream_offsets[0] = (MB_NC - 16) / 2;
	for (int i = 0; i < 4; i++) t.preamble_tones[i] = pre16[i];
	for (int i = 0; i < 8; i++) t.ack_tones[i] = ack16[i], t.break_tones[i] = brk16[i];
	return t;
}
}  // namespace

extern "C" {

int mercury_b200_generate_pattern_passband(mercury_b200_t *h, int use_break_tones, double *out, uint64_t *passband_start_sample)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (!out) return fail(h, MERCURY_B200_EINVAL, "null buffer");
	rc = tx_ensure_mode(h);  // (front-end constants, TX stream)
	if (rc) return rc;
	const int total = 16 * MB_FE_SYM;
	double *d = nullptr;
	MB_CUDA(h, cudaMalloc(&d, (size_t)(16 * MB_NOFDM * 2 + 2 * total + 2 * ((total + 255) / 256)) * sizeof(double)));
	double2 *d_bb = reinterpret_cast<double2 *>(d);
	double *d_pb = d + 16 * MB_NOFDM * 2, *d_out = d_pb + total, *d_pp = d_out + total;
	const uint64_t start = passband_start_sample ? *passband_start_sample : 0;
	cudaError_t e = mb_tx_pattern(ack_mfsk_plan(), use_break_tones, h->fe_const.fc, h->fe_const.Ts, h->fe_const.amp, start, d_bb, d_pb, d_pp, d_out, h->tx.stream);
	if (e == cudaSuccess) e = cudaMemcpy(out, d_out, total * sizeof(double), cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e != cudaSuccess) return cuda_fail(h, e, "generate_pattern_passband");
	if (passband_start_sample) *passband_start_sample = start + total;
	h->launches += 3;
	return total;
}

int mercury_b200_detect_patterns_from_passband_batch(mercury_b200_t *h, const void *passband, int sample_format, size_t n_buffers, int n_samples,
						     mercury_b200_mfsk_pattern_result *out)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n_buffers == 0) return MERCURY_B200_OK;
	if (!passband || !out || n_samples < 16 * MB_FE_SYM || fe_sample_bytes(sample_format) == 0) return fail(h, MERCURY_B200_EINVAL, "bad argument");
	const MbMode &m = cur_mode(h);
	const size_t ss = fe_sample_bytes(sample_format);
	const size_t chunk = std::min<size_t>(n_buffers, std::max<size_t>(1, (1u << 30) / ((size_t)n_samples * 16)));
	rc = fe_ensure(h, chunk, n_samples, m, false, chunk * n_samples * ss);
	if (rc) return rc;
	FeWork &w = h->fe;
	const int nsymb = n_samples / MB_FE_SYM;
	if (h->mfsk_cap_buffers < chunk || h->mfsk_cap_energies < chunk * nsymb * MB_NC) {
		MB_CUDA(h, cudaDeviceSynchronize());
		if (h->d_mfsk_out) cudaFree(h->d_mfsk_out);
		if (h->d_mfsk_energies) cudaFree(h->d_mfsk_energies);
		h->d_mfsk_out = nullptr, h->d_mfsk_energies = nullptr;
		h->mfsk_cap_buffers = h->mfsk_cap_energies = 0;
		MB_CUDA(h, cudaMalloc(&h->d_mfsk_out, chunk * sizeof(MbMfskPatternResult)));
		MB_CUDA(h, cudaMalloc(&h->d_mfsk_energies, chunk * nsymb * MB_NC * sizeof(double)));
		h->mfsk_cap_buffers = chunk, h->mfsk_cap_energies = chunk * nsymb * MB_NC;
	}
	const MbMfsk plan = ack_mfsk_plan();
	for (size_t done = 0; done < n_buffers; done += chunk) {
		const size_t c = std::min(chunk, n_buffers - done);
		MbFeArgs a;
		memset(&a, 0, sizeof(a));
		a.x = w.d_x[0], a.x_format = sample_format, a.n = (int)c, a.buf = n_samples, a.carrier = w.carrier, a.bbi = w.bbi, a.energy_part = w.energy_part;
		MB_CUDA(h, cudaMemcpyAsync(w.d_x[0], static_cast<const uint8_t *>(passband) + done * n_samples * ss, c * n_samples * ss, cudaMemcpyHostToDevice, w.stream));
		MB_CUDA(h, mb_fe_p2b_data(a, w.stream));
		MB_CUDA(h, mb_launch_mfsk_patterns(w.bbi, 0, c, n_samples, 0, nullptr, 0, plan, 4, h->d_mfsk_energies, h->d_mfsk_out, w.stream));
		MB_CUDA(h, cudaMemcpyAsync(out + done, h->d_mfsk_out, c * sizeof(MbMfskPatternResult), cudaMemcpyDeviceToHost, w.stream));
		MB_CUDA(h, cudaStreamSynchronize(w.stream));
		h->launches += 3;
	}
	return MERCURY_B200_OK;
}

void *mercury_b200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

void mercury_b200_host_free(void *p)
{
	if (p) cudaFreeHost(p);
}

void *mercury_b200_device_alloc(mercury_b200_t *h, size_t bytes)
{
	if (!h || cudaSetDevice(h->device) != cudaSuccess) return nullptr;
	void *p = nullptr;
	if (cudaMalloc(&p, bytes) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

void mercury_b200_device_free(mercury_b200_t *h, void *p)
{
	if (h && p) {
		cudaSetDevice(h->device);
		cudaFree(p);
	}
}

int mercury_b200_memcpy_h2d(mercury_b200_t *h, void *dst, const void *src, size_t bytes)
{
	if (!h) return MERCURY_B200_EINVAL;
	MB_CUDA(h, cudaSetDevice(h->device));
	MB_CUDA(h, cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
	return MERCURY_B200_OK;
}

int mercury_b200_memcpy_d2h(mercury_b200_t *h, void *dst, const void *src, size_t bytes)
{
	if (!h) return MERCURY_B200_EINVAL;
	MB_CUDA(h, cudaSetDevice(h->device));
	MB_CUDA(h, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
	return MERCURY_B200_OK;
}

int mercury_b200_synchronize(mercury_b200_t *h)
{
	if (!h) return MERCURY_B200_EINVAL;
	MB_CUDA(h, cudaSetDevice(h->device));
	MB_CUDA(h, cudaDeviceSynchronize());
	return MERCURY_B200_OK;
}

uint64_t mercury_b200_kernel_launches(const mercury_b200_t *h) { return h ? h->launches : 0; }

int mercury_b200_synth_frames(const char *path, int config, size_t n_frames, uint64_t seed, double esn0_db, const uint8_t *payload_in,
			      float *baseband_out, uint8_t *payload_out, int n_threads)
{
	if (!path || !baseband_out) return MERCURY_B200_EINVAL;
	static std::vector<uint8_t> blob;  // built once per process; the tables do not depend on config
	static std::string blob_path;
	if (blob.empty() || blob_path != path) {
		std::vector<uint8_t> b;
		std::string e = mb_build_blob(path, b);
		if (!e.empty()) return MERCURY_B200_EIO;
		blob.swap(b);
		blob_path = path;
	}
	std::string e = mb_synth_frames(blob, config, n_frames, seed, esn0_db, payload_in, baseband_out, payload_out, n_threads);
	return e.empty() ? MERCURY_B200_OK : MERCURY_B200_EINVAL;
}

}  // extern "C"
