// mb_api.cu -- the C ABI (include/mercury_b200.h) over the two kernels: handle, resident tables, O(1) mode
// switch, device-resident batch entry points and the pipelined host-buffer batch path.
//
// Mirrors the reference's cl_telecom_system surface for the RX tail (see the header for file:line anchors).
// There is deliberately no CPU implementation behind any compute entry point.
#include <cuda_runtime.h>

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

struct mercury_b200 {
	int device = -1;
	std::vector<uint8_t> blob;
	MbBlobHeader hdr;
	uint8_t *d_blob = nullptr;
	int config = -1, ldpc_iters = 50, decoder = MERCURY_B200_DECODER_SPA;
	int cheap_test_threads = 16;  // tuning knob of the decoder's early syndrome test (MERCURY_B200_CHEAP_TEST in the environment)
	std::string err;
	uint64_t launches = 0;
	Slot slots[kSlots];
	void *d_scratch_llr = nullptr;  // internal-order LLRs of demod_decode_batch_device
	size_t scratch_frames = 0;
	float2 *dbg_Y = nullptr, *dbg_H = nullptr, *dbg_Z = nullptr;
	float *h_stage = nullptr;  // pinned staging for receive_baseband
	size_t h_stage_bytes = 0;
};

namespace {

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
	MB_CUDA(h, cudaMalloc(&h->d_blob, h->blob.size()));
	MB_CUDA(h, cudaMemcpy(h->d_blob, h->blob.data(), h->blob.size(), cudaMemcpyHostToDevice));
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
		 bool gi_removed = false)
{
	MbDemodArgs a;
	memset(&a, 0, sizeof(a));
	const MbMode &m = h->hdr.modes[h->config];
	a.x = static_cast<const float2 *>(d_x);
	a.sym_stride = gi_removed ? MB_NFFT : MB_NOFDM;
	a.sym_skip = gi_removed ? 0 : MB_NGI;
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
	const MbMode &m = h->hdr.modes[h->config];
	a.llr = static_cast<const float *>(d_llr);
	a.payload = static_cast<uint8_t *>(d_payload);
	a.stats = static_cast<MbRxStats *>(d_stats);
	a.blob = h->d_blob;
	a.mode = m;
	a.rate = h->hdr.rates[m.rate_idx];
	a.max_iters = h->ldpc_iters;
	a.check_gate = 1;
	a.cheap_test_threads = h->cheap_test_threads;
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
	if (h->d_blob) cudaFree(h->d_blob);
	if (h->h_stage) cudaFreeHost(h->h_stage);
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

int mercury_b200_load_configuration(mercury_b200_t *h, int config, int ldpc_iters)
{
	if (!h) return MERCURY_B200_EINVAL;
	if (h->blob.empty()) return fail(h, MERCURY_B200_ESTATE, "tables not loaded");
	// telecom_system.cc:2494-2497: out-of-range configurations are ignored by the reference; here they are an error
	if (config < 0 || config >= MB_NMODES) return fail(h, MERCURY_B200_EINVAL, "configuration must be 0..16 (CONFIG_0..CONFIG_16)");
	h->config = config;
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
	const MbMode &m = h->hdr.modes[h->config];
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
	return h->hdr.modes[h->config].frame_bytes;
}

int mercury_b200_get_frame_size_bits(const mercury_b200_t *h)
{
	if (!h || h->blob.empty() || h->config < 0) return MERCURY_B200_ESTATE;
	return h->hdr.modes[h->config].nReal - 16;
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
	rc = launch_demod(h, d_x, n, h->d_scratch_llr, d_stats, d_llr_cw, 0, s);
	if (rc) return rc;
	return launch_ldpc(h, h->d_scratch_llr, n, d_payload, d_stats, s);
}

int mercury_b200_demod_decode_batch(mercury_b200_t *h, const float *x, size_t n, uint8_t *payload, mercury_b200_rx_stats *stats, float *llr_cw)
{
	int rc = check_ready(h);
	if (rc) return rc;
	if (n == 0) return MERCURY_B200_OK;
	if (!x || !payload || !stats) return fail(h, MERCURY_B200_EINVAL, "null host buffer");
	const MbMode &m = h->hdr.modes[h->config];
	const size_t frame_x = (size_t)m.Nsymb * MB_NOFDM * sizeof(float2);
	// The guard interval never crosses PCIe: a strided (2-D) copy moves the 2,048 useful bytes of every 2,176-byte symbol, so the
	// device copy of a chunk is [frames][Nsymb][256] and the demodulator is told that the GI is already gone (-5.9 % of the bytes
	// of the link that bounds this path).
	const size_t sym_in = MB_NOFDM * sizeof(float2), sym_dev = MB_NFFT * sizeof(float2), gi = MB_NGI * sizeof(float2);
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
		rc = launch_demod(h, s.d_x, c, s.d_llr, s.d_stats, llr_cw ? s.d_llr_cw : nullptr, done, s.stream, /*gi_removed=*/true);
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
	const MbMode &m = h->hdr.modes[h->config];
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
