// mb_batcher.cpp -- multi-link frame batcher: many concurrent receive paths share one GPU batch.
//
// In the reference every link owns one cl_telecom_system and decodes one frame per receive_byte() call on its own thread:
// the audio thread shifts a symbol into passband_delayed_data under capture_prep_mutex (audioio.c:999-1069), the main loop
// snapshots the buffer under the same mutex and calls receive_byte() (arq_common.cc:2619-2668, telecom_system.cc:2207-2262).
// That single-frame call pattern is what makes a GPU pointless for ONE link; a gateway that terminates many links has
// thousands of such calls in flight.  The batcher keeps the call a link makes synchronous and per-frame (same contract as
// mercury_b200_receive_baseband) and turns the concurrency into batch size:
//
//   * callers copy their frame into the next free slot of a pinned staging buffer (two buffers, ping-pong) and sleep on
//     the buffer's condition variable;
//   (the pass-band flavour, mercury_b200_batcher_create_passband / _receive_byte, does the same with whole capture buffers and the
//   whole receive_byte(): front-end + tail, link state in and out)
//   * one worker thread closes a buffer when it is full or when its oldest frame has waited max_wait_us, runs ONE
//     mercury_b200_demod_decode_batch over it (H2D, two kernels, D2H, pipelined in chunks), and wakes the callers, who copy
//     their own payload / stats out.
//
// Everything here is host C++ over the library's own C ABI; no CUDA calls of its own.
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/mercury_b200.h"

namespace {

struct Buffer {
	uint8_t *x = nullptr;                     // pinned: capacity x in_bytes (one frame / one capture per slot)
	uint8_t *payload = nullptr;               // pinned: capacity x frame_bytes
	uint8_t *stats = nullptr;                 // pinned: capacity x stats_bytes records
	size_t filled = 0;                        // slots handed out
	size_t written = 0;                       // slots whose copy-in has finished
	size_t readers = 0;                       // callers that still have to copy their result out
	uint64_t generation = 0;                  // incremented when a batch completes
	bool closed = false;                      // no more slots: being decoded
	int rc = 0;                               // result of the batch call
	std::chrono::steady_clock::time_point first;
	std::condition_variable done;
};

}  // namespace

typedef int (*mb_batch_fn)(void *ctx, const void *x, size_t n, uint8_t *payload, void *stats);

struct mercury_b200_batcher {
	mercury_b200_t *h = nullptr;
	mb_batch_fn run = nullptr;  // what decodes a closed batch: the library's host-buffer batch call, or a test double
	void *run_ctx = nullptr;
	bool pinned = true;
	size_t capacity = 0, in_bytes = 0, frame_bytes = 0, stats_bytes = 0;
	bool stats_in = false;  // the caller's stats record is an input too (the link state of the whole receive_byte())
	int sample_format = 0;
	int (*user_run)(void *, const float *, size_t, uint8_t *, mercury_b200_rx_stats *) = nullptr;  // test double of the baseband flavour
	void *user_ctx = nullptr;
	std::chrono::microseconds max_wait{200};
	std::mutex mu;
	std::condition_variable work;   // worker: a buffer has frames / shutdown
	std::condition_variable space;  // callers: a buffer accepts frames again
	Buffer buf[2];
	int cur = 0;  // buffer currently accepting frames
	bool stop = false;
	size_t in_flight = 0;            // callers currently inside submit() (destroy waits for them before it frees anything)
	std::condition_variable idle;    // destroy: in_flight dropped to 0
	std::thread worker;
	uint64_t n_batches = 0, n_frames = 0, n_full = 0;
};

namespace {

void worker_loop(mercury_b200_batcher *b)
{
	std::unique_lock<std::mutex> lk(b->mu);
	for (;;) {
		Buffer &B = b->buf[b->cur];
		if (B.closed) {  // its previous batch is still being read out by the callers; the last one re-opens it and wakes us
			b->work.wait(lk);
			continue;
		}
		if (b->stop && B.filled == 0) return;
		if (B.filled == 0) {
			b->work.wait(lk);
			continue;
		}
		// close when full, on shutdown, or when the oldest frame has waited long enough
		if (B.filled < b->capacity && !b->stop) {
			const auto deadline = B.first + b->max_wait;
			if (std::chrono::steady_clock::now() < deadline) {
				b->work.wait_until(lk, deadline);
				continue;
			}
		}
		B.closed = true;
		// the other buffer takes over as soon as its previous batch has been read out
		b->cur ^= 1;
		b->space.notify_all();
		while (B.written < B.filled) b->work.wait(lk);  // copy-ins of the last slots still running
		const size_t n = B.filled;
		if (n == b->capacity) b->n_full++;
		lk.unlock();
		const int rc = b->run(b->run_ctx, B.x, n, B.payload, B.stats);
		lk.lock();
		B.rc = rc;
		B.readers = n;
		B.generation++;
		b->n_batches++;
		b->n_frames += n;
		B.done.notify_all();
	}
}

}  // namespace

extern "C" {

static int run_on_gpu(void *ctx, const void *x, size_t n, uint8_t *payload, void *stats)
{
	mercury_b200_batcher *b = static_cast<mercury_b200_batcher *>(ctx);
	return mercury_b200_demod_decode_batch(b->h, static_cast<const float *>(x), n, payload, static_cast<mercury_b200_rx_stats *>(stats), nullptr);
}

static int run_passband_on_gpu(void *ctx, const void *x, size_t n, uint8_t *payload, void *stats)
{
	mercury_b200_batcher *b = static_cast<mercury_b200_batcher *>(ctx);
	return mercury_b200_receive_byte_batch(b->h, x, b->sample_format, n, payload, static_cast<mercury_b200_receive_stats *>(stats), nullptr);
}

static int run_user(void *ctx, const void *x, size_t n, uint8_t *payload, void *stats)
{
	mercury_b200_batcher *b = static_cast<mercury_b200_batcher *>(ctx);
	return b->user_run(b->user_ctx, static_cast<const float *>(x), n, payload, static_cast<mercury_b200_rx_stats *>(stats));
}

static void free_buffers(mercury_b200_batcher *b)
{
	for (Buffer &B : b->buf) {
		if (b->pinned) {
			mercury_b200_host_free(B.x);
			mercury_b200_host_free(B.payload);
			mercury_b200_host_free(B.stats);
		} else {
			free(B.x);
			free(B.payload);
			free(B.stats);
		}
	}
}

static int create(size_t in_bytes, size_t frame_bytes, size_t stats_bytes, size_t max_batch, unsigned max_wait_us, mb_batch_fn run, bool pinned,
		  mercury_b200_batcher_t **out)
{
	mercury_b200_batcher *b = new (std::nothrow) mercury_b200_batcher;
	if (!b) return MERCURY_B200_ENOMEM;
	b->run = run, b->run_ctx = b, b->pinned = pinned;
	b->capacity = max_batch;
	b->in_bytes = in_bytes;
	b->frame_bytes = frame_bytes;
	b->stats_bytes = stats_bytes;
	b->max_wait = std::chrono::microseconds(max_wait_us);
	for (Buffer &B : b->buf) {
		auto alloc = [&](size_t bytes) { return pinned ? mercury_b200_host_alloc(bytes) : malloc(bytes); };
		B.x = static_cast<uint8_t *>(alloc(max_batch * in_bytes));
		B.payload = static_cast<uint8_t *>(alloc(max_batch * frame_bytes));
		B.stats = static_cast<uint8_t *>(alloc(max_batch * stats_bytes));
	}
	for (Buffer &B : b->buf)
		if (!B.x || !B.payload || !B.stats) {
			free_buffers(b);
			delete b;
			return MERCURY_B200_ENOMEM;
		}
	*out = b;
	return MERCURY_B200_OK;
}

static void start(mercury_b200_batcher *b) { b->worker = std::thread(worker_loop, b); }

int mercury_b200_batcher_create(mercury_b200_t *h, size_t max_batch, unsigned max_wait_us, mercury_b200_batcher_t **out)
{
	if (!h || !out || max_batch == 0) return MERCURY_B200_EINVAL;
	*out = nullptr;
	mercury_b200_geometry g;
	const int rc = mercury_b200_get_geometry(h, &g);
	if (rc != MERCURY_B200_OK) return rc;
	const int rc2 = create((size_t)g.Nsymb * MERCURY_B200_NOFDM * 2 * sizeof(float), (size_t)g.frame_bytes, sizeof(mercury_b200_rx_stats), max_batch, max_wait_us,
			       run_on_gpu, true, out);
	if (rc2 != MERCURY_B200_OK) return rc2;
	(*out)->h = h;
	start(*out);
	return MERCURY_B200_OK;
}

// The same machinery one level up (SURVEY.md 8f rows 1 + 4): every link hands its whole pass-band capture buffer to a synchronous
// receive_byte()-shaped call (arq_common.cc:2619-2668); the batch runs the GPU front-end + tail over all of them at once.
int mercury_b200_batcher_create_passband(mercury_b200_t *h, int sample_format, size_t max_batch, unsigned max_wait_us, mercury_b200_batcher_t **out)
{
	if (!h || !out || max_batch == 0) return MERCURY_B200_EINVAL;
	*out = nullptr;
	mercury_b200_geometry g;
	const int rc = mercury_b200_get_geometry(h, &g);
	if (rc != MERCURY_B200_OK) return rc;
	size_t ss;
	switch (sample_format) {
	case MERCURY_B200_SAMPLES_F64: ss = 8; break;
	case MERCURY_B200_SAMPLES_F32: ss = 4; break;
	case MERCURY_B200_SAMPLES_I16: ss = 2; break;
	case MERCURY_B200_SAMPLES_I32: ss = 4; break;
	default: return MERCURY_B200_EINVAL;
	}
	const int cs = mercury_b200_get_capture_samples(h);
	if (cs <= 0) return MERCURY_B200_ESTATE;
	const int rc2 = create((size_t)cs * ss, (size_t)g.frame_bytes, sizeof(mercury_b200_receive_stats), max_batch, max_wait_us, run_passband_on_gpu, true, out);
	if (rc2 != MERCURY_B200_OK) return rc2;
	(*out)->h = h;
	(*out)->sample_format = sample_format;
	(*out)->stats_in = true;
	start(*out);
	return MERCURY_B200_OK;
}

// Test hook (not part of the product surface): the same batching machinery in front of a caller-supplied batch function, so
// that its concurrency logic can be exercised -- and run under ThreadSanitizer -- on a machine without a GPU.
int mercury_b200_batcher_create_with_backend(size_t frame_floats, size_t frame_bytes, size_t max_batch, unsigned max_wait_us,
					     int (*run)(void *, const float *, size_t, uint8_t *, mercury_b200_rx_stats *), void *ctx,
					     mercury_b200_batcher_t **out)
{
	if (!run || !out || max_batch == 0 || frame_floats == 0 || frame_bytes == 0) return MERCURY_B200_EINVAL;
	*out = nullptr;
	const int rc = create(frame_floats * sizeof(float), frame_bytes, sizeof(mercury_b200_rx_stats), max_batch, max_wait_us, run_user, false, out);
	if (rc != MERCURY_B200_OK) return rc;
	(*out)->user_run = run, (*out)->user_ctx = ctx;
	start(*out);
	return MERCURY_B200_OK;
}

static int submit(mercury_b200_batcher_t *b, const void *in, uint8_t *payload, void *stats)
{
	std::unique_lock<std::mutex> lk(b->mu);
	if (b->stop) return MERCURY_B200_ESTATE;
	b->in_flight++;
	struct Leave {  // every exit below runs with the lock held
		mercury_b200_batcher *b;
		~Leave()
		{
			if (--b->in_flight == 0) b->idle.notify_all();
		}
	} leave{b};
	Buffer *B;
	for (;;) {  // a buffer that accepts frames: not closed, not full, previous results all read out
		if (b->stop) return MERCURY_B200_ESTATE;
		B = &b->buf[b->cur];
		if (!B->closed && B->filled < b->capacity && B->readers == 0) break;
		b->space.wait(lk);
	}
	const size_t slot = B->filled++;
	const uint64_t gen = B->generation;
	if (slot == 0) B->first = std::chrono::steady_clock::now();
	if (slot == 0 || B->filled == b->capacity) b->work.notify_one();
	lk.unlock();
	memcpy(B->x + slot * b->in_bytes, in, b->in_bytes);  // outside the lock: links copy in parallel
	if (b->stats_in) memcpy(B->stats + slot * b->stats_bytes, stats, b->stats_bytes);
	lk.lock();
	B->written++;
	if (B->closed && B->written == B->filled) b->work.notify_one();
	B->done.wait(lk, [&] { return B->generation != gen; });
	const int rc = B->rc;
	lk.unlock();
	if (rc == MERCURY_B200_OK) {
		memcpy(payload, B->payload + slot * b->frame_bytes, b->frame_bytes);
		memcpy(stats, B->stats + slot * b->stats_bytes, b->stats_bytes);
	}
	lk.lock();
	if (--B->readers == 0) {  // last reader re-opens the buffer
		B->filled = B->written = 0;
		B->closed = false;
		b->space.notify_all();
		b->work.notify_one();
	}
	return rc;
}

int mercury_b200_batcher_receive_baseband(mercury_b200_batcher_t *b, const float *baseband, uint8_t *payload, mercury_b200_rx_stats *stats)
{
	if (!b || !baseband || !payload || !stats || b->stats_in) return MERCURY_B200_EINVAL;
	return submit(b, baseband, payload, stats);
}

int mercury_b200_batcher_receive_byte(mercury_b200_batcher_t *b, const void *passband, uint8_t *payload, mercury_b200_receive_stats *stats)
{
	if (!b || !passband || !payload || !stats || !b->stats_in) return MERCURY_B200_EINVAL;
	return submit(b, passband, payload, stats);
}

int mercury_b200_batcher_get_counters(mercury_b200_batcher_t *b, uint64_t *batches, uint64_t *frames, uint64_t *full_batches)
{
	if (!b) return MERCURY_B200_EINVAL;
	std::lock_guard<std::mutex> lk(b->mu);
	if (batches) *batches = b->n_batches;
	if (frames) *frames = b->n_frames;
	if (full_batches) *full_batches = b->n_full;
	return MERCURY_B200_OK;
}

void mercury_b200_batcher_destroy(mercury_b200_batcher_t *b)
{
	if (!b) return;
	{
		std::lock_guard<std::mutex> lk(b->mu);
		b->stop = true;
	}
	b->work.notify_all();
	b->space.notify_all();
	if (b->worker.joinable()) b->worker.join();
	{
		// The worker leaves as soon as the buffer it looks at is empty; callers of the other buffer may still be copying their results
		// out, or be on their way out of submit().  Nothing is freed before the last of them has left (they hold the mutex when they do).
		std::unique_lock<std::mutex> lk(b->mu);
		b->idle.wait(lk, [&] { return b->in_flight == 0; });
	}
	free_buffers(b);
	delete b;
}

}  // extern "C"
