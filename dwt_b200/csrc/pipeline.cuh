// pipeline.cuh -- the per-device codec context (host orchestration of the kernels)
#pragma once
#include <mutex>
#include "coder.cuh"
#include "hilbert.cuh"
#include "lift.cuh"

#include <stddef.h>

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes); // grows (never shrinks); contents are lost on growth
	void release();
	template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct PinBuf {
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes);
	void release();
	template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct dwt_stats;

struct dwt_ctx {
	int device = 0;
	cudaStream_t st = nullptr;
	long long launches = 0;
	cudaEvent_t ev[9] = {};   // 0..3 stage timers, 4..7 caller slots, 8 cross-context waits
	struct XferGate *gate = nullptr; // set by a pool: see ctx_copy
	cudaEvent_t sync_ev = nullptr; // ctx_stream_sync: polled briefly, then waited for with the thread asleep
	bool sleepy_wait = false;      // set for the contexts of a pool (many waiting threads); a lone context spins
	cudaEvent_t xfer_ev[2] = {};   // completion of this context's latest gated copy, per direction

	// geometry cache
	bool have_geom = false;
	Geom geom;
	HilbertPlan plan;
	Sched sched;

	// image / transform buffers
	DevBuf img;        // u8 interleaved
	DevBuf pyr;        // planar int32 Mallat pyramid
	DevBuf ll[2];      // ping-pong LL
	int root_buf = 0;  // which of them holds the root after a forward transform
	DevBuf small;      // maxabs[4] | missing[48] | misc
	int img_w = 0, img_h = 0, img_ch = 0;
	bool img_resident = false;

	// coder buffers
	DevBuf bs, sig, ent, Z, signbuf, specbuf, refbuf, tiles, thr_state, chunks, info, dsched, out, stream;
	DevBuf mem_pref, ref_pref, ones_rank, sign_rank, dstate, win, flush;
	DevBuf dec_scan, dec_seg, dec_chunks; // decoder: per-slice chain tables, segment and chunk records
	DevBuf dec_lut;    // decoder: order-0 token table
	bool dec_lut_ready = false;
	int sm_count = 1;
	int scan_mode = 0;   // dwt_ctx_set_decoder_scan: 0 auto, 1 parallel, 2 serial
	int in_flight = 1;   // contexts the caller keeps busy on this device (throughput- vs latency-oriented kernels)
	PinBuf pin_small, pin_io, pin_stream;

	// last encode result (device resident)
	size_t out_bytes = 0;      // valid bytes in `out` (already truncated to the capacity)
	// last decode result
	int dec_w = 0, dec_h = 0, dec_ch = 0;
	size_t stream_len = 0, stream_head = 0;
	bool stream_resident = false;
};

int ctx_set_geometry(dwt_ctx *c, int w, int h, int ch);
void build_schedule(const Geom &g, const int *planes, Sched *s);
// img (or a planar int32 image when planar_in != NULL) -> pyr (+ root LL in an ll buffer), maxabs in small
// counters_zeroed: the caller already issued ctx_zero_transform_counters on the stream (keeps the small memset out of a
// timed lifting region)
int ctx_zero_transform_counters(dwt_ctx *c, bool forward);
int ctx_forward_transform(dwt_ctx *c, const int *planar_in, bool counters_zeroed = false);
// root LL in ll[0] + details in pyr (pitch w[levels_used]) -> u8 image in img (to_u8) or planar int32
int ctx_inverse_transform(dwt_ctx *c, int levels_used, bool to_u8, int *planar_out, bool counters_zeroed = false);
int ensure_transform_buffers(dwt_ctx *c);
// One large host<->device copy per direction at a time among the contexts of a pool: copies issued together share the link,
// so every job of a batch would get its data only when all of them have it (an idle GPU for the first ~20 ms of a batch
// of 8K frames); one after the other, the first job computes after one copy time.
// The turn-taking is an event chain on the device's timeline, not a lock held by a waiting host thread: a gated copy
// waits (cudaStreamWaitEvent) for the previous gated copy of its direction and leaves its own completion event behind.
// The mutex only covers the few microseconds of queueing those three calls.
struct XferGate {
	std::mutex dir[2];                      // [0] host -> device, [1] device -> host
	cudaEvent_t last[2] = {nullptr, nullptr}; // completion of the most recent gated copy (owned by the context that queued it)
};
// copy on the context's stream and, if `wait`, wait for it.  Gated device -> host copies first wait for the stream's
// kernels, so that the copies queued behind them in the chain are not held up by this context's compute.
cudaError_t ctx_copy(dwt_ctx *c, void *dst, const void *src, size_t n, cudaMemcpyKind kind, bool wait);
// wait for the context's stream.  A lone context spins in cudaStreamSynchronize: a frame has three such waits and a
// sleeping thread wakes up 100-300 us late (measured: 6.8 -> 10.1 ms per 8K round trip).  The contexts of a pool with more workers
// than this GPU's share of the box's cores poll an event for ~100 us and then sleep in cudaEventSynchronize
// (cudaEventBlockingSync); the other contexts keep the GPU busy meanwhile.  (Sleeping always was measured too: +3 % end
// to end on 8K frames, -15 % on batches of 1080p images whose waits are short, at 24 cores for one GPU.)
// DWT_SYNC=spin / block forces one behaviour, DWT_SPIN_US sets the poll time.
cudaError_t ctx_stream_sync(dwt_ctx *c);
int ctx_upload_stream(dwt_ctx *c, const uint8_t *stream, size_t len, bool wait);
const int *ctx_root_ll(dwt_ctx *c);      // device pointer to the planar root LL after ctx_forward_transform
