// capi.cu -- transform entry points of the C ABI: cdf53 / icdf53 (cdf53.h:9,36), the two `transformation`
// drivers (encode.c:16-30, decode.c:16-30) and the colour transforms (image.h:67-79) on host buffers.
#include "pipeline.cuh"
#include "layout.cuh"

#include "dwt_b200.h"

#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>
#include <string.h>


static thread_local dwt_ctx *g_default_ctx = nullptr;

// a temporary device allocation that is released on every return path
struct DevTmp {
	int *p = nullptr;
	cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
	~DevTmp()
	{
		if (p)
			cudaFree(p);
	}
};

static dwt_ctx *default_ctx()
{
	if (!g_default_ctx)
		g_default_ctx = dwt_ctx_create(-1);
	return g_default_ctx;
}

// root LL in ll[0] (planar, pitch w[0]) + details in pyr -> image
int ctx_inverse_transform(dwt_ctx *c, int levels_used, bool to_u8, int *planar_out, bool counters_zeroed)
{
	const Geom &g = c->geom;
	cudaStream_t st = c->st;
	const int top = levels_used > 0 ? levels_used : 0;
	const long long pyr_stride = g.pix[top];
	const int pyr_pitch = g.w[top];
	if (!counters_zeroed && ctx_zero_transform_counters(c, false)) // work counters of the level launches
		return -1;
	if (levels_used == 0) {
		// decode.c:258 with levels == 0 still runs one inverse level on the w0 x h0 root, reading the root
		// itself as a one-level Mallat pyramid
		LiftLevel p;
		p.W = g.w[0];
		p.H = g.h[0];
		p.channels = g.channels;
		p.in = c->ll[0].p;
		p.in_chan_stride = g.pix[0];
		p.in_pitch = g.w[0];
		p.pyr = c->ll[0].as<int>();
		p.pyr_chan_stride = g.pix[0];
		p.pyr_pitch = g.w[0];
		p.maxabs = nullptr;
		p.work = c->small.as<int>() + 64;
		int mode = 2;
		if (to_u8) {
			p.out = c->img.p;
			p.out_chan_stride = 0;
			p.out_pitch = g.w[0];
			mode = g.channels == 3 ? 0 : 1;
		} else {
			p.out = planar_out;
			p.out_chan_stride = g.pix[0];
			p.out_pitch = g.w[0];
		}
		return lift_inverse_level(p, mode, st, &c->launches);
	}
	// the coarsest levels that fit one CTA's shared memory run fused (root in ll[0] -> level T in ll[1])
	int first = 1, cur = 0;
	for (int T = levels_used - 1; T >= 1; --T) {
		if (!lift_tail_fits(g.w[T], g.h[T]))
			continue;
		LiftTail t;
		t.ll_in = c->ll[0].as<int>();
		t.in_chan_stride = g.pix[0];
		t.in_pitch = g.w[0];
		t.ll_out = c->ll[1].as<int>();
		t.out_chan_stride = g.pix[T];
		t.out_pitch = g.w[T];
		t.pyr = c->pyr.as<int>();
		t.pyr_chan_stride = pyr_stride;
		t.pyr_pitch = pyr_pitch;
		t.W = g.w[T];
		t.H = g.h[T];
		t.nlev = T;
		t.channels = g.channels;
		t.maxabs = nullptr;
		if (lift_tail(t, true, st, &c->launches))
			return -1;
		first = T + 1;
		cur = 1;
		break;
	}
	for (int lv = first; lv <= levels_used; ++lv) {
		LiftLevel p;
		p.W = g.w[lv];
		p.H = g.h[lv];
		p.channels = g.channels;
		p.in = c->ll[cur].p;
		p.in_chan_stride = g.pix[lv - 1];
		p.in_pitch = g.w[lv - 1];
		p.pyr = c->pyr.as<int>();
		p.pyr_chan_stride = pyr_stride;
		p.pyr_pitch = pyr_pitch;
		p.maxabs = nullptr;
		p.work = c->small.as<int>() + 64 + lv;
		p.chained = !(lv == first && cur == 0); // all but the first launch of the chain
		int mode = 2;
		if (lv == levels_used && to_u8) {
			p.out = c->img.p;
			p.out_chan_stride = 0;
			p.out_pitch = g.w[lv];
			mode = g.channels == 3 ? 0 : 1;
		} else if (lv == levels_used) {
			p.out = planar_out;
			p.out_chan_stride = g.pix[lv];
			p.out_pitch = g.w[lv];
		} else {
			p.out = c->ll[cur ^ 1].p;
			p.out_chan_stride = g.pix[lv];
			p.out_pitch = g.w[lv];
		}
		if (lift_inverse_level(p, mode, st, &c->launches))
			return -1;
		cur ^= 1;
	}
	return 0;
}

// ------------------------------------------------------------------------------------------------ 1-D entry points

static void run_1d(int *out, int *in, int N, int SO, int SI, int CH, bool inverse)
{
	dwt_ctx *c = default_ctx();
	if (!c || N < 2 || CH < 1) {
		if (!c)
			fprintf(stderr, "libdwt_b200: %s\n", dwt_last_error());
		return; // the reference signature has no error path
	}
	cudaStream_t st = c->st;
	const size_t span_in = (size_t)(N - 1) * SI + CH, span_out = (size_t)(N - 1) * SO + CH;
	DevTmp d_in, d_out;
	// the reference signature returns nothing: a CUDA failure is reported on stderr and leaves `out` untouched
	cudaError_t e = cudaSetDevice(c->device);
	if (e == cudaSuccess)
		e = d_in.alloc(span_in * sizeof(int));
	if (e == cudaSuccess)
		e = d_out.alloc(span_out * sizeof(int));
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(d_in.p, in, span_in * sizeof(int), cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(d_out.p, out, span_out * sizeof(int), cudaMemcpyHostToDevice, st); // untouched gaps survive
	if (e == cudaSuccess && lift_cdf53_1d(d_out.p, d_in.p, N, SO, SI, CH, inverse, st))
		e = cudaErrorUnknown;
	c->launches += 1;
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(out, d_out.p, span_out * sizeof(int), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess && !inverse) // cdf53.h:12-23 lifts in place: the caller sees the lifted samples in `in`
		e = cudaMemcpyAsync(in, d_in.p, span_in * sizeof(int), cudaMemcpyDeviceToHost, st);
	const cudaError_t e2 = cudaStreamSynchronize(st);
	if (e == cudaSuccess)
		e = e2;
	if (e != cudaSuccess) {
		cudaGetLastError();
		fprintf(stderr, "libdwt_b200: %s failed: %s\n", inverse ? "icdf53" : "cdf53",
		        e == cudaErrorUnknown ? dwt_last_error() : cudaGetErrorString(e));
	}
}

extern "C" void cdf53(int *out, int *in, int N, int SO, int SI, int CH)
{
	run_1d(out, in, N, SO, SI, CH, false);
}

extern "C" void icdf53(int *out, int *in, int N, int SO, int SI, int CH)
{
	run_1d(out, in, N, SO, SI, CH, true);
}

// ------------------------------------------------------------------------------------------------ 2-D drivers

extern "C" int dwt_forward(int *out, const int *in, int W, int H, int CH)
{
	dwt_ctx *c = default_ctx();
	if (!c)
		return -1;
	if (W < 2 || H < 2 || CH < 1 || CH > 3) {
		dwt_set_error("dwt_forward: unsupported shape %dx%dx%d", W, H, CH);
		return -1;
	}
	CUDA_OK(cudaSetDevice(c->device));
	if (ctx_set_geometry(c, W, H, CH))
		return -1;
	const Geom &g = c->geom;
	const int L = g.levels;
	cudaStream_t st = c->st;
	const long long npix = (long long)W * H;
	const size_t n = (size_t)npix * CH;
	DevTmp ta, tb;
	CUDA_OK(ta.alloc(n * sizeof(int)));
	CUDA_OK(tb.alloc(n * sizeof(int)));
	int *d_a = ta.p, *d_b = tb.p;
	CUDA_OK(cudaMemcpyAsync(d_a, in, n * sizeof(int), cudaMemcpyHostToDevice, st));
	deinterleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a, d_b, npix, CH);
	int r = ctx_forward_transform(c, d_b);
	if (!r) {
		export_pyramid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->pyr.as<int>(), ctx_root_ll(c), d_a, W, H, CH,
		                                                                   g.w[0], g.h[0]);
		c->launches += 2;
		cudaMemcpyAsync(out, d_a, n * sizeof(int), cudaMemcpyDeviceToHost, st);
	}
	cudaError_t e = cudaStreamSynchronize(st);
	(void)L;
	if (r || e != cudaSuccess) {
		if (!r)
			dwt_set_error("dwt_forward: %s", cudaGetErrorString(e));
		return -1;
	}
	return 0;
}

extern "C" int dwt_inverse(int *out, const int *in, int W, int H, int CH)
{
	dwt_ctx *c = default_ctx();
	if (!c)
		return -1;
	if (W < 2 || H < 2 || CH < 1 || CH > 3) {
		dwt_set_error("dwt_inverse: unsupported shape %dx%dx%d", W, H, CH);
		return -1;
	}
	CUDA_OK(cudaSetDevice(c->device));
	if (ctx_set_geometry(c, W, H, CH))
		return -1;
	const Geom &g = c->geom;
	const int L = g.levels;
	cudaStream_t st = c->st;
	if (ensure_transform_buffers(c))
		return -1;
	const long long npix = (long long)W * H;
	const size_t n = (size_t)npix * CH;
	DevTmp ta, tb;
	CUDA_OK(ta.alloc(n * sizeof(int)));
	CUDA_OK(tb.alloc(n * sizeof(int)));
	int *d_a = ta.p, *d_b = tb.p;
	CUDA_OK(cudaMemcpyAsync(d_a, in, n * sizeof(int), cudaMemcpyHostToDevice, st));
	// planar pyramid; the root rectangle is copied to ll[0] with its own pitch
	deinterleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a, c->pyr.as<int>(), npix, CH);
	for (int ch = 0; ch < CH; ++ch)
		CUDA_OK(cudaMemcpy2DAsync(c->ll[0].as<int>() + (size_t)ch * g.pix[0], sizeof(int) * g.w[0],
		                          c->pyr.as<int>() + (size_t)ch * npix, sizeof(int) * W, sizeof(int) * g.w[0], g.h[0],
		                          cudaMemcpyDeviceToDevice, st));
	int r = ctx_inverse_transform(c, L, false, d_b);
	if (!r) {
		interleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_b, d_a, npix, CH);
		c->launches += 2;
		cudaMemcpyAsync(out, d_a, n * sizeof(int), cudaMemcpyDeviceToHost, st);
	}
	cudaError_t e = cudaStreamSynchronize(st);
	if (r || e != cudaSuccess) {
		if (!r)
			dwt_set_error("dwt_inverse: %s", cudaGetErrorString(e));
		return -1;
	}
	return 0;
}

static int colour_host(int *buffer, int total, bool inverse)
{
	dwt_ctx *c = default_ctx();
	if (!c)
		return -1;
	if (total <= 0)
		return 0;
	CUDA_OK(cudaSetDevice(c->device));
	DevTmp td;
	size_t bytes = (size_t)total * 3 * sizeof(int);
	CUDA_OK(td.alloc(bytes));
	int *d = td.p;
	CUDA_OK(cudaMemcpyAsync(d, buffer, bytes, cudaMemcpyHostToDevice, c->st));
	int r = lift_colour(d, total, inverse, c->st);
	c->launches += 1;
	if (!r)
		CUDA_OK(cudaMemcpyAsync(buffer, d, bytes, cudaMemcpyDeviceToHost, c->st));
	cudaError_t e = cudaStreamSynchronize(c->st);
	if (r || e != cudaSuccess) {
		if (!r)
			dwt_set_error("colour transform: %s", cudaGetErrorString(e));
		return -1;
	}
	return 0;
}

extern "C" int dwt_ycocg_from_rgb(int *buffer, int total)
{
	return colour_host(buffer, total, false);
}

extern "C" int dwt_rgb_from_ycocg(int *buffer, int total)
{
	return colour_host(buffer, total, true);
}

// ------------------------------------------------------------------------------------------------ host staging + timing helpers

// Pinned (page-locked) host memory for callers that want PCIe-rate transfers: buffers handed to
// dwt_encode_into / dwt_decode_into are copied with cudaMemcpyAsync directly, without a pageable staging pass.
extern "C" void *dwt_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
		cudaGetLastError();
		dwt_set_error("cudaMallocHost(%zu) failed", bytes);
		return nullptr;
	}
	return p;
}

extern "C" void dwt_host_free(void *p)
{
	if (p)
		cudaFreeHost(p);
}

extern "C" int dwt_encode_into(dwt_ctx *c, const uint8_t *pixels, int width, int height, int channels, int capacity,
                               uint8_t *out, size_t out_room, size_t *out_len, struct dwt_stats *stats)
{
	if (!c) {
		dwt_set_error("null context (no CUDA device?)");
		return -1;
	}
	if (dwt_ctx_upload_image(c, pixels, width, height, channels))
		return -1;
	if (dwt_ctx_encode_resident(c, capacity, stats)) {
		cudaStreamSynchronize(c->st); // the upload may still be reading `pixels`
		return -1;
	}
	if (c->out_bytes > out_room) {
		cudaStreamSynchronize(c->st);
		dwt_set_error("output buffer too small: need %zu bytes", c->out_bytes);
		*out_len = c->out_bytes;
		return -1;
	}
	if (c->out_bytes) {
		CUDA_OK(ctx_copy(c, out, c->out.p, c->out_bytes, cudaMemcpyDeviceToHost, true));
	}
	*out_len = c->out_bytes;
	return 0;
}

extern "C" int dwt_decode_into(dwt_ctx *c, const uint8_t *stream, size_t len, int pixels_max, uint8_t *pixels,
                               size_t pixels_room, int *width, int *height, int *channels, struct dwt_stats *stats)
{
	if (!c) {
		dwt_set_error("null context (no CUDA device?)");
		return -1;
	}
	if (ctx_upload_stream(c, stream, len, false)) // `stream` outlives the call: no separate wait for the upload
		return -1;
	int r = dwt_ctx_decode_resident(c, pixels_max, stats);
	if (r) {
		cudaStreamSynchronize(c->st); // the upload may still be reading `stream`
		return r;
	}
	const size_t n = (size_t)c->dec_w * c->dec_h * c->dec_ch;
	*width = c->dec_w;
	*height = c->dec_h;
	*channels = c->dec_ch;
	if (n > pixels_room) {
		cudaStreamSynchronize(c->st);
		dwt_set_error("pixel buffer too small: need %zu bytes", n);
		return -1;
	}
	CUDA_OK(ctx_copy(c, pixels, c->img.p, n, cudaMemcpyDeviceToHost, true));
	return 0;
}

// ------------------------------------------------------------------------------------------------ batches
//
// Images are independent: a pool owns `workers` contexts on one device (one CUDA stream each) and codes
// the items of a batch on as many host threads, so the copies of one item overlap the kernels of the others and the
// single-warp stages of one frame hide behind the wide kernels of its neighbours.  One pool per GPU; no collective.

struct dwt_pool {
	std::vector<int> devices;              // the GPUs of the pool, in the caller's order
	int workers = 1;                       // contexts (and host threads) per device
	std::vector<dwt_ctx *> ctx;            // device d owns ctx[d * workers .. (d + 1) * workers)
	std::vector<std::unique_ptr<XferGate>> gate; // one per device: large copies of a direction take turns on that GPU's link
	std::mutex err_lock;
	char err[512] = "";
};

extern "C" const char *dwt_pool_last_error(const dwt_pool *p)
{
	return p ? p->err : "";
}

// How many contexts the caller keeps busy on this device at the same time (default 1).  With several frames in flight the
// library prefers kernels that do less total work over kernels that finish one frame sooner (decoder scan).
// this GPU's share of the box's cores: more waiting threads than that should sleep, not spin (ctx_stream_sync)
static bool oversubscribed(int threads_on_this_gpu)
{
	int count = 1;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1)
		count = 1;
	cpu_set_t cpus;
	const int ncpu = sched_getaffinity(0, sizeof(cpus), &cpus) == 0 ? CPU_COUNT(&cpus) : 1;
	return threads_on_this_gpu > 1 && threads_on_this_gpu > ncpu / count;
}

extern "C" int dwt_ctx_set_in_flight(dwt_ctx *c, int contexts)
{
	if (!c || contexts < 1)
		return -1;
	c->in_flight = contexts;
	c->sleepy_wait = oversubscribed(contexts);
	return 0;
}

extern "C" int dwt_ctx_set_decoder_scan(dwt_ctx *c, int mode)
{
	if (!c || mode < 0 || mode > 2)
		return -1;
	c->scan_mode = mode;
	return 0;
}

extern "C" dwt_pool *dwt_pool_create_multi(const int *devices, int n_devices, int workers)
{
	if (workers < 1)
		workers = 1;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
		dwt_set_error("no CUDA device: libdwt_b200 has no CPU fallback");
		return nullptr;
	}
	dwt_pool *p = new dwt_pool();
	p->workers = workers;
	const bool sleepy = oversubscribed(workers);
	if (!devices || n_devices <= 0) { // every visible device
		for (int d = 0; d < count; ++d)
			p->devices.push_back(d);
	} else {
		for (int i = 0; i < n_devices; ++i)
			p->devices.push_back(devices[i]);
	}
	for (size_t d = 0; d < p->devices.size(); ++d) {
		p->gate.emplace_back(new XferGate());
		for (int i = 0; i < workers; ++i) {
			dwt_ctx *c = dwt_ctx_create(p->devices[d]);
			if (!c) {
				for (dwt_ctx *x : p->ctx)
					dwt_ctx_destroy(x);
				delete p;
				return nullptr;
			}
			c->in_flight = workers;
			c->sleepy_wait = sleepy;
			c->gate = workers > 1 ? p->gate[d].get() : nullptr;
			p->ctx.push_back(c);
			p->devices[d] = c->device; // device < 0 resolved to the current device
		}
	}
	return p;
}

extern "C" dwt_pool *dwt_pool_create(int device, int workers)
{
	return dwt_pool_create_multi(&device, 1, workers);
}

extern "C" void dwt_pool_destroy(dwt_pool *p)
{
	if (!p)
		return;
	for (dwt_ctx *c : p->ctx)
		dwt_ctx_destroy(c);
	delete p;
}

extern "C" int dwt_pool_workers(const dwt_pool *p)
{
	return p ? (int)p->ctx.size() : 0;
}

extern "C" int dwt_pool_devices(const dwt_pool *p, int *devices, int room)
{
	if (!p)
		return 0;
	for (int i = 0; i < room && i < (int)p->devices.size(); ++i)
		devices[i] = p->devices[(size_t)i];
	return (int)p->devices.size();
}

// Item i of a batch belongs to device i mod G (G = devices of the pool): images are independent, so the GPUs share nothing
// (no collective).  The `workers` threads of a device hand its items out among themselves.
template <typename F>
static int pool_run(dwt_pool *p, int n, F &&one)
{
	if (!p || n < 0) {
		dwt_set_error("bad batch arguments");
		return -1;
	}
	const int G = (int)p->devices.size();
	std::vector<std::atomic<int>> next((size_t)G);
	for (auto &a : next)
		a.store(0);
	std::atomic<int> failed(0);
	auto work = [&](int d, int w) {
		for (;;) {
			const int i = d + G * next[(size_t)d].fetch_add(1);
			if (i >= n)
				break;
			if (one(p->ctx[(size_t)(d * p->workers + w)], i)) {
				failed.fetch_add(1);
				std::lock_guard<std::mutex> hold(p->err_lock); // dwt_last_error() is per thread: keep the reason with the pool
				snprintf(p->err, sizeof(p->err), "item %d: %s", i, dwt_last_error());
			}
		}
	};
	std::vector<std::thread> th;
	for (int d = 0; d < G; ++d) {
		const int items = n > d ? (n - d + G - 1) / G : 0;
		const int nw = p->workers < items ? p->workers : items;
		for (int w = 0; w < nw; ++w)
			th.emplace_back(work, d, w);
	}
	for (auto &t : th)
		t.join();
	return failed.load();
}

extern "C" int dwt_pool_encode(dwt_pool *p, struct dwt_encode_item *items, int n)
{
	return pool_run(p, n, [&](dwt_ctx *c, int i) {
		dwt_encode_item &it = items[i];
		it.status = dwt_encode_into(c, it.pixels, it.width, it.height, it.channels, it.capacity, it.out, it.out_room, &it.out_len,
		                            nullptr);
		return it.status != 0;
	});
}

extern "C" int dwt_pool_decode(dwt_pool *p, struct dwt_decode_item *items, int n)
{
	return pool_run(p, n, [&](dwt_ctx *c, int i) {
		dwt_decode_item &it = items[i];
		it.status = dwt_decode_into(c, it.stream, it.len, it.pixels_max, it.pixels, it.pixels_room, &it.width, &it.height,
		                            &it.channels, nullptr);
		return it.status != 0;
	});
}

// encode and decode items of one batch in one go, interleaved on the workers (a transcoding front end has both kinds
// of job in flight: uploads of pixels overlap downloads of pixels, streams flow the other way)
extern "C" int dwt_pool_run(dwt_pool *p, struct dwt_encode_item *enc, int n_enc, struct dwt_decode_item *dec, int n_dec)
{
	if (n_enc < 0 || n_dec < 0 || (n_enc && !enc) || (n_dec && !dec)) {
		dwt_set_error("bad batch arguments");
		return -1;
	}
	const int lo = n_enc < n_dec ? n_enc : n_dec;
	return pool_run(p, n_enc + n_dec, [&](dwt_ctx *c, int i) {
		// jobs 0 .. 2*lo-1 alternate encode / decode, the rest is whatever kind is left
		bool is_enc;
		int k;
		if (i < 2 * lo) {
			is_enc = (i & 1) == 0;
			k = i >> 1;
		} else {
			is_enc = n_enc > n_dec;
			k = lo + (i - 2 * lo);
		}
		if (is_enc) {
			dwt_encode_item &it = enc[k];
			it.status = dwt_encode_into(c, it.pixels, it.width, it.height, it.channels, it.capacity, it.out, it.out_room,
			                            &it.out_len, nullptr);
			return it.status != 0;
		}
		dwt_decode_item &it = dec[k];
		it.status = dwt_decode_into(c, it.stream, it.len, it.pixels_max, it.pixels, it.pixels_room, &it.width, &it.height,
		                            &it.channels, nullptr);
		return it.status != 0;
	});
}

// write a buffer larger than L2 (126 MB) so the next timed step starts from HBM
extern "C" int dwt_ctx_flush_l2(dwt_ctx *c)
{
	CUDA_OK(cudaSetDevice(c->device));
	const size_t bytes = 256u << 20;
	if (c->flush.ensure(bytes))
		return -1;
	CUDA_OK(cudaMemsetAsync(c->flush.p, 0x5a, bytes, c->st));
	CUDA_OK(cudaStreamSynchronize(c->st));
	return 0;
}

// CUDA events on the context's stream (slots 0..3) for callers that time several calls as one region
extern "C" int dwt_ctx_event_record(dwt_ctx *c, int slot)
{
	if (slot < 0 || slot > 3)
		return -1;
	CUDA_OK(cudaSetDevice(c->device));
	CUDA_OK(cudaEventRecord(c->ev[4 + slot], c->st));
	return 0;
}

// make `c`'s stream wait for everything queued so far on `other`'s stream (frames coded concurrently on several
// contexts are timed as one region with events of one context)
extern "C" int dwt_ctx_wait_for(dwt_ctx *c, dwt_ctx *other)
{
	if (!c || !other)
		return -1;
	if (c == other)
		return 0;
	CUDA_OK(cudaSetDevice(c->device));
	CUDA_OK(cudaEventRecord(other->ev[8], other->st));
	CUDA_OK(cudaStreamWaitEvent(c->st, other->ev[8], 0));
	return 0;
}

extern "C" float dwt_ctx_event_elapsed_ms(dwt_ctx *c, int slot_a, int slot_b)
{
	float ms = -1.f;
	if (slot_a < 0 || slot_a > 3 || slot_b < 0 || slot_b > 3)
		return ms;
	cudaSetDevice(c->device);
	cudaEventSynchronize(c->ev[4 + slot_b]);
	if (cudaEventElapsedTime(&ms, c->ev[4 + slot_a], c->ev[4 + slot_b]) != cudaSuccess) {
		cudaGetLastError();
		return -1.f;
	}
	return ms;
}
