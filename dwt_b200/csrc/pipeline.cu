// pipeline.cu -- codec context and the encode side of the C ABI (include/dwt_b200.h).
//
// Host orchestration mirrors main() of encode.c:133-232: geometry (utils.h:28-40), colour + multi-level
// lifting, linearisation, per-channel plane counts, header / root image / plane counts (host, through the
// reference-shaped stream entry points of host/streamio.c), then the chunk schedule of encode.c:183-221
// handed to the GPU bit-plane coder.  The byte capacity is applied as a prefix cut (bytes.h:75-85).
#include "pipeline.cuh"
#include "layout.cuh"

#include "../host/streamio_internal.h"
#include "dwt_b200.h"

#include <atomic>
#include <chrono>

#include <sched.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------ errors

static thread_local char g_err[512];

void dwt_set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *dwt_last_error(void)
{
	return g_err;
}

int dwt_device_sms()
{
	static std::atomic<int> cache[64];
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0)
		dev = 0;
	const int slot = dev & 63;
	int n = cache[slot].load(std::memory_order_relaxed);
	if (n <= 0) {
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
			n = 1;
		cache[slot].store(n, std::memory_order_relaxed);
	}
	return n;
}

// ------------------------------------------------------------------------------------------------ buffers

int DevBuf::ensure(size_t bytes)
{
	if (bytes <= cap)
		return 0;
	if (p)
		cudaFree(p);
	p = nullptr;
	cap = 0;
	size_t want = bytes + bytes / 8 + 256;
	CUDA_OK(cudaMalloc(&p, want));
	cap = want;
	return 0;
}

void DevBuf::release()
{
	if (p)
		cudaFree(p);
	p = nullptr;
	cap = 0;
}

int PinBuf::ensure(size_t bytes)
{
	if (bytes <= cap)
		return 0;
	if (p)
		cudaFreeHost(p);
	p = nullptr;
	cap = 0;
	size_t want = bytes + bytes / 8 + 256;
	CUDA_OK(cudaMallocHost(&p, want));
	cap = want;
	return 0;
}

void PinBuf::release()
{
	if (p)
		cudaFreeHost(p);
	p = nullptr;
	cap = 0;
}

// ------------------------------------------------------------------------------------------------ context

extern "C" dwt_ctx *dwt_ctx_create(int device)
{
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
		dwt_set_error("no CUDA device: libdwt_b200 has no CPU fallback");
		return nullptr;
	}
	if (device < 0 && cudaGetDevice(&device) != cudaSuccess)
		device = 0;
	if (device >= count) {
		dwt_set_error("device %d out of range (%d devices)", device, count);
		return nullptr;
	}
	if (cudaSetDevice(device) != cudaSuccess) {
		dwt_set_error("cudaSetDevice(%d) failed", device);
		return nullptr;
	}
	dwt_ctx *c = new dwt_ctx();
	c->device = device;
	if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) {
		dwt_set_error("cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
		delete c;
		return nullptr;
	}
	for (auto &e : c->ev)
		cudaEventCreate(&e);
	cudaEventCreateWithFlags(&c->sync_ev, cudaEventDisableTiming | cudaEventBlockingSync);
	for (auto &e : c->xfer_ev)
		cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
	c->plan.cell_base = nullptr;
	return c;
}

cudaError_t ctx_stream_sync(dwt_ctx *c)
{
	static const int forced = !getenv("DWT_SYNC") ? 0 : (!strcmp(getenv("DWT_SYNC"), "spin") ? 1 : (!strcmp(getenv("DWT_SYNC"), "block") ? 2 : 0));
	static const long spin_us = getenv("DWT_SPIN_US") ? atol(getenv("DWT_SPIN_US")) : 100;
	const bool sleepy = forced ? forced == 2 : c->sleepy_wait;
	if (!sleepy || !c->sync_ev)
		return cudaStreamSynchronize(c->st);
	cudaError_t e = cudaEventRecord(c->sync_ev, c->st);
	if (e != cudaSuccess)
		return e;
	if (spin_us > 0) {
		const auto t0 = std::chrono::steady_clock::now();
		for (;;) {
			e = cudaEventQuery(c->sync_ev);
			if (e != cudaErrorNotReady)
				return e;
			if (std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= spin_us)
				break;
		}
	}
	return cudaEventSynchronize(c->sync_ev); // the event was created with cudaEventBlockingSync: the thread sleeps
}

cudaError_t ctx_copy(dwt_ctx *c, void *dst, const void *src, size_t n, cudaMemcpyKind kind, bool wait)
{
	static const bool ungated = getenv("DWT_XFER_GATE") && !strcmp(getenv("DWT_XFER_GATE"), "0");
	// below 16 MB a copy is over before the others notice (1080p frames measured 2-3 % slower gated: the extra wait costs more)
	if (!c->gate || ungated || n < (16u << 20)) {
		cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, c->st);
		return e == cudaSuccess && wait ? ctx_stream_sync(c) : e;
	}
	const int d = kind == cudaMemcpyDeviceToHost ? 1 : 0;
	cudaError_t e = cudaSuccess;
	if (d == 1 && (e = ctx_stream_sync(c)) != cudaSuccess) // the stream's kernels first: the chain only ever waits for copies
		return e;
	{
		std::lock_guard<std::mutex> hold(c->gate->dir[d]);
		if (c->gate->last[d] && c->gate->last[d] != c->xfer_ev[d])
			e = cudaStreamWaitEvent(c->st, c->gate->last[d], 0);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(dst, src, n, kind, c->st);
		if (e == cudaSuccess)
			e = cudaEventRecord(c->xfer_ev[d], c->st);
		if (e == cudaSuccess)
			c->gate->last[d] = c->xfer_ev[d];
	}
	return e == cudaSuccess && wait ? ctx_stream_sync(c) : e;
}

extern "C" void dwt_ctx_destroy(dwt_ctx *c)
{
	if (!c)
		return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->st);
	DevBuf *bufs[] = {&c->img, &c->pyr, &c->ll[0], &c->ll[1], &c->small, &c->bs, &c->sig, &c->ent, &c->Z,
	                  &c->signbuf, &c->specbuf, &c->refbuf, &c->tiles, &c->thr_state, &c->chunks, &c->info,
	                  &c->dsched, &c->out, &c->stream, &c->mem_pref, &c->ref_pref, &c->ones_rank, &c->sign_rank,
	                  &c->dstate, &c->win, &c->flush, &c->dec_scan, &c->dec_seg, &c->dec_chunks, &c->dec_lut};
	for (DevBuf *b : bufs)
		b->release();
	c->pin_small.release();
	c->pin_io.release();
	c->pin_stream.release();
	hilbert_plan_free(&c->plan);
	for (auto &e : c->ev)
		cudaEventDestroy(e);
	for (auto &e : c->xfer_ev)
		if (e)
			cudaEventDestroy(e);
	if (c->sync_ev)
		cudaEventDestroy(c->sync_ev);
	cudaStreamDestroy(c->st);
	delete c;
}

extern "C" long long dwt_ctx_launch_count(const dwt_ctx *c)
{
	return c ? c->launches : 0;
}

extern "C" int dwt_ctx_sync(dwt_ctx *c)
{
	CUDA_OK(cudaSetDevice(c->device));
	CUDA_OK(cudaStreamSynchronize(c->st));
	return 0;
}

extern "C" void dwt_free(void *p)
{
	free(p);
}

// ------------------------------------------------------------------------------------------------ geometry + schedule

int ctx_set_geometry(dwt_ctx *c, int w, int h, int ch)
{
	if (c->have_geom && c->geom.w[c->geom.levels] == w && c->geom.h[c->geom.levels] == h && c->geom.channels == ch)
		return 0;
	Geom &g = c->geom;
	memset(&g, 0, sizeof(g));
	int lengths[16], pixels[16], widths[16], heights[16];
	g.levels = compute_lengths(lengths, pixels, widths, heights, w, h, 8);
	g.channels = ch;
	for (int l = 0; l <= g.levels; ++l) {
		g.w[l] = widths[l];
		g.h[l] = heights[l];
		g.len[l] = lengths[l];
		g.pix[l] = (long long)widths[l] * heights[l];
	}
	int gb = 0, tb = 0;
	for (int l = 0; l < g.levels; ++l) {
		g.num[l] = g.pix[l + 1] - g.pix[l];
		g.G[l] = (int)((g.num[l] + 31) / 32);
		g.gbase[l] = gb;
		gb += g.G[l];
		g.ntile[l] = (g.G[l] + DWT_TILE_GROUPS - 1) / DWT_TILE_GROUPS;
		g.tbase[l] = tb;
		tb += g.ntile[l];
	}
	g.gbase[g.levels] = gb;
	g.GT = gb;
	g.tbase[g.levels] = tb;
	hilbert_plan_free(&c->plan);
	if (hilbert_plan_build(g, &c->plan, c->st, &c->launches))
		return -1;
	c->have_geom = true;
	return 0;
}

// chunk emission order of encode.c:183-221 (decode.c:187-243 walks the same order)
void build_schedule(const Geom &g, const int *planes, Sched *s)
{
	memset(s, 0, sizeof(*s));
	memset(s->chunk_of, 0xff, sizeof(s->chunk_of));
	int planes_max = 0;
	long long base = 0;
	for (int c = 0; c < g.channels; ++c) {
		s->planes[c] = planes[c];
		if (planes[c] > planes_max)
			planes_max = planes[c];
		s->bsbase[c] = base;
		base += (long long)(planes[c] + 1) * g.GT;
	}
	for (int c = g.channels; c < 4; ++c)
		s->bsbase[c] = base;
	int n = 0, e = 0;
	auto emit = [&](int c, int l, int p) {
		s->chan[n] = (short)c;
		s->level[n] = (short)l;
		s->plane[n] = (short)p;
		s->chunk_of[c][l][p] = (short)n;
		s->ebase[n] = e;
		e += g.ntile[l];
		++n;
	};
	if (planes_max > 0) {
		const int levels = g.levels;
		const int maximum = levels > planes_max ? levels : planes_max;
		const int layers_max = 2 * maximum - 1;
		if (planes_max == planes[0])
			emit(0, 0, planes[0] - 1);
		for (int layers = 0; layers < layers_max; ++layers) {
			for (int l = 0; l < levels && l <= layers + 1; ++l) {
				int p = planes_max - 1 - (layers + 1 - l);
				if (p >= 0 && p < planes[0])
					emit(0, l, p);
			}
			for (int l = 0; l < levels && l <= layers; ++l) {
				int p = planes_max - 1 - (layers - l);
				for (int c = 1; c < g.channels; ++c)
					if (p >= 0 && p < planes[c])
						emit(c, l, p);
			}
		}
	}
	s->nchunks = n;
	s->ebase[n] = e;
}

// ------------------------------------------------------------------------------------------------ transforms

int ensure_transform_buffers(dwt_ctx *c)
{
	const Geom &g = c->geom;
	const int L = g.levels;
	size_t full = (size_t)g.pix[L] * g.channels;
	if (c->pyr.ensure(full * sizeof(int)))
		return -1;
	size_t llsz = (size_t)g.pix[L - 1] * g.channels * sizeof(int);
	if (c->ll[0].ensure(llsz) || c->ll[1].ensure(llsz))
		return -1;
	if (c->small.ensure(1024))
		return -1;
	return 0;
}

int ctx_zero_transform_counters(dwt_ctx *c, bool forward)
{
	if (ensure_transform_buffers(c))
		return -1;
	// maxabs[4] at ints 0..3 and the work counters of the level launches at 64..95 (16..63 belong to the decoder)
	if (forward)
		CUDA_OK(cudaMemsetAsync(c->small.p, 0, 96 * sizeof(int), c->st));
	else
		CUDA_OK(cudaMemsetAsync(c->small.as<int>() + 64, 0, 32 * sizeof(int), c->st));
	return 0;
}

int ctx_forward_transform(dwt_ctx *c, const int *planar_in, bool counters_zeroed)
{
	const Geom &g = c->geom;
	const int L = g.levels;
	if (ensure_transform_buffers(c))
		return -1;
	if (!counters_zeroed && ctx_zero_transform_counters(c, true))
		return -1;
	int cur = 0;
	for (int lv = L; lv >= 1; --lv) {
		if (lv < L && lift_tail_fits(g.w[lv], g.h[lv])) {
			// the remaining levels in one launch, one CTA per channel (shared-memory resident)
			LiftTail t;
			t.ll_in = c->ll[cur ^ 1].as<int>();
			t.in_chan_stride = g.pix[lv];
			t.in_pitch = g.w[lv];
			t.ll_out = c->ll[cur].as<int>();
			t.out_chan_stride = g.pix[0];
			t.out_pitch = g.w[0];
			t.pyr = c->pyr.as<int>();
			t.pyr_chan_stride = g.pix[L];
			t.pyr_pitch = g.w[L];
			t.W = g.w[lv];
			t.H = g.h[lv];
			t.nlev = lv;
			t.channels = g.channels;
			t.maxabs = c->small.as<int>();
			t.chained = 1; // behind the level kernel that wrote ll_in
			if (lift_tail(t, false, c->st, &c->launches))
				return -1;
			c->root_buf = cur;
			return 0;
		}
		LiftLevel p;
		p.W = g.w[lv];
		p.H = g.h[lv];
		p.channels = g.channels;
		int mode;
		if (lv == L && !planar_in) {
			p.in = c->img.p;
			p.in_chan_stride = 0;
			p.in_pitch = g.w[L];
			mode = g.channels == 3 ? 0 : 1;
		} else if (lv == L) {
			p.in = planar_in;
			p.in_chan_stride = g.pix[L];
			p.in_pitch = g.w[L];
			mode = 2;
		} else {
			p.in = c->ll[cur ^ 1].p;
			p.in_chan_stride = g.pix[lv];
			p.in_pitch = g.w[lv];
			mode = 2;
		}
		p.out = c->ll[cur].p;
		p.out_chan_stride = g.pix[lv - 1];
		p.out_pitch = g.w[lv - 1];
		p.pyr = c->pyr.as<int>();
		p.pyr_chan_stride = g.pix[L];
		p.pyr_pitch = g.w[L];
		p.maxabs = c->small.as<int>();
		p.work = c->small.as<int>() + 64 + (L - lv);
		p.chained = lv != L;
		if (lift_forward_level(p, mode, c->st, &c->launches))
			return -1;
		c->root_buf = cur;
		cur ^= 1;
	}
	return 0;
}

const int *ctx_root_ll(dwt_ctx *c)
{
	// after ctx_forward_transform the last written LL buffer holds the root (planar, pitch w[0])
	return c->ll[c->root_buf].as<int>();
}

// ------------------------------------------------------------------------------------------------ encode

extern "C" int dwt_ctx_upload_image(dwt_ctx *c, const uint8_t *pixels, int width, int height, int channels)
{
	if (!c || !pixels || width < 8 || height < 8 || width > 65536 || height > 65536 ||
	    (channels != 1 && channels != 3)) {
		dwt_set_error("bad image arguments (%dx%dx%d)", width, height, channels);
		return -1; // encode.c:140-146
	}
	if ((long long)width * height > 0x7fffffffLL / 4) {
		dwt_set_error("image too large for the reference's int arithmetic");
		return -1;
	}
	CUDA_OK(cudaSetDevice(c->device));
	size_t n = (size_t)width * height * channels;
	if (c->img.ensure(n + 16))
		return -1;
	CUDA_OK(ctx_copy(c, c->img.p, pixels, n, cudaMemcpyHostToDevice, false));
	c->img_w = width;
	c->img_h = height;
	c->img_ch = channels;
	c->img_resident = true;
	return 0;
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
		cudaGetLastError();
		return 0.f;
	}
	return ms;
}

static inline size_t round_up(size_t v, size_t a)
{
	return (v + a - 1) / a * a;
}

// the host-side writers of the stream prefix: released on every return path
struct PrefixWriters {
	struct bytes_writer *bw = nullptr;
	struct bits_writer *bits = nullptr;
	struct vli_writer *vli = nullptr;
	void finish() // pads the last byte into bw; bw stays readable until the guard goes out of scope
	{
		if (vli)
			delete_vli_writer(vli);
		if (bits)
			close_bits_writer(bits);
		vli = nullptr;
		bits = nullptr;
	}
	~PrefixWriters()
	{
		finish();
		if (bw)
			close_bytes_writer(bw);
	}
};

// header, root image and plane counts through the reference-shaped writers, return values ignored like encode.c:169-182
// does.  capacity > 0 makes the byte sink refuse like bytes.h:77-78 (used to reproduce the stderr counters of a cut that
// lands inside the prefix: every write call then stops at its first refused byte, so the counts are not a closed form).
static void write_prefix(PrefixWriters &pw, int capacity, const Geom &g, const int *root, const int *planes, long long *meta_bits,
                         long long *root_end)
{
	const int C = g.channels, L = g.levels;
	struct bytes_writer *bw = pw.bw = bytes_writer_mem(capacity);
	put_byte(bw, 'W');
	put_byte(bw, C == 3 ? '6' : '5');
	write_bytes(bw, g.w[L] - 1, 2);
	write_bytes(bw, g.h[L] - 1, 2);
	struct bits_writer *bits = pw.bits = bits_writer(bw);
	struct vli_writer *vli = pw.vli = vli_writer(bits);
	*meta_bits = bits_count(bits);
	for (int ch = 0; ch < C; ++ch) {
		const int *v = root + (size_t)ch * g.pix[0];
		int mx = 0;
		for (int i = 0; i < (int)g.pix[0]; ++i)
			if (abs(v[i]) > mx)
				mx = abs(v[i]);
		int cnt = 1 + ilog2(mx);
		put_vli(vli, cnt);
		for (int i = 0; cnt && i < (int)g.pix[0]; ++i) {
			vli_write_bits(vli, abs(v[i]), cnt);
			if (v[i])
				vli_put_bit(vli, v[i] < 0);
		}
	}
	*root_end = bits_count(bits);
	for (int ch = 0; ch < C; ++ch)
		put_vli(vli, planes[ch]);
}

extern "C" int dwt_ctx_encode_resident(dwt_ctx *c, int capacity, struct dwt_stats *stt)
{
	if (!c || !c->img_resident) {
		dwt_set_error("no image uploaded");
		return -1;
	}
	CUDA_OK(cudaSetDevice(c->device));
	if (ctx_set_geometry(c, c->img_w, c->img_h, c->img_ch))
		return -1;
	const Geom &g = c->geom;
	const int C = g.channels, L = g.levels;
	cudaStream_t st = c->st;
	if (ctx_zero_transform_counters(c, true))
		return -1;
	CUDA_OK(cudaEventRecord(c->ev[0], st));
	if (ctx_forward_transform(c, nullptr, true))
		return -1;

	// ---- plane counts and root image come back to the host (a few hundred bytes)
	const int nroot = (int)g.pix[0] * C;
	if (c->pin_small.ensure(sizeof(int) * (4 + nroot) + sizeof(EncInfo) + 64))
		return -1;
	int *h_small = c->pin_small.as<int>();
	CUDA_OK(cudaEventRecord(c->ev[1], st)); // ev0..ev1 = the lifting kernels only
	CUDA_OK(cudaMemcpyAsync(h_small, c->small.p, 16, cudaMemcpyDeviceToHost, st));
	CUDA_OK(cudaMemcpyAsync(h_small + 4, ctx_root_ll(c), sizeof(int) * nroot, cudaMemcpyDeviceToHost, st));
	CUDA_OK(ctx_stream_sync(c));
	int planes[3] = {0, 0, 0}, planes_max = 0;
	for (int ch = 0; ch < C; ++ch) {
		planes[ch] = 1 + ilog2(h_small[ch]); // encode.c:130
		if (planes[ch] > planes_max)
			planes_max = planes[ch];
		if (planes[ch] > DWT_MAX_PLANES - 1) {
			dwt_set_error("coefficient magnitude beyond the reference's 29-bit range");
			return -1;
		}
	}
	build_schedule(g, planes, &c->sched);
	const Sched &S = c->sched;

	// ---- stream prefix on the host: header (encode.c:169-172), root image (encode.c:97-110), planes (encode.c:181-182)
	PrefixWriters pw;
	const int *root = h_small + 4;
	long long meta_bits = 0, root_end = 0;
	write_prefix(pw, 0, g, root, planes, &meta_bits, &root_end);
	struct bits_writer *bits = pw.bits;
	struct vli_writer *vli = pw.vli;
	struct bytes_writer *bw = pw.bw;
	u64 prefix_bits = (u64)bits_count(bits);
	int k0 = dwt_vli_writer_order(vli);
	u64 total_bits = 0;
	if (planes_max == 0) {
		// all detail coefficients are zero: the reference codes one all-zero pseudo
		// plane of the coarsest luma level and flushes the run; the payload is that single VLI.
		struct rle_writer *rle = rle_writer(vli);
		for (long long i = 0; i < g.num[0]; ++i)
			put_rle(rle, 0);
		rle_flush(rle);
		delete_rle_writer(rle);
		total_bits = (u64)bits_count(bits);
		prefix_bits = total_bits;
	}
	pw.finish(); // pads the last byte
	size_t prefix_len = 0;
	const uint8_t *prefix = bytes_writer_data(bw, &prefix_len);

	u64 tot_ref = 0;
	if (planes_max > 0) {
		// ---- bit-sliced store + linearisation
		const size_t bs_words = (size_t)S.bsbase[C];
		if (c->bs.ensure(bs_words * 4 + 64))
			return -1;
		CUDA_OK(cudaMemsetAsync(c->bs.p, 0, bs_words * 4, st));
		if (c->dsched.ensure(sizeof(Sched)))
			return -1;
		CUDA_OK(cudaMemcpyAsync(c->dsched.p, &c->sched, sizeof(Sched), cudaMemcpyHostToDevice, st));
		if (hilbert_linearize(g, c->plan, S, c->pyr.as<int>(), g.pix[L], g.w[L], c->bs.as<u32>(), L, st, &c->launches))
			return -1;
		CUDA_OK(cudaEventRecord(c->ev[2], st));

		// ---- coder working set
		const long long ndet = g.pix[L] - g.pix[0];
		const long long max_tok_ll = ndet * C + S.nchunks + 1;
		if (max_tok_ll >= 0xfff00000LL) {
			dwt_set_error("image too large: more than 2^32 tokens possible");
			return -1;
		}
		EncBuffers b;
		memset(&b, 0, sizeof(b));
		// a capacity bounds how many tokens and refinement bits can reach the output (enc_chunk_setup_kernel cuts the rest):
		// the buffers that are zeroed per frame and the token-tile grids are sized by what can survive
		const u64 limit_bits = capacity > 0 ? 8ull * (u64)capacity : 0ull;
		const u64 tok_cap = enc_token_bound(g, S, prefix_bits, limit_bits);
		b.max_tokens = (u32)((u64)max_tok_ll < tok_cap ? (u64)max_tok_ll : tok_cap);
		const size_t tok_room = round_up(b.max_tokens, DWT_TOK_TILE) + 32;
		const size_t ntile_max = tok_room / DWT_TOK_TILE + 1;
		b.nent = S.ebase[S.nchunks];
		long long ref_bound = 0;
		for (int ch = 0; ch < C; ++ch)
			ref_bound += ndet * (planes[ch] > 1 ? planes[ch] - 1 : 0);
		const u64 ref_cap = enc_ref_bound(g, S, prefix_bits, limit_bits);
		const size_t ref_words = (size_t)(((u64)ref_bound < ref_cap ? (u64)ref_bound : ref_cap) / 32) + 4;
		const size_t bit_words = tok_room / 32 + 4;
		if (c->ent.ensure((size_t)b.nent * 12 + 64 + (size_t)(b.nent / 4096 + 2) * 24) || // + block totals of the scan
		    c->Z.ensure(tok_room * 4 + 64) ||
		    c->signbuf.ensure(bit_words * 4) || c->specbuf.ensure(bit_words * 4) || c->refbuf.ensure(ref_words * 4) ||
		    c->tiles.ensure(ntile_max * 8 + (ntile_max + 8) * 16 + 512) || c->thr_state.ensure(ntile_max * 256) ||
		    c->chunks.ensure(sizeof(EncChunks)) || c->info.ensure(sizeof(EncInfo)))
			return -1;
		CUDA_OK(cudaMemsetAsync(c->signbuf.p, 0, bit_words * 4, st));
		CUDA_OK(cudaMemsetAsync(c->specbuf.p, 0, bit_words * 4, st));
		CUDA_OK(cudaMemsetAsync(c->refbuf.p, 0, ref_words * 4, st));
		b.bs = c->bs.as<u32>();
		b.ent_z = c->ent.as<u32>();
		b.ent_1 = b.ent_z + b.nent;
		b.ent_r = b.ent_1 + b.nent;
		b.Z = c->Z.as<u32>();
		b.signbuf = c->signbuf.as<u32>();
		b.specbuf = c->specbuf.as<u32>();
		b.refbuf = c->refbuf.as<u32>();
		b.tile_bitbase = c->tiles.as<u64>();
		b.tile_flags = (u32 *)(b.tile_bitbase + ntile_max); // 16-byte aligned: ntile_max * 8 bytes in front
		b.tile_flag_bytes = (ntile_max + 8) * 16 + 64;
		b.thr_state = c->thr_state.as<unsigned char>();
		b.chunks = c->chunks.as<EncChunks>();
		b.info = c->info.as<EncInfo>();
		b.sched = c->dsched.as<Sched>();
		if (enc_count(g, S, b, st, &c->launches) || enc_scan_and_setup(g, S, b, prefix_bits, limit_bits, st, &c->launches) ||
		    enc_emit(g, S, b, st, &c->launches) || enc_vli_orders(b, k0, st, &c->launches))
			return -1;
		EncInfo *h_info = (EncInfo *)(h_small + 4 + nroot + 2);
		h_info = (EncInfo *)(((uintptr_t)h_info + 15) & ~(uintptr_t)15);
		CUDA_OK(cudaMemcpyAsync(h_info, b.info, sizeof(EncInfo), cudaMemcpyDeviceToHost, st));
		CUDA_OK(ctx_stream_sync(c));
		if (h_info->error) {
			dwt_set_error("coder: input outside the supported range (code %d)", h_info->error);
			return -1;
		}
		// chunks that provably start behind the capacity were not coded (enc_chunk_setup_kernel): total_bits is then a
		// lower bound of the untruncated length that still lies >= 64 bits behind the capacity
		tot_ref = h_info->jcut < S.nchunks ? h_info->ref_cut : h_info->tot_ref;
		total_bits = prefix_bits + h_info->tok_bits + tot_ref;

		// ---- output stream: zero, prefix bytes, then scatter
		const size_t full_bytes = (size_t)((total_bits + 7) / 8);
		const size_t out_bytes = capacity > 0 && (size_t)capacity < full_bytes ? (size_t)capacity : full_bytes;
		const size_t out_room = round_up(out_bytes, 4) + 64;
		if (c->out.ensure(out_room))
			return -1;
		CUDA_OK(cudaMemsetAsync(c->out.p, 0, out_room, st));
		if (c->pin_io.ensure(prefix_len + 16))
			return -1;
		memcpy(c->pin_io.p, prefix, prefix_len);
		size_t pre_copy = prefix_len < out_bytes + 8 ? prefix_len : out_bytes + 8;
		CUDA_OK(cudaMemcpyAsync(c->out.p, c->pin_io.p, pre_copy, cudaMemcpyHostToDevice, st));
		b.out = c->out.as<u32>();
		b.out_limit_bits = (u64)out_bytes * 8;
		if (enc_scatter(S, b, prefix_bits, tot_ref, st, &c->launches))
			return -1;
		c->out_bytes = out_bytes;
	} else {
		CUDA_OK(cudaEventRecord(c->ev[2], st));
		const size_t full_bytes = prefix_len;
		const size_t out_bytes = capacity > 0 && (size_t)capacity < full_bytes ? (size_t)capacity : full_bytes;
		if (c->out.ensure(round_up(full_bytes, 4) + 64) || c->pin_io.ensure(prefix_len + 16))
			return -1;
		memcpy(c->pin_io.p, prefix, prefix_len);
		CUDA_OK(cudaMemcpyAsync(c->out.p, c->pin_io.p, prefix_len, cudaMemcpyHostToDevice, st));
		c->out_bytes = out_bytes;
	}
	CUDA_OK(cudaEventRecord(c->ev[3], st));
	// the stage timers need the stream drained; a caller without stats (the pool) waits once, behind its download
	if (stt)
		CUDA_OK(ctx_stream_sync(c));

	if (stt) {
		memset(stt, 0, sizeof(*stt));
		const long long full = (long long)total_bits;
		stt->full_bits = full;
		stt->levels = L;
		for (int ch = 0; ch < 3; ++ch)
			stt->planes[ch] = planes[ch];
		// the three stderr counters of encode.c:175-180,226-230.  bits_count() = bits->cnt + 8 * bytes_count() (bits.h:41-44):
		// once the byte sink refuses (bytes.h:77-78) the byte count stays at the capacity, and every prefix write call stops
		// at its first refused byte -- the counters of a cut inside the prefix come from a replay against a capped sink
		const long long cap8 = 8LL * capacity;
		stt->meta_bits = meta_bits;
		stt->root_bits = root_end - meta_bits;
		if (capacity > 0 && cap8 < (long long)prefix_bits) { // the cut lands inside the prefix: replay it against a capped sink
			PrefixWriters sim;
			long long m = 0, r = 0;
			write_prefix(sim, capacity, g, root, planes, &m, &r);
			stt->meta_bits = m;
			stt->root_bits = r - m;
		}
		const bool cut = capacity > 0 && full >= 8LL * (capacity + 1); // a payload byte was refused: `goto end` at a byte boundary
		stt->total_bits = cut ? cap8 : full;
		long long byte_cnt = (full + 7) / 8; // bytes_count() after close_bits_writer() wrote the padded last byte
		if (capacity > 0 && byte_cnt > capacity)
			byte_cnt = capacity;
		stt->kib = (byte_cnt + 512) / 1024;
		stt->ms_lift = ev_ms(c->ev[0], c->ev[1]);
		stt->ms_linearize = ev_ms(c->ev[1], c->ev[2]);
		stt->ms_coder = ev_ms(c->ev[2], c->ev[3]);
		stt->ms_total = ev_ms(c->ev[0], c->ev[3]);
	}
	return 0;
}

extern "C" int dwt_ctx_download_stream(dwt_ctx *c, uint8_t **out, size_t *out_len)
{
	CUDA_OK(cudaSetDevice(c->device));
	uint8_t *buf = (uint8_t *)malloc(c->out_bytes ? c->out_bytes : 1);
	if (!buf) {
		dwt_set_error("out of host memory");
		return -1;
	}
	if (c->out_bytes) {
		CUDA_OK(cudaMemcpyAsync(buf, c->out.p, c->out_bytes, cudaMemcpyDeviceToHost, c->st));
		CUDA_OK(ctx_stream_sync(c));
	}
	*out = buf;
	*out_len = c->out_bytes;
	return 0;
}

extern "C" int dwt_encode(dwt_ctx *c, const uint8_t *pixels, int width, int height, int channels, int capacity,
                          uint8_t **out, size_t *out_len, struct dwt_stats *stats)
{
	if (!c) {
		dwt_set_error("null context (no CUDA device?)");
		return -1;
	}
	if (dwt_ctx_upload_image(c, pixels, width, height, channels))
		return -1;
	if (dwt_ctx_encode_resident(c, capacity, stats))
		return -1;
	return dwt_ctx_download_stream(c, out, out_len);
}

// ------------------------------------------------------------------------------------------------ debug taps

extern "C" int dwt_debug_front_end(dwt_ctx *c, const uint8_t *pixels, int width, int height, int channels,
                                   int *pyramid, int *planar, int *planes_out)
{
	if (dwt_ctx_upload_image(c, pixels, width, height, channels))
		return -1;
	if (ctx_set_geometry(c, width, height, channels))
		return -1;
	const Geom &g = c->geom;
	const int C = g.channels, L = g.levels;
	cudaStream_t st = c->st;
	if (ctx_forward_transform(c, nullptr))
		return -1;
	int h_max[4];
	CUDA_OK(cudaMemcpyAsync(h_max, c->small.p, 16, cudaMemcpyDeviceToHost, st));
	CUDA_OK(ctx_stream_sync(c));
	int planes[3] = {0, 0, 0};
	for (int ch = 0; ch < C; ++ch)
		planes[ch] = 1 + ilog2(h_max[ch]);
	for (int ch = 0; ch < 3; ++ch)
		planes_out[ch] = planes[ch];
	const long long npix = g.pix[L];
	const size_t n = (size_t)npix * C;
	struct TmpGuard { // released on every return path
		int *p = nullptr;
		~TmpGuard()
		{
			if (p)
				cudaFree(p);
		}
	} guard;
	CUDA_OK(cudaMalloc(&guard.p, n * sizeof(int)));
	int *tmp = guard.p;
	if (pyramid) {
		export_pyramid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->pyr.as<int>(), ctx_root_ll(c), tmp, g.w[L],
		                                                                   g.h[L], C, g.w[0], g.h[0]);
		CUDA_OK(cudaMemcpyAsync(pyramid, tmp, n * sizeof(int), cudaMemcpyDeviceToHost, st));
		CUDA_OK(ctx_stream_sync(c));
	}
	if (planar) {
		build_schedule(g, planes, &c->sched);
		const Sched &S = c->sched;
		const size_t bs_words = (size_t)S.bsbase[C];
		if (c->bs.ensure(bs_words * 4 + 64))
			return -1;
		CUDA_OK(cudaMemsetAsync(c->bs.p, 0, bs_words * 4, st));
		CUDA_OK(cudaMemsetAsync(tmp, 0, n * sizeof(int), st));
		if (hilbert_linearize(g, c->plan, S, c->pyr.as<int>(), npix, g.w[L], c->bs.as<u32>(), L, st, &c->launches) ||
		    hilbert_unslice(g, S, c->bs.as<u32>(), tmp, npix, st))
			return -1;
		// root raster first (encode.c:37-45)
		for (int ch = 0; ch < C; ++ch)
			CUDA_OK(cudaMemcpyAsync(tmp + (size_t)ch * npix, ctx_root_ll(c) + (size_t)ch * g.pix[0],
			                        sizeof(int) * g.pix[0], cudaMemcpyDeviceToDevice, st));
		CUDA_OK(cudaMemcpyAsync(planar, tmp, n * sizeof(int), cudaMemcpyDeviceToHost, st));
		CUDA_OK(ctx_stream_sync(c));
	}
	return 0;
}
