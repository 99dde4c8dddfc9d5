// pipeline_dec.cu -- decode side of the C ABI (include/dwt_b200.h).
//
// Host orchestration mirrors main() of decode.c:136-268: magic + size (decode.c:145-159), geometry, the
// optional PIXELS limit (decode.c:165-171), root image and plane counts (decode.c:119-134,180-186; host,
// through the reference-shaped stream entry points), then the chunk schedule of decode.c:187-243 on the GPU
// with a device-resident cursor (bit position, Rice order, pending run, missing[], level), and finally
// reconstruction + inverse lifting + inverse colour for the resolution the stream reached (decode.c:249-263).
#include "pipeline.cuh"

#include "../host/streamio_internal.h"
#include "dwt_b200.h"

#include <stdlib.h>
#include <string.h>

#include <vector>

static inline size_t round_up(size_t v, size_t a)
{
	return (v + a - 1) / a * a;
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

// wait = false: the caller keeps `stream` alive and unchanged until it has waited for the context itself
int ctx_upload_stream(dwt_ctx *c, const uint8_t *stream, size_t len, bool wait)
{
	if (!c) {
		dwt_set_error("null context");
		return -1;
	}
	CUDA_OK(cudaSetDevice(c->device));
	const size_t room = round_up(len, 4) + 64;
	// the host only parses the stream prefix (header, root image < 16x16, plane counts: a few KB)
	const size_t head = len < (1u << 20) ? len : (1u << 20);
	if (c->stream.ensure(room) || c->pin_stream.ensure(head + 16))
		return -1;
	// the parse peeks up to 12 bytes past its position: keep the tail zeroed
	CUDA_OK(cudaMemsetAsync((char *)c->stream.p + (len / 4) * 4, 0, room - (len / 4) * 4, c->st));
	if (len) {
		memcpy(c->pin_stream.p, stream, head);
		CUDA_OK(ctx_copy(c, c->stream.p, stream, len, cudaMemcpyHostToDevice, wait));
	}
	c->stream_head = head;
	c->stream_len = len;
	c->stream_resident = true;
	return 0;
}

extern "C" int dwt_ctx_upload_stream(dwt_ctx *c, const uint8_t *stream, size_t len)
{
	return ctx_upload_stream(c, stream, len, true); // the caller's buffer is free again when this returns
}

// chunk list of decode.c:199-243: the encoder's order, cut where a level loop reaches l >= levels_max
static int decode_schedule_length(const Geom &g, const Sched &S, int levels_max)
{
	if (levels_max >= g.levels)
		return S.nchunks;
	if (levels_max <= 0)
		return 0;
	// walk the same loops and count the chunks emitted before the first l >= levels_max
	int planes_max = 0;
	for (int c = 0; c < g.channels; ++c)
		if (S.planes[c] > planes_max)
			planes_max = S.planes[c];
	if (planes_max == 0)
		return 0;
	const int levels = g.levels;
	const int maximum = levels > planes_max ? levels : planes_max;
	const int layers_max = 2 * maximum - 1;
	int n = 0;
	if (planes_max == S.planes[0])
		++n;
	for (int layers = 0; layers < layers_max; ++layers) {
		for (int l = 0; l < levels && l <= layers + 1; ++l) {
			if (l >= levels_max)
				return n;
			int p = planes_max - 1 - (layers + 1 - l);
			if (p >= 0 && p < S.planes[0])
				++n;
		}
		for (int l = 0; l < levels && l <= layers; ++l) {
			if (l >= levels_max)
				return n;
			int p = planes_max - 1 - (layers - l);
			for (int c = 1; c < g.channels; ++c)
				if (p >= 0 && p < S.planes[c])
					++n;
		}
	}
	return n;
}

// the host-side readers of the stream prefix: released on every return path
struct PrefixReaders {
	struct bytes_reader *br = nullptr;
	struct bits_reader *bits = nullptr;
	struct vli_reader *vli = nullptr;
	~PrefixReaders()
	{
		if (vli)
			delete_vli_reader(vli);
		if (bits)
			close_bits_reader(bits);
		if (br)
			close_bytes_reader(br);
	}
};

// order-0 token table of the decoder: built once per process, immutable afterwards
static const u32 *dec_token_table_host()
{
	static const std::vector<u32> table = [] {
		std::vector<u32> t(DWT_DEC_LUT_WORDS);
		dec_token_table(t.data());
		return t;
	}();
	return table.data();
}

extern "C" int dwt_ctx_decode_resident(dwt_ctx *c, int pixels_max, struct dwt_stats *stt)
{
	if (!c || !c->stream_resident) {
		dwt_set_error("no stream uploaded");
		return -1;
	}
	CUDA_OK(cudaSetDevice(c->device));
	cudaStream_t st = c->st;
	// ---- header + root image + plane counts on the host (the bytes are still in the pinned staging buffer)
	const uint8_t *bytes = c->pin_stream.as<uint8_t>();
	const size_t len = c->stream_len;
	PrefixReaders pr;
	struct bytes_reader *br = pr.br = bytes_reader_mem(bytes, c->stream_head); // prefix only (see dwt_ctx_upload_stream)
	// decode.c:145-159: the reference exits 1 either silently (wrong magic, size below 8: return 2 here) or after get_byte
	// printed "reached end of file" (bytes.h:99-103: return 1 here)
	int width = 0, height = 0;
	const int letter = get_byte(br);
	if (letter < 0)
		return 1;
	if (letter != 'W')
		return 2;
	const int number = get_byte(br);
	if (number < 0)
		return 1;
	if (number != '5' && number != '6')
		return 2;
	if (read_bytes(br, &width, 2) || read_bytes(br, &height, 2))
		return 1;
	++width;
	++height;
	if (width < 8 || height < 8)
		return 2;
	const int C = number == '6' ? 3 : 1;
	if (ctx_set_geometry(c, width, height, C))
		return -1;
	const Geom &g = c->geom;
	const int L = g.levels;
	int levels_max = L;
	if (pixels_max >= 0)
		while (levels_max > 0 && g.pix[levels_max] > pixels_max)
			--levels_max; // decode.c:165-171
	struct bits_reader *bits = pr.bits = bits_reader(br);
	struct vli_reader *vli = pr.vli = vli_reader(bits);
	const int nroot = (int)g.pix[0];
	if (c->pin_small.ensure(sizeof(int) * (size_t)(nroot * C + 64) + sizeof(DecState) + 64))
		return -1;
	int *h_root = c->pin_small.as<int>();
	memset(h_root, 0, sizeof(int) * (size_t)nroot * C);
	int planes[3] = {0, 0, 0};
	bool bad = false;
	for (int ch = 0; ch < C && !bad; ++ch) { // decode_root decode.c:119-134
		int cnt = get_vli(vli);
		if (cnt < 0) {
			bad = true;
			break;
		}
		for (int i = 0; cnt && i < nroot; ++i) {
			int v = 0, r = 0;
			if (vli_read_bits(vli, &v, cnt)) {
				bad = true;
				break;
			}
			if (v && (r = vli_get_bit(vli)) > 0)
				v = -v;
			if (r < 0) {
				bad = true;
				break;
			}
			h_root[ch * nroot + i] = v;
		}
	}
	for (int ch = 0; ch < C && !bad; ++ch)
		if ((planes[ch] = get_vli(vli)) < 0)
			bad = true; // decode.c:183-186
	const int k0 = dwt_vli_reader_order(vli);
	const long long b0 = dwt_bits_reader_position(bits);
	if (bad)
		return 1; // EOF inside the root image or the plane counts (decode.c:180-186)
	int planes_max = 0;
	for (int ch = 0; ch < C; ++ch) {
		if (planes[ch] > planes_max)
			planes_max = planes[ch];
		if (planes[ch] > DWT_MAX_PLANES - 1) {
			dwt_set_error("stream declares %d bit planes: outside the reference's range", planes[ch]);
			return -1;
		}
	}
	CUDA_OK(cudaEventRecord(c->ev[0], st));
	build_schedule(g, planes, &c->sched);
	const Sched &S = c->sched;
	const int nchunks = decode_schedule_length(g, S, levels_max);

	// ---- device state
	DecState *h_state = (DecState *)(((uintptr_t)(h_root + nroot * C + 8) + 15) & ~(uintptr_t)15);
	memset(h_state, 0, sizeof(DecState));
	h_state->bitpos = (u64)b0;
	h_state->end_bits = (u64)len * 8;
	h_state->order = k0;
	h_state->level = -1;
	for (int ch = 0; ch < C; ++ch)
		for (int l = 0; l < L; ++l)
			h_state->missing[ch * 16 + l] = planes[ch];
	int level = -1;
	if (planes_max == 0) {
		// all-zero detail: the reference decodes one pseudo plane of the coarsest luma
		// level; whatever the stream holds there, every coefficient stays zero and `level` becomes 0
		if (levels_max > 0)
			level = 0;
	} else if (nchunks > 0) {
		const size_t bs_words = (size_t)S.bsbase[C];
		const size_t sig_words = (size_t)g.GT * C;
		const size_t ntiles = (size_t)g.tbase[L] * C;
		const u64 end_bits = (u64)len * 8;
		const size_t nwin = (size_t)((end_bits / 64 + 2 + DWT_DEC_WS - 1) / DWT_DEC_WS);
		const size_t nslice = nwin * DWT_DEC_WS;
		const size_t rank_words = (size_t)(dec_rank_bits(g, S, nchunks) / 32) + 8;
		// per slice: E 4 B + P 16 B + TK 4 B; per window: X 4 B + PT 16 B + TT 4 B + two link records
		const size_t nsuper = (nwin + DWT_DEC_SUPER - 1) / DWT_DEC_SUPER;
		const size_t scan_bytes = nslice * 24 + nwin * (24 + 2 * sizeof(DecLink)) + nsuper * 2 * sizeof(DecSuper) +
		                          (nsuper + (size_t)nchunks + 8) * sizeof(DecBulk) + nwin * 4 + round_up(2 * nwin, 16) + nwin * 8 + 64 + 16 + 256;
		if (c->bs.ensure(bs_words * 4 + 64) || c->sig.ensure(sig_words * 4 + 64) || c->dstate.ensure(sizeof(DecState)) ||
		    c->mem_pref.ensure(ntiles * 8 + 64) || c->ref_pref.ensure(ntiles * 8 + 64) ||
		    c->ones_rank.ensure(rank_words * 8 + 64) || c->dec_scan.ensure(scan_bytes) ||
		    c->dec_seg.ensure((nwin + 2 * (size_t)nchunks + 64) * sizeof(DecSeg)) ||
		    c->dec_chunks.ensure((size_t)DWT_MAX_CHUNKS * sizeof(DecChunk)) || c->dsched.ensure(sizeof(Sched)))
			return -1;
		if (!c->dec_lut_ready) {
			const size_t table_bytes = sizeof(u32) * DWT_DEC_LUT_WORDS;
			if (c->dec_lut.ensure(table_bytes))
				return -1;
			CUDA_OK(cudaMemcpyAsync(c->dec_lut.p, dec_token_table_host(), table_bytes, cudaMemcpyHostToDevice, st));
			CUDA_OK(ctx_stream_sync(c));
			c->dec_lut_ready = true;
		}
		CUDA_OK(cudaMemsetAsync(c->bs.p, 0, bs_words * 4, st));
		CUDA_OK(cudaMemsetAsync(c->sig.p, 0, sig_words * 4, st));
		CUDA_OK(cudaMemsetAsync(c->ones_rank.p, 0, rank_words * 8, st));
		CUDA_OK(cudaMemsetAsync(c->dec_chunks.p, 0, (size_t)DWT_MAX_CHUNKS * sizeof(DecChunk), st));
		CUDA_OK(cudaMemcpyAsync(c->dstate.p, h_state, sizeof(DecState), cudaMemcpyHostToDevice, st));
		CUDA_OK(cudaMemcpyAsync(c->dsched.p, &c->sched, sizeof(Sched), cudaMemcpyHostToDevice, st));
		DecBuffers b;
		b.bs = c->bs.as<u32>();
		b.sig = c->sig.as<u32>();
		b.stream = c->stream.as<u32>();
		b.toklut = c->dec_lut.as<u32>();
		b.end_bits = end_bits;
		b.nwin = (u32)nwin;
		b.in_flight = c->in_flight;
		b.scan_mode = c->scan_mode;
		char *sp = c->dec_scan.as<char>();
		b.P = (ulonglong2 *)sp;
		sp += nslice * 16;
		b.winPT = (ulonglong2 *)sp;
		sp += nwin * 16;
		b.link = (DecLink *)sp;
		sp += nwin * 2 * sizeof(DecLink);
		b.super = (DecSuper *)sp;
		sp += nsuper * 2 * sizeof(DecSuper);
		b.nsuper = (u32)nsuper;
		b.bulk = (DecBulk *)sp;
		sp += (nsuper + (size_t)nchunks + 8) * sizeof(DecBulk);
		b.E = (u32 *)sp;
		sp += nslice * 4;
		b.TK = (u32 *)sp;
		sp += nslice * 4;
		b.winX = (u32 *)sp;
		sp += nwin * 4;
		b.winTT = (u32 *)sp;
		sp += nwin * 4;
		b.winX2 = (u32 *)sp;
		sp += nwin * 4;
		b.chg = (unsigned char *)sp;
		sp += round_up(2 * nwin, 16);
		sp = (char *)(((uintptr_t)sp + 15) & ~(uintptr_t)15); // the u32 arrays above leave it 4-byte aligned
		b.ext_list = (uint2 *)sp;
		sp += nwin * 8;
		b.ext_count = (u32 *)sp;
		CUDA_OK(cudaMemsetAsync(b.ext_count, 0, 64, st));
		b.seg = c->dec_seg.as<DecSeg>();
		b.chunks = c->dec_chunks.as<DecChunk>();
		b.tile_sums = c->mem_pref.as<u32>();
		b.tile_base = c->ref_pref.as<u32>();
		b.ones_rank = c->ones_rank.as<u32>();
		b.sign_rank = b.ones_rank + rank_words;
		b.state = c->dstate.as<DecState>();
		b.sched = c->dsched.as<Sched>();
		if (dec_run(g, S, b, nchunks, st, &c->launches))
			return -1;
		CUDA_OK(cudaMemcpyAsync(h_state, c->dstate.p, sizeof(DecState), cudaMemcpyDeviceToHost, st));
		CUDA_OK(cudaEventRecord(c->ev[1], st));
		CUDA_OK(ctx_stream_sync(c));
		if (getenv("DWT_DEBUG")) { // windows each lineage pass walked
			u32 cnt[16];
			if (cudaMemcpy(cnt, b.ext_count, sizeof(cnt), cudaMemcpyDeviceToHost) == cudaSuccess) {
				fprintf(stderr, "lineage: windows walked per pass:");
				for (int t = 0; t < 16; ++t)
					fprintf(stderr, " %u", cnt[t]);
				fprintf(stderr, " (of %zu)\n", nwin);
			}
		}
		if (h_state->guard_tripped) {
			dwt_set_error("decoder: the resolver's iteration guard fired (code %d): internal error", h_state->guard_tripped);
			return -1;
		}
		level = h_state->level;
	}
	if (planes_max == 0 || nchunks == 0)
		CUDA_OK(cudaEventRecord(c->ev[1], st));

	// ---- reconstruction for the resolution reached (decode.c:249-263)
	const int levels_used = level + 1;
	const int ow = g.w[levels_used], oh = g.h[levels_used];
	if (ensure_transform_buffers(c))
		return -1;
	if (c->img.ensure((size_t)ow * oh * C + 16))
		return -1;
	CUDA_OK(cudaMemcpyAsync(c->ll[0].p, h_root, sizeof(int) * (size_t)nroot * C, cudaMemcpyHostToDevice, st));
	if (ctx_zero_transform_counters(c, false))
		return -1;
	if (levels_used > 0) {
		int *d_missing = c->small.as<int>() + 16;
		CUDA_OK(cudaMemcpyAsync(d_missing, h_state->missing, sizeof(int) * 48, cudaMemcpyHostToDevice, st));
		if (planes_max == 0) {
			// no coefficient was coded: the detail bands of the levels used are all zero
			CUDA_OK(cudaMemsetAsync(c->pyr.p, 0, sizeof(int) * (size_t)g.pix[levels_used] * C, st));
		} else if (hilbert_reconstruct(g, c->plan, S, c->bs.as<u32>(), d_missing, c->pyr.as<int>(), g.pix[levels_used], ow,
		                               levels_used, st, &c->launches)) {
			return -1;
		}
	}
	CUDA_OK(cudaEventRecord(c->ev[2], st));
	if (ctx_inverse_transform(c, levels_used, true, nullptr, true))
		return -1;
	CUDA_OK(cudaEventRecord(c->ev[3], st));
	if (stt) // stage timers; a caller without stats waits once, behind its download
		CUDA_OK(ctx_stream_sync(c));
	c->dec_w = ow;
	c->dec_h = oh;
	c->dec_ch = C;
	if (stt) {
		memset(stt, 0, sizeof(*stt));
		stt->levels = L;
		for (int ch = 0; ch < 3; ++ch)
			stt->planes[ch] = planes[ch];
		stt->level_reached = level;
		stt->ms_coder = ev_ms(c->ev[0], c->ev[1]);
		stt->ms_linearize = ev_ms(c->ev[1], c->ev[2]);
		stt->ms_lift = ev_ms(c->ev[2], c->ev[3]);
		stt->ms_total = ev_ms(c->ev[0], c->ev[3]);
		stt->full_bits = (long long)h_state->bitpos;
		stt->parse_windows = h_state->nseg;
		stt->parse_jumps = h_state->slow_entries;
		stt->parse_exact = h_state->exact_steps;
		if (getenv("DWT_DEBUG"))
			fprintf(stderr, "resolver: %u segments, %u bulk super-windows, %u exact-step entries, %u exact slice steps, %u super rounds, "
			                "%u window rounds, %u end searches\n", h_state->nseg, h_state->nbulk, h_state->slow_entries,
			        h_state->exact_steps, h_state->n_super, h_state->n_window, h_state->n_search);
		if (getenv("DWT_DEBUG")) {
			fprintf(stderr, "  exact steps per window visit (0,1,2,<=4,..,<=128 | joined):");
			for (int t = 0; t < 10; ++t)
				fprintf(stderr, " %u", h_state->dbg_hist[t]);
			fprintf(stderr, "\n  entries: %u chunk starts, %u stray, %u end-before-join, %u window-not-joined\n", h_state->dbg_reason[0],
			        h_state->dbg_reason[1], h_state->dbg_reason[2], h_state->dbg_reason[3]);
		}
	}
	return 0;
}

extern "C" int dwt_ctx_download_image(dwt_ctx *c, uint8_t **pixels, int *width, int *height, int *channels)
{
	CUDA_OK(cudaSetDevice(c->device));
	const size_t n = (size_t)c->dec_w * c->dec_h * c->dec_ch;
	uint8_t *buf = (uint8_t *)malloc(n ? n : 1);
	if (!buf) {
		dwt_set_error("out of host memory");
		return -1;
	}
	CUDA_OK(cudaMemcpyAsync(buf, c->img.p, n, cudaMemcpyDeviceToHost, c->st));
	CUDA_OK(ctx_stream_sync(c));
	*pixels = buf;
	*width = c->dec_w;
	*height = c->dec_h;
	*channels = c->dec_ch;
	return 0;
}

extern "C" int dwt_decode(dwt_ctx *c, const uint8_t *stream, size_t len, int pixels_max, uint8_t **pixels, int *width,
                          int *height, int *channels, struct dwt_stats *stats)
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
	return dwt_ctx_download_image(c, pixels, width, height, channels);
}
