/*
 * dwt_oracle.c -- TEST INFRASTRUCTURE ONLY (see dwt_oracle.h).
 *
 * A CPU restatement of the xdsopl/dwt codec, written from the behaviour of the reference
 * (file:line citations are relative to the reference tree).  It is NOT on the product path.
 * It is pinned against the unmodified reference binaries in tests/test_oracle.py.
 */
#include "dwt_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ geometry (utils.h:9-40) */

static int floor_log2(int x) /* utils.h:9-15: -1 for x <= 0 */
{
	int l = -1;
	while (x > 0) {
		x >>= 1;
		++l;
	}
	return l;
}

int orc_geometry(int w, int h, int *lengths, int *pixels, int *widths, int *heights)
{
	/* utils.h:17-26: halve (ceil) while both halves stay >= 8; level 0 is the root */
	int ws[32], hs[32], n = 0;
	ws[0] = w;
	hs[0] = h;
	for (;;) {
		int w2 = (ws[n] + 1) / 2, h2 = (hs[n] + 1) / 2;
		++n;
		ws[n] = w2;
		hs[n] = h2;
		/* the reference recurses on (w2,h2) only if both >= N0: the first halved size that
		 * fails the test is the root (it may be smaller than 8) */
		if (!(w2 >= 8 && h2 >= 8) || n >= 15)
			break;
	}
	/* ws[n],hs[n] is the root; ws[0],hs[0] the full image; levels = n */
	int levels = n;
	for (int l = 0; l <= levels; ++l) {
		widths[l] = ws[levels - l];
		heights[l] = hs[levels - l];
		pixels[l] = widths[l] * heights[l];
		int a = 1 << (floor_log2(widths[l] - 1) + 1); /* utils.h:34-38 */
		int b = 1 << (floor_log2(heights[l] - 1) + 1);
		lengths[l] = a > b ? a : b;
	}
	return levels;
}

/* ------------------------------------------------------------------ colour (image.h:34-65) */

static int clampi(int x, int a, int b)
{
	return x < a ? a : x > b ? b : x;
}

void orc_rgb_to_ycocg(int *buf, int npix)
{
	for (int i = 0; i < npix; ++i) {
		int *p = buf + 3 * i;
		int r = p[0], g = p[1], b = p[2];
		int u = r - b;     /* image.h:58 */
		int t = b + u / 2; /* image.h:59  (C division: truncates toward zero) */
		int v = g - t;     /* image.h:60 */
		int y = t + v / 2; /* image.h:61 */
		p[0] = y;
		p[1] = u;
		p[2] = v;
	}
}

void orc_ycocg_to_rgb(int *buf, int npix)
{
	for (int i = 0; i < npix; ++i) {
		int *p = buf + 3 * i;
		int y = clampi(p[0], 0, 255); /* image.h:41-43: clamp BEFORE the inverse */
		int u = clampi(p[1], -255, 255);
		int v = clampi(p[2], -255, 255);
		int t = y - v / 2;
		int g = v + t;
		int b = t - u / 2;
		int r = b + u;
		p[0] = r;
		p[1] = g;
		p[2] = b;
	}
}

/* ------------------------------------------------------------------ 1-D lifting (cdf53.h) */

void orc_cdf53(int *out, int *in, int N, int SO, int SI, int CH)
{
	int half = (N + 1) / 2; /* cdf53.h:11: K/2 */
	for (int c = 0; c < CH; ++c) {
		int *x = in + c;
		/* predict (cdf53.h:12-17): interior odds, then the unpaired last odd when N is even */
		for (int i = 1; i + 1 < N; i += 2)
			x[i * SI] -= (x[(i - 1) * SI] + x[(i + 1) * SI]) / 2;
		if ((N & 1) == 0)
			x[(N - 1) * SI] -= x[(N - 2) * SI];
		/* update (cdf53.h:19-23): first even uses one neighbour; an odd-N tail even is untouched */
		x[0] += x[SI] / 2;
		for (int i = 2; i < (N & ~1); i += 2)
			x[i * SI] += (x[(i - 1) * SI] + x[(i + 1) * SI]) / 4;
		/* deinterleave (cdf53.h:25-33) */
		for (int i = 0; i < N; ++i) {
			int dst = (i & 1) ? half + i / 2 : i / 2;
			out[dst * SO + c] = x[i * SI];
		}
	}
}

void orc_icdf53(int *out, int *in, int N, int SO, int SI, int CH)
{
	int half = (N + 1) / 2;
	for (int c = 0; c < CH; ++c) {
		int *x = out + c;
		for (int i = 0; i < N; ++i) { /* cdf53.h:38-47 */
			int src = (i & 1) ? half + i / 2 : i / 2;
			x[i * SO] = in[src * SI + c];
		}
		x[0] -= x[SO] / 2; /* cdf53.h:49-53 */
		for (int i = 2; i < (N & ~1); i += 2)
			x[i * SO] -= (x[(i - 1) * SO] + x[(i + 1) * SO]) / 4;
		for (int i = 1; i + 1 < N; i += 2) /* cdf53.h:55-60 */
			x[i * SO] += (x[(i - 1) * SO] + x[(i + 1) * SO]) / 2;
		if ((N & 1) == 0)
			x[(N - 1) * SO] += x[(N - 2) * SO];
	}
}

/* ------------------------------------------------------------------ 2-D drivers */

void orc_forward2d(int *out, int *in, int w, int h, int ch)
{
	/* encode.c:16-30, iterative form.  SW = full row stride in samples. */
	int SW = w * ch;
	int W = w, H = h;
	for (;;) {
		for (int j = 0; j < H; ++j) { /* rows, then copy back (encode.c:18-22) */
			orc_cdf53(out + SW * j, in + SW * j, W, ch, ch, ch);
			memcpy(in + SW * j, out + SW * j, sizeof(int) * (size_t)W * ch);
		}
		orc_cdf53(out, in, H, SW, SW, W * ch); /* all columns at once (encode.c:23) */
		int W2 = (W + 1) / 2, H2 = (H + 1) / 2;
		for (int j = 0; j < H2; ++j) /* LL back to `in` (encode.c:25-27) */
			memcpy(in + SW * j, out + SW * j, sizeof(int) * (size_t)W2 * ch);
		if (!(W2 >= 8 && H2 >= 8))
			break;
		W = W2;
		H = H2;
	}
}

void orc_inverse2d(int *out, int *in, int w, int h, int ch)
{
	/* decode.c:16-30: coarsest level first; per level columns then rows */
	int SW = w * ch;
	int Ws[32], Hs[32], n = 0;
	Ws[0] = w;
	Hs[0] = h;
	for (;;) {
		int W2 = (Ws[n] + 1) / 2, H2 = (Hs[n] + 1) / 2;
		if (!(W2 >= 8 && H2 >= 8))
			break;
		++n;
		Ws[n] = W2;
		Hs[n] = H2;
	}
	for (int k = n; k >= 0; --k) {
		int W = Ws[k], H = Hs[k];
		orc_icdf53(out, in, H, SW, SW, W * ch);
		for (int j = 0; j < H; ++j)
			memcpy(in + SW * j, out + SW * j, sizeof(int) * (size_t)W * ch);
		for (int j = 0; j < H; ++j) {
			orc_icdf53(out + SW * j, in + SW * j, W, ch, ch, ch);
			memcpy(in + SW * j, out + SW * j, sizeof(int) * (size_t)W * ch);
		}
	}
}

/* ------------------------------------------------------------------ Hilbert (hilbert.h:15-34) */

void orc_hilbert(int n, int d, int *px, int *py)
{
	int x = 0, y = 0;
	for (int s = 1; s < n; s *= 2, d /= 4) {
		int rx = (d / 2) & 1;
		int ry = (d ^ rx) & 1;
		if (!ry) {
			if (rx) {
				x = s - 1 - x;
				y = s - 1 - y;
			}
			int t = x;
			x = y;
			y = t;
		}
		x += s * rx;
		y += s * ry;
	}
	*px = x;
	*py = y;
}

void orc_linearize(int *planar, const int *pyr, int w, int h, int ch)
{
	int lengths[16], pixels[16], widths[16], heights[16];
	int levels = orc_geometry(w, h, lengths, pixels, widths, heights);
	long long total = (long long)w * h;
	long long k = 0;
	for (int y = 0; y < heights[0]; ++y) /* root raster (encode.c:37-45) */
		for (int x = 0; x < widths[0]; ++x, ++k)
			for (int c = 0; c < ch; ++c)
				planar[c * total + k] = pyr[ch * ((long long)w * y + x) + c];
	for (int l = 0; l < levels; ++l) { /* detail levels in Hilbert order (encode.c:46-57) */
		long long nn = (long long)lengths[l + 1] * lengths[l + 1];
		for (long long d = 0; d < nn; ++d) {
			int x, y;
			orc_hilbert(lengths[l + 1], (int)d, &x, &y);
			if (x >= widths[l + 1] || y >= heights[l + 1])
				continue;
			if (x < widths[l] && y < heights[l])
				continue; /* LL rectangle belongs to the coarser levels */
			for (int c = 0; c < ch; ++c)
				planar[c * total + k] = pyr[ch * ((long long)w * y + x) + c];
			++k;
		}
	}
}

/* ------------------------------------------------------------------ bit sink with capacity
 * bytes.h:75-85 (cap), bits.h:58-78 (LSB-first), vli.h:67-84 (adaptive Rice), rle.h:56-89 (runs) */

struct sink {
	uint8_t *buf;
	long long room, cnt, cap;
	int acc, nacc;
	int order;
	int run;
	int overflow;
};

static int sink_byte(struct sink *s, int b)
{
	if (s->cap > 0 && s->cnt >= s->cap)
		return -2;
	if (s->cnt >= s->room) {
		s->overflow = 1;
		return -1;
	}
	s->buf[s->cnt++] = (uint8_t)b;
	return 0;
}

static int sink_bit(struct sink *s, int b)
{
	s->acc |= (b ? 1 : 0) << s->nacc++;
	if (s->nacc >= 8) {
		s->nacc -= 8;
		int v = s->acc;
		s->acc >>= 8;
		return sink_byte(s, v & 255);
	}
	return 0;
}

static int sink_bits(struct sink *s, int v, int n)
{
	for (int i = 0; i < n; ++i) {
		int r = sink_bit(s, (v >> i) & 1);
		if (r)
			return r;
	}
	return 0;
}

static int sink_vli(struct sink *s, int val)
{
	int r;
	while (val >= (1 << s->order)) {
		if ((r = sink_bit(s, 0)))
			return r;
		val -= 1 << s->order;
		s->order += 1;
	}
	if ((r = sink_bit(s, 1)))
		return r;
	if ((r = sink_bits(s, val, s->order)))
		return r;
	s->order = s->order >= 2 ? s->order - 2 : 0;
	return 0;
}

static int sink_symbol(struct sink *s, int one) /* rle.h:56-64 */
{
	if (s->run < 0)
		return s->run;
	if (one)
		return s->run = sink_vli(s, s->run);
	s->run++;
	return 0;
}

static int sink_raw(struct sink *s, int bit) /* rle.h:79-89: phantom one before a raw bit */
{
	if (s->run < 0)
		return s->run;
	if (s->run > 0) {
		int r = sink_symbol(s, 1);
		if (r)
			return r;
	}
	return sink_bit(s, bit);
}

static long long sink_bitcount(const struct sink *s)
{
	return s->nacc + 8 * s->cnt;
}

/* one (channel, level, plane) chunk in the stateless form of encode.c:60-95:
 * "not yet significant" <=> no magnitude bit above `plane` is set */
static int code_chunk(struct sink *s, const int *v, long long num, int plane)
{
	for (long long i = 0; i < num; ++i) {
		unsigned mag = (unsigned)(v[i] < 0 ? -v[i] : v[i]);
		unsigned above = plane >= 0 ? (plane >= 31 ? 0 : mag >> (plane + 1)) : mag;
		if (above)
			continue;
		int bit = plane >= 0 ? (int)((mag >> plane) & 1) : 0;
		int r = sink_symbol(s, bit);
		if (r)
			return r;
		if (bit && (r = sink_raw(s, v[i] < 0)))
			return r;
	}
	for (long long i = 0; i < num; ++i) {
		unsigned mag = (unsigned)(v[i] < 0 ? -v[i] : v[i]);
		unsigned above = plane >= 0 ? (plane >= 31 ? 0 : mag >> (plane + 1)) : mag;
		if (!above)
			continue;
		int r = sink_raw(s, plane >= 0 ? (int)((mag >> plane) & 1) : 0);
		if (r)
			return r;
	}
	return 0;
}

int orc_front_end(const uint8_t *pix, int w, int h, int ch, int *pyramid, int *planar, int *planes)
{
	if (w < 8 || h < 8 || w > 65536 || h > 65536 || (ch != 1 && ch != 3))
		return -1;
	int lengths[16], pixels[16], widths[16], heights[16];
	orc_geometry(w, h, lengths, pixels, widths, heights);
	size_t n = (size_t)w * h * ch;
	int *img = malloc(sizeof(int) * n);
	int *pyr = pyramid ? pyramid : malloc(sizeof(int) * n);
	int *lin = planar ? planar : malloc(sizeof(int) * n);
	if (!img || !pyr || !lin)
		return -1;
	for (size_t i = 0; i < n; ++i)
		img[i] = pix[i];
	if (ch == 3)
		orc_rgb_to_ycocg(img, w * h);
	orc_forward2d(pyr, img, w, h, ch);
	orc_linearize(lin, pyr, w, h, ch);
	long long total = (long long)w * h;
	for (int c = 0; c < ch; ++c) { /* encode.c:112-131, 163-165 */
		int mx = 0;
		for (long long i = pixels[0]; i < total; ++i) {
			int a = lin[c * total + i];
			a = a < 0 ? -a : a;
			if (a > mx)
				mx = a;
		}
		planes[c] = 1 + floor_log2(mx);
	}
	free(img);
	if (!pyramid)
		free(pyr);
	if (!planar)
		free(lin);
	return 0;
}

long long orc_encode(const uint8_t *pix, int w, int h, int ch, int capacity,
                     uint8_t *out, long long out_room, struct orc_stats *st)
{
	if (w < 8 || h < 8 || w > 65536 || h > 65536 || (ch != 1 && ch != 3))
		return -1;
	int lengths[16], pixels[16], widths[16], heights[16];
	int levels = orc_geometry(w, h, lengths, pixels, widths, heights);
	long long total = (long long)w * h;
	int *lin = malloc(sizeof(int) * (size_t)total * ch);
	int planes[3] = { 0, 0, 0 };
	if (!lin || orc_front_end(pix, w, h, ch, NULL, lin, planes))
		return -1;

	struct sink s;
	memset(&s, 0, sizeof(s));
	s.buf = out;
	s.room = out_room;
	s.cap = capacity;
	/* header encode.c:169-172 (return values ignored, like the reference) */
	sink_byte(&s, 'W');
	sink_byte(&s, ch == 3 ? '6' : '5');
	sink_byte(&s, (w - 1) & 255);
	sink_byte(&s, ((w - 1) >> 8) & 255);
	sink_byte(&s, (h - 1) & 255);
	sink_byte(&s, ((h - 1) >> 8) & 255);
	long long meta = sink_bitcount(&s);
	for (int c = 0; c < ch; ++c) { /* encode_root encode.c:97-110 */
		const int *v = lin + c * total;
		int mx = 0;
		for (int i = 0; i < pixels[0]; ++i) {
			int a = v[i] < 0 ? -v[i] : v[i];
			if (a > mx)
				mx = a;
		}
		int cnt = 1 + floor_log2(mx);
		sink_vli(&s, cnt);
		for (int i = 0; cnt && i < pixels[0]; ++i) {
			sink_bits(&s, v[i] < 0 ? -v[i] : v[i], cnt);
			if (v[i])
				sink_bit(&s, v[i] < 0);
		}
	}
	long long root = sink_bitcount(&s);
	for (int c = 0; c < ch; ++c)
		sink_vli(&s, planes[c]);
	int planes_max = 0;
	for (int c = 0; c < ch; ++c)
		if (planes[c] > planes_max)
			planes_max = planes[c];
	int maximum = levels > planes_max ? levels : planes_max;
	int layers_max = 2 * maximum - 1;
	int stop = 0;
	/* schedule encode.c:183-221 */
	if (planes_max == planes[0])
		stop = code_chunk(&s, lin + pixels[0], pixels[1] - pixels[0], planes[0] - 1);
	for (int layers = 0; !stop && layers < layers_max; ++layers) {
		for (int l = 0; !stop && l < levels && l <= layers + 1; ++l) {
			int plane = planes_max - 1 - (layers + 1 - l);
			if (plane < 0 || plane >= planes[0])
				continue;
			stop = code_chunk(&s, lin + pixels[l], pixels[l + 1] - pixels[l], plane);
		}
		for (int l = 0; !stop && l < levels && l <= layers; ++l) {
			int plane = planes_max - 1 - (layers - l);
			for (int c = 1; !stop && c < ch; ++c) {
				if (plane < 0 || plane >= planes[c])
					continue;
				stop = code_chunk(&s, lin + c * total + pixels[l], pixels[l + 1] - pixels[l], plane);
			}
		}
	}
	if (!stop)
		s.run = sink_vli(&s, s.run); /* rle_flush rle.h:37-40: always one final VLI */
	long long bits = sink_bitcount(&s);
	if (s.nacc) /* close_bits_writer bits.h:51-56 */
		sink_byte(&s, s.acc & 255);
	free(lin);
	if (st) {
		st->meta_bits = meta;
		st->root_bits = root - meta;
		st->total_bits = bits;
		st->bytes = s.cnt;
		st->levels = levels;
		for (int c = 0; c < 3; ++c)
			st->planes[c] = c < ch ? planes[c] : 0;
	}
	if (s.overflow)
		return -1;
	return s.cnt;
}

/* ------------------------------------------------------------------ bit source (bits.h:80-106, vli.h:86-101, rle.h:66-103) */

struct source {
	const uint8_t *buf;
	long long len, pos;
	int acc, nacc;
	int order;
	int run;
};

static int src_bit(struct source *s)
{
	if (!s->nacc) {
		if (s->pos >= s->len)
			return -1; /* EOF only when a new byte is needed (bits.h:82-88) */
		s->acc = s->buf[s->pos++];
		s->nacc = 8;
	}
	int b = s->acc & 1;
	s->acc >>= 1;
	s->nacc -= 1;
	return b;
}

static int src_bits(struct source *s, int *v, int n)
{
	int a = 0;
	for (int i = 0; i < n; ++i) {
		int b = src_bit(s);
		if (b < 0)
			return b;
		a |= b << i;
	}
	*v = a;
	return 0;
}

static int src_vli(struct source *s)
{
	int val = 0, sum = 0, r;
	while ((r = src_bit(s)) == 0) {
		sum += 1 << s->order;
		s->order += 1;
	}
	if (r < 0)
		return r;
	if ((r = src_bits(s, &val, s->order)))
		return r;
	s->order = s->order >= 2 ? s->order - 2 : 0;
	return val + sum;
}

static int src_symbol(struct source *s) /* rle.h:66-77 */
{
	if (s->run < 0)
		return s->run;
	if (!s->run) {
		s->run = src_vli(s);
		if (s->run < 0)
			return s->run;
		return !s->run;
	}
	return s->run-- == 1;
}

static int src_raw(struct source *s) /* rle.h:91-103 */
{
	if (s->run < 0)
		return s->run;
	if (s->run > 0) {
		int r = src_symbol(s);
		if (r < 0)
			return r;
		if (r != 1)
			return -1;
	}
	return src_bit(s);
}

/* decode.c:67-100 on a (magnitude, sign) pair of arrays instead of flag bits */
static int parse_chunk(struct source *s, unsigned *mag, uint8_t *neg, long long num, int plane)
{
	/* significance is evaluated against the state at chunk entry: the reference marks newly
	 * significant values with a flag that only turns into "refinable" after the chunk */
	for (long long i = 0; i < num; ++i) {
		unsigned above = plane >= 0 ? (plane >= 31 ? 0 : mag[i] >> (plane + 1)) : mag[i];
		if (above)
			continue;
		int bit = src_symbol(s);
		if (bit < 0)
			return bit;
		if (bit) {
			if (plane >= 0)
				mag[i] |= 1u << plane;
			int sg = src_raw(s);
			if (sg < 0)
				return sg;
			neg[i] = (uint8_t)sg;
		}
	}
	for (long long i = 0; i < num; ++i) {
		unsigned above = plane >= 0 ? (plane >= 31 ? 0 : mag[i] >> (plane + 1)) : mag[i];
		if (!above)
			continue;
		int bit = src_raw(s);
		if (bit < 0)
			return bit;
		if (plane >= 0)
			mag[i] |= (unsigned)bit << plane;
	}
	return 0;
}

int orc_decode_coeffs(const uint8_t *stream, long long len, int pixels_max,
                      int **planar_out, int *missing, int *level_out, int *pw, int *ph, int *pch)
{
	struct source s;
	memset(&s, 0, sizeof(s));
	s.buf = stream;
	s.len = len;
	/* header decode.c:145-159 */
	if (len < 1 || stream[0] != 'W')
		return 1;
	if (len < 2 || (stream[1] != '5' && stream[1] != '6'))
		return 1;
	if (len < 6)
		return 1;
	int color = stream[1] == '6';
	int width = (stream[2] | (stream[3] << 8)) + 1;
	int height = (stream[4] | (stream[5] << 8)) + 1;
	s.pos = 6;
	if (width < 8 || height < 8)
		return 1;
	int lengths[16], pixels[16], widths[16], heights[16];
	int levels = orc_geometry(width, height, lengths, pixels, widths, heights);
	int levels_max = levels;
	if (pixels_max >= 0) { /* decode.c:165-171 */
		while (levels_max > 0 && pixels[levels_max] > pixels_max)
			--levels_max;
	}
	long long total = (long long)widths[levels_max] * heights[levels_max];
	int ch = color ? 3 : 1;
	unsigned *mag = calloc((size_t)total * ch, sizeof(unsigned));
	uint8_t *neg = calloc((size_t)total * ch, 1);
	int *root = calloc((size_t)pixels[0] * ch, sizeof(int));
	if (!mag || !neg || !root)
		return 1;
	for (int c = 0; c < ch; ++c) { /* decode_root decode.c:119-134 */
		int cnt = src_vli(&s);
		if (cnt < 0)
			goto fail;
		for (int i = 0; cnt && i < pixels[0]; ++i) {
			int v = 0, r;
			if (src_bits(&s, &v, cnt))
				goto fail;
			r = 0;
			if (v && (r = src_bit(&s)) > 0)
				v = -v;
			if (r < 0)
				goto fail;
			root[c * pixels[0] + i] = v;
		}
	}
	int planes[3] = { 0, 0, 0 };
	for (int c = 0; c < ch; ++c)
		if ((planes[c] = src_vli(&s)) < 0)
			goto fail;
	int planes_max = 0;
	for (int c = 0; c < ch; ++c)
		if (planes[c] > planes_max)
			planes_max = planes[c];
	int maximum = levels > planes_max ? levels : planes_max;
	int layers_max = 2 * maximum - 1;
	for (int i = 0; i < 48; ++i)
		missing[i] = 0;
	for (int c = 0; c < ch; ++c)
		for (int l = 0; l < levels; ++l)
			missing[c * 16 + l] = planes[c];
	int level = -1;
	if (!levels_max)
		goto end;
	if (planes_max == planes[0]) { /* decode.c:201-207 */
		level = 0;
		if (parse_chunk(&s, mag + pixels[0], neg + pixels[0], pixels[1] - pixels[0], planes[0] - 1))
			goto end;
		--missing[0];
	}
	for (int layers = 0; layers < layers_max; ++layers) {
		for (int l = 0; l < levels && l <= layers + 1; ++l) {
			if (l >= levels_max)
				goto end;
			int plane = planes_max - 1 - (layers + 1 - l);
			if (plane < 0 || plane >= planes[0])
				continue;
			if (level < l)
				level = l;
			if (parse_chunk(&s, mag + pixels[l], neg + pixels[l], pixels[l + 1] - pixels[l], plane))
				goto end;
			--missing[l];
		}
		for (int l = 0; l < levels && l <= layers; ++l) {
			if (l >= levels_max)
				goto end;
			int plane = planes_max - 1 - (layers - l);
			for (int c = 1; c < ch; ++c) {
				if (plane < 0 || plane >= planes[c])
					continue;
				if (level < l)
					level = l;
				if (parse_chunk(&s, mag + c * total + pixels[l], neg + c * total + pixels[l],
				                pixels[l + 1] - pixels[l], plane))
					goto end;
				--missing[c * 16 + l];
			}
		}
	}
end:;
	/* process decode.c:102-117 + output geometry decode.c:251-254 */
	int out_levels = level + 1;
	long long out_total = pixels[out_levels];
	int *planar = calloc((size_t)out_total * ch, sizeof(int));
	if (!planar)
		goto fail;
	for (int c = 0; c < ch; ++c) {
		for (int i = 0; i < pixels[0]; ++i)
			planar[c * out_total + i] = root[c * pixels[0] + i];
		for (long long i = pixels[0]; i < out_total; ++i) {
			int m = (int)(mag[c * total + i] & 0x1fffffffu);
			planar[c * out_total + i] = neg[c * total + i] ? -m : m;
		}
	}
	free(mag);
	free(neg);
	free(root);
	*planar_out = planar;
	*level_out = level;
	*pw = width;
	*ph = height;
	*pch = ch;
	return 0;
fail:
	free(mag);
	free(neg);
	free(root);
	return 1;
}

int orc_decode(const uint8_t *stream, long long len, int pixels_max,
               uint8_t **pix, int *pw, int *ph, int *pch)
{
	int *planar = NULL, missing[48], level = -1, fw, fh, ch;
	if (orc_decode_coeffs(stream, len, pixels_max, &planar, missing, &level, &fw, &fh, &ch))
		return 1;
	int lengths[16], pixels[16], widths[16], heights[16];
	orc_geometry(fw, fh, lengths, pixels, widths, heights);
	int levels = level + 1;
	int w = widths[levels], h = heights[levels];
	long long total = pixels[levels];
	int *temp = malloc(sizeof(int) * (size_t)total * ch);
	int *img = malloc(sizeof(int) * (size_t)total * ch);
	uint8_t *out = malloc((size_t)total * ch);
	if (!temp || !img || !out)
		return 1;
	/* reconstruction decode.c:32-65 (inverse scatter + dequantisation bias) */
	long long k = 0;
	for (int y = 0; y < heights[0]; ++y)
		for (int x = 0; x < widths[0]; ++x, ++k)
			for (int c = 0; c < ch; ++c)
				temp[ch * ((long long)w * y + x) + c] = planar[c * total + k];
	for (int l = 0; l < levels; ++l) {
		long long nn = (long long)lengths[l + 1] * lengths[l + 1];
		for (long long d = 0; d < nn; ++d) {
			int x, y;
			orc_hilbert(lengths[l + 1], (int)d, &x, &y);
			if (x >= widths[l + 1] || y >= heights[l + 1] || (x < widths[l] && y < heights[l]))
				continue;
			for (int c = 0; c < ch; ++c) {
				int v = planar[c * total + k];
				int m = missing[c * 16 + l] - 2;
				if (m >= 0) {
					int bias = 1 << m;
					if (v < 0)
						v -= bias;
					else if (v > 0)
						v += bias;
				}
				temp[ch * ((long long)w * y + x) + c] = v;
			}
			++k;
		}
	}
	orc_inverse2d(img, temp, w, h, ch);
	if (ch == 3)
		orc_ycocg_to_rgb(img, (int)total);
	for (long long i = 0; i < total * ch; ++i) /* write_pnm clamp pnm.h:108 */
		out[i] = (uint8_t)clampi(img[i], 0, 255);
	free(planar);
	free(temp);
	free(img);
	*pix = out;
	*pw = w;
	*ph = h;
	*pch = ch;
	return 0;
}

void orc_free(void *p)
{
	free(p);
}

/* ------------------------------------------------------------------ synthetic inputs */

static uint32_t hash32(uint32_t x, uint32_t y, uint32_t c, uint32_t seed)
{
	uint32_t h = seed ^ (x * 0x9E3779B1u) ^ (y * 0x85EBCA77u) ^ (c * 0xC2B2AE3Du);
	h ^= h >> 15;
	h *= 0x2C1B3C6Du;
	h ^= h >> 12;
	h *= 0x297A2D39u;
	h ^= h >> 15;
	return h;
}

void orc_synth(uint8_t *out, int w, int h, int kind, uint32_t seed)
{
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x)
			for (int c = 0; c < 3; ++c) {
				uint32_t hv = hash32((uint32_t)x, (uint32_t)y, (uint32_t)c, seed);
				int v;
				if (kind == 1) {
					v = (int)(hv >> 24);
				} else {
					int s = ((x * (c + 2) + y * (5 - c)) >> 3) & 511;
					int tri = s < 256 ? s : 511 - s;
					int blk = (((x >> 6) ^ (y >> 6)) & 1) * 40;
					int n = (int)(hv & 7) - 4;
					v = clampi((tri * 3) / 4 + blk + n, 0, 255);
				}
				out[((size_t)y * w + x) * 3 + c] = (uint8_t)v;
			}
}
