/*
 * streamio.c -- host-side byte / bit / VLI / run-length sinks and sources of libdwt_b200.
 *
 * Same entry-point names, argument meaning, ownership rules and error codes as the reference headers
 * (bytes.h:23-118, bits.h:23-106, vli.h:21-101, rle.h:21-103): constructors malloc a small object that
 * the matching delete_ / close_ call frees, lower layers are borrowed, writers return 0 / -1 (I/O) /
 * -2 (capacity reached), readers return a value >= 0 or a negative error, errors latch in the run
 * counter.  The codec uses them for the serial stream prefix (header, root image, plane counts); the
 * bit-plane payload is produced and consumed by the CUDA kernels.
 *
 * Added here (not in the reference): memory-backed byte sinks/sources, so the library never needs a file.
 */
#include "dwt_b200.h"

#include <stdlib.h>
#include <string.h>

struct bytes_reader {
	FILE *file;          /* NULL for a memory source */
	char *name;
	const uint8_t *mem;
	size_t len, pos;
};

struct bytes_writer {
	FILE *file;          /* NULL for a memory sink */
	char *name;
	int cnt, cap;
	uint8_t *mem;
	size_t room;
};

struct bits_reader { struct bytes_reader *bytes; int acc, cnt; };
struct bits_writer { struct bytes_writer *bytes; int acc, cnt; };
struct vli_reader { struct bits_reader *bits; int order; };
struct vli_writer { struct bits_writer *bits; int order; };
struct rle_reader { struct vli_reader *vli; int cnt; };
struct rle_writer { struct vli_writer *vli; int cnt; };

static const char *real_name(const char *name, const char *dash)
{
	return (name[0] == '-' && !name[1]) ? dash : name; /* bytes.h:26-28,42-44 */
}

/* ---- bytes ---- */

struct bytes_reader *bytes_reader(char *name)
{
	const char *fname = real_name(name, "/dev/stdin");
	FILE *f = fopen(fname, "r");
	if (!f) {
		fprintf(stderr, "could not open \"%s\" file to read\n", fname);
		return 0;
	}
	struct bytes_reader *b = calloc(1, sizeof(*b));
	b->file = f;
	b->name = name;
	return b;
}

struct bytes_reader *bytes_reader_mem(const uint8_t *data, size_t len)
{
	struct bytes_reader *b = calloc(1, sizeof(*b));
	b->name = (char *)"<memory>";
	b->mem = data;
	b->len = len;
	return b;
}

struct bytes_writer *bytes_writer(char *name, int capacity)
{
	const char *fname = real_name(name, "/dev/stdout");
	FILE *f = fopen(fname, "w");
	if (!f) {
		fprintf(stderr, "could not open \"%s\" file to write\n", fname);
		return 0;
	}
	struct bytes_writer *b = calloc(1, sizeof(*b));
	b->file = f;
	b->name = name;
	b->cap = capacity;
	return b;
}

struct bytes_writer *bytes_writer_mem(int capacity)
{
	struct bytes_writer *b = calloc(1, sizeof(*b));
	b->name = (char *)"<memory>";
	b->cap = capacity;
	return b;
}

const uint8_t *bytes_writer_data(struct bytes_writer *bytes, size_t *len)
{
	if (len)
		*len = (size_t)bytes->cnt;
	return bytes->mem;
}

int bytes_count(struct bytes_writer *bytes)
{
	return bytes->cnt;
}

void close_bytes_reader(struct bytes_reader *bytes)
{
	if (bytes->file)
		fclose(bytes->file);
	free(bytes);
}

void close_bytes_writer(struct bytes_writer *bytes)
{
	if (bytes->file)
		fclose(bytes->file);
	free(bytes->mem);
	free(bytes);
}

int put_byte(struct bytes_writer *bytes, int b)
{
	if (bytes->cap > 0 && bytes->cnt >= bytes->cap) /* bytes.h:77-78: refuse, write nothing */
		return -2;
	if (bytes->file) {
		if (fputc(b & 255, bytes->file) == EOF) {
			fprintf(stderr, "could not write to file \"%s\"\n", bytes->name);
			return -1;
		}
	} else {
		if ((size_t)bytes->cnt >= bytes->room) {
			size_t room = bytes->room ? 2 * bytes->room : 256;
			uint8_t *m = realloc(bytes->mem, room);
			if (!m)
				return -1;
			bytes->mem = m;
			bytes->room = room;
		}
		bytes->mem[bytes->cnt] = (uint8_t)(b & 255);
	}
	bytes->cnt += 1;
	return 0;
}

int write_bytes(struct bytes_writer *bytes, int b, int n)
{
	for (int k = 0; k < n; ++k) { /* little endian, bytes.h:87-95 */
		int ret = put_byte(bytes, b >> (8 * k));
		if (ret)
			return ret;
	}
	return 0;
}

int get_byte(struct bytes_reader *bytes)
{
	if (bytes->file) {
		int b = fgetc(bytes->file);
		if (b != EOF)
			return b;
	} else if (bytes->pos < bytes->len) {
		return bytes->mem[bytes->pos++];
	}
	if (bytes->file) /* memory sources stay quiet: the library is not a program */
		fprintf(stderr, "reached end of file \"%s\"\n", bytes->name);
	return -1;
}

int read_bytes(struct bytes_reader *bytes, int *b, int n)
{
	int a = 0;
	for (int k = 0; k < n; ++k) {
		int v = get_byte(bytes);
		if (v < 0)
			return v;
		a |= v << (8 * k);
	}
	*b = a;
	return 0;
}

/* ---- bits: LSB first inside each byte (bits.h:58-93) ---- */

struct bits_reader *bits_reader(struct bytes_reader *bytes)
{
	struct bits_reader *r = calloc(1, sizeof(*r));
	r->bytes = bytes;
	return r;
}

struct bits_writer *bits_writer(struct bytes_writer *bytes)
{
	struct bits_writer *w = calloc(1, sizeof(*w));
	w->bytes = bytes;
	return w;
}

int bits_count(struct bits_writer *bits)
{
	return bits->cnt + 8 * bytes_count(bits->bytes);
}

void close_bits_reader(struct bits_reader *bits)
{
	free(bits);
}

void close_bits_writer(struct bits_writer *bits)
{
	if (bits->cnt) /* zero-padded last byte, bits.h:51-56 */
		put_byte(bits->bytes, bits->acc);
	free(bits);
}

int put_bit(struct bits_writer *bits, int b)
{
	if (b)
		bits->acc |= 1 << bits->cnt;
	if (++bits->cnt < 8)
		return 0;
	int full = bits->acc & 255;
	bits->acc >>= 8;
	bits->cnt -= 8;
	return put_byte(bits->bytes, full);
}

int write_bits(struct bits_writer *bits, int b, int n)
{
	for (int k = 0; k < n; ++k) {
		int ret = put_bit(bits, (b >> k) & 1);
		if (ret)
			return ret;
	}
	return 0;
}

int get_bit(struct bits_reader *bits)
{
	if (bits->cnt == 0) { /* EOF shows up only when a fresh byte is needed, bits.h:82-88 */
		int v = get_byte(bits->bytes);
		if (v < 0)
			return v;
		bits->acc = v;
		bits->cnt = 8;
	}
	int b = bits->acc & 1;
	bits->acc >>= 1;
	bits->cnt -= 1;
	return b;
}

int read_bits(struct bits_reader *bits, int *b, int n)
{
	int a = 0;
	for (int k = 0; k < n; ++k) {
		int v = get_bit(bits);
		if (v < 0)
			return v;
		a |= v << k;
	}
	*b = a;
	return 0;
}

/* ---- adaptive Rice code (vli.h:67-101) ---- */

struct vli_reader *vli_reader(struct bits_reader *bits)
{
	struct vli_reader *v = calloc(1, sizeof(*v));
	v->bits = bits;
	return v;
}

struct vli_writer *vli_writer(struct bits_writer *bits)
{
	struct vli_writer *v = calloc(1, sizeof(*v));
	v->bits = bits;
	return v;
}

void delete_vli_reader(struct vli_reader *vli) { free(vli); }
void delete_vli_writer(struct vli_writer *vli) { free(vli); }
int vli_put_bit(struct vli_writer *vli, int bit) { return put_bit(vli->bits, bit); }
int vli_get_bit(struct vli_reader *vli) { return get_bit(vli->bits); }
int vli_write_bits(struct vli_writer *vli, int b, int n) { return write_bits(vli->bits, b, n); }
int vli_read_bits(struct vli_reader *vli, int *b, int n) { return read_bits(vli->bits, b, n); }

static int relax(int order)
{
	return order >= 2 ? order - 2 : 0;
}

int put_vli(struct vli_writer *vli, int val)
{
	int ret;
	for (; val >= (1 << vli->order); vli->order++) { /* one 0 per escalation of the order */
		if ((ret = put_bit(vli->bits, 0)))
			return ret;
		val -= 1 << vli->order;
	}
	if ((ret = put_bit(vli->bits, 1)))
		return ret;
	if ((ret = write_bits(vli->bits, val, vli->order)))
		return ret;
	vli->order = relax(vli->order);
	return 0;
}

int get_vli(struct vli_reader *vli)
{
	int sum = 0, val = 0, ret;
	for (;;) {
		ret = get_bit(vli->bits);
		if (ret)
			break;
		sum += 1 << vli->order;
		vli->order++;
	}
	if (ret < 0)
		return ret;
	if ((ret = read_bits(vli->bits, &val, vli->order)))
		return ret;
	vli->order = relax(vli->order);
	return sum + val;
}

/* ---- zero-run coder (rle.h:37-103) ---- */

struct rle_reader *rle_reader(struct vli_reader *vli)
{
	struct rle_reader *r = calloc(1, sizeof(*r));
	r->vli = vli;
	return r;
}

struct rle_writer *rle_writer(struct vli_writer *vli)
{
	struct rle_writer *w = calloc(1, sizeof(*w));
	w->vli = vli;
	return w;
}

int rle_flush(struct rle_writer *rle)
{
	rle->cnt = put_vli(rle->vli, rle->cnt); /* always one VLI, even for an empty run */
	return rle->cnt;
}

void delete_rle_reader(struct rle_reader *rle)
{
	if (rle->cnt > 1)
		fprintf(stderr, "%d zeros not read.\n", rle->cnt);
	free(rle);
}

void delete_rle_writer(struct rle_writer *rle)
{
	if (rle->cnt > 0)
		fprintf(stderr, "forgot to flush counter for %d zeros.\n", rle->cnt);
	free(rle);
}

int put_rle(struct rle_writer *rle, int b)
{
	if (rle->cnt < 0)
		return rle->cnt;
	if (!b) {
		rle->cnt++;
		return 0;
	}
	rle->cnt = put_vli(rle->vli, rle->cnt);
	return rle->cnt;
}

int get_rle(struct rle_reader *rle)
{
	if (rle->cnt < 0)
		return rle->cnt;
	if (rle->cnt == 0) {
		rle->cnt = get_vli(rle->vli);
		if (rle->cnt < 0)
			return rle->cnt;
		return rle->cnt == 0; /* the first zero of a fresh run is reported without counting down */
	}
	int last = rle->cnt == 1;
	rle->cnt--;
	return last;
}

int rle_put_bit(struct rle_writer *rle, int bit)
{
	if (rle->cnt < 0)
		return rle->cnt;
	if (rle->cnt > 0) { /* a pending run is closed by a phantom one before any raw bit */
		int ret = put_rle(rle, 1);
		if (ret)
			return ret;
	}
	return vli_put_bit(rle->vli, bit);
}

int rle_get_bit(struct rle_reader *rle)
{
	if (rle->cnt < 0)
		return rle->cnt;
	if (rle->cnt > 0) {
		int ret = get_rle(rle);
		if (ret < 0)
			return ret;
		if (ret != 1)
			return -1;
	}
	return vli_get_bit(rle->vli);
}

/* ---- geometry (utils.h:9-40) ---- */

int ilog2(int x)
{
	int l = -1;
	while (x > 0) {
		x /= 2;
		++l;
	}
	return l;
}

int compute_lengths(int *lengths, int *pixels, int *widths, int *heights, int W, int H, int N0)
{
	int ws[20], hs[20], n = 0;
	ws[0] = W;
	hs[0] = H;
	do { /* halve with ceil until a halved size fails the N0 test: that size is the root */
		ws[n + 1] = (ws[n] + 1) / 2;
		hs[n + 1] = (hs[n] + 1) / 2;
		++n;
	} while (ws[n] >= N0 && hs[n] >= N0 && n < 15);
	for (int l = 0; l <= n; ++l) {
		widths[l] = ws[n - l];
		heights[l] = hs[n - l];
		pixels[l] = widths[l] * heights[l];
		int a = 1 << (ilog2(widths[l] - 1) + 1);
		int b = 1 << (ilog2(heights[l] - 1) + 1);
		lengths[l] = a > b ? a : b;
	}
	return n;
}

/* ---- accessors for the pipeline (streamio_internal.h) ---- */

int dwt_vli_writer_order(struct vli_writer *vli) { return vli->order; }
int dwt_vli_reader_order(struct vli_reader *vli) { return vli->order; }

long long dwt_bits_reader_position(struct bits_reader *bits)
{
	struct bytes_reader *b = bits->bytes;
	long long bytes = b->file ? (long long)ftell(b->file) : (long long)b->pos;
	return 8 * bytes - bits->cnt;
}
