/*
 * pnm.c -- Netpbm P5/P6 file I/O for the drop-in CLIs (behaviour of pnm.h:14-117, bulk I/O instead of
 * one fgetc/fputc per byte).
 *
 * Accepted input, like the reference: "P5" / "P6", then width, height, maxval separated by whitespace with
 * optional '#' comment lines, maxval must be 255 (pnm.h:63-67), exactly one byte after maxval, then
 * width*height*channels raw bytes.  "-" means stdin / stdout (pnm.h:16-18,93-95).
 * The writer emits "P%d %d %d 255\n" (pnm.h:102).
 */
#include "pnm.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char *real_name(const char *name, const char *dash)
{
	return (name[0] == '-' && !name[1]) ? dash : name;
}

/* header tokenizer: `cur` is the byte most recently taken from the file */
struct cursor {
	FILE *file;
	int cur;
};

static int advance(struct cursor *k)
{
	k->cur = fgetc(k->file);
	return k->cur != EOF;
}

static int is_digit(int c)
{
	return c >= '0' && c <= '9';
}

/* One decimal header field.  Entered with `cur` = the byte that ended the previous field (or the single byte
 * behind the magic), which is dropped.  '#' right at the start of a field opens a comment line, any number of
 * them (pnm.h:38-41); everything else that is not a digit is skipped; at most 15 digits are taken (pnm.h:45-54);
 * the byte that ends the field is consumed -- behind the third field that is the one separator in front of the
 * pixel data.  Returns 0 when the file ends first. */
static int header_field(struct cursor *k, long *value)
{
	if (!advance(k))
		return 0;
	while (k->cur == '#') {
		do {
			if (!advance(k))
				return 0;
		} while (k->cur != '\n');
		if (!advance(k))
			return 0;
	}
	while (!is_digit(k->cur))
		if (!advance(k))
			return 0;
	long v = 0;
	for (int digits = 0; is_digit(k->cur) && digits < 15; ++digits) {
		v = 10 * v + (k->cur - '0');
		if (!advance(k))
			return 0;
	}
	*value = v;
	return 1;
}

uint8_t *dwt_read_pnm(const char *name, int *width, int *height, int *channels)
{
	const char *fname = real_name(name, "/dev/stdin");
	struct cursor k = {fopen(fname, "r"), 0};
	if (!k.file) {
		fprintf(stderr, "could not open \"%s\" file to read.\n", fname);
		return 0;
	}
	uint8_t *pix = 0;
	const char *complaint = 0;
	int truncated = 0;
	const int magic0 = fgetc(k.file), magic1 = fgetc(k.file);
	long field[3] = {0, 0, 0}; /* width, height, maxval */
	if (magic0 != 'P' || (magic1 != '5' && magic1 != '6')) {
		complaint = "file \"%s\" neither P5 nor P6 image.\n";
	} else if (!advance(&k) || !header_field(&k, &field[0]) || !header_field(&k, &field[1]) || !header_field(&k, &field[2])) {
		truncated = 1;
	} else if (!field[0] || !field[1] || !field[2]) {
		complaint = "could not read image file \"%s\".\n";
	} else if (field[2] != 255) { /* pnm.h:63-67 */
		complaint = "cant read \"%s\", only 8 bit per channel SRGB supported at the moment.\n";
	} else if (field[0] > 0x7fffffffL || field[1] > 0x7fffffffL) {
		complaint = "could not read image file \"%s\".\n";
	} else {
		const int ch = magic1 == '5' ? 1 : 3;
		const size_t total = (size_t)field[0] * (size_t)field[1] * ch;
		pix = malloc(total ? total : 1);
		if (!pix || fread(pix, 1, total, k.file) != total) {
			truncated = 1;
		} else {
			*width = (int)field[0];
			*height = (int)field[1];
			*channels = ch;
		}
	}
	if (truncated)
		fprintf(stderr, "EOF while reading from \"%s\".\n", fname);
	else if (complaint)
		fprintf(stderr, complaint, fname);
	fclose(k.file);
	if (truncated || complaint) {
		free(pix);
		return 0;
	}
	return pix;
}

int dwt_write_pnm(const char *name, const uint8_t *pixels, int width, int height, int channels)
{
	const char *fname = real_name(name, "/dev/stdout");
	FILE *file = fopen(fname, "w");
	if (!file) {
		fprintf(stderr, "could not open \"%s\" file to write.\n", fname);
		return 0;
	}
	size_t total = (size_t)width * height * channels;
	if (fprintf(file, "P%d %d %d 255\n", channels == 1 ? 5 : 6, width, height) < 0 ||
	    fwrite(pixels, 1, total, file) != total) {
		fprintf(stderr, "EOF while writing to \"%s\".\n", fname);
		fclose(file);
		return 0;
	}
	fclose(file);
	return 1;
}

uint8_t *dwt_read_file(const char *name, size_t *len)
{
	const char *fname = real_name(name, "/dev/stdin");
	FILE *file = fopen(fname, "r");
	if (!file) {
		fprintf(stderr, "could not open \"%s\" file to read\n", fname);
		return 0;
	}
	size_t room = 1 << 20, n = 0;
	uint8_t *buf = malloc(room);
	for (;;) {
		size_t got = buf ? fread(buf + n, 1, room - n, file) : 0;
		n += got;
		if (n < room)
			break;
		room *= 2;
		uint8_t *nb = realloc(buf, room);
		if (!nb) {
			free(buf);
			buf = 0;
			break;
		}
		buf = nb;
	}
	fclose(file);
	*len = n;
	return buf;
}
