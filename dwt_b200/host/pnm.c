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

uint8_t *dwt_read_pnm(const char *name, int *width, int *height, int *channels)
{
	const char *fname = real_name(name, "/dev/stdin");
	FILE *file = fopen(fname, "r");
	if (!file) {
		fprintf(stderr, "could not open \"%s\" file to read.\n", fname);
		return 0;
	}
	int letter = fgetc(file), number = fgetc(file);
	if (letter != 'P' || (number != '5' && number != '6')) {
		fprintf(stderr, "file \"%s\" neither P5 nor P6 image.\n", fname);
		fclose(file);
		return 0;
	}
	int ch = number == '5' ? 1 : 3;
	int integer[3];
	uint8_t *pix = 0;
	int c = fgetc(file);
	if (c == EOF)
		goto eof;
	for (int i = 0; i < 3; ++i) {
		/* the reference looks at the next byte for '#': comment lines are skipped (pnm.h:38-41) */
		while ((c = fgetc(file)) == '#')
			while ((c = fgetc(file)) != '\n')
				if (c == EOF)
					goto eof;
		while (c < '0' || c > '9')
			if ((c = fgetc(file)) == EOF)
				goto eof;
		char str[16];
		int n = 0;
		while (c >= '0' && c <= '9' && n < 15) {
			str[n++] = (char)c;
			if ((c = fgetc(file)) == EOF)
				goto eof;
		}
		str[n] = 0;
		integer[i] = atoi(str);
	}
	if (!(integer[0] && integer[1] && integer[2])) {
		fprintf(stderr, "could not read image file \"%s\".\n", fname);
		fclose(file);
		return 0;
	}
	if (integer[2] != 255) {
		fprintf(stderr, "cant read \"%s\", only 8 bit per channel SRGB supported at the moment.\n", fname);
		fclose(file);
		return 0;
	}
	{
		size_t total = (size_t)integer[0] * integer[1] * ch;
		pix = malloc(total ? total : 1);
		if (!pix || fread(pix, 1, total, file) != total)
			goto eof;
	}
	fclose(file);
	*width = integer[0];
	*height = integer[1];
	*channels = ch;
	return pix;
eof:
	fprintf(stderr, "EOF while reading from \"%s\".\n", fname);
	fclose(file);
	free(pix);
	return 0;
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
