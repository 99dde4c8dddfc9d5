/*
 * decode -- drop-in for the reference decoder program (decode.c:136-268): same argv and exit codes; the
 * codec itself runs on the GPU through libdwt_b200.
 *
 *   decode input.dwt output.pnm [PIXELS]          ("-" = stdin / stdout)
 *
 * Exit 1 without output on a bad magic, short header, size < 8 or EOF inside the root image / plane counts
 * (decode.c:143-159,180-186); otherwise a PNM is always written, possibly at a lower resolution when the
 * stream was truncated (decode.c:251-255).
 */
#include "dwt_b200.h"
#include "pnm.h"

#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv)
{
	if (argc < 3 || argc > 4) {
		fprintf(stderr, "usage: %s input.dwt output.pnm [PIXELS]\n", argv[0]);
		return 1;
	}
	size_t len = 0;
	uint8_t *stream = dwt_read_file(argv[1], &len);
	if (!stream)
		return 1;
	int pixels_max = argc >= 4 ? atoi(argv[3]) : -1;
	if (argc >= 4 && pixels_max < 0)
		pixels_max = 0; /* a negative PIXELS behaves like 0 in decode.c:165-171 */
	const char *dev = getenv("DWT_DEVICE");
	dwt_ctx *ctx = dwt_ctx_create(dev ? atoi(dev) : -1);
	if (!ctx) {
		fprintf(stderr, "%s: %s\n", argv[0], dwt_last_error());
		return 1;
	}
	uint8_t *pixels = 0;
	int width, height, channels;
	struct dwt_stats st;
	int r = dwt_decode(ctx, stream, len, pixels_max, &pixels, &width, &height, &channels, &st);
	if (r) {
		if (r < 0)
			fprintf(stderr, "%s: %s\n", argv[0], dwt_last_error());
		else if (r == 1) /* the stream ended inside header, root image or plane counts: bytes.h:99-103 prints this */
			fprintf(stderr, "reached end of file \"%s\"\n", argv[1]);
		/* r == 2: wrong magic or size below 8 -- the reference exits 1 silently (decode.c:146-159) */
		return 1;
	}
	if (!dwt_write_pnm(argv[2], pixels, width, height, channels))
		return 1;
	dwt_free(pixels);
	free(stream);
	dwt_ctx_destroy(ctx);
	return 0;
}
