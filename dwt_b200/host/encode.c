/*
 * encode -- drop-in for the reference encoder program (encode.c:133-232): same argv, exit codes, output
 * file and stderr counter lines; the codec itself runs on the GPU through libdwt_b200.
 *
 *   encode input.pnm output.dwt [CAPACITY]        ("-" = stdin / stdout)
 *
 * Environment: DWT_DEVICE=<n> selects the CUDA device (default: current device).
 */
#include "dwt_b200.h"
#include "pnm.h"

#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv)
{
	if (argc != 3 && argc != 4) {
		fprintf(stderr, "usage: %s input.pnm output.dwt [CAPACITY]\n", argv[0]);
		return 1;
	}
	int width, height, channels;
	uint8_t *pixels = dwt_read_pnm(argv[1], &width, &height, &channels);
	if (!pixels || width > 65536 || height > 65536) /* encode.c:139-141 */
		return 1;
	if (width < 8 || height < 8)                    /* encode.c:144-146 */
		return 1;
	int capacity = argc >= 4 ? atoi(argv[3]) : 0;   /* encode.c:150-152 */
	const char *dev = getenv("DWT_DEVICE");
	dwt_ctx *ctx = dwt_ctx_create(dev ? atoi(dev) : -1);
	if (!ctx) {
		fprintf(stderr, "%s: %s\n", argv[0], dwt_last_error());
		return 1;
	}
	uint8_t *stream = 0;
	size_t len = 0;
	struct dwt_stats st;
	if (dwt_encode(ctx, pixels, width, height, channels, capacity, &stream, &len, &st)) {
		fprintf(stderr, "%s: %s\n", argv[0], dwt_last_error());
		return 1;
	}
	/* the output file is opened only after the image was accepted, like bytes_writer() at encode.c:166 */
	const char *fname = (argv[2][0] == '-' && !argv[2][1]) ? "/dev/stdout" : argv[2];
	FILE *file = fopen(fname, "w");
	if (!file) {
		fprintf(stderr, "could not open \"%s\" file to write\n", fname);
		return 1;
	}
	fprintf(stderr, "%d bits for meta data\n", (int)st.meta_bits);
	fprintf(stderr, "%d bits for root image\n", (int)st.root_bits);
	if (len && fwrite(stream, 1, len, file) != len)
		fprintf(stderr, "could not write to file \"%s\"\n", argv[2]);
	fclose(file);
	fprintf(stderr, "%d bits (%d KiB) encoded\n", (int)st.total_bits, (int)st.kib);
	dwt_free(stream);
	free(pixels);
	dwt_ctx_destroy(ctx);
	return 0;
}
