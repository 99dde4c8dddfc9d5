/*
 * dwtbatch -- many images through one process (the reference programs code one image per process,
 * encode.c:133-232 / decode.c:136-268).  Every file is coded exactly as the drop-in `encode` / `decode` would code
 * it -- same streams, same pixels, same acceptance rules -- but the images share a pool of GPU contexts, so file
 * I/O, host<->device copies and kernels of different images overlap.
 *
 *   dwtbatch encode [-c CAPACITY] [-j WORKERS] [-g GPUS] OUTDIR [FILE.pnm ...]     -> OUTDIR/<name>.dwt
 *   dwtbatch decode [-p PIXELS]   [-j WORKERS] [-g GPUS] OUTDIR [FILE.dwt ...]     -> OUTDIR/<name>.pnm
 *
 * -g all | N | a,b,c : the GPUs of this box to shard the files over (file i -> GPU i mod G, WORKERS contexts on each;
 * images are independent, so there is no exchange between the GPUs).  Default: one GPU (DWT_DEVICE or the current one).
 *
 * Without FILE arguments the paths are read from stdin, one per line.  Exit status: 0 when every file was coded,
 * 1 otherwise (each failure is reported on stderr and the other files are still coded; no output file is created
 * for a rejected input, like the reference).  Environment: DWT_DEVICE=<n> selects the CUDA device.
 */
#include "dwt_b200.h"
#include "pnm.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct slot { /* page-locked staging of one item, reused from wave to wave */
	uint8_t *in, *out;
	size_t in_room, out_room;
};

static int slot_room(uint8_t **p, size_t *room, size_t need)
{
	if (*room >= need)
		return 0;
	if (*p)
		dwt_host_free(*p);
	*room = need + need / 8 + 4096;
	*p = (uint8_t *)dwt_host_alloc(*room);
	if (!*p) {
		*room = 0;
		fprintf(stderr, "dwtbatch: %s\n", dwt_last_error());
		return -1;
	}
	return 0;
}

/* OUTDIR/<basename of path without its extension><ext> */
static char *out_name(const char *dir, const char *path, const char *ext)
{
	const char *base = strrchr(path, '/');
	base = base ? base + 1 : path;
	const char *dot = strrchr(base, '.');
	size_t stem = dot && dot != base ? (size_t)(dot - base) : strlen(base);
	char *name = (char *)malloc(strlen(dir) + 1 + stem + strlen(ext) + 1);
	if (name)
		sprintf(name, "%s/%.*s%s", dir, (int)stem, base, ext);
	return name;
}

static int write_file(const char *name, const uint8_t *data, size_t len)
{
	FILE *f = fopen(name, "w");
	if (!f) {
		fprintf(stderr, "could not open \"%s\" file to write\n", name);
		return 0;
	}
	int ok = !len || fwrite(data, 1, len, f) == len;
	if (!ok)
		fprintf(stderr, "could not write to file \"%s\"\n", name);
	fclose(f);
	return ok;
}

static char **stdin_paths(int *n)
{
	char **list = 0, line[4096];
	int cap = 0;
	*n = 0;
	while (fgets(line, sizeof(line), stdin)) {
		size_t l = strlen(line);
		while (l && (line[l - 1] == '\n' || line[l - 1] == '\r'))
			line[--l] = 0;
		if (!l)
			continue;
		if (*n == cap) {
			cap = cap ? 2 * cap : 64;
			list = (char **)realloc(list, sizeof(char *) * cap);
			if (!list)
				return 0;
		}
		char *copy = (char *)malloc(l + 1);
		if (!copy)
			return 0;
		memcpy(copy, line, l + 1);
		list[(*n)++] = copy;
	}
	return list;
}

int main(int argc, char **argv)
{
	const char *usage = "usage: %s encode [-c CAPACITY] [-j WORKERS] [-g all|N|a,b,..] OUTDIR [FILE.pnm ...]\n"
	                    "       %s decode [-p PIXELS] [-j WORKERS] [-g all|N|a,b,..] OUTDIR [FILE.dwt ...]\n";
	if (argc < 3 || (strcmp(argv[1], "encode") && strcmp(argv[1], "decode"))) {
		fprintf(stderr, usage, argv[0], argv[0]);
		return 1;
	}
	const int enc = !strcmp(argv[1], "encode");
	int capacity = 0, pixels_max = -1, workers = 8, a = 2;
	const char *gpus = 0;
	for (; a + 1 < argc && argv[a][0] == '-' && argv[a][1] && !argv[a][2]; a += 2) {
		if (argv[a][1] == 'c' && enc)
			capacity = atoi(argv[a + 1]); /* <= 0: unlimited, encode.c:150-152 */
		else if (argv[a][1] == 'p' && !enc)
			pixels_max = atoi(argv[a + 1]) < 0 ? 0 : atoi(argv[a + 1]); /* decode.c:165-171 */
		else if (argv[a][1] == 'j')
			workers = atoi(argv[a + 1]);
		else if (argv[a][1] == 'g')
			gpus = argv[a + 1];
		else
			break;
	}
	if (a >= argc || workers < 1) {
		fprintf(stderr, usage, argv[0], argv[0]);
		return 1;
	}
	const char *outdir = argv[a++];
	int nfiles = argc - a;
	char **files = argv + a;
	if (nfiles == 0) {
		files = stdin_paths(&nfiles);
		if (!files && nfiles)
			return 1;
	}
	const char *dev = getenv("DWT_DEVICE");
	dwt_pool *pool;
	if (!gpus) {
		pool = dwt_pool_create(dev ? atoi(dev) : -1, workers);
	} else if (!strcmp(gpus, "all")) {
		pool = dwt_pool_create_multi(0, 0, workers);
	} else {
		int list[64], n = 0;
		if (!strchr(gpus, ',')) { /* a count: devices 0 .. N-1 */
			for (int d = 0; d < atoi(gpus) && n < 64; ++d)
				list[n++] = d;
		} else {
			for (const char *q = gpus; *q && n < 64; q = strchr(q, ',') ? strchr(q, ',') + 1 : q + strlen(q))
				list[n++] = atoi(q);
		}
		if (n < 1) {
			fprintf(stderr, usage, argv[0], argv[0]);
			return 1;
		}
		pool = dwt_pool_create_multi(list, n, workers);
	}
	if (!pool) {
		fprintf(stderr, "%s: %s\n", argv[0], dwt_last_error());
		return 1;
	}
	const int wave = 2 * dwt_pool_workers(pool); /* items staged at a time: bounds the page-locked memory */
	struct slot *slots = (struct slot *)calloc(wave, sizeof(struct slot));
	struct dwt_encode_item *ei = (struct dwt_encode_item *)calloc(wave, sizeof(*ei));
	struct dwt_decode_item *di = (struct dwt_decode_item *)calloc(wave, sizeof(*di));
	int *which = (int *)calloc(wave, sizeof(int));
	if (!slots || !ei || !di || !which)
		return 1;
	int failed = 0;
	for (int f0 = 0; f0 < nfiles; f0 += wave) {
		int n = 0;
		for (int f = f0; f < nfiles && f < f0 + wave; ++f) {
			struct slot *s = slots + n;
			if (enc) {
				int w, h, ch;
				uint8_t *px = dwt_read_pnm(files[f], &w, &h, &ch);
				if (!px || w > 65536 || h > 65536 || w < 8 || h < 8) { /* encode.c:139-146 */
					if (px)
						fprintf(stderr, "%s: unsupported image size\n", files[f]);
					free(px);
					++failed;
					continue;
				}
				const size_t raw = (size_t)w * h * ch;
				if (slot_room(&s->in, &s->in_room, raw) || slot_room(&s->out, &s->out_room, raw + raw / 2 + 4096)) {
					free(px);
					++failed;
					continue;
				}
				memcpy(s->in, px, raw);
				free(px);
				memset(ei + n, 0, sizeof(*ei));
				ei[n].pixels = s->in;
				ei[n].width = w;
				ei[n].height = h;
				ei[n].channels = ch;
				ei[n].capacity = capacity;
				ei[n].out = s->out;
				ei[n].out_room = s->out_room;
			} else {
				size_t len = 0;
				uint8_t *st = dwt_read_file(files[f], &len);
				if (!st) {
					++failed;
					continue;
				}
				/* size from the header ('W' '5'|'6', width-1, height-1: decode.c:145-159); a bad header is
				 * rejected by the decoder itself, it only has to get a buffer of some size */
				size_t raw = 4096;
				if (len >= 6 && st[0] == 'W' && (st[1] == '5' || st[1] == '6'))
					raw = (size_t)(1 + st[2] + 256 * st[3]) * (1 + st[4] + 256 * st[5]) * (st[1] == '6' ? 3 : 1);
				if (slot_room(&s->in, &s->in_room, len + 1) || slot_room(&s->out, &s->out_room, raw)) {
					free(st);
					++failed;
					continue;
				}
				memcpy(s->in, st, len);
				free(st);
				memset(di + n, 0, sizeof(*di));
				di[n].stream = s->in;
				di[n].len = len;
				di[n].pixels_max = pixels_max;
				di[n].pixels = s->out;
				di[n].pixels_room = s->out_room;
			}
			which[n++] = f;
		}
		if (!n)
			continue;
		if ((enc ? dwt_pool_encode(pool, ei, n) : dwt_pool_decode(pool, di, n)) < 0) {
			fprintf(stderr, "%s: %s\n", argv[0], dwt_last_error());
			return 1;
		}
		for (int i = 0; i < n; ++i) {
			const char *path = files[which[i]];
			const int status = enc ? ei[i].status : di[i].status;
			if (status) {
				if (status == 1)
					fprintf(stderr, "reached end of file \"%s\"\n", path); /* decode.c:180-186, bytes.h:99-103 */
				else if (status > 1)
					fprintf(stderr, "%s: not a .dwt stream\n", path);   /* decode.c:146-159 (the reference is silent) */
				else
					fprintf(stderr, "%s: coding failed: %s\n", path, dwt_pool_last_error(pool));
				++failed;
				continue;
			}
			char *name = out_name(outdir, path, enc ? ".dwt" : ".pnm");
			int ok = name && (enc ? write_file(name, ei[i].out, ei[i].out_len)
			                      : dwt_write_pnm(name, di[i].pixels, di[i].width, di[i].height, di[i].channels));
			if (!ok)
				++failed;
			free(name);
		}
	}
	for (int i = 0; i < wave; ++i) {
		if (slots[i].in)
			dwt_host_free(slots[i].in);
		if (slots[i].out)
			dwt_host_free(slots[i].out);
	}
	dwt_pool_destroy(pool);
	if (failed)
		fprintf(stderr, "%s: %d of %d files failed\n", argv[0], failed, nfiles);
	return failed ? 1 : 0;
}
