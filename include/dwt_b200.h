/*
 * dwt_b200.h -- C ABI of libdwt_b200.so: a B200 (sm_100a) implementation of the xdsopl/dwt
 * encode/decode hot path.  Plain pointers and sizes only; no CUDA or torch types in any signature.
 *
 * Every entry point cites the reference interface it replaces (file:line relative to the reference
 * tree).  The reference has no library boundary of its own -- its functions are defined in headers and
 * compiled into the two programs -- so the boundary is: (1) the coarse codec calls the two `main`s are
 * made of, (2) the cdf53 / rle / vli / bits / bytes entry points with their reference signatures, and
 * (3) the .dwt wire format.  INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * There is NO CPU fallback: every codec call fails with -1 when no CUDA device is usable.
 */
#ifndef DWT_B200_H
#define DWT_B200_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ coarse codec API
 * replaces the body of main() in encode.c:133-232 and decode.c:136-268 (everything between
 * read_pnm()/bytes_reader() and write_pnm()/close_bytes_writer()). */

struct dwt_stats {
	long long meta_bits;  /* "%d bits for meta data"   encode.c:175-176 */
	long long root_bits;  /* "%d bits for root image"  encode.c:179-180 */
	long long total_bits; /* "%d bits (%d KiB) encoded" encode.c:226-230 (first number)  */
	long long kib;        /*                                    (second number)          */
	long long full_bits;  /* size of the untruncated stream in bits (header included); a lower bound (still behind
	                         the capacity) when a capacity let the encoder skip the chunks that cannot reach the output */
	int levels;
	int planes[3];
	/* device time of the last call, milliseconds (CUDA events on the context's stream) */
	float ms_h2d, ms_lift, ms_linearize, ms_coder, ms_d2h, ms_total;
	/* decoder only */
	int level_reached;    /* `level` of decode.c:197: highest detail level started (-1: root only) */
	/* decoder parse statistics: windows up to the end of each pass, chain jumps, exact slice steps */
	long long parse_windows, parse_jumps, parse_exact;
};

typedef struct dwt_ctx dwt_ctx;

/* One context = one CUDA device + one stream + reusable device/pinned buffers.  device < 0: current. */
dwt_ctx *dwt_ctx_create(int device);
void dwt_ctx_destroy(dwt_ctx *ctx);
const char *dwt_last_error(void);

/* Encode an 8-bit image (channels 1 = 'W5' gray, 3 = 'W6' RGB, interleaved, row-major) to a .dwt stream.
 * capacity <= 0: unlimited (bytes.h:77); otherwise the result is exactly the first `capacity` bytes of
 * the unlimited stream, like the reference's byte sink.  *out is malloc()ed; release with dwt_free().
 * Returns 0, or -1 where the reference exits 1 (width/height < 8 or > 65536: encode.c:140-146) and on
 * CUDA errors.  Hitting the capacity is not an error (the reference exits 0). */
int dwt_encode(dwt_ctx *ctx, const uint8_t *pixels, int width, int height, int channels, int capacity,
               uint8_t **out, size_t *out_len, struct dwt_stats *stats);

/* Decode a (possibly truncated) .dwt stream.  pixels_max < 0: no PIXELS argument (decode.c:165-171).
 * Returns 0 and a malloc()ed interleaved u8 image whose size may be smaller than the coded size when
 * the stream was truncated (decode.c:251-255).  Where the reference exits 1 without output the return value is
 * positive: 1 when the stream ended inside the header, the root image or the plane counts (the reference's
 * get_byte prints "reached end of file" there, bytes.h:99-103), 2 when the magic is wrong or a dimension is
 * below 8 (decode.c:146-159: silent).  -1 on CUDA errors. */
int dwt_decode(dwt_ctx *ctx, const uint8_t *stream, size_t len, int pixels_max,
               uint8_t **pixels, int *width, int *height, int *channels, struct dwt_stats *stats);

void dwt_free(void *p);

/* Device-resident variants used by the benchmark: the image / stream is uploaded once, the kernels run
 * on data already in HBM, and results stay in HBM until downloaded.  Same return conventions.
 * dwt_ctx_upload_image() only queues its copy: the caller's buffer must stay untouched until the next
 * dwt_ctx_encode_resident() or dwt_ctx_sync() on the context has returned (dwt_ctx_upload_stream() waits
 * for its copy itself). */
int dwt_ctx_upload_image(dwt_ctx *ctx, const uint8_t *pixels, int width, int height, int channels);
int dwt_ctx_encode_resident(dwt_ctx *ctx, int capacity, struct dwt_stats *stats);
int dwt_ctx_download_stream(dwt_ctx *ctx, uint8_t **out, size_t *out_len);
int dwt_ctx_upload_stream(dwt_ctx *ctx, const uint8_t *stream, size_t len);
int dwt_ctx_decode_resident(dwt_ctx *ctx, int pixels_max, struct dwt_stats *stats);
int dwt_ctx_download_image(dwt_ctx *ctx, uint8_t **pixels, int *width, int *height, int *channels);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long dwt_ctx_launch_count(const dwt_ctx *ctx);
int dwt_ctx_sync(dwt_ctx *ctx);
/* number of contexts the caller keeps busy on this device at the same time (default 1; dwt_pool sets its worker count):
 * with several frames in flight the library prefers kernels that do less total work over lower single-frame latency */
int dwt_ctx_set_in_flight(dwt_ctx *ctx, int contexts);
/* testing aid: pin the decoder's stream-scan kernel.  0 = choose by stream size and the in-flight hint (default),
 * 1 = one CTA per scan window (lowest latency), 2 = one thread per window and chain (least work).  Both produce the
 * same tables; the parity tests run every corrupted stream through both. */
int dwt_ctx_set_decoder_scan(dwt_ctx *ctx, int mode);

/* Caller-owned buffer variants of dwt_encode / dwt_decode (same semantics and return values; -1 with
 * *out_len = needed size when the buffer is too small).  With buffers from dwt_host_alloc() (page-locked
 * memory) the host<->device copies run at PCIe rate; this is what the CLIs' file buffers use. */
void *dwt_host_alloc(size_t bytes);
void dwt_host_free(void *p);
int dwt_encode_into(dwt_ctx *ctx, const uint8_t *pixels, int width, int height, int channels, int capacity,
                    uint8_t *out, size_t out_room, size_t *out_len, struct dwt_stats *stats);
int dwt_decode_into(dwt_ctx *ctx, const uint8_t *stream, size_t len, int pixels_max, uint8_t *pixels,
                    size_t pixels_room, int *width, int *height, int *channels, struct dwt_stats *stats);
/* Batches of independent images (the reference codes one image per process): a pool owns `workers`
 * contexts on one device and codes the items on as many host threads, so host<->device copies and kernels of different
 * items overlap.  Buffers should come from dwt_host_alloc().  status per item: the value dwt_encode_into /
 * dwt_decode_into would have returned.  Return value: number of items whose status is non-zero, -1 on bad arguments.
 * The contexts of a device take turns for host<->device copies of 16 MB and more (one per direction at a time, chained
 * with events on the device, never a lock held across a wait), so the first items of a batch start computing after one
 * copy time.  Waiting host threads poll for ~20 us and then sleep (cudaEventBlockingSync).  Environment (debugging
 * aids): DWT_XFER_GATE=0 switches the turn-taking off, DWT_SYNC=spin makes waiting threads spin in the driver,
 * DWT_SPIN_US=<n> sets the poll time. */
struct dwt_encode_item {
	const uint8_t *pixels;
	int width, height, channels, capacity;
	uint8_t *out;
	size_t out_room, out_len;
	int status;
};
struct dwt_decode_item {
	const uint8_t *stream;
	size_t len;
	int pixels_max; /* < 0: none */
	uint8_t *pixels;
	size_t pixels_room;
	int width, height, channels;
	int status;
};
typedef struct dwt_pool dwt_pool;
dwt_pool *dwt_pool_create(int device, int workers);
/* The same over several GPUs of one box (images are independent, no exchange step, no collective):
 * `workers` contexts and host threads on each of the n_devices devices; item i of a batch is coded on
 * devices[i mod n_devices].  devices == NULL or n_devices <= 0: every visible device.  dwt_pool_workers() then
 * returns the total number of contexts, dwt_pool_devices() the device list (returns its length). */
dwt_pool *dwt_pool_create_multi(const int *devices, int n_devices, int workers);
int dwt_pool_devices(const dwt_pool *pool, int *devices, int room);
void dwt_pool_destroy(dwt_pool *pool);
int dwt_pool_workers(const dwt_pool *pool);
/* why the most recent failed item of the pool failed (dwt_last_error() is per thread and the items run on the
 * pool's own threads); "" when nothing has failed yet.  Valid until the next batch call on the pool. */
const char *dwt_pool_last_error(const dwt_pool *pool);
int dwt_pool_encode(dwt_pool *pool, struct dwt_encode_item *items, int n);
int dwt_pool_decode(dwt_pool *pool, struct dwt_decode_item *items, int n);
/* both kinds of item in one call, interleaved on the workers (pixel uploads overlap pixel downloads) */
int dwt_pool_run(dwt_pool *pool, struct dwt_encode_item *enc, int n_enc, struct dwt_decode_item *dec, int n_dec);

/* benchmark helpers: evict L2 (writes 256 MB), CUDA events on the context's stream (slots 0..3) */
int dwt_ctx_flush_l2(dwt_ctx *ctx);
int dwt_ctx_event_record(dwt_ctx *ctx, int slot);
float dwt_ctx_event_elapsed_ms(dwt_ctx *ctx, int slot_a, int slot_b);
/* ctx's stream waits for the work queued so far on other's stream (several contexts timed as one region) */
int dwt_ctx_wait_for(dwt_ctx *ctx, dwt_ctx *other);

/* ------------------------------------------------------------------ transform entry points */

/* cdf53.h:9 and cdf53.h:36 -- exact reference signatures and semantics: host buffers, strides in ints,
 * cdf53() clobbers `in` with the lifted (not yet deinterleaved) samples, icdf53() leaves `in` alone.
 * Both run on the GPU of the calling thread's default context. */
void cdf53(int *out, int *in, int N, int SO, int SI, int CH);
void icdf53(int *out, int *in, int N, int SO, int SI, int CH);

/* `transformation` of encode.c:16-30 / decode.c:16-30 (the two programs define different functions of
 * the same name, so the library names them apart).  Interleaved int[H][W][CH] host buffers, row stride
 * W*CH, min_len = 8.  `in` is left unchanged (the reference uses it as scratch). Returns 0 / -1. */
int dwt_forward(int *out, const int *in, int W, int H, int CH);
int dwt_inverse(int *out, const int *in, int W, int H, int CH);

/* colour transforms image.h:67-79 on interleaved int triples (host buffers) */
int dwt_ycocg_from_rgb(int *buffer, int total);
int dwt_rgb_from_ycocg(int *buffer, int total);

/* level geometry utils.h:28-40; arrays of 16 ints; returns levels */
int compute_lengths(int *lengths, int *pixels, int *widths, int *heights, int W, int H, int N0);
int ilog2(int x);

/* debugging / parity taps (tests only): stages of the encoder front end in the reference's layouts */
int dwt_debug_front_end(dwt_ctx *ctx, const uint8_t *pixels, int width, int height, int channels,
                        int *pyramid /* interleaved int[h][w][ch] or NULL */,
                        int *planar /* ch * w*h linearised coefficients or NULL */, int *planes /* 3 */);

/* ------------------------------------------------------------------ stream entry points
 * Same names, signatures, ownership and error codes as the reference headers (0 / -1 I/O / -2 capacity).
 * They are the host-side serial sinks/sources the CLI uses for the header, root image and plane counts;
 * the bit-plane payload itself is produced / consumed on the GPU. */

struct bytes_reader;
struct bytes_writer;
struct bits_reader;
struct bits_writer;
struct vli_reader;
struct vli_writer;
struct rle_reader;
struct rle_writer;

/* bytes.h:23-118 */
struct bytes_reader *bytes_reader(char *name);
struct bytes_writer *bytes_writer(char *name, int capacity);
int bytes_count(struct bytes_writer *bytes);
void close_bytes_reader(struct bytes_reader *bytes);
void close_bytes_writer(struct bytes_writer *bytes);
int put_byte(struct bytes_writer *bytes, int b);
int write_bytes(struct bytes_writer *bytes, int b, int n);
int get_byte(struct bytes_reader *bytes);
int read_bytes(struct bytes_reader *bytes, int *b, int n);
/* memory-backed variants (not in the reference): the sink grows with realloc, the source borrows */
struct bytes_writer *bytes_writer_mem(int capacity);
const uint8_t *bytes_writer_data(struct bytes_writer *bytes, size_t *len);
struct bytes_reader *bytes_reader_mem(const uint8_t *data, size_t len);

/* bits.h:23-106 */
struct bits_reader *bits_reader(struct bytes_reader *bytes);
struct bits_writer *bits_writer(struct bytes_writer *bytes);
int bits_count(struct bits_writer *bits);
void close_bits_reader(struct bits_reader *bits);
void close_bits_writer(struct bits_writer *bits);
int put_bit(struct bits_writer *bits, int b);
int write_bits(struct bits_writer *bits, int b, int n);
int get_bit(struct bits_reader *bits);
int read_bits(struct bits_reader *bits, int *b, int n);

/* vli.h:21-101 */
struct vli_reader *vli_reader(struct bits_reader *bits);
struct vli_writer *vli_writer(struct bits_writer *bits);
void delete_vli_reader(struct vli_reader *vli);
void delete_vli_writer(struct vli_writer *vli);
int vli_put_bit(struct vli_writer *vli, int bit);
int vli_get_bit(struct vli_reader *vli);
int vli_write_bits(struct vli_writer *vli, int b, int n);
int vli_read_bits(struct vli_reader *vli, int *b, int n);
int put_vli(struct vli_writer *vli, int val);
int get_vli(struct vli_reader *vli);

/* rle.h:21-103 */
struct rle_reader *rle_reader(struct vli_reader *vli);
struct rle_writer *rle_writer(struct vli_writer *vli);
int rle_flush(struct rle_writer *rle);
void delete_rle_reader(struct rle_reader *rle);
void delete_rle_writer(struct rle_writer *rle);
int put_rle(struct rle_writer *rle, int b);
int get_rle(struct rle_reader *rle);
int rle_put_bit(struct rle_writer *rle, int bit);
int rle_get_bit(struct rle_reader *rle);

#ifdef __cplusplus
}
#endif
#endif
