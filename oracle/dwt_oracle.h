/*
 * dwt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99, single-threaded) of the xdsopl/dwt encode/decode path.
 * It exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the
 * CUDA product path stage by stage.  Nothing in dwt_b200/ (the product) includes, links or calls it.
 *
 * Parity pin: this restatement is checked byte-for-byte against the unmodified reference programs
 * (oracle/_ref/encode, oracle/_ref/decode, built from /root/reference by oracle/Makefile) and against
 * the sha256 pins of SURVEY.md App. E.1 in tests/test_oracle.py.
 */
#ifndef DWT_ORACLE_H
#define DWT_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* level geometry: utils.h:17-40.  arrays need 16 entries.  returns levels. */
int orc_geometry(int w, int h, int *lengths, int *pixels, int *widths, int *heights);

/* colour: image.h:53-65 (forward), image.h:34-51 (inverse, clamps first). interleaved int triples */
void orc_rgb_to_ycocg(int *buf, int npix);
void orc_ycocg_to_rgb(int *buf, int npix);

/* 1-D lifting: cdf53.h:9-34 / 36-61 (same argument meaning as the reference) */
void orc_cdf53(int *out, int *in, int N, int SO, int SI, int CH);
void orc_icdf53(int *out, int *in, int N, int SO, int SI, int CH);

/* multi-level 2-D drivers: encode.c:16-30 / decode.c:16-30.
 * interleaved int[h][w][ch] buffers with row stride w*ch; `in` is scratch and is clobbered. */
void orc_forward2d(int *out, int *in, int w, int h, int ch);
void orc_inverse2d(int *out, int *in, int w, int h, int ch);

/* Hilbert d -> (x,y): hilbert.h:15-34 */
void orc_hilbert(int n, int d, int *x, int *y);

/* linearisation encode.c:32-58: interleaved pyramid -> planar per-channel arrays (ch * w*h ints) */
void orc_linearize(int *planar, const int *pyramid, int w, int h, int ch);

struct orc_stats {
	long long meta_bits, root_bits, total_bits; /* the three stderr counters encode.c:175-176,179-180,226-230 */
	long long bytes;                            /* bytes actually written                                   */
	int planes[3];
	int levels;
};

/* Stage dump of the encoder front end (everything before the bit-plane coder).
 * pyramid  : ch*w*h ints, interleaved Mallat pyramid after transformation()   (may be NULL)
 * planar   : ch*w*h ints, planar linearised two's-complement coefficients     (may be NULL)
 * planes   : per channel plane counts (encode.c:163-165)                      */
int orc_front_end(const uint8_t *pix, int w, int h, int ch, int *pyramid, int *planar, int *planes);

/* Whole encoder, memory to memory.  capacity <= 0: unlimited (bytes.h:77).
 * out must hold out_room bytes; returns bytes written or -1 if out_room was too small / bad args. */
long long orc_encode(const uint8_t *pix, int w, int h, int ch, int capacity,
                     uint8_t *out, long long out_room, struct orc_stats *st);

/* Whole decoder, memory to memory.  pixels_max < 0: no PIXELS argument (decode.c:165-171).
 * On success returns 0 and a malloc()ed u8 image (free with orc_free); returns 1 where the reference
 * program exits 1 without output (bad magic, short header, EOF inside root / planes). */
int orc_decode(const uint8_t *stream, long long len, int pixels_max,
               uint8_t **pix, int *w, int *h, int *ch);

/* decoder stage dump: coefficients after process() in planar order + missing[] + level reached */
int orc_decode_coeffs(const uint8_t *stream, long long len, int pixels_max,
                      int **planar /* ch * total ints, malloc */, int *missing /*48*/, int *level,
                      int *w, int *h, int *ch);

void orc_free(void *p);

/* synthetic generator of SURVEY.md App. E.2.  kind 0 = photo, 1 = noise.  out: w*h*3 bytes */
void orc_synth(uint8_t *out, int w, int h, int kind, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
