/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY.  The reference's header-only modules compiled, unmodified and where
 * they lie (-I/root/reference), into oracle/_ref/libref_shim.so, so that tests can call the reference's own
 *   cdf53 / icdf53                                   cdf53.h:9,36
 *   bytes_* / bits_* / vli_* / rle_* entry points    bytes.h:23-118, bits.h:23-106, vli.h:21-101, rle.h:21-103
 *   ilog2 / compute_lengths / hilbert                utils.h:9-40, hilbert.h:15-34
 *   rgb2ycocg / ycocg2rgb                            image.h:39-65
 * function by function, next to the same-named entry points of libdwt_b200.so (each library is loaded with its
 * own ctypes handle, so the identical names do not clash).  Nothing of the reference is copied into this
 * repository: this file only names the headers.  Built by oracle/Makefile when /root/reference exists; the
 * product never links or loads it.
 */
#include "image.h"
#include "cdf53.h"
#include "utils.h"
#include "hilbert.h"
#include "rle.h"

/* hilbert() returns a struct by value: a plain-pointer wrapper for ctypes */
void ref_hilbert_xy(int n, int d, int *x, int *y)
{
	struct position p = hilbert(n, d);
	*x = p.x;
	*y = p.y;
}
