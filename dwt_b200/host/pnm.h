/* pnm.h -- P5/P6 file helpers of the drop-in CLIs (see pnm.c) */
#ifndef DWT_HOST_PNM_H
#define DWT_HOST_PNM_H
#include <stddef.h>
#include <stdint.h>
/* returns malloc()ed interleaved 8-bit pixels or NULL (message on stderr, like read_pnm pnm.h:14-87) */
uint8_t *dwt_read_pnm(const char *name, int *width, int *height, int *channels);
/* returns 1 on success, 0 on failure (like write_pnm pnm.h:89-117) */
int dwt_write_pnm(const char *name, const uint8_t *pixels, int width, int height, int channels);
/* whole file (or stdin for "-") into a malloc()ed buffer */
uint8_t *dwt_read_file(const char *name, size_t *len);
#endif
