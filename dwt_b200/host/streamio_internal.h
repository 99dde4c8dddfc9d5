/* streamio_internal.h -- accessors the codec pipeline needs on top of the reference-shaped stream API */
#ifndef DWT_STREAMIO_INTERNAL_H
#define DWT_STREAMIO_INTERNAL_H
#include "dwt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int dwt_vli_writer_order(struct vli_writer *vli);  /* adaptive Rice order after the last put_vli */
int dwt_vli_reader_order(struct vli_reader *vli);
long long dwt_bits_reader_position(struct bits_reader *bits); /* stream bits consumed so far */
#ifdef __cplusplus
}
#endif
#endif
