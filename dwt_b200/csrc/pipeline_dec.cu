// pipeline_dec.cu -- placeholder until the decoder lands
#include "pipeline.cuh"
#include "dwt_b200.h"
extern "C" int dwt_ctx_upload_stream(dwt_ctx *, const uint8_t *, size_t) { dwt_set_error("decoder not built yet"); return -1; }
extern "C" int dwt_ctx_decode_resident(dwt_ctx *, int, struct dwt_stats *) { dwt_set_error("decoder not built yet"); return -1; }
extern "C" int dwt_ctx_download_image(dwt_ctx *, uint8_t **, int *, int *, int *) { dwt_set_error("decoder not built yet"); return -1; }
extern "C" int dwt_decode(dwt_ctx *, const uint8_t *, size_t, int, uint8_t **, int *, int *, int *, struct dwt_stats *) { dwt_set_error("decoder not built yet"); return -1; }
