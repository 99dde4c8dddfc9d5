// hilbert.cuh -- Hilbert-order linearisation of the Mallat pyramid into the bit-sliced coefficient store
// (encode.c:32-58 + encode.c:112-131) and its inverse (decode.c:32-65 + decode.c:102-117).
#pragma once
#include "common.cuh"

// Geometry-only plan, built on the host once per image size: for every detail level the curve is cut into aligned
// cells of cs x cs positions (cs = min(32, side)).
//   cell_base[cell_off[l] + q]  number of valid positions (inside the level's w x h domain, outside its LL
//                               rectangle) that precede cell q on the curve -- the closed-form rank of SURVEY.md
//                               App. C.3 evaluated per cell
//   cell_info[...]              cell column | cell row << 12 | orientation << 24: inside a cell the curve is the
//                               32 x 32 base curve transposed (bit 0) and / or point-reflected (bit 1)
//   full_list / part_list       cells whose 1024 positions are all valid (warp-per-cell fast kernels) and cells cut
//                               by the image or LL boundary (generic kernel); empty cells are in neither.
//                               Entry = level << 28 | q, all levels back to back.
struct HilbertPlan {
	int cell_off[DWT_MAX_LEVELS + 1];
	int ncell[DWT_MAX_LEVELS];
	int cs[DWT_MAX_LEVELS];
	int full_off[DWT_MAX_LEVELS + 1], part_off[DWT_MAX_LEVELS + 1]; // list ranges per level
	u32 *cell_base; // device (one allocation holds all the arrays)
	u32 *cell_info;
	u32 *full_list, *part_list;
	unsigned short *tile_lut; // 4 orientations x 1024: word offset of curve index 32 k + i inside a TMA-staged (128-byte
	                          // swizzled) 32 x 32 cell, stored at [orientation][i][k] (hilbert.cu: the TMA kernels)
};

int hilbert_plan_build(const Geom &g, HilbertPlan *plan, cudaStream_t st, long long *launches);
void hilbert_plan_free(HilbertPlan *plan);

// pyramid (planar int32 [c][H][W], Mallat layout) -> bit-sliced store.
// store word (c, p, gAll) = bs[bsbase[c] + p * GT + gAll]; plane index planes[c] holds the sign bits.
int hilbert_linearize(const Geom &g, const HilbertPlan &plan, const Sched &s, const int *pyr,
                      long long pyr_chan_stride, int pyr_pitch, u32 *bs, int levels_used, cudaStream_t st,
                      long long *launches);

// bit-sliced store -> pyramid, adding the dequantisation bias of decode.c:51-58 (missing[c*16+l]).
// Only detail levels < levels_used are written; the root is written by the caller.
int hilbert_reconstruct(const Geom &g, const HilbertPlan &plan, const Sched &s, const u32 *bs,
                        const int *missing_dev /* 48 ints */, int *pyr, long long pyr_chan_stride, int pyr_pitch,
                        int levels_used, cudaStream_t st, long long *launches);

// bit-sliced store -> planar two's-complement coefficients (tests / debugging): out[c * total + pix[0] + k]
int hilbert_unslice(const Geom &g, const Sched &s, const u32 *bs, int *planar, long long total, cudaStream_t st);
