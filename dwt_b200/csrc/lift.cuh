// lift.cuh -- launch interface of the colour + CDF 5/3 lifting kernels (lift.cu)
#pragma once
#include "common.cuh"

// One 2-D lifting level.  Device layouts:
//   planar int32 LL buffers   : [channel][row][col], pitch = width of that level
//   Mallat pyramid (details)  : planar int32 [channel][H_full][W_full]; the detail bands of the level
//                               whose input is W x H live at the reference's Mallat positions
//                               (x >= ceil(W/2) and/or y >= ceil(H/2)), cf. encode.c:16-30
struct LiftLevel {
	const void *in;            // forward: u8 image (first level) or int32 planar LL; inverse: int32 planar LL (coarse)
	long long in_chan_stride;  // elements between channels of `in` (planar modes)
	int in_pitch;              // elements per row of `in` (pixels for interleaved u8)
	void *out;                 // forward: int32 planar LL (ceil(W/2) x ceil(H/2)); inverse: int32 planar (W x H) or u8 image
	long long out_chan_stride;
	int out_pitch;
	int *pyr;                  // Mallat pyramid
	long long pyr_chan_stride;
	int pyr_pitch;
	int W, H;                  // size of the fine side of this level
	int channels;
	int *maxabs;               // forward only: per-channel max |detail| (atomicMax)
	int *work;                 // zeroed device counter: work items are handed out dynamically
	int chained = 0;           // 1 = the previous launch on the stream produced this level's input: launch as a
	                           // programmatic dependent (its prologue overlaps the producer's drain)
};

// mode: 0 = u8 interleaved RGB with the colour transform fused (image.h:53-65 / 34-51),
//       1 = u8 gray, 2 = int32 planar
int lift_forward_level(const LiftLevel &lv, int mode, cudaStream_t st, long long *launches);
int lift_inverse_level(const LiftLevel &lv, int mode, cudaStream_t st, long long *launches);

// The remaining levels of a channel fused into one CTA (image resident in shared memory).  W x H is the size of
// the finest fused level; nlev levels lead from / to the (W >> nlev) x (H >> nlev) image (halving with ceil).
struct LiftTail {
	const int *ll_in;          // forward: planar W x H LL; inverse: planar root
	long long in_chan_stride;
	int in_pitch;
	int *ll_out;               // forward: planar root; inverse: planar W x H
	long long out_chan_stride;
	int out_pitch;
	int *pyr;
	long long pyr_chan_stride;
	int pyr_pitch;
	int W, H, nlev, channels;
	int *maxabs;               // forward only
	int chained = 0;           // as LiftLevel::chained
};
bool lift_tail_fits(int W, int H); // shared memory and lane budget of the fused kernel
int lift_tail(const LiftTail &t, bool inverse, cudaStream_t st, long long *launches);

// generic strided 1-D lifting on device buffers (cdf53.h:9-34 / 36-61 semantics, all CH lanes in parallel)
int lift_cdf53_1d(int *d_out, int *d_in, int N, int SO, int SI, int CH, bool inverse, cudaStream_t st);
int lift_colour(int *d_buf, int total, bool inverse, cudaStream_t st);
