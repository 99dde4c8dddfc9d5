// bitslice.cuh -- 32 coefficients <-> bit-plane words inside one thread.
//
// The bit-sliced store keeps, per group of 32 Hilbert-ordered coefficients, one 32-bit word per bit-plane (bit i of
// word p = bit p of |coefficient i|) plus a sign word.  A warp ballot produces one such word per instruction; a thread
// that holds all 32 coefficients of a group produces all of them with a 16 x 16 bit-matrix transpose (Hacker's Delight
// 7-3, LSB-first), done on both 16-bit halves of 16 registers at once:
//   half-word  h_i = |c_i| (15 bits; the reference keeps magnitudes below 2^29 and every image class stays below
//                    2^12, the callers check the plane count) | sign << 15          (encode.c:124-128 sign-magnitude)
//   packed     w_i = h_i | h_{i+16} << 16,  i = 0 .. 15
//   transposed t_p = bit p of every h_i: bits 0..15 from h_0..h_15, bits 16..31 from h_16..h_31  = plane word p,
//              t_15 = sign word
// Host-callable so that tests/test_bitslice.py can check it on the CPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DWT_HD __host__ __device__ __forceinline__
#else
#define DWT_HD inline
#endif

constexpr int BITSLICE_MAX_PLANES = 15; // magnitude bits a packed half-word can carry

DWT_HD uint32_t bitslice_half(int c) // sign-magnitude half-word of one coefficient
{
	const uint32_t mag = (uint32_t)(c < 0 ? -c : c);
	return (mag & 0x7fffu) | (c < 0 ? 0x8000u : 0u);
}

DWT_HD int bitslice_value(uint32_t h) // inverse of bitslice_half
{
	const int mag = (int)(h & 0x7fffu);
	return (h & 0x8000u) ? -mag : mag;
}

// in place: w[i] (i = 0..15) packed half-words in, plane words out (and back: the transpose is an involution)
DWT_HD void bitslice_transpose16(uint32_t (&w)[16])
{
#define DWT_BS_STAGE(J, M)                                                                                              \
	_Pragma("unroll") for (int k = 0; k < 16; ++k)                                                                      \
	{                                                                                                                   \
		if (!(k & J)) {                                                                                                 \
			const uint32_t t = ((w[k] >> J) ^ w[k + J]) & M;                                                            \
			w[k + J] ^= t;                                                                                              \
			w[k] ^= t << J;                                                                                             \
		}                                                                                                               \
	}
	DWT_BS_STAGE(8, 0x00ff00ffu)
	DWT_BS_STAGE(4, 0x0f0f0f0fu)
	DWT_BS_STAGE(2, 0x33333333u)
	DWT_BS_STAGE(1, 0x55555555u)
#undef DWT_BS_STAGE
}
