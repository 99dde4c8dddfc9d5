// common.cuh -- shared device helpers and host-side plan structures of libdwt_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DWT_MAX_LEVELS 16
#define DWT_MAX_PLANES 30     // magnitudes stay below 2^29 in the reference (encode.c:116-128)
#define DWT_MAX_CHUNKS (DWT_MAX_LEVELS * 3 * DWT_MAX_PLANES + 1)
#define DWT_TILE_GROUPS 256   // 32-coefficient groups per coder tile (one per thread)
#define DWT_TOK_TILE 2048     // tokens per VLI tile (256 threads x 8)
#define DWT_TOK_PER_THREAD 8

typedef unsigned int u32;
typedef unsigned long long u64;

// Level geometry (utils.h:17-40) plus the bit-sliced coefficient store layout.
// Level index l: 0 = root .. levels = full image.  Detail level l (0 .. levels-1) holds the coefficients
// of the w[l+1] x h[l+1] domain outside the w[l] x h[l] LL rectangle: num[l] = pix[l+1] - pix[l] values
// per channel, stored Hilbert-ordered in G[l] = ceil(num/32) groups of 32.
struct Geom {
	int levels, channels;
	int w[DWT_MAX_LEVELS + 1], h[DWT_MAX_LEVELS + 1], len[DWT_MAX_LEVELS + 1];
	long long pix[DWT_MAX_LEVELS + 1];
	long long num[DWT_MAX_LEVELS];
	int G[DWT_MAX_LEVELS];           // groups per detail level
	int gbase[DWT_MAX_LEVELS + 1];   // first group of level l inside a channel's concatenated group axis
	int GT;                          // groups per channel = gbase[levels]
	int ntile[DWT_MAX_LEVELS];       // coder tiles per level
	int tbase[DWT_MAX_LEVELS + 1];   // first tile of level l inside a channel
};

// Chunk schedule (encode.c:183-221): chunk j = (channel, level, plane), in emission order.
struct Sched {
	int nchunks;
	int planes[3];
	long long bsbase[4];             // word offset of channel c in the bit-sliced store = sum (planes+1)*GT
	short chan[DWT_MAX_CHUNKS], level[DWT_MAX_CHUNKS], plane[DWT_MAX_CHUNKS];
	int ebase[DWT_MAX_CHUNKS + 1];   // first scan entry (tile) of chunk j
	short chunk_of[3][DWT_MAX_LEVELS][DWT_MAX_PLANES]; // (c,l,p) -> j, -1 when absent
};

#define CUDA_OK(x)                                                                                       \
	do {                                                                                                 \
		cudaError_t e_ = (x);                                                                            \
		if (e_ != cudaSuccess) {                                                                         \
			dwt_set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_));               \
			return -1;                                                                                   \
		}                                                                                                \
	} while (0)

void dwt_set_error(const char *fmt, ...);
int dwt_device_sms(); // multiprocessors of the calling thread's current device (cached per device)

#ifdef __CUDACC__

__device__ __forceinline__ u32 lanemask_lt()
{
	u32 m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

// Hacker's Delight 7-4 "compress": gather the bits of x selected by m to the low end (software pext).
__device__ __forceinline__ u32 bit_compress(u32 x, u32 m)
{
	x &= m;
	u32 mk = ~m << 1;
#pragma unroll
	for (int i = 0; i < 5; ++i) {
		u32 mp = mk ^ (mk << 1);
		mp ^= mp << 2;
		mp ^= mp << 4;
		mp ^= mp << 8;
		mp ^= mp << 16;
		u32 mv = mp & m;
		m = (m ^ mv) | (mv >> (1 << i));
		u32 t = x & mv;
		x = (x ^ t) | (t >> (1 << i));
		mk &= ~mp;
	}
	return x;
}

// inverse of bit_compress: spread the low popc(m) bits of x to the positions selected by m (software pdep)
__device__ __forceinline__ u32 bit_expand(u32 x, u32 m)
{
	u32 m0 = m, mk = ~m << 1, a[5];
#pragma unroll
	for (int i = 0; i < 5; ++i) {
		u32 mp = mk ^ (mk << 1);
		mp ^= mp << 2;
		mp ^= mp << 4;
		mp ^= mp << 8;
		mp ^= mp << 16;
		u32 mv = mp & m;
		a[i] = mv;
		m = (m ^ mv) | (mv >> (1 << i));
		mk &= ~mp;
	}
#pragma unroll
	for (int i = 4; i >= 0; --i) {
		u32 mv = a[i];
		u32 t = x << (1 << i);
		x = (x & ~mv) | (t & mv);
	}
	return x & m0;
}

// bit_expand split in two for several words that go through the same mask: the mask-only part (the five move masks) ...
struct ExpandPlan {
	u32 mv[5], m0;
};
__device__ __forceinline__ ExpandPlan bit_expand_plan(u32 m)
{
	ExpandPlan pl;
	pl.m0 = m;
	u32 mk = ~m << 1;
#pragma unroll
	for (int i = 0; i < 5; ++i) {
		u32 mp = mk ^ (mk << 1);
		mp ^= mp << 2;
		mp ^= mp << 4;
		mp ^= mp << 8;
		mp ^= mp << 16;
		const u32 mv = mp & m;
		pl.mv[i] = mv;
		m = (m ^ mv) | (mv >> (1 << i));
		mk &= ~mp;
	}
	return pl;
}
// ... and the five conditional moves of one word
__device__ __forceinline__ u32 bit_expand_apply(u32 x, const ExpandPlan &pl)
{
#pragma unroll
	for (int i = 4; i >= 0; --i) {
		const u32 t = x << (1 << i);
		x = (x & ~pl.mv[i]) | (t & pl.mv[i]);
	}
	return x & pl.m0;
}

// OR `n` (0..32) bits into an LSB-first bit array of zero-initialised 32-bit words at bit offset `off`.
__device__ __forceinline__ void bits_or(u32 *buf, u64 off, u32 bits, int n)
{
	if (n <= 0)
		return;
	if (n < 32)
		bits &= (1u << n) - 1u;
	u64 w = off >> 5;
	int s = (int)(off & 31);
	u32 lo = bits << s;
	if (lo)
		atomicOr(buf + w, lo);
	if (s && s + n > 32) {
		u32 hi = bits >> (32 - s);
		if (hi)
			atomicOr(buf + w + 1, hi);
	}
}

// fetch 32 bits starting at bit offset `off` from an LSB-first bit array (reads words w and w+1)
__device__ __forceinline__ u32 bits_get32(const u32 *buf, u64 off)
{
	u64 w = off >> 5;
	int s = (int)(off & 31);
	u32 lo = buf[w];
	u32 hi = buf[w + 1];
	return __funnelshift_r(lo, hi, s);
}

__device__ __forceinline__ int ilog2_u32(u32 x) // x > 0
{
	return 31 - __clz(x);
}

// adaptive Rice step (vli.h:67-84 in closed form): value v at order k is coded with
// e = ilog2(v + 2^k): (e-k) zeros, a one, e payload bits; next order max(e-2, 0).
__device__ __forceinline__ int vli_e(u32 v, int k)
{
	return ilog2_u32(v + (1u << k));
}
__device__ __forceinline__ int vli_next(int e)
{
	return e >= 2 ? e - 2 : 0;
}

// block-wide exclusive scan of one u64 per thread (three packed 21-bit fields are fine); blockDim <= 1024.
// `warp_sums` must hold 32 u64.  Returns the exclusive prefix; *total receives the block total.
__device__ __forceinline__ u64 block_exscan_u64(u64 v, u64 *warp_sums, u64 *total)
{
	int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	u64 inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		u64 t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d)
			inc += t;
	}
	if (lane == 31)
		warp_sums[wid] = inc;
	__syncthreads();
	u64 base = 0, tot = 0;
	for (int i = 0; i < nw; ++i) {
		u64 s = warp_sums[i];
		if (i < wid)
			base += s;
		tot += s;
	}
	__syncthreads();
	*total = tot;
	return base + inc - v;
}

// exclusive scan over the 256 threads of a tile of two counts packed as lo | hi << 16 (tile totals < 2^16): one warp
// shuffle scan, the eight warp totals through shared memory, ONE block barrier (`ws` holds two sets of warp totals used
// alternately by consecutive calls: a warp can only be one call ahead of the slowest).  The generic 64-bit
// block_exscan_u64 cost a quarter of enc_emit's instructions.
__device__ __forceinline__ u32 tile_exscan_2x16(u32 v, u32 (*ws)[8], int parity)
{
	static_assert(DWT_TILE_GROUPS == 256, "eight warps per tile");
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	u32 inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const u32 t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d)
			inc += t;
	}
	if (lane == 31)
		ws[parity][wid] = inc;
	__syncthreads();
	const uint4 a = *reinterpret_cast<const uint4 *>(&ws[parity][0]), b = *reinterpret_cast<const uint4 *>(&ws[parity][4]);
	u32 base = 0;
	base += wid > 0 ? a.x : 0u;
	base += wid > 1 ? a.y : 0u;
	base += wid > 2 ? a.z : 0u;
	base += wid > 3 ? a.w : 0u;
	base += wid > 4 ? b.x : 0u;
	base += wid > 5 ? b.y : 0u;
	base += wid > 6 ? b.z : 0u;
	return base + inc - v;
}

#endif // __CUDACC__
