// hilbert.cu -- Hilbert-order gather / scatter between the Mallat pyramid and the bit-sliced store.
//
// Reference behaviour: encode.c:32-58 (linearization), encode.c:112-131 (sign-magnitude), decode.c:32-65
// (reconstruction incl. dequantisation bias), decode.c:102-117, hilbert.h:15-34 (d -> x,y).
//
// The reference walks all side^2 curve indices serially.  Here the curve is cut into aligned cells of
// 32x32 positions (one CTA each).  The number of valid positions before a cell is a closed-form
// rectangle count that is scanned once per geometry; inside a cell the rank is a
// ballot/popc prefix.  A cell is a compact 2-D patch, so the pyramid reads/writes stay sector-local, and
// the output is written as 32-coefficient groups, bit-sliced: one 32-bit word per bit-plane plus a sign
// word, which is what the coder kernels consume.
#include "hilbert.cuh"
#include "bitslice.cuh"

#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <vector>

namespace {

__host__ __device__ __forceinline__ void hilbert_d2xy(int n, u32 d, int &x, int &y) // hilbert.h:15-34
{
	x = 0;
	y = 0;
	for (int s = 1; s < n; s <<= 1, d >>= 2) {
		int rx = (d >> 1) & 1;
		int ry = (d ^ rx) & 1;
		if (!ry) {
			if (rx) {
				x = s - 1 - x;
				y = s - 1 - y;
			}
			int t = x;
			x = y;
			y = t;
		}
		x += s * rx;
		y += s * ry;
	}
}

inline int overlap(int o, int cs, int lim) // |[o, o+cs) n [0, lim)|
{
	int hi = o + cs < lim ? o + cs : lim;
	return hi > o ? hi - o : 0;
}

struct ChanLayout {
	int planes[3];
	long long bsbase[4];
};

struct HLevel {
	int n, cs, w1, h1, w0, h0, gbase;
	u32 cell_off;
};

struct HParams {
	HLevel lv[DWT_MAX_LEVELS];
	int channels, GT;
	ChanLayout lay;
	const u32 *cell_base, *cell_info;
};

constexpr int PMAX = 12; // bit planes (+ sign) the warp-per-cell kernels keep in registers

__device__ __forceinline__ void build_lut(unsigned short *lut)
{
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
		int x, y;
		hilbert_d2xy(32, (u32)i, x, y);
		lut[i] = (unsigned short)(x | (y << 8));
	}
	__syncthreads();
}

// position of curve index dloc of a cell with orientation o (bit 0 transpose, bit 1 point reflection)
__device__ __forceinline__ void cell_xy(const unsigned short *lut, int dloc, u32 o, int &x, int &y)
{
	const u32 e = lut[dloc];
	int lx = (int)(e & 255u), ly = (int)(e >> 8);
	if (o & 1u) {
		const int t = lx;
		lx = ly;
		ly = t;
	}
	if (o & 2u) {
		lx = 31 - lx;
		ly = 31 - ly;
	}
	x = lx;
	y = ly;
}

// ---- cells whose 1024 positions are all valid: one warp per cell.  The warp walks the cell's groups of 32 ranks;
// lane i of iteration k holds the coefficient of rank 32 (g0 + k) + i of all channels, a ballot per bit-plane gives the
// group's word, lane 0 parks it in the warp's shared-memory tile [plane][group], and at the end every plane row
// leaves as one contiguous 128-byte run.
constexpr int LIN_ROWS = 3 * (PMAX + 1); // plane rows of the three channels

// NP = largest plane count of the channels: the ballots of planes 0 .. NP-1 are unrolled with immediate masks and
// offsets for every channel (a channel with fewer planes parks zero words that are never flushed); signs use row PMAX
template <int NP>
__global__ void __launch_bounds__(256) linearize_full_kernel(const __grid_constant__ HParams P, const u32 *__restrict__ list,
                                                              int nlist, const int *__restrict__ pyr, long long chan_stride,
                                                              int pitch, u32 *bs)
{
	__shared__ unsigned short lut[1024];
	__shared__ u32 tile[8][LIN_ROWS][32];
	build_lut(lut);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	u32(*my)[32] = tile[wid];
	const int C = P.channels;
	for (int item = blockIdx.x * 8 + wid; item < nlist; item += gridDim.x * 8) {
		const u32 ent = list[item];
		const HLevel &L = P.lv[ent >> 28];
		const u32 q = ent & 0x0fffffffu;
		const u32 R0 = P.cell_base[L.cell_off + q], info = P.cell_info[L.cell_off + q];
		const int ox = (int)(info & 0xfffu) * 32, oy = (int)((info >> 12) & 0xfffu) * 32;
		const u32 orient = info >> 24;
		const u32 g0 = R0 >> 5;
		const int s = (int)(R0 & 31u);
		const int *base = pyr + (size_t)oy * pitch + ox;
		const int niter = s ? 33 : 32; // the 33rd group holds the cell's last s positions
		__syncwarp();
		for (int k = 0; k < niter; ++k) {
			const int dloc = 32 * k + lane - s;
			u32 v[3] = {0u, 0u, 0u};
			if (dloc >= 0 && dloc < 1024) {
				int x, y;
				cell_xy(lut, dloc, orient, x, y);
				const int *src = base + (size_t)y * pitch + x;
#pragma unroll
				for (int c = 0; c < 3; ++c) {
					if (c < C) {
						const int t = __ldg(src + (size_t)c * chan_stride);
						v[c] = (t < 0 ? 0x80000000u : 0u) | (u32)abs(t); // encode.c:124-128
					}
				}
			}
			const int col = k & 31; // group 32 reuses column 0 after the first 32 columns are flushed
			if (k == 32) {
				// flush the 32 whole-or-first groups before the tail group overwrites column 0
				__syncwarp();
				for (int c = 0; c < C; ++c) {
					const int planes = P.lay.planes[c];
					u32 *dst = bs + P.lay.bsbase[c] + L.gbase + g0 + lane;
					for (int p = 0; p <= planes; ++p) {
						const u32 w = my[c * (PMAX + 1) + (p == planes ? PMAX : p)][lane];
						u32 *d = dst + (long long)p * P.GT;
						if (lane > 0)
							*d = w; // whole group
						else if (w)
							atomicOr(d, w); // first group of a cell that does not start on a group boundary
					}
				}
				__syncwarp();
			}
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				if (c < C) {
					u32(*rows)[32] = my + c * (PMAX + 1);
#pragma unroll
					for (int p = 0; p < NP; ++p) {
						const u32 w = __ballot_sync(0xffffffffu, v[c] & (1u << p));
						if (lane == 0)
							rows[p][col] = w;
					}
					const u32 w = __ballot_sync(0xffffffffu, v[c] >> 31);
					if (lane == 0)
						rows[PMAX][col] = w; // sign row
				}
			}
		}
		__syncwarp();
		if (s == 0) {
			for (int c = 0; c < C; ++c) {
				const int planes = P.lay.planes[c];
				u32 *dst = bs + P.lay.bsbase[c] + L.gbase + g0 + lane;
				for (int p = 0; p <= planes; ++p)
					dst[(long long)p * P.GT] = my[c * (PMAX + 1) + (p == planes ? PMAX : p)][lane];
			}
		} else if (lane == 0) {
			// tail group (column 0 now): shared with the next cell on the curve
			for (int c = 0; c < C; ++c) {
				const int planes = P.lay.planes[c];
				u32 *dst = bs + P.lay.bsbase[c] + L.gbase + g0 + 32;
				for (int p = 0; p <= planes; ++p) {
					const u32 w = my[c * (PMAX + 1) + (p == planes ? PMAX : p)][0];
					if (w)
						atomicOr(dst + (long long)p * P.GT, w);
				}
			}
		}
	}
}

// inverse: the plane rows of the cell's 33 groups are staged in the warp's shared-memory tile (128-byte loads), then
// lane i of iteration k assembles the coefficient of rank 32 (g0 + k) + i: per plane one broadcast read, one rotate
// that brings bit i to bit p, one merge
template <int NP> // NP = largest plane count of the channels (rows a channel does not have are staged as zero)
__global__ void __launch_bounds__(256) reconstruct_full_kernel(const __grid_constant__ HParams P,
                                                                const u32 *__restrict__ list, int nlist,
                                                                const u32 *__restrict__ bs, const int *__restrict__ missing,
                                                                int levels_used, int *pyr, long long chan_stride, int pitch)
{
	__shared__ unsigned short lut[1024];
	__shared__ u32 tile[8][LIN_ROWS][33];
	build_lut(lut);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	u32(*my)[33] = tile[wid];
	const int C = P.channels;
	int rot[NP]; // rotate right by rot[p]: bit `lane` lands on bit p
#pragma unroll
	for (int p = 0; p < NP; ++p)
		rot[p] = (lane - p) & 31;
	for (int item = blockIdx.x * 8 + wid; item < nlist; item += gridDim.x * 8) {
		const u32 ent = list[item];
		const int level = (int)(ent >> 28);
		if (level >= levels_used)
			continue;
		const HLevel &L = P.lv[level];
		const u32 q = ent & 0x0fffffffu;
		const u32 R0 = P.cell_base[L.cell_off + q], info = P.cell_info[L.cell_off + q];
		const int ox = (int)(info & 0xfffu) * 32, oy = (int)((info >> 12) & 0xfffu) * 32;
		const u32 orient = info >> 24;
		const u32 g0 = R0 >> 5;
		const int s = (int)(R0 & 31u);
		int *base = pyr + (size_t)oy * pitch + ox;
		int bias[3];
		__syncwarp();
		for (int c = 0; c < C; ++c) {
			const int planes = P.lay.planes[c];
			const u32 *src = bs + P.lay.bsbase[c] + L.gbase + g0;
			for (int p = 0; p <= NP; ++p) { // p == NP stands for the sign row
				const int row = c * (PMAX + 1) + (p == NP ? PMAX : p);
				const bool have = p == NP || p < planes;
				const u32 *r = src + (long long)(p == NP ? planes : p) * P.GT;
				my[row][lane] = have ? __ldg(r + lane) : 0u;
				if (s && lane == 0)
					my[row][32] = have ? __ldg(r + 32) : 0u;
			}
			const int m = missing[c * 16 + level] - 2; // decode.c:51-58
			bias[c] = m >= 0 ? 1 << m : 0;
		}
		__syncwarp();
		const int niter = s ? 33 : 32;
		for (int k = 0; k < niter; ++k) {
			const int dloc = 32 * k + lane - s;
			const bool act = dloc >= 0 && dloc < 1024;
			int x = 0, y = 0;
			if (act)
				cell_xy(lut, dloc, orient, x, y);
			int *dstp = base + (size_t)y * pitch + x;
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				if (c < C) {
					const u32(*rows)[33] = my + c * (PMAX + 1);
					u32 mag = 0;
#pragma unroll
					for (int p = 0; p < NP; ++p)
						mag |= __funnelshift_r(rows[p][k], rows[p][k], rot[p]) & (1u << p);
					const u32 neg = (rows[PMAX][k] >> lane) & 1u;
					int v = neg ? -(int)mag : (int)mag;
					if (v != 0)
						v += v < 0 ? -bias[c] : bias[c];
					if (act)
						dstp[(size_t)c * chan_stride] = v;
				}
			}
		}
	}
}

// ---- cells cut by the image or the LL boundary: one CTA per cell, ranks by ballot / popc
__global__ void __launch_bounds__(256) linearize_kernel(const __grid_constant__ HParams P, const u32 *__restrict__ list,
                                                         const int *pyr, long long chan_stride, int pitch, u32 *bs)
{
	__shared__ u32 vals[3][1024];
	__shared__ int wcnt[4][8];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const u32 ent = list[blockIdx.x];
	const HLevel &L = P.lv[ent >> 28];
	const int q = (int)(ent & 0x0fffffffu);
	const int npos = L.cs * L.cs;
	const u32 R0 = P.cell_base[L.cell_off + q];
	int px[4], py[4], rk[4];
	bool ok[4];
#pragma unroll
	for (int it = 0; it < 4; ++it) {
		int dloc = it * 256 + tid;
		bool act = dloc < npos;
		int x = 0, y = 0;
		if (act)
			hilbert_d2xy(L.n, (u32)q * (u32)npos + (u32)dloc, x, y);
		bool v = act && x < L.w1 && y < L.h1 && (x >= L.w0 || y >= L.h0);
		u32 bal = __ballot_sync(0xffffffffu, v);
		px[it] = x;
		py[it] = y;
		ok[it] = v;
		rk[it] = __popc(bal & lanemask_lt());
		if (lane == 0)
			wcnt[it][wid] = __popc(bal);
	}
	__syncthreads();
	int nv = 0;
#pragma unroll
	for (int it = 0; it < 4; ++it) {
		int base = 0;
		for (int j = 0; j < 32; ++j) {
			int c = wcnt[j >> 3][j & 7];
			if (j < it * 8 + wid)
				base += c;
		}
		rk[it] += base;
	}
	for (int j = 0; j < 32; ++j)
		nv += wcnt[j >> 3][j & 7];
#pragma unroll
	for (int it = 0; it < 4; ++it) {
		if (ok[it]) {
			for (int c = 0; c < P.channels; ++c) {
				int v = __ldg(pyr + (size_t)c * chan_stride + (size_t)py[it] * pitch + px[it]);
				vals[c][rk[it]] = (v < 0 ? 0x80000000u : 0u) | (u32)abs(v); // encode.c:124-128
			}
		}
	}
	__syncthreads();
	if (nv == 0)
		return;
	const u32 g_first = R0 >> 5, g_last = (R0 + (u32)nv - 1) >> 5;
	const int ngr = (int)(g_last - g_first) + 1;
	for (int task = wid; task < P.channels * ngr; task += 8) {
		int c = task / ngr;
		u32 g = g_first + (u32)(task - c * ngr);
		long long r = (long long)g * 32 + lane - (long long)R0;
		u32 v = (r >= 0 && r < nv) ? vals[c][r] : 0u;
		const int planes = P.lay.planes[c];
		u32 mine = 0;
		for (int p = 0; p < planes; ++p) {
			u32 b = __ballot_sync(0xffffffffu, (v >> p) & 1u);
			if (lane == p)
				mine = b;
		}
		u32 sb = __ballot_sync(0xffffffffu, v >> 31);
		if (lane == planes)
			mine = sb;
		bool whole = (u64)g * 32 >= R0 && (u64)g * 32 + 32 <= (u64)R0 + (u64)nv;
		if (lane <= planes) {
			u32 *dst = bs + P.lay.bsbase[c] + (long long)lane * P.GT + L.gbase + g;
			if (whole)
				*dst = mine;
			else if (mine)
				atomicOr(dst, mine);
		}
	}
}

__global__ void __launch_bounds__(256) reconstruct_kernel(const __grid_constant__ HParams P, const u32 *__restrict__ list,
                                                           const u32 *bs, const int *missing, int levels_used, int *pyr,
                                                           long long chan_stride, int pitch)
{
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	__shared__ int wcnt[4][8];
	const u32 ent = list[blockIdx.x];
	const int level = (int)(ent >> 28);
	if (level >= levels_used)
		return;
	const HLevel &L = P.lv[level];
	const int q = (int)(ent & 0x0fffffffu);
	const int npos = L.cs * L.cs;
	const u32 R0 = P.cell_base[L.cell_off + q];
	int px[4], py[4], rk[4];
	bool ok[4];
#pragma unroll
	for (int it = 0; it < 4; ++it) {
		int dloc = it * 256 + tid;
		bool act = dloc < npos;
		int x = 0, y = 0;
		if (act)
			hilbert_d2xy(L.n, (u32)q * (u32)npos + (u32)dloc, x, y);
		bool v = act && x < L.w1 && y < L.h1 && (x >= L.w0 || y >= L.h0);
		u32 bal = __ballot_sync(0xffffffffu, v);
		px[it] = x;
		py[it] = y;
		ok[it] = v;
		rk[it] = __popc(bal & lanemask_lt());
		if (lane == 0)
			wcnt[it][wid] = __popc(bal);
	}
	__syncthreads();
#pragma unroll
	for (int it = 0; it < 4; ++it) {
		int base = 0;
		for (int j = 0; j < it * 8 + wid; ++j)
			base += wcnt[j >> 3][j & 7];
		rk[it] += base;
	}
#pragma unroll
	for (int it = 0; it < 4; ++it) {
		if (!ok[it])
			continue;
		u32 r = R0 + (u32)rk[it];
		u32 g = r >> 5, bit = r & 31;
		for (int c = 0; c < P.channels; ++c) {
			const int planes = P.lay.planes[c];
			const u32 *src = bs + P.lay.bsbase[c] + L.gbase + g;
			int mag = 0;
			for (int p = 0; p < planes; ++p)
				mag |= (int)((__ldg(src + (long long)p * P.GT) >> bit) & 1u) << p;
			int neg = (int)((__ldg(src + (long long)planes * P.GT) >> bit) & 1u);
			int v = neg ? -mag : mag;
			int m = missing[c * 16 + level] - 2; // decode.c:51-58
			if (m >= 0 && v != 0)
				v += v < 0 ? -(1 << m) : (1 << m);
			pyr[(size_t)c * chan_stride + (size_t)py[it] * pitch + px[it]] = v;
		}
	}
}

// ------------------------------------------------------------------------------------------------ TMA-staged full cells
//
// One channel of a full cell is a 32 x 32 x 1 box of the pyramid: exactly what the tensor memory accelerator moves.  One
// cp.async.bulk.tensor per (cell, channel) brings the box into shared memory (128-byte swizzle, so that the curve-ordered
// reads below spread over the banks), an mbarrier tells the warp when it has landed, and the next box is already in
// flight in the warp's second buffer.  Lane k then owns the cell's k-th group of 32 curve-consecutive coefficients and
// turns them into bit-plane words with a 16 x 16 bit-matrix transpose in registers (bitslice.cuh) instead of one warp
// ballot per plane and group: ~6 instead of ~30 warp instructions per group and channel.  Every plane row of the cell
// leaves as one 128-byte run.  The inverse (reconstruct) mirrors it and sends each box back with a TMA store.
// Requirements: rows 16-byte aligned in HBM (width % 4 == 0), at most 15 bit-planes; everything else takes the
// ballot kernels above.

constexpr int TMA_WARPS = 8;

__device__ __forceinline__ u32 smem_addr(const void *p)
{
	return (u32)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(u64 *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
	asm volatile("{\n"
	             ".reg .pred p;\n"
	             "WAIT_%=:\n"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	             "@p bra DONE_%=;\n"
	             "bra WAIT_%=;\n"
	             "DONE_%=:\n"
	             "}" ::"r"(smem_addr(bar)),
	             "r"(parity)
	             : "memory");
}

__device__ __forceinline__ void tma_load_cell(void *dst, const CUtensorMap *map, int x, int y, int z, u64 *bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
	                 smem_addr(dst)),
	             "l"(reinterpret_cast<u64>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_addr(bar))
	             : "memory");
}

__device__ __forceinline__ void tma_store_cell(const CUtensorMap *map, int x, int y, int z, const void *src)
{
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(reinterpret_cast<u64>(map)),
	             "r"(x), "r"(y), "r"(z), "r"(smem_addr(src))
	             : "memory");
}

struct TmaSmem { // carved out of the dynamic shared memory, box buffers first (1024-byte aligned for the swizzle)
	int *cells;            // [TMA_WARPS][2][1024]
	unsigned short *lut;   // [4][1024]
	u64 *bars;             // [TMA_WARPS][2]
};

__device__ __forceinline__ TmaSmem tma_smem(unsigned char *raw, const unsigned short *__restrict__ tile_lut)
{
	TmaSmem s;
	const u32 base = smem_addr(raw);
	raw += (1024u - (base & 1023u)) & 1023u;
	s.cells = reinterpret_cast<int *>(raw);
	s.lut = reinterpret_cast<unsigned short *>(raw + (size_t)TMA_WARPS * 2 * 4096);
	s.bars = reinterpret_cast<u64 *>(s.lut + 4096);
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) // 8 KB of offsets, two per word
		reinterpret_cast<u32 *>(s.lut)[i] = __ldg(reinterpret_cast<const u32 *>(tile_lut) + i);
	if (threadIdx.x < TMA_WARPS * 2)
		mbar_init(s.bars + threadIdx.x, 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();
	return s;
}

constexpr size_t tma_smem_bytes()
{
	return (size_t)TMA_WARPS * 2 * 4096 + 8192 + TMA_WARPS * 2 * 8 + 1024;
}

// The unit of work is ONE channel of a cell (a 32 x 32 x 1 box, 4 KB): two boxes per warp are in flight or in use, 8
// warps per CTA and 3 CTAs per SM keep 24 warps on an SM (whole three-channel cells of 12 KB allowed 8: the kernel then
// ran at two warps per scheduler and 38 % issue utilisation).  The cell descriptors (list entry -> rank / origin words:
// two dependent loads) are fetched one and two cells ahead.
constexpr int LIN_CTAS_PER_SM = 3;

template <int NC>
__global__ void __launch_bounds__(TMA_WARPS * 32, LIN_CTAS_PER_SM) linearize_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                                         const __grid_constant__ HParams P,
                                                                                         const u32 *__restrict__ list, int nlist,
                                                                                         const unsigned short *__restrict__ tile_lut,
                                                                                         u32 *bs)
{
	extern __shared__ unsigned char tma_raw[];
	const TmaSmem sm = tma_smem(tma_raw, tile_lut);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int *mycells = sm.cells + (size_t)wid * 2 * 1024;
	u64 *mybar = sm.bars + wid * 2;
	const int stride = gridDim.x * TMA_WARPS;
	int item = blockIdx.x * TMA_WARPS + wid;
	if (item >= nlist)
		return;
	auto entry_of = [&](int it) { return it < nlist ? __ldg(list + it) : 0u; };
	auto words_of = [&](u32 ent, u32 &R0, u32 &info) {
		const HLevel &L = P.lv[ent >> 28];
		R0 = __ldg(P.cell_base + L.cell_off + (ent & 0x0fffffffu));
		info = __ldg(P.cell_info + L.cell_off + (ent & 0x0fffffffu));
	};
	auto issue = [&](u32 info, int ch, int buf) { // lane 0: channel ch of the cell with origin word `info` into buffer `buf`
		mbar_expect_tx(mybar + buf, 4096u);
		tma_load_cell(mycells + (size_t)buf * 1024, &tmap, (int)(info & 0xfffu) * 32, (int)((info >> 12) & 0xfffu) * 32, ch,
		              mybar + buf);
	};
	// descriptors: the current cell, the next one, and the list entry of the one after
	u32 ent_c = entry_of(item), R0_c, info_c;
	words_of(ent_c, R0_c, info_c);
	u32 ent_n = entry_of(item + stride), R0_n = 0, info_n = 0;
	if (item + stride < nlist)
		words_of(ent_n, R0_n, info_n);
	u32 ent_nn = entry_of(item + 2 * stride);
	if (lane == 0)
		issue(info_c, 0, 0);
	u32 phases = 0;
	int buf = 0;
	for (; item < nlist; item += stride) {
		const u32 g0 = (u32)P.lv[ent_c >> 28].gbase + (R0_c >> 5);
		const int s = (int)(R0_c & 31u);
		const unsigned short *lo = sm.lut + (info_c >> 24) * 1024;
		const bool split = s != 0 && lane == 0;
		const u32 lo_mask = (1u << s) - 1u; // ranks below s of lane 0's word belong to the group behind the cell's last whole one
#pragma unroll 1
		for (int ch = 0; ch < NC; ++ch, buf ^= 1) {
			// the next unit's box goes into the other buffer, which the __syncwarp at the end of the last round released
			if (lane == 0) {
				if (ch + 1 < NC)
					issue(info_c, ch + 1, buf ^ 1);
				else if (item + stride < nlist)
					issue(info_n, 0, buf ^ 1);
			}
			mbar_wait(mybar + buf, (phases >> buf) & 1u);
			phases ^= 1u << buf;
			const int *cell = mycells + (size_t)buf * 1024;
			// lane k owns the ranks 32 k - s .. 32 k - s + 31 of the cell (mod 1024: for s > 0 lane 0 holds the head of the cell's
			// first group in its high bits and the tail of the cell's last group in its low bits)
			u32 w[16];
#pragma unroll
			for (int i = 0; i < 32; ++i) {
				const int ip = (i - s) & 31;
				const int kp = (lane - (i < s ? 1 : 0)) & 31;
				const u32 h = bitslice_half(cell[lo[ip * 32 + kp]]);
				if (i < 16)
					w[i] = h;
				else
					w[i - 16] |= h << 16;
			}
			bitslice_transpose16(w);
			const int planes = P.lay.planes[ch];
			u32 *dst = bs + P.lay.bsbase[ch] + g0 + lane;
#pragma unroll
			for (int p = 0; p < 16; ++p) {
				if (p < 15 && p >= planes)
					continue;
				const u32 v = w[p];
				u32 *d = dst + (long long)(p == 15 ? planes : p) * P.GT;
				if (!split) {
					*d = v;
				} else {
					if (v & ~lo_mask)
						atomicOr(d, v & ~lo_mask);
					if (v & lo_mask)
						atomicOr(d + 32, v & lo_mask);
				}
			}
			__syncwarp(); // every lane is done with the buffer before lane 0 lets the next box land in it
		}
		ent_c = ent_n;
		R0_c = R0_n;
		info_c = info_n;
		ent_n = ent_nn;
		if (item + 2 * stride < nlist)
			words_of(ent_n, R0_n, info_n);
		ent_nn = entry_of(item + 3 * stride);
	}
}

// The inverse has no box to wait for -- its input are the plane rows, fetched with plain loads -- so what hides the load
// latency is the number of warps.  A warp therefore works on ONE channel of a cell at a time: a 4 KB cell buffer per warp
// instead of 12 KB and 16 plane words in registers instead of 48, which lets 32 warps share an SM (the three-channel
// version fitted 16 and spent 77 % of its stall samples waiting for the plane rows).  The TMA store of channel c reads
// the buffer while the warp loads and transposes the planes of channel c + 1; the cell descriptors (list entry ->
// rank / origin words, two dependent loads) are fetched one and two cells ahead.
constexpr int REC_WARPS = 8;
constexpr int REC_CTAS_PER_SM = 4;

constexpr size_t rec_smem_bytes()
{
	return (size_t)REC_WARPS * 4096 + 8192 + 1024;
}

template <int NC>
__global__ void __launch_bounds__(REC_WARPS * 32, REC_CTAS_PER_SM) reconstruct_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                                           const __grid_constant__ HParams P,
                                                                                           const u32 *__restrict__ list, int nlist,
                                                                                           const unsigned short *__restrict__ tile_lut,
                                                                                           const u32 *__restrict__ bs,
                                                                                           const int *__restrict__ missing,
                                                                                           int levels_used)
{
	extern __shared__ unsigned char tma_raw[];
	unsigned char *raw = tma_raw;
	raw += (1024u - (smem_addr(raw) & 1023u)) & 1023u;
	int *cells = reinterpret_cast<int *>(raw);
	unsigned short *lut = reinterpret_cast<unsigned short *>(raw + (size_t)REC_WARPS * 4096);
	for (int i = threadIdx.x; i < 2048; i += blockDim.x)
		reinterpret_cast<u32 *>(lut)[i] = __ldg(reinterpret_cast<const u32 *>(tile_lut) + i);
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int *cell = cells + (size_t)wid * 1024;
	const int stride = gridDim.x * REC_WARPS;
	int item = blockIdx.x * REC_WARPS + wid;
	// descriptor pipeline: `ent` belongs to the cell after the next, (R0, info) to the next cell
	auto entry_of = [&](int it) { return it < nlist ? __ldg(list + it) : 0u; };
	u32 ent0 = entry_of(item), ent1 = entry_of(item + stride);
	u32 R0 = 0, info = 0;
	if (item < nlist) {
		const HLevel &L = P.lv[ent0 >> 28];
		R0 = __ldg(P.cell_base + L.cell_off + (ent0 & 0x0fffffffu));
		info = __ldg(P.cell_info + L.cell_off + (ent0 & 0x0fffffffu));
	}
	for (; item < nlist; item += stride) {
		const int level = (int)(ent0 >> 28);
		const int ox = (int)(info & 0xfffu) * 32, oy = (int)((info >> 12) & 0xfffu) * 32;
		const u32 orient = info >> 24;
		const u32 g0 = (u32)P.lv[level].gbase + (R0 >> 5);
		const int s = (int)(R0 & 31u);
		// next cell's descriptor words, and the list entry behind it
		ent0 = ent1;
		ent1 = entry_of(item + 2 * stride);
		if (item + stride < nlist) {
			const HLevel &L = P.lv[ent0 >> 28];
			R0 = __ldg(P.cell_base + L.cell_off + (ent0 & 0x0fffffffu));
			info = __ldg(P.cell_info + L.cell_off + (ent0 & 0x0fffffffu));
		}
		if (level >= levels_used)
			continue;
		const unsigned short *lo = lut + orient * 1024;
		const bool split = s != 0 && lane == 0;
		const u32 lo_mask = (1u << s) - 1u;
#pragma unroll 1
		for (int ch = 0; ch < NC; ++ch) {
			u32 w[16];
			const int planes = P.lay.planes[ch];
			const u32 *src = bs + P.lay.bsbase[ch] + g0 + lane;
			// every plane row is requested before the first one is used; the second word of a cell that starts inside a
			// group (s != 0: lane 0 owns the head of the first group and the tail of the last) is a pass of its own behind
			// a warp-uniform branch -- merged into the loop above, each plane's load waited for the one in front
#pragma unroll
			for (int p = 0; p < 16; ++p)
				w[p] = (p == 15 || p < planes) ? __ldg(src + (long long)(p == 15 ? planes : p) * P.GT) : 0u;
			if (s != 0) {
				u32 t[16];
#pragma unroll
				for (int p = 0; p < 16; ++p)
					t[p] = (split && (p == 15 || p < planes)) ? __ldg(src + (long long)(p == 15 ? planes : p) * P.GT + 32) : 0u;
#pragma unroll
				for (int p = 0; p < 16; ++p)
					if (split)
						w[p] = (w[p] & ~lo_mask) | (t[p] & lo_mask);
			}
			const int m = missing[ch * 16 + level] - 2; // decode.c:51-58
			const int bias = m >= 0 ? 1 << m : 0;
			bitslice_transpose16(w);
			// the TMA store of the previous channel must have finished reading the buffer
			if (lane == 0)
				asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
			__syncwarp();
#pragma unroll
			for (int i = 0; i < 32; ++i) {
				const int ip = (i - s) & 31;
				const int kp = (lane - (i < s ? 1 : 0)) & 31;
				const int off = lo[ip * 32 + kp];
				int v = bitslice_value(i < 16 ? (w[i] & 0xffffu) : (w[i - 16] >> 16));
				if (v != 0)
					v += v < 0 ? -bias : bias;
				cell[off] = v;
			}
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // my writes before the bulk copy reads them
			__syncwarp();
			if (lane == 0) {
				tma_store_cell(&tmap, ox, oy, ch, cell);
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			}
		}
	}
	if (lane == 0)
		asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the stores that read it
	__syncwarp();
}

// 3-D tensor map of a planar int32 pyramid [C][H][W] with boxes of one cell: 32 x 32 x C, 128-byte swizzle
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
	static const EncodeTiledFn fn = [] {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
			cudaGetLastError();
			p = nullptr;
		}
		return reinterpret_cast<EncodeTiledFn>(p);
	}();
	return fn;
}

bool make_cell_map(CUtensorMap *map, const int *pyr, int W, int H, int C, long long chan_stride, int pitch, int box_c)
{
	const EncodeTiledFn fn = encode_tiled_fn();
	if (!fn || (pitch & 3) || (chan_stride & 3) || (reinterpret_cast<uintptr_t>(pyr) & 15) || W < 32 || H < 32)
		return false;
	const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C};
	const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)chan_stride * 4};
	const cuuint32_t box[3] = {32, 32, (cuuint32_t)box_c}, estr[3] = {1, 1, 1};
	return fn(map, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<int *>(pyr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tma_wanted()
{
	const char *v = getenv("DWT_HILBERT"); // A/B and test aid: "ballot" sends full cells through the ballot kernels
	return !(v && !strcmp(v, "ballot"));
}

bool planes_fit_tma(const Geom &g, const Sched &s)
{
	for (int c = 0; c < g.channels; ++c)
		if (s.planes[c] > BITSLICE_MAX_PLANES)
			return false;
	return g.channels == 1 || g.channels == 3;
}

int tma_grid(int nlist)
{
	const int want = (nlist + TMA_WARPS - 1) / TMA_WARPS;
	const int cap = dwt_device_sms() * LIN_CTAS_PER_SM; // CTAs of ~73 KB
	return want < cap ? want : cap;
}

int rec_grid(int nlist)
{
	const int want = (nlist + REC_WARPS - 1) / REC_WARPS;
	const int cap = dwt_device_sms() * REC_CTAS_PER_SM;
	return want < cap ? want : cap;
}

template <typename K>
int tma_configure(K kernel, size_t bytes)
{
	CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
	return 0;
}

__global__ void unslice_kernel(Geom G, ChanLayout lay, const u32 *bs, int *planar, long long total)
{
	long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; // index inside the detail range
	int c = blockIdx.y;
	long long ndet = G.pix[G.levels] - G.pix[0];
	if (k >= ndet)
		return;
	int l = 0;
	while (l + 1 < G.levels && k + G.pix[0] >= G.pix[l + 1])
		++l;
	long long r = k + G.pix[0] - G.pix[l];
	long long g = G.gbase[l] + (r >> 5);
	int bit = (int)(r & 31);
	const u32 *src = bs + lay.bsbase[c] + g;
	int planes = lay.planes[c], mag = 0;
	for (int p = 0; p < planes; ++p)
		mag |= (int)((src[(long long)p * G.GT] >> bit) & 1u) << p;
	int neg = (int)((src[(long long)planes * G.GT] >> bit) & 1u);
	planar[(long long)c * total + G.pix[0] + k] = neg ? -mag : mag;
}

ChanLayout make_layout(const Sched &s, int channels)
{
	ChanLayout lay;
	for (int c = 0; c < 3; ++c)
		lay.planes[c] = c < channels ? s.planes[c] : 0;
	for (int c = 0; c < 4; ++c)
		lay.bsbase[c] = s.bsbase[c];
	return lay;
}

HParams make_params(const Geom &g, const HilbertPlan &plan, const Sched &s)
{
	HParams P;
	for (int l = 0; l < DWT_MAX_LEVELS; ++l) {
		HLevel &L = P.lv[l];
		if (l < g.levels) {
			L.n = g.len[l + 1];
			L.cs = plan.cs[l];
			L.w1 = g.w[l + 1];
			L.h1 = g.h[l + 1];
			L.w0 = g.w[l];
			L.h0 = g.h[l];
			L.gbase = g.gbase[l];
			L.cell_off = (u32)plan.cell_off[l];
		} else {
			L.n = L.cs = L.w1 = L.h1 = L.w0 = L.h0 = L.gbase = 0;
			L.cell_off = 0;
		}
	}
	P.channels = g.channels;
	P.GT = g.GT;
	P.lay = make_layout(s, g.channels);
	P.cell_base = plan.cell_base;
	P.cell_info = plan.cell_info;
	return P;
}

bool planes_fit(const Geom &g, const Sched &s)
{
	for (int c = 0; c < g.channels; ++c)
		if (s.planes[c] > PMAX)
			return false;
	return true;
}

} // namespace

int hilbert_plan_build(const Geom &g, HilbertPlan *plan, cudaStream_t st, long long *launches)
{
	(void)launches;
	int off = 0;
	for (int l = 0; l < g.levels; ++l) {
		int n = g.len[l + 1];
		int cs = n < 32 ? n : 32;
		plan->cs[l] = cs;
		plan->ncell[l] = (n / cs) * (n / cs);
		plan->cell_off[l] = off;
		off += plan->ncell[l];
	}
	plan->cell_off[g.levels] = off;
	const size_t ntot = (size_t)(off > 0 ? off : 1);
	std::vector<u32> h_base(ntot), h_info(ntot), h_full, h_part;
	for (int l = 0; l < g.levels; ++l) {
		const int n = g.len[l + 1], cs = plan->cs[l], nc = plan->ncell[l];
		const int w1 = g.w[l + 1], h1 = g.h[l + 1], w0 = g.w[l], h0 = g.h[l];
		plan->full_off[l] = (int)h_full.size();
		plan->part_off[l] = (int)h_part.size();
		u32 run = 0;
		for (int q = 0; q < nc; ++q) {
			int x0, y0, x1, y1;
			hilbert_d2xy(n, (u32)q * (u32)(cs * cs), x0, y0);
			hilbert_d2xy(n, (u32)q * (u32)(cs * cs) + 1u, x1, y1);
			const int ox = x0 & ~(cs - 1), oy = y0 & ~(cs - 1);
			// the base curve starts at (0,0) and steps to (0,1): a start in the far corner means the cell is
			// point-reflected, a first step along x means it is transposed
			const u32 flip = (x0 - ox) != 0 ? 2u : 0u;
			const u32 swap = (y1 == y0) ? 1u : 0u;
			const int cnt = overlap(ox, cs, w1) * overlap(oy, cs, h1) - overlap(ox, cs, w0) * overlap(oy, cs, h0);
			const size_t idx = (size_t)plan->cell_off[l] + q;
			h_base[idx] = run;
			h_info[idx] = (u32)(ox / cs) | ((u32)(oy / cs) << 12) | ((swap | flip) << 24);
			run += (u32)cnt;
			const u32 ent = ((u32)l << 28) | (u32)q;
			if (cs == 32 && cnt == 1024)
				h_full.push_back(ent);
			else if (cnt > 0)
				h_part.push_back(ent);
		}
	}
	plan->full_off[g.levels] = (int)h_full.size();
	plan->part_off[g.levels] = (int)h_part.size();
	// word offsets of the curve inside a TMA-staged cell, per orientation: entry [o][i][k] belongs to curve index 32 k + i;
	// the box arrives with the 128-byte swizzle (16-byte chunk index XOR row mod 8)
	std::vector<unsigned short> h_lut(4096);
	for (u32 o = 0; o < 4; ++o)
		for (int d = 0; d < 1024; ++d) {
			int x, y;
			hilbert_d2xy(32, (u32)d, x, y);
			if (o & 1u) {
				const int t = x;
				x = y;
				y = t;
			}
			if (o & 2u) {
				x = 31 - x;
				y = 31 - y;
			}
			const int off = y * 32 + ((((x >> 2) ^ (y & 7)) << 2) | (x & 3));
			h_lut[o * 1024 + (d & 31) * 32 + (d >> 5)] = (unsigned short)off;
		}
	const size_t words = 2 * ntot + h_full.size() + h_part.size() + 4 + 2048;
	plan->cell_base = nullptr;
	CUDA_OK(cudaMalloc(&plan->cell_base, sizeof(u32) * words));
	plan->cell_info = plan->cell_base + ntot;
	plan->full_list = plan->cell_info + ntot;
	plan->part_list = plan->full_list + h_full.size();
	plan->tile_lut = reinterpret_cast<unsigned short *>(plan->part_list + h_part.size() + ((h_full.size() + h_part.size()) & 1) + 2);
	CUDA_OK(cudaMemcpyAsync(plan->tile_lut, h_lut.data(), sizeof(unsigned short) * 4096, cudaMemcpyHostToDevice, st));
	CUDA_OK(cudaMemcpyAsync(plan->cell_base, h_base.data(), sizeof(u32) * ntot, cudaMemcpyHostToDevice, st));
	CUDA_OK(cudaMemcpyAsync(plan->cell_info, h_info.data(), sizeof(u32) * ntot, cudaMemcpyHostToDevice, st));
	if (!h_full.empty())
		CUDA_OK(cudaMemcpyAsync(plan->full_list, h_full.data(), sizeof(u32) * h_full.size(), cudaMemcpyHostToDevice, st));
	if (!h_part.empty())
		CUDA_OK(cudaMemcpyAsync(plan->part_list, h_part.data(), sizeof(u32) * h_part.size(), cudaMemcpyHostToDevice, st));
	CUDA_OK(cudaStreamSynchronize(st)); // the host vectors go out of scope
	return 0;
}

void hilbert_plan_free(HilbertPlan *plan)
{
	if (plan->cell_base)
		cudaFree(plan->cell_base);
	plan->cell_base = nullptr;
}

static int full_grid(int nlist)
{
	const int sms = dwt_device_sms();
	const int want = (nlist + 7) / 8;
	return want < sms * 8 ? want : sms * 8;
}

int hilbert_linearize(const Geom &g, const HilbertPlan &plan, const Sched &s, const int *pyr,
                      long long pyr_chan_stride, int pyr_pitch, u32 *bs, int levels_used, cudaStream_t st,
                      long long *launches)
{
	const HParams P = make_params(g, plan, s);
	const bool fast = planes_fit(g, s);
	// lists are ordered by level: the first levels_used levels are a prefix
	const int nfull = plan.full_off[levels_used], npart = plan.part_off[levels_used];
	CUtensorMap tmap;
	if (nfull > 0 && tma_wanted() && planes_fit_tma(g, s) &&
	    make_cell_map(&tmap, pyr, g.w[g.levels], g.h[g.levels], g.channels, pyr_chan_stride, pyr_pitch, 1)) {
		// full cells through the tensor memory accelerator (see the TMA section above)
		if (g.channels == 3) {
			if (tma_configure(linearize_tma_kernel<3>, tma_smem_bytes()))
				return -1;
			linearize_tma_kernel<3><<<tma_grid(nfull), TMA_WARPS * 32, tma_smem_bytes(), st>>>(tmap, P, plan.full_list, nfull,
			                                                                                plan.tile_lut, bs);
		} else {
			if (tma_configure(linearize_tma_kernel<1>, tma_smem_bytes()))
				return -1;
			linearize_tma_kernel<1><<<tma_grid(nfull), TMA_WARPS * 32, tma_smem_bytes(), st>>>(tmap, P, plan.full_list, nfull,
			                                                                                plan.tile_lut, bs);
		}
		if (launches)
			++*launches;
	} else if (fast && nfull > 0) {
		int np = 1;
		for (int c = 0; c < g.channels; ++c)
			if (s.planes[c] > np)
				np = s.planes[c];
		const int grid = full_grid(nfull);
#define LIN_CASE(N)                                                                                                      \
	case N:                                                                                                              \
		linearize_full_kernel<N><<<grid, 256, 0, st>>>(P, plan.full_list, nfull, pyr, pyr_chan_stride, pyr_pitch, bs);    \
		break;
		switch (np) {
			LIN_CASE(1) LIN_CASE(2) LIN_CASE(3) LIN_CASE(4) LIN_CASE(5) LIN_CASE(6) LIN_CASE(7) LIN_CASE(8) LIN_CASE(9)
			LIN_CASE(10) LIN_CASE(11) LIN_CASE(12)
		}
#undef LIN_CASE
		if (launches)
			++*launches;
	} else if (nfull > 0) {
		linearize_kernel<<<nfull, 256, 0, st>>>(P, plan.full_list, pyr, pyr_chan_stride, pyr_pitch, bs);
		if (launches)
			++*launches;
	}
	if (npart > 0) {
		linearize_kernel<<<npart, 256, 0, st>>>(P, plan.part_list, pyr, pyr_chan_stride, pyr_pitch, bs);
		if (launches)
			++*launches;
	}
	CUDA_OK(cudaGetLastError());
	return 0;
}

int hilbert_reconstruct(const Geom &g, const HilbertPlan &plan, const Sched &s, const u32 *bs,
                        const int *missing_dev, int *pyr, long long pyr_chan_stride, int pyr_pitch,
                        int levels_used, cudaStream_t st, long long *launches)
{
	const HParams P = make_params(g, plan, s);
	const bool fast = planes_fit(g, s);
	const int nfull = plan.full_off[levels_used], npart = plan.part_off[levels_used];
	CUtensorMap tmap;
	if (nfull > 0 && tma_wanted() && planes_fit_tma(g, s) &&
	    make_cell_map(&tmap, pyr, g.w[levels_used], g.h[levels_used], g.channels, pyr_chan_stride, pyr_pitch, 1)) {
		if (g.channels == 3) {
			if (tma_configure(reconstruct_tma_kernel<3>, rec_smem_bytes()))
				return -1;
			reconstruct_tma_kernel<3><<<rec_grid(nfull), REC_WARPS * 32, rec_smem_bytes(), st>>>(tmap, P, plan.full_list, nfull,
			                                                                                  plan.tile_lut, bs, missing_dev,
			                                                                                  levels_used);
		} else {
			if (tma_configure(reconstruct_tma_kernel<1>, rec_smem_bytes()))
				return -1;
			reconstruct_tma_kernel<1><<<rec_grid(nfull), REC_WARPS * 32, rec_smem_bytes(), st>>>(tmap, P, plan.full_list, nfull,
			                                                                                  plan.tile_lut, bs, missing_dev,
			                                                                                  levels_used);
		}
		if (launches)
			++*launches;
	} else if (fast && nfull > 0) {
		int np = 1;
		for (int c = 0; c < g.channels; ++c)
			if (s.planes[c] > np)
				np = s.planes[c];
		const int grid = full_grid(nfull);
#define REC_CASE(N)                                                                                                      \
	case N:                                                                                                              \
		reconstruct_full_kernel<N><<<grid, 256, 0, st>>>(P, plan.full_list, nfull, bs, missing_dev, levels_used, pyr,     \
		                                                 pyr_chan_stride, pyr_pitch);                                    \
		break;
		switch (np) {
			REC_CASE(1) REC_CASE(2) REC_CASE(3) REC_CASE(4) REC_CASE(5) REC_CASE(6) REC_CASE(7) REC_CASE(8) REC_CASE(9)
			REC_CASE(10) REC_CASE(11) REC_CASE(12)
		}
#undef REC_CASE
		if (launches)
			++*launches;
	} else if (nfull > 0) {
		reconstruct_kernel<<<nfull, 256, 0, st>>>(P, plan.full_list, bs, missing_dev, levels_used, pyr, pyr_chan_stride,
		                                          pyr_pitch);
		if (launches)
			++*launches;
	}
	if (npart > 0) {
		reconstruct_kernel<<<npart, 256, 0, st>>>(P, plan.part_list, bs, missing_dev, levels_used, pyr, pyr_chan_stride,
		                                          pyr_pitch);
		if (launches)
			++*launches;
	}
	CUDA_OK(cudaGetLastError());
	return 0;
}

int hilbert_unslice(const Geom &g, const Sched &s, const u32 *bs, int *planar, long long total, cudaStream_t st)
{
	long long ndet = g.pix[g.levels] - g.pix[0];
	if (ndet <= 0)
		return 0;
	dim3 grid((unsigned)((ndet + 255) / 256), g.channels);
	unslice_kernel<<<grid, 256, 0, st>>>(g, make_layout(s, g.channels), bs, planar, total);
	CUDA_OK(cudaGetLastError());
	return 0;
}
