// coder_enc.cu -- the bit-plane coder of encode.c:60-95,183-221 + rle.h + vli.h + bits.h as a parallel pipeline.
//
// The reference emits one bit at a time through put_rle -> put_vli -> put_bit -> put_byte.  The stream it
// produces is: for every chunk (channel, level, plane) in schedule order, a
// significance pass (zero runs, adaptive-Rice coded, each closed by a 1 and followed by a raw sign bit)
// and a refinement pass (raw bits, preceded by a "phantom one" that closes a pending run).  Runs and the
// Rice order cross chunk boundaries.  Parallel formulation:
//
//   count    per (chunk, tile of 256 groups): #zero symbols, #ones, #refinement bits, from popc on the
//            bit-sliced store ("already significant" <=> a higher plane has a 1)            [enc_count_kernel]
//   scan     exclusive prefixes in schedule order; per-chunk token index / run bookkeeping   [enc_scan_kernel,
//            (flush candidates = phantom ones, final rle_flush token)                         enc_chunk_setup_kernel]
//   emit     every 1 becomes a token: Z[t] = number of zero symbols before it, so its run is Z[t]-Z[t-1];
//            sign bits and the dense refinement bit vector are packed with a software pext    [enc_emit_kernel]
//   orders   the Rice order is a serial recurrence k' = max(ilog2(v + 2^k) - 2, 0).  The maps are monotone
//            in k, so a tile whose end order agrees for entry orders 0 and 31 is a constant map; tiles are
//            solved by a block-local fixed-point iteration, the (rare) non-constant tiles are chained
//            exactly                                                              [enc_vli_kernel]
//   offsets  exclusive scan of token bit lengths (decoupled look-back inside enc_vli_kernel)
//   scatter  tokens and refinement bits are OR-ed into the zero-initialised stream          [enc_scatter_kernel,
//                                                                                            enc_refcopy_kernel]
#include "coder.cuh"

#include <stdlib.h>

namespace {

constexpr int TG = DWT_TILE_GROUPS;
constexpr int TT = DWT_TOK_TILE;
constexpr int TPT = DWT_TOK_PER_THREAD;

__device__ __forceinline__ void tile_coords(const Geom &G, int b, int &c, int &l, int &i)
{
	int tpc = G.tbase[G.levels];
	c = b / tpc;
	int r = b - c * tpc;
	l = G.levels - 1; // from the top: three quarters of the tiles belong to the finest level
	while (l > 0 && G.tbase[l] > r)
		--l;
	i = r - G.tbase[l];
}

__device__ __forceinline__ u32 group_valid_mask(const Geom &G, int l, int g)
{
	if (g >= G.G[l])
		return 0u;
	long long rem = G.num[l] - (long long)g * 32;
	return rem >= 32 ? 0xffffffffu : ((1u << (int)rem) - 1u);
}

// ------------------------------------------------------------------------------------------------ count

__global__ void __launch_bounds__(TG) enc_count_kernel(const __grid_constant__ Geom G, const Sched *__restrict__ S,
                                                        const u32 *__restrict__ bs, u32 *ent_z, u32 *ent_1, u32 *ent_r)
{
	__shared__ u32 acc[DWT_MAX_PLANES][3];
	int c, l, i;
	tile_coords(G, blockIdx.x, c, l, i);
	const int P = S->planes[c];
	for (int k = threadIdx.x; k < P * 3; k += TG)
		(&acc[0][0])[k] = 0;
	__syncthreads();
	const int g = i * TG + threadIdx.x;
	const u32 vm = group_valid_mask(G, l, g);
	const u32 *base = bs + S->bsbase[c] + G.gbase[l] + g;
	u32 sig = 0;
	u32 Bn = vm && P > 0 ? __ldg(base + (long long)(P - 1) * G.GT) : 0u; // the next plane's word is always one plane ahead
	for (int p = P - 1; p >= 0; --p) {
		const u32 B = Bn;
		Bn = vm && p > 0 ? __ldg(base + (long long)(p - 1) * G.GT) : 0u;
		u32 member = vm & ~sig;
		u32 n1 = __popc(B & member), nz = __popc(member) - n1, nr = __popc(sig);
		nz = __reduce_add_sync(0xffffffffu, nz);
		n1 = __reduce_add_sync(0xffffffffu, n1);
		nr = __reduce_add_sync(0xffffffffu, nr);
		if ((threadIdx.x & 31) == 0) {
			if (nz)
				atomicAdd(&acc[p][0], nz);
			if (n1)
				atomicAdd(&acc[p][1], n1);
			if (nr)
				atomicAdd(&acc[p][2], nr);
		}
		sig |= B;
	}
	__syncthreads();
	if (threadIdx.x < P) {
		int j = S->chunk_of[c][l][threadIdx.x];
		int e = S->ebase[j] + i;
		ent_z[e] = acc[threadIdx.x][0];
		ent_1[e] = acc[threadIdx.x][1];
		ent_r[e] = acc[threadIdx.x][2];
	}
}

// ------------------------------------------------------------------------------------------------ scans

// exclusive scan of the three count arrays in two passes over blocks of SCAN_TILE entries: block-local prefixes and
// block totals first, then every block adds the totals of the blocks in front of it (a few dozen values)
constexpr int SCAN_TILE = 4096;

__global__ void __launch_bounds__(1024) enc_scan_local_kernel(u32 *ez, u32 *e1, u32 *er, int n, u64 *bsum)
{
	__shared__ u64 ws[32];
	u32 *arr[3] = {ez, e1, er};
	const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
	for (int a = 0; a < 3; ++a) {
		u32 v[4];
#pragma unroll
		for (int k = 0; k < 4; ++k)
			v[k] = base + k < n ? arr[a][base + k] : 0u;
		u64 tot;
		u32 run = (u32)block_exscan_u64((u64)v[0] + v[1] + v[2] + v[3], ws, &tot);
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			if (base + k < n)
				arr[a][base + k] = run; // prefixes are used modulo 2^32 (runs are differences)
			run += v[k];
		}
		if (threadIdx.x == 0)
			bsum[a * gridDim.x + blockIdx.x] = tot;
	}
}

__global__ void __launch_bounds__(1024) enc_scan_fix_kernel(u32 *ez, u32 *e1, u32 *er, int n, const u64 *bsum, EncInfo *info)
{
	__shared__ u64 ws[32];
	u32 *arr[3] = {ez, e1, er};
	const int nb = gridDim.x;
	const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
	u64 tots[3];
	for (int a = 0; a < 3; ++a) {
		u64 before = 0, all = 0;
		for (int i = threadIdx.x; i < nb; i += 1024) {
			const u64 v = bsum[a * nb + i];
			all += v;
			if (i < (int)blockIdx.x)
				before += v;
		}
		u64 t0, t1;
		block_exscan_u64(before, ws, &t0);
		block_exscan_u64(all, ws, &t1);
		tots[a] = t1;
		const u32 add = (u32)t0;
#pragma unroll
		for (int k = 0; k < 4; ++k)
			if (base + k < n)
				arr[a][base + k] += add;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		info->tot_zero = tots[0];
		info->tot_one = tots[1];
		info->tot_ref = tots[2];
		info->error = 0;
		info->opaque_tiles = 0;
		if (tots[1] >= 0xfffff000ull || tots[2] >= 0xfffff000ull)
			info->error = 2; // more than 2^32 tokens / refinement bits: outside the supported range
	}
}

// slack behind the capacity before a chunk counts as unreachable: keeps "the stream is longer than the capacity" certain
// for the host's stderr counters (a cut stream reports 8 * capacity bits, encode.c:226-230)
constexpr u64 CUT_SLACK_BITS = 64;

__global__ void __launch_bounds__(1024) enc_chunk_setup_kernel(const Sched *__restrict__ S, const u32 *ez, const u32 *e1,
                                                                const u32 *er, int nent, EncChunks *C, EncInfo *info, u32 *Z,
                                                                u32 *specbuf, u32 max_tokens, u64 prefix_bits, u64 limit_bits)
{
	__shared__ u64 ws[32];
	__shared__ int s_jcut, s_icut, s_ecut;
	const int J = S->nchunks;
	const int per = (J + 1 + 1023) / 1024;
	const int b = threadIdx.x * per, e = min(b + per, J);
	const u32 tz = (u32)info->tot_zero, t1 = (u32)info->tot_one, tr = (u32)info->tot_ref;
	if (threadIdx.x == 0) {
		// capacity cut: a one costs at least 2 bits (VLI >= 1 bit, sign), a refinement bit 1, and the refinement block of a
		// chunk lies behind all of its tokens -- so the first token of tile i of chunk j cannot start in front of
		//     prefix + 2 * (ones in front of it) + (refinement bits of the chunks in front of j).
		// The first (chunk, tile) entry whose bound lies behind the capacity, and everything after it, cannot reach the
		// output (bytes.h:77-78).  The bound is non-decreasing along the entries: binary search, chunks first, then tiles.
		const u64 L = limit_bits + CUT_SLACK_BITS;
		int jc = J, ic = 0, lo = nent;
		if (limit_bits) {
			int a = 0, z = J; // first chunk whose first entry is bounded behind L (J: none)
			while (a < z) {
				const int mid = (a + z) >> 1;
				const int em = S->ebase[mid];
				if (prefix_bits + 2ull * e1[em] + er[em] >= L)
					z = mid;
				else
					a = mid + 1;
			}
			const int jhi = a;
			jc = jhi;
			ic = 0;
			lo = jhi < J ? S->ebase[jhi] : nent;
			if (jhi > 0) { // the cut may fall inside the last chunk that starts in front of L
				const int jp = jhi - 1, e0 = S->ebase[jp], en = S->ebase[jp + 1];
				const u64 base = prefix_bits + er[e0];
				int x = e0, y = en; // first entry of chunk jp bounded behind L (en: none)
				while (x < y) {
					const int mid = (x + y) >> 1;
					if (base + 2ull * e1[mid] >= L)
						y = mid;
					else
						x = mid + 1;
				}
				if (x < en) {
					jc = jp;
					ic = x - e0;
					lo = x;
				}
			}
		}
		s_jcut = jc;
		s_icut = ic;
		s_ecut = lo;
	}
	u64 nf = 0;
	for (int j = b; j < e; ++j) {
		u32 r0 = er[S->ebase[j]], r1 = j + 1 < J ? er[S->ebase[j + 1]] : tr;
		nf += (r1 != r0);
	}
	u64 tot;
	u64 F = block_exscan_u64(nf, ws, &tot); // syncs: the cut is known behind it
	const int jcut = s_jcut;
	for (int j = b; j < e; ++j) {
		const int e0 = S->ebase[j];
		u32 r0 = er[e0], r1 = j + 1 < J ? er[S->ebase[j + 1]] : tr;
		u32 o0 = e1[e0], o1 = j + 1 < J ? e1[S->ebase[j + 1]] : t1;
		u32 z1 = j + 1 < J ? ez[S->ebase[j + 1]] : tz;
		C->tok_start[j] = o0 + (u32)F;
		C->tok_adj[j] = (u32)F;
		C->ref_start[j] = r0;
		C->nref[j] = r1 - r0;
		C->ref_pos[j] = 0;
		if (r1 != r0) { // flush candidate right after the chunk's last 1 (rle.h:79-89)
			u32 t = o1 + (u32)F;
			if (t < max_tokens && j < jcut) {
				Z[t + 1] = z1;
				atomicOr(specbuf + (t >> 5), 1u << (t & 31));
			}
			++F;
		}
	}
	__syncthreads(); // tok_adj / ref_start of chunk jcut (written by its owner) are read by thread 0
	if (threadIdx.x == 0) {
		u32 tf = t1 + (u32)tot; // final rle_flush token (rle.h:37-40)
		u32 zf = tz;
		C->tok_start[J] = tf;
		C->tok_adj[J] = (u32)tot;
		C->ref_start[J] = tr;
		info->jcut = jcut;
		info->icut = s_icut;
		info->ref_cut = tr;
		if (jcut < J) {
			// the token list ends in front of tile icut of chunk jcut; its last entry stands for "whatever follows": it is
			// placed behind the capacity, so its bits never reach the output (enc_scatter clips at the capacity).  The
			// refinement block of chunk jcut lies behind all of its tokens, hence behind the capacity as well.
			tf = e1[s_ecut] + C->tok_adj[jcut];
			zf = ez[s_ecut];
			info->ref_cut = C->ref_start[jcut];
		}
		Z[0] = 0;
		if (tf < max_tokens) {
			Z[tf + 1] = zf;
			atomicOr(specbuf + (tf >> 5), 1u << (tf & 31));
		} else {
			info->error = 3;
		}
		info->ntok = tf + 1;
		info->ntiles = (tf + 1 + TT - 1) / TT;
	}
}

// ------------------------------------------------------------------------------------------------ emit

__global__ void __launch_bounds__(TG) enc_emit_kernel(const __grid_constant__ Geom G, const Sched *__restrict__ S,
                                                       const u32 *__restrict__ bs, const u32 *__restrict__ ez,
                                                       const u32 *__restrict__ e1, const u32 *__restrict__ er,
                                                       const EncChunks *__restrict__ C, const EncInfo *__restrict__ info, u32 *Z,
                                                       u32 *signbuf, u32 *refbuf)
{
	__shared__ __align__(16) u32 ws[2][8];
	int c, l, i;
	tile_coords(G, blockIdx.x, c, l, i);
	const int P = S->planes[c];
	// chunks behind jcut, and the tiles from icut on of chunk jcut, lie behind the capacity (enc_chunk_setup_kernel)
	const int jcut = info->jcut, icut = info->icut;
	if (P == 0)
		return;
	{
		const int jtop = S->chunk_of[c][l][P - 1];
		if (jtop > jcut || (jtop == jcut && i >= icut))
			return; // the planes of a (channel, level) are coded top down: nothing of this tile is needed
	}
	const int g = i * TG + threadIdx.x;
	const u32 vm = group_valid_mask(G, l, g);
	const u32 *base = bs + S->bsbase[c] + G.gbase[l] + g;
	const u32 sgn = vm ? __ldg(base + (long long)P * G.GT) : 0u;
	u32 sig = 0;
	u32 Bn = vm ? __ldg(base + (long long)(P - 1) * G.GT) : 0u; // the next plane's word is always one plane ahead
	for (int p = P - 1; p >= 0; --p) {
		const u32 B = Bn;
		Bn = vm && p > 0 ? __ldg(base + (long long)(p - 1) * G.GT) : 0u;
		u32 member = vm & ~sig;
		u32 ones = B & member, zeros = member & ~B;
		u32 n1 = __popc(ones), nz = __popc(zeros), nr = __popc(sig);
		const int j = S->chunk_of[c][l][p];
		if (j > jcut || (j == jcut && i >= icut))
			break; // block-uniform: this plane and the ones below it are cut
		// zeros and ones before this group inside the tile; the refinement bits follow from them: the groups in front of a
		// valid group are all full, so zeros + ones + refinement bits in front of it = 32 per group
		const u32 ex = tile_exscan_2x16(nz | (n1 << 16), ws, p & 1);
		const u32 ex_z = ex & 0xffffu, ex_1 = ex >> 16, ex_r = 32u * threadIdx.x - ex_z - ex_1;
		const int e = S->ebase[j] + i;
		if (n1) {
			u32 zb = ez[e] + ex_z;
			u32 tb = e1[e] + C->tok_adj[j] + ex_1;
			// the signs of the ones, in token order: gathered in the walk over the ones (cheaper than a parallel-suffix
			// compress of the sign word for the two or three ones a group has per plane)
			u32 o = ones, t = tb, sb = 0;
			while (o) {
				int b = __ffs(o) - 1;
				Z[t + 1] = zb + __popc(zeros & ((1u << b) - 1u));
				sb |= ((sgn >> b) & 1u) << (t - tb);
				++t;
				o &= o - 1;
			}
			bits_or(signbuf, tb, sb, (int)n1);
		}
		if (nr && j < jcut) { // the refinement block of chunk jcut lies behind all of its tokens: behind the cut
			u32 rb = er[e] + ex_r;
			bits_or(refbuf, rb, bit_compress(B, sig), (int)nr);
		}
		sig |= B;
	}
}

// ------------------------------------------------------------------------------------------------ VLI orders

struct Tok {
	u32 v[TPT];
	u32 kinds; // 2 bits per token: 0 = one (+sign), 1 = flush candidate, 2 = final, 3 = inactive
	u32 signs;
};

__device__ __forceinline__ void load_tokens(const u32 *__restrict__ Z, const u32 *__restrict__ specbuf,
                                            const u32 *__restrict__ signbuf, u32 ntok, u32 t0, Tok &T)
{
	T.kinds = 0;
	T.signs = 0;
	if (t0 >= ntok) {
		T.kinds = 0xffffu;
#pragma unroll
		for (int k = 0; k < TPT; ++k)
			T.v[k] = 0;
		return;
	}
	const uint4 a = *reinterpret_cast<const uint4 *>(Z + t0), b = *reinterpret_cast<const uint4 *>(Z + t0 + 4);
	const u32 last = Z[t0 + 8];
	u32 z[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, last};
	const u32 spec = (specbuf[t0 >> 5] >> (t0 & 31)) & 0xffu;
	if (signbuf)
		T.signs = (signbuf[t0 >> 5] >> (t0 & 31)) & 0xffu;
	if (spec == 0 && t0 + TPT < ntok) { // eight ordinary tokens (a one and its sign each): what all but ~300 threads of a frame hold
#pragma unroll
		for (int k = 0; k < TPT; ++k)
			T.v[k] = z[k + 1] - z[k];
		return;
	}
#pragma unroll
	for (int k = 0; k < TPT; ++k) {
		u32 t = t0 + k;
		T.v[k] = z[k + 1] - z[k];
		u32 kind = t >= ntok ? 3u : (t == ntok - 1 ? 2u : ((spec >> k) & 1u));
		if (kind == 3u)
			T.v[k] = 0;
		T.kinds |= kind << (2 * k);
	}
}

// run the tokens of one thread from order k; returns the order after them and adds their bit lengths
__device__ __forceinline__ int run_tokens(const Tok &T, int k, u32 *bits, u32 *longest = nullptr)
{
	u32 sum = 0;
	if (T.kinds == 0) { // ordinary tokens only: no kind to look at
		u32 mx = 0;
#pragma unroll
		for (int i = 0; i < TPT; ++i) {
			const int e = vli_e(T.v[i], k);
			const u32 len = (u32)(2 * e - k + 2);
			sum += len;
			mx = max(mx, len);
			k = vli_next(e);
		}
		if (bits)
			*bits = sum;
		if (longest)
			*longest = mx;
		return k;
	}
	if (longest)
		*longest = 64;
#pragma unroll
	for (int i = 0; i < TPT; ++i) {
		u32 kind = (T.kinds >> (2 * i)) & 3u;
		u32 v = T.v[i];
		if (kind == 3u || (kind == 1u && v == 0))
			continue; // an empty phantom run emits nothing and leaves the order alone (rle.h:83)
		int e = vli_e(v, k);
		sum += 2 * e - k + 1 + (kind == 0u);
		k = vli_next(e);
	}
	if (bits)
		*bits = sum;
	return k;
}

// block-local fixed point: thread t starts from the order thread t-1 ended with; `first` is the tile's entry order
__device__ __forceinline__ int solve_tile(const Tok &T, int first, unsigned char *ends, int *start_out)
{
	int s = threadIdx.x == 0 ? first : 0, e = 0;
	bool dirty = true;
	for (;;) {
		if (dirty)
			e = run_tokens(T, s, nullptr);
		ends[threadIdx.x] = (unsigned char)e;
		__syncthreads();
		int ns = threadIdx.x == 0 ? first : ends[threadIdx.x - 1];
		dirty = ns != s;
		s = ns;
		if (!__syncthreads_or(dirty))
			break;
	}
	*start_out = s;
	return e;
}

// One pass over the token tiles (2048 tokens each), tiles taken in launch order through a ticket so that a tile only
// ever waits for tiles that are already running:
//   1. end order of the tile for entry orders 0 and 31 (the order map is monotone: equal ends = constant map), published;
//   2. the tile's entry order: the predecessor's end if its map is constant, else the predecessor's resolved end
//      (a chain only through the rare non-constant tiles);
//   3. exact orders of the tile's threads, published resolved end, bit lengths and the tile's bit total.
// (A decoupled look-back for the bit prefix inside this kernel was measured and lost: every tile then holds its SM slot
// until its 32 predecessors have their totals; the prefix is a separate two-pass scan.)
#define VA_READY 0x80000000u

__global__ void __launch_bounds__(256) enc_vli_kernel(const u32 *__restrict__ Z, const u32 *__restrict__ specbuf, EncInfo *info,
                                                       u32 *ticket, volatile u32 *tileA, volatile u32 *tileB,
                                                       unsigned char *thr_state, u32 *tile_bits, int k0)
{
	__shared__ unsigned char ends[256];
	__shared__ u32 s_tile, total;
	__shared__ int s_start;
	const int tid = threadIdx.x;
	if (tid == 0) {
		s_tile = atomicAdd(ticket, 1u);
		total = 0;
	}
	__syncthreads();
	const u32 tile = s_tile;
	const u32 ntok = info->ntok;
	if ((u64)tile * TT >= ntok)
		return;
	Tok T;
	load_tokens(Z, specbuf, nullptr, ntok, tile * TT + tid * TPT, T);
	bool big = false;
#pragma unroll
	for (int k = 0; k < TPT; ++k)
		big |= T.v[k] >= (1u << 30);
	if (big)
		info->error = 4; // run length beyond the reference's int range
	int s_lo, s_hi;
	const int lo = solve_tile(T, 0, ends, &s_lo);
	__syncthreads();
	const int hi = solve_tile(T, 31, ends, &s_hi);
	if (tid == 255)
		tileA[tile] = VA_READY | (u32)lo | ((u32)hi << 8);
	if (tid == 0) {
		int start = k0;
		if (tile > 0) {
			u32 a;
			while (!((a = tileA[tile - 1]) & VA_READY))
				;
			const int plo = (int)(a & 0xffu), phi = (int)((a >> 8) & 0xffu);
			if (plo == phi) {
				start = plo; // constant map: the predecessor's entry order does not matter
			} else {
				u32 bprev;
				while (!((bprev = tileB[tile - 1]) & VA_READY))
					;
				start = (int)(bprev & 0xffu);
				atomicAdd(&info->opaque_tiles, 1u);
			}
		}
		s_start = start;
	}
	__syncthreads();
	const int start = s_start;
	int s, end;
	if (start == 0) {
		s = s_lo;
		end = lo;
	} else if (start == 31) {
		s = s_hi;
		end = hi;
	} else {
		__syncthreads();
		end = solve_tile(T, start, ends, &s);
	}
	if (tid == 255)
		tileB[tile] = VA_READY | (u32)end;
	thr_state[(size_t)tile * 256 + tid] = (unsigned char)s;
	u32 bits;
	run_tokens(T, s, &bits);
	bits = __reduce_add_sync(0xffffffffu, bits);
	if ((tid & 31) == 0)
		atomicAdd(&total, bits);
	__syncthreads();
	if (tid == 0)
		tile_bits[tile] = total;
}

// exclusive prefix (64-bit) of the tile bit totals, two passes over blocks of SCAN_TILE tiles
__global__ void __launch_bounds__(1024) enc_bitscan_local_kernel(const u32 *__restrict__ tile_bits, u64 *tile_bitbase,
                                                                 const EncInfo *__restrict__ info, u64 *bsum)
{
	__shared__ u64 ws[32];
	const int n = (int)((info->ntok + TT - 1) / TT);
	const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
	u32 v[4];
#pragma unroll
	for (int k = 0; k < 4; ++k)
		v[k] = base + k < n ? tile_bits[base + k] : 0u;
	u64 tot;
	u64 run = block_exscan_u64((u64)v[0] + v[1] + v[2] + v[3], ws, &tot);
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		if (base + k < n)
			tile_bitbase[base + k] = run;
		run += v[k];
	}
	if (threadIdx.x == 0)
		bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) enc_bitscan_fix_kernel(u64 *tile_bitbase, EncInfo *info, const u64 *__restrict__ bsum)
{
	__shared__ u64 ws[32];
	const int n = (int)((info->ntok + TT - 1) / TT);
	const int nb = gridDim.x;
	const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
	u64 before = 0, all = 0;
	for (int i = threadIdx.x; i < nb; i += 1024) {
		const u64 v = bsum[i];
		all += v;
		if (i < (int)blockIdx.x)
			before += v;
	}
	u64 t0, t1;
	block_exscan_u64(before, ws, &t0);
	block_exscan_u64(all, ws, &t1);
#pragma unroll
	for (int k = 0; k < 4; ++k)
		if (base + k < n)
			tile_bitbase[base + k] += t0;
	if (blockIdx.x == 0 && threadIdx.x == 0)
		info->tok_bits = t1;
}

// ------------------------------------------------------------------------------------------------ scatter

__device__ __forceinline__ int chunk_of_index(const u32 *starts, int J, u32 t) // max j in [0,J] with starts[j] <= t
{
	int lo = 0, hi = J;
	while (lo < hi) {
		int mid = (lo + hi + 1) >> 1;
		if (starts[mid] <= t)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

__global__ void __launch_bounds__(256) enc_scatter_kernel(const u32 *__restrict__ Z, const u32 *__restrict__ specbuf,
                                                           const u32 *__restrict__ signbuf, EncInfo *info,
                                                           const unsigned char *__restrict__ thr_state,
                                                           const u64 *__restrict__ tile_bitbase, EncChunks *C, int J,
                                                           u32 *out, u64 prefix_bits, u64 limit_bits)
{
	__shared__ u64 ws[32];
	const u32 ntok = info->ntok;
	const u32 tile = blockIdx.x;
	if ((u64)tile * TT >= ntok)
		return;
	const u32 t0 = tile * TT + threadIdx.x * TPT;
	Tok T;
	load_tokens(Z, specbuf, signbuf, ntok, t0, T);
	int k = thr_state[(size_t)tile * 256 + threadIdx.x];
	u32 mybits, longest;
	run_tokens(T, k, &mybits, &longest);
	u64 tot;
	u64 off = prefix_bits + tile_bitbase[tile] + block_exscan_u64(mybits, ws, &tot);
	// chunk of the tile's first token: one binary search per block, the threads only walk forward from it
	__shared__ int s_j0;
	if (threadIdx.x == 0)
		s_j0 = chunk_of_index(C->tok_start, J, tile * TT);
	__syncthreads();
	if (t0 >= ntok)
		return;
	int j = s_j0;
	u32 next_start = j < J ? C->tok_start[j + 1] : 0xffffffffu; // first token of the next chunk
	while (j < J && t0 >= next_start) {
		++j;
		next_start = j < J ? C->tok_start[j + 1] : 0xffffffffu;
	}
	u64 ref_start = C->ref_start[j];
	if (T.kinds == 0 && longest <= 24 && (j >= J || t0 + TPT <= next_start)) {
		// the common case: eight ordinary tokens of one chunk, no code longer than 24 bits.  Their codes are contiguous in
		// the stream: a 64-bit shift register, one atomic per 32-bit word that fills up
		const u64 pos = off + ref_start;
		u64 w = pos >> 5;
		int fill = (int)(pos & 31);
		u64 acc = 0;
#pragma unroll
		for (int i = 0; i < TPT; ++i) {
			const u32 x = T.v[i] + (1u << k); // vli.h:67-84 in closed form: e - k zeros, a one, the e low bits of x, then the sign
			const int e = ilog2_u32(x);
			const int nz = e - k;
			const u32 code = ((((x ^ (1u << e)) << 1) | 1u) << nz) | (((T.signs >> i) & 1u) << (nz + 1 + e));
			acc |= (u64)code << fill;
			fill += nz + e + 2;
			k = vli_next(e);
			if (fill >= 32) {
				if ((w << 5) < limit_bits && (u32)acc)
					atomicOr(out + w, (u32)acc);
				acc >>= 32;
				fill -= 32;
				++w;
			}
		}
		if (fill > 0 && (w << 5) < limit_bits && (u32)acc)
			atomicOr(out + w, (u32)acc);
		return;
	}
	// the codes of a thread's tokens are contiguous in the stream (except across a chunk's refinement block): they are
	// gathered in a 64-bit window and sent with one atomic per 32-bit word instead of one or two per token
	u64 acc = 0, acc_pos = 0;
	int acc_len = 0;
	auto flush = [&]() {
		if (acc_len > 0 && acc_pos < limit_bits) {
			bits_or(out, acc_pos, (u32)acc, acc_len < 32 ? acc_len : 32);
			if (acc_len > 32)
				bits_or(out, acc_pos + 32, (u32)(acc >> 32), acc_len - 32);
		}
		acc = 0;
		acc_len = 0;
	};
#pragma unroll
	for (int i = 0; i < TPT; ++i) {
		const u32 t = t0 + i;
		const u32 kind = (T.kinds >> (2 * i)) & 3u;
		if (kind == 3u)
			break;
		while (j < J && t >= next_start) {
			++j;
			next_start = j < J ? C->tok_start[j + 1] : 0xffffffffu;
			ref_start = C->ref_start[j];
		}
		const u32 v = T.v[i];
		int len = 0;
		if (!(kind == 1u && v == 0)) {
			const int e = vli_e(v, k);
			const int nz = e - k;
			const u32 payload = v - ((1u << e) - (1u << k));
			u64 code = (1ull << nz) | ((u64)payload << (nz + 1));
			len = nz + 1 + e;
			if (kind == 0u) {
				code |= (u64)((T.signs >> i) & 1u) << len;
				++len;
			}
			const u64 pos = off + ref_start;
			if (acc_len == 0 || pos != acc_pos + (u64)acc_len || acc_len + len > 64) {
				flush();
				acc_pos = pos;
			}
			acc |= code << acc_len;
			acc_len += len;
			k = vli_next(e);
		}
		off += len;
		if (kind == 1u)
			C->ref_pos[j] = off + ref_start; // the chunk's refinement block starts right after the phantom one
		if (kind == 2u)
			info->total_bits = off + ref_start;
	}
	flush();
}

__global__ void __launch_bounds__(256) enc_refcopy_kernel(const u32 *__restrict__ refbuf, const EncChunks *__restrict__ C,
                                                           int J, u64 tot_ref, u32 *out, u64 limit_bits)
{
	const u64 sw = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	u64 pos = sw * 32;
	if (pos >= tot_ref)
		return;
	const u64 end = min(pos + 32, tot_ref);
	const u32 word = refbuf[sw];
	int j = chunk_of_index(C->ref_start, J, (u32)pos);
	while (pos < end) {
		while (j < J && (u64)C->ref_start[j + 1] <= pos)
			++j;
		u64 seg_end = min(end, (u64)C->ref_start[j + 1]);
		int n = (int)(seg_end - pos);
		u32 bits = word >> (int)(pos - sw * 32);
		u64 dst = C->ref_pos[j] + (pos - C->ref_start[j]);
		if (dst < limit_bits)
			bits_or(out, dst, bits, n);
		pos = seg_end;
	}
}

} // namespace

int enc_count(const Geom &g, const Sched &hs, const EncBuffers &b, cudaStream_t st, long long *launches)
{
	(void)hs;
	int blocks = g.channels * g.tbase[g.levels];
	enc_count_kernel<<<blocks, TG, 0, st>>>(g, b.sched, b.bs, b.ent_z, b.ent_1, b.ent_r);
	++*launches;
	CUDA_OK(cudaGetLastError());
	return 0;
}

int enc_scan_and_setup(const Geom &g, const Sched &hs, const EncBuffers &b, u64 prefix_bits, u64 limit_bits, cudaStream_t st,
                       long long *launches)
{
	(void)g;
	(void)hs;
	const int nb = (b.nent + SCAN_TILE - 1) / SCAN_TILE;
	u64 *bsum = reinterpret_cast<u64 *>((reinterpret_cast<uintptr_t>(b.ent_r + b.nent) + 15) & ~(uintptr_t)15); // room kept by the context
	enc_scan_local_kernel<<<nb, 1024, 0, st>>>(b.ent_z, b.ent_1, b.ent_r, b.nent, bsum);
	enc_scan_fix_kernel<<<nb, 1024, 0, st>>>(b.ent_z, b.ent_1, b.ent_r, b.nent, bsum, b.info);
	++*launches;
	enc_chunk_setup_kernel<<<1, 1024, 0, st>>>(b.sched, b.ent_z, b.ent_1, b.ent_r, b.nent, b.chunks, b.info, b.Z, b.specbuf,
	                                           b.max_tokens, prefix_bits, limit_bits);
	*launches += 2;
	CUDA_OK(cudaGetLastError());
	return 0;
}

// Tokens / refinement bits that can survive a capacity of limit_bits (0: no limit): every entry in front of the cut has
// a lower bound below limit + slack, i.e. fewer than (limit + slack) / 2 ones and (limit + slack) refinement bits in front of
// it; the entry itself adds at most a tile of ones, the flush candidates one token per chunk, and the refinement block of
// the last whole chunk at most one level's worth of bits.
u64 enc_token_bound(const Geom &g, const Sched &hs, u64 prefix_bits, u64 limit_bits)
{
	(void)g;
	if (!limit_bits)
		return ~0ull;
	const u64 room = limit_bits + CUT_SLACK_BITS > prefix_bits ? limit_bits + CUT_SLACK_BITS - prefix_bits : 0;
	return room / 2 + (u64)TG * 32 + (u64)hs.nchunks + 2;
}

u64 enc_ref_bound(const Geom &g, const Sched &hs, u64 prefix_bits, u64 limit_bits)
{
	(void)hs;
	if (!limit_bits)
		return ~0ull;
	u64 biggest = 0;
	for (int l = 0; l < g.levels; ++l)
		if ((u64)g.num[l] > biggest)
			biggest = (u64)g.num[l];
	const u64 room = limit_bits + CUT_SLACK_BITS > prefix_bits ? limit_bits + CUT_SLACK_BITS - prefix_bits : 0;
	return room + biggest + 64;
}

int enc_emit(const Geom &g, const Sched &hs, const EncBuffers &b, cudaStream_t st, long long *launches)
{
	(void)hs;
	int blocks = g.channels * g.tbase[g.levels];
	enc_emit_kernel<<<blocks, TG, 0, st>>>(g, b.sched, b.bs, b.ent_z, b.ent_1, b.ent_r, b.chunks, b.info, b.Z, b.signbuf,
	                                       b.refbuf);
	++*launches;
	CUDA_OK(cudaGetLastError());
	return 0;
}

int enc_vli_orders(const EncBuffers &b, int k0, cudaStream_t st, long long *launches)
{
	// the token count lives on the device: launch for the upper bound, surplus tiles exit at once
	const unsigned tiles = (b.max_tokens + TT - 1) / TT;
	CUDA_OK(cudaMemsetAsync(b.tile_flags, 0, b.tile_flag_bytes, st));
	u32 *ticket = b.tile_flags;
	u32 *tileA = ticket + 8, *tileB = tileA + (tiles + 4), *tile_bits = tileB + (tiles + 4);
	u64 *bsum = reinterpret_cast<u64 *>((reinterpret_cast<uintptr_t>(tile_bits + (tiles + 4)) + 15) & ~(uintptr_t)15);
	enc_vli_kernel<<<tiles, 256, 0, st>>>(b.Z, b.specbuf, b.info, ticket, tileA, tileB, b.thr_state, tile_bits, k0);
	const unsigned nb = (tiles + SCAN_TILE - 1) / SCAN_TILE;
	enc_bitscan_local_kernel<<<nb, 1024, 0, st>>>(tile_bits, b.tile_bitbase, b.info, bsum);
	enc_bitscan_fix_kernel<<<nb, 1024, 0, st>>>(b.tile_bitbase, b.info, bsum);
	*launches += 3;
	CUDA_OK(cudaGetLastError());
	return 0;
}

int enc_scatter(const Sched &hs, const EncBuffers &b, u64 prefix_bits, u64 tot_ref, cudaStream_t st, long long *launches)
{
	unsigned tiles = (b.max_tokens + TT - 1) / TT;
	enc_scatter_kernel<<<tiles, 256, 0, st>>>(b.Z, b.specbuf, b.signbuf, b.info, b.thr_state, b.tile_bitbase, b.chunks,
	                                          hs.nchunks, b.out, prefix_bits, b.out_limit_bits);
	++*launches;
	if (tot_ref) {
		u64 words = (tot_ref + 31) / 32;
		enc_refcopy_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(b.refbuf, b.chunks, hs.nchunks, tot_ref, b.out,
		                                                                      b.out_limit_bits);
		++*launches;
	}
	CUDA_OK(cudaGetLastError());
	return 0;
}
