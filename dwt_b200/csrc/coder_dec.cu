// coder_dec.cu -- the bit-plane decoder of decode.c:67-100,187-243 + rle.h + vli.h + bits.h on the GPU.
//
// Chunks (channel, level, plane) are decoded in schedule order; within a chunk:
//   prep      per tile of 256 groups: how many coefficients are still insignificant (members of the
//             significance pass) and how many are already significant (refinement bits)   [dec_prep_kernel]
//   tilescan  exclusive prefixes of those counts, chunk totals, parse control reset     [dec_tilescan_kernel]
//   parse     the significance pass is a serial chain of [adaptive-Rice run][sign] tokens.  The stream is cut
//             into windows of 1024 slices of 64 bits; persistent CTAs (one per SM) take windows in order.
//             A CTA first builds, for each of its slices, the exact transfer table "a token starts at offset
//             d with order 0 -> where (and with which order) does the chain leave the slice" by dynamic
//             programming from the slice end.  With the tables the two canonical chains of the window (even
//             / odd start) are resolved by a parity-predicted fixed-point iteration of table look-ups.  The
//             only serial step is then: wait for the previous window's exit state, follow the tables until
//             the true chain joins a canonical chain (a few look-ups), publish the exit.  Token run
//             lengths are scanned into member ranks (a second look-back chain across windows) and ones /
//             signs are set in rank space.  EOF, the run carried across chunks (rle.h:66-77) and the phantom
//             one before refinement bits (rle.h:91-103) follow the reference exactly.      [dec_parse_kernel]
//   deposit   rank-space bits are expanded into the insignificant positions of each group (software pdep),
//             refinement bits are taken straight from the stream, significance is updated [dec_deposit_kernel]
#include "coder.cuh"

namespace {

constexpr int TG = DWT_TILE_GROUPS;
constexpr int PT = 1024;          // parse threads = slices per window
constexpr int SLICE = 64;         // stream bits per slice; >= the longest token (31 zeros + 1 + 31 payload + sign)
constexpr int WIN_BITS = PT * SLICE;
constexpr int ROW = 66;           // u16 entries per table row (64 + padding against bank conflicts)
constexpr int KDEAD = 255;        // a chain that cannot continue (EOF inside a token, impossible order)
constexpr unsigned short PDEAD = 0xffffu;
constexpr u64 FLAG = 1ull << 63;
constexpr int REFINE_ROUNDS = 6;

// Parse windows grow 64, 128, 256, 512, 1024, 1024, ... slices, so that the many small chunks of the coarse
// levels do not pay for a full window.
__host__ __device__ __forceinline__ int win_size(u32 w)
{
	return w < 4 ? 64 << w : PT;
}
__host__ __device__ __forceinline__ u64 win_start(u32 w) // first slice of window w, relative to the chunk's first slice
{
	return w <= 4 ? 64ull * ((1u << w) - 1u) : 960ull + (u64)(w - 4) * PT;
}
__host__ __device__ __forceinline__ u64 windows_for(u64 slices) // windows needed to cover `slices` slices
{
	if (slices <= 960)
		for (u32 w = 0; w <= 4; ++w)
			if (win_start(w) >= slices)
				return w;
	return 4 + (slices - 960 + PT - 1) / PT;
}

__device__ __forceinline__ unsigned short pack_state(int off, int k)
{
	return (unsigned short)(off | (k << 6));
}
__device__ __forceinline__ int st_off(unsigned short s)
{
	return s & 63;
}
__device__ __forceinline__ int st_k(unsigned short s)
{
	return s == PDEAD ? KDEAD : (s >> 6) & 63;
}

__device__ __forceinline__ u32 group_valid_mask(const Geom &G, int l, int g)
{
	if (g >= G.G[l])
		return 0u;
	long long rem = G.num[l] - (long long)g * 32;
	return rem >= 32 ? 0xffffffffu : ((1u << (int)rem) - 1u);
}

__global__ void __launch_bounds__(TG) dec_prep_kernel(const __grid_constant__ Geom G, int l, const u32 *__restrict__ sig,
                                                       u32 *tile_sums)
{
	__shared__ u32 acc[2];
	if (threadIdx.x < 2)
		acc[threadIdx.x] = 0;
	__syncthreads();
	const int g = blockIdx.x * TG + threadIdx.x;
	const u32 vm = group_valid_mask(G, l, g);
	const u32 s = vm ? sig[g] : 0u;
	u32 m = __reduce_add_sync(0xffffffffu, (u32)__popc(vm & ~s));
	u32 r = __reduce_add_sync(0xffffffffu, (u32)__popc(s));
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&acc[0], m);
		atomicAdd(&acc[1], r);
	}
	__syncthreads();
	if (threadIdx.x < 2)
		tile_sums[2 * blockIdx.x + threadIdx.x] = acc[threadIdx.x];
}

__global__ void __launch_bounds__(1024) dec_tilescan_kernel(DecState *st, const u32 *__restrict__ tile_sums, u32 *tile_base,
                                                             int ntile, u64 *win_state, u64 *win_rank, int nwin_cap,
                                                             int level)
{
	__shared__ u64 ws[32];
	const int tid = threadIdx.x;
	if (st->stopped)
		return;
	const int per = (ntile + 1023) / 1024;
	const int b = tid * per, e = min(b + per, ntile);
	u64 sm = 0, sr = 0;
	for (int i = b; i < e; ++i) {
		sm += tile_sums[2 * i];
		sr += tile_sums[2 * i + 1];
	}
	u64 tm, tr;
	u64 bm = block_exscan_u64(sm, ws, &tm);
	u64 br = block_exscan_u64(sr, ws, &tr);
	u32 rm = (u32)bm, rr = (u32)br;
	for (int i = b; i < e; ++i) {
		u32 m = tile_sums[2 * i], r = tile_sums[2 * i + 1];
		tile_base[2 * i] = rm;
		tile_base[2 * i + 1] = rr;
		rm += m;
		rr += r;
	}
	// windows this chunk can touch at most: from its first slice to the end of the stream
	const u64 first = st->bitpos & ~63ull;
	u64 nwin = windows_for((st->end_bits > first ? (st->end_bits - first) >> 6 : 0) + 2) + 1;
	if (nwin > (u64)nwin_cap)
		nwin = nwin_cap;
	for (u64 w = tid; w < nwin; w += 1024) {
		win_state[w] = 0;
		win_rank[w] = 0;
	}
	if (tid == 0) {
		st->n_member = (u32)tm;
		st->n_ref = (u32)tr;
		if (st->level < level)
			st->level = level; // decode.c:203,219-220,236-237: the chunk is started
		st->ticket = 0;
		st->published = 0;
		st->done = 0;
		st->c_bitpos = st->bitpos;
		st->c_order = st->order;
		st->c_pending = st->pending;
	}
}

__device__ __forceinline__ u64 peek64(const u32 *__restrict__ s, u64 pos)
{
	const u64 w = pos >> 5;
	const int sh = (int)(pos & 31);
	const u64 lo = (u64)__ldg(s + w) | ((u64)__ldg(s + w + 1) << 32);
	u64 v = lo >> sh;
	if (sh)
		v |= (u64)__ldg(s + w + 2) << (64 - sh);
	return v;
}

// one [VLI] token at (pos, k): returns false when it cannot be read completely (EOF / invalid)
__device__ __forceinline__ bool read_vli(const u32 *__restrict__ s, u64 end_bits, u64 pos, int k, u64 *n, int *len,
                                         int *knext, u64 *word)
{
	if (pos >= end_bits)
		return false;
	const u64 w = peek64(s, pos);
	const int u = w ? __ffsll((long long)w) - 1 : 64;
	const int e = k + u;
	if (e > 31)
		return false;
	const int L = u + 1 + e;
	if (pos + L > end_bits)
		return false;
	const u32 payload = (u32)(w >> (u + 1)) & (u32)((1ull << e) - 1ull);
	*n = ((1ull << e) - (1ull << k)) + payload;
	*len = L;
	*knext = e >= 2 ? e - 2 : 0;
	*word = w;
	return true;
}

// slice-local token step on the two 64-bit words of a slice (w0 = the slice, w1 = the 64 bits behind it):
// token at offset d (< 64) with order k; avail = stream bits left from the slice start.
// Returns the offset behind [VLI][sign] (may be >= 64) or -1 when the chain dies.
// Only the unary prefix decides a token's length, and it is at most 31 zeros, so a 32-bit window suffices.
__device__ __forceinline__ int slice_step(u64 w0, u64 w1, int avail, int d, int &k)
{
	const u32 lo = d < 32 ? (u32)w0 : (u32)(w0 >> 32);
	const u32 hi = d < 32 ? (u32)(w0 >> 32) : (u32)w1;
	const u32 bits = __funnelshift_r(lo, hi, d & 31);
	const int u = bits ? __ffs((int)bits) - 1 : 32;
	const int e = k + u;
	const int L = u + 1 + e;
	if (e > 31 || d >= avail || d + L > avail)
		return -1;
	k = e >= 2 ? e - 2 : 0;
	return d + L + 1;
}

__device__ __forceinline__ int clamp_avail(u64 end_bits, u64 lo_bit)
{
	const long long av = (long long)end_bits - (long long)lo_bit;
	return av > (1 << 30) ? (1 << 30) : (av < -(1 << 30) ? -(1 << 30) : (int)av);
}

// the two 64-bit words a slice's token steps can touch (zero behind the padded end of the stream)
__device__ __forceinline__ void load_slice(const u32 *__restrict__ stream, u64 end_bits, u64 slice, u64 &w0, u64 &w1)
{
	const u64 lo = slice << 6;
	w0 = lo < end_bits + 128 ? __ldg((const u64 *)stream + slice) : 0ull;
	w1 = lo < end_bits + 64 ? __ldg((const u64 *)stream + slice + 1) : 0ull;
}

// exit state of a slice for the entry (off, k), using the slice's order-0 table for the order-0 part
__device__ __forceinline__ unsigned short slice_exit(const unsigned short *row, u64 w0, u64 w1, int avail,
                                                     unsigned short entry)
{
	if (entry == PDEAD)
		return PDEAD;
	int d = st_off(entry), k = st_k(entry);
	while (k != 0) { // excursion at a non-zero order: plain token steps until the order is back to 0
		d = slice_step(w0, w1, avail, d, k);
		if (d < 0)
			return PDEAD;
		if (d >= 64)
			return pack_state(d - 64, k);
	}
	return row[d];
}

// members consumed by the tokens that start inside the slice [lo, lo+64), beginning at (pos, k)
__device__ __forceinline__ u64 count_slice(const u32 *__restrict__ s, u64 end_bits, u64 lo, u64 pos, int k)
{
	const u64 lim = lo + SLICE;
	u64 csum = 0;
	while (k != KDEAD && pos < lim) {
		u64 n, w;
		int len, kn;
		if (!read_vli(s, end_bits, pos, k, &n, &len, &kn, &w))
			break;
		csum += n + 1;
		pos += len + 1;
		k = kn;
	}
	return csum;
}

enum { EV_NONE = 0, EV_COVERED = 1, EV_PENDING = 2, EV_STOP = 3 };

// refinement pass bookkeeping of one chunk (decode.c:89-98,206,223,240); runs in exactly one thread
__device__ void finish_chunk(DecState *st, bool stop, u64 bitpos, int order, u32 pending, u32 nref, int chan, int level)
{
	const u64 end_bits = st->end_bits;
	const bool sig_done = !stop; // every member has its symbol (or is covered by the carried run)
	bool complete = sig_done;
	int ref_valid = 0;
	const u64 ref_pos = bitpos;
	if (sig_done && nref > 0) {
		if (pending > 1) {
			stop = true; // rle.h:98-99: a pending run must end exactly at the phantom one
			complete = false;
		} else {
			pending = 0;
			ref_valid = 1;
			if (bitpos + nref > end_bits) {
				stop = true; // partial refinement: the deposit keeps the bits before EOF
				complete = false;
			} else {
				bitpos += nref;
			}
		}
	}
	st->ref_bitpos = ref_pos;
	st->ref_valid = ref_valid;
	st->bitpos = bitpos;
	st->order = order == KDEAD ? 0 : order;
	st->pending = pending;
	st->stopped = stop ? 1 : 0;
	st->chunk_done += 1;
	if (complete)
		st->missing[chan * 16 + level] -= 1;
	__threadfence();
	st->done = 1;
}

__global__ void __launch_bounds__(PT, 1) dec_parse_kernel(DecState *st, const u32 *__restrict__ stream, u32 *ones_rank,
                                                           u32 *sign_rank, u64 *win_state, u64 *win_rank, int nwin_cap,
                                                           int chan, int level)
{
	extern __shared__ unsigned short T[];        // [PT][ROW] order-0 transfer tables of the window's slices
	__shared__ u64 ws[32];
	__shared__ u32 wmap[32];
	__shared__ unsigned short ecan[2][PT];       // entries of the two canonical chains
	__shared__ unsigned short xcan[2][PT];       // their exits
	__shared__ unsigned short etrue[PT];         // entries of the true chain where it was stepped exactly
	__shared__ u64 winfo[PT];                    // per slice: predicted entries of both chains + their next wrong link
	__shared__ u32 xpair[PT];                    // per slice: exact exits of both chains for those entries
	__shared__ unsigned char mark[PT];           // walker marks: 1/2 = follows predicted chain 0/1 from here, 3 = exact, 4 = dead
	__shared__ u64 sw[PT + 1];                   // the window's stream words
	__shared__ u32 s_w;
	__shared__ int s_done, s_merge_at, winner;
	__shared__ u64 s_rank_excl;
	__shared__ u64 f_pos;
	__shared__ int f_k, f_event;
	__shared__ u32 f_pending;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (st->stopped)
		return;

	// ---- chunk entry state (identical in every CTA; side effects only in block 0)
	const u64 end_bits = st->end_bits;
	const u64 R = st->n_member;
	const u32 nref = st->n_ref;
	u64 bitpos = st->c_bitpos; // entry snapshot taken by dec_tilescan_kernel: never written while this kernel runs
	const int order = st->c_order;
	u32 pending = st->c_pending;
	u64 r0 = 0; // members already accounted for
	bool stop = false;
	// a run carried in from earlier chunks (rle.h:66-77): (pending-1) zeros, then a one
	if (pending > 0) {
		if ((u64)pending - 1 >= R) {
			pending -= (u32)R;
			r0 = R;
		} else {
			const u64 rk = pending - 1;
			if (blockIdx.x == 0 && tid == 0)
				atomicOr(ones_rank + (rk >> 5), 1u << (rk & 31));
			if (bitpos < end_bits) {
				if (blockIdx.x == 0 && tid == 0 && (peek64(stream, bitpos) & 1ull))
					atomicOr(sign_rank + (rk >> 5), 1u << (rk & 31));
				bitpos += 1;
			} else {
				stop = true; // the sign bit hits EOF: the magnitude bit stays (decode.c:80-86)
			}
			r0 = (u64)pending;
			pending = 0;
		}
	}
	if (stop || pending != 0 || r0 >= R) { // no token to parse in this chunk
		__syncthreads(); // every thread has read the entry state before it is overwritten
		if (blockIdx.x == 0 && tid == 0)
			finish_chunk(st, stop, bitpos, order, pending, nref, chan, level);
		return;
	}
	const u64 Rrem = R - r0;
	const u64 S0 = bitpos >> 6; // first slice of window 0
	u64 nwin = windows_for(((end_bits - (S0 << 6)) >> 6) + 2) + 1;
	{
		// a token is at most 33 bits per member it covers (order <= 31), so the pass cannot reach further
		const u64 bound = windows_for((Rrem * 33 + 63) / 64 + 2) + 1;
		if (nwin > bound)
			nwin = bound;
	}
	if (nwin > (u64)nwin_cap)
		nwin = nwin_cap;

	for (;;) {
		__syncthreads();
		if (tid == 0) {
			const u32 tk = atomicAdd(&st->ticket, 1u);
			// speculation is throttled: at most as many unconfirmed windows as confirmed ones (+2), so a chunk
			// that ends in its first window does not pay for 148 table builds behind it
			volatile u32 *pub = &st->published;
			volatile int *dn = &st->done;
			int d = *dn;
			while (!d && tk < nwin && tk >= 2u * *pub + 2u)
				d = *dn;
			s_w = tk;
			s_done = d;
			winner = PT;
		}
		__syncthreads();
		const u32 w = s_w;
		if (s_done || w >= nwin) {
			// a ticket holder must publish even when the pass is over: a later window may be waiting on it
			if (tid == 0 && w < nwin) {
				win_state[w] = FLAG | PDEAD;
				win_rank[w] = FLAG | (FLAG - 1);
				__threadfence();
			}
			break;
		}
		const long long t_win0 = clock64();
		const int sz = win_size(w);                 // active slices (threads) of this window
		const u64 wslice0 = S0 + win_start(w);      // its first slice
		const u64 slice = wslice0 + tid;
		const u64 sub_lo = slice << 6;
		const int avail = clamp_avail(end_bits, sub_lo);
		u64 w0, w1;
		load_slice(stream, end_bits, slice, w0, w1);
		unsigned short *row = T + tid * ROW;

		// (1) order-0 transfer table of my slice, by dynamic programming from the slice end
		for (int d = tid < sz ? 63 : -1; d >= 0; --d) {
			int k = 0;
			int nd = slice_step(w0, w1, avail, d, k);
			unsigned short v = PDEAD;
			while (nd >= 0) {
				if (nd >= 64) {
					v = pack_state(nd - 64, k);
					break;
				}
				if (k == 0) {
					v = row[nd];
					break;
				}
				nd = slice_step(w0, w1, avail, nd, k);
			}
			row[d] = v;
		}
		// (2) predicted chains.  Chain c starts with a token at offset c of the window at order 0.  The parity
		// class every slice is entered with comes from a scan of 2-state maps; the predicted entry of a slice is
		// the table exit of its predecessor, and ONE exact evaluation per slice tells which links
		// (slice i -> i+1) the prediction got right.  No iteration: wrong links are repaired by the walker.
		sw[tid] = w0;
		if (tid == sz - 1)
			sw[sz] = w1;
		mark[tid] = 0;
		u32 mymap = 0;
#pragma unroll
		for (int c = 0; c < 2; ++c) {
			unsigned short x = row[c];
			u32 out = (x != PDEAD && st_k(x) == 0) ? (u32)(st_off(x) & 1) : 0u;
			mymap |= out << c;
		}
		u32 inc = mymap; // bit b = class after this slice when entered with class b
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			u32 t = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d)
				inc = ((inc >> (t & 1u)) & 1u) | (((inc >> ((t >> 1) & 1u)) & 1u) << 1);
		}
		if (lane == 31)
			wmap[wid] = inc;
		__syncthreads();
		u32 pre = 2u; // identity: class 0 -> 0, class 1 -> 1
		for (int i = 0; i < wid; ++i) {
			u32 m = wmap[i];
			pre = ((m >> (pre & 1u)) & 1u) | (((m >> ((pre >> 1) & 1u)) & 1u) << 1);
		}
		u32 excl = __shfl_up_sync(0xffffffffu, inc, 1);
		if (lane > 0)
			pre = ((excl >> (pre & 1u)) & 1u) | (((excl >> ((pre >> 1) & 1u)) & 1u) << 1);
		// pre: class my slice is entered with on chain 0 (bit 0) and chain 1 (bit 1)
		xcan[0][tid] = row[pre & 1u];
		xcan[1][tid] = row[(pre >> 1) & 1u];
		__syncthreads();
		unsigned short e_c[2], x_c[2];
#pragma unroll
		for (int c = 0; c < 2; ++c)
			e_c[c] = tid == 0 ? pack_state(c, 0) : xcan[c][tid - 1];
		__syncthreads();
#pragma unroll
		for (int c = 0; c < 2; ++c)
			x_c[c] = tid < sz ? slice_exit(row, w0, w1, avail, e_c[c]) : PDEAD;
		// a fixed number of refinement rounds (entry <- predecessor's exact exit) repairs the links where the
		// predicted entry needed more than one slice to join the chain
		for (int it = 0; it < REFINE_ROUNDS; ++it) {
			xcan[0][tid] = x_c[0];
			xcan[1][tid] = x_c[1];
			__syncthreads();
			if (tid > 0 && tid < sz) {
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					const unsigned short ne = xcan[c][tid - 1];
					if (ne != e_c[c]) {
						e_c[c] = ne;
						x_c[c] = slice_exit(row, w0, w1, avail, ne);
					}
				}
			}
			__syncthreads();
		}
#pragma unroll
		for (int c = 0; c < 2; ++c) {
			ecan[c][tid] = e_c[c];
			xcan[c][tid] = x_c[c];
		}
		__syncthreads();
		// nbad[c] = first slice j >= i whose exit does not match the predicted entry of slice j+1;
		// everything the walker needs about slice i goes into one 64-bit record
		u32 nb[2];
#pragma unroll
		for (int c = 0; c < 2; ++c) {
			const bool ok = tid < sz - 1 && x_c[c] == ecan[c][tid + 1];
			u32 v = ok ? 0xffffu : (u32)tid;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				u32 t = __shfl_down_sync(0xffffffffu, v, d);
				if (lane + d < 32)
					v = min(v, t);
			}
			if (lane == 0)
				wmap[wid] = v;
			__syncthreads();
			for (int i = wid + 1; i < PT / 32; ++i)
				v = min(v, wmap[i]);
			nb[c] = v;
			__syncthreads();
		}
		winfo[tid] = (u64)e_c[0] | ((u64)e_c[1] << 16) | ((u64)nb[0] << 32) | ((u64)nb[1] << 48);
		xpair[tid] = (u32)x_c[0] | ((u32)x_c[1] << 16);
		__syncthreads();

		// (3) the serial step: previous window's exit -> walk the true chain.  Where it coincides with a predicted
		// chain it jumps to that chain's next wrong link; elsewhere it steps exactly, one slice at a time.
		if (tid == 0) {
			const long long t_wait0 = clock64();
			unsigned short state = PDEAD;
			bool over = false;
			if (w == 0) {
				state = pack_state((int)(bitpos & 63), order);
			} else {
				volatile u64 *src = win_state + (w - 1);
				volatile int *dn = &st->done;
				u64 v;
				u32 spins = 0;
				while (!((v = *src) & FLAG))
					if ((++spins & 31u) == 0 && *dn) {
						over = true;
						break;
					}
				state = over ? PDEAD : (unsigned short)(v & 0xffffu);
			}
			int i = 0;
			u32 nexact = 0, njump = 0;
			const long long t_walk0 = clock64();
			if (over) {
				i = -1; // the pass ended in an earlier window: nothing to do here
			} else if (state == PDEAD) {
				i = -2; // dead on arrival
			} else {
				long long av = (long long)end_bits - (long long)(wslice0 << 6);
				const int availw = av > (1 << 30) ? (1 << 30) : (av < -(1 << 30) ? -(1 << 30) : (int)av);
				u32 st32 = state;
				while (i < sz) {
					if (st32 == PDEAD) { // the chain died inside this window: later slices have no tokens
						mark[i] = 4;
						break;
					}
					const u64 rec = winfo[i];
					const int c = st32 == (u32)(rec & 0xffffu) ? 0 : (st32 == (u32)((rec >> 16) & 0xffffu) ? 1 : -1);
					if (c >= 0) {
						const int j = (int)((rec >> (32 + 16 * c)) & 0xffffu);
						mark[i] = (unsigned char)(1 + c);
						st32 = (xpair[j] >> (16 * c)) & 0xffffu;
						i = j + 1;
						++njump;
						continue;
					}
					mark[i] = 3;
					etrue[i] = (unsigned short)st32;
					++nexact;
					// exact step through slice i: table look-up at order 0, plain token steps otherwise
					int off = st32 & 63, k = (st32 >> 6) & 63;
					const unsigned short *trow = T + i * ROW;
					if (k != 0) {
						const u64 a = sw[i], b = sw[i + 1];
						const int avail_i = availw - 64 * i;
						for (;;) {
							const u32 lo = off < 32 ? (u32)a : (u32)(a >> 32);
							const u32 hi = off < 32 ? (u32)(a >> 32) : (u32)b;
							const u32 bits = __funnelshift_r(lo, hi, off & 31);
							const int u = bits ? __ffs((int)bits) - 1 : 32;
							const int e = k + u;
							const int L = u + 1 + e;
							if (e > 31 || off >= avail_i || off + L > avail_i) {
								off = -1;
								break;
							}
							k = e >= 2 ? e - 2 : 0;
							off += L + 1;
							if (off >= 64 || k == 0)
								break;
						}
					}
					if (off < 0)
						st32 = PDEAD;
					else if (off >= 64)
						st32 = (u32)pack_state(off - 64, k);
					else
						st32 = trow[off];
					++i;
				}
				state = (unsigned short)st32;
				i = 0;
			}
			win_state[w] = FLAG | state; // flag and value travel in one 64-bit word: no fence needed
			*(volatile u32 *)&st->published = w + 1;
			if (over)
				win_rank[w] = FLAG | (FLAG - 1);
			const long long t_walk1 = clock64();
			atomicAdd(&st->dbg_walk, nexact);
			atomicAdd(&st->dbg_iters, njump);
			atomicAdd(&st->dbg_cyc[0], (u64)(t_wait0 - t_win0));
			atomicAdd(&st->dbg_cyc[1], (u64)(t_walk0 - t_wait0));
			atomicAdd(&st->dbg_cyc[2], (u64)(t_walk1 - t_walk0));
			s_merge_at = i;
		}
		__syncthreads();
		if (s_merge_at == -1)
			break;
		const bool dead_window = s_merge_at == -2;
		// every slice is governed by the last mark at or before it
		unsigned short my_entry = PDEAD;
		{
			int gm = mark[tid] ? tid : -1;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				int t = __shfl_up_sync(0xffffffffu, gm, d);
				if (lane >= d)
					gm = max(gm, t);
			}
			if (lane == 31)
				wmap[wid] = (u32)gm;
			__syncthreads();
			for (int i = 0; i < wid; ++i)
				gm = max(gm, (int)wmap[i]);
			if (!dead_window && gm >= 0) {
				const int m = mark[gm];
				if (m == 3)
					my_entry = etrue[tid];
				else if (m == 1 || m == 2)
					my_entry = ecan[m - 1][tid];
			}
		}
		if (tid >= sz)
			my_entry = PDEAD;
		const u64 e_pos = sub_lo + st_off(my_entry);
		const int e_k = st_k(my_entry);

		// (4) members consumed per slice -> ranks inside the window; inclusive count chained across windows
		const u64 csum = count_slice(stream, end_bits, sub_lo, e_pos, e_k);
		u64 total;
		u64 cum = block_exscan_u64(csum, ws, &total); // syncs inside
		if (tid == 0) {
			u64 base = 0;
			if (w > 0) {
				volatile u64 *src = win_rank + (w - 1);
				u64 v;
				while (!((v = *src) & FLAG))
					;
				base = v & ~FLAG;
			}
			u64 incl = base + total;
			if (incl >= FLAG)
				incl = FLAG - 1;
			win_rank[w] = FLAG | incl;
			s_rank_excl = base;
			f_event = EV_NONE;
		}
		__syncthreads();
		const u64 rank_excl = s_rank_excl;
		if (rank_excl > Rrem || dead_window)
			continue; // this window lies behind the end of the chunk's significance pass
		cum += rank_excl;

		// (5) the walk that writes ones and signs, and finds where the pass ends
		{
			u64 pos = e_pos;
			int k = e_k;
			int ev = EV_NONE;
			u64 ev_pos = 0;
			int ev_k = 0;
			u32 ev_pending = 0;
			const u64 lim = sub_lo + SLICE;
			while (k != KDEAD && pos < lim) {
				if (cum >= Rrem) {
					ev = EV_COVERED;
					ev_pos = pos;
					ev_k = k;
					break;
				}
				u64 n, wd;
				int len, kn;
				if (!read_vli(stream, end_bits, pos, k, &n, &len, &kn, &wd)) {
					ev = EV_STOP;
					break;
				}
				const u64 one = cum + n;
				if (one < Rrem) {
					const u64 rk = r0 + one;
					atomicOr(ones_rank + (rk >> 5), 1u << (rk & 31));
					if (pos + len + 1 > end_bits) {
						ev = EV_STOP; // sign bit beyond EOF: the magnitude bit stays
						break;
					}
					if ((wd >> len) & 1ull)
						atomicOr(sign_rank + (rk >> 5), 1u << (rk & 31));
					cum = one + 1;
					pos += len + 1;
					k = kn;
				} else {
					ev = EV_PENDING; // the run reaches past this chunk's members (rle.h:74-76)
					ev_pending = (u32)(n - (Rrem - cum) + 1);
					ev_pos = pos + len;
					ev_k = kn;
					break;
				}
			}
			if (ev != EV_NONE)
				atomicMin(&winner, tid);
			__syncthreads();
			if (ev != EV_NONE && winner == tid) {
				f_event = ev;
				f_pos = ev_pos;
				f_k = ev_k;
				f_pending = ev_pending;
			}
			__syncthreads();
		}
		if (f_event != EV_NONE) {
			if (tid == 0) {
				atomicAdd(&st->dbg_windows, w + 1);
				if (f_event == EV_STOP)
					finish_chunk(st, true, bitpos, order, 0, nref, chan, level);
				else
					finish_chunk(st, false, f_pos, f_k, f_event == EV_PENDING ? f_pending : 0u, nref, chan, level);
			}
			break;
		}
	}
}

__global__ void __launch_bounds__(TG) dec_deposit_kernel(const __grid_constant__ Geom G, int l, int chunk_seq,
                                                          u32 *plane_words, u32 *sign_words, u32 *sig,
                                                          const u32 *__restrict__ tile_base,
                                                          const u32 *__restrict__ ones_rank,
                                                          const u32 *__restrict__ sign_rank,
                                                          const u32 *__restrict__ stream, const DecState *st)
{
	__shared__ u64 ws[32];
	if (st->chunk_done != chunk_seq)
		return; // the parse of this chunk never ran (decoding stopped earlier)
	const int g = blockIdx.x * TG + threadIdx.x;
	const u32 vm = group_valid_mask(G, l, g);
	const u32 s = vm ? sig[g] : 0u;
	const u32 member = vm & ~s;
	const u32 nm = __popc(member), nr = __popc(s);
	u64 tot;
	const u64 ex = block_exscan_u64((u64)nm | ((u64)nr << 32), ws, &tot);
	if (!vm)
		return;
	u32 B = 0;
	if (nm) {
		const u64 off = (u64)tile_base[2 * blockIdx.x] + (u32)ex;
		u32 ob = bits_get32(ones_rank, off);
		if (nm < 32)
			ob &= (1u << nm) - 1u;
		if (ob) {
			B = bit_expand(ob, member);
			u32 sb = bits_get32(sign_rank, off) & ob;
			if (sb)
				sign_words[g] |= bit_expand(sb, member);
		}
	}
	if (nr && st->ref_valid) {
		const u64 pos = st->ref_bitpos + tile_base[2 * blockIdx.x + 1] + (u32)(ex >> 32);
		const u64 end = st->end_bits;
		if (pos < end) {
			u32 rb = (u32)peek64(stream, pos);
			u64 avail = end - pos;
			u32 take = nr;
			if (avail < take)
				take = (u32)avail;
			if (take < 32)
				rb &= (1u << take) - 1u;
			B |= bit_expand(rb, s);
		}
	}
	plane_words[g] = B;
	if (B)
		sig[g] = s | B;
}

constexpr size_t PARSE_SMEM = (size_t)PT * ROW * sizeof(unsigned short);

} // namespace

int dec_setup(void)
{
	CUDA_OK(cudaFuncSetAttribute(dec_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PARSE_SMEM));
	return 0;
}

int dec_chunk(const Geom &g, const Sched &hs, const DecBuffers &b, int j, cudaStream_t st, long long *launches)
{
	const int c = hs.chan[j], l = hs.level[j], p = hs.plane[j];
	const int ntile = g.ntile[l];
	const size_t rank_words = (size_t)g.G[l] + 4;
	// ones_rank and sign_rank are adjacent halves of one buffer
	CUDA_OK(cudaMemsetAsync(b.ones_rank, 0, rank_words * 4, st));
	CUDA_OK(cudaMemsetAsync(b.sign_rank, 0, rank_words * 4, st));
	u32 *sig = b.sig + (size_t)c * g.GT + g.gbase[l];
	u32 *plane_words = b.bs + hs.bsbase[c] + (long long)p * g.GT + g.gbase[l];
	u32 *sign_words = b.bs + hs.bsbase[c] + (long long)hs.planes[c] * g.GT + g.gbase[l];
	dec_prep_kernel<<<ntile, TG, 0, st>>>(g, l, sig, b.tile_sums);
	dec_tilescan_kernel<<<1, 1024, 0, st>>>(b.state, b.tile_sums, b.tile_base, ntile, b.win_state, b.win_rank, b.nwin_cap,
	                                        l);
	dec_parse_kernel<<<b.parse_ctas, PT, PARSE_SMEM, st>>>(b.state, b.stream, b.ones_rank, b.sign_rank, b.win_state,
	                                                       b.win_rank, b.nwin_cap, c, l);
	dec_deposit_kernel<<<ntile, TG, 0, st>>>(g, l, j + 1, plane_words, sign_words, sig, b.tile_base, b.ones_rank,
	                                         b.sign_rank, b.stream, b.state);
	*launches += 4;
	CUDA_OK(cudaGetLastError());
	return 0;
}
