// coder_dec.cu -- the bit-plane decoder of decode.c:67-100,187-243 + rle.h + vli.h + bits.h on the GPU.
//
// The significance pass of a chunk is a serial chain of [adaptive-Rice run][sign] tokens whose end is only known
// once the runs cover the chunk's members, and the chunks follow each other in the stream.  The work is split so
// that everything expensive is independent of the chunks and runs once over the whole stream:
//
//   scan     (dec_scan_kernel, one CTA per window of 512 slices of 64 bits)  two canonical token chains per
//            window, seeded at bit 0 / bit 1 of the window with order 0, are resolved exactly by a parity-class
//            prediction and a neighbour fixed-point iteration.  Per slice: entry state of both chains, prefix of
//            members (run + 1) and of tokens along each chain.  Chains that start anywhere else join one of the
//            two within a few slices (the code is self-synchronising), so these tables describe the true chain.
//   link     (dec_link_kernel, one thread per window and class)  where does the exit state of class q of window
//            w-1 join the canonical chains of window w, and what does the whole window consume on that path.
//   resolve  (dec_resolve_kernel, ONE warp)  the only serial step: walks the chunk schedule.  A chunk start is
//            followed token by token until it joins a canonical chain; whole windows are then consumed 32 at a
//            time through the link records (a warp scan over 2-state class maps); the window where the members
//            run out is searched on the per-slice prefixes and the last slice is stepped exactly.  EOF, the run
//            carried across chunks (rle.h:66-77) and the phantom one before refinement bits (rle.h:91-103)
//            follow the reference exactly.  Output: per chunk its rank offset / refinement position, per
//            (chunk, window) visit a segment record.
//   emit     (dec_emit_kernel, one CTA per segment)  every slice knows its entry state and member rank: ones and
//            signs are set in rank space.
//   deposit  (per plane depth, all channels and levels at once)  rank-space bits are expanded into the
//            insignificant positions of each group (software pdep), refinement bits come straight from the
//            stream, significance is updated.
#include "coder.cuh"

#include <stdlib.h>

namespace {

constexpr int TG = DWT_TILE_GROUPS;
constexpr int WS = DWT_DEC_WS;
constexpr u32 PDEAD = 0xffffu;    // a chain that cannot continue (EOF inside a token, impossible order)
constexpr int LINK_CAP = DWT_DEC_WS; // exact slice steps the link pass spends on one (window, chain): the whole window
constexpr u32 LINK_OPEN = 0xfffeu;  // link record: the chain had not joined after LINK_CAP slices
constexpr int EXACT_FIRST = 4;     // resolver: slices of a chunk start that are stepped with full bookkeeping before it changes gear
constexpr u64 DEATH = 1ull << 48; // member count charged to a slice in which a canonical chain dies: more than any chunk holds

enum { EV_NONE = 0, EV_COVERED = 1, EV_PENDING = 2, EV_STOP = 3 };

__device__ __forceinline__ int clamp_avail(u64 end_bits, u64 lo_bit)
{
	const long long av = (long long)end_bits - (long long)lo_bit;
	return av > (1 << 30) ? (1 << 30) : (av < -(1 << 30) ? -(1 << 30) : (int)av);
}

// the two 64-bit words a slice's tokens can touch (zero behind the padded end of the stream)
__device__ __forceinline__ void load_slice(const u32 *__restrict__ stream, u64 end_bits, u64 slice, u64 &w0, u64 &w1)
{
	const u64 lo = slice << 6;
	w0 = lo < end_bits + 128 ? __ldg((const u64 *)stream + slice) : 0ull;
	w1 = lo < end_bits + 64 ? __ldg((const u64 *)stream + slice + 1) : 0ull;
}

__device__ __forceinline__ u64 shfl_up_u64(u64 v, int d)
{
	const u32 lo = __shfl_up_sync(0xffffffffu, (u32)v, d), hi = __shfl_up_sync(0xffffffffu, (u32)(v >> 32), d);
	return (u64)lo | ((u64)hi << 32);
}

__device__ __forceinline__ u64 bits_from(u64 a, u64 b, int d) // 64 stream bits from offset d (0..63) of the pair
{
	return d ? (a >> d) | (b << (64 - d)) : a;
}

// exit state of a slice entered with `entry` (first token at offset entry & 63 with order entry >> 6):
// the state the next slice is entered with.  Only the unary prefix decides a token's length, and a valid
// one has at most 31 zeros, so a 32-bit window is enough.  avail = stream bits left from the slice start.
__device__ __forceinline__ u32 slice_exit(u64 a, u64 b, int avail, u32 entry)
{
	if (entry == PDEAD)
		return PDEAD;
	int d = (int)(entry & 63u), k = (int)(entry >> 6);
	do {
		const u32 lo = d < 32 ? (u32)a : (u32)(a >> 32);
		const u32 hi = d < 32 ? (u32)(a >> 32) : (u32)b;
		const u32 bits = __funnelshift_r(lo, hi, d & 31);
		const int u = bits ? __ffs((int)bits) - 1 : 32;
		const int e = k + u;
		const int L = u + 1 + e;
		if (e > 31 || d + L > avail)
			return PDEAD;
		k = e >= 2 ? e - 2 : 0;
		d += L + 1;
	} while (d < 64);
	return (u32)(d - 64) | ((u32)k << 6);
}

// same walk, also counting the members (run + 1) and tokens of the tokens that start in the slice
__device__ __forceinline__ u32 slice_walk(u64 a, u64 b, int avail, u32 entry, u64 &mem, u32 &ntok)
{
	mem = 0;
	ntok = 0;
	if (entry == PDEAD)
		return PDEAD;
	int d = (int)(entry & 63u), k = (int)(entry >> 6);
	do {
		const u64 w = bits_from(a, b, d);
		const u32 lo = (u32)w;
		const int u = lo ? __ffs((int)lo) - 1 : 32;
		const int e = k + u;
		const int L = u + 1 + e;
		if (e > 31 || d + L > avail)
			return PDEAD;
		const u32 payload = (u32)(w >> (u + 1)) & ((1u << e) - 1u);
		mem += (u64)((1u << e) - (1u << k)) + payload + 1ull;
		++ntok;
		k = e >= 2 ? e - 2 : 0;
		d += L + 1;
	} while (d < 64);
	return (u32)(d - 64) | ((u32)k << 6);
}

// ---- order-0 token table.  93 % of the tokens are coded at order 0 with at most two leading zeros (run < 7):
// "1s", "01ps", "001pps" -- 2, 4 or 6 bits, and the order stays 0.  TOKLUT[12 stream bits] decodes as many of
// those as fit completely: bits consumed (4 bits) | start of the last token (4) | tokens (3) | members (6).
// A window that starts with anything else (longer run, non-zero order) yields 0 and takes the generic step.
constexpr int LUT_BITS = 12;

__host__ __device__ inline u32 toklut_entry(u32 bits)
{
	int pos = 0, last = 0, ntok = 0, mem = 0;
	for (;;) {
		int u = 0;
		while (u < 3 && pos + u < LUT_BITS && !((bits >> (pos + u)) & 1u))
			++u;
		if (u >= 3 || pos + 2 * u + 2 > LUT_BITS)
			break;
		const u32 payload = (bits >> (pos + u + 1)) & ((1u << u) - 1u);
		mem += (1 << u) - 1 + (int)payload + 1;
		++ntok;
		last = pos;
		pos += 2 * u + 2;
	}
	return (u32)pos | ((u32)last << 4) | ((u32)ntok << 8) | ((u32)mem << 11);
}

// second word of the table: which of the entry's members are ones (bit i = member i, 16 bits) and, on the same
// positions, the sign bits of those ones (16 bits)
__host__ __device__ inline u32 toklut_masks(u32 bits)
{
	int pos = 0, mem = 0;
	u32 ones = 0, signs = 0;
	for (;;) {
		int u = 0;
		while (u < 3 && pos + u < LUT_BITS && !((bits >> (pos + u)) & 1u))
			++u;
		if (u >= 3 || pos + 2 * u + 2 > LUT_BITS)
			break;
		const u32 payload = (bits >> (pos + u + 1)) & ((1u << u) - 1u);
		const int n = (1 << u) - 1 + (int)payload;
		ones |= 1u << (mem + n);
		if ((bits >> (pos + 2 * u + 1)) & 1u)
			signs |= 1u << (mem + n);
		mem += n + 1;
		pos += 2 * u + 2;
	}
	return ones | (signs << 16);
}

__device__ __forceinline__ u32 window32(u64 a, u64 b, int d) // 32 stream bits from offset d (0..63) of the pair
{
	const u32 lo = d < 32 ? (u32)a : (u32)(a >> 32);
	const u32 hi = d < 32 ? (u32)(a >> 32) : (u32)b;
	return __funnelshift_r(lo, hi, d & 31);
}

__device__ __forceinline__ u32 slice_exit_lut(const u32 *lut, u64 a, u64 b, int avail, u32 entry)
{
	if (entry == PDEAD)
		return PDEAD;
	int d = (int)(entry & 63u), k = (int)(entry >> 6);
	const bool lut_ok = avail >= 64 + LUT_BITS + 4; // every window looked up lies inside the stream
	do {
		const u32 bits = window32(a, b, d);
		if (k == 0 && lut_ok) {
			const u32 t = lut[bits & ((1u << LUT_BITS) - 1u)];
			if ((t & 15u) && d + (int)((t >> 4) & 15u) < 64) { // every token of the entry starts in this slice
				d += (int)(t & 15u);
				continue;
			}
		}
		const int u = bits ? __ffs((int)bits) - 1 : 32;
		const int e = k + u;
		const int L = u + 1 + e;
		if (e > 31 || d + L > avail)
			return PDEAD;
		k = e >= 2 ? e - 2 : 0;
		d += L + 1;
	} while (d < 64);
	return (u32)(d - 64) | ((u32)k << 6);
}

__device__ __forceinline__ u32 slice_walk_lut(const u32 *lut, u64 a, u64 b, int avail, u32 entry, u64 &mem, u32 &ntok)
{
	mem = 0;
	ntok = 0;
	if (entry == PDEAD)
		return PDEAD;
	int d = (int)(entry & 63u), k = (int)(entry >> 6);
	u32 msum = 0;
	const bool lut_ok = avail >= 64 + LUT_BITS + 4;
	do {
		const u32 bits = window32(a, b, d);
		if (k == 0 && lut_ok) {
			const u32 t = lut[bits & ((1u << LUT_BITS) - 1u)];
			if ((t & 15u) && d + (int)((t >> 4) & 15u) < 64) {
				d += (int)(t & 15u);
				ntok += (t >> 8) & 7u;
				msum += t >> 11;
				continue;
			}
		}
		const int u = bits ? __ffs((int)bits) - 1 : 32;
		const int e = k + u;
		const int L = u + 1 + e;
		if (e > 31 || d + L > avail) {
			mem += msum;
			return PDEAD;
		}
		// the payload lies inside the 32-bit window for all but the long runs: no 64-bit shifts then
		const u32 payload = (L <= 32 ? bits >> (u + 1) : (u32)(bits_from(a, b, d) >> (u + 1))) & ((1u << e) - 1u);
		mem += (u64)((1u << e) - (1u << k)) + payload + 1ull;
		++ntok;
		k = e >= 2 ? e - 2 : 0;
		d += L + 1;
	} while (d < 64);
	mem += msum;
	return (u32)(d - 64) | ((u32)k << 6);
}

// ---------------------------------------------------------------------------------------------- scan

__global__ void __launch_bounds__(WS) dec_scan_kernel(const u32 *__restrict__ stream, u64 end_bits,
                                                       const u32 *__restrict__ toklut, u32 *E, ulonglong2 *P, u32 *TK,
                                                       u32 *winX, ulonglong2 *winPT, u32 *winTT)
{
	__shared__ u32 lut[1 << LUT_BITS];
	__shared__ u32 sx[WS];
	__shared__ u32 wmap[WS / 32];
	__shared__ u64 ws[32];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const u64 gs = (u64)blockIdx.x * WS + tid;
	u64 a, b;
	load_slice(stream, end_bits, gs, a, b);
	const int avail = clamp_avail(end_bits, gs << 6);
	for (int i = tid; i < (1 << LUT_BITS); i += WS)
		lut[i] = __ldg(toklut + i);
	__syncthreads();
	// exits for a first token at offset 0 / 1 with order 0
	const u32 xs0 = slice_exit_lut(lut, a, b, avail, 0u), xs1 = slice_exit_lut(lut, a, b, avail, 1u);
	// At order 0 every token has an even length, so a chain keeps the parity of its offsets.  Chain c enters the
	// window at offset c; the parity class it enters every later slice with comes from a scan of 2-state maps.
	u32 mymap = 0;
	if (xs0 != PDEAD && (xs0 >> 6) == 0)
		mymap |= xs0 & 1u;
	if (xs1 != PDEAD && (xs1 >> 6) == 0)
		mymap |= (xs1 & 1u) << 1;
	u32 inc = mymap; // bit b = class behind this slice when it is entered with class b
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const u32 t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d)
			inc = ((inc >> (t & 1u)) & 1u) | (((inc >> ((t >> 1) & 1u)) & 1u) << 1);
	}
	if (lane == 31)
		wmap[wid] = inc;
	__syncthreads();
	u32 pre = 2u; // identity: class 0 -> 0, class 1 -> 1
	for (int i = 0; i < wid; ++i) {
		const u32 m = wmap[i];
		pre = ((m >> (pre & 1u)) & 1u) | (((m >> ((pre >> 1) & 1u)) & 1u) << 1);
	}
	const u32 excl = __shfl_up_sync(0xffffffffu, inc, 1);
	if (lane > 0)
		pre = ((excl >> (pre & 1u)) & 1u) | (((excl >> ((pre >> 1) & 1u)) & 1u) << 1);
	// predicted exits: as if the chains entered at offset 0 / 1 of their class
	sx[tid] = ((pre & 1u) ? xs1 : xs0) | ((((pre >> 1) & 1u) ? xs1 : xs0) << 16);
	__syncthreads();
	// a canonical chain that dies (garbage in front of a chunk, EOF) starts again at its seed offset in the next
	// slice, so that chunks further on in the window still find a chain to join; the slice it dies in is charged
	// DEATH members, which ends every search for a chain that was following it
	u32 e0 = 0u, e1 = 1u;
	if (tid > 0) {
		const u32 v = sx[tid - 1];
		e0 = v & 0xffffu;
		e1 = v >> 16;
		if (e0 == PDEAD)
			e0 = 0u;
		if (e1 == PDEAD)
			e1 = 1u;
	}
	// exits of the predicted entries, with the members / tokens they consume (kept up to date while the entries change)
	u64 m0, m1;
	u32 t0, t1;
	u32 x0 = slice_walk_lut(lut, a, b, avail, e0, m0, t0), x1 = x0;
	if (e1 == e0) {
		m1 = m0;
		t1 = t0;
	} else {
		x1 = slice_walk_lut(lut, a, b, avail, e1, m1, t1);
	}
	// fixed point: entry <- the predecessor's exact exit.  A wrong prediction is repaired where the chains join
	// again, a few slices further on, so the repairs are local: each warp iterates on its own 32 slices with
	// shuffles (no block barrier) and only the warp-boundary states travel through shared memory.  Inside the loops
	// only the exits are recomputed; the member / token counts of the slices whose entry changed follow afterwards.
	__shared__ u32 sb[WS / 32];
	bool dirty0 = false, dirty1 = false;
	u32 bound = tid > 0 ? sx[tid - 1] : 0u; // what lane 0 of a warp takes as its predecessor's exits
	__syncthreads();
	for (;;) {
		for (;;) { // warp-local fixed point
			u32 v = __shfl_up_sync(0xffffffffu, x0 | (x1 << 16), 1);
			if (lane == 0)
				v = bound;
			int changed = 0;
			if (tid > 0) {
				u32 n0 = v & 0xffffu, n1 = v >> 16;
				if (n0 == PDEAD)
					n0 = 0u;
				if (n1 == PDEAD)
					n1 = 1u;
				if (n0 != e0) {
					e0 = n0;
					x0 = slice_exit_lut(lut, a, b, avail, e0);
					dirty0 = true;
					changed = 1;
				}
				if (n1 != e1) {
					e1 = n1;
					x1 = e1 == e0 ? x0 : slice_exit_lut(lut, a, b, avail, e1);
					dirty1 = true;
					changed = 1;
				}
			}
			if (!__any_sync(0xffffffffu, changed))
				break;
		}
		if (lane == 31)
			sb[wid] = x0 | (x1 << 16);
		__syncthreads();
		int moved = 0;
		if (lane == 0 && wid > 0) {
			const u32 nb = sb[wid - 1];
			moved = nb != bound;
			bound = nb;
		}
		if (!__syncthreads_or(moved))
			break;
	}
	if (dirty0)
		slice_walk_lut(lut, a, b, avail, e0, m0, t0);
	if (dirty1) {
		if (e1 == e0) {
			m1 = m0;
			t1 = t0;
		} else {
			slice_walk_lut(lut, a, b, avail, e1, m1, t1);
		}
	}
	if (x0 == PDEAD)
		m0 += DEATH;
	if (x1 == PDEAD)
		m1 += DEATH;
	u64 tot0, tot1, tott;
	const u64 p0 = block_exscan_u64(m0, ws, &tot0);
	const u64 p1 = block_exscan_u64(m1, ws, &tot1);
	const u64 pt = block_exscan_u64((u64)t0 | ((u64)t1 << 32), ws, &tott);
	E[gs] = e0 | (e1 << 16);
	P[gs] = make_ulonglong2(p0, p1);
	TK[gs] = (u32)(pt & 0xffffu) | ((u32)(pt >> 32) << 16);
	if (tid == WS - 1)
		winX[blockIdx.x] = x0 | (x1 << 16);
	if (tid == 0) {
		winPT[blockIdx.x] = make_ulonglong2(tot0, tot1);
		winTT[blockIdx.x] = (u32)(tott & 0xffffu) | ((u32)(tott >> 32) << 16);
	}
}

// The same tables, computed the other way round: ONE thread per (window, chain) walks its 512 slices serially from
// the seed, so every token is decoded once per chain (the kernel above decodes it ~4.5 times: seeds, evaluation,
// repairs, at 13 of 32 lanes).  A stream of a few MB has thousands of windows, which is all the parallelism a
// latency-bound walk needs: the kernel takes ~512 x the latency of one slice (1.8 ms) whatever the stream size, uses a
// third of the instructions and leaves the machine's issue slots to the kernels of other frames.  It is the scan of
// choice when several frames are in flight (8K: +19 % frames per second with 8 contexts); alone on the GPU the parallel
// kernel above is faster (1.1 ms).  Used when the caller announced >= DEC_SERIAL_MIN_IN_FLIGHT busy contexts
// (dwt_ctx_set_in_flight; dwt_pool does) and the stream has >= DEC_SERIAL_MIN_WINDOWS windows.
constexpr u32 DEC_SERIAL_MIN_WINDOWS = 2048;
constexpr int DEC_SERIAL_MIN_IN_FLIGHT = 4;

// chain q of window w walked serially from the state (d0, k0) at its first slice; fills the chain's per-slice tables and
// the window's exit state / totals.  One flat loop over the token steps of the whole window: the lanes of a warp
// (different windows) do not wait for each other at slice boundaries, a lane that crosses one closes the slice and
// opens the next on the side.
__device__ __forceinline__ void walk_window_chain(const u32 *__restrict__ stream, u64 end_bits, const u32 *lut, u32 w, u32 q, int d0,
                                                  int k0, u32 *E, ulonglong2 *P, u32 *TK, u32 *winX, ulonglong2 *winPT, u32 *winTT)
{
	unsigned short *E16 = reinterpret_cast<unsigned short *>(E), *TK16 = reinterpret_cast<unsigned short *>(TK);
	u64 *P64 = reinterpret_cast<u64 *>(P);
	const u64 gs0 = (u64)w * WS;
	auto word = [&](u64 slice) { return (slice << 6) < end_bits + 128 ? __ldg((const u64 *)stream + slice) : 0ull; };
	u64 a = word(gs0), b = word(gs0 + 1), c = word(gs0 + 2), dnext = word(gs0 + 3); // two slices of look-ahead
	int i = 0, d = d0, k = k0;
	int avail = clamp_avail(end_bits, gs0 << 6);
	bool lut_ok = avail >= 64 + LUT_BITS + 4;
	u32 x = (u32)d0 | ((u32)k0 << 6);
	u64 pm = 0, m = 0;  // members before the slice / inside it so far
	u32 ptok = 0, n = 0;
	E16[2 * gs0 + q] = (unsigned short)x;
	P64[2 * gs0 + q] = 0;
	TK16[2 * gs0 + q] = 0;
	for (;;) {
		bool dead = false;
		const u32 bits = window32(a, b, d);
		bool stepped = false;
		if (k == 0 && lut_ok) {
			const u32 t = lut[bits & ((1u << LUT_BITS) - 1u)];
			if ((t & 15u) && d + (int)((t >> 4) & 15u) < 64) { // every token of the entry starts in this slice
				d += (int)(t & 15u);
				n += (t >> 8) & 7u;
				m += t >> 11;
				stepped = true;
			}
		}
		if (!stepped) {
			const int u = bits ? __ffs((int)bits) - 1 : 32;
			const int e = k + u;
			const int L = u + 1 + e;
			if (e > 31 || d + L > avail) {
				dead = true;
			} else {
				// the payload lies inside the 32-bit window for all but the long runs: no 64-bit shifts then
				const u32 payload = (L <= 32 ? bits >> (u + 1) : (u32)(bits_from(a, b, d) >> (u + 1))) & ((1u << e) - 1u);
				m += (u64)((1u << e) - (1u << k)) + payload + 1ull;
				++n;
				k = e >= 2 ? e - 2 : 0;
				d += L + 1;
			}
		}
		if (dead || d >= 64) { // the slice is finished
			x = dead ? PDEAD : ((u32)(d - 64) | ((u32)k << 6));
			pm += m + (dead ? DEATH : 0ull); // see dec_scan_kernel: a dead chain restarts at its seed offset
			ptok += n;
			m = 0;
			n = 0;
			if (++i == WS)
				break;
			if (dead) {
				d = (int)q;
				k = 0;
			} else {
				d -= 64;
			}
			const u64 gs = gs0 + i;
			E16[2 * gs + q] = (unsigned short)((u32)d | ((u32)k << 6));
			P64[2 * gs + q] = pm;
			TK16[2 * gs + q] = (unsigned short)ptok;
			a = b;
			b = c;
			c = dnext;
			dnext = word(gs + 3);
			avail = clamp_avail(end_bits, gs << 6);
			lut_ok = avail >= 64 + LUT_BITS + 4;
		}
	}
	reinterpret_cast<unsigned short *>(winX)[2 * w + q] = (unsigned short)x;
	reinterpret_cast<u64 *>(winPT)[2 * w + q] = pm;
	reinterpret_cast<unsigned short *>(winTT)[2 * w + q] = (unsigned short)ptok;
}

// ONE thread per (window, chain) walks its 128 slices serially from the seed (bit q of the window, order 0), so every
// token is decoded once per chain.  A stream of a few MB has tens of thousands of windows, which is all the parallelism
// the latency-bound walks need.
__global__ void __launch_bounds__(128) dec_scan_serial_kernel(const u32 *__restrict__ stream, u64 end_bits,
                                                               const u32 *__restrict__ toklut, u32 nwin, u32 *E, ulonglong2 *P,
                                                               u32 *TK, u32 *winX, ulonglong2 *winPT, u32 *winTT)
{
	__shared__ u32 lut[1 << LUT_BITS];
	for (int i = threadIdx.x; i < (1 << LUT_BITS); i += 128)
		lut[i] = __ldg(toklut + i);
	__syncthreads();
	const u32 t = blockIdx.x * 128 + threadIdx.x;
	const u32 w = t >> 1, q = t & 1u;
	if (w >= nwin)
		return;
	walk_window_chain(stream, end_bits, lut, w, q, (int)q, 0, E, P, TK, winX, winPT, winTT);
}

// Lineage pass.  A chain seeded at a window start needs a while to fall onto the stream's true token grid: a few
// tokens where the Rice order is 0 (dense planes), but on the order of a hundred slices where runs are long (the top
// planes of the big levels: the order there is ~10 and a chain at order 0 reads payload bits as short tokens until a
// long zero string lifts it).  With 128-slice windows such regions would never have a canonical chain old enough to be
// the true one.  So class 1 of window w is replaced by the CONTINUATION of class 1 of window w - 1 wherever that
// continuation does not fall onto one of w's chains within a few slices; every pass makes the oldest chain of such a
// region one window older (pass p reads the exits of pass p - 1: no ordering between the threads of a pass).  Dense
// regions are left alone: there the continuation joins at once.
constexpr int EXTEND_JOIN_SLICES = 6;

// part 1, one thread per window: does the chain that leaves window w - 1 on class 1 fall onto one of w's chains within a
// few slices?  If not, w goes on the list of windows whose class-1 chain is to be replaced by that continuation.
__global__ void __launch_bounds__(128) dec_extend_find_kernel(const u32 *__restrict__ stream, u64 end_bits, u32 nwin,
                                                               const u32 *__restrict__ E, const u32 *__restrict__ winX_in,
                                                               u32 *winX_out, const unsigned char *changed_in,
                                                               unsigned char *changed_out, u32 *list_count, uint2 *list)
{
	const u32 w = blockIdx.x * 128 + threadIdx.x;
	if (w >= nwin)
		return;
	winX_out[w] = winX_in[w];
	changed_out[w] = 0;
	if (w < 1 || (changed_in && !changed_in[w - 1]))
		return;
	const u32 st = winX_in[w - 1] >> 16; // where the class-1 chain of the window in front ends
	if (st == PDEAD)
		return;
	const u64 gs0 = (u64)w * WS;
	u32 s = st;
	for (int i = 0; i < EXTEND_JOIN_SLICES; ++i) { // no table: a handful of tokens
		const u32 e = E[gs0 + i];
		if (s == (e & 0xffffu) || s == (e >> 16))
			return;
		u64 a, b;
		load_slice(stream, end_bits, gs0 + i, a, b);
		s = slice_exit(a, b, clamp_avail(end_bits, (gs0 + i) << 6), s);
		if (s == PDEAD)
			return; // a chain that dies here is not worth keeping
	}
	list[atomicAdd(list_count, 1u)] = make_uint2(w, st);
}

// part 2, one WARP per listed window.  A lone thread that walks a window with full bookkeeping needs ~0.18 ms; here lane 0
// only follows the token POSITIONS from slice to slice (a few instructions per token, ~17 us per window) and leaves every
// slice's entry state in shared memory, then the 32 lanes count what four slices each consume from their entry states, and
// a warp scan turns the counts into the chain's tables.  Same tables as walk_window_chain.  The serial walk is the
// latency of a pass, so what matters is how many windows are walked at the same time: a warp per window keeps up to 64 of
// them on an SM (a CTA per window, the first version, kept 4 and a pass took 85 - 290 us).
constexpr int XW_WARPS = 8;

__global__ void __launch_bounds__(XW_WARPS * 32) dec_extend_walk_kernel(const u32 *__restrict__ stream, u64 end_bits,
                                                                        const u32 *__restrict__ toklut, u32 *E, ulonglong2 *P,
                                                                        u32 *TK, u32 *winX_out, ulonglong2 *winPT, u32 *winTT,
                                                                        unsigned char *changed_out,
                                                                        const u32 *__restrict__ list_count,
                                                                        const uint2 *__restrict__ list)
{
	static_assert(WS == 128, "four slices per lane");
	__shared__ u32 lut[1 << LUT_BITS];
	__shared__ u64 sw_all[XW_WARPS][WS + 1];
	__shared__ unsigned short sentry_all[XW_WARPS][WS];
	const u32 count = *list_count;
	if (blockIdx.x * XW_WARPS >= count)
		return;
	for (int i = threadIdx.x; i < (1 << LUT_BITS); i += XW_WARPS * 32)
		lut[i] = __ldg(toklut + i);
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	u64 *sw = sw_all[wid];
	unsigned short *sentry = sentry_all[wid];
	unsigned short *E16 = reinterpret_cast<unsigned short *>(E), *TK16 = reinterpret_cast<unsigned short *>(TK);
	u64 *P64 = reinterpret_cast<u64 *>(P);
	for (u32 item = blockIdx.x * XW_WARPS + wid; item < count; item += gridDim.x * XW_WARPS) {
		const uint2 it = list[item];
		const u32 w = it.x;
		const u64 gs0 = (u64)w * WS;
		__syncwarp(); // the previous item's shared arrays are no longer read
		for (int i = lane; i < WS + 1; i += 32)
			sw[i] = ((gs0 + i) << 6) < end_bits + 128 ? __ldg((const u64 *)stream + gs0 + i) : 0ull;
		__syncwarp();
		if (lane == 0) { // the serial part: positions only
			u32 state = it.y;
			if (((gs0 + WS + 4) << 6) < end_bits) {
				// The window and everything a token of it can touch lie inside the stream.  The walk is ONE dependent chain
				// (measured: ~700 tokens and 125 000 cycles per listed window, 180 cycles per token through slice_exit_lut),
				// so it is written for the length of that chain: the slice's words in registers, a token's length straight
				// from the unary prefix (u zeros, the one, e = k + u payload bits, the sign: 2u + k + 2 bits), no end-of-stream
				// bookkeeping, no table (the listed windows are the ones with orders well above 0).
				const u32 *w32 = reinterpret_cast<const u32 *>(sw);
				int d = (int)(state & 63u), k = (int)(state >> 6);
				u32 w0 = w32[0], w1 = w32[1];
				for (int t = 0; t < WS; ++t) {
					const u32 w2 = w32[2 * t + 2], w3 = w32[2 * t + 3]; // the next slice's words (independent of the chain)
					sentry[t] = (unsigned short)((u32)d | ((u32)k << 6));
					do {
						const u32 bits = __funnelshift_r(d < 32 ? w0 : w1, d < 32 ? w1 : w2, d);
						const int u = __ffs((int)bits) - 1; // -1: no one within 32 bits
						const int e = k + u;
						if (bits == 0u || e > 31) { // dead: restart at class 1's seed offset with order 0 in the next slice
							d = 64 + 1;
							k = 0;
							break;
						}
						d += 2 * u + k + 2;
						k = max(e - 2, 0);
					} while (d < 64);
					d -= 64;
					w0 = w2;
					w1 = w3;
				}
			} else {
				for (int t = 0; t < WS; ++t) {
					sentry[t] = (unsigned short)state;
					const u64 b2 = ((gs0 + t) << 6) < end_bits + 64 ? sw[t + 1] : 0ull; // load_slice's rule for the second word
					const u32 x = slice_exit_lut(lut, sw[t], b2, clamp_avail(end_bits, (gs0 + t) << 6), state);
					state = x == PDEAD ? 1u : x; // a dead chain restarts at class 1's seed offset with order 0 (dec_scan_kernel)
				}
			}
		}
		__syncwarp();
		// lane l counts slices 4 l .. 4 l + 3 from their entry states
		u64 mem[4], msum = 0;
		u32 tok[4], tsum = 0, xlast = 0;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const int t = 4 * lane + q;
			const u64 gs = gs0 + t;
			const u64 b2 = (gs << 6) < end_bits + 64 ? sw[t + 1] : 0ull;
			xlast = slice_walk_lut(lut, sw[t], b2, clamp_avail(end_bits, gs << 6), (u32)sentry[t], mem[q], tok[q]);
			if (xlast == PDEAD)
				mem[q] += DEATH;
			msum += mem[q];
			tsum += tok[q];
		}
		u64 minc = msum;
		u32 tinc = tsum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const u64 tm = shfl_up_u64(minc, d);
			const u32 tt = __shfl_up_sync(0xffffffffu, tinc, d);
			if (lane >= d) {
				minc += tm;
				tinc += tt;
			}
		}
		u64 pm = minc - msum;
		u32 pt = tinc - tsum;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const u64 gs = gs0 + 4 * lane + q;
			E16[2 * gs + 1] = sentry[4 * lane + q];
			P64[2 * gs + 1] = pm;
			TK16[2 * gs + 1] = (unsigned short)pt;
			pm += mem[q];
			pt += tok[q];
		}
		if (lane == 31) {
			reinterpret_cast<unsigned short *>(winX_out)[2 * w + 1] = (unsigned short)xlast;
			reinterpret_cast<u64 *>(winPT)[2 * w + 1] = minc;
			reinterpret_cast<unsigned short *>(winTT)[2 * w + 1] = (unsigned short)tinc;
			changed_out[w] = 1;
		}
	}
}

// ---------------------------------------------------------------------------------------------- link

__global__ void __launch_bounds__(128) dec_link_kernel(const u32 *__restrict__ stream, u64 end_bits, u32 nwin,
                                                        const u32 *__restrict__ E, const ulonglong2 *__restrict__ P,
                                                        const u32 *__restrict__ TK, const u32 *__restrict__ winX,
                                                        const ulonglong2 *__restrict__ winPT,
                                                        const u32 *__restrict__ winTT, DecLink *link)
{
	const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
	const u32 w = t >> 1, q = t & 1u;
	if (w >= nwin)
		return;
	DecLink L;
	L.mem = 0;
	L.tok = 0;
	L.stray_mem = 0;
	L.stray_tok = 0;
	L.exit_state = (unsigned short)PDEAD;
	L.m = WS;
	L.qn = 3;
	L.qm = 0;
	if (w > 0) {
		u32 st = (winX[w - 1] >> (16 * q)) & 0xffffu;
		u64 smem = 0;
		u32 stok = 0;
		int i = 0, qm = -1;
		if (st != PDEAD) {
			// true chains join within a few slices, but where both canonical chains have fallen onto one bit parity
			// (dense planes without long runs) a chain of the other parity can run through the whole window: such
			// a window is walked to its end here, in parallel, rather than by the resolver (LINK_CAP = window).
			// A lower cap is honoured by the resolver (LINK_OPEN records are stepped exactly).
			// the loads do not depend on the state: the next slice is fetched while this one is walked
			u64 na, nb;
			u32 ne = E[(u64)w * WS];
			load_slice(stream, end_bits, (u64)w * WS, na, nb);
			for (; i < LINK_CAP; ++i) {
				const u64 gs = (u64)w * WS + i;
				const u32 e = ne;
				const u64 a = na, b = nb;
				if (i + 1 < LINK_CAP) {
					ne = E[gs + 1];
					load_slice(stream, end_bits, gs + 1, na, nb);
				}
				if (st == (e & 0xffffu)) {
					qm = 0;
					break;
				}
				if (st == (e >> 16)) {
					qm = 1;
					break;
				}
				u64 mm;
				u32 tt;
				st = slice_walk(a, b, clamp_avail(end_bits, gs << 6), st, mm, tt);
				smem += mm;
				stok += tt;
				if (st == PDEAD)
					break;
			}
		}
		L.stray_mem = smem > 0xffffffffull ? 0xffffffffu : (u32)smem;
		L.stray_tok = stok;
		if (qm >= 0) {
			const u64 gs = (u64)w * WS + i;
			const ulonglong2 pm = P[gs], pt = winPT[w];
			const u32 tkm = TK[gs], tt = winTT[w];
			L.m = (unsigned short)i;
			L.qm = (unsigned char)qm;
			L.mem = smem + (qm ? pt.y - pm.y : pt.x - pm.x);
			L.tok = stok + (qm ? (tt >> 16) - (tkm >> 16) : (tt & 0xffffu) - (tkm & 0xffffu));
			const u32 x = (winX[w] >> (16 * qm)) & 0xffffu;
			L.qn = x == PDEAD ? 3 : (unsigned char)qm;
		} else {
			L.mem = smem;
			L.tok = stok;
			L.qn = st == PDEAD ? 3 : 2;
			L.exit_state = (unsigned short)(st == PDEAD ? PDEAD : (i >= WS ? st : LINK_OPEN));
		}
	}
	uint4 r0, r1;
	r0.x = (u32)L.mem;
	r0.y = (u32)(L.mem >> 32);
	r0.z = L.tok;
	r0.w = L.stray_mem;
	r1.x = L.stray_tok;
	r1.y = (u32)L.exit_state | ((u32)L.m << 16);
	r1.z = (u32)L.qn | ((u32)L.qm << 8);
	r1.w = 0;
	((uint4 *)link)[2 * (size_t)t] = r0;
	((uint4 *)link)[2 * (size_t)t + 1] = r1;
}

// ---------------------------------------------------------------------------------------------- resolve

// tokens that start in the slice (a, b) from offset d with order k, against the member budget T.
// Returns EV_NONE when the slice is left (d >= 64 then), otherwise the event that ends the pass.
// lut: order-0 token table in shared memory (the resolver runs in one warp, so table and generic steps never diverge)
// 32 stream bits from offset `off` (0 .. 95) of the 128-bit pair (a, b)
__device__ __forceinline__ u32 window_at(u64 a, u64 b, int off)
{
	const u32 w0 = (u32)a, w1 = (u32)(a >> 32), w2 = (u32)b, w3 = (u32)(b >> 32);
	const int i = off >> 5;
	const u32 lo = i == 0 ? w0 : (i == 1 ? w1 : w2);
	const u32 hi = i == 0 ? w1 : (i == 1 ? w2 : w3);
	return __funnelshift_r(lo, hi, off & 31);
}

// The resolver's exact walk is one dependent chain per token, so what counts is the length of that chain: all state is
// 32 bit (a chunk has fewer than 2^31 members, a run is below 2^32), the payload comes from a second 32-bit window instead
// of 64-bit shifts, and the end-of-stream checks live in a separate variant that only runs within three slices of EOF.
template <bool CHECKED>
__device__ __forceinline__ int walk_events_t(const u32 *lut, u64 a, u64 b, int avail, u32 T, int &d, int &k, u32 &cum, u32 &ones,
                                             u32 &ev_pending)
{
	const bool lut_ok = !CHECKED || avail >= 64 + LUT_BITS + 4;
	for (;;) {
		if (cum >= T)
			return EV_COVERED; // every member has its symbol; the pass ends in front of the token at d
		if (d >= 64)
			return EV_NONE;
		const u32 bits = window32(a, b, d);
		if (k == 0 && lut_ok) { // several short tokens at once, as long as the pass cannot end inside them
			const u32 t = lut[bits & ((1u << LUT_BITS) - 1u)];
			const u32 mem = t >> 11;
			if ((t & 15u) && d + (int)((t >> 4) & 15u) < 64 && mem < T - cum) {
				d += (int)(t & 15u);
				ones += (t >> 8) & 7u;
				cum += mem;
				continue;
			}
		}
		const int u = bits ? __ffs((int)bits) - 1 : 32;
		const int e = k + u;
		const int L = u + 1 + e;
		if (e > 31 || (CHECKED && d + L > avail))
			return EV_STOP; // the token cannot be read completely (vli.h:88-95): decoding ends
		// the payload lies inside the first window for all but the long runs (u + 1 + e <= 32)
		const u32 payload = (L <= 32 ? bits >> (u + 1) : window_at(a, b, d + u + 1)) & ((1u << e) - 1u);
		const u32 n = (1u << e) - (1u << k) + payload;
		const int kn = e >= 2 ? e - 2 : 0;
		const u32 room = T - cum;
		if (n >= room) { // the run reaches past this chunk's members (rle.h:74-76)
			ev_pending = n - room + 1;
			d += L;
			k = kn;
			return EV_PENDING;
		}
		++ones;
		if (CHECKED && d + L + 1 > avail)
			return EV_STOP; // sign bit beyond EOF: the magnitude bit stays (decode.c:80-86)
		cum += n + 1;
		d += L + 1;
		k = kn;
	}
}

// tokens that start in the slice (a, b) from offset d with order k, against the member budget T (cum < T on entry).
// Returns EV_NONE when the slice is left (d >= 64 then), otherwise the event that ends the pass.
// lut: order-0 token table in shared memory (the resolver runs in one warp, so table and generic steps never diverge)
__device__ __forceinline__ int walk_events(const u32 *lut, u64 a, u64 b, int avail, u64 T64, int &d, int &k, u64 &cum64, u32 &ones,
                                           u32 &ev_pending)
{
	if (cum64 >= T64)
		return EV_COVERED;
	const u32 T = (u32)T64; // T <= 2^31 (a level has fewer than 2^31 coefficients), cum < T
	u32 cum = (u32)cum64;
	const int r = avail >= 192 ? walk_events_t<false>(lut, a, b, avail, T, d, k, cum, ones, ev_pending)
	                           : walk_events_t<true>(lut, a, b, avail, T, d, k, cum, ones, ev_pending);
	cum64 = cum;
	return r;
}

__device__ __forceinline__ u32 map_compose(u32 first, u32 second) // class maps: 2 bits per entry class 0 / 1
{
	const u32 a0 = first & 3u, a1 = (first >> 2) & 3u;
	const u32 r0 = a0 >= 2u ? a0 : (second >> (2 * a0)) & 3u;
	const u32 r1 = a1 >= 2u ? a1 : (second >> (2 * a1)) & 3u;
	return r0 | (r1 << 2);
}

__device__ __forceinline__ void load_link(const DecLink *src, DecLink &L) // two 16-byte loads, unpacked in registers
{
	const uint4 r0 = __ldg((const uint4 *)src), r1 = __ldg((const uint4 *)src + 1);
	L.mem = (u64)r0.x | ((u64)r0.y << 32);
	L.tok = r0.z;
	L.stray_mem = r0.w;
	L.stray_tok = r1.x;
	L.exit_state = (unsigned short)(r1.y & 0xffffu);
	L.m = (unsigned short)(r1.y >> 16);
	L.qn = (unsigned char)(r1.z & 0xffu);
	L.qm = (unsigned char)((r1.z >> 8) & 0xffu);
}

__device__ __forceinline__ void dead_link(DecLink &L) // what a window behind the stream looks like
{
	L.mem = 0;
	L.tok = 0;
	L.stray_mem = 0;
	L.stray_tok = 0;
	L.exit_state = (unsigned short)PDEAD;
	L.m = WS;
	L.qn = 3;
	L.qm = 0;
}

__device__ __forceinline__ u64 shfl_u64(u64 v, int src)
{
	const u32 lo = __shfl_sync(0xffffffffu, (u32)v, src), hi = __shfl_sync(0xffffffffu, (u32)(v >> 32), src);
	return (u64)lo | ((u64)hi << 32);
}

__device__ __forceinline__ u32 scan_class_maps(u32 inc, int lane) // inclusive scan of 2-state class maps over the warp
{
#pragma unroll
	for (int dd = 1; dd < 32; dd <<= 1) {
		const u32 t = __shfl_up_sync(0xffffffffu, inc, dd);
		if (lane >= dd)
			inc = map_compose(t, inc);
	}
	return inc;
}

// ---------------------------------------------------------------------------------------------- super records

// One warp per super-window of DWT_DEC_SUPER windows: what the 32 windows consume when the chain enters the first of
// them with class 0 / class 1 -- the link records composed, so that the resolver crosses 32 windows with one record.
__global__ void __launch_bounds__(128) dec_super_kernel(u32 nwin, u32 nsuper, const DecLink *__restrict__ link, DecSuper *super)
{
	static_assert(DWT_DEC_SUPER == 32, "one lane per window of a super-window");
	const u32 FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const u32 s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (s >= nsuper)
		return;
	const u32 w = s * DWT_DEC_SUPER + lane;
	const bool valid = w < nwin;
	DecLink L0, L1;
	if (valid) {
		load_link(link + 2 * (u64)w, L0);
		load_link(link + 2 * (u64)w + 1, L1);
	} else {
		dead_link(L0);
		dead_link(L1);
	}
	const u32 inc = scan_class_maps((u32)L0.qn | ((u32)L1.qn << 2), lane);
	const u32 excl = __shfl_up_sync(FULL, inc, 1);
	const u32 last = __shfl_sync(FULL, inc, 31);
#pragma unroll
	for (int q = 0; q < 2; ++q) {
		const u32 cls = lane == 0 ? (u32)q : (excl >> (2 * q)) & 3u; // class the chain enters my window with
		const bool canon = cls < 2u && valid;
		const DecLink &my = cls == 1u ? L1 : L0;
		u64 mem = canon ? my.mem : 0ull;
		const u32 tok = __reduce_add_sync(FULL, canon ? my.tok : 0u);
#pragma unroll
		for (int dd = 16; dd >= 1; dd >>= 1)
			mem += shfl_u64(mem, lane ^ dd);
		const bool clean = __all_sync(FULL, canon);
		const u32 ex = __shfl_sync(FULL, (u32)my.exit_state, 31);
		if (lane == 0) {
			uint4 r;
			r.x = (u32)mem;
			r.y = (u32)(mem >> 32);
			r.z = tok;
			r.w = ex | (((last >> (2 * q)) & 3u) << 16) | ((clean ? 1u : 0u) << 24);
			((uint4 *)super)[2 * (u64)s + q] = r;
		}
	}
}

__device__ __forceinline__ void load_super(const DecSuper *src, DecSuper &S)
{
	const uint4 r = __ldg((const uint4 *)src);
	S.mem = (u64)r.x | ((u64)r.y << 32);
	S.tok = r.z;
	S.exit_state = (unsigned short)(r.w & 0xffffu);
	S.qn = (unsigned char)((r.w >> 16) & 3u);
	S.clean = (unsigned char)((r.w >> 24) & 1u);
}

// ---------------------------------------------------------------------------------------------- resolve

// make EXTRA=-DDWT_RESOLVE_PROFILE: cycles per part of the resolver (development aid; not in the shipped library)
#ifdef DWT_RESOLVE_PROFILE
#define RP_BEGIN() const long long rp_t0 = clock64()
#define RP_END(slot) rp_cyc[slot] += clock64() - rp_t0
#define RP_STAT(stmt) stmt
#define RP_HIST(t) hist[t]
#define RP_REASON(t) reason[t]
#else
#define RP_BEGIN()
#define RP_END(slot)
#define RP_STAT(stmt)
#define RP_HIST(t) 0u
#define RP_REASON(t) 0u
#endif

__global__ void __launch_bounds__(32) dec_resolve_kernel(const __grid_constant__ Geom G, int nchunks,
                                                          const __grid_constant__ DecBuffers B)
{
	static_assert(WS % 32 == 0 && WS >= 32 && WS <= 256, "the end search loads WS / 32 slices per lane");
	constexpr int NPL = WS / 32;
#ifdef DWT_RESOLVE_PROFILE
	long long rp_cyc[5] = {0, 0, 0, 0, 0};
	const long long rp_start = clock64();
#endif
	__shared__ u32 sigcount[48];
	__shared__ int missing[48];
	__shared__ u32 lut[1 << LUT_BITS];
	__shared__ unsigned char s_chan[DWT_MAX_CHUNKS], s_level[DWT_MAX_CHUNKS];
	const int lane = threadIdx.x;
	const u32 FULL = 0xffffffffu;
	DecState *st = B.state;
	const Sched *S = B.sched;
	const u32 *__restrict__ stream = B.stream;
	const u64 end_bits = B.end_bits;
	const u32 nwin = B.nwin, nsuper = B.nsuper;
	for (int i = lane; i < 48; i += 32) {
		sigcount[i] = 0;
		missing[i] = st->missing[i];
	}
	for (int i = lane; i < (1 << LUT_BITS); i += 32)
		lut[i] = __ldg(B.toklut + i);
	for (int i = lane; i < nchunks; i += 32) {
		s_chan[i] = (unsigned char)S->chan[i];
		s_level[i] = (unsigned char)S->level[i];
	}
	__syncwarp();
	u64 bitpos = st->bitpos;
	int order = st->order;
	u32 pending = 0;
	int level = -1;
	bool stopped = false;
	u64 rank_base = 0;
	u32 nseg = 0, nbulk = 0, slow_entries = 0, exact_steps = 0, n_super = 0, n_window = 0, n_search = 0;
#ifdef DWT_RESOLVE_PROFILE
	u32 hist[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, reason[4] = {0, 0, 0, 0};
#endif
	int tripped = 0;
	// every pass of the loops below moves on by at least one window or ends the chunk; the guard only exists so that a
	// bug can never hang the device
	const u32 guard_max = 8u * nwin + 4096u;

	for (int j = 0; j < nchunks && !stopped; ++j) {
		const int c = s_chan[j], l = s_level[j];
		if (level < l)
			level = l; // decode.c:203,219-220,236-237: the chunk is started
		const u32 nsig = sigcount[c * 16 + l];
		const u64 R = (u64)G.num[l] - nsig;
		const u32 nref = nsig;
		const u64 my_rank_base = rank_base;
		rank_base += ((u64)G.num[l] + 63) & ~63ull;
		u64 r0 = 0;
		bool stop = false;
		u32 ones = 0;
		// a run carried in from earlier chunks (rle.h:66-77): (pending - 1) zeros, then a one
		if (pending > 0) {
			if ((u64)pending - 1 >= R) {
				pending -= (u32)R;
				r0 = R;
			} else {
				const u64 rk = my_rank_base + pending - 1;
				++ones;
				if (lane == 0)
					atomicOr(B.ones_rank + (rk >> 5), 1u << (rk & 31));
				if (bitpos < end_bits) {
					const u32 wd = __ldg(stream + (bitpos >> 5));
					if (lane == 0 && ((wd >> (bitpos & 31)) & 1u))
						atomicOr(B.sign_rank + (rk >> 5), 1u << (rk & 31));
					bitpos += 1;
				} else {
					stop = true; // the sign bit hits EOF: the magnitude bit stays (decode.c:80-86)
				}
				r0 = (u64)pending;
				pending = 0;
			}
		}
		int event = EV_NONE;
		u64 f_pos = bitpos;
		int f_k = order;
		u32 f_pending = pending;
		const u64 T = R > r0 ? R - r0 : 0;
		if (stop) {
			event = EV_STOP;
		} else if (pending == 0 && r0 < R) {
			// ------------------------------------------------ the chunk's own tokens: find where the pass ends
			u64 cum = 0;
			u64 gs = bitpos >> 6;
			int d = (int)(bitpos & 63), k = order;
			int mode = 0;        // 0 stray (exact steps), 1 linked (whole windows / super-windows)
			u32 w = 0, q = 0;    // linked mode: next window, class of the chain behind window w-1
			u32 carry_state = PDEAD;
			bool force_window = false; // linked mode: the super-window at w has to be looked at window by window
			// end search on class sq of window sw from slice sm (cum = members before slice sm)
			bool search = false;
			u32 sw = 0, sq = 0;
			int sm = 0;
			// link records fetched together with a window's slices, for the window round that usually follows
			u32 pf_w = 0xffffffffu, pfX = 0;
			DecLink pfL0, pfL1;
			dead_link(pfL0);
			dead_link(pfL1);
			u32 guard = 0;
			bool long_walk = false; // the walk has already crossed a window without joining: no point in careful first steps
			++slow_entries;
			RP_STAT(++reason[0]);
			while (event == EV_NONE) {
				if (++guard > guard_max) {
					tripped = 1;
					event = EV_STOP;
					break;
				}
				if (mode == 0) {
					// ---- exact steps from slice gs until the chain joins a canonical chain, the window ends or the pass ends
					RP_BEGIN();
					const u32 cw = (u32)(gs / WS);
					const int i0 = (int)(gs - (u64)cw * WS);
					if (cw >= nwin) {
						event = EV_STOP; // behind the scanned stream: nothing left to read
						break;
					}
					// everything the visit can need is requested at once: the window's totals, the link records of the windows
					// behind it, and (in the batch loop) slices with their tables -- one round trip instead of three
					const ulonglong2 ptw = B.winPT[cw];
					const u32 ttw = B.winTT[cw], xw = B.winX[cw];
					{
						const u32 wl = cw + 1 + (u32)lane;
						if (wl < nwin) {
							load_link(B.link + 2 * (u64)wl, pfL0);
							load_link(B.link + 2 * (u64)wl + 1, pfL1);
							pfX = B.winX[wl - 1];
						} else {
							dead_link(pfL0);
							dead_link(pfL1);
							pfX = 0xffffffffu;
						}
						pf_w = cw + 1;
					}
					const u32 seg_state = (u32)d | ((u32)k << 6);
					const u64 seg_cum0 = cum;
					RP_STAT(const u32 steps0 = exact_steps;)
					int i = i0, m = WS, qm = 0;
					u64 pm = 0;   // members of class qm before the join slice
					u32 tkm = 0;  // tokens of both classes before the join slice
					while (i < WS && event == EV_NONE && m == WS) {
						const int nb = min(32, WS - i);
						const u64 mgs = (u64)cw * WS + i + lane;
						u64 ma = 0, mb = 0;
						u32 me = 0, mtk = 0;
						ulonglong2 mp = make_ulonglong2(0ull, 0ull);
						if (lane < nb) {
							load_slice(stream, end_bits, mgs, ma, mb);
							me = B.E[mgs];
							mp = B.P[mgs];
							mtk = B.TK[mgs];
						}
						// Most visits join within a slice or two: those slices are stepped with full bookkeeping right away.
						// A walk that is still going after EXACT_FIRST slices (a region where the canonical chains are slow to
						// fall onto the token grid: long runs, high Rice orders) changes gear: the serial part only follows the
						// token POSITIONS (offset and order from slice to slice: a few instructions per token) until the chain
						// joins or a token cannot be read; what the slices consume is then counted by all lanes at once, one
						// slice each, from the entry states the walk left behind; a scan finds the slice in which the members
						// run out, and only that slice is stepped with full bookkeeping.
						const int avail_l = clamp_avail(end_bits, mgs << 6);
						int t0 = 0;
						if (i == i0 && !long_walk) {
							for (; t0 < nb && t0 < EXACT_FIRST; ++t0) {
								const u32 e = __shfl_sync(FULL, me, t0);
								const u32 state = (u32)d | ((u32)k << 6);
								if (state == (e & 0xffffu) || state == (e >> 16)) {
									m = i + t0;
									qm = state == (e & 0xffffu) ? 0 : 1;
									pm = shfl_u64(qm ? mp.y : mp.x, t0);
									tkm = __shfl_sync(FULL, mtk, t0);
									break;
								}
								const u64 a = shfl_u64(ma, t0), b = shfl_u64(mb, t0);
								++exact_steps;
								const u64 sgs = (u64)cw * WS + i + t0;
								event = walk_events(lut, a, b, __shfl_sync(FULL, avail_l, t0), T, d, k, cum, ones, f_pending);
								if (event != EV_NONE) {
									f_pos = (sgs << 6) + (u64)d;
									f_k = k;
									break;
								}
								d -= 64;
							}
							if (event != EV_NONE || m != WS)
								break;
						}
						u32 my_entry = PDEAD;
						int nwalk = 0, stop_t = -1;
						for (int t = t0; t < nb; ++t) {
							const u32 e = __shfl_sync(FULL, me, t);
							const u32 state = (u32)d | ((u32)k << 6);
							if (state == (e & 0xffffu) || state == (e >> 16)) {
								m = i + t;
								qm = state == (e & 0xffffu) ? 0 : 1;
								pm = shfl_u64(qm ? mp.y : mp.x, t);
								tkm = __shfl_sync(FULL, mtk, t);
								break;
							}
							const u64 a = shfl_u64(ma, t), b = shfl_u64(mb, t);
							if (lane == t)
								my_entry = state;
							const u32 x = slice_exit_lut(lut, a, b, __shfl_sync(FULL, avail_l, t), state);
							++nwalk;
							if (x == PDEAD) {
								stop_t = t; // a token of this slice cannot be read completely: the exact step below says why
								break;
							}
							d = (int)(x & 63u);
							k = (int)(x >> 6);
						}
						exact_steps += (u32)nwalk;
						if (nwalk > 0) {
							const bool mine = lane >= t0 && lane < t0 + nwalk;
							u64 lmem = 0;
							u32 ltok = 0;
							if (mine)
								slice_walk_lut(lut, ma, mb, avail_l, my_entry, lmem, ltok);
							u64 incm = lmem;
							u32 inct = ltok;
#pragma unroll
							for (int dd = 1; dd < 32; dd <<= 1) {
								const u64 tm = shfl_u64(incm, max(lane - dd, 0));
								const u32 tt = __shfl_up_sync(FULL, inct, dd);
								if (lane >= dd) {
									incm += tm;
									inct += tt;
								}
							}
							const u32 crossed = __ballot_sync(FULL, mine && cum + incm >= T);
							int xt = crossed ? __ffs((int)crossed) - 1 : -1;
							if (xt < 0)
								xt = stop_t;
							if (xt >= 0) {
								// the pass ends (or the stream does) inside slice i + xt: members and ones in front of it, then
								// the exact step from the slice's entry state
								cum += shfl_u64(incm - lmem, xt);
								ones += __shfl_sync(FULL, inct - ltok, xt);
								const u32 en = __shfl_sync(FULL, my_entry, xt);
								const u64 a = shfl_u64(ma, xt), b = shfl_u64(mb, xt);
								const u64 sgs = (u64)cw * WS + i + xt;
								d = (int)(en & 63u);
								k = (int)(en >> 6);
								event = walk_events(lut, a, b, clamp_avail(end_bits, sgs << 6), T, d, k, cum, ones, f_pending);
								if (event == EV_NONE && cum >= T)
									event = EV_COVERED; // the pass ends exactly with the slice's last token
								if (event == EV_NONE) {
									event = EV_STOP; // cannot happen: the scan saw the members run out or a token fail here
									tripped = 3;
								}
								f_pos = (sgs << 6) + (u64)d;
								f_k = k;
								m = WS; // the pass ended in front of a join the position walk may have seen further on
								qm = 0;
								break;
							}
							cum += shfl_u64(incm, t0 + nwalk - 1);
							ones += __shfl_sync(FULL, inct, t0 + nwalk - 1);
						}
						i += nb;
					}
					if (lane == 0) {
						DecSeg sg;
						sg.w = cw;
						sg.j = (unsigned short)j;
						sg.state = (unsigned short)seg_state;
						sg.i0 = (unsigned short)i0;
						sg.m = (unsigned short)m;
						sg.qm = (u32)qm;
						sg.cum0 = (u32)seg_cum0;
						sg.cum_m = (u32)cum;
						B.seg[nseg] = sg;
					}
					++nseg;
					RP_END(0);
#ifdef DWT_RESOLVE_PROFILE
					{
						const u32 n = exact_steps - steps0;
						const int b = n <= 2 ? (int)n : (n <= 4 ? 3 : (n <= 8 ? 4 : (n <= 16 ? 5 : (n <= 32 ? 6 : (n <= 64 ? 7 : 8)))));
#pragma unroll
						for (int t = 0; t < 9; ++t)
							hist[t] += b == t;
						hist[9] += m < WS;
					}
#endif
					if (event != EV_NONE)
						break;
					if (m == WS) { // the window ended before the chain joined: go on exactly in the next one
						long_walk = true;
						RP_STAT(++reason[3]);
						gs = ((u64)cw + 1) * WS;
						continue;
					}
					// joined class qm at slice m: is the rest of the window enough to end the pass?
					const u64 rest = (qm ? ptw.y : ptw.x) - pm;
					if (cum + rest >= T) {
						search = true;
						sw = cw;
						sq = (u32)qm;
						sm = m;
					} else {
						cum += rest;
						ones += qm ? (ttw >> 16) - (tkm >> 16) : (ttw & 0xffffu) - (tkm & 0xffffu);
						const u32 x = (xw >> (16 * qm)) & 0xffffu;
						if (x == PDEAD) {
							event = EV_STOP; // the chain dies inside this window
							break;
						}
						mode = 1;
						w = cw + 1;
						q = (u32)qm;
						force_window = false;
					}
				} else if ((w & (DWT_DEC_SUPER - 1)) == 0 && !force_window) {
					// ---- whole super-windows, 32 at a time: lane i looks at super-window w / 32 + i
					++n_super;
					RP_BEGIN();
					const u32 sl = w / DWT_DEC_SUPER + (u32)lane;
					DecSuper R0, R1;
					if (sl < nsuper) {
						load_super(B.super + 2 * (u64)sl, R0);
						load_super(B.super + 2 * (u64)sl + 1, R1);
					} else {
						R0.mem = R1.mem = 0;
						R0.tok = R1.tok = 0;
						R0.exit_state = R1.exit_state = (unsigned short)PDEAD;
						R0.qn = R1.qn = 3;
						R0.clean = R1.clean = 0;
					}
					const u32 inc = scan_class_maps((u32)R0.qn | ((u32)R1.qn << 2), lane);
					const u32 excl = __shfl_up_sync(FULL, inc, 1);
					const u32 cls = q >= 2u ? q : (lane == 0 ? q : (excl >> (2 * q)) & 3u);
					const DecSuper &my = cls == 1u ? R1 : R0;
					const bool ok = cls < 2u && my.clean != 0;
					const u64 mymem = ok ? my.mem : 0ull;
					const u32 mytok = ok ? my.tok : 0u;
					u64 incm = mymem;
					u32 inct = mytok;
#pragma unroll
					for (int dd = 1; dd < 32; dd <<= 1) {
						const u64 tm = shfl_u64(incm, max(lane - dd, 0));
						const u32 tt = __shfl_up_sync(FULL, inct, dd);
						if (lane >= dd) {
							incm += tm;
							inct += tt;
						}
					}
					const bool reached = ok && cum + incm >= T;
					const u32 bal = __ballot_sync(FULL, reached || !ok);
					const int f = bal ? __ffs((int)bal) - 1 : 32;
					if (lane < f) { // consumed whole: dec_bulk_kernel writes its 32 segment records
						DecBulk bk;
						bk.w0 = sl * DWT_DEC_SUPER;
						bk.seg_base = nseg + 32u * (u32)lane;
						bk.cum0 = (u32)(cum + (incm - mymem));
						bk.j = (unsigned short)j;
						bk.cls = (unsigned short)cls;
						B.bulk[nbulk + lane] = bk;
					}
					nseg += 32u * (u32)f;
					nbulk += (u32)f;
					if (f > 0) {
						cum += shfl_u64(incm, f - 1);
						ones += __shfl_sync(FULL, inct, f - 1);
						carry_state = __shfl_sync(FULL, (u32)my.exit_state, f - 1);
						q = (__shfl_sync(FULL, inc, f - 1) >> (2 * q)) & 3u;
						w += DWT_DEC_SUPER * (u32)f;
					}
					if (f < 32)
						force_window = true; // something happens inside super-window w / 32: look at its windows
					RP_END(1);
				} else {
					// ---- whole windows, lane i looks at window w + i, up to the end of the super-window
					const int nlim = DWT_DEC_SUPER - (int)(w & (DWT_DEC_SUPER - 1));
					force_window = false;
					++n_window;
					RP_BEGIN();
					const u32 wl = w + (u32)lane;
					const bool valid = wl < nwin && lane < nlim;
					DecLink L0, L1;
					u32 xprev;
					if (pf_w == w) {
						L0 = pfL0;
						L1 = pfL1;
						xprev = pfX;
					} else if (wl < nwin) {
						load_link(B.link + 2 * (u64)wl, L0);
						load_link(B.link + 2 * (u64)wl + 1, L1);
						xprev = B.winX[wl - 1]; // w >= 1 in linked mode
					} else {
						dead_link(L0);
						dead_link(L1);
						xprev = 0xffffffffu;
					}
					if (!valid) {
						dead_link(L0);
						dead_link(L1);
					}
					const u32 inc = scan_class_maps((u32)L0.qn | ((u32)L1.qn << 2), lane);
					const u32 excl = __shfl_up_sync(FULL, inc, 1);
					// class the chain enters my window with
					const u32 cls = q >= 2u ? q : (lane == 0 ? q : (excl >> (2 * q)) & 3u);
					const DecLink &my = cls == 1u ? L1 : L0;
					const bool canon = cls < 2u && valid;
					const u64 mymem = canon ? my.mem : 0ull;
					const u32 mytok = canon ? my.tok : 0u;
					u64 incm = mymem;
					u32 inct = mytok;
#pragma unroll
					for (int dd = 1; dd < 32; dd <<= 1) {
						const u64 tm = shfl_u64(incm, max(lane - dd, 0));
						const u32 tt = __shfl_up_sync(FULL, inct, dd);
						if (lane >= dd) {
							incm += tm;
							inct += tt;
						}
					}
					const u64 pre = incm - mymem;
					const bool reached = canon && cum + incm >= T;
					const u32 bal = __ballot_sync(FULL, reached || !canon);
					const int f = bal ? __ffs((int)bal) - 1 : 32;
					const u32 my_entry = (xprev >> (16 * (cls & 1u))) & 0xffffu;
					if (lane < f) { // consumed completely
						DecSeg sg;
						sg.w = wl;
						sg.j = (unsigned short)j;
						sg.state = (unsigned short)my_entry;
						sg.i0 = 0;
						sg.m = my.m;
						sg.qm = my.qm;
						sg.cum0 = (u32)(cum + pre);
						sg.cum_m = (u32)(cum + pre + my.stray_mem);
						B.seg[nseg + lane] = sg;
					}
					nseg += (u32)f;
					RP_END(2);
					const u32 exits = (u32)L0.exit_state | ((u32)L1.exit_state << 16);
					if (f >= nlim) { // every window up to the super-window boundary was consumed (f == nlim)
						const int src = nlim - 1;
						cum += shfl_u64(incm, src);
						ones += __shfl_sync(FULL, inct, src);
						const u32 ex = __shfl_sync(FULL, exits, src);
						carry_state = (ex >> (16 * (__shfl_sync(FULL, cls, src) & 1u))) & 0xffffu;
						q = (__shfl_sync(FULL, inc, src) >> (2 * q)) & 3u;
						w += (u32)nlim;
						continue;
					}
					// lane f holds the window where something happens
					const u32 wf = w + (u32)f;
					const u32 cls_f = __shfl_sync(FULL, cls, f);
					const bool valid_f = __shfl_sync(FULL, (int)valid, f) != 0;
					const u64 cum_f = cum + shfl_u64(pre, f);
					const u32 ones_f = ones + __shfl_sync(FULL, inct - mytok, f);
					if (f > 0) {
						const u32 pcls = __shfl_sync(FULL, cls, f - 1);
						const u32 ex = __shfl_sync(FULL, exits, f - 1);
						carry_state = (ex >> (16 * (pcls & 1u))) & 0xffffu;
					}
					cum = cum_f;
					ones = ones_f;
					if (!valid_f || cls_f == 3u) {
						event = EV_STOP; // the chain died, or the stream is used up
						break;
					}
					if (cls_f == 2u) { // the chain left the previous window without joining: exact steps again
						mode = 0;
						gs = (u64)wf * WS;
						d = (int)(carry_state & 63u);
						k = (int)(carry_state >> 6);
						++slow_entries;
						RP_STAT(++reason[1]);
						continue;
					}
					const u32 m_f = __shfl_sync(FULL, (u32)my.m, f), qm_f = __shfl_sync(FULL, (u32)my.qm, f);
					const u32 smem_f = __shfl_sync(FULL, my.stray_mem, f), stok_f = __shfl_sync(FULL, my.stray_tok, f);
					const u32 entry_f = __shfl_sync(FULL, my_entry, f);
					if (m_f >= (u32)WS || T - cum_f <= (u64)smem_f) { // the pass ends before the chain joins
						mode = 0;
						gs = (u64)wf * WS;
						d = (int)(entry_f & 63u);
						k = (int)(entry_f >> 6);
						++slow_entries;
						RP_STAT(++reason[2]);
						continue;
					}
					if (lane == 0) {
						DecSeg sg;
						sg.w = wf;
						sg.j = (unsigned short)j;
						sg.state = (unsigned short)entry_f;
						sg.i0 = 0;
						sg.m = (unsigned short)m_f;
						sg.qm = qm_f;
						sg.cum0 = (u32)cum_f;
						sg.cum_m = (u32)(cum_f + smem_f);
						B.seg[nseg] = sg;
					}
					++nseg;
					cum = cum_f + smem_f;
					ones = ones_f + stok_f;
					search = true;
					sw = wf;
					sq = qm_f;
					sm = (int)m_f;
				}
				if (search) {
					// ---- the pass ends on class sq of window sw at or behind slice sm: smallest slice i >= sm with
					// cum + P(i + 1) - P(sm) >= T.  Every lane fetches the tables and the stream words of its WS / 32 slices
					// in one go, a ballot finds the slice, and that slice is stepped exactly.
					search = false;
					++n_search;
					RP_BEGIN();
					const u64 wbase = (u64)sw * WS;
					const u64 g4 = wbase + (u32)(NPL * lane);
					u64 pv[NPL + 1]; // members of class sq before my slices and before the slice behind them
					u32 tk4[NPL], en4[NPL];
					u64 s5[NPL + 1]; // stream words of my slices and the one behind them
#pragma unroll
					for (int t = 0; t < NPL; ++t) {
						const ulonglong2 p2 = B.P[g4 + t];
						pv[t] = sq ? p2.y : p2.x;
						en4[t] = B.E[g4 + t];
						tk4[t] = B.TK[g4 + t];
					}
#pragma unroll
					for (int t = 0; t < NPL + 1; ++t)
						s5[t] = ((g4 + t) << 6) < end_bits + 128 ? __ldg((const u64 *)stream + g4 + t) : 0ull;
					const ulonglong2 pt2 = B.winPT[sw];
					const u64 ptot = sq ? pt2.y : pt2.x;
					pv[NPL] = shfl_u64(pv[0], min(lane + 1, 31));
					if (lane == 31)
						pv[NPL] = ptot;
					// the value a lane holds for its t-th slice, t uniform across the warp
					auto pick64 = [&](const u64 *v, int t) {
						u64 r = v[0];
#pragma unroll
						for (int q2 = 1; q2 < NPL + 1; ++q2)
							r = t == q2 ? v[q2] : r;
						return r;
					};
					auto pick32 = [&](const u32 *v, int t) {
						u32 r = v[0];
#pragma unroll
						for (int q2 = 1; q2 < NPL; ++q2)
							r = t == q2 ? v[q2] : r;
						return r;
					};
					// P and TK at the start slice sm
					const int sm_lane = sm / NPL, sm_k = sm % NPL;
					const u64 pm = shfl_u64(pick64(pv, sm_k), sm_lane);
					const u32 tkm = __shfl_sync(FULL, pick32(tk4, sm_k), sm_lane);
					u32 hit = 0;
#pragma unroll
					for (int t = 0; t < NPL; ++t) {
						const int idx = NPL * lane + t;
						if (idx >= sm && cum + (pv[t + 1] - pm) >= T)
							hit |= 1u << t;
					}
					const u32 bal = __ballot_sync(FULL, hit != 0);
					if (!bal) {
						event = EV_STOP; // cannot happen: the caller saw that the window covers the rest
						tripped = 2;
						break;
					}
					const int fl = __ffs((int)bal) - 1;
					const int fk = __ffs((int)__shfl_sync(FULL, hit, fl)) - 1;
					const int lo = NPL * fl + fk;
					const u64 egs = wbase + lo;
					const u64 pi = shfl_u64(pick64(pv, fk), fl);
					const u32 tki = __shfl_sync(FULL, pick32(tk4, fk), fl);
					const u32 e2 = __shfl_sync(FULL, pick32(en4, fk), fl);
					const u64 a = shfl_u64(pick64(s5, fk), fl);
					u64 b = shfl_u64(pick64(s5, fk + 1), fl);
					if (((egs + 1) << 6) >= end_bits + 64)
						b = 0ull; // load_slice's rule for the second word
					const u32 e = (e2 >> (16 * sq)) & 0xffffu;
					cum += pi - pm;
					ones += sq ? (tki >> 16) - (tkm >> 16) : (tki & 0xffffu) - (tkm & 0xffffu);
					d = (int)(e & 63u);
					k = (int)(e >> 6);
					if (e == PDEAD) {
						event = EV_STOP; // cannot happen: the prefixes say tokens start here
					} else {
						event = walk_events(lut, a, b, clamp_avail(end_bits, egs << 6), T, d, k, cum, ones, f_pending);
						if (event == EV_NONE && cum >= T)
							event = EV_COVERED; // the pass ends exactly with the slice's last token
						if (event == EV_NONE)
							event = EV_STOP; // cannot happen
					}
					f_pos = (egs << 6) + (u64)d;
					f_k = k;
					RP_END(3);
				}
			}
			if (event == EV_COVERED)
				f_pending = 0;
		}
		// ---------------------------------------------------- refinement pass bookkeeping (decode.c:89-98,206,223,240)
		if (event == EV_STOP) {
			stop = true;
			f_pos = bitpos;
			f_k = order;
			f_pending = 0;
		}
		bool complete = !stop;
		int ref_valid = 0;
		const u64 ref_pos = f_pos;
		if (!stop && nref > 0) {
			if (f_pending > 1) {
				stop = true; // rle.h:98-99: a pending run must end exactly at the phantom one
				complete = false;
			} else {
				f_pending = 0;
				ref_valid = 1;
				if (f_pos + nref > end_bits) {
					stop = true; // partial refinement: the deposit keeps the bits before EOF
					complete = false;
				} else {
					f_pos += nref;
				}
			}
		}
		bitpos = f_pos;
		order = f_k;
		pending = f_pending;
		stopped = stop;
		if (lane == 0) {
			DecChunk ck;
			ck.rank_base = my_rank_base;
			ck.ref_bitpos = ref_pos;
			ck.r0 = (u32)r0;
			ck.T = (u32)T;
			ck.parsed = 1;
			ck.ref_valid = ref_valid;
			B.chunks[j] = ck;
			if (complete)
				missing[c * 16 + l] -= 1;
			sigcount[c * 16 + l] = nsig + ones;
		}
		__syncwarp();
	}
	__syncwarp();
	for (int i = lane; i < 48; i += 32)
		st->missing[i] = missing[i];
	if (lane == 0) {
		st->bitpos = bitpos;
		st->order = order;
		st->pending = pending;
		st->stopped = stopped ? 1 : 0;
		st->level = level;
		st->nseg = nseg;
		st->nbulk = nbulk;
		st->slow_entries = slow_entries;
		st->exact_steps = exact_steps;
		st->n_super = n_super;
		st->n_window = n_window;
		st->n_search = n_search;
		for (int t = 0; t < 10; ++t)
			st->dbg_hist[t] = RP_HIST(t);
		for (int t = 0; t < 4; ++t)
			st->dbg_reason[t] = RP_REASON(t);
		st->guard_tripped = tripped;
#ifdef DWT_RESOLVE_PROFILE
		printf("resolver cycles: exact-step visits %lld, super rounds %lld, window rounds %lld, end searches %lld, total %lld\n", rp_cyc[0],
		       rp_cyc[1], rp_cyc[2], rp_cyc[3], clock64() - rp_start);
#endif
	}
}

// ---------------------------------------------------------------------------------------------- bulk segments

// The 32 segment records of every super-window the resolver consumed whole: the same class scan the resolver runs for a
// window round, one warp per super-window, all of them in parallel.
__global__ void __launch_bounds__(128) dec_bulk_kernel(const __grid_constant__ DecBuffers B)
{
	const u32 FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const u32 nbulk = B.state->nbulk;
	for (u32 r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nbulk; r += gridDim.x * (blockDim.x >> 5)) {
		const DecBulk bk = B.bulk[r];
		const u32 wl = bk.w0 + (u32)lane; // a clean super-window: every window exists, w0 >= 32
		DecLink L0, L1;
		load_link(B.link + 2 * (u64)wl, L0);
		load_link(B.link + 2 * (u64)wl + 1, L1);
		const u32 xprev = B.winX[wl - 1];
		const u32 inc = scan_class_maps((u32)L0.qn | ((u32)L1.qn << 2), lane);
		const u32 excl = __shfl_up_sync(FULL, inc, 1);
		const u32 q = bk.cls;
		const u32 cls = lane == 0 ? q : (excl >> (2 * q)) & 3u;
		const DecLink &my = cls == 1u ? L1 : L0;
		u64 incm = my.mem;
#pragma unroll
		for (int dd = 1; dd < 32; dd <<= 1) {
			const u64 tm = shfl_u64(incm, max(lane - dd, 0));
			if (lane >= dd)
				incm += tm;
		}
		const u64 pre = incm - my.mem;
		DecSeg sg;
		sg.w = wl;
		sg.j = bk.j;
		sg.state = (unsigned short)((xprev >> (16 * (cls & 1u))) & 0xffffu);
		sg.i0 = 0;
		sg.m = my.m;
		sg.qm = my.qm;
		sg.cum0 = (u32)((u64)bk.cum0 + pre);
		sg.cum_m = (u32)((u64)bk.cum0 + pre + my.stray_mem);
		B.seg[bk.seg_base + lane] = sg;
	}
}

// ---------------------------------------------------------------------------------------------- emit

// rank-space bits of consecutive tokens mostly fall into the same 32-bit word: they are collected in registers and
// sent with one atomic per word.  All of a slice's bookkeeping is 32 bit: a chunk has fewer than 2^31 members, ranks are
// kept relative to the word the chunk's rank space starts in (the callers pass the two bit vectors already offset by that
// word), and a token that fits a 32-bit stream window together with its sign bit (all but the long runs) is taken apart
// without 64-bit shifts.
struct RankAcc {
	u32 word; // relative to the chunk's first rank word
	u32 ones, signs;
};

__device__ __forceinline__ void acc_flush(RankAcc &A, u32 *ones_w, u32 *sign_w)
{
	if (A.ones)
		atomicOr(ones_w + A.word, A.ones);
	if (A.signs)
		atomicOr(sign_w + A.word, A.signs);
	A.ones = A.signs = 0;
}

// ones and signs of the tokens that start in one slice; false when the chain ends here.  T = members the chunk's own
// tokens cover, cum (< T on entry) = members consumed so far, r0 = bit of the chunk's first rank inside its word.
// (A table-driven walk like the scan's was measured here and lost: staging 32 KB per CTA and the divergence between
// table and generic steps cost more than the per-token work they save.)
__device__ __forceinline__ bool emit_slice(u64 a, u64 b, int avail, u32 T, u32 r0, int &d, int &k, u32 &cum, RankAcc &A,
                                           u32 *ones_w, u32 *sign_w)
{
	while (d < 64) {
		const u32 lo = window32(a, b, d);
		const int u = lo ? __ffs((int)lo) - 1 : 32;
		const int e = k + u;
		const int L = u + 1 + e;
		if (e > 31 || d + L > avail)
			return false;
		u32 payload, sign;
		if (L < 32) {
			payload = (lo >> (u + 1)) & ((1u << e) - 1u);
			sign = (lo >> L) & 1u;
		} else {
			const u64 w = bits_from(a, b, d);
			payload = (u32)(w >> (u + 1)) & ((1u << e) - 1u);
			sign = (u32)(w >> L) & 1u;
		}
		const u32 n = (1u << e) - (1u << k) + payload; // the run in front of this one (< 2^32)
		if (n >= T - cum)
			return false;
		const u32 one = cum + n;
		const u32 rel = r0 + one;
		if ((rel >> 5) != A.word) {
			acc_flush(A, ones_w, sign_w);
			A.word = rel >> 5;
		}
		const u32 bit = 1u << (rel & 31);
		A.ones |= bit;
		if (d + L + 1 > avail)
			return false;
		if (sign)
			A.signs |= bit;
		cum = one + 1;
		d += L + 1;
		k = e >= 2 ? e - 2 : 0;
	}
	return cum < T;
}

__global__ void __launch_bounds__(WS) dec_emit_kernel(const __grid_constant__ DecBuffers B)
{
	const u32 nseg = B.state->nseg;
	const int i = threadIdx.x;
	for (u32 sidx = blockIdx.x; sidx < nseg; sidx += gridDim.x) {
		const DecSeg s = B.seg[sidx];
		const DecChunk ck = B.chunks[s.j];
		const int m = s.m;
		const u32 T = ck.T;
		const u64 base = ck.rank_base + ck.r0;
		const u32 r0 = (u32)(base & 31u);
		u32 *ones_w = B.ones_rank + (base >> 5), *sign_w = B.sign_rank + (base >> 5);
		const u64 wbase = (u64)s.w * WS;
		if (i == s.i0 && m > i) {
			// exact steps through the slices in front of the join
			int d = (int)(s.state & 63u), k = (int)(s.state >> 6);
			u32 cum = s.cum0;
			RankAcc A = {~0u, 0u, 0u};
			if (cum < T) {
				for (int ii = i; ii < m; ++ii) {
					u64 a, b;
					const u64 gs = wbase + ii;
					load_slice(B.stream, B.end_bits, gs, a, b);
					if (!emit_slice(a, b, clamp_avail(B.end_bits, gs << 6), T, r0, d, k, cum, A, ones_w, sign_w))
						break;
					d -= 64;
				}
			}
			acc_flush(A, ones_w, sign_w);
		} else if (i >= m) {
			const u64 gs = wbase + i;
			const u32 e = (B.E[gs] >> (16 * s.qm)) & 0xffffu;
			if (e == PDEAD)
				continue;
			const ulonglong2 pi = B.P[gs], pm = B.P[wbase + m];
			const u64 cum64 = (u64)s.cum_m + (s.qm ? pi.y - pm.y : pi.x - pm.x);
			if (cum64 >= T)
				continue;
			u32 cum = (u32)cum64;
			u64 a, b;
			load_slice(B.stream, B.end_bits, gs, a, b);
			int d = (int)(e & 63u), k = (int)(e >> 6);
			RankAcc A = {~0u, 0u, 0u};
			emit_slice(a, b, clamp_avail(B.end_bits, gs << 6), T, r0, d, k, cum, A, ones_w, sign_w);
			acc_flush(A, ones_w, sign_w);
		}
	}
}

// ---------------------------------------------------------------------------------------------- deposit

__device__ __forceinline__ u32 group_valid_mask(const Geom &G, int l, int g)
{
	if (g >= G.G[l])
		return 0u;
	long long rem = G.num[l] - (long long)g * 32;
	return rem >= 32 ? 0xffffffffu : ((1u << (int)rem) - 1u);
}

__device__ __forceinline__ int level_of_tile(const Geom &G, int tile)
{
	int l = G.levels - 1; // from the top: three quarters of the tiles belong to the finest level
	while (l > 0 && G.tbase[l] > tile)
		--l;
	return l;
}

// chunk of (channel c, level l) at plane depth `depth` (0 = the channel's top plane), or -1
__device__ __forceinline__ int chunk_at(const Sched *S, const DecChunk *chunks, int nchunks, int c, int l, int depth,
                                        int *plane)
{
	const int p = S->planes[c] - 1 - depth;
	if (p < 0)
		return -1;
	const int j = S->chunk_of[c][l][p];
	if (j < 0 || j >= nchunks || !chunks[j].parsed)
		return -1;
	*plane = p;
	return j;
}

// members (still insignificant) and refinement positions per tile of 256 groups
__global__ void __launch_bounds__(TG) dec_prep_kernel(const __grid_constant__ Geom G, const __grid_constant__ DecBuffers B,
                                                       int nchunks, int depth)
{
	__shared__ u32 acc[2];
	const int c = blockIdx.y, tile = blockIdx.x;
	const int l = level_of_tile(G, tile);
	int p;
	if (chunk_at(B.sched, B.chunks, nchunks, c, l, depth, &p) < 0)
		return;
	if (threadIdx.x < 2)
		acc[threadIdx.x] = 0;
	__syncthreads();
	const int g = (tile - G.tbase[l]) * TG + threadIdx.x;
	const u32 vm = group_valid_mask(G, l, g);
	const u32 s = vm ? B.sig[(size_t)c * G.GT + G.gbase[l] + g] : 0u;
	const u32 m = __reduce_add_sync(0xffffffffu, (u32)__popc(vm & ~s));
	const u32 r = __reduce_add_sync(0xffffffffu, (u32)__popc(s));
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&acc[0], m);
		atomicAdd(&acc[1], r);
	}
	__syncthreads();
	if (threadIdx.x < 2)
		B.tile_sums[2 * ((size_t)c * G.tbase[G.levels] + tile) + threadIdx.x] = acc[threadIdx.x];
}

// exclusive prefixes of the tile counts inside every (channel, level)
__global__ void __launch_bounds__(1024) dec_tilescan_kernel(const __grid_constant__ Geom G,
                                                             const __grid_constant__ DecBuffers B, int nchunks, int depth)
{
	__shared__ u64 ws[32];
	const int l = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
	int p;
	if (chunk_at(B.sched, B.chunks, nchunks, c, l, depth, &p) < 0)
		return;
	const int ntile = G.ntile[l];
	const u32 *sums = B.tile_sums + 2 * ((size_t)c * G.tbase[G.levels] + G.tbase[l]);
	u32 *base = B.tile_base + 2 * ((size_t)c * G.tbase[G.levels] + G.tbase[l]);
	const int per = (ntile + 1023) / 1024;
	const int b = tid * per, e = min(b + per, ntile);
	u64 sm = 0, sr = 0;
	for (int i = b; i < e; ++i) {
		sm += sums[2 * i];
		sr += sums[2 * i + 1];
	}
	u64 tm, tr;
	const u64 bm = block_exscan_u64(sm, ws, &tm);
	const u64 br = block_exscan_u64(sr, ws, &tr);
	u32 rm = (u32)bm, rr = (u32)br;
	for (int i = b; i < e; ++i) {
		const u32 m = sums[2 * i], r = sums[2 * i + 1];
		base[2 * i] = rm;
		base[2 * i + 1] = rr;
		rm += m;
		rr += r;
	}
}

// also leaves the tile counts of the NEXT depth (members / refinement positions after this plane) in tile_sums,
// so that dec_prep_kernel only runs for the top plane
__global__ void __launch_bounds__(TG) dec_deposit_kernel(const __grid_constant__ Geom G, const __grid_constant__ DecBuffers B,
                                                          int nchunks, int depth)
{
	__shared__ __align__(16) u32 ws[2][8];
	__shared__ u32 acc[2];
	const int c = blockIdx.y, tile = blockIdx.x;
	const int l = level_of_tile(G, tile);
	int p;
	const int j = chunk_at(B.sched, B.chunks, nchunks, c, l, depth, &p);
	if (j < 0)
		return;
	if (threadIdx.x < 2)
		acc[threadIdx.x] = 0;
	const DecChunk ck = B.chunks[j];
	const Sched *S = B.sched;
	const int g = (tile - G.tbase[l]) * TG + threadIdx.x;
	const size_t gi = (size_t)G.gbase[l] + g;
	u32 *sig = B.sig + (size_t)c * G.GT;
	const u32 vm = group_valid_mask(G, l, g);
	const u32 s = vm ? sig[gi] : 0u;
	const u32 member = vm & ~s;
	const u32 nm = __popc(member), nr = __popc(s);
	const u32 ex2 = tile_exscan_2x16(nm | (nr << 16), ws, 0); // syncs: acc is initialised behind it
	const u64 ex = (u64)(ex2 & 0xffffu) | ((u64)(ex2 >> 16) << 32);
	const size_t ti = 2 * ((size_t)c * G.tbase[G.levels] + tile);
	u32 Bw = 0;
	if (vm) {
		const u32 *tb = B.tile_base + ti;
		u32 *plane_words = B.bs + S->bsbase[c] + (long long)p * G.GT;
		u32 *sign_words = B.bs + S->bsbase[c] + (long long)S->planes[c] * G.GT;
		if (nm) {
			const u64 off = ck.rank_base + tb[0] + (u32)ex;
			u32 ob = bits_get32(B.ones_rank, off);
			if (nm < 32)
				ob &= (1u << nm) - 1u;
			if (ob) {
				const ExpandPlan pl = bit_expand_plan(member); // ones and their signs go through the same mask
				Bw = bit_expand_apply(ob, pl);
				const u32 sb = bits_get32(B.sign_rank, off) & ob;
				if (sb)
					sign_words[gi] |= bit_expand_apply(sb, pl);
			}
		}
		if (nr && ck.ref_valid) {
			const u64 pos = ck.ref_bitpos + tb[1] + (u32)(ex >> 32);
			const u64 end = B.end_bits;
			if (pos < end) {
				const u64 w = pos >> 5;
				const int sh = (int)(pos & 31);
				u32 rb = __funnelshift_r(__ldg(B.stream + w), __ldg(B.stream + w + 1), sh);
				const u64 avail = end - pos;
				u32 take = nr;
				if (avail < take)
					take = (u32)avail;
				if (take < 32)
					rb &= (1u << take) - 1u;
				Bw |= bit_expand(rb, s);
			}
		}
		plane_words[gi] = Bw;
		if (Bw)
			sig[gi] = s | Bw;
	}
	const u32 ns = s | Bw;
	const u32 m2 = __reduce_add_sync(0xffffffffu, (u32)__popc(vm & ~ns));
	const u32 r2 = __reduce_add_sync(0xffffffffu, (u32)__popc(ns));
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&acc[0], m2);
		atomicAdd(&acc[1], r2);
	}
	__syncthreads();
	if (threadIdx.x < 2)
		B.tile_sums[ti + threadIdx.x] = acc[threadIdx.x];
}

} // namespace

u64 dec_rank_bits(const Geom &g, const Sched &hs, int nchunks)
{
	u64 bits = 0;
	for (int j = 0; j < nchunks; ++j)
		bits += ((u64)g.num[hs.level[j]] + 63) & ~63ull;
	return bits;
}

void dec_token_table(u32 *host_table)
{
	for (u32 i = 0; i < (1u << LUT_BITS); ++i) {
		host_table[i] = toklut_entry(i);
		host_table[(1u << LUT_BITS) + i] = toklut_masks(i);
	}
}

// A lineage pass costs ~0.13 ms of latency whatever the stream size (one window's positions walked by one thread) and
// saves exact resolver steps in proportion to the stream's slow-to-synchronise regions -- the top planes of the big
// levels, which every stream of a large image starts with.  Measured (single frame, coder stage, B200): 8K photo 4.35 /
// 4.04 / 3.80 / 3.62 / 3.43 / 3.57 / 3.61 ms for 0 / 1 / 2 / 3 / 5 / 8 / 12 passes; its first 1 MiB 3.26 / 2.57 / 2.19 / 2.24 ms
// for 0 / 2 / 4 / 8; 4K photo 1.82 / 1.83 / 1.97 ms for 0 / 1 / 2.
static int lineage_passes(const Geom &g, u32 nwin)
{
	if (g.pix[g.levels] < 20000000LL || nwin < 512u)
		return 0;
	return nwin >= 8192u ? 5 : 4;
}

int dec_run(const Geom &g, const Sched &hs, const DecBuffers &b_in, int nchunks, cudaStream_t st, long long *launches)
{
	DecBuffers b = b_in;
	// windows of 128 slices give the one-thread-per-chain walk enough warps to hide its latencies at any stream size, and it
	// decodes every token once: it is the default; the CTA-per-window kernel stays as the cross-check (scan_mode 1)
	const bool serial = b.scan_mode != 1;
	if (serial)
		dec_scan_serial_kernel<<<(2 * b.nwin + 127) / 128, 128, 0, st>>>(b.stream, b.end_bits, b.toklut, b.nwin, b.E, b.P, b.TK,
		                                                                  b.winX, b.winPT, b.winTT);
	else
		dec_scan_kernel<<<b.nwin, WS, 0, st>>>(b.stream, b.end_bits, b.toklut, b.E, b.P, b.TK, b.winX, b.winPT, b.winTT);
	++*launches;
	{
		u32 *xin = b.winX, *xout = b.winX2;
		unsigned char *cin = nullptr, *cout = b.chg;
		static const int forced = getenv("DWT_LINEAGE") ? atoi(getenv("DWT_LINEAGE")) : -1; // tuning aid
		// several frames in flight: the resolver's latency hides behind the other frames' kernels, the passes are only work
		int passes = forced >= 0 ? forced : (b.in_flight >= 4 ? 0 : lineage_passes(g, b.nwin));
		if (passes > DWT_DEC_MAX_LINEAGE)
			passes = DWT_DEC_MAX_LINEAGE;
		// a warp per listed window, eight per CTA; CTAs without a window leave at once
		const u32 ext_want = (b.nwin + XW_WARPS - 1) / XW_WARPS, ext_cap = 8u * (u32)dwt_device_sms();
		const unsigned ext_grid = (unsigned)(ext_want < ext_cap ? ext_want : ext_cap);
		for (int pass = 0; pass < passes; ++pass) {
			u32 *cnt = b.ext_count + pass; // one counter per pass, zeroed by the caller
			dec_extend_find_kernel<<<(b.nwin + 127) / 128, 128, 0, st>>>(b.stream, b.end_bits, b.nwin, b.E, xin, xout, cin, cout, cnt,
			                                                             b.ext_list);
			dec_extend_walk_kernel<<<ext_grid, XW_WARPS * 32, 0, st>>>(b.stream, b.end_bits, b.toklut, b.E, b.P, b.TK, xout, b.winPT, b.winTT, cout,
			                                                cnt, b.ext_list);
			*launches += 2;
			u32 *t = xin;
			xin = xout;
			xout = t;
			cin = cout;
			cout = cout == b.chg ? b.chg + b.nwin : b.chg;
		}
		b.winX = xin; // the exits after the last pass
	}
	dec_link_kernel<<<(2 * b.nwin + 127) / 128, 128, 0, st>>>(b.stream, b.end_bits, b.nwin, b.E, b.P, b.TK, b.winX, b.winPT,
	                                                           b.winTT, b.link);
	dec_super_kernel<<<(b.nsuper + 3) / 4, 128, 0, st>>>(b.nwin, b.nsuper, b.link, b.super);
	dec_resolve_kernel<<<1, 32, 0, st>>>(g, nchunks, b);
	dec_bulk_kernel<<<(b.nsuper + 3) / 4 < 1184u ? (b.nsuper + 3) / 4 : 1184u, 128, 0, st>>>(b);
	{
		const int sms = dwt_device_sms();
		const u32 want = b.nwin + 2u * (u32)nchunks, cap = (u32)sms * 64u;
		dec_emit_kernel<<<want < cap ? want : cap, WS, 0, st>>>(b);
	}
	*launches += 5;
	int depth_max = 0;
	for (int c = 0; c < g.channels; ++c)
		if (hs.planes[c] > depth_max)
			depth_max = hs.planes[c];
	const dim3 tiles((unsigned)g.tbase[g.levels], (unsigned)g.channels);
	const dim3 lv((unsigned)g.levels, (unsigned)g.channels);
	if (depth_max > 0) {
		dec_prep_kernel<<<tiles, TG, 0, st>>>(g, b, nchunks, 0);
		++*launches;
	}
	for (int depth = 0; depth < depth_max; ++depth) {
		dec_tilescan_kernel<<<lv, 1024, 0, st>>>(g, b, nchunks, depth);
		dec_deposit_kernel<<<tiles, TG, 0, st>>>(g, b, nchunks, depth); // leaves the tile counts of depth + 1
		*launches += 2;
	}
	CUDA_OK(cudaGetLastError());
	return 0;
}
