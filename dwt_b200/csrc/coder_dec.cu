// coder_dec.cu -- the bit-plane decoder of decode.c:67-100,187-243 + rle.h + vli.h + bits.h on the GPU.
//
// Chunks (channel, level, plane) are decoded in schedule order; within a chunk:
//   prep     per tile of 256 groups: how many coefficients are still insignificant (members of the
//            significance pass) and how many are already significant (refinement bits)   [dec_prep_kernel]
//   parse    the significance pass is a serial chain of [adaptive-Rice run][sign] tokens.  One CTA walks the
//            stream in windows of 1024 slices of 64 bits.  Every slice is parsed speculatively from its first
//            two bit offsets at order 0 (at order 0 all tokens have even length, so a chain keeps its parity)
//            and remembers WHERE its chains had token starts (a 64-bit "visited" mask).  A slice then
//            classifies each possible exit of its predecessor by looking it up in its masks (exact merge
//            detection; an exit that joins no chain is parsed on the spot as an extra hypothesis), which
//            turns the serial chain into a scan over 5-state maps.  Token run lengths are scanned into member
//            ranks and ones / signs are set in rank space.  EOF, the run carried across chunks
//            (rle.h:66-77) and the phantom one before refinement bits (rle.h:91-103) follow the reference
//            exactly.                                                                      [dec_parse_kernel]
//   deposit  rank-space bits are expanded into the insignificant positions of each group (software pdep),
//            refinement bits are taken straight from the stream, significance is updated [dec_deposit_kernel]
#include "coder.cuh"

namespace {

constexpr int TG = DWT_TILE_GROUPS;
constexpr int PT = 1024;   // parse threads = slices per window
constexpr int SLICE = 64;  // stream bits per slice; >= the longest token (31 zeros + 1 + 31 payload + sign)
constexpr int DEAD = 255;  // a chain that cannot continue (EOF inside a token, impossible order)
constexpr int ABSENT = 254;
constexpr u32 ST_U = 4;    // "undetermined" chain state (absorbing)

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

// Parse the [VLI][sign] tokens that start inside the slice [lo, lo+SLICE), beginning at (pos, k).
// Returns the exit (first token start at or behind the slice end), the members consumed, and the mask of
// slice offsets at which this chain had a token start while at order 0.
__device__ __forceinline__ void run_slice(const u32 *__restrict__ s, u64 end_bits, u64 lo, u64 &pos, int &k, u64 &csum,
                                          u64 &visited)
{
	const u64 lim = lo + SLICE;
	csum = 0;
	visited = 0;
	while (k != DEAD && pos < lim) {
		if (k == 0)
			visited |= 1ull << (int)(pos - lo);
		u64 n, w;
		int len, kn;
		if (!read_vli(s, end_bits, pos, k, &n, &len, &kn, &w)) {
			k = DEAD;
			break;
		}
		csum += n + 1;
		pos += len + 1;
		k = kn;
	}
}

enum { EV_NONE = 0, EV_COVERED = 1, EV_PENDING = 2, EV_STOP = 3 };

// chain-state maps: input = hypothesis (0..3) the predecessor slice's chain follows, output = hypothesis this
// slice's chain follows (0..3) or ST_U.  3 bits per input.
__device__ __forceinline__ u32 map_apply(u32 m, u32 s)
{
	return s >= ST_U ? ST_U : (m >> (3 * s)) & 7u;
}
__device__ __forceinline__ u32 map_compose(u32 first, u32 then)
{
	u32 r = 0;
#pragma unroll
	for (u32 s = 0; s < 4; ++s)
		r |= map_apply(then, map_apply(first, s)) << (3 * s);
	return r;
}
constexpr u32 MAP_ID = 0u | (1u << 3) | (2u << 6) | (3u << 9);

__global__ void __launch_bounds__(PT) dec_parse_kernel(DecState *st, const u32 *__restrict__ stream,
                                                        const u32 *__restrict__ tile_sums, u32 *tile_base, int ntile,
                                                        u32 *ones_rank, u32 *sign_rank, int chan, int level)
{
	__shared__ u64 ws[32];
	__shared__ u32 wmap[32];
	__shared__ u32 x_off[4][PT];          // exits of the (up to) four hypotheses of every slice, relative to the window
	__shared__ unsigned char x_k[4][PT];
	__shared__ unsigned char sigma[PT];   // hypothesis the true chain follows in every slice
	__shared__ int first_u, winner;
	__shared__ u64 f_pos; // final state written by the winning thread
	__shared__ int f_k, f_event;
	__shared__ u32 f_pending;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (st->stopped)
		return;

	// ---- exclusive prefixes of the per-tile member / refinement counts
	{
		const int per = (ntile + PT - 1) / PT;
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
		if (tid == 0) {
			st->n_member = (u32)tm;
			st->n_ref = (u32)tr;
			if (st->level < level)
				st->level = level; // decode.c:203,219-220,236-237: the chunk is started
			f_event = EV_NONE;
		}
	}
	__syncthreads();
	const u64 end_bits = st->end_bits;
	const u64 R = st->n_member;
	const u32 nref = st->n_ref;
	u64 bitpos = st->bitpos;
	int order = st->order;
	u32 pending = st->pending;
	u64 r0 = 0; // members already accounted for
	bool stop = false;
	u32 n_windows = 0, n_short = 0;
	__syncthreads();

	// ---- a run carried in from earlier chunks (rle.h:66-77): (pending-1) zeros, then a one
	if (pending > 0) {
		if ((u64)pending - 1 >= R) {
			pending -= (u32)R;
			r0 = R;
		} else {
			const u64 rk = pending - 1;
			if (tid == 0)
				atomicOr(ones_rank + (rk >> 5), 1u << (rk & 31));
			if (bitpos < end_bits) {
				if (tid == 0 && (peek64(stream, bitpos) & 1ull))
					atomicOr(sign_rank + (rk >> 5), 1u << (rk & 31));
				bitpos += 1;
			} else {
				stop = true; // the sign bit hits EOF: the magnitude bit stays (decode.c:80-86)
			}
			r0 = (u64)pending;
			pending = 0;
		}
	}

	// ---- significance pass: windows of PT slices
	while (!stop && pending == 0 && r0 < R) {
		++n_windows;
		const u64 Rrem = R - r0;
		const u64 wb = bitpos;
		const u64 sub_lo = wb + (u64)tid * SLICE;
		u64 hv[4] = {0, 0, 0, 0};      // visited masks of my hypotheses
		u64 he_pos[4] = {0, 0, 0, 0};  // their entries
		int he_k[4] = {ABSENT, ABSENT, ABSENT, ABSENT};
		u64 dummy;

		// (1) two hypotheses per slice: a token starts at slice offset 0 / 1 at order 0 (slice 0 knows the truth)
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			u64 p = tid == 0 ? wb : sub_lo + h;
			int k = tid == 0 ? order : 0;
			he_pos[h] = p;
			he_k[h] = k;
			if (tid == 0 && h == 1) {
				x_off[1][0] = x_off[0][0];
				x_k[1][0] = x_k[0][0];
				hv[1] = hv[0];
			} else {
				run_slice(stream, end_bits, sub_lo, p, k, dummy, hv[h]);
				x_off[h][tid] = (u32)(p - wb);
				x_k[h][tid] = (unsigned char)k;
			}
		}
		x_k[2][tid] = ABSENT;
		x_k[3][tid] = ABSENT;
		if (tid == 0)
			first_u = PT;
		__syncthreads();

		// which of my chains does an entry (p, k) join?  exact: it must coincide with one of their token starts
		auto classify = [&](u64 p, int k) -> u32 {
			if (k == DEAD || k == ABSENT || p < sub_lo || p >= sub_lo + SLICE)
				return ST_U;
#pragma unroll
			for (int h = 0; h < 4; ++h) {
				if (he_k[h] == ABSENT)
					continue;
				if (k == 0 && ((hv[h] >> (int)(p - sub_lo)) & 1ull))
					return (u32)h;
				if (p == he_pos[h] && k == he_k[h])
					return (u32)h;
			}
			return ST_U;
		};

		// (2) classify the predecessor's two exits; an exit that joins none of my chains becomes a new hypothesis
		u32 cls[4] = {ST_U, ST_U, ST_U, ST_U};
		if (tid > 0) {
#pragma unroll
			for (int s = 0; s < 2; ++s) {
				u64 p = wb + x_off[s][tid - 1];
				int k = x_k[s][tid - 1];
				if (k == DEAD)
					continue;
				u32 c = classify(p, k);
				if (c == ST_U && p >= sub_lo && p < sub_lo + SLICE) {
					const int h = 2 + s;
					he_pos[h] = p;
					he_k[h] = k;
					run_slice(stream, end_bits, sub_lo, p, k, dummy, hv[h]);
					x_off[h][tid] = (u32)(p - wb);
					x_k[h][tid] = (unsigned char)k;
					c = (u32)h;
				}
				cls[s] = c;
			}
		}
		__syncthreads();
		// (3) classify the predecessor's extra hypotheses (no further parsing: an unknown ends the window early)
		u32 mymap;
		if (tid > 0) {
#pragma unroll
			for (int s = 2; s < 4; ++s) {
				int k = x_k[s][tid - 1];
				if (k != ABSENT)
					cls[s] = classify(wb + x_off[s][tid - 1], k);
			}
			mymap = cls[0] | (cls[1] << 3) | (cls[2] << 6) | (cls[3] << 9);
		} else {
			mymap = 0; // slice 0 follows its hypothesis 0 (the true entry) whatever comes in
		}
		// inclusive scan of the maps: sigma_i = (M_i o ... o M_0)(0)
		u32 inc = mymap;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			u32 t = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d)
				inc = map_compose(t, inc);
		}
		if (lane == 31)
			wmap[wid] = inc;
		__syncthreads();
		u32 pre = MAP_ID;
		for (int i = 0; i < wid; ++i)
			pre = map_compose(pre, wmap[i]);
		const u32 sg = map_apply(map_compose(pre, inc), 0);
		sigma[tid] = (unsigned char)sg;
		if (sg == ST_U)
			atomicMin(&first_u, tid);
		if (tid == 0)
			winner = PT;
		__syncthreads();
		const int J = first_u; // slices [0, J) have an exact entry
		if (J < PT)
			++n_short;

		// (4) count pass from the exact entry, scan into member ranks, then the walk that writes ones and signs
		u64 e_pos = wb;
		int e_k = order;
		if (tid > 0 && tid < J) {
			const int sp = sigma[tid - 1];
			e_pos = wb + x_off[sp][tid - 1];
			e_k = x_k[sp][tid - 1];
		}
		u64 csum = 0;
		if (tid < J) {
			u64 p = e_pos, v;
			int k = e_k;
			run_slice(stream, end_bits, sub_lo, p, k, csum, v);
		}
		u64 total;
		u64 cum = block_exscan_u64(csum, ws, &total); // members consumed before this slice (syncs inside)
		{
			u64 pos = e_pos;
			int k = tid < J ? e_k : DEAD;
			int ev = EV_NONE;
			u64 ev_pos = 0;
			int ev_k = 0;
			u32 ev_pending = 0;
			const u64 lim = sub_lo + SLICE;
			while (k != DEAD && pos < lim) {
				if (cum >= Rrem) {
					ev = EV_COVERED;
					ev_pos = pos;
					ev_k = k;
					break;
				}
				u64 n, w;
				int len, kn;
				if (!read_vli(stream, end_bits, pos, k, &n, &len, &kn, &w)) {
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
					if ((w >> len) & 1ull)
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
			if (f_event == EV_STOP) {
				stop = true;
			} else {
				bitpos = f_pos;
				order = f_k;
				pending = f_event == EV_PENDING ? f_pending : 0;
				r0 = R;
			}
			break;
		}
		// no end inside this window: continue behind the last exact slice
		{
			const int sl = sigma[J - 1];
			bitpos = wb + x_off[sl][J - 1];
			order = x_k[sl][J - 1];
		}
		r0 += total;
		__syncthreads();
		if (order == DEAD) {
			stop = true;
			break;
		}
	}

	// ---- refinement pass (raw bits) and bookkeeping, decode.c:89-98,206,223,240
	if (tid == 0) {
		const bool sig_done = !stop; // every member has its symbol (or is covered by the carried run)
		bool complete = sig_done;
		int ref_valid = 0;
		u64 ref_pos = bitpos;
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
		st->order = order == DEAD ? 0 : order;
		st->pending = pending;
		st->stopped = stop ? 1 : 0;
		st->chunk_done += 1;
		st->dbg_windows += n_windows;
		st->dbg_iters += n_short;
		if (complete)
			st->missing[chan * 16 + level] -= 1;
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

} // namespace

int dec_chunk(const Geom &g, const Sched &hs, const DecBuffers &b, int j, cudaStream_t st, long long *launches)
{
	const int c = hs.chan[j], l = hs.level[j], p = hs.plane[j];
	const int ntile = g.ntile[l];
	const size_t rank_words = (size_t)g.G[l] + 4;
	CUDA_OK(cudaMemsetAsync(b.ones_rank, 0, rank_words * 4, st));
	CUDA_OK(cudaMemsetAsync(b.sign_rank, 0, rank_words * 4, st));
	u32 *sig = b.sig + (size_t)c * g.GT + g.gbase[l];
	u32 *plane_words = b.bs + hs.bsbase[c] + (long long)p * g.GT + g.gbase[l];
	u32 *sign_words = b.bs + hs.bsbase[c] + (long long)hs.planes[c] * g.GT + g.gbase[l];
	dec_prep_kernel<<<ntile, TG, 0, st>>>(g, l, sig, b.mem_pref);
	dec_parse_kernel<<<1, PT, 0, st>>>(b.state, b.stream, b.mem_pref, b.ref_pref, ntile, b.ones_rank, b.sign_rank, c, l);
	dec_deposit_kernel<<<ntile, TG, 0, st>>>(g, l, j + 1, plane_words, sign_words, sig, b.ref_pref, b.ones_rank,
	                                         b.sign_rank, b.stream, b.state);
	*launches += 3;
	CUDA_OK(cudaGetLastError());
	return 0;
}
