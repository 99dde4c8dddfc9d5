// coder_dec.cu -- the bit-plane decoder of decode.c:67-100,187-243 + rle.h + vli.h + bits.h on the GPU.
//
// Chunks (channel, level, plane) are decoded in schedule order; within a chunk:
//   prep     per tile of 256 groups: how many coefficients are still insignificant (members of the
//            significance pass) and how many are already significant (refinement bits)   [dec_prep_kernel]
//   parse    the significance pass is a chain of [adaptive-Rice run][sign] tokens.  One CTA walks the stream
//            in windows of 1024 x 64 bits: every thread parses speculatively from the start of its 64-bit
//            slice, then entries are corrected to the predecessor's exit until nothing changes (the parses
//            re-synchronise after a few tokens), a scan of the run lengths turns tokens into member ranks,
//            and ones / signs are set in rank space.  EOF, the run carried across chunks (rle.h:66-77) and the
//            phantom one before refinement bits (rle.h:91-103) follow the reference exactly [dec_parse_kernel]
//   deposit  rank-space bits are expanded into the insignificant positions of each group (software pdep),
//            refinement bits are taken straight from the stream, significance is updated [dec_deposit_kernel]
#include "coder.cuh"

namespace {

constexpr int TG = DWT_TILE_GROUPS;
constexpr int PT = 1024;        // parse threads
constexpr int SLICE = 64;       // stream bits per thread and window
constexpr int DEAD = 255;

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

// speculative run over one slice: parse [VLI][sign] tokens that start before `lim`
__device__ __forceinline__ void spec_run(const u32 *__restrict__ s, u64 end_bits, u64 lim, u64 &pos, int &k, u64 &csum)
{
	csum = 0;
	while (k != DEAD && pos < lim) {
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

__global__ void __launch_bounds__(PT) dec_parse_kernel(DecState *st, const u32 *__restrict__ stream,
                                                        const u32 *__restrict__ tile_sums, u32 *tile_base, int ntile,
                                                        u32 *ones_rank, u32 *sign_rank, int chan, int level)
{
	__shared__ u64 ws[32];
	__shared__ u64 x_pos[PT];
	__shared__ unsigned char x_k[PT];
	__shared__ int winner;
	__shared__ u64 f_pos;     // final state written by the winning thread
	__shared__ int f_k, f_event;
	__shared__ u32 f_pending;
	const int tid = threadIdx.x;
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
			winner = PT;
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
				if (tid == 0 && ((peek64(stream, bitpos)) & 1ull))
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
		const u64 Rrem = R - r0;
		const u64 sub_lo = bitpos + (u64)tid * SLICE, sub_hi = sub_lo + SLICE;
		u64 e_pos = tid == 0 ? bitpos : sub_lo; // entry
		int e_k = tid == 0 ? order : 0;
		u64 xp = e_pos, csum = 0;
		int xk = e_k;
		bool dirty = true;
		for (;;) {
			if (dirty) {
				xp = e_pos;
				xk = e_k;
				spec_run(stream, end_bits, sub_hi, xp, xk, csum);
			}
			x_pos[tid] = xp;
			x_k[tid] = (unsigned char)xk;
			__syncthreads();
			if (tid > 0) {
				u64 np = x_pos[tid - 1];
				int nk = x_k[tid - 1];
				dirty = np != e_pos || nk != e_k;
				e_pos = np;
				e_k = nk;
			} else {
				dirty = false;
			}
			if (!__syncthreads_or(dirty))
				break;
		}
		// member ranks: exclusive scan of the members consumed per slice
		u64 total;
		u64 cum = block_exscan_u64(csum, ws, &total);
		// final walk with output
		{
			u64 pos = e_pos;
			int k = e_k;
			int ev = EV_NONE;
			u64 ev_pos = 0;
			int ev_k = 0;
			u32 ev_pending = 0;
			if (k == DEAD && (tid == 0 || x_k[tid - 1] == DEAD) && tid == 0)
				ev = EV_STOP;
			while (k != DEAD && pos < sub_hi) {
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
						ev = EV_STOP; // sign bit beyond EOF
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
		// no end inside this window: continue behind the last slice
		bitpos = x_pos[PT - 1];
		order = x_k[PT - 1];
		r0 += total;
		__syncthreads();
		if (order == DEAD) {
			stop = true;
			break;
		}
	}

	// ---- refinement pass (raw bits) and bookkeeping, decode.c:89-98,206,223,240
	if (tid == 0) {
		bool complete = !stop;
		u64 ref_pos = bitpos;
		if (!stop && nref > 0) {
			if (pending > 1) {
				stop = true; // rle.h:98-99: a pending run must end exactly here
				complete = false;
			} else {
				pending = 0; // the phantom one
				if (bitpos + nref > end_bits) {
					stop = true;
					complete = false;
				} else {
					bitpos += nref;
				}
			}
		}
		st->ref_bitpos = ref_pos;
		st->ref_valid = (!stop || ref_pos < end_bits) && nref > 0 && (complete || bitpos == ref_pos) ? 1 : 0;
		if (!complete && nref > 0 && pending <= 1 && ref_pos <= end_bits)
			st->ref_valid = 1; // partial refinement: bits before EOF are kept, the deposit clips at end_bits
		if (stop && (r0 < R || pending > 1))
			st->ref_valid = 0; // the refinement pass was never reached
		st->bitpos = bitpos;
		st->order = order == DEAD ? 0 : order;
		st->pending = pending;
		st->stopped = stop ? 1 : 0;
		st->chunk_done += 1;
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
