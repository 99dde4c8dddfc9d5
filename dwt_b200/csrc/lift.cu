// lift.cu -- reversible colour transform + integer CDF 5/3 lifting, forward and inverse, for sm_100a.
//
// What it computes (bit-exact with the reference):
//   colour   image.h:53-65 (forward), image.h:34-51 (inverse, clamps first), pnm.h:108 (final clamp)
//   1-D      cdf53.h:9-34 / cdf53.h:36-61, including the asymmetric boundary rules and C's truncating
//            division
//   2-D      encode.c:16-30: rows then columns per level; decode.c:16-30: columns then rows
//
// How (streaming strips, no intermediate round trip through HBM):
//   A warp owns a strip of 64 input columns (60 produce output, 2+2 are halo) and walks down a row
//   segment.  Each lane holds one (even, odd) column pair.  The horizontal lifting step of a row takes the
//   two neighbour taps with warp shuffles; the vertical step is a sliding window held in registers, so a
//   sample is read from HBM once per level (plus the halo overlap that L2 serves) and every global access
//   of a warp is one contiguous row segment.
#include "lift.cuh"

#include <atomic>
#include <type_traits>

namespace {

constexpr int STRIP_OUT = 120; // output columns per warp: 30 lanes x 4 columns (one halo lane on each side)
constexpr int WARPS_PER_BLOCK = 4;
constexpr u32 FULLMASK = 0xffffffffu;

__device__ __forceinline__ int clampi(int x, int a, int b)
{
	return min(max(x, a), b);
}

__device__ __forceinline__ int clamp_u8(int x) // clamp to 0..255 in one instruction (VIMNMX with relu)
{
	return __vimin_s32_relu(x, 255);
}

template <int MODE>
struct FwdTraits {
	static constexpr int NC = MODE == 0 ? 3 : 1;
};

// four consecutive columns of one row as they sit in memory: 12 bytes of RGB, 4 bytes of gray, or 4 ints
template <int MODE>
struct Raw {
	u32 w[MODE == 0 ? 3 : (MODE == 1 ? 1 : 4)];
};

__device__ __forceinline__ void store2(int *q, int a, int b, bool vec)
{
	if (vec) {
		*reinterpret_cast<int2 *>(q) = make_int2(a, b);
	} else {
		q[0] = a;
		q[1] = b;
	}
}

// ------------------------------------------------------------------------------------------------ forward

// columns x .. x+3 of row y (clamped to the image); `fast`: all four exist and the row is 4-byte / 16-byte aligned
template <int MODE, bool ROW_INSIDE = false>
__device__ __forceinline__ void fwd_load(const LiftLevel &p, int ch, int y, int x, bool fast, Raw<MODE> &r)
{
	const int cy = ROW_INSIDE ? y : clampi(y, 0, p.H - 1);
	if constexpr (MODE == 0) {
		const uint8_t *row = (const uint8_t *)p.in + (size_t)cy * p.in_pitch * 3;
		if (fast) {
			const u32 *q = reinterpret_cast<const u32 *>(row + (size_t)x * 3);
			r.w[0] = __ldg(q);
			r.w[1] = __ldg(q + 1);
			r.w[2] = __ldg(q + 2);
		} else {
			r.w[0] = r.w[1] = r.w[2] = 0;
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const uint8_t *q = row + (size_t)clampi(x + k, 0, p.W - 1) * 3;
#pragma unroll
				for (int b = 0; b < 3; ++b) {
					const int n = 3 * k + b;
					r.w[n >> 2] |= (u32)__ldg(q + b) << (8 * (n & 3));
				}
			}
		}
	} else if constexpr (MODE == 1) {
		const uint8_t *row = (const uint8_t *)p.in + (size_t)cy * p.in_pitch;
		if (fast) {
			r.w[0] = __ldg(reinterpret_cast<const u32 *>(row + x));
		} else {
			r.w[0] = 0;
#pragma unroll
			for (int k = 0; k < 4; ++k)
				r.w[0] |= (u32)__ldg(row + clampi(x + k, 0, p.W - 1)) << (8 * k);
		}
	} else {
		const int *row = (const int *)p.in + (size_t)ch * p.in_chan_stride + (size_t)cy * p.in_pitch;
		if (fast) {
			const int4 v = __ldg(reinterpret_cast<const int4 *>(row + x));
			r.w[0] = (u32)v.x;
			r.w[1] = (u32)v.y;
			r.w[2] = (u32)v.z;
			r.w[3] = (u32)v.w;
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k)
				r.w[k] = (u32)__ldg(row + clampi(x + k, 0, p.W - 1));
		}
	}
}

// raw columns -> samples v[channel][column]; the colour transform image.h:58-61 is fused for MODE 0
template <int MODE>
__device__ __forceinline__ void fwd_expand(const Raw<MODE> &r, int (&v)[FwdTraits<MODE>::NC][4])
{
	if constexpr (MODE == 0) {
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			int R = (int)((r.w[(3 * k) >> 2] >> (8 * ((3 * k) & 3))) & 255u);
			int G = (int)((r.w[(3 * k + 1) >> 2] >> (8 * ((3 * k + 1) & 3))) & 255u);
			int B = (int)((r.w[(3 * k + 2) >> 2] >> (8 * ((3 * k + 2) & 3))) & 255u);
			// hide the 8-bit range from the compiler: it would narrow the whole transform to 16-bit arithmetic,
			// which costs a mask and a sign extension per operation on this machine
			asm("" : "+r"(R));
			asm("" : "+r"(G));
			asm("" : "+r"(B));
			const int U = R - B, T = B + U / 2, V = G - T;
			v[0][k] = T + V / 2;
			v[1][k] = U;
			v[2][k] = V;
		}
	} else if constexpr (MODE == 1) {
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			int t = (int)((r.w[0] >> (8 * k)) & 255u);
			asm("" : "+r"(t));
			v[0][k] = t;
		}
	} else {
#pragma unroll
		for (int k = 0; k < 4; ++k)
			v[0][k] = (int)r.w[k];
	}
}

struct HPred { // horizontal boundary rules of this lane's four columns (cdf53.h:12-23 / 49-60), row independent
	bool d1_int, d3_int, s0_first, s0_upd, s2_upd;
};

// horizontal lifting of one row: in v[c] = X[x..x+3]; out v[c] = s[x], d[x+1], s[x+2], d[x+3]
template <int NC>
__device__ __forceinline__ void fwd_hlift(int (&v)[NC][4], const HPred &hp)
{
#pragma unroll
	for (int c = 0; c < NC; ++c) {
		const int X0 = v[c][0], X1 = v[c][1], X2 = v[c][2], X3 = v[c][3];
		const int X4 = __shfl_down_sync(FULLMASK, X0, 1);
		const int d1 = hp.d1_int ? X1 - (X0 + X2) / 2 : X1 - X0; // x+1 == N-1 (N even); garbage beyond the row
		const int d3 = hp.d3_int ? X3 - (X2 + X4) / 2 : X3 - X2;
		const int dm1 = __shfl_up_sync(FULLMASK, d3, 1);
		// x == 0: first sample; x == N-1 (N odd): the tail even sample is not updated (cdf53.h:21)
		const int s0 = hp.s0_first ? X0 + d1 / 2 : (hp.s0_upd ? X0 + (dm1 + d1) / 4 : X0);
		const int s2 = hp.s2_upd ? X2 + (d1 + d3) / 4 : X2;
		v[c][0] = s0;
		v[c][1] = d1;
		v[c][2] = s2;
		v[c][3] = d3;
	}
}

// HI: the whole strip is horizontally interior and every buffer is aligned for vector access, so the boundary
// rules, clamps and alignment fall-backs are compiled out (all but the two edge strips of a level take this path)
template <int MODE, bool HI>
__device__ __forceinline__ void fwd_body(const LiftLevel &p, const int y0, const int y1, const int strip, const int ch,
                                         const int lane)
{
	constexpr int NC = FwdTraits<MODE>::NC;
	const int W = p.W, H = p.H;
	const int x = strip * STRIP_OUT - 4 + 4 * lane; // this lane's first column (a multiple of 4)
	const int w2 = (W + 1) / 2, h2 = (H + 1) / 2;
	const bool vs = lane >= 1 && lane <= 30 && (HI || x < W); // this lane stores
	const bool full = vs && (HI || x + 3 < W);                // ... all four columns
	const bool aligned_in = MODE == 2 ? ((p.in_pitch | (int)p.in_chan_stride) & 3) == 0 : (p.in_pitch & 3) == 0;
	const bool fast = HI || (aligned_in && x >= 0 && x + 3 < W);
	const bool v_ll = HI || ((p.out_pitch | (int)p.out_chan_stride) & 1) == 0;
	const bool v_lh = HI || ((p.pyr_pitch | (int)p.pyr_chan_stride) & 1) == 0;
	const bool v_hl = HI || (v_lh && (w2 & 1) == 0);
	HPred hp;
	hp.d1_int = HI || x + 1 < W - 1;
	hp.d3_int = HI || x + 3 < W - 1;
	hp.s0_first = !HI && x == 0;
	hp.s0_upd = HI || x + 1 <= W - 1;
	hp.s2_upd = HI || x + 3 <= W - 1;

	// vertical window: e = row-lifted current even row, dp = vertical detail of the odd row above
	int e[NC][4], dp[NC][4];
	Raw<MODE> raw;
	fwd_load<MODE>(p, ch, y0, x, fast, raw);
	fwd_expand<MODE>(raw, e);
	fwd_hlift<NC>(e, hp);
	if (y0 > 0) {
		int m2[NC][4], m1[NC][4];
		fwd_load<MODE>(p, ch, y0 - 2, x, fast, raw);
		fwd_expand<MODE>(raw, m2);
		fwd_hlift<NC>(m2, hp);
		fwd_load<MODE>(p, ch, y0 - 1, x, fast, raw);
		fwd_expand<MODE>(raw, m1);
		fwd_hlift<NC>(m1, hp);
#pragma unroll
		for (int c = 0; c < NC; ++c)
#pragma unroll
			for (int k = 0; k < 4; ++k)
				dp[c][k] = m1[c][k] - (m2[c][k] + e[c][k]) / 2; // y0-1 is an interior odd row
	} else {
#pragma unroll
		for (int c = 0; c < NC; ++c)
#pragma unroll
			for (int k = 0; k < 4; ++k)
				dp[c][k] = 0;
	}
	int mx[NC];
#pragma unroll
	for (int c = 0; c < NC; ++c)
		mx[c] = 0;

	const int xl = x >> 1, xh = w2 + (x >> 1);
	// raw rows are fetched two iterations ahead (r+1, r+2 in no / ne, r+3, r+4 in no2 / ne2): one iteration of
	// distance left the first use of a row waiting on DRAM for half of all stall samples
	Raw<MODE> no, ne, no2, ne2;
	fwd_load<MODE>(p, ch, y0 + 1, x, fast, no);
	fwd_load<MODE>(p, ch, y0 + 2, x, fast, ne);
	fwd_load<MODE>(p, ch, y0 + 3, x, fast, no2);
	fwd_load<MODE>(p, ch, y0 + 4, x, fast, ne2);
	// one row pair.  STEADY (interior strips only): not the first pair, rows r+3 / r+4 exist and are prefetched --
	// no clamp, no boundary rule, no branch in the body; this is what almost every iteration runs
	auto step = [&](const int r, auto steady_tag) {
		constexpr bool STEADY = decltype(steady_tag)::value;
		int o[NC][4], f[NC][4];
		fwd_expand<MODE>(no, o);
		fwd_expand<MODE>(ne, f);
		no = no2;
		ne = ne2;
		if (STEADY || r + 4 < y1) {
			fwd_load<MODE, STEADY>(p, ch, r + 5, x, fast, no2);
			fwd_load<MODE, STEADY>(p, ch, r + 6, x, fast, ne2);
		}
		fwd_hlift<NC>(o, hp);
		fwd_hlift<NC>(f, hp);
		const bool has_odd = STEADY || r + 1 < H;
		const bool v_int = STEADY || (r > 0 && r + 1 < H - 1); // neither the first row pair nor the last odd row
		const size_t lo_row = (size_t)(r >> 1), hi_row = (size_t)(h2 + (r >> 1));
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			const int cc = MODE == 2 ? ch : c;
			int D[4], S[4];
			if (v_int) {
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					D[k] = o[c][k] - (e[c][k] + f[c][k]) / 2;
					S[k] = e[c][k] + (dp[c][k] + D[k]) / 4;
				}
			} else {
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					// r+1 == H-1 (H even): plain difference; unused when r+1 >= H
					D[k] = r + 1 < H - 1 ? o[c][k] - (e[c][k] + f[c][k]) / 2 : o[c][k] - e[c][k];
					// r == 0: first row; r == H-1 (H odd): untouched
					S[k] = r == 0 ? e[c][k] + D[k] / 2 : (has_odd ? e[c][k] + (dp[c][k] + D[k]) / 4 : e[c][k]);
				}
			}
			int *ll = (int *)p.out + (size_t)cc * p.out_chan_stride + lo_row * p.out_pitch;
			int *pl = p.pyr + (size_t)cc * p.pyr_chan_stride + lo_row * p.pyr_pitch;
			int *ph = p.pyr + (size_t)cc * p.pyr_chan_stride + hi_row * p.pyr_pitch;
			if (full) {
				store2(ll + xl, S[0], S[2], v_ll);  // LL
				store2(pl + xh, S[1], S[3], v_hl);  // HL (high-x, low-y)
				int m = max(abs(S[1]), abs(S[3]));
				if (has_odd) {
					store2(ph + xl, D[0], D[2], v_lh); // LH (low-x, high-y)
					store2(ph + xh, D[1], D[3], v_hl); // HH
					m = max(max(m, abs(D[0])), max(abs(D[2]), max(abs(D[1]), abs(D[3]))));
				}
				mx[c] = max(mx[c], m);
			} else if (!HI && vs) {
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					if (x + k >= W)
						continue;
					if (!(k & 1)) {
						ll[xl + (k >> 1)] = S[k];
						if (has_odd) {
							ph[xl + (k >> 1)] = D[k];
							mx[c] = max(mx[c], abs(D[k]));
						}
					} else {
						pl[xh + (k >> 1)] = S[k];
						mx[c] = max(mx[c], abs(S[k]));
						if (has_odd) {
							ph[xh + (k >> 1)] = D[k];
							mx[c] = max(mx[c], abs(D[k]));
						}
					}
				}
			}
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				e[c][k] = f[c][k];
				dp[c][k] = D[k];
			}
		}
	};
#pragma unroll 1
	for (int r = y0; r < y1; r += 2) {
		if (HI && r > 0 && r + 6 < H)
			step(r, std::true_type());
		else
			step(r, std::false_type());
	}
#pragma unroll
	for (int c = 0; c < NC; ++c) {
		const int m = __reduce_max_sync(FULLMASK, mx[c]);
		if (lane == 0 && m > 0)
			atomicMax(p.maxabs + (MODE == 2 ? ch : c), m);
	}
}

// Work items (strip, row segment, channel) are handed out through a counter, strips fastest.  The grid is sized to
// what is resident and every warp keeps taking items until they run out.  The schedule allows a second, smaller
// segment height for the last rows (rss); equal heights measured best, so rss == rs.
struct LiftSched {
	int nstrip, planes;
	int rs, nbig;    // segments of rs rows cover rows [0, nbig * rs)
	int rss, nsmall; // then segments of rss rows
	int fixed;       // 1 = the grid holds a warp per item: no hand-out
};

// programmatic dependent launch: let the next kernel of the stream get resident early / wait for the producer
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int MODE>
__device__ __forceinline__ bool next_item(const LiftLevel &p, const LiftSched &q, int lane, int &strip, int &y0, int &y1, int &ch)
{
	int item = 0;
	if (MODE == 2 && q.fixed) { // the caller starts with y1 < 0; a served item leaves y1 > 0
		if (y1 >= 0)
			return false;
		item = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
	} else {
		if (lane == 0)
			item = atomicAdd(p.work, 1);
		item = __shfl_sync(FULLMASK, item, 0);
	}
	const int big = q.nstrip * q.nbig * q.planes;
	if (item < big) {
		strip = item % q.nstrip;
		const int r = item / q.nstrip;
		y0 = (r % q.nbig) * q.rs;
		y1 = min(y0 + q.rs, min(q.nbig * q.rs, p.H));
		ch = r / q.nbig;
		return true;
	}
	item -= big;
	if (item >= q.nstrip * q.nsmall * q.planes)
		return false;
	strip = item % q.nstrip;
	const int r = item / q.nstrip;
	y0 = q.nbig * q.rs + (r % q.nsmall) * q.rss;
	y1 = min(y0 + q.rss, p.H);
	ch = r / q.nsmall;
	return true;
}

template <int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MODE == 2 ? 6 : 5) lift_fwd_kernel(const __grid_constant__ LiftLevel p,
                                                                            const __grid_constant__ LiftSched q)
{
	const int lane = threadIdx.x & 31;
	const int W = p.W;
	const int w2 = (W + 1) / 2;
	const bool aligned = (MODE == 2 ? ((p.in_pitch | (int)p.in_chan_stride) & 3) == 0 : (p.in_pitch & 3) == 0) &&
	                     ((p.out_pitch | (int)p.out_chan_stride | p.pyr_pitch | (int)p.pyr_chan_stride | w2) & 1) == 0;
	pdl_launch_dependents();
	pdl_wait();
	int strip, y0, y1 = -1, ch;
	while (next_item<MODE>(p, q, lane, strip, y0, y1, ch)) {
		if (aligned && strip > 0 && strip * STRIP_OUT + 124 < W)
			fwd_body<MODE, true>(p, y0, y1, strip, ch, lane);
		else
			fwd_body<MODE, false>(p, y0, y1, strip, ch, lane);
	}
}

// ------------------------------------------------------------------------------------------------ inverse

struct IPred { // horizontal boundary rules of this lane's two column pairs (cdf53.h:49-60)
	bool e0_first, e0_upd, e1_upd, o0_int, o1_int;
};

// horizontal inverse of one row: in s[c][0..1] = s_j, s_j+1 and d[c][0..1] = d_j, d_j+1;
// out v[c] = x[2j], x[2j+1], x[2j+2], x[2j+3]
template <int NC>
__device__ __forceinline__ void inv_hlift(const int (&s)[NC][2], const int (&d)[NC][2], int (&v)[NC][4], const IPred &ip)
{
#pragma unroll
	for (int c = 0; c < NC; ++c) {
		const int dm1 = __shfl_up_sync(FULLMASK, d[c][1], 1); // d_{j-1}
		const int e0 = ip.e0_first ? s[c][0] - d[c][0] / 2 : (ip.e0_upd ? s[c][0] - (dm1 + d[c][0]) / 4 : s[c][0]);
		const int e1 = ip.e1_upd ? s[c][1] - (d[c][0] + d[c][1]) / 4 : s[c][1];
		const int e2 = __shfl_down_sync(FULLMASK, e0, 1); // x[2j+4]
		v[c][0] = e0;
		v[c][1] = ip.o0_int ? d[c][0] + (e0 + e1) / 2 : d[c][0] + e0; // 2j+1 == N-1 (N even)
		v[c][2] = e1;
		v[c][3] = ip.o1_int ? d[c][1] + (e1 + e2) / 2 : d[c][1] + e1;
	}
}

template <int MODE>
__device__ __forceinline__ void inv_store_row(const LiftLevel &p, int ch, int y, int x, bool vs, bool full, bool vec,
                                              const int (&v)[FwdTraits<MODE>::NC][4])
{
	if (!vs)
		return;
	if constexpr (MODE == 2) {
		int *row = (int *)p.out + (size_t)ch * p.out_chan_stride + (size_t)y * p.out_pitch;
		if (full && vec) {
			*reinterpret_cast<int4 *>(row + x) = make_int4(v[0][0], v[0][1], v[0][2], v[0][3]);
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k)
				if (x + k < p.W)
					row[x + k] = v[0][k];
		}
	} else if constexpr (MODE == 1) {
		uint8_t *row = (uint8_t *)p.out + (size_t)y * p.out_pitch;
		u32 w = 0;
#pragma unroll
		for (int k = 0; k < 4; ++k)
			w |= (u32)clamp_u8(v[0][k]) << (8 * k);
		if (full && vec) {
			*reinterpret_cast<u32 *>(row + x) = w;
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k)
				if (x + k < p.W)
					row[x + k] = (uint8_t)(w >> (8 * k));
		}
	} else {
		uint8_t *row = (uint8_t *)p.out + (size_t)y * p.out_pitch * 3;
		u32 w[3] = {0u, 0u, 0u};
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			// image.h:41-50 then pnm.h:108
			int Y = clamp_u8(v[0][k]), U = clampi(v[1][k], -255, 255), V = clampi(v[2][k], -255, 255);
			asm("" : "+r"(Y)); // keep the compiler from narrowing the transform to 16-bit arithmetic (see fwd_expand)
			asm("" : "+r"(U));
			asm("" : "+r"(V));
			const int T = Y - V / 2, G = V + T, B = T - U / 2, R = B + U;
			const u32 px[3] = {(u32)clamp_u8(R), (u32)clamp_u8(G), (u32)clamp_u8(B)};
#pragma unroll
			for (int b = 0; b < 3; ++b) {
				const int n = 3 * k + b;
				w[n >> 2] |= px[b] << (8 * (n & 3));
			}
		}
		if (full && vec) {
			u32 *q = reinterpret_cast<u32 *>(row + (size_t)x * 3);
			q[0] = w[0];
			q[1] = w[1];
			q[2] = w[2];
		} else {
#pragma unroll
			for (int n = 0; n < 12; ++n)
				if (x + n / 3 < p.W)
					row[(size_t)x * 3 + n] = (uint8_t)(w[n >> 2] >> (8 * (n & 3)));
		}
	}
}

template <int MODE, bool HI>
__device__ __forceinline__ void inv_body(const LiftLevel &p, const int y0, const int y1, const int strip, const int chz,
                                         const int lane)
{
	constexpr int NC = FwdTraits<MODE>::NC;
	const int W = p.W, H = p.H;
	const int x = strip * STRIP_OUT - 4 + 4 * lane; // output columns x .. x+3 = column pairs j, j+1
	const int j = x >> 1;
	const int w2 = (W + 1) / 2, h2 = (H + 1) / 2, wd = W / 2, hd = H / 2;
	const bool vs = lane >= 1 && lane <= 30 && (HI || x < W);
	const bool full = vs && (HI || x + 3 < W);
	// clamped band columns: low-x pair and high-x pair (Mallat position)
	const int cl0 = HI ? j : clampi(j, 0, w2 - 1), cl1 = HI ? j + 1 : clampi(j + 1, 0, w2 - 1);
	const int chh0 = w2 + (HI ? j : clampi(j, 0, max(wd - 1, 0))), chh1 = w2 + (HI ? j + 1 : clampi(j + 1, 0, max(wd - 1, 0)));
	const bool in_range = HI || (j >= 0 && j + 1 < wd);
	const bool v_ll = HI || (in_range && ((p.in_pitch | (int)p.in_chan_stride) & 1) == 0);
	const bool v_lh = HI || (in_range && ((p.pyr_pitch | (int)p.pyr_chan_stride) & 1) == 0);
	const bool v_hl = HI || (v_lh && (w2 & 1) == 0);
	const bool vec_out = HI || (MODE == 2 ? ((p.out_pitch | (int)p.out_chan_stride) & 3) == 0 : (p.out_pitch & 3) == 0);
	IPred ip;
	ip.e0_first = !HI && j == 0;
	ip.e0_upd = HI || x + 1 <= W - 1;
	ip.e1_upd = HI || x + 3 <= W - 1;
	ip.o0_int = HI || x + 1 < W - 1;
	ip.o1_int = HI || x + 3 < W - 1;

	// the four band samples of pair row m: S (low-y) and D (high-y), each for the low-x pair and the high-x pair
	auto load = [&](int m, int(&S_s)[NC][2], int(&S_d)[NC][2], int(&D_s)[NC][2], int(&D_d)[NC][2]) {
		const int ms = clampi(m, 0, h2 - 1), md = h2 + clampi(m, 0, max(hd - 1, 0));
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			const int cc = MODE == 2 ? chz : c;
			const int *ll = (const int *)p.in + (size_t)cc * p.in_chan_stride + (size_t)ms * p.in_pitch;
			const int *ps = p.pyr + (size_t)cc * p.pyr_chan_stride + (size_t)ms * p.pyr_pitch;
			const int *pd = p.pyr + (size_t)cc * p.pyr_chan_stride + (size_t)md * p.pyr_pitch;
			if (v_ll) {
				const int2 t = __ldg(reinterpret_cast<const int2 *>(ll + cl0));
				S_s[c][0] = t.x;
				S_s[c][1] = t.y;
			} else {
				S_s[c][0] = __ldg(ll + cl0);
				S_s[c][1] = __ldg(ll + cl1);
			}
			if (v_hl) {
				const int2 t = __ldg(reinterpret_cast<const int2 *>(ps + chh0));
				const int2 u = __ldg(reinterpret_cast<const int2 *>(pd + chh0));
				S_d[c][0] = t.x;
				S_d[c][1] = t.y;
				D_d[c][0] = u.x;
				D_d[c][1] = u.y;
			} else {
				S_d[c][0] = __ldg(ps + chh0);
				S_d[c][1] = __ldg(ps + chh1);
				D_d[c][0] = __ldg(pd + chh0);
				D_d[c][1] = __ldg(pd + chh1);
			}
			if (v_lh) {
				const int2 t = __ldg(reinterpret_cast<const int2 *>(pd + cl0));
				D_s[c][0] = t.x;
				D_s[c][1] = t.y;
			} else {
				D_s[c][0] = __ldg(pd + cl0);
				D_s[c][1] = __ldg(pd + cl1);
			}
		}
	};

	const int m0 = y0 >> 1;
	int Dm_s[NC][2], Dm_d[NC][2], Ep_s[NC][2], Ep_d[NC][2];
#pragma unroll
	for (int c = 0; c < NC; ++c)
#pragma unroll
		for (int k = 0; k < 2; ++k)
			Dm_s[c][k] = Dm_d[c][k] = Ep_s[c][k] = Ep_d[c][k] = 0;
	// the pair row above the segment and its first pair row are requested together (rows are clamped, so the load for the
	// first segment is harmless and its result dropped): behind a branch the compiler sinks the first load to its use and
	// the item start waits for memory twice
	int nS_s[NC][2], nS_d[NC][2], nD_s[NC][2], nD_d[NC][2];
	{
		int t0[NC][2], t1[NC][2], pD_s[NC][2], pD_d[NC][2];
		load(m0 - 1, t0, t1, pD_s, pD_d);
		load(m0, nS_s, nS_d, nD_s, nD_d);
#pragma unroll
		for (int c = 0; c < NC; ++c)
#pragma unroll
			for (int k = 0; k < 2; ++k) {
				Dm_s[c][k] = m0 > 0 ? pD_s[c][k] : 0;
				Dm_d[c][k] = m0 > 0 ? pD_d[c][k] : 0;
			}
	}
	for (int m = m0; 2 * m - 1 < y1; ++m) {
		int S_s[NC][2], S_d[NC][2], D_s[NC][2], D_d[NC][2];
#pragma unroll
		for (int c = 0; c < NC; ++c)
#pragma unroll
			for (int k = 0; k < 2; ++k) {
				S_s[c][k] = nS_s[c][k];
				S_d[c][k] = nS_d[c][k];
				D_s[c][k] = nD_s[c][k];
				D_d[c][k] = nD_d[c][k];
			}
		if (2 * m + 1 < y1)
			load(m + 1, nS_s, nS_d, nD_s, nD_d);
		const bool hasE = 2 * m < H;
		int E_s[NC][2], E_d[NC][2], O_s[NC][2], O_d[NC][2];
#pragma unroll
		for (int c = 0; c < NC; ++c)
#pragma unroll
			for (int k = 0; k < 2; ++k) {
				// vertical inverse (cdf53.h:49-60 along columns)
				if (m == 0) {
					E_s[c][k] = S_s[c][k] - D_s[c][k] / 2;
					E_d[c][k] = S_d[c][k] - D_d[c][k] / 2;
				} else if (2 * m + 1 <= H - 1) {
					E_s[c][k] = S_s[c][k] - (Dm_s[c][k] + D_s[c][k]) / 4;
					E_d[c][k] = S_d[c][k] - (Dm_d[c][k] + D_d[c][k]) / 4;
				} else { // 2m == H-1 (H odd)
					E_s[c][k] = S_s[c][k];
					E_d[c][k] = S_d[c][k];
				}
				if (hasE) {
					O_s[c][k] = Dm_s[c][k] + (Ep_s[c][k] + E_s[c][k]) / 2;
					O_d[c][k] = Dm_d[c][k] + (Ep_d[c][k] + E_d[c][k]) / 2;
				} else { // 2m-1 == H-1 (H even)
					O_s[c][k] = Dm_s[c][k] + Ep_s[c][k];
					O_d[c][k] = Dm_d[c][k] + Ep_d[c][k];
				}
			}
		int v[NC][4];
		if (m > m0) { // odd row 2m-1 (uniform branch: shuffles stay converged)
			inv_hlift<NC>(O_s, O_d, v, ip);
			inv_store_row<MODE>(p, chz, 2 * m - 1, x, vs, full, vec_out, v);
		}
		if (hasE && 2 * m < y1) { // even row 2m
			inv_hlift<NC>(E_s, E_d, v, ip);
			inv_store_row<MODE>(p, chz, 2 * m, x, vs, full, vec_out, v);
		}
#pragma unroll
		for (int c = 0; c < NC; ++c)
#pragma unroll
			for (int k = 0; k < 2; ++k) {
				Ep_s[c][k] = E_s[c][k];
				Ep_d[c][k] = E_d[c][k];
				Dm_s[c][k] = D_s[c][k];
				Dm_d[c][k] = D_d[c][k];
			}
	}
}

template <int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) lift_inv_kernel(const __grid_constant__ LiftLevel p,
                                                                         const __grid_constant__ LiftSched q)
{
	const int lane = threadIdx.x & 31;
	const int W = p.W;
	const int w2 = (W + 1) / 2;
	const bool aligned = ((p.in_pitch | (int)p.in_chan_stride | p.pyr_pitch | (int)p.pyr_chan_stride | w2) & 1) == 0 &&
	                     (MODE == 2 ? ((p.out_pitch | (int)p.out_chan_stride) & 3) == 0 : (p.out_pitch & 3) == 0);
	pdl_launch_dependents();
	pdl_wait();
	int strip, y0, y1 = -1, ch;
	while (next_item<MODE>(p, q, lane, strip, y0, y1, ch)) {
		if (aligned && strip > 0 && strip * STRIP_OUT + 124 < W)
			inv_body<MODE, true>(p, y0, y1, strip, ch, lane);
		else
			inv_body<MODE, false>(p, y0, y1, strip, ch, lane);
	}
}

// ------------------------------------------------------------------------------------------------ fused tail levels
//
// Below ~256 columns a level is launch-latency bound, so the remaining levels of a channel run in ONE CTA with
// the LL image resident in shared memory: rows in place (a warp owns a row: read, lift, write back deinterleaved),
// then columns into a second buffer that only has to keep the new LL quadrant -- the detail bands go straight
// to the pyramid.  The inverse mirrors it.  Same arithmetic and order as the per-level kernels (encode.c:16-30).

template <typename F>
__device__ __forceinline__ int fwd_d_at(F X, int i, int N) // detail at odd i (cdf53.h:12-16)
{
	return i < N - 1 ? X(i) - (X(i - 1) + X(i + 1)) / 2 : X(i) - X(i - 1);
}

template <typename F>
__device__ __forceinline__ int fwd_s_at(F X, int i, int N) // smooth at even i (cdf53.h:17-23)
{
	if (i == 0)
		return X(0) + fwd_d_at(X, 1, N) / 2;
	if (i + 1 <= N - 1)
		return X(i) + (fwd_d_at(X, i - 1, N) + fwd_d_at(X, i + 1, N)) / 4;
	return X(i);
}

template <typename FS, typename FD>
__device__ __forceinline__ int inv_e_at(FS S, FD D, int j, int N) // x[2j] (cdf53.h:49-55)
{
	if (j == 0)
		return S(0) - D(0) / 2;
	if (2 * j + 1 <= N - 1)
		return S(j) - (D(j - 1) + D(j)) / 4;
	return S(j);
}

template <typename FS, typename FD>
__device__ __forceinline__ int inv_o_at(FS S, FD D, int j, int N) // x[2j+1] (cdf53.h:56-60)
{
	const int e = inv_e_at(S, D, j, N);
	return 2 * j + 1 < N - 1 ? D(j) + (e + inv_e_at(S, D, j + 1, N)) / 2 : D(j) + e;
}

constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_MAX_W = 64;       // wider levels are faster as strip launches (short walks, dependent launches)
constexpr int TAIL_MAX_INTS = 56000; // shared memory budget (both buffers), in ints

__global__ void __launch_bounds__(TAIL_THREADS) lift_tail_fwd_kernel(LiftTail t)
{
	pdl_launch_dependents();
	pdl_wait();
	extern __shared__ int sm[];
	__shared__ int smax;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int ch = blockIdx.x;
	int w = t.W, h = t.H;
	int *src = sm, *dst = sm + t.W * t.H;
	int ps = t.W;
	if (tid == 0)
		smax = 0;
	{
		const int *in = t.ll_in + (size_t)ch * t.in_chan_stride;
		for (int i = tid; i < w * h; i += TAIL_THREADS) {
			const int y = i / w, x = i - y * w;
			src[i] = in[(size_t)y * t.in_pitch + x];
		}
	}
	__syncthreads();
	int *pyr = t.pyr + (size_t)ch * t.pyr_chan_stride;
	int mx = 0;
	for (int lev = 0; lev < t.nlev; ++lev) {
		const int w2 = (w + 1) / 2, h2 = (h + 1) / 2;
		for (int y = wid; y < h; y += TAIL_THREADS / 32) { // rows, in place
			int *row = src + y * ps;
			auto X = [&](int i) { return row[i]; };
			int sv[4], dv[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const int i = lane + 32 * q;
				sv[q] = 2 * i < w ? fwd_s_at(X, 2 * i, w) : 0;
				dv[q] = 2 * i + 1 < w ? fwd_d_at(X, 2 * i + 1, w) : 0;
			}
			__syncwarp();
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const int i = lane + 32 * q;
				if (2 * i < w)
					row[i] = sv[q];
				if (2 * i + 1 < w)
					row[w2 + i] = dv[q];
			}
		}
		__syncthreads();
		const int pd = w2;
		for (int i = tid; i < h2 * w; i += TAIL_THREADS) { // columns: LL -> dst, everything else -> pyramid
			const int j = i / w, x = i - j * w;
			auto X = [&](int y) { return src[y * ps + x]; };
			const int S = fwd_s_at(X, 2 * j, h);
			if (x < w2) {
				dst[j * pd + x] = S;
			} else {
				pyr[(size_t)j * t.pyr_pitch + x] = S;
				mx = max(mx, abs(S));
			}
			if (2 * j + 1 < h) {
				const int D = fwd_d_at(X, 2 * j + 1, h);
				pyr[(size_t)(h2 + j) * t.pyr_pitch + x] = D;
				mx = max(mx, abs(D));
			}
		}
		__syncthreads();
		int *tmp = src;
		src = dst;
		dst = tmp;
		ps = pd;
		w = w2;
		h = h2;
	}
	{
		int *out = t.ll_out + (size_t)ch * t.out_chan_stride;
		for (int i = tid; i < w * h; i += TAIL_THREADS) {
			const int y = i / w, x = i - y * w;
			out[(size_t)y * t.out_pitch + x] = src[y * ps + x];
		}
	}
	mx = __reduce_max_sync(FULLMASK, mx);
	if (lane == 0 && mx > 0)
		atomicMax(&smax, mx);
	__syncthreads();
	if (tid == 0 && smax > 0)
		atomicMax(t.maxabs + ch, smax);
}

// inverse: t.W x t.H is the size of the FINEST fused level; t.ll_in holds the root, t.ll_out receives W x H
__global__ void __launch_bounds__(TAIL_THREADS) lift_tail_inv_kernel(LiftTail t)
{
	pdl_launch_dependents();
	pdl_wait();
	extern __shared__ int sm[];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int ch = blockIdx.x;
	// level sizes, coarse to fine: ws[0] x hs[0] is the root
	int ws[DWT_MAX_LEVELS + 1], hs[DWT_MAX_LEVELS + 1];
	ws[t.nlev] = t.W;
	hs[t.nlev] = t.H;
	for (int l = t.nlev; l > 0; --l) {
		ws[l - 1] = (ws[l] + 1) / 2;
		hs[l - 1] = (hs[l] + 1) / 2;
	}
	int *big = sm, *small = sm + t.W * t.H;
	// buffers alternate so that the finest level lands in `big`
	int *cur = (t.nlev & 1) ? small : big;
	int pc = ws[0];
	{
		const int *in = t.ll_in + (size_t)ch * t.in_chan_stride;
		const int w = ws[0], h = hs[0];
		for (int i = tid; i < w * h; i += TAIL_THREADS) {
			const int y = i / w, x = i - y * w;
			cur[y * pc + x] = in[(size_t)y * t.in_pitch + x];
		}
	}
	__syncthreads();
	const int *pyr = t.pyr + (size_t)ch * t.pyr_chan_stride;
	for (int lev = 1; lev <= t.nlev; ++lev) {
		const int w = ws[lev], h = hs[lev], w2 = ws[lev - 1], h2 = hs[lev - 1];
		int *nxt = cur == big ? small : big;
		const int pn = w;
		for (int i = tid; i < h2 * w; i += TAIL_THREADS) { // columns (decode.c:21): pair row j of column x
			const int j = i / w, x = i - j * w;
			auto S = [&](int m) { return x < w2 ? cur[m * pc + x] : __ldg(pyr + (size_t)m * t.pyr_pitch + x); };
			auto D = [&](int m) { return __ldg(pyr + (size_t)(h2 + m) * t.pyr_pitch + x); };
			nxt[(2 * j) * pn + x] = inv_e_at(S, D, j, h);
			if (2 * j + 1 < h)
				nxt[(2 * j + 1) * pn + x] = inv_o_at(S, D, j, h);
		}
		__syncthreads();
		for (int y = wid; y < h; y += TAIL_THREADS / 32) { // rows, in place (decode.c:25-29)
			int *row = nxt + y * pn;
			auto S = [&](int m) { return row[m]; };
			auto D = [&](int m) { return row[w2 + m]; };
			int ev[4], ov[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const int jj = lane + 32 * q;
				ev[q] = 2 * jj < w ? inv_e_at(S, D, jj, w) : 0;
				ov[q] = 2 * jj + 1 < w ? inv_o_at(S, D, jj, w) : 0;
			}
			__syncwarp();
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const int jj = lane + 32 * q;
				if (2 * jj < w)
					row[2 * jj] = ev[q];
				if (2 * jj + 1 < w)
					row[2 * jj + 1] = ov[q];
			}
		}
		__syncthreads();
		cur = nxt;
		pc = pn;
	}
	{
		int *out = t.ll_out + (size_t)ch * t.out_chan_stride;
		const int w = t.W, h = t.H;
		for (int i = tid; i < w * h; i += TAIL_THREADS) {
			const int y = i / w, x = i - y * w;
			out[(size_t)y * t.out_pitch + x] = cur[y * pc + x];
		}
	}
}

int resident_blocks()
{
	return dwt_device_sms() * 8;
}

int pick_rows(int W, int H, int planes)
{
	// aim for >= ~12k items (four per resident warp: measured best with dependent launches); rows per segment even
	long long strips = (W + STRIP_OUT - 1) / STRIP_OUT;
	int rs = 64;
	while (rs > 8 && strips * planes * ((H + rs - 1) / rs) < 12000)
		rs >>= 1;
	// a level small enough to give every item a resident warp is latency bound (a launch plus the dependent L2 round
	// trips of one warp's walk): the shortest walk wins there, the extra halo rows cost nothing that matters
	const long long cap = (long long)resident_blocks() * WARPS_PER_BLOCK;
	while (rs <= 8 && rs > 2 && strips * planes * ((H + rs / 2 - 1) / (rs / 2)) <= cap)
		rs >>= 1;
	return rs;
}

// ------------------------------------------------------------------------------------------------ 1-D entry points

// forward: one thread per (pair index, lane c).  Writes the lifted interleaved samples to `lifted`
// (what the reference leaves in `in`) and the deinterleaved result to `out`.
__global__ void cdf53_1d_kernel(int *out, const int *in, int *lifted, int N, int SO, int SI, int CH)
{
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int half = (N + 1) / 2;
	if (t >= (long long)half * CH)
		return;
	int c = (int)(t % CH), k = (int)(t / CH), i = 2 * k;
	auto X = [&](int j) { return in[(size_t)j * SI + c]; };
	auto detail = [&](int j) { // odd j < N
		return j < N - 1 ? X(j) - (X(j - 1) + X(j + 1)) / 2 : X(j) - X(j - 1);
	};
	int s;
	if (i == 0)
		s = X(0) + detail(1) / 2;
	else if (i + 1 <= N - 1)
		s = X(i) + (detail(i - 1) + detail(i + 1)) / 4;
	else
		s = X(i);
	out[(size_t)k * SO + c] = s;
	lifted[(size_t)i * SI + c] = s;
	if (i + 1 < N) {
		int d = detail(i + 1);
		out[(size_t)(half + k) * SO + c] = d;
		lifted[(size_t)(i + 1) * SI + c] = d;
	}
}

__global__ void icdf53_1d_kernel(int *out, const int *in, int N, int SO, int SI, int CH)
{
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int half = (N + 1) / 2;
	if (t >= (long long)half * CH)
		return;
	int c = (int)(t % CH), k = (int)(t / CH);
	auto S = [&](int j) { return in[(size_t)j * SI + c]; };
	auto D = [&](int j) { return in[(size_t)(half + j) * SI + c]; };
	auto even = [&](int j) { // reconstructed x[2j]
		if (j == 0)
			return S(0) - D(0) / 2;
		if (2 * j + 1 <= N - 1)
			return S(j) - (D(j - 1) + D(j)) / 4;
		return S(j);
	};
	int e = even(k);
	out[(size_t)(2 * k) * SO + c] = e;
	if (2 * k + 1 < N) {
		int o = 2 * k + 1 < N - 1 ? D(k) + (e + even(k + 1)) / 2 : D(k) + e;
		out[(size_t)(2 * k + 1) * SO + c] = o;
	}
}

__global__ void colour_kernel(int *buf, int total, bool inverse)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= total)
		return;
	int *q = buf + (size_t)t * 3;
	if (!inverse) {
		int R = q[0], G = q[1], B = q[2];
		int U = R - B, T = B + U / 2, V = G - T;
		q[0] = T + V / 2;
		q[1] = U;
		q[2] = V;
	} else {
		int Y = clampi(q[0], 0, 255), U = clampi(q[1], -255, 255), V = clampi(q[2], -255, 255);
		int T = Y - V / 2, G = V + T, B = T - U / 2;
		q[0] = B + U;
		q[1] = G;
		q[2] = B;
	}
}

} // namespace

// plain launch, or a programmatic dependent launch behind the previous kernel of the stream
template <typename... P, typename... A>
static cudaError_t launch_chained(void (*k)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int chained, const A &...a)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = chained ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, k, a...);
}


static LiftSched make_sched(const LiftLevel &lv, int planes, long long *items)
{
	LiftSched q;
	q.nstrip = (lv.W + STRIP_OUT - 1) / STRIP_OUT;
	q.planes = planes;
	q.rs = pick_rows(lv.W, lv.H, planes);
	q.rss = q.rs; // a finer last quarter (rs / 4) was measured and lost: the extra halo rows cost more than the idle tail
	const int nseg = (lv.H + q.rs - 1) / q.rs;
	q.nbig = q.rss < q.rs ? nseg * 3 / 4 : nseg; // small levels: one size
	const int rest = lv.H - q.nbig * q.rs;
	q.nsmall = rest > 0 ? (rest + q.rss - 1) / q.rss : 0;
	*items = (long long)q.nstrip * planes * (q.nbig + q.nsmall);
	q.fixed = (*items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK <= resident_blocks();
	return q;
}

// lv.work must point at a zeroed device counter (one per launch)
int lift_forward_level(const LiftLevel &lv, int mode, cudaStream_t st, long long *launches)
{
	const int planes = mode == 2 ? lv.channels : 1;
	long long items;
	const LiftSched q = make_sched(lv, planes, &items);
	const int grid = (int)min((items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, (long long)resident_blocks());
	dim3 block(WARPS_PER_BLOCK * 32);
	if (mode == 0)
		CUDA_OK(launch_chained(lift_fwd_kernel<0>, grid, block, 0, st, lv.chained, lv, q));
	else if (mode == 1)
		CUDA_OK(launch_chained(lift_fwd_kernel<1>, grid, block, 0, st, lv.chained, lv, q));
	else
		CUDA_OK(launch_chained(lift_fwd_kernel<2>, grid, block, 0, st, lv.chained, lv, q));
	if (launches)
		++*launches;
	return 0;
}

int lift_inverse_level(const LiftLevel &lv, int mode, cudaStream_t st, long long *launches)
{
	const int planes = mode == 2 ? lv.channels : 1;
	long long items;
	const LiftSched q = make_sched(lv, planes, &items);
	const int grid = (int)min((items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, (long long)resident_blocks());
	dim3 block(WARPS_PER_BLOCK * 32);
	if (mode == 0)
		CUDA_OK(launch_chained(lift_inv_kernel<0>, grid, block, 0, st, lv.chained, lv, q));
	else if (mode == 1)
		CUDA_OK(launch_chained(lift_inv_kernel<1>, grid, block, 0, st, lv.chained, lv, q));
	else
		CUDA_OK(launch_chained(lift_inv_kernel<2>, grid, block, 0, st, lv.chained, lv, q));
	if (launches)
		++*launches;
	return 0;
}

bool lift_tail_fits(int W, int H)
{
	const long long ints = (long long)W * H + (long long)((W + 1) / 2) * ((H + 1) / 2);
	return W <= TAIL_MAX_W && ints <= TAIL_MAX_INTS;
}

int lift_tail(const LiftTail &t, bool inverse, cudaStream_t st, long long *launches)
{
	static std::atomic<unsigned long long> configured{0}; // one bit per device: the attribute is per device
	const size_t smem = sizeof(int) * ((size_t)t.W * t.H + (size_t)((t.W + 1) / 2) * ((t.H + 1) / 2));
	int dev = 0;
	CUDA_OK(cudaGetDevice(&dev));
	const unsigned long long bit = 1ull << (dev & 63);
	if (!(configured.load(std::memory_order_acquire) & bit)) {
		CUDA_OK(cudaFuncSetAttribute(lift_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                             (int)(TAIL_MAX_INTS * sizeof(int))));
		CUDA_OK(cudaFuncSetAttribute(lift_tail_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                             (int)(TAIL_MAX_INTS * sizeof(int))));
		configured.fetch_or(bit, std::memory_order_release);
	}
	CUDA_OK(launch_chained(inverse ? lift_tail_inv_kernel : lift_tail_fwd_kernel, t.channels, TAIL_THREADS, smem, st, t.chained, t));
	if (launches)
		++*launches;
	return 0;
}

int lift_cdf53_1d(int *d_out, int *d_in, int N, int SO, int SI, int CH, bool inverse, cudaStream_t st)
{
	long long n = (long long)((N + 1) / 2) * CH;
	if (n <= 0)
		return 0;
	int block = 256;
	unsigned grid = (unsigned)((n + block - 1) / block);
	if (!inverse) {
		// the lifted samples must not overwrite inputs other threads still read: stage them in a copy
		size_t span = ((size_t)(N - 1) * SI + CH) * sizeof(int);
		int *lifted = nullptr;
		CUDA_OK(cudaMallocAsync(&lifted, span, st));
		CUDA_OK(cudaMemcpyAsync(lifted, d_in, span, cudaMemcpyDeviceToDevice, st));
		cdf53_1d_kernel<<<grid, block, 0, st>>>(d_out, d_in, lifted, N, SO, SI, CH);
		CUDA_OK(cudaGetLastError());
		CUDA_OK(cudaMemcpyAsync(d_in, lifted, span, cudaMemcpyDeviceToDevice, st));
		CUDA_OK(cudaFreeAsync(lifted, st));
	} else {
		icdf53_1d_kernel<<<grid, block, 0, st>>>(d_out, d_in, N, SO, SI, CH);
		CUDA_OK(cudaGetLastError());
	}
	return 0;
}

int lift_colour(int *d_buf, int total, bool inverse, cudaStream_t st)
{
	if (total <= 0)
		return 0;
	colour_kernel<<<(total + 255) / 256, 256, 0, st>>>(d_buf, total, inverse);
	CUDA_OK(cudaGetLastError());
	return 0;
}
