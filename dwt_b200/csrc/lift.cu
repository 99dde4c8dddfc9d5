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

namespace {

constexpr int STRIP_OUT = 60; // output columns per warp
constexpr int WARPS_PER_BLOCK = 4;

__device__ __forceinline__ int clampi(int x, int a, int b)
{
	return min(max(x, a), b);
}

// ------------------------------------------------------------------------------------------------ forward

template <int MODE>
struct FwdTraits {
	static constexpr int NC = MODE == 0 ? 3 : 1;
};

// raw samples of this lane's even / odd column in row y (colour transform fused for MODE 0)
template <int MODE>
__device__ __forceinline__ void fwd_load(const LiftLevel &p, int ch, int y, int cx0, int cx1,
                                         int (&a)[FwdTraits<MODE>::NC], int (&b)[FwdTraits<MODE>::NC])
{
	int cy = clampi(y, 0, p.H - 1);
	if constexpr (MODE == 0) {
		const uint8_t *row = (const uint8_t *)p.in + (size_t)cy * p.in_pitch * 3;
		const uint8_t *q0 = row + (size_t)cx0 * 3, *q1 = row + (size_t)cx1 * 3;
		int r0 = __ldg(q0), g0 = __ldg(q0 + 1), b0 = __ldg(q0 + 2);
		int r1 = __ldg(q1), g1 = __ldg(q1 + 1), b1 = __ldg(q1 + 2);
		// image.h:58-61
		int u0 = r0 - b0, t0 = b0 + u0 / 2, v0 = g0 - t0;
		int u1 = r1 - b1, t1 = b1 + u1 / 2, v1 = g1 - t1;
		a[0] = t0 + v0 / 2;
		a[1] = u0;
		a[2] = v0;
		b[0] = t1 + v1 / 2;
		b[1] = u1;
		b[2] = v1;
	} else if constexpr (MODE == 1) {
		const uint8_t *row = (const uint8_t *)p.in + (size_t)cy * p.in_pitch;
		a[0] = __ldg(row + cx0);
		b[0] = __ldg(row + cx1);
	} else {
		const int *row = (const int *)p.in + (size_t)ch * p.in_chan_stride + (size_t)cy * p.in_pitch;
		a[0] = __ldg(row + cx0);
		b[0] = __ldg(row + cx1);
	}
}

// horizontal lifting of one row: in (a,b) = x[xe], x[xe+1]; out (a,b) = s[xe], d[xe+1]  (cdf53.h:12-23)
template <int NC>
__device__ __forceinline__ void fwd_hlift(int (&a)[NC], int (&b)[NC], int xe, int N)
{
#pragma unroll
	for (int c = 0; c < NC; ++c) {
		int right = __shfl_down_sync(0xffffffffu, a[c], 1); // x[xe+2]
		int d;
		if (xe + 1 < N - 1)
			d = b[c] - (a[c] + right) / 2;
		else
			d = b[c] - a[c]; // xe+1 == N-1 (N even); garbage beyond the row, never stored
		int left = __shfl_up_sync(0xffffffffu, d, 1); // d[xe-1]
		int s;
		if (xe == 0)
			s = a[c] + d / 2;
		else if (xe + 1 <= N - 1)
			s = a[c] + (left + d) / 4;
		else
			s = a[c]; // xe == N-1 (N odd): the tail even sample is not updated (cdf53.h:21)
		a[c] = s;
		b[c] = d;
	}
}

template <int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) lift_fwd_kernel(LiftLevel p, int RS)
{
	constexpr int NC = FwdTraits<MODE>::NC;
	const int lane = threadIdx.x & 31;
	const int strip = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
	const int W = p.W, H = p.H;
	if (strip * STRIP_OUT >= W)
		return;
	const int ch = MODE == 2 ? blockIdx.z : 0;
	const int xe = strip * STRIP_OUT - 2 + 2 * lane; // this lane's even column
	const int cx0 = clampi(xe, 0, W - 1), cx1 = clampi(xe + 1, 0, W - 1);
	const int y0 = blockIdx.y * RS;
	if (y0 >= H)
		return;
	const int y1 = min(y0 + RS, H);
	const int w2 = (W + 1) / 2, h2 = (H + 1) / 2;
	const bool vs = lane >= 1 && lane <= 30 && xe < W; // this lane stores its even column
	const bool vd = vs && xe + 1 < W;                  // ... and its odd column

	// vertical window: e = row-lifted sample of the current even row, dp = vertical detail of the row above
	int e_s[NC], e_d[NC], dp_s[NC], dp_d[NC];
#pragma unroll
	for (int c = 0; c < NC; ++c)
		dp_s[c] = dp_d[c] = 0;
	fwd_load<MODE>(p, ch, y0, cx0, cx1, e_s, e_d);
	fwd_hlift<NC>(e_s, e_d, xe, W);
	if (y0 > 0) {
		int m2s[NC], m2d[NC], m1s[NC], m1d[NC];
		fwd_load<MODE>(p, ch, y0 - 2, cx0, cx1, m2s, m2d);
		fwd_load<MODE>(p, ch, y0 - 1, cx0, cx1, m1s, m1d);
		fwd_hlift<NC>(m2s, m2d, xe, W);
		fwd_hlift<NC>(m1s, m1d, xe, W);
#pragma unroll
		for (int c = 0; c < NC; ++c) { // y0-1 is an interior odd row
			dp_s[c] = m1s[c] - (m2s[c] + e_s[c]) / 2;
			dp_d[c] = m1d[c] - (m2d[c] + e_d[c]) / 2;
		}
	}
	int mx[NC];
#pragma unroll
	for (int c = 0; c < NC; ++c)
		mx[c] = 0;

	int no_s[NC], no_d[NC], ne_s[NC], ne_d[NC]; // prefetched raw rows r+1, r+2
	fwd_load<MODE>(p, ch, y0 + 1, cx0, cx1, no_s, no_d);
	fwd_load<MODE>(p, ch, y0 + 2, cx0, cx1, ne_s, ne_d);
	for (int r = y0; r < y1; r += 2) {
		int o_s[NC], o_d[NC], f_s[NC], f_d[NC];
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			o_s[c] = no_s[c];
			o_d[c] = no_d[c];
			f_s[c] = ne_s[c];
			f_d[c] = ne_d[c];
		}
		if (r + 2 < y1) {
			fwd_load<MODE>(p, ch, r + 3, cx0, cx1, no_s, no_d);
			fwd_load<MODE>(p, ch, r + 4, cx0, cx1, ne_s, ne_d);
		}
		fwd_hlift<NC>(o_s, o_d, xe, W);
		fwd_hlift<NC>(f_s, f_d, xe, W);
		const bool has_odd = r + 1 < H;
		const size_t lo_row = (size_t)(r >> 1), hi_row = (size_t)(h2 + (r >> 1));
		const int xl = xe >> 1, xh = w2 + (xe >> 1);
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			const int cc = MODE == 2 ? ch : c;
			int D_s, D_d, S_s, S_d;
			if (r + 1 < H - 1) {
				D_s = o_s[c] - (e_s[c] + f_s[c]) / 2;
				D_d = o_d[c] - (e_d[c] + f_d[c]) / 2;
			} else { // r+1 == H-1 (H even); unused when r+1 >= H
				D_s = o_s[c] - e_s[c];
				D_d = o_d[c] - e_d[c];
			}
			if (r == 0) {
				S_s = e_s[c] + D_s / 2;
				S_d = e_d[c] + D_d / 2;
			} else if (has_odd) {
				S_s = e_s[c] + (dp_s[c] + D_s) / 4;
				S_d = e_d[c] + (dp_d[c] + D_d) / 4;
			} else { // r == H-1 (H odd): untouched
				S_s = e_s[c];
				S_d = e_d[c];
			}
			int *ll = (int *)p.out + (size_t)cc * p.out_chan_stride;
			int *py = p.pyr + (size_t)cc * p.pyr_chan_stride;
			if (vs) {
				ll[lo_row * p.out_pitch + xl] = S_s; // LL
				if (has_odd) {
					py[hi_row * p.pyr_pitch + xl] = D_s; // LH (low-x, high-y)
					mx[c] = max(mx[c], abs(D_s));
				}
			}
			if (vd) {
				py[lo_row * p.pyr_pitch + xh] = S_d; // HL (high-x, low-y)
				mx[c] = max(mx[c], abs(S_d));
				if (has_odd) {
					py[hi_row * p.pyr_pitch + xh] = D_d; // HH
					mx[c] = max(mx[c], abs(D_d));
				}
			}
			e_s[c] = f_s[c];
			e_d[c] = f_d[c];
			dp_s[c] = D_s;
			dp_d[c] = D_d;
		}
	}
#pragma unroll
	for (int c = 0; c < NC; ++c) {
		int m = __reduce_max_sync(0xffffffffu, mx[c]);
		if (lane == 0 && m > 0)
			atomicMax(p.maxabs + (MODE == 2 ? ch : c), m);
	}
}

// ------------------------------------------------------------------------------------------------ inverse

// horizontal inverse of one row (cdf53.h:49-60): in (a,b) = s_i, d_i; out (a,b) = x[2i], x[2i+1]
template <int NC>
__device__ __forceinline__ void inv_hlift(int (&a)[NC], int (&b)[NC], int i, int N)
{
#pragma unroll
	for (int c = 0; c < NC; ++c) {
		int left = __shfl_up_sync(0xffffffffu, b[c], 1); // d_{i-1}
		int e;
		if (i == 0)
			e = a[c] - b[c] / 2;
		else if (2 * i + 1 <= N - 1)
			e = a[c] - (left + b[c]) / 4;
		else
			e = a[c]; // 2i == N-1 (N odd)
		int right = __shfl_down_sync(0xffffffffu, e, 1); // x[2i+2]
		int o;
		if (2 * i + 1 < N - 1)
			o = b[c] + (e + right) / 2;
		else
			o = b[c] + e; // 2i+1 == N-1 (N even)
		a[c] = e;
		b[c] = o;
	}
}

template <int MODE>
__device__ __forceinline__ void inv_store_row(const LiftLevel &p, int ch, int y, int i, bool v0, bool v1,
                                              int (&a)[FwdTraits<MODE>::NC], int (&b)[FwdTraits<MODE>::NC])
{
	if constexpr (MODE == 2) {
		int *row = (int *)p.out + (size_t)ch * p.out_chan_stride + (size_t)y * p.out_pitch;
		if (v0)
			row[2 * i] = a[0];
		if (v1)
			row[2 * i + 1] = b[0];
	} else if constexpr (MODE == 1) {
		uint8_t *row = (uint8_t *)p.out + (size_t)y * p.out_pitch;
		if (v0)
			row[2 * i] = (uint8_t)clampi(a[0], 0, 255);
		if (v1)
			row[2 * i + 1] = (uint8_t)clampi(b[0], 0, 255);
	} else {
		uint8_t *row = (uint8_t *)p.out + (size_t)y * p.out_pitch * 3;
#pragma unroll
		for (int k = 0; k < 2; ++k) {
			int(&q)[3] = k ? b : a;
			if (k ? v1 : v0) {
				// image.h:41-50 then pnm.h:108
				int Y = clampi(q[0], 0, 255), U = clampi(q[1], -255, 255), V = clampi(q[2], -255, 255);
				int T = Y - V / 2, G = V + T, B = T - U / 2, R = B + U;
				uint8_t *px = row + (size_t)(2 * i + k) * 3;
				px[0] = (uint8_t)clampi(R, 0, 255);
				px[1] = (uint8_t)clampi(G, 0, 255);
				px[2] = (uint8_t)clampi(B, 0, 255);
			}
		}
	}
}

template <int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) lift_inv_kernel(LiftLevel p, int RS)
{
	constexpr int NC = FwdTraits<MODE>::NC;
	const int lane = threadIdx.x & 31;
	const int strip = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
	const int W = p.W, H = p.H;
	if (strip * STRIP_OUT >= W)
		return;
	const int chz = MODE == 2 ? blockIdx.z : 0;
	const int i = strip * (STRIP_OUT / 2) - 1 + lane; // column pair index: output columns 2i, 2i+1
	const int w2 = (W + 1) / 2, h2 = (H + 1) / 2, wd = W / 2, hd = H / 2;
	const int y0 = blockIdx.y * RS;
	if (y0 >= H)
		return;
	const int y1 = min(y0 + RS, H);
	const bool v0 = lane >= 1 && lane <= 30 && 2 * i < W;
	const bool v1 = v0 && 2 * i + 1 < W;
	const int ci = clampi(i, 0, w2 - 1);                  // clamped low-x column
	const int cj = w2 + clampi(i, 0, max(wd - 1, 0));     // clamped high-x column (Mallat position)

	// loads the four band samples of pair row m: S (low-y) and D (high-y) for the low-x and high-x column
	auto load = [&](int m, int(&S_s)[NC], int(&S_d)[NC], int(&D_s)[NC], int(&D_d)[NC]) {
		int ms = clampi(m, 0, h2 - 1), md = h2 + clampi(m, 0, max(hd - 1, 0));
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			const int cc = MODE == 2 ? chz : c;
			const int *ll = (const int *)p.in + (size_t)cc * p.in_chan_stride;
			const int *py = p.pyr + (size_t)cc * p.pyr_chan_stride;
			S_s[c] = __ldg(ll + (size_t)ms * p.in_pitch + ci);
			S_d[c] = __ldg(py + (size_t)ms * p.pyr_pitch + cj);
			D_s[c] = __ldg(py + (size_t)md * p.pyr_pitch + ci);
			D_d[c] = __ldg(py + (size_t)md * p.pyr_pitch + cj);
		}
	};

	const int m0 = y0 >> 1;
	int Dm_s[NC], Dm_d[NC], Ep_s[NC], Ep_d[NC];
#pragma unroll
	for (int c = 0; c < NC; ++c)
		Dm_s[c] = Dm_d[c] = Ep_s[c] = Ep_d[c] = 0;
	if (m0 > 0) {
		int t0[NC], t1[NC];
		load(m0 - 1, t0, t1, Dm_s, Dm_d);
	}
	int nS_s[NC], nS_d[NC], nD_s[NC], nD_d[NC];
	load(m0, nS_s, nS_d, nD_s, nD_d);
	for (int m = m0; 2 * m - 1 < y1; ++m) {
		int S_s[NC], S_d[NC], D_s[NC], D_d[NC];
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			S_s[c] = nS_s[c];
			S_d[c] = nS_d[c];
			D_s[c] = nD_s[c];
			D_d[c] = nD_d[c];
		}
		if (2 * m + 1 < y1)
			load(m + 1, nS_s, nS_d, nD_s, nD_d);
		const bool hasE = 2 * m < H;
		int E_s[NC], E_d[NC], O_s[NC], O_d[NC];
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			// vertical inverse (cdf53.h:49-60 along columns)
			if (m == 0) {
				E_s[c] = S_s[c] - D_s[c] / 2;
				E_d[c] = S_d[c] - D_d[c] / 2;
			} else if (2 * m + 1 <= H - 1) {
				E_s[c] = S_s[c] - (Dm_s[c] + D_s[c]) / 4;
				E_d[c] = S_d[c] - (Dm_d[c] + D_d[c]) / 4;
			} else { // 2m == H-1 (H odd)
				E_s[c] = S_s[c];
				E_d[c] = S_d[c];
			}
			if (hasE) {
				O_s[c] = Dm_s[c] + (Ep_s[c] + E_s[c]) / 2;
				O_d[c] = Dm_d[c] + (Ep_d[c] + E_d[c]) / 2;
			} else { // 2m-1 == H-1 (H even)
				O_s[c] = Dm_s[c] + Ep_s[c];
				O_d[c] = Dm_d[c] + Ep_d[c];
			}
		}
		if (m > m0) { // odd row 2m-1 (uniform branch: shuffles stay converged)
			inv_hlift<NC>(O_s, O_d, i, W);
			inv_store_row<MODE>(p, chz, 2 * m - 1, i, v0, v1, O_s, O_d);
		}
		if (hasE && 2 * m < y1) { // even row 2m
			int T_s[NC], T_d[NC];
#pragma unroll
			for (int c = 0; c < NC; ++c) {
				T_s[c] = E_s[c];
				T_d[c] = E_d[c];
			}
			inv_hlift<NC>(T_s, T_d, i, W);
			inv_store_row<MODE>(p, chz, 2 * m, i, v0, v1, T_s, T_d);
		}
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			Ep_s[c] = E_s[c];
			Ep_d[c] = E_d[c];
			Dm_s[c] = D_s[c];
			Dm_d[c] = D_d[c];
		}
	}
}

int pick_rows(int W, int H, int planes)
{
	// aim for >= ~4k warps in flight; rows per segment even, 8..64
	long long strips = (W + STRIP_OUT - 1) / STRIP_OUT;
	int rs = 64;
	while (rs > 8 && strips * planes * ((H + rs - 1) / rs) < 4096)
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

int lift_forward_level(const LiftLevel &lv, int mode, cudaStream_t st, long long *launches)
{
	int planes = mode == 2 ? lv.channels : 1;
	int RS = pick_rows(lv.W, lv.H, planes);
	int strips = (lv.W + STRIP_OUT - 1) / STRIP_OUT;
	dim3 grid((strips + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, (lv.H + RS - 1) / RS, planes);
	dim3 block(WARPS_PER_BLOCK * 32);
	if (mode == 0)
		lift_fwd_kernel<0><<<grid, block, 0, st>>>(lv, RS);
	else if (mode == 1)
		lift_fwd_kernel<1><<<grid, block, 0, st>>>(lv, RS);
	else
		lift_fwd_kernel<2><<<grid, block, 0, st>>>(lv, RS);
	if (launches)
		++*launches;
	CUDA_OK(cudaGetLastError());
	return 0;
}

int lift_inverse_level(const LiftLevel &lv, int mode, cudaStream_t st, long long *launches)
{
	int planes = mode == 2 ? lv.channels : 1;
	int RS = pick_rows(lv.W, lv.H, planes);
	int strips = (lv.W + STRIP_OUT - 1) / STRIP_OUT;
	dim3 grid((strips + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, (lv.H + RS - 1) / RS, planes);
	dim3 block(WARPS_PER_BLOCK * 32);
	if (mode == 0)
		lift_inv_kernel<0><<<grid, block, 0, st>>>(lv, RS);
	else if (mode == 1)
		lift_inv_kernel<1><<<grid, block, 0, st>>>(lv, RS);
	else
		lift_inv_kernel<2><<<grid, block, 0, st>>>(lv, RS);
	if (launches)
		++*launches;
	CUDA_OK(cudaGetLastError());
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
