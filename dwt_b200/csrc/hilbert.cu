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

namespace {

__device__ __forceinline__ void hilbert_d2xy(int n, u32 d, int &x, int &y) // hilbert.h:15-34
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

__device__ __forceinline__ int overlap(int o, int cs, int lim) // |[o, o+cs) n [0, lim)|
{
	int hi = min(o + cs, lim);
	return hi > o ? hi - o : 0;
}

__global__ void cell_count_kernel(int n, int cs, int w1, int h1, int w0, int h0, int ncell, u32 *counts)
{
	int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= ncell)
		return;
	int x, y;
	hilbert_d2xy(n, (u32)q * (u32)(cs * cs), x, y);
	int ox = x & ~(cs - 1), oy = y & ~(cs - 1);
	int inside = overlap(ox, cs, w1) * overlap(oy, cs, h1);
	int ll = overlap(ox, cs, w0) * overlap(oy, cs, h0);
	counts[q] = (u32)(inside - ll);
}

// single-block in-place exclusive scan
__global__ void __launch_bounds__(1024) exscan_u32_kernel(u32 *data, int n)
{
	__shared__ u64 ws[32];
	int per = (n + blockDim.x - 1) / blockDim.x;
	int b = threadIdx.x * per, e = min(b + per, n);
	u64 s = 0;
	for (int i = b; i < e; ++i)
		s += data[i];
	u64 tot;
	u64 base = block_exscan_u64(s, ws, &tot);
	u32 run = (u32)base;
	for (int i = b; i < e; ++i) {
		u32 v = data[i];
		data[i] = run;
		run += v;
	}
}

struct ChanLayout {
	int planes[3];
	long long bsbase[4];
};

struct CellParams {
	int n, cs, w1, h1, w0, h0;
	int channels;
	int GT, gbase;   // groups per channel, first group of this level
	const u32 *cell_base;
	ChanLayout lay;
};

__global__ void __launch_bounds__(256) linearize_kernel(CellParams P, const int *pyr, long long chan_stride,
                                                         int pitch, u32 *bs)
{
	__shared__ u32 vals[3][1024];
	__shared__ int wcnt[4][8];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int q = blockIdx.x;
	const int npos = P.cs * P.cs;
	const u32 R0 = P.cell_base[q];
	int px[4], py[4], rk[4];
	bool ok[4];
#pragma unroll
	for (int it = 0; it < 4; ++it) {
		int dloc = it * 256 + tid;
		bool act = dloc < npos;
		int x = 0, y = 0;
		if (act)
			hilbert_d2xy(P.n, (u32)q * (u32)npos + (u32)dloc, x, y);
		bool v = act && x < P.w1 && y < P.h1 && (x >= P.w0 || y >= P.h0);
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
			u32 *dst = bs + P.lay.bsbase[c] + (long long)lane * P.GT + P.gbase + g;
			if (whole)
				*dst = mine;
			else if (mine)
				atomicOr(dst, mine);
		}
	}
}

__global__ void __launch_bounds__(256) reconstruct_kernel(CellParams P, const u32 *bs, const int *missing, int level,
                                                           int *pyr, long long chan_stride, int pitch)
{
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	__shared__ int wcnt[4][8];
	const int q = blockIdx.x;
	const int npos = P.cs * P.cs;
	const u32 R0 = P.cell_base[q];
	int px[4], py[4], rk[4];
	bool ok[4];
#pragma unroll
	for (int it = 0; it < 4; ++it) {
		int dloc = it * 256 + tid;
		bool act = dloc < npos;
		int x = 0, y = 0;
		if (act)
			hilbert_d2xy(P.n, (u32)q * (u32)npos + (u32)dloc, x, y);
		bool v = act && x < P.w1 && y < P.h1 && (x >= P.w0 || y >= P.h0);
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
			const u32 *src = bs + P.lay.bsbase[c] + P.gbase + g;
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

CellParams make_cell_params(const Geom &g, const HilbertPlan &plan, const Sched &s, int l)
{
	CellParams P;
	P.n = g.len[l + 1];
	P.cs = plan.cs[l];
	P.w1 = g.w[l + 1];
	P.h1 = g.h[l + 1];
	P.w0 = g.w[l];
	P.h0 = g.h[l];
	P.channels = g.channels;
	P.GT = g.GT;
	P.gbase = g.gbase[l];
	P.cell_base = plan.cell_base + plan.cell_off[l];
	P.lay = make_layout(s, g.channels);
	return P;
}

} // namespace

int hilbert_plan_build(const Geom &g, HilbertPlan *plan, cudaStream_t st, long long *launches)
{
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
	plan->cell_base = nullptr;
	CUDA_OK(cudaMalloc(&plan->cell_base, sizeof(u32) * (size_t)(off > 0 ? off : 1)));
	for (int l = 0; l < g.levels; ++l) {
		u32 *cb = plan->cell_base + plan->cell_off[l];
		int nc = plan->ncell[l];
		cell_count_kernel<<<(nc + 255) / 256, 256, 0, st>>>(g.len[l + 1], plan->cs[l], g.w[l + 1], g.h[l + 1], g.w[l],
		                                                    g.h[l], nc, cb);
		exscan_u32_kernel<<<1, 1024, 0, st>>>(cb, nc);
		if (launches)
			*launches += 2;
	}
	CUDA_OK(cudaGetLastError());
	return 0;
}

void hilbert_plan_free(HilbertPlan *plan)
{
	if (plan->cell_base)
		cudaFree(plan->cell_base);
	plan->cell_base = nullptr;
}

int hilbert_linearize(const Geom &g, const HilbertPlan &plan, const Sched &s, const int *pyr,
                      long long pyr_chan_stride, int pyr_pitch, u32 *bs, int levels_used, cudaStream_t st,
                      long long *launches)
{
	for (int l = 0; l < levels_used; ++l) {
		CellParams P = make_cell_params(g, plan, s, l);
		linearize_kernel<<<plan.ncell[l], 256, 0, st>>>(P, pyr, pyr_chan_stride, pyr_pitch, bs);
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
	for (int l = 0; l < levels_used; ++l) {
		CellParams P = make_cell_params(g, plan, s, l);
		reconstruct_kernel<<<plan.ncell[l], 256, 0, st>>>(P, bs, missing_dev, l, pyr, pyr_chan_stride, pyr_pitch);
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
