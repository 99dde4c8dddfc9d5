// layout.cuh -- small layout converters between the reference's interleaved int buffers and the planar
// device layout (static so each translation unit gets its own copy; no relocatable device code needed)
#pragma once

static __global__ void deinterleave_kernel(const int *in, int *out, long long npix, int ch)
{
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= npix * ch)
		return;
	long long px = t / ch;
	int c = (int)(t - px * ch);
	out[(long long)c * npix + px] = in[t];
}

static __global__ void interleave_kernel(const int *in, int *out, long long npix, int ch)
{
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= npix * ch)
		return;
	long long px = t / ch;
	int c = (int)(t - px * ch);
	out[t] = in[(long long)c * npix + px];
}

// interleaved Mallat pyramid in the reference's layout: details from pyr, root from the LL buffer
static __global__ void export_pyramid_kernel(const int *pyr, const int *root, int *out, int W, int H, int ch, int w0,
                                             int h0)
{
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	long long npix = (long long)W * H;
	if (t >= npix * ch)
		return;
	long long px = t / ch;
	int c = (int)(t - px * ch);
	int y = (int)(px / W), x = (int)(px - (long long)y * W);
	int v;
	if (x < w0 && y < h0)
		v = root[(long long)c * w0 * h0 + (long long)y * w0 + x];
	else
		v = pyr[(long long)c * npix + px];
	out[t] = v;
}
