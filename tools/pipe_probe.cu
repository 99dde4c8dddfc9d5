// pipe_probe.cu -- integer instruction throughput on the machine at hand (development aid for the lifting kernels):
// warp instructions per clock and SM for the candidates of the C-division sequences.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu && tools/pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int OP>
__device__ __forceinline__ void op(unsigned &x, unsigned y)
{
	if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y));            // IMAD
	if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y));            // IMAD.HI.U32
	if (OP == 2) asm volatile("mad.hi.s32 %0, %0, %1, %1;" : "+r"(x) : "r"(y));            // IMAD.HI
	if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));                   // IADD3 / IMAD.IADD (ptxas chooses)
	if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x) : "r"(y));         // LOP3
	if (OP == 5) asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y));         // SHF
	if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(x) : "r"(y));           // PRMT
	if (OP == 7) asm volatile("{.reg .s32 t; max.s32 t, %0, %1; max.s32 %0, t, 77;}" : "+r"(x) : "r"(y)); // VIMNMX(3)
	if (OP == 8) { // x + (x >>> 31): LEA.HI
		unsigned t;
		asm volatile("shr.u32 %0, %1, 31;" : "=r"(t) : "r"(x));
		asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(t));
	}
	if (OP == 9) { // y + (x >> 1) arithmetic: LEA.HI.SX32
		int t;
		asm volatile("shr.s32 %0, %1, 1;" : "=r"(t) : "r"(x));
		asm volatile("add.s32 %0, %1, %2;" : "=r"(x) : "r"(t), "r"(y));
	}
	if (OP == 10) { // one IMAD and one LOP3, independent: do the two pipes overlap?
		asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x) : "r"(y));
	}
	if (OP == 11) asm volatile("mul.hi.u32 %0, %0, 2;" : "+r"(x));                          // IMAD.HI.U32 with an immediate
	if (OP == 12) asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(x) : "r"(y));             // IMAD with an immediate
	if (OP == 13) asm volatile("abs.s32 %0, %0;" : "+r"(x));                                // IABS
}

template <int OP>
__global__ void __launch_bounds__(1024) probe(unsigned *out, unsigned seed, long long *cycles)
{
	unsigned v[CHAINS], w[CHAINS];
#pragma unroll
	for (int k = 0; k < CHAINS; ++k) {
		v[k] = seed * (threadIdx.x + k + 1);
		w[k] = seed + k;
	}
	const long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < ITERS; ++i) {
#pragma unroll
		for (int k = 0; k < CHAINS; ++k) {
			op<OP>(v[k], seed);
			if (OP == 10)
				asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(w[k]) : "r"(seed));
		}
	}
	const long long t1 = clock64();
	unsigned s = 0;
#pragma unroll
	for (int k = 0; k < CHAINS; ++k)
		s += v[k] + w[k];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		*cycles = t1 - t0;
}

template <int OP>
static void run(const char *name, int per_iter, unsigned *out, long long *cyc)
{
	int sms = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	probe<OP><<<sms * 2, 1024>>>(out, 12345u, cyc); // 64 warps per SM: 16 per scheduler
	cudaDeviceSynchronize();
	probe<OP><<<sms * 2, 1024>>>(out, 12345u, cyc);
	cudaDeviceSynchronize();
	long long c = 0;
	cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
	// block 0 shares its SM with one other block: 64 warps run ITERS * CHAINS * per_iter instructions each in c cycles
	const double winst = 64.0 * ITERS * CHAINS * per_iter;
	printf("%-34s %7.2f warp instr / clk / SM   (%lld cycles)\n", name, winst / (double)c, c);
}

int main()
{
	unsigned *out;
	long long *cyc;
	cudaMalloc(&out, 4u << 20);
	cudaMalloc(&cyc, 8);
	run<0>("IMAD (mad.lo r,r,r)", 1, out, cyc);
	run<12>("IMAD (mad.lo r,imm,r)", 1, out, cyc);
	run<1>("IMAD.HI.U32 (mad.hi.u32)", 1, out, cyc);
	run<2>("IMAD.HI (mad.hi.s32)", 1, out, cyc);
	run<11>("mul.hi.u32 x, 2", 1, out, cyc);
	run<3>("add.u32", 1, out, cyc);
	run<4>("LOP3", 1, out, cyc);
	run<5>("SHF", 1, out, cyc);
	run<6>("PRMT", 1, out, cyc);
	run<7>("max3 (2 PTX max -> VIMNMX3?)", 1, out, cyc);
	run<13>("IABS", 1, out, cyc);
	run<8>("x + (x >>> 31)  (LEA.HI?)", 1, out, cyc);
	run<9>("y + (x >> 1)    (LEA.HI.SX32?)", 1, out, cyc);
	run<10>("LOP3 + IMAD pair (2 instr)", 2, out, cyc);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		printf("CUDA error: %s\n", cudaGetErrorString(e));
	return e != cudaSuccess;
}
