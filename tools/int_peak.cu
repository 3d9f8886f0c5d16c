// int_peak.cu -- micro-benchmark of the B200 integer pipes ("the lanes/SM figure
// must be confirmed by a micro-benchmark on the box").  For each instruction class it runs ILP
// independent dependency chains per thread and reports lane-ops per clock per SM and T op/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/int_peak tools/int_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define ITERS 4096
#define ILP 8

enum { OP_IADD, OP_MAX, OP_ADDMAX, OP_MAX3, OP_LOP3, OP_IMAD, OP_MAX16, OP_ADD16, OP_ADDMAX16, OP_MIX_ALU_FMA, OP_SHFL, OP_SEL, OP_N };
static const char *NAMES[] = {"IADD3 (a+b)", "VIMNMX (max s32)", "VIADDMNMX (max(a+b,c))", "VIMNMX3 (max3)", "LOP3 (xor)",
                              "IMAD (a*b+c)", "VIMNMX.U16x2", "VIADD.16x2", "VIADDMNMX.S16x2", "mix VIMNMX+IMAD", "SHFL.UP", "ISETP+SEL"};

template <int OP>
__global__ void k(int *out, int a0, int b0, long long *clk)
{
	int x[ILP];
#pragma unroll
	for (int i = 0; i < ILP; ++i) x[i] = a0 + threadIdx.x + i;
	int b = b0, c = b0 * 3 + 1;
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) {
			if (OP == OP_IADD) x[i] = x[i] + b;
			else if (OP == OP_MAX) x[i] = max(x[i], b + i);
			else if (OP == OP_ADDMAX) x[i] = __viaddmax_s32(x[i], b, c);
			else if (OP == OP_MAX3) x[i] = __vimax3_s32(x[i], b, c + i);
			else if (OP == OP_LOP3) x[i] = x[i] ^ b;
			else if (OP == OP_IMAD) x[i] = x[i] * b + c;
			else if (OP == OP_MAX16) x[i] = (int)__vmaxu2((unsigned)x[i], (unsigned)(b + i));
			else if (OP == OP_ADD16) x[i] = (int)__vadd2((unsigned)x[i], (unsigned)b);
			else if (OP == OP_ADDMAX16) x[i] = (int)__viaddmax_s16x2((unsigned)x[i], (unsigned)b, (unsigned)c);
			else if (OP == OP_MIX_ALU_FMA) { if (i & 1) x[i] = x[i] * b + c; else x[i] = max(x[i], b + i); }
			else if (OP == OP_SHFL) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
			else if (OP == OP_SEL) x[i] = (x[i] > c) ? b : x[i] + 0;
		}
		b += 1;
	}
	long long t1 = clock64();
	int s = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i) s ^= x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP> static void run(int sms, int *d_out, long long *d_clk)
{
	const int blocks = sms, threads = 1024;      // one 1024-thread block per SM (two do not co-reside: they ran back to back)
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<OP><<<blocks, threads>>>(d_out, 1, 3, d_clk);
	cudaEventRecord(e0);
	k<OP><<<blocks, threads>>>(d_out, 1, 3, d_clk);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long clk0; cudaMemcpy(&clk0, d_clk, sizeof clk0, cudaMemcpyDeviceToHost);
	const double ops = (double)blocks * threads * ITERS * ILP;
	// per-SM rate from the in-kernel cycle counter of block 0 (1024 threads resident per SM)
	const double per_clk_sm = 1024.0 * ITERS * ILP / (double)clk0;
	printf("%-26s %8.3f ms  %7.2f Tlane-op/s  %6.1f lane-ops/clk/SM  (block0 %lld clk)\n", NAMES[OP], ms, ops / ms / 1e9, per_clk_sm, clk0);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	printf("device: %s, %d SMs, clockRate %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
	int *d_out; long long *d_clk;
	cudaMalloc(&d_out, sizeof(int) * p.multiProcessorCount * 1024);
	cudaMalloc(&d_clk, sizeof(long long) * p.multiProcessorCount * 2);
	run<OP_IADD>(p.multiProcessorCount, d_out, d_clk);
	run<OP_MAX>(p.multiProcessorCount, d_out, d_clk);
	run<OP_ADDMAX>(p.multiProcessorCount, d_out, d_clk);
	run<OP_MAX3>(p.multiProcessorCount, d_out, d_clk);
	run<OP_LOP3>(p.multiProcessorCount, d_out, d_clk);
	run<OP_IMAD>(p.multiProcessorCount, d_out, d_clk);
	run<OP_MAX16>(p.multiProcessorCount, d_out, d_clk);
	run<OP_ADD16>(p.multiProcessorCount, d_out, d_clk);
	run<OP_ADDMAX16>(p.multiProcessorCount, d_out, d_clk);
	run<OP_MIX_ALU_FMA>(p.multiProcessorCount, d_out, d_clk);
	run<OP_SHFL>(p.multiProcessorCount, d_out, d_clk);
	run<OP_SEL>(p.multiProcessorCount, d_out, d_clk);
	cudaError_t e = cudaDeviceSynchronize();
	printf("status: %s\n", cudaGetErrorString(e));
	return e != cudaSuccess;
}
