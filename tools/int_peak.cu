// int_peak.cu -- micro-benchmark of the B200 integer pipes ("the lanes/SM figure
// must be confirmed by a micro-benchmark on the box").  For each instruction class it runs ILP
// independent dependency chains per thread, 32 warps per SM, and reports lane-ops per clock per SM
// (from the in-kernel cycle counter of block 0) and T lane-op/s (CUDA events, all SMs).
// The mixes answer the question the fill kernels' design rests on: do the ALU pipe (VIMNMX / VIADDMNMX /
// LOP3 / IADD3) and the FMA pipe (IMAD) issue side by side, i.e. is an IMAD free next to an ALU-bound stream?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/int_peak tools/int_peak.cu
//   (check the loop bodies with: cuobjdump -sass tools/_build/int_peak | grep -A40 'Function : _Z1kILi<n>')
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define ITERS 4096
#define ILP 8

enum { OP_IADD, OP_MAX, OP_ADDMAX, OP_MAX3, OP_LOP3, OP_IMAD, OP_MAX16, OP_ADDMAX16, OP_MAX3_16, OP_MIX_1_1, OP_MIX_2_1, OP_MIX16_1_1, OP_SHFL, OP_LDS, OP_CELL_MIX, OP_N };
static const char *NAMES[] = {"IADD3 (a+b)", "VIMNMX (max s32)", "VIADDMNMX (max(a+b,c))", "VIMNMX3 (max3)", "LOP3 ((a&b)|c)",
                              "IMAD (a*b+c)", "VIMNMX.U16x2", "VIADDMNMX.U16x2", "VIMNMX3.U16x2", "mix VIADDMNMX : IMAD 1:1",
                              "mix (VIADDMNMX,LOP3) : IMAD 2:1", "mix VIADDMNMX.U16x2 : IMAD 1:1", "SHFL.UP", "LDS.32", "packed cell mix 8 ALU : 7 IMAD"};
static const int OPS_PER_SLOT[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 3, 2, 1, 1, 15};     // instructions one loop slot issues

template <int OP>
__global__ void k(int *out, int a0, int b0, long long *clk)
{
	__shared__ int sh[1024];
	sh[threadIdx.x] = threadIdx.x;
	__syncthreads();
	int x[ILP], y[ILP];
#pragma unroll
	for (int i = 0; i < ILP; ++i) { x[i] = a0 + threadIdx.x + i; y[i] = a0 * 7 + i; }
	int b = b0, c = b0 * 3 + 1;
	asm volatile("" : "+r"(b), "+r"(c));
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) {
			if (OP == OP_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
			else if (OP == OP_MAX) asm volatile("max.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
			else if (OP == OP_ADDMAX) x[i] = __viaddmax_s32(x[i], b, c);
			else if (OP == OP_MAX3) x[i] = __vimax3_s32(x[i], b, c);
			else if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(x[i]) : "r"(b), "r"(c));
			else if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
			else if (OP == OP_MAX16) x[i] = (int)__vmaxu2((unsigned)x[i], (unsigned)b);
			else if (OP == OP_ADDMAX16) x[i] = (int)__viaddmax_u16x2((unsigned)x[i], (unsigned)b, (unsigned)c);
			else if (OP == OP_MAX3_16) x[i] = (int)__vimax3_u16x2((unsigned)x[i], (unsigned)b, (unsigned)c);
			else if (OP == OP_MIX_1_1) { x[i] = __viaddmax_s32(x[i], b, c); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c)); }
			else if (OP == OP_MIX_2_1) { x[i] = __viaddmax_s32(x[i], b, c); asm volatile("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(x[i]) : "r"(b), "r"(c));
			                             asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c)); }
			else if (OP == OP_MIX16_1_1) {      // volatile statements keep their order: ALU and FMA instructions alternate in the SASS
				asm volatile("{.reg .b32 t1;\n\tadd.u16x2 t1, %0, %1;\n\tmax.u16x2 %0, t1, %2;}" : "+r"(x[i]) : "r"(b), "r"(c));
				asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c));
			}
			else if (OP == OP_CELL_MIX) {       // the instruction mix of one packed cell of at_cell.cuh: 5 max-type + 3 LOP3 on the ALU pipe, 7 IMAD on the FMA pipe
				asm volatile("{.reg .b32 t1;\n\tadd.u16x2 t1, %0, %1;\n\tmax.u16x2 %0, t1, %2;}" : "+r"(x[i]) : "r"(b), "r"(c));
				asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(x[i]) : "r"(b), "r"(c));
				asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c));
				asm volatile("{.reg .b32 t1;\n\tadd.u16x2 t1, %0, %1;\n\tmax.u16x2 %0, t1, %2;}" : "+r"(x[i]) : "r"(b), "r"(c));
				asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(x[i]) : "r"(b), "r"(c));
				asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c));
				asm volatile("{.reg .b32 t1;\n\tadd.u16x2 t1, %0, %1;\n\tmax.u16x2 %0, t1, %2;}" : "+r"(x[i]) : "r"(b), "r"(c));
				asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(x[i]) : "r"(b), "r"(c));
				asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c));
				x[i] = (int)__vimax3_u16x2((unsigned)x[i], (unsigned)b, (unsigned)c);
				asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c));
				asm volatile("{.reg .b32 t1;\n\tadd.u16x2 t1, %0, %1;\n\tmax.u16x2 %0, t1, %2;}" : "+r"(x[i]) : "r"(b), "r"(c));
			}
			else if (OP == OP_SHFL) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
			else if (OP == OP_LDS) x[i] = sh[x[i] & 1023];
		}
	}
	long long t1 = clock64();
	int s = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i) s ^= x[i] ^ y[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP> static void run(int sms, int *d_out, long long *d_clk)
{
	const int blocks = sms, threads = 1024;      // one 1024-thread block per SM
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<OP><<<blocks, threads>>>(d_out, 1, 3, d_clk);
	cudaEventRecord(e0);
	k<OP><<<blocks, threads>>>(d_out, 1, 3, d_clk);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long clk0; cudaMemcpy(&clk0, d_clk, sizeof clk0, cudaMemcpyDeviceToHost);
	const double instr_lanes = (double)threads * ITERS * ILP * OPS_PER_SLOT[OP];      // lane-instructions of one block = one SM
	printf("%-34s %8.3f ms  %7.2f Tlane-op/s  %6.1f lane-ops/clk/SM  %5.2f warp-instr/clk/SM  (block0 %lld clk)\n", NAMES[OP], ms,
	       instr_lanes * blocks / (ms * 1e-3) / 1e12, instr_lanes / (double)clk0, instr_lanes / 32.0 / (double)clk0, clk0);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	printf("device: %s, %d SMs, clockRate %d kHz; 1024 threads/SM, ILP %d, %d iterations\n", p.name, p.multiProcessorCount, khz, ILP, ITERS);
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
	run<OP_ADDMAX16>(p.multiProcessorCount, d_out, d_clk);
	run<OP_MAX3_16>(p.multiProcessorCount, d_out, d_clk);
	run<OP_MIX_1_1>(p.multiProcessorCount, d_out, d_clk);
	run<OP_MIX_2_1>(p.multiProcessorCount, d_out, d_clk);
	run<OP_MIX16_1_1>(p.multiProcessorCount, d_out, d_clk);
	run<OP_SHFL>(p.multiProcessorCount, d_out, d_clk);
	run<OP_LDS>(p.multiProcessorCount, d_out, d_clk);
	run<OP_CELL_MIX>(p.multiProcessorCount, d_out, d_clk);
	cudaError_t e = cudaDeviceSynchronize();
	printf("status: %s\n", cudaGetErrorString(e));
	return e != cudaSuccess;
}
