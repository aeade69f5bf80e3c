// Microbenchmark: issue / pipe cost of the packed FP32 instructions of sm_100a
// (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2) against their scalar forms,
// alone and mixed with ALU-pipe (FMNMX) and XU (MUFU.RSQ) work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define DEV __device__ __forceinline__
DEV u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
DEV void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
DEV u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
DEV u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
DEV u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
DEV float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
DEV float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
DEV float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
DEV float fmin_(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
DEV float rsq(float a) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

#define CH 8      // independent chains per thread
#define ITERS 4096

// MODE: 0 FADD  1 FADD2  2 FMUL  3 FMUL2  4 FFMA  5 FFMA2
//       6 FADD+FMNMX (1:1)  7 FADD2+FMNMX (1:1)  8 FADD2+2 FMNMX  9 FFMA2 + RSQ (8:1)
//       10 FFMA(scalar, 2 per) + FMNMX   11 FADD2 + FMUL2 alternating  12 FMNMX alone
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, float seed, float one) {
	float a[CH], b[CH];
	u64 A[CH];
	for (int i = 0; i < CH; ++i) {
		a[i] = seed + threadIdx.x + i;
		b[i] = seed * i;
		A[i] = pk(a[i], b[i]);
	}
	const u64 O = pk(one, one);
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int i = 0; i < CH; ++i) {
			if (MODE == 0) a[i] = fadd(a[i], one);
			if (MODE == 1) A[i] = add2(A[i], O);
			if (MODE == 2) a[i] = fmul(a[i], one);
			if (MODE == 3) A[i] = mul2(A[i], O);
			if (MODE == 4) a[i] = ffma(a[i], one, one);
			if (MODE == 5) A[i] = fma2(A[i], O, O);
			if (MODE == 6) { a[i] = fadd(a[i], one); b[i] = fmin_(b[i], a[(i + 1) % CH]); }
			if (MODE == 7) { A[i] = add2(A[i], O); b[i] = fmin_(b[i], a[(i + 1) % CH]); }
			if (MODE == 8) { A[i] = add2(A[i], O); b[i] = fmin_(b[i], a[(i + 1) % CH]); a[i] = fmin_(a[i], b[(i + 3) % CH]); }
			if (MODE == 9) { A[i] = fma2(A[i], O, O); if (i == 0) b[0] = rsq(b[0]); }
			if (MODE == 10) { a[i] = ffma(a[i], one, one); b[i] = ffma(b[i], one, one); a[i] = fmin_(a[i], b[(i + 3) % CH]); }
			if (MODE == 11) { A[i] = (i & 1) ? add2(A[i], O) : mul2(A[i], O); }
			if (MODE == 12) a[i] = fmin_(a[i], b[i]);
		}
	}
	float s = 0;
	for (int i = 0; i < CH; ++i) {
		float x, y;
		upk(A[i], x, y);
		s += a[i] + b[i] + x + y;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, double inst_per_iter_per_chain, double flop_per_iter_per_chain, float* out, int sms, double mhz) {
	const int ctas = sms * 8;  // 2048 threads/SM requested; 8 x 256
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k<MODE><<<ctas, 256>>>(out, 1.0f, 1.0f);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int r = 0; r < 5; ++r) {
		cudaEventRecord(e0);
		k<MODE><<<ctas, 256>>>(out, 1.0f, 1.0f);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	const double warps = (double)ctas * 8;
	const double winst = warps * ITERS * CH * inst_per_iter_per_chain;
	const double cycles = best * 1e-3 * mhz * 1e6;
	printf("%-34s %8.3f ms  %6.3f warp-inst/clk/SMSP  %7.2f TFLOP/s\n", name, best, winst / cycles / sms / 4,
	       warps * 32 * ITERS * CH * flop_per_iter_per_chain / (best * 1e-3) / 1e12);
}

int main() {
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double mhz = khz / 1000.0;
	printf("%s, %d SMs, %.0f MHz (max; rates assume the GPU runs at it)\n", p.name, p.multiProcessorCount, mhz);
	float* out;
	cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
	const int s = p.multiProcessorCount;
	run<0>("FADD", 1, 1, out, s, mhz);
	run<1>("FADD2", 1, 2, out, s, mhz);
	run<2>("FMUL", 1, 1, out, s, mhz);
	run<3>("FMUL2", 1, 2, out, s, mhz);
	run<4>("FFMA", 1, 2, out, s, mhz);
	run<5>("FFMA2", 1, 4, out, s, mhz);
	run<12>("FMNMX", 1, 1, out, s, mhz);
	run<6>("FADD + FMNMX", 2, 2, out, s, mhz);
	run<7>("FADD2 + FMNMX", 2, 3, out, s, mhz);
	run<8>("FADD2 + 2 FMNMX", 3, 4, out, s, mhz);
	run<9>("8 FFMA2 + 1 MUFU.RSQ", 1.125, 4, out, s, mhz);
	run<10>("2 FFMA + FMNMX", 3, 5, out, s, mhz);
	run<11>("FADD2 / FMUL2 alternating", 1, 2, out, s, mhz);
	cudaError_t e = cudaDeviceSynchronize();
	printf("%s\n", cudaGetErrorString(e));
	return e != cudaSuccess;
}
