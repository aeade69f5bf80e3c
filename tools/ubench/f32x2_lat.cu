// Dependent-chain latency and occupancy-limited throughput of FADD2/FMUL2/FFMA2 vs scalar
// (how many warps per scheduler does a packed-FP32 kernel need?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_lat f32x2_lat.cu && ./f32x2_lat
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
DEV float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
#define ITERS 8192
// MODE 0: FADD chain, 1: FADD2 chain, 2: FFMA chain, 3: FFMA2 chain (CH independent chains per thread)
template <int MODE, int CH> __global__ void k(float* out, float one, float b1, float b2) {
	float a[CH];
	u64 A[CH];
	for (int i = 0; i < CH; ++i) { a[i] = one + threadIdx.x + i; A[i] = pk(a[i], a[i] + 1); }
	const u64 O = pk(one, one), B1 = pk(b1, b1 + 1), B2 = pk(b2, b2 + 2);
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int i = 0; i < CH; ++i) {
			if (MODE == 0) a[i] = fadd(a[i], one);
			if (MODE == 1) A[i] = add2(A[i], O);
			if (MODE == 2) a[i] = ffma(a[i], b1, b2);
			if (MODE == 3) A[i] = fma2(A[i], B1, B2);
		}
	}
	float s = 0;
	for (int i = 0; i < CH; ++i) { float x, y; upk(A[i], x, y); s += a[i] + x + y; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int CH> void run(const char* name, int warps_per_smsp, float* out, int sms, double mhz) {
	const int threads = warps_per_smsp * 4 * 32;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<MODE, CH><<<sms, threads>>>(out, 1.0f, 1.0f, 0.5f);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int r = 0; r < 5; ++r) {
		cudaEventRecord(e0);
		k<MODE, CH><<<sms, threads>>>(out, 1.0f, 1.0f, 0.5f);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	const double cycles = best * 1e-3 * mhz * 1e6;
	printf("%-8s chains/thread %d  warps/SMSP %d : %7.2f cycles per instruction per warp, %6.3f warp-inst/clk/SMSP\n", name, CH,
	       warps_per_smsp, cycles / ((double)ITERS * CH), (double)ITERS * CH * warps_per_smsp / cycles);
}
int main() {
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double mhz = khz / 1000.0; const int s = p.multiProcessorCount;
	float* out; cudaMalloc(&out, (size_t)s * 1024 * 4);
	printf("%s %d SMs %.0f MHz\n", p.name, s, mhz);
	// warm the clocks
	for (int i = 0; i < 20; ++i) k<3, 4><<<s, 1024>>>(out, 1.f, 1.f, .5f);
	cudaDeviceSynchronize();
	run<0, 1>("FADD", 1, out, s, mhz);  run<1, 1>("FADD2", 1, out, s, mhz);
	run<2, 1>("FFMA", 1, out, s, mhz);  run<3, 1>("FFMA2", 1, out, s, mhz);
	for (int w = 2; w <= 8; w *= 2) {
		run<2, 1>("FFMA", w, out, s, mhz); run<3, 1>("FFMA2", w, out, s, mhz);
		run<2, 2>("FFMA", w, out, s, mhz); run<3, 2>("FFMA2", w, out, s, mhz);
	}
	run<3, 4>("FFMA2", 4, out, s, mhz); run<3, 4>("FFMA2", 8, out, s, mhz);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
