// Microbenchmark: do the FMA pipe (FFMA2) and the ALU pipe (FMNMX3 and friends) overlap on sm_100?
//
// ncu on k1_filter says FMA pipe 57 % + ALU pipe 36 % of cycles and never both; the loop's cost is the SUM of its
// FFMA2 and FMNMX3 costs. This isolates the question in register-only loops (no shared memory, no barriers): per
// iteration and per accumulator k (K of them, independent), NF packed FMAs on loop-carried accumulators followed by
// NM "reduction" operations of kind OP that consume the fresh FMA results. Reported: cycles per SMSP per k-step
// (a k-step = NF FFMA2 + NM ops issued by one warp), so that "sum" or "max" of the isolated costs can be read off.
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_pipes.cu -o tools/ubench_pipes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

enum { OP_NONE = 0, OP_FMNMX3, OP_FMNMX, OP_VIMNMX3, OP_LOP3, OP_IADD3, OP_FSETP_OR, OP_FADD, OP_IMAD, OP_HMNMX2, OP_F2FP, OP_SCALAR_FFMA_MIN3, OP_PRMT };

// SCALAR = 1: the FMAs are 2*NF scalar FFMAs instead of NF FFMA2s (same lane-operations)
template <int K, int NF, int NM, int OP, int SCALAR>
__global__ void __launch_bounds__(256) pipes(float* out, int iters, float fa, float fb)
{
	u64 acc[NF][K];
	float m[K]; int mi[K]; unsigned pr = 0;
	const u64 A = pack2(fa, fa * 1.0001f), B = pack2(fb, fb * 0.5f);
#pragma unroll
	for (int k = 0; k < K; k++) {
		m[k] = 1e30f + k; mi[k] = 0x7fffffff - k;
#pragma unroll
		for (int f = 0; f < NF; f++) acc[f][k] = pack2(threadIdx.x * 1e-3f + k, f + 0.25f * k);
	}
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < K; k++) {
#pragma unroll
			for (int f = 0; f < NF; f++) {
				if (SCALAR) {
					float lo, hi; unpack2(acc[f][k], lo, hi);
					lo = fmaf(lo, fa, fb); hi = fmaf(hi, fa, fb);
					acc[f][k] = pack2(lo, hi);
				} else acc[f][k] = fma2(acc[f][k], A, B);
			}
#pragma unroll
			for (int q = 0; q < NM; q++) {
				float lo, hi; unpack2(acc[q % NF][k], lo, hi);
				if (OP == OP_FMNMX3) m[k] = min3(m[k], lo, hi);
				if (OP == OP_FMNMX) { m[k] = fminf(m[k], lo); }
				if (OP == OP_VIMNMX3) mi[k] = min(mi[k], min(__float_as_int(lo), __float_as_int(hi)));
				if (OP == OP_LOP3) mi[k] = (mi[k] | __float_as_int(lo)) ^ __float_as_int(hi);
				if (OP == OP_IADD3) mi[k] = mi[k] + __float_as_int(lo) + __float_as_int(hi);
				if (OP == OP_FSETP_OR) { unsigned p; asm("{ .reg .pred q; setp.ne.u32 q, %1, 0; setp.le.or.f32 q, %2, %3, q; selp.u32 %0, 1, 0, q; }" : "=r"(p) : "r"(pr), "f"(lo), "f"(m[k])); pr = p; }
				if (OP == OP_FADD) m[k] = m[k] + lo;
				if (OP == OP_IMAD) mi[k] = mi[k] * __float_as_int(lo) + __float_as_int(hi);
				if (OP == OP_HMNMX2) { unsigned r; asm("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"((unsigned)mi[k]), "r"(__float_as_uint(lo))); mi[k] = (int)r; }
				if (OP == OP_F2FP) { unsigned r; asm("cvt.rz.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(lo), "f"(hi)); mi[k] ^= (int)r; }
				if (OP == OP_PRMT) mi[k] = (int)__byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632) + 0 * mi[k];
			}
		}
	}
	float s = (float)pr;
#pragma unroll
	for (int k = 0; k < K; k++) {
		s += m[k] + (float)mi[k];
#pragma unroll
		for (int f = 0; f < NF; f++) { float lo, hi; unpack2(acc[f][k], lo, hi); s += lo + hi; }
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int K, int NF, int NM, int OP, int SCALAR = 0> void run(const char* name, float* out, int sms, int ctas_per_sm)
{
	const int iters = 4096, blocks = sms * ctas_per_sm;
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	float best = 1e30f;
	for (int r = 0; r < 6; r++) {
		cudaEventRecord(a); pipes<K, NF, NM, OP, SCALAR><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(b);
		CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (r > 0 && ms < best) best = ms;
	}
	// warps per SMSP = ctas_per_sm * 8 / 4; k-steps per warp = iters * K
	const double ksteps_per_smsp = (double)ctas_per_sm * 2.0 * iters * K;
	const double cyc = best * 1e-3 * 1.965e9 / ksteps_per_smsp;
	printf("%-44s K=%d CTAs/SM=%d: %7.3f ms  %6.2f cycles per (%d FFMA%s + %d op) step per SMSP\n", name, K, ctas_per_sm, best, cyc, SCALAR ? 2 * NF : NF, SCALAR ? "" : "2", NM);
}

int main()
{
	int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
	const int sms = prop.multiProcessorCount;
	float* out; CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 4));
	printf("device %s SMs=%d; cycles assume 1.965 GHz\n", prop.name, sms);
	for (int c = 1; c <= 4; c *= 2) {
		run<8, 2, 0, OP_NONE>("2 FFMA2", out, sms, c);
		run<8, 0 + 1, 1, OP_FMNMX3>("1 FFMA2 + 1 FMNMX3", out, sms, c);
		run<8, 2, 1, OP_FMNMX3>("2 FFMA2 + 1 FMNMX3 (the planar loop's mix)", out, sms, c);
		run<8, 3, 1, OP_FMNMX3>("3 FFMA2 + 1 FMNMX3 (the full bound's mix)", out, sms, c);
		run<8, 2, 2, OP_FMNMX>("2 FFMA2 + 2 FMNMX", out, sms, c);
		run<8, 2, 1, OP_VIMNMX3>("2 FFMA2 + 1 VIMNMX3", out, sms, c);
		run<8, 2, 1, OP_LOP3>("2 FFMA2 + 1 LOP3", out, sms, c);
		run<8, 2, 1, OP_IADD3>("2 FFMA2 + 1 IADD3", out, sms, c);
		run<8, 2, 2, OP_FSETP_OR>("2 FFMA2 + 2 FSETP.OR", out, sms, c);
		run<8, 2, 1, OP_FADD>("2 FFMA2 + 1 FADD (FMA pipe)", out, sms, c);
		run<8, 2, 1, OP_IMAD>("2 FFMA2 + 1 IMAD (FMA pipe)", out, sms, c);
		run<8, 2, 1, OP_HMNMX2>("2 FFMA2 + 1 HMNMX2.BF16", out, sms, c);
		run<8, 2, 1, OP_F2FP>("2 FFMA2 + 1 F2FP.BF16 + LOP", out, sms, c);
		run<8, 2, 1, OP_PRMT>("2 FFMA2 + 1 PRMT", out, sms, c);
		run<8, 2, 0, OP_NONE, 1>("4 FFMA (scalar)", out, sms, c);
		run<8, 2, 1, OP_FMNMX3, 1>("4 FFMA (scalar) + 1 FMNMX3", out, sms, c);
		run<4, 2, 1, OP_FMNMX3>("2 FFMA2 + 1 FMNMX3", out, sms, c);
		run<16, 2, 1, OP_FMNMX3>("2 FFMA2 + 1 FMNMX3", out, sms, c);
	}
	// ALU alone: how long does an FMNMX3 take when nothing else competes?
	run<8, 1, 4, OP_FMNMX3>("1 FFMA2 + 4 FMNMX3", out, sms, 2);
	run<8, 1, 4, OP_LOP3>("1 FFMA2 + 4 LOP3", out, sms, 2);
	run<8, 1, 4, OP_IADD3>("1 FFMA2 + 4 IADD3", out, sms, 2);
	return 0;
}
