// Microbenchmark: can the tensor cores serve as the lower-bound FILTER of the matching step?
//
// North star: "Tensor cores are used only if ncu shows the |p|^2+|q|^2-2p.q contraction beats the FP32 FMA path at K=3".
// The contraction is not the reference's arithmetic, so it can only be the filter e~ = w_j - 2 p.q_j (nn_filter.cu); its
// result lands in TMEM as a [128 sources x N targets] FP32 tile and the filter needs, per source (= per TMEM lane =
// per thread), the MINIMUM over the targets (columns). There is no tcgen05.ld.red on sm_100a, so every accumulator
// must be read into registers (tcgen05.ld) and min-reduced on the ALU pipe. Two questions, answered separately:
//   part A (this file, mode "epi"): the epilogue ceiling — pairs/clk/SM that tcgen05.ld + FMNMX3 can consume, with no
//       MMA at all (reads whatever TMEM holds). If this is below the FP32 filter (37 pairs/clk/SM measured in k1_filter,
//       9.2 per SMSP), the tensor-core path is dead whatever the MMA rate.
//   part B (mode "mma"): tcgen05.mma kind::tf32 M=128 N=256 K=8 issue rate with the same epilogue running behind it
//       (A/B tiles are zero-filled shared memory, canonical K-major no-swizzle layout; timing only).
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_tc_filter.cu -o tools/ubench_tc_filter
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
	uint32_t* u = reinterpret_cast<uint32_t*>(v);
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
	             : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]),
	               "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]),
	               "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
	             : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// REDUCE: 0 = one FMNMX per load (keeps the load alive, measures TMEM read alone), 1 = full min over every value (FMNMX3 per 2 values),
//         2 = threshold test accumulated in one predicate (FSETP.LE.OR per value): all the filter needs is "any e~ <= tau"
template <int COLS, int REDUCE>
__global__ void __launch_bounds__(256) epi_kernel(float* out, int reps)
{
	__shared__ uint32_t tmem_base_s;
	const int warp = threadIdx.x >> 5;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(COLS));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);      // a warp reads the 32 lanes of its quarter
	// warps 4..7 (if present) read the second half of the columns of the same lanes
	const int nw = blockDim.x >> 5;
	const int c_lo = (nw > 4 && warp >= 4) ? COLS / 2 : 0;
	const int c_hi = (nw > 4 && warp < 4) ? COLS / 2 : COLS;
	float m = 3e38f;
	float a[32], b[32];
	const float tau = -1e30f + (float)reps;                 // never true on the garbage TMEM holds... whatever it holds, the work is the same
	if (REDUCE == 2) { asm volatile(".reg .pred pe;"); asm volatile("setp.ne.u32 pe, 0, 0;"); }
	for (int r = 0; r < reps; r++) {
		tmem_ld32(base + c_lo, a);
		for (int c = c_lo; c < c_hi; c += 64) {
			tmem_wait_ld();
			tmem_ld32(base + c + 32, b);
			if (REDUCE == 1) {
#pragma unroll
				for (int k = 0; k < 32; k += 2) m = min3(m, a[k], a[k + 1]);
			} else if (REDUCE == 2) {
#pragma unroll
				for (int k = 0; k < 32; k++) asm volatile("setp.le.or.f32 pe, %0, %1, pe;" :: "f"(a[k]), "f"(tau));
			} else m = fminf(m, a[0]);
			tmem_wait_ld();
			if (c + 64 < c_hi) tmem_ld32(base + c + 64, a);
			if (REDUCE == 1) {
#pragma unroll
				for (int k = 0; k < 32; k += 2) m = min3(m, b[k], b[k + 1]);
			} else if (REDUCE == 2) {
#pragma unroll
				for (int k = 0; k < 32; k++) asm volatile("setp.le.or.f32 pe, %0, %1, pe;" :: "f"(b[k]), "f"(tau));
			} else m = fminf(m, b[31]);
		}
	}
	if (REDUCE == 2) { unsigned f; asm volatile("selp.u32 %0, 1, 0, pe;" : "=r"(f)); m = (float)f; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = m;
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "n"(COLS));
}

template <int COLS, int REDUCE> static void run_epi(const char* name, float* out, int sms, int ctas_per_sm, int threads)
{
	const int reps = 2000;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int r = 0; r < 5; r++) {
		cudaEventRecord(e0); epi_kernel<COLS, REDUCE><<<sms * ctas_per_sm, threads>>>(out, reps); cudaEventRecord(e1);
		CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
	}
	const double vals_per_sm = (double)ctas_per_sm * 128.0 * COLS * reps;       // one value = one (source, target) pair of the filter
	const double cyc = best * 1e-3 * 1.965e9;
	printf("%-40s cols=%3d CTAs/SM=%d threads=%3d: %7.3f ms  %7.1f pairs/clk/SM  (%6.1f B/clk/SM of TMEM reads)\n", name, COLS, ctas_per_sm, threads, best,
	       vals_per_sm / cyc, 4.0 * vals_per_sm / cyc);
}

// ---------------------------------------------------------------------------------------------------------------
// part B: tcgen05.mma kind::tf32, M=128, N=256, K=8 per instruction, operands in shared memory (no swizzle, K-major).
// One elected thread issues MMAs into accumulator stage s while the other warps min-reduce stage s^1.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
	// sm_100 shared-memory matrix descriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 0b001 at [46,49), swizzle NONE
	return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count)); }
// bounded wait: a descriptor mistake must end the kernel with an error flag, not hang the GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity)
{
	for (int spin = 0; spin < (1 << 22); spin++) {
		uint32_t ok;
		asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
		if (ok) return true;
	}
	return false;
}
__device__ int g_timeout = 0;

// KSTEPS = MMAs of K=8 accumulated per tile (1: plain 3-D bound padded to 8; 2: hi/lo split, K=16)
template <int KSTEPS, int REDUCE>
__global__ void __launch_bounds__(160) mma_kernel(float* out, int tiles)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint32_t tmem_base_s;
	__shared__ __align__(8) uint64_t full_bar[2], empty_bar[2];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	// A: 128 x 8 tf32 = 4 KB, B: 256 x 8 tf32 = 8 KB, zero-filled (values are irrelevant to the timing)
	for (int i = threadIdx.x; i < (4096 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
	if (threadIdx.x == 0) { for (int s = 0; s < 2; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 128); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	if (warp == 4) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy zero fill -> async proxy (tensor core reads)
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tbase = tmem_base_s;
	const uint32_t sA = (uint32_t)__cvta_generic_to_shared(smem), sB = sA + 4096;
	// instruction descriptor, kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
	// canonical no-swizzle K-major: core matrix = 8 rows x 16 B (128 B contiguous); K = 8 tf32 = 2 core matrices per row group
	const uint64_t dA = smem_desc(sA, 128, 256), dB = smem_desc(sB, 128, 256);
	float m = 3e38f;
	if (warp == 4) {
		if (lane == 0) {
			for (int t = 0; t < tiles; t++) {
				const int s = t & 1;
				if (t >= 2 && !mbar_wait(&empty_bar[s], (uint32_t)(((t >> 1) - 1) & 1))) { g_timeout = 1; break; }
				asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
				for (int k = 0; k < KSTEPS; k++) {
					asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
					             :: "r"(tbase + (uint32_t)(s * 256)), "l"(dA), "l"(dB), "r"(idesc), "r"((uint32_t)(k > 0)));
				}
				asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(&full_bar[s])) : "memory");
			}
		}
	} else {
		const uint32_t base = tbase + ((uint32_t)(warp * 32) << 16);
		float a[32], b[32];
		for (int t = 0; t < tiles; t++) {
			const int s = t & 1;
			if (!mbar_wait(&full_bar[s], (uint32_t)((t >> 1) & 1))) { g_timeout = 2; break; }
			asm volatile("tcgen05.fence::after_thread_sync;");
			const uint32_t cb = base + (uint32_t)(s * 256);
			tmem_ld32(cb, a);
			for (int c = 0; c < 256; c += 64) {
				tmem_wait_ld();
				tmem_ld32(cb + c + 32, b);
				if (REDUCE) {
#pragma unroll
					for (int k = 0; k < 32; k += 2) m = min3(m, a[k], a[k + 1]);
				} else m = fminf(m, a[t & 31]);
				tmem_wait_ld();
				if (c + 64 < 256) tmem_ld32(cb + c + 64, a);
				if (REDUCE) {
#pragma unroll
					for (int k = 0; k < 32; k += 2) m = min3(m, b[k], b[k + 1]);
				} else m = fminf(m, b[t & 31]);
			}
			asm volatile("tcgen05.fence::before_thread_sync;");
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(&empty_bar[s])) : "memory");
		}
		out[blockIdx.x * 128 + threadIdx.x] = m;
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "n"(512));
}

template <int KSTEPS, int REDUCE> static void run_mma(const char* name, float* out, int sms)
{
	const int tiles = 4000;
	const size_t smem = 4096 + 8192 + 1024;
	CK(cudaFuncSetAttribute(mma_kernel<KSTEPS, REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int r = 0; r < 4; r++) {
		cudaEventRecord(e0); mma_kernel<KSTEPS, REDUCE><<<sms, 160, smem>>>(out, tiles); cudaEventRecord(e1);
		CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
	}
	int to = 0; CK(cudaMemcpyFromSymbol(&to, g_timeout, sizeof to));
	if (to) { printf("%-40s TIMED OUT waiting on an mbarrier (%d): the MMA never completed\n", name, to); return; }
	const double pairs_per_sm = 128.0 * 256.0 * tiles;
	const double cyc = best * 1e-3 * 1.965e9;
	printf("%-40s K=%2d: %7.3f ms  %7.1f pairs/clk/SM  (%.0f cycles per 128x256 tile)\n", name, 8 * KSTEPS, best, pairs_per_sm / cyc, cyc / tiles);
}

// ---------------------------------------------------------------------------------------------------------------
// part C (mode "check"): numerical validation of the descriptors. A [128 x 16] and B [256 x 16] hold random values that are
// exact in TF32; D = A B^T is accumulated over two K=8 instructions, read back with tcgen05.ld and compared with a
// double-precision product on the host. Also reports the largest accumulation error relative to sum |a_k b_k|
// (the tensor core's FP32 accumulate is not IEEE round-to-nearest; the filter's error bound needs a measured constant).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) check_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint32_t tmem_base_s;
	__shared__ __align__(8) uint64_t bar;
	const int warp = threadIdx.x >> 5;
	float* sAf = reinterpret_cast<float*>(smem);                 // 128 x 16 floats, canonical K-major no-swizzle: LBO 128 B, SBO 512 B
	float* sBf = sAf + 128 * 16;                                 // 256 x 16
	for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) { const int r = i / 16, k = i % 16; sAf[(r / 8) * 128 + (k / 4) * 32 + (r % 8) * 4 + (k % 4)] = A[i]; }
	for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) { const int r = i / 16, k = i % 16; sBf[(r / 8) * 128 + (k / 4) * 32 + (r % 8) * 4 + (k % 4)] = B[i]; }
	if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(256));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tbase = tmem_base_s;
	const uint32_t sA = (uint32_t)__cvta_generic_to_shared(sAf), sB = (uint32_t)__cvta_generic_to_shared(sBf);
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
	if (threadIdx.x == 0) {
		for (int k = 0; k < 2; k++) {
			const uint64_t dA = smem_desc(sA + k * 256, 128, 512), dB = smem_desc(sB + k * 256, 128, 512);     // K advances by 2 core matrices = 256 B
			asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
			             :: "r"(tbase), "l"(dA), "l"(dB), "r"(idesc), "r"((uint32_t)(k > 0)));
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
	}
	const bool ok = mbar_wait(&bar, 0);
	asm volatile("tcgen05.fence::after_thread_sync;");
	if (ok) {
		const uint32_t base = tbase + ((uint32_t)(warp * 32) << 16);
		float v[32];
		for (int c = 0; c < 256; c += 32) {
			tmem_ld32(base + c, v);
			tmem_wait_ld();
			for (int k = 0; k < 32; k++) D[(size_t)threadIdx.x * 256 + c + k] = v[k];
		}
	} else if (threadIdx.x == 0) g_timeout = 3;
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "n"(256));
}

// part D (mode "lat"): latency of one tcgen05.mma batch — issue, tcgen05.commit, mbarrier wait — as the issuing thread sees it
template <int N, int NMMA>
__global__ void __launch_bounds__(128) lat_kernel(long long* out, int reps)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint32_t tmem_base_s;
	__shared__ __align__(8) uint64_t bar;
	const int warp = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < (4096 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
	if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(256));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tbase = tmem_base_s;
	const uint32_t sA = (uint32_t)__cvta_generic_to_shared(smem), sB = sA + 4096;
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
	const uint64_t dA = smem_desc(sA, 128, 256), dB = smem_desc(sB, 128, 256);
	if (threadIdx.x == 0) {
		long long best = 1ll << 60, sum = 0;
		for (int r = 0; r < reps; r++) {
			const long long t0 = clock64();
#pragma unroll
			for (int k = 0; k < NMMA; k++)
				asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" :: "r"(tbase), "l"(dA), "l"(dB), "r"(idesc), "r"((uint32_t)(k > 0)));
			asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
			if (!mbar_wait(&bar, (uint32_t)(r & 1))) { g_timeout = 4; break; }
			const long long dt = clock64() - t0;
			if (dt < best) best = dt;
			sum += dt;
		}
		out[0] = best; out[1] = sum / reps;
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "n"(256));
}
template <int N, int NMMA> static void run_lat()
{
	long long* d; CK(cudaMalloc(&d, 16));
	const size_t smem = 4096 + 8192 + 1024;
	CK(cudaFuncSetAttribute(lat_kernel<N, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	lat_kernel<N, NMMA><<<1, 128, smem>>>(d, 200);
	CK(cudaDeviceSynchronize());
	long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
	printf("latency issue -> commit -> mbarrier seen: %d x tcgen05.mma tf32 128x%dx8: min %lld cycles, mean %lld\n", NMMA, N, h[0], h[1]);
	cudaFree(d);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

static int run_check()
{
	const int M = 128, N = 256, K = 16;
	float *hA = new float[M * K], *hB = new float[N * K], *hD = new float[M * N];
	uint32_t st = 12345u;
	auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (float)((st >> 8) & 0xffff) / 32768.0f - 1.0f; };
	for (int i = 0; i < M * K; i++) hA[i] = tf32_trunc(rnd() * 7.0f);
	for (int i = 0; i < N * K; i++) hB[i] = tf32_trunc(rnd() * 3.0f);
	float *dA, *dB, *dD;
	CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dB, N * K * 4)); CK(cudaMalloc(&dD, M * N * 4));
	CK(cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice));
	CK(cudaMemset(dD, 0xff, M * N * 4));
	const size_t smem = (128 + 256) * 16 * 4 + 1024;
	CK(cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	check_kernel<<<1, 128, smem>>>(dA, dB, dD);
	CK(cudaDeviceSynchronize());
	int to = 0; CK(cudaMemcpyFromSymbol(&to, g_timeout, sizeof to));
	if (to) { printf("check: TIMED OUT (%d)\n", to); return 1; }
	CK(cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost));
	double max_abs = 0, max_rel = 0; int bad = 0;
	for (int i = 0; i < M; i++)
		for (int j = 0; j < N; j++) {
			double ref = 0, mag = 0;
			for (int k = 0; k < K; k++) { ref += (double)hA[i * K + k] * hB[j * K + k]; mag += fabs((double)hA[i * K + k] * hB[j * K + k]); }
			const double err = fabs((double)hD[i * N + j] - ref);
			if (err > max_abs) max_abs = err;
			if (err / mag > max_rel) max_rel = err / mag;
			if (err > 1e-4 * mag + 1e-6) { if (bad < 5) printf("  D[%d][%d] = %g, expected %g\n", i, j, hD[i * N + j], ref); bad++; }
		}
	printf("check: tcgen05.mma kind::tf32 128x256x16 (2 instructions, no-swizzle K-major descriptors): %s; max |err| %.3e, max |err| / sum|a_k b_k| = %.3e = %.2f x 2^-24\n",
	       bad ? "MISMATCH" : "matches the host product", max_abs, max_rel, max_rel * 16777216.0);
	return bad ? 1 : 0;
}

int main(int argc, char** argv)
{
	if (argc > 1 && !strcmp(argv[1], "check")) return run_check();
	if (argc > 1 && !strcmp(argv[1], "lat")) { run_lat<128, 1>(); run_lat<128, 2>(); run_lat<256, 1>(); run_lat<256, 2>(); run_lat<256, 4>(); run_lat<64, 2>(); return 0; }
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	float* out; CK(cudaMalloc(&out, (size_t)sms * 4 * 256 * 4));
	const bool do_mma = argc > 1 && !strcmp(argv[1], "mma");
	printf("device %s SMs=%d; FP32 filter today: 37 pairs/clk/SM (k1_filter planar, 9.2 per SMSP)\n", prop.name, sms);
	if (!do_mma) {
		run_epi<512, 0>("tcgen05.ld only", out, sms, 1, 128);
		run_epi<512, 1>("tcgen05.ld + FMNMX3 min", out, sms, 1, 128);
		run_epi<512, 0>("tcgen05.ld only", out, sms, 1, 256);
		run_epi<512, 1>("tcgen05.ld + FMNMX3 min", out, sms, 1, 256);
		run_epi<256, 0>("tcgen05.ld only", out, sms, 2, 128);
		run_epi<256, 1>("tcgen05.ld + FMNMX3 min", out, sms, 2, 128);
		run_epi<256, 1>("tcgen05.ld + FMNMX3 min", out, sms, 2, 256);
		run_epi<128, 1>("tcgen05.ld + FMNMX3 min", out, sms, 4, 128);
		run_epi<512, 2>("tcgen05.ld + FSETP.OR threshold test", out, sms, 1, 128);
		run_epi<512, 2>("tcgen05.ld + FSETP.OR threshold test", out, sms, 1, 256);
		run_epi<256, 2>("tcgen05.ld + FSETP.OR threshold test", out, sms, 2, 128);
		run_epi<256, 2>("tcgen05.ld + FSETP.OR threshold test", out, sms, 2, 256);
		run_epi<128, 2>("tcgen05.ld + FSETP.OR threshold test", out, sms, 4, 128);
	} else {
		run_mma<1, 0>("mma tf32 128x256x8 + ld only", out, sms);
		run_mma<1, 1>("mma tf32 128x256x8 + ld + min", out, sms);
		run_mma<2, 1>("mma tf32 128x256x16 + ld + min", out, sms);
	}
	CK(cudaDeviceSynchronize());
	printf("done\n");
	return 0;
}
