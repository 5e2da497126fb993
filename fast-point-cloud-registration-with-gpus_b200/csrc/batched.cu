// batched.cu — K9: many small independent registrations, one per CTA, the whole ICP loop in ONE kernel.
//
// BASELINE.json config 5 (4096 pairs of 2048-point clouds). The reference has no batched mode; the
// nearest relative is its shared-memory block reduction test (src/tests/centroid.cu:14-47). Each CTA
// runs, for its registration, exactly the loop of src/ICP_point_to_point.cu:295-423:
//   matching (K1's packed inner loop, target resident in shared memory, 8 sources per thread in registers,
//   sub-tile tracking + re-scan for the index) -> FP64 block reduction of {sum p, sum q, sum q p^T} ->
//   3x3 Jacobi SVD by one thread (solve_device.cuh) -> transform of the registers with RyT's arithmetic ->
//   FP64 block reduction of the squared residual -> the reference's stop test,
// with __syncthreads() as the only synchronisation: no kernel launches, no host round trips, no global
// memory traffic inside the loop except the final results. Two CTAs (16 warps) share an SM; registrations are drawn
// from an atomic counter because their iteration counts differ (static round-robin left CTAs idle at the end).
#include "common.cuh"
#include <algorithm>
#include "k1_device.cuh"
#include "solve_device.cuh"
#include <cstdlib>

namespace icpb {

constexpr int K9_S = 8, K9_THREADS = 256;
constexpr int K9_MAX_N = K9_S * K9_THREADS;     // 2048 sources per registration
constexpr int K9_MAX_M = 4096;                  // targets per registration (48 KB of shared memory)
constexpr int K9_TRK = 64;                      // tracking sub-tile (small clouds: keep the re-scan cheap)

struct BatchParams {
	const float* sources;   // [batch][n][3]
	const float* targets;   // [batch][m][3]
	int batch, n, m, max_iter, stop_early, mode, flags;
	float sentinel, thr0;
	double tol;
	float* errors;          // [batch][max_iter+1]
	int* iterations;        // [batch]
	int* iterations_run;    // [batch]
	double* R;              // [batch][9]
	double* t;              // [batch][3]
	int* idx;               // [batch][n] or nullptr
	int* next;              // zeroed before the launch: CTAs draw registrations from it (their iteration counts differ)
	// streaming upload (config 5 end to end): the clouds arrive in chunks of `chunk` registrations WHILE the kernel runs;
	// the copy engine sets ready[c] = 1 after chunk c has landed (a 4-byte H2D copy queued behind the chunk's data on
	// the same stream). nullptr: everything is resident before the launch.
	const int* ready;
	int chunk;
	int* fail;              // set when a chunk never arrived (bounded wait instead of a hang)
};

// The drawing thread waits until registration b's chunk has landed. DMA writes by the copy engine need no SM, so a
// kernel that occupies every SM cannot starve them (unlike a kernel waiting for another kernel).
__device__ __forceinline__ int k9_draw(const BatchParams& p)
{
	const int b = atomicAdd(p.next, 1);
	if (p.ready == nullptr || b >= p.batch) return b;
	const int* flag = p.ready + b / p.chunk;
	const long long t0 = clock64();
	for (;;) {
		int v;
		asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
		if (v != 0) return b;
		if (clock64() - t0 > 20000000000ll) { *p.fail = 1; return p.batch; }      // ~10 s: the upload died
		__nanosleep(200);
	}
}

__device__ __forceinline__ double block_sum(double v, double* red /* [THREADS/32] */)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0.0;
#pragma unroll
	for (int w = 0; w < K9_THREADS / 32; w++) s += red[w];
	return s;
}

template <int MODE>
__global__ void __launch_bounds__(K9_THREADS, 2) icp_batched_kernel(const BatchParams p)
{
	extern __shared__ __align__(16) unsigned char k9_smem[];
	const int mpad = ((p.m + K9_TRK - 1) / K9_TRK) * K9_TRK;
	float* tx = reinterpret_cast<float*>(k9_smem);      // [mpad + 4] each, padded with +inf
	float* ty = tx + mpad + 4;
	float* tz = ty + mpad + 4;
	__shared__ double red[K9_THREADS / 32];
	__shared__ double mom[16];
	__shared__ IterState st;                            // reuse of the streaming path's control block + solver
	__shared__ float err_prev;
	const int tid = threadIdx.x;
	const float inf = __int_as_float(0x7f800000);

	__shared__ int s_b;
	while (true) {
		__syncthreads();
		if (tid == 0) s_b = k9_draw(p);
		__syncthreads();
		const int b = s_b;
		if (b >= p.batch) break;
		const float* T = p.targets + (size_t)b * p.m * 3;
		for (int j = tid; j < mpad + 4; j += K9_THREADS) {
			const bool ok = j < p.m;
			tx[j] = ok ? T[3 * (size_t)j] : inf; ty[j] = ok ? T[3 * (size_t)j + 1] : inf; tz[j] = ok ? T[3 * (size_t)j + 2] : inf;
		}
		float sx[K9_S], sy[K9_S], sz[K9_S];
		int idx[K9_S];
		const float* Sg = p.sources + (size_t)b * p.n * 3;
#pragma unroll
		for (int s = 0; s < K9_S; s++) {
			const int i = s * K9_THREADS + tid;
			const bool ok = i < p.n;
			sx[s] = ok ? Sg[3 * (size_t)i] : 0.f; sy[s] = ok ? Sg[3 * (size_t)i + 1] : 0.f; sz[s] = ok ? Sg[3 * (size_t)i + 2] : 0.f;
			idx[s] = 0;
		}
		if (tid == 0) {
			st.done = 0; st.iteration = 0; st.iters_run = 0; st.flags = p.flags;
			for (int k = 0; k < 9; k++) st.Rtot[k] = (k % 4 == 0) ? 1.0 : 0.0;
			for (int k = 0; k < 3; k++) st.ttot[k] = 0.0;
			err_prev = 0.f;
			p.errors[(size_t)b * (p.max_iter + 1)] = 0.f;
		}
		__syncthreads();

		while (true) {
			// ---- matching: K1's inner loop over the shared-memory target ----
			float m[K9_S], thr[K9_S];
			int best[K9_S];
#pragma unroll
			for (int s = 0; s < K9_S; s++) { m[s] = p.thr0; thr[s] = p.thr0; best[s] = -1; }
			const float4* X4 = reinterpret_cast<const float4*>(tx);
			const float4* Y4 = reinterpret_cast<const float4*>(ty);
			const float4* Z4 = reinterpret_cast<const float4*>(tz);
			const int nsub = mpad / K9_TRK;
#pragma unroll 1
			for (int sub = 0; sub < nsub; sub++) {
				const int j0 = sub * (K9_TRK / 4), j1 = j0 + K9_TRK / 4;
				float4 X = X4[j0], Y = Y4[j0], Z = Z4[j0];
#pragma unroll 2
				for (int j = j0; j < j1; j++) {
					const float4 Xn = X4[j + 1], Yn = Y4[j + 1], Zn = Z4[j + 1];   // +4 floats of padding make this safe
					const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
					const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
					const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
#pragma unroll
					for (int s = 0; s < K9_S; s++) {
						const u64 PX = bcast2v(sx[s]), PY = bcast2v(sy[s]), PZ = bcast2v(sz[s]);
						u64 dx = sub2(PX, x01), dy = sub2(PY, y01), dz = sub2(PZ, z01);
						u64 d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
						float a, c;
						unpack2(d, a, c);
						m[s] = min3(m[s], a, c);
						dx = sub2(PX, x23); dy = sub2(PY, y23); dz = sub2(PZ, z23);
						d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
						unpack2(d, a, c);
						m[s] = min3(m[s], a, c);
					}
					X = Xn; Y = Yn; Z = Zn;
				}
#pragma unroll
				for (int s = 0; s < K9_S; s++) {
					if (m[s] < thr[s]) { best[s] = sub; thr[s] = lower_threshold<MODE>(m[s]); m[s] = thr[s]; }
				}
			}
			// index recovery: first j of the remembered sub-tile that attains the minimum
#pragma unroll
			for (int s = 0; s < K9_S; s++) {
				if (best[s] >= 0) {
					const float target = (MODE == ICPB_DIST_SQRT) ? __fsqrt_rn(thr[s]) : thr[s];
					const int base = best[s] * K9_TRK;
					int found = -1;
					for (int j = 0; j < K9_TRK && found < 0; j++) {
						float d = dist_chain(sx[s], sy[s], sz[s], tx[base + j], ty[base + j], tz[base + j]);
						if (MODE == ICPB_DIST_SQRT) d = __fsqrt_rn(d);
						if (d <= target) found = j;
					}
					if (found >= 0) idx[s] = base + found;
				}
			}
			// ---- moments (FP64), SVD by thread 0 ----
			double acc[15];
#pragma unroll
			for (int k = 0; k < 15; k++) acc[k] = 0.0;
#pragma unroll
			for (int s = 0; s < K9_S; s++) {
				if (s * K9_THREADS + tid < p.n) {
					const double x = sx[s], y = sy[s], z = sz[s];
					const double qx = tx[idx[s]], qy = ty[idx[s]], qz = tz[idx[s]];
					acc[0] += x; acc[1] += y; acc[2] += z; acc[3] += qx; acc[4] += qy; acc[5] += qz;
					acc[6] += qx * x; acc[7] += qy * x; acc[8] += qz * x;
					acc[9] += qx * y; acc[10] += qy * y; acc[11] += qz * y;
					acc[12] += qx * z; acc[13] += qy * z; acc[14] += qz * z;
				}
			}
#pragma unroll
			for (int k = 0; k < 15; k++) { const double v = block_sum(acc[k], red); if (tid == 0) mom[k] = v; }
			if (tid == 0) {
				for (int k = 0; k < 15; k++) st.moments[k] = mom[k];
				st.moments[15] = (double)p.n;
				solve_p2p(&st);
			}
			__syncthreads();
			// ---- transform (RyT arithmetic) + residual against the same correspondences ----
			const float r0 = st.R[0], r1 = st.R[1], r2 = st.R[2], r3 = st.R[3], r4 = st.R[4], r5 = st.R[5], r6 = st.R[6], r7 = st.R[7], r8 = st.R[8];
			const float t0 = st.T[0], t1 = st.T[1], t2 = st.T[2];
			double e = 0.0;
#pragma unroll
			for (int s = 0; s < K9_S; s++) {
				const float x = sx[s], y = sy[s], z = sz[s];
				sx[s] = __fadd_rn(__fmaf_rn(r6, z, __fmaf_rn(r0, x, __fmul_rn(r3, y))), t0);
				sy[s] = __fadd_rn(__fmaf_rn(r7, z, __fmaf_rn(r1, x, __fmul_rn(r4, y))), t1);
				sz[s] = __fadd_rn(__fmaf_rn(r8, z, __fmaf_rn(r2, x, __fmul_rn(r5, y))), t2);
				if (s * K9_THREADS + tid < p.n) {
					const float ex = __fsub_rn(sx[s], tx[idx[s]]), ey = __fsub_rn(sy[s], ty[idx[s]]), ez = __fsub_rn(sz[s], tz[idx[s]]);
					e += (double)ex * (double)ex + (double)ey * (double)ey + (double)ez * (double)ez;
				}
			}
			const double esum = block_sum(e, red);
			if (tid == 0) {
				// finish_iteration of the streaming path (src/ICP_point_to_point.cu:415-422)
				const float err = (float)(sqrt(esum) / sqrt((double)p.n));
				const int it = st.iteration;
				p.errors[(size_t)b * (p.max_iter + 1) + it + 1] = err;
				st.iters_run += 1;
				const bool stop = p.stop_early && (((double)err < p.tol) || ((double)(float)fabs((double)err - (double)err_prev) < p.tol));
				err_prev = err;
				if (stop) st.done = 1;
				else { st.iteration = it + 1; if (it + 1 >= p.max_iter) st.done = 1; }
			}
			__syncthreads();
			if (st.done) break;
		}
		if (tid == 0) {
			p.iterations[b] = st.iteration;
			p.iterations_run[b] = st.iters_run;
			for (int k = 0; k < 9; k++) p.R[(size_t)b * 9 + k] = st.Rtot[k];
			for (int k = 0; k < 3; k++) p.t[(size_t)b * 3 + k] = st.ttot[k];
		}
		if (p.idx) {
#pragma unroll
			for (int s = 0; s < K9_S; s++) { const int i = s * K9_THREADS + tid; if (i < p.n) p.idx[(size_t)b * p.n + i] = idx[s]; }
		}
	}
}


// ------------------------------------------------------------------------------------------------
// K9F: the same loop with K1F's lower-bound filter in front of the exact chain (nn_filter.cu; derivation in
// DESIGN.md §K1F). Per registration the target is centred on its bounding box once (Xc, Yc, Zc, W = |q-c|^2 next to
// X, Y, Z in shared memory, Rq = max |q-c|); per iteration every source starts from the exact distance to its
// previous correspondence (strictly above it, so that match — or an equal one with a lower index — is found again),
// a 32-target sub-tile is skipped when the 3-FMA bound proves every chain value in it exceeds the running exact
// threshold, and the sub-tiles that cannot be excluded are evaluated by the whole warp with K1's packed chain. Every
// index therefore comes from the exact chain with the reference's tie rule: trajectories are bitwise those of the
// direct kernel above (tests/test_gpu_batched.py::test_batched_filter_is_bitwise_the_direct_kernel). Degenerate
// targets (non-finite coordinates, radius outside [1e-15, 1e15]) skip the filter: every sub-tile takes the exact pass.
// Sources, thresholds and indices live in per-thread shared-memory slots so that the inner loop fits 128 registers.
// ------------------------------------------------------------------------------------------------
constexpr float K9_U = 5.9604644775390625e-08f;            // 2^-24

template <int MODE, int TRK>
__global__ void __launch_bounds__(K9_THREADS, 2) icp_batched_filter_kernel(const BatchParams p)
{
	extern __shared__ __align__(16) unsigned char k9_smem[];
	constexpr int S = K9_S, T_ = K9_THREADS;
	const int mpad = ((p.m + 127) / 128) * 128;            // a multiple of every supported sub-tile size (host: mpad_f)
	const int stride = mpad + 4;                           // +4 floats: the quad prefetch of the last sub-tile stays in bounds
	float* tx  = reinterpret_cast<float*>(k9_smem);
	float* ty  = tx + stride;
	float* tz  = ty + stride;
	float* txc = tz + stride;
	float* tyc = txc + stride;
	float* tzc = tyc + stride;
	float* tw  = tzc + stride;
	const int tid = threadIdx.x;
	float* slots  = tw + stride;                           // [6][S][T_] per-thread slots
	float* thr_s  = slots + tid;
	int*   best_s = reinterpret_cast<int*>(slots + S * T_) + tid;
	float* ox_s   = slots + 2 * S * T_ + tid;
	float* oy_s   = slots + 3 * S * T_ + tid;
	float* oz_s   = slots + 4 * S * T_ + tid;
	int*   idx_s  = reinterpret_cast<int*>(slots + 5 * S * T_) + tid;
	__shared__ double red[K9_THREADS / 32];
	__shared__ double mom[16];
	__shared__ IterState st;
	__shared__ float err_prev;
	__shared__ int s_b;
	__shared__ float s_lo[K9_THREADS / 32][3], s_hi[K9_THREADS / 32][3];
	__shared__ unsigned s_r2[K9_THREADS / 32];
	__shared__ float s_ctr[3], s_rq;
	__shared__ int s_filter_ok;
	const float inf = __int_as_float(0x7f800000);
	const float one8u = 1.0f + 8.0f * K9_U;
	const int lane = tid & 31, warp = tid >> 5;

	while (true) {
		__syncthreads();
		if (tid == 0) s_b = k9_draw(p);
		__syncthreads();
		const int b = s_b;
		if (b >= p.batch) break;

		// ---- target: originals, bounding box -> centre, centred copy + W, radius bound ----
		const float* Tg = p.targets + (size_t)b * p.m * 3;
		float lo[3] = { inf, inf, inf }, hi[3] = { -inf, -inf, -inf };
		for (int j = tid; j < stride; j += T_) {
			const bool ok = j < p.m;
			const float x = ok ? Tg[3 * (size_t)j] : inf, y = ok ? Tg[3 * (size_t)j + 1] : inf, z = ok ? Tg[3 * (size_t)j + 2] : inf;
			tx[j] = x; ty[j] = y; tz[j] = z;
			if (ok) { lo[0] = fminf(lo[0], x); lo[1] = fminf(lo[1], y); lo[2] = fminf(lo[2], z); hi[0] = fmaxf(hi[0], x); hi[1] = fmaxf(hi[1], y); hi[2] = fmaxf(hi[2], z); }
		}
#pragma unroll
		for (int k = 0; k < 3; k++) {
			for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
			if (lane == 0) { s_lo[warp][k] = lo[k]; s_hi[warp][k] = hi[k]; }
		}
		__syncthreads();
		if (tid < 3) {
			float a = inf, c = -inf;
			for (int w = 0; w < K9_THREADS / 32; w++) { a = fminf(a, s_lo[w][tid]); c = fmaxf(c, s_hi[w][tid]); }
			float ctr = 0.5f * a + 0.5f * c;
			if (!isfinite(ctr)) ctr = 0.0f;
			s_ctr[tid] = ctr;
		}
		__syncthreads();
		const float cx = s_ctr[0], cy = s_ctr[1], cz = s_ctr[2];
		unsigned r2 = 0u;
		for (int j = tid; j < stride; j += T_) {
			float xc = 1e18f, yc = 1e18f, zc = 1e18f, w = 3e36f;
			if (j < p.m) {
				xc = __fsub_rn(tx[j], cx); yc = __fsub_rn(ty[j], cy); zc = __fsub_rn(tz[j], cz);
				w = __fmaf_rn(zc, zc, __fmaf_rn(xc, xc, __fmul_rn(yc, yc)));
				r2 = max(r2, __float_as_uint(w));          // w >= 0 or NaN: uint order = float order, NaN on top
			}
			txc[j] = xc; tyc[j] = yc; tzc[j] = zc; tw[j] = w;
		}
		for (int o = 16; o > 0; o >>= 1) r2 = max(r2, __shfl_xor_sync(0xffffffffu, r2, o));
		if (lane == 0) s_r2[warp] = r2;
		__syncthreads();
		if (tid == 0) {
			unsigned m2 = 0u;
			for (int w = 0; w < K9_THREADS / 32; w++) m2 = max(m2, s_r2[w]);
			const float rq = __fmul_ru(__fsqrt_ru(__uint_as_float(m2)), one8u);
			s_rq = rq;
			s_filter_ok = (isfinite(rq) && rq <= 1e15f && rq >= 1e-15f) ? 1 : 0;
			st.done = 0; st.iteration = 0; st.iters_run = 0; st.flags = p.flags;
			for (int k = 0; k < 9; k++) st.Rtot[k] = (k % 4 == 0) ? 1.0 : 0.0;
			for (int k = 0; k < 3; k++) st.ttot[k] = 0.0;
			err_prev = 0.f;
			p.errors[(size_t)b * (p.max_iter + 1)] = 0.f;
		}
		const float* Sg = p.sources + (size_t)b * p.n * 3;
#pragma unroll
		for (int s = 0; s < S; s++) {
			const int i = s * T_ + tid;
			const bool ok = i < p.n;
			ox_s[s * T_] = ok ? Sg[3 * (size_t)i] : 0.f; oy_s[s * T_] = ok ? Sg[3 * (size_t)i + 1] : 0.f; oz_s[s * T_] = ok ? Sg[3 * (size_t)i + 2] : 0.f;
			idx_s[s * T_] = 0;
		}
		__syncthreads();
		const float rq = s_rq;
		const bool filter_ok = s_filter_ok != 0;
		const float4* X4  = reinterpret_cast<const float4*>(tx);
		const float4* Y4  = reinterpret_cast<const float4*>(ty);
		const float4* Z4  = reinterpret_cast<const float4*>(tz);
		const float4* XC4 = reinterpret_cast<const float4*>(txc);
		const float4* YC4 = reinterpret_cast<const float4*>(tyc);
		const float4* ZC4 = reinterpret_cast<const float4*>(tzc);
		const float4* W4  = reinterpret_cast<const float4*>(tw);
		const int nsub = mpad / TRK;

		while (true) {
			// ---- matching ----
			float ax[S], ay[S], az[S], tau[S], kk[S];
#pragma unroll
			for (int s = 0; s < S; s++) {
				const float x = ox_s[s * T_], y = oy_s[s * T_], z = oz_s[s * T_];
				const float pcx = __fsub_rn(x, cx), pcy = __fsub_rn(y, cy), pcz = __fsub_rn(z, cz);
				ax[s] = -2.0f * pcx; ay[s] = -2.0f * pcy; az[s] = -2.0f * pcz;
				const float p2 = __fmaf_rn(pcz, pcz, __fmaf_rn(pcx, pcx, __fmul_rn(pcy, pcy)));
				const float p2lo = __fmul_rd(p2, 1.0f - 8.0f * K9_U);
				const float rp = __fmul_ru(__fsqrt_ru(p2), one8u);
				float e = __fmul_ru(8.0f * rq, rq);
				e = __fmaf_ru(10.0f * rp, rq, e);
				e = __fmaf_ru(2.0f * rp, rp, e);
				e = __fmul_ru(e, 1.05f * K9_U);
				kk[s] = __fsub_ru(e, p2lo);
				// warm start from the previous correspondence (index 0 in the first iteration: any index is an upper bound)
				float th = p.thr0;
				const int j0 = idx_s[s * T_];
				const float us = dist_chain(x, y, z, tx[j0], ty[j0], tz[j0]);
				const float up = (MODE == ICPB_DIST_SQRT) ? __fmul_ru(us, one8u) : us;
				const float nx = __uint_as_float(__float_as_uint(up) + 1u);
				if (us == us && nx < th) th = nx;
				thr_s[s * T_] = th; best_s[s * T_] = -1;
				tau[s] = __fadd_ru(__fmul_ru(th, one8u), kk[s]);
			}
#pragma unroll 1
			for (int sub = 0; sub < nsub; sub++) {
				const int j0 = sub * (TRK / 4), j1 = j0 + TRK / 4;
				unsigned need = (1u << S) - 1u;
				if (filter_ok) {
					float em[S];
#pragma unroll
					for (int s = 0; s < S; s++) em[s] = inf;
					float4 X = XC4[j0], Y = YC4[j0], Z = ZC4[j0], W = W4[j0];
#pragma unroll 2
					for (int j = j0; j < j1; j++) {
						const float4 Xn = XC4[j + 1], Yn = YC4[j + 1], Zn = ZC4[j + 1], Wn = W4[j + 1];
						const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
						const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
						const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
						const u64 w01 = pack2(W.x, W.y), w23 = pack2(W.z, W.w);
#pragma unroll
						for (int s = 0; s < S; s++) {
							const u64 AX = bcast2v(ax[s]), AY = bcast2v(ay[s]), AZ = bcast2v(az[s]);
							u64 e = fma2(AX, x01, fma2(AY, y01, fma2(AZ, z01, w01)));
							float a, c;
							unpack2(e, a, c);
							em[s] = min3(em[s], a, c);
							e = fma2(AX, x23, fma2(AY, y23, fma2(AZ, z23, w23)));
							unpack2(e, a, c);
							em[s] = min3(em[s], a, c);
						}
						X = Xn; Y = Yn; Z = Zn; W = Wn;
					}
					need = 0u;
#pragma unroll
					for (int s = 0; s < S; s++) {
						const unsigned bal = __ballot_sync(0xffffffffu, em[s] <= tau[s]);
						if (bal) need |= (1u << s);
					}
				}
				if (need) {
#pragma unroll
					for (int s = 0; s < S; s++) {
						if (need & (1u << s)) {          // warp-uniform
							const float sx = ox_s[s * T_], sy = oy_s[s * T_], sz = oz_s[s * T_];
							const float th = thr_s[s * T_];
							const u64 PX = pack2(sx, sx), PY = pack2(sy, sy), PZ = pack2(sz, sz);
							float mm = th;
#pragma unroll 4
							for (int j = j0; j < j1; j++) {
								const float4 Xo = X4[j], Yo = Y4[j], Zo = Z4[j];
								u64 dx = sub2(PX, pack2(Xo.x, Xo.y)), dy = sub2(PY, pack2(Yo.x, Yo.y)), dz = sub2(PZ, pack2(Zo.x, Zo.y));
								u64 d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
								float a, c;
								unpack2(d, a, c);
								mm = min3(mm, a, c);
								dx = sub2(PX, pack2(Xo.z, Xo.w)); dy = sub2(PY, pack2(Yo.z, Yo.w)); dz = sub2(PZ, pack2(Zo.z, Zo.w));
								d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
								unpack2(d, a, c);
								mm = min3(mm, a, c);
							}
							if (mm < th) {
								const float nt_ = lower_threshold<MODE>(mm);
								thr_s[s * T_] = nt_; best_s[s * T_] = sub;
								tau[s] = __fadd_ru(__fmul_ru(nt_, one8u), kk[s]);
							}
						}
					}
				}
			}
			// index recovery: first j of the remembered sub-tile that attains the minimum
#pragma unroll
			for (int s = 0; s < S; s++) {
				const int bs = best_s[s * T_];
				if (bs >= 0) {
					const float th = thr_s[s * T_];
					const float target = (MODE == ICPB_DIST_SQRT) ? __fsqrt_rn(th) : th;
					const float sx = ox_s[s * T_], sy = oy_s[s * T_], sz = oz_s[s * T_];
					const int base = bs * TRK;
					int found = -1;
					for (int j = 0; j < TRK && found < 0; j++) {
						float d = dist_chain(sx, sy, sz, tx[base + j], ty[base + j], tz[base + j]);
						if (MODE == ICPB_DIST_SQRT) d = __fsqrt_rn(d);
						if (d <= target) found = j;
					}
					if (found >= 0) idx_s[s * T_] = base + found;
				}
			}
			// ---- moments (FP64), SVD by thread 0 ----
			double acc[15];
#pragma unroll
			for (int k = 0; k < 15; k++) acc[k] = 0.0;
#pragma unroll
			for (int s = 0; s < S; s++) {
				if (s * T_ + tid < p.n) {
					const int j = idx_s[s * T_];
					const double x = ox_s[s * T_], y = oy_s[s * T_], z = oz_s[s * T_];
					const double qx = tx[j], qy = ty[j], qz = tz[j];
					acc[0] += x; acc[1] += y; acc[2] += z; acc[3] += qx; acc[4] += qy; acc[5] += qz;
					acc[6] += qx * x; acc[7] += qy * x; acc[8] += qz * x;
					acc[9] += qx * y; acc[10] += qy * y; acc[11] += qz * y;
					acc[12] += qx * z; acc[13] += qy * z; acc[14] += qz * z;
				}
			}
#pragma unroll
			for (int k = 0; k < 15; k++) { const double v = block_sum(acc[k], red); if (tid == 0) mom[k] = v; }
			if (tid == 0) {
				for (int k = 0; k < 15; k++) st.moments[k] = mom[k];
				st.moments[15] = (double)p.n;
				solve_p2p(&st);
			}
			__syncthreads();
			// ---- transform (RyT arithmetic) + residual against the same correspondences ----
			const float r0 = st.R[0], r1 = st.R[1], r2_ = st.R[2], r3 = st.R[3], r4 = st.R[4], r5 = st.R[5], r6 = st.R[6], r7 = st.R[7], r8 = st.R[8];
			const float t0 = st.T[0], t1 = st.T[1], t2 = st.T[2];
			double e = 0.0;
#pragma unroll
			for (int s = 0; s < S; s++) {
				const float x = ox_s[s * T_], y = oy_s[s * T_], z = oz_s[s * T_];
				const float nx = __fadd_rn(__fmaf_rn(r6, z, __fmaf_rn(r0, x, __fmul_rn(r3, y))), t0);
				const float ny = __fadd_rn(__fmaf_rn(r7, z, __fmaf_rn(r1, x, __fmul_rn(r4, y))), t1);
				const float nz = __fadd_rn(__fmaf_rn(r8, z, __fmaf_rn(r2_, x, __fmul_rn(r5, y))), t2);
				ox_s[s * T_] = nx; oy_s[s * T_] = ny; oz_s[s * T_] = nz;
				if (s * T_ + tid < p.n) {
					const int j = idx_s[s * T_];
					const float ex = __fsub_rn(nx, tx[j]), ey = __fsub_rn(ny, ty[j]), ez = __fsub_rn(nz, tz[j]);
					e += (double)ex * (double)ex + (double)ey * (double)ey + (double)ez * (double)ez;
				}
			}
			const double esum = block_sum(e, red);
			if (tid == 0) {
				const float err = (float)(sqrt(esum) / sqrt((double)p.n));
				const int it = st.iteration;
				p.errors[(size_t)b * (p.max_iter + 1) + it + 1] = err;
				st.iters_run += 1;
				const bool stop = p.stop_early && (((double)err < p.tol) || ((double)(float)fabs((double)err - (double)err_prev) < p.tol));
				err_prev = err;
				if (stop) st.done = 1;
				else { st.iteration = it + 1; if (it + 1 >= p.max_iter) st.done = 1; }
			}
			__syncthreads();
			if (st.done) break;
		}
		if (tid == 0) {
			p.iterations[b] = st.iteration;
			p.iterations_run[b] = st.iters_run;
			for (int k = 0; k < 9; k++) p.R[(size_t)b * 9 + k] = st.Rtot[k];
			for (int k = 0; k < 3; k++) p.t[(size_t)b * 3 + k] = st.ttot[k];
		}
		if (p.idx) {
#pragma unroll
			for (int s = 0; s < S; s++) { const int i = s * T_ + tid; if (i < p.n) p.idx[(size_t)b * p.n + i] = idx_s[s * T_]; }
		}
	}
}

} // namespace icpb

using namespace icpb;

static float sqrt_threshold_host(float sentinel)
{
	if (!(sentinel > 0.0f)) return 0.0f;
	float y = sentinel * sentinel;
	if (std::isinf(y)) return y;
	while (sqrtf(y) < sentinel) y = nextafterf(y, INFINITY);
	while (y > 0.0f && sqrtf(nextafterf(y, 0.0f)) >= sentinel) y = nextafterf(y, 0.0f);
	return y;
}

extern "C" int icpb_run_batched(icpb_ctx* ctx, const icpb_params* params, int batch, const float* sources, int n, const float* targets, int m,
                                float* errors, int* iterations, double* R, double* t, float* elapsed_ms)
{
	ICPB_NVTX("icpb_run_batched");
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = reinterpret_cast<Ctx*>(ctx);
	ICPB_CUDA(c, cudaSetDevice(c->device));
	auto bad = [&](const char* msg) { snprintf(c->err, sizeof c->err, "icpb_run_batched: %s", msg); return ICPB_ERR_BADARG; };
	if (!params || !sources || !targets || !errors || !iterations || !R || !t) return bad("NULL argument");
	if (batch < 1 || n < 1 || m < 1) return bad("empty batch");
	if (n > K9_MAX_N || m > K9_MAX_M) return bad("clouds larger than 2048 sources / 4096 targets: use icpb_run per pair");
	if (params->metric != ICPB_POINT_TO_POINT) return bad("only point-to-point is batched");
	if (params->dist_mode != ICPB_DIST_SQ && params->dist_mode != ICPB_DIST_SQRT) return bad("dist_mode must be SQ or SQRT");
	if (params->max_iter < 1 || params->max_iter > 4096) return bad("max_iter out of range");

	const size_t sb = sizeof(float) * 3 * (size_t)batch * n, tb = sizeof(float) * 3 * (size_t)batch * m;
	const size_t eb = sizeof(float) * (size_t)batch * (params->max_iter + 1);
	// Device buffers are kept in the context and only ever grow: a serving loop calls this with the same shapes over and
	// over, and cudaMalloc/cudaFree pairs cost more than the 25 MB upload of a 512-pair batch.
	K9Buffers& kb = c->k9;
	auto grow = [&](void** ptr, size_t& cap, size_t need) -> cudaError_t {
		if (need <= cap) return cudaSuccess;
		cudaFree(*ptr); *ptr = nullptr; cap = 0;
		const cudaError_t e = cudaMalloc(ptr, need);
		if (e == cudaSuccess) cap = need;
		return e;
	};
#define K9_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail_cuda(c, e__, #call, __FILE__, __LINE__); } while (0)
	constexpr int CHUNK = 32;                            // registrations per upload chunk (1.5 MB at 2048 points: ~60 us of PCIe)
	const int nchunks = (batch + CHUNK - 1) / CHUNK;
	// where do the clouds live? device pointers are used in place (the device-resident measurement), pinned host memory is
	// streamed in chunks behind the running kernel, pageable host memory is uploaded first (its copies are staged by the CPU)
	auto mem_type = [](const void* ptr) { cudaPointerAttributes a; if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return cudaMemoryTypeUnregistered; } return a.type; };
	const cudaMemoryType ts = mem_type(sources), tt_ = mem_type(targets);
	const bool resident = (ts == cudaMemoryTypeDevice || ts == cudaMemoryTypeManaged) && (tt_ == cudaMemoryTypeDevice || tt_ == cudaMemoryTypeManaged);
	bool streaming = !resident && ts == cudaMemoryTypeHost && tt_ == cudaMemoryTypeHost && batch > CHUNK;
	if (const char* e = getenv("ICPB_K9_STREAM")) if (atoi(e) == 0) streaming = false;
	if (!resident) { K9_TRY(grow((void**)&kb.s, kb.s_cap, sb)); K9_TRY(grow((void**)&kb.t, kb.t_cap, tb)); }
	K9_TRY(grow((void**)&kb.e, kb.e_cap, eb));
	K9_TRY(grow((void**)&kb.ints, kb.ints_cap, sizeof(int) * ((size_t)2 * batch + 8 + (size_t)nchunks)));
	K9_TRY(grow((void**)&kb.dbl, kb.dbl_cap, sizeof(double) * 12 * (size_t)batch));
	if (!kb.copy_stream) { K9_TRY(cudaStreamCreateWithFlags(&kb.copy_stream, cudaStreamNonBlocking)); K9_TRY(cudaEventCreateWithFlags(&kb.ev, cudaEventDisableTiming)); }
	if (!kb.one) { K9_TRY(cudaMallocHost((void**)&kb.one, 2 * sizeof(int))); kb.one[0] = 1; kb.one[1] = 0; }
	const float* d_s = resident ? sources : kb.s; const float* d_t = resident ? targets : kb.t;
	float* d_e = kb.e;
	int* d_it = kb.ints; int* d_run = kb.ints + batch; int* d_next = kb.ints + 2 * (size_t)batch; int* d_fail = d_next + 1; int* d_ready = d_next + 8;
	double* d_R = kb.dbl; double* d_tt = kb.dbl + 9 * (size_t)batch;
	K9_TRY(cudaMemsetAsync(d_e, 0, eb, c->stream));
	K9_TRY(cudaMemsetAsync(d_next, 0, sizeof(int) * (8 + (size_t)nchunks), c->stream));
	if (!resident && !streaming) {
		K9_TRY(cudaMemcpyAsync(kb.s, sources, sb, cudaMemcpyHostToDevice, c->stream));
		K9_TRY(cudaMemcpyAsync(kb.t, targets, tb, cudaMemcpyHostToDevice, c->stream));
	}

	BatchParams p;
	p.sources = d_s; p.targets = d_t; p.batch = batch; p.n = n; p.m = m;
	p.max_iter = params->max_iter; p.stop_early = params->stop_early; p.mode = params->dist_mode; p.flags = params->flags;
	p.sentinel = params->sentinel; p.tol = params->tol;
	p.thr0 = (params->dist_mode == ICPB_DIST_SQRT) ? sqrt_threshold_host(params->sentinel) : params->sentinel;
	p.errors = d_e; p.iterations = d_it; p.iterations_run = d_run; p.R = d_R; p.t = d_tt; p.idx = nullptr; p.next = d_next;
	p.ready = streaming ? d_ready : nullptr; p.chunk = CHUNK; p.fail = d_fail;
	const int mpad = ((m + K9_TRK - 1) / K9_TRK) * K9_TRK;
	// default: the filter kernel (K9F); ICPB_K9_FILTER=0 selects the direct kernel (every pair through the exact chain)
	bool use_filter = true;
	if (const char* e = getenv("ICPB_K9_FILTER")) use_filter = atoi(e) != 0;
	// K9F's sub-tile (targets per filter test / exact pass): 32 (measured on B200, 2368 pairs of 2048 points: 32 -> 25.5 ms,
	// 64 -> 27.1 ms, 128 -> 30.7 ms: finer sub-tiles send fewer irrelevant targets through the exact chain);
	// ICPB_K9_TRK=32|64|128 overrides. mpad is a multiple of 128 so that every choice tiles it.
	int trk = 32;
	if (const char* e = getenv("ICPB_K9_TRK")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128) trk = v; }
	const int mpad_f = ((m + 127) / 128) * 128;
	const size_t smem = use_filter ? sizeof(float) * (7 * (size_t)(mpad_f + 4) + 6 * (size_t)K9_S * K9_THREADS)
	                               : sizeof(float) * 3 * (size_t)(mpad + 4);
	const bool sq = params->dist_mode == ICPB_DIST_SQRT;
	auto kern = !use_filter ? (sq ? icp_batched_kernel<ICPB_DIST_SQRT> : icp_batched_kernel<ICPB_DIST_SQ>)
	          : trk == 64   ? (sq ? icp_batched_filter_kernel<ICPB_DIST_SQRT, 64> : icp_batched_filter_kernel<ICPB_DIST_SQ, 64>)
	          : trk == 128  ? (sq ? icp_batched_filter_kernel<ICPB_DIST_SQRT, 128> : icp_batched_filter_kernel<ICPB_DIST_SQ, 128>)
	                        : (sq ? icp_batched_filter_kernel<ICPB_DIST_SQRT, 32> : icp_batched_filter_kernel<ICPB_DIST_SQ, 32>);
	K9_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	K9_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K9_THREADS, smem));
	if (per_sm < 1) per_sm = 1;
	int grid = c->sm_count * per_sm;
	if (grid > batch) grid = batch;
	if (streaming) K9_TRY(cudaEventRecord(kb.ev, c->stream));        // the flags are zeroed before any chunk may set one
	K9_TRY(cudaEventRecord(c->ev[2], c->stream));
	kern<<<grid, K9_THREADS, smem, c->stream>>>(p);
	c->launches++;
	K9_TRY(cudaGetLastError());
	K9_TRY(cudaEventRecord(c->ev[3], c->stream));
	if (streaming) {
		// the kernel is already running (or queued) on the compute stream; the chunks follow on the copy stream, each one
		// trailed by its 4-byte "landed" flag. CTAs that draw a registration whose chunk is still in flight wait for it.
		K9_TRY(cudaStreamWaitEvent(kb.copy_stream, kb.ev, 0));
		for (int ch = 0; ch < nchunks; ch++) {
			const size_t b0 = (size_t)ch * CHUNK, cnt = (size_t)std::min(CHUNK, batch - ch * CHUNK);
			K9_TRY(cudaMemcpyAsync(kb.s + b0 * n * 3, sources + b0 * n * 3, sizeof(float) * 3 * cnt * n, cudaMemcpyHostToDevice, kb.copy_stream));
			K9_TRY(cudaMemcpyAsync(kb.t + b0 * m * 3, targets + b0 * m * 3, sizeof(float) * 3 * cnt * m, cudaMemcpyHostToDevice, kb.copy_stream));
			K9_TRY(cudaMemcpyAsync(d_ready + ch, kb.one, sizeof(int), cudaMemcpyHostToDevice, kb.copy_stream));
		}
	}
	int h_fail = 0;
	K9_TRY(cudaMemcpyAsync(errors, d_e, eb, cudaMemcpyDeviceToHost, c->stream));
	K9_TRY(cudaMemcpyAsync(iterations, d_it, sizeof(int) * batch, cudaMemcpyDeviceToHost, c->stream));
	K9_TRY(cudaMemcpyAsync(R, d_R, sizeof(double) * 9 * batch, cudaMemcpyDeviceToHost, c->stream));
	K9_TRY(cudaMemcpyAsync(t, d_tt, sizeof(double) * 3 * batch, cudaMemcpyDeviceToHost, c->stream));
	K9_TRY(cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	if (streaming) K9_TRY(cudaStreamSynchronize(kb.copy_stream));
	K9_TRY(cudaStreamSynchronize(c->stream));
	if (h_fail) { snprintf(c->err, sizeof c->err, "icpb_run_batched: a chunk of the streamed upload never arrived"); return ICPB_ERR_CUDA; }
	if (elapsed_ms) K9_TRY(cudaEventElapsedTime(elapsed_ms, c->ev[2], c->ev[3]));
#undef K9_TRY
	return ICPB_OK;
}
