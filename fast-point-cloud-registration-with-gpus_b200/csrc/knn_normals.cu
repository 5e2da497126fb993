// knn_normals.cu — K5 (exact k-nearest neighbours of every target point) and K6 (PCA normals).
//
// Replaces, for point-to-plane ICP (reference file:line):
//   K5  knn<<<>>> + minimum()                      src/ICP_point_to_plane.cu:30-70, called at :406 with P = Q = target
//   K6  Normals<<<>>> steps 1-2 + the host loop of LAPACKE_ssyev / cblas_isamin   src/ICP_point_to_plane.cu:80-101, 431-439
//
// The reference writes an M x M float distance matrix (40 GB at 100k points) and then runs k+1 linear
// argmin scans per point, invalidating each winner with 10000.0. The result of that procedure is simply
// the k+1 smallest entries under the order (sqrt.rn distance, index): that is what K5 computes, in one
// pass, with a per-thread sorted candidate list in registers and no matrix.
//   * Candidates are kept by (squared distance, index) — exact floats, no sqrt in the loop. Because
//     sqrt.rn can merge different squares into one float, KNN_CAND = 8 candidates are kept for k+1 <= 7
//     results and the selection is certified afterwards: if sqrt(cand[7]) > sqrt(cand[k]) every point
//     that could tie with the k-th result is among the candidates and re-ranking them by
//     (sqrt, index) is exact. A query that fails the certificate (rare) is redone by knn_exact_kernel,
//     a literal restatement of the reference procedure, one warp per query.
//   * Tiles are visited outwards from the block's own tile (clouds are stored in scan order, so the
//     true neighbours are met first and almost every later quad fails the `min4 <= worst` test).
//   * Distances use the packed FADD2/FMUL2/FFMA2 chain of K1 (same rounding as the reference's
//     (float)sqrt(dx*dx+dy*dy+dz*dz) before the sqrt).
// K6: per target point the centroid of neighbours 1..k (each term divided by (float)k, then added, in
// order), the 6 upper-triangle covariance sums with FFMA in the reference's order, a cyclic Jacobi
// eigen-solve in FP64 registers, eigenvector of the eigenvalue of smallest magnitude (first on ties,
// eigenvalues ascending — cblas_isamin on ssyev's output).
#include "common.cuh"
#include <climits>

namespace icpb {

constexpr int KNN_CAND    = 8;
constexpr int KNN_THREADS = 256;

__device__ __forceinline__ u64 kpack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void kunpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 ksub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 kmul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 kfma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__device__ __forceinline__ float kchain(float xp, float yp, float zp, float xq, float yq, float zq)
{
	const float dx = __fsub_rn(xp, xq), dy = __fsub_rn(yp, yq), dz = __fsub_rn(zp, zq);
	return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ bool key_less(float d, int i, float d2, int i2) { return (d < d2) || (d == d2 && i < i2); }

__device__ __forceinline__ void cand_insert(float (&kd)[KNN_CAND], int (&ki)[KNN_CAND], float d, int idx)
{
	kd[KNN_CAND - 1] = d; ki[KNN_CAND - 1] = idx;
#pragma unroll
	for (int p = KNN_CAND - 1; p > 0; p--) {
		if (key_less(kd[p], ki[p], kd[p - 1], ki[p - 1])) {
			const float td = kd[p]; kd[p] = kd[p - 1]; kd[p - 1] = td;
			const int ti = ki[p]; ki[p] = ki[p - 1]; ki[p - 1] = ti;
		}
	}
}

// SQRT_RANK: the canonical program ranks sqrt.rn distances (src/ICP_point_to_plane.cu:54-57); the dataset programs and the
// "clean" variant rank the squared chain itself (src/CUDA/GPU_point_to_plane_bunny.cu:63,72), for which the candidate
// list ordered by (d^2, index) already is the answer and no certificate is needed.
template <bool SQRT_RANK>
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const float* __restrict__ qtiles, const float4* __restrict__ q4, int m, int nt, int k1,
                                                           int* __restrict__ nbr, int* __restrict__ flags)
{
	__shared__ __align__(16) float tile[3 * K1_TT];
	const int tid = threadIdx.x;
	const int i = blockIdx.x * KNN_THREADS + tid;
	const bool valid = i < m;
	const float4 me = valid ? q4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
	const u64 PX = kpack2(me.x, me.x), PY = kpack2(me.y, me.y), PZ = kpack2(me.z, me.z);
	float kd[KNN_CAND]; int ki[KNN_CAND];
	const float inf = __int_as_float(0x7f800000);
#pragma unroll
	for (int p = 0; p < KNN_CAND; p++) { kd[p] = inf; ki[p] = INT_MAX; }

	int home = (blockIdx.x * KNN_THREADS + KNN_THREADS / 2) / K1_TT;
	if (home > nt - 1) home = nt - 1;
	int lo = home - 1, hi = home + 1, t = home;
	bool up = true;
	for (int step = 0; step < nt; step++) {
		__syncthreads();
		const float4* src = reinterpret_cast<const float4*>(qtiles + (size_t)t * 3 * K1_TT);
		float4* dst = reinterpret_cast<float4*>(tile);
		for (int q = tid; q < 3 * K1_TT / 4; q += KNN_THREADS) dst[q] = __ldg(src + q);
		__syncthreads();
		const float4* X4 = reinterpret_cast<const float4*>(tile);
		const float4* Y4 = X4 + K1_TT / 4;
		const float4* Z4 = Y4 + K1_TT / 4;
		const int base = t * K1_TT;
#pragma unroll 4
		for (int j = 0; j < K1_TT / 4; j++) {
			const float4 X = X4[j], Y = Y4[j], Z = Z4[j];
			u64 dx = ksub2(PX, kpack2(X.x, X.y)), dy = ksub2(PY, kpack2(Y.x, Y.y)), dz = ksub2(PZ, kpack2(Z.x, Z.y));
			u64 d01 = kfma2(dz, dz, kfma2(dx, dx, kmul2(dy, dy)));
			dx = ksub2(PX, kpack2(X.z, X.w)); dy = ksub2(PY, kpack2(Y.z, Y.w)); dz = ksub2(PZ, kpack2(Z.z, Z.w));
			u64 d23 = kfma2(dz, dz, kfma2(dx, dx, kmul2(dy, dy)));
			float a, b, c, d;
			kunpack2(d01, a, b); kunpack2(d23, c, d);
			const float mn = fminf(fminf(a, b), fminf(c, d));
			if (mn <= kd[KNN_CAND - 1]) {
				const int j0 = base + 4 * j;
				if (j0 < m && key_less(a, j0, kd[KNN_CAND - 1], ki[KNN_CAND - 1])) cand_insert(kd, ki, a, j0);
				if (j0 + 1 < m && key_less(b, j0 + 1, kd[KNN_CAND - 1], ki[KNN_CAND - 1])) cand_insert(kd, ki, b, j0 + 1);
				if (j0 + 2 < m && key_less(c, j0 + 2, kd[KNN_CAND - 1], ki[KNN_CAND - 1])) cand_insert(kd, ki, c, j0 + 2);
				if (j0 + 3 < m && key_less(d, j0 + 3, kd[KNN_CAND - 1], ki[KNN_CAND - 1])) cand_insert(kd, ki, d, j0 + 3);
			}
		}
		// next tile, alternating outwards from `home`
		if (up) { if (hi < nt) t = hi++; else t = lo--; }
		else    { if (lo >= 0) t = lo--; else t = hi++; }
		up = !up;
	}
	if (!valid) return;

	if constexpr (!SQRT_RANK) {
		for (int p = 0; p < k1; p++) nbr[(size_t)i * k1 + p] = (kd[p] < 10000.0f && ki[p] != INT_MAX) ? ki[p] : 0;
		flags[i] = 0;
	} else {
	// re-rank the candidates by (sqrt.rn distance, index), certify, emit
	float ks[KNN_CAND];
#pragma unroll
	for (int p = 0; p < KNN_CAND; p++) ks[p] = __fsqrt_rn(kd[p]);
	const bool complete = (ki[KNN_CAND - 1] == INT_MAX);              // fewer than KNN_CAND points exist: all are candidates
	const bool certified = complete || (ks[KNN_CAND - 1] > ks[k1 - 1]);
#pragma unroll
	for (int a = 1; a < KNN_CAND; a++) {
#pragma unroll
		for (int p = KNN_CAND - 1; p > 0; p--) {
			if (p >= a && key_less(ks[p], ki[p], ks[p - 1], ki[p - 1])) {
				const float td = ks[p]; ks[p] = ks[p - 1]; ks[p - 1] = td;
				const int ti = ki[p]; ki[p] = ki[p - 1]; ki[p - 1] = ti;
			}
		}
	}
	for (int p = 0; p < k1; p++) nbr[(size_t)i * k1 + p] = (ks[p] < 10000.0f && ki[p] != INT_MAX) ? ki[p] : 0;
	flags[i] = certified ? 0 : 1;
	}
}

// Literal restatement of the reference procedure (src/ICP_point_to_plane.cu:30-44,61-69) for the queries
// the fast kernel could not certify: k1 passes, each the argmin over sqrt'ed distances with strict `<`
// from 10000.0 (lowest index on ties), previous winners excluded. One warp per query.
__global__ void knn_exact_kernel(const float4* __restrict__ q4, int m, int k1, int* __restrict__ nbr, const int* __restrict__ flags)
{
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int i = warp; i < m; i += nwarps) {
		if (!flags[i]) continue;
		const float4 me = q4[i];
		int chosen[KNN_CAND];
		for (int pass = 0; pass < k1; pass++) {
			float bd = 10000.0f; int bj = INT_MAX;
			for (int j = lane; j < m; j += 32) {
				bool skip = false;
				for (int c = 0; c < pass; c++) skip |= (chosen[c] == j);
				if (skip) continue;
				const float4 q = __ldg(q4 + j);
				const float d = __fsqrt_rn(kchain(me.x, me.y, me.z, q.x, q.y, q.z));
				if (d < 10000.0f && key_less(d, j, bd, bj)) { bd = d; bj = j; }
			}
			for (int o = 16; o > 0; o >>= 1) {
				const float od = __shfl_xor_sync(0xffffffffu, bd, o);
				const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
				if (key_less(od, oj, bd, bj)) { bd = od; bj = oj; }
			}
			chosen[pass] = (bj == INT_MAX) ? 0 : bj;
			if (lane == 0) nbr[(size_t)i * k1 + pass] = chosen[pass];
			if (bj == INT_MAX) chosen[pass] = -1;     // nothing was invalidated except slot 0, which stays selectable in the reference only if < 10000: it is not
		}
	}
}

// ------------------------------------------------------------------------------------------------
// K6: PCA normal
// ------------------------------------------------------------------------------------------------
__device__ void eig3_smallest(const float A[6] /* xx xy xz yy yz zz */, float n[3])
{
	double a[3][3] = { { A[0], A[1], A[2] }, { A[1], A[3], A[4] }, { A[2], A[4], A[5] } };
	double v[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
	for (int sweep = 0; sweep < 60; sweep++) {
		const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
		const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
		if (off <= 1e-18 * diag || off == 0.0) break;
#pragma unroll
		for (int pq = 0; pq < 3; pq++) {
			const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
			if (a[p][q] == 0.0) continue;
			const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
			const double t = ((theta >= 0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
			const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
			for (int k = 0; k < 3; k++) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
#pragma unroll
			for (int k = 0; k < 3; k++) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
#pragma unroll
			for (int k = 0; k < 3; k++) { const double x = v[k][p], y = v[k][q]; v[k][p] = c * x - s * y; v[k][q] = s * x + c * y; }
		}
	}
	// ascending eigenvalues (ssyev), then the first one of smallest |w| in float (cblas_isamin)
	int ord[3] = { 0, 1, 2 };
	for (int x = 0; x < 2; x++)
		for (int y = x + 1; y < 3; y++)
			if (a[ord[y]][ord[y]] < a[ord[x]][ord[x]]) { const int t = ord[x]; ord[x] = ord[y]; ord[y] = t; }
	int im = 0;
	for (int c = 1; c < 3; c++)
		if (fabsf((float)a[ord[c]][ord[c]]) < fabsf((float)a[ord[im]][ord[im]])) im = c;
	const int col = ord[im];
	n[0] = (float)v[0][col]; n[1] = (float)v[1][col]; n[2] = (float)v[2][col];
}

__global__ void __launch_bounds__(128) normals_kernel(const float4* __restrict__ q4, const int* __restrict__ nbr, int m, int k, float4* __restrict__ nrm4)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const float fk = (float)k;
	float bx = 0.f, by = 0.f, bz = 0.f;
	for (int j = 1; j < k + 1; j++) {
		const float4 q = __ldg(q4 + nbr[(size_t)i * (k + 1) + j]);
		bx = __fadd_rn(bx, __fdiv_rn(q.x, fk)); by = __fadd_rn(by, __fdiv_rn(q.y, fk)); bz = __fadd_rn(bz, __fdiv_rn(q.z, fk));
	}
	float A[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
	for (int j = 1; j < k + 1; j++) {
		const float4 q = __ldg(q4 + nbr[(size_t)i * (k + 1) + j]);
		const float dx = __fsub_rn(q.x, bx), dy = __fsub_rn(q.y, by), dz = __fsub_rn(q.z, bz);
		A[0] = __fmaf_rn(dx, dx, A[0]); A[1] = __fmaf_rn(dx, dy, A[1]); A[2] = __fmaf_rn(dx, dz, A[2]);
		A[3] = __fmaf_rn(dy, dy, A[3]); A[4] = __fmaf_rn(dy, dz, A[4]); A[5] = __fmaf_rn(dz, dz, A[5]);
	}
	float n[3];
	eig3_smallest(A, n);
	nrm4[i] = make_float4(n[0], n[1], n[2], 0.f);
}

__global__ void normals_pack_kernel(const float* __restrict__ xyz, int m, float4* __restrict__ nrm4)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m) nrm4[i] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.f);
}
__global__ void normals_unpack_kernel(const float4* __restrict__ nrm4, int m, float* __restrict__ xyz)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m) { const float4 v = nrm4[i]; xyz[3 * (size_t)i] = v.x; xyz[3 * (size_t)i + 1] = v.y; xyz[3 * (size_t)i + 2] = v.z; }
}

} // namespace icpb

using namespace icpb;
static inline Ctx* C(icpb_ctx* p) { return reinterpret_cast<Ctx*>(p); }

template <typename T> static int kn_alloc(Ctx* c, T** p, size_t count)
{
	if (*p) { cudaFree(*p); *p = nullptr; }
	cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
	if (e != cudaSuccess) { fail_cuda(c, e, "cudaMalloc", __FILE__, __LINE__); return ICPB_ERR_NOMEM; }
	return ICPB_OK;
}
static int kn_stage(Ctx* c, size_t bytes)
{
	if (bytes <= c->stage_cap) return ICPB_OK;
	int rc = kn_alloc(c, reinterpret_cast<unsigned char**>(&c->stage_xyz), bytes);
	c->stage_cap = rc == ICPB_OK ? bytes : 0;
	return rc;
}

extern "C" {

int icpb_estimate_normals(icpb_ctx* ctx, int k, float* elapsed_ms) { return icpb_estimate_normals_ex(ctx, k, ICPB_DIST_SQRT, elapsed_ms); }

int icpb_estimate_normals_ex(icpb_ctx* ctx, int k, int knn_dist_mode, float* elapsed_ms)
{
	ICPB_NVTX("icpb_estimate_normals");
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = C(ctx);
	if (knn_dist_mode != ICPB_DIST_SQ && knn_dist_mode != ICPB_DIST_SQRT) { snprintf(c->err, sizeof c->err, "icpb_estimate_normals_ex: knn_dist_mode must be ICPB_DIST_SQ or ICPB_DIST_SQRT"); return ICPB_ERR_BADARG; }
	ICPB_CUDA(c, cudaSetDevice(c->device));
	if (c->m <= 0) { snprintf(c->err, sizeof c->err, "icpb_estimate_normals: set the target first"); return ICPB_ERR_STATE; }
	if (k < 1 || k + 1 >= KNN_CAND) { snprintf(c->err, sizeof c->err, "icpb_estimate_normals: k must be in [1,%d]", KNN_CAND - 2); return ICPB_ERR_BADARG; }
	int rc;
	const int k1 = k + 1;
	if (!c->nbr || c->knn_k != k) { if ((rc = kn_alloc(c, &c->nbr, (size_t)c->m * k1)) != ICPB_OK) return rc; }
	if (!c->nrm4) { if ((rc = kn_alloc(c, &c->nrm4, (size_t)c->m)) != ICPB_OK) return rc; }
	if ((rc = kn_stage(c, sizeof(int) * (size_t)c->m)) != ICPB_OK) return rc;
	int* flags = reinterpret_cast<int*>(c->stage_xyz);
	ICPB_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
	if (c->knn_pyramid && c->grid_pyramid) {
		if ((rc = launch_knn_tree(c, k1, knn_dist_mode, c->nbr)) != ICPB_OK) return rc;
		c->launches--;                      // counted by the launcher; the common increment below stays
	} else if (knn_dist_mode == ICPB_DIST_SQRT) {
		knn_kernel<true><<<(c->m + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, c->stream>>>(c->qtiles, c->q4, c->m, c->nt, k1, c->nbr, flags);
		c->launches++;
		ICPB_CUDA(c, cudaGetLastError());
		knn_exact_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->q4, c->m, k1, c->nbr, flags);
	} else {
		knn_kernel<false><<<(c->m + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, c->stream>>>(c->qtiles, c->q4, c->m, c->nt, k1, c->nbr, flags);
	}
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	normals_kernel<<<(c->m + 127) / 128, 128, 0, c->stream>>>(c->q4, c->nbr, c->m, k, c->nrm4);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	ICPB_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (elapsed_ms) ICPB_CUDA(c, cudaEventElapsedTime(elapsed_ms, c->ev[2], c->ev[3]));
	c->knn_k = k; c->have_normals = true;
	return ICPB_OK;
}

int icpb_get_neighbors(icpb_ctx* ctx, int* nbr, int on_device)
{
	if (!ctx || !nbr) return ICPB_ERR_BADARG;
	Ctx* c = C(ctx);
	ICPB_CUDA(c, cudaSetDevice(c->device));
	if (!c->nbr || c->knn_k <= 0) { snprintf(c->err, sizeof c->err, "no neighbour lists: call icpb_estimate_normals"); return ICPB_ERR_STATE; }
	ICPB_CUDA(c, cudaMemcpyAsync(nbr, c->nbr, sizeof(int) * (size_t)c->m * (c->knn_k + 1), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return ICPB_OK;
}

int icpb_get_normals(icpb_ctx* ctx, float* normals, int on_device)
{
	if (!ctx || !normals) return ICPB_ERR_BADARG;
	Ctx* c = C(ctx);
	ICPB_CUDA(c, cudaSetDevice(c->device));
	if (!c->have_normals) { snprintf(c->err, sizeof c->err, "no normals: call icpb_estimate_normals or icpb_set_normals"); return ICPB_ERR_STATE; }
	int rc;
	float* dst = normals;
	if (!on_device) { if ((rc = kn_stage(c, sizeof(float) * 3 * (size_t)c->m)) != ICPB_OK) return rc; dst = c->stage_xyz; }
	normals_unpack_kernel<<<(c->m + 255) / 256, 256, 0, c->stream>>>(c->nrm4, c->m, dst);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	if (!on_device) ICPB_CUDA(c, cudaMemcpyAsync(normals, dst, sizeof(float) * 3 * (size_t)c->m, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return ICPB_OK;
}

int icpb_set_normals(icpb_ctx* ctx, const float* normals, int on_device)
{
	if (!ctx || !normals) return ICPB_ERR_BADARG;
	Ctx* c = C(ctx);
	ICPB_CUDA(c, cudaSetDevice(c->device));
	if (c->m <= 0) { snprintf(c->err, sizeof c->err, "icpb_set_normals: set the target first"); return ICPB_ERR_STATE; }
	int rc;
	if (!c->nrm4) { if ((rc = kn_alloc(c, &c->nrm4, (size_t)c->m)) != ICPB_OK) return rc; }
	const float* src = normals;
	if (!on_device) {
		if ((rc = kn_stage(c, sizeof(float) * 3 * (size_t)c->m)) != ICPB_OK) return rc;
		ICPB_CUDA(c, cudaMemcpyAsync(c->stage_xyz, normals, sizeof(float) * 3 * (size_t)c->m, cudaMemcpyHostToDevice, c->stream));
		src = c->stage_xyz;
	}
	normals_pack_kernel<<<(c->m + 255) / 256, 256, 0, c->stream>>>(src, c->m, c->nrm4);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	c->have_normals = true;
	return ICPB_OK;
}

} // extern "C"
