// reduce_kernels.cu — K2 (moments), K3 (3x3 Jacobi SVD -> R,T), K7/K8 (6x6 normal equations, Cholesky,
// Euler -> R) and K4 (transform + residual) of the ICP iteration, plus the pack/unpack kernels.
//
// They replace, per iteration (reference file:line):
//   K2  Q_index<<<>>> + cublasSgemv x2 + deviation<<<>>> + cublasSgemm(3,3,N)   src/ICP_point_to_point.cu:308-357
//       (racy hand-written variant: centroid<<<>>>, src/ICP_standard.cu:41-91)
//   K3  cusolverDnSgesvd + cublasSgemm x2 + cublasScopy                         src/ICP_point_to_point.cu:369-397
//   K7  Q_index + Cxb<<<>>> + cublasSgemv(36,N) + cublasSgemv(6,N)              src/ICP_point_to_plane.cu:532-556
//   K8  cusolverDnSpotrf/Spotrs + host cos/sin -> R,T + H2D                     src/ICP_point_to_plane.cu:576-601
//   K4  RyT<<<>>> + cublasScopy + Scopy/Saxpy/Snrm2 + the host convergence test src/ICP_point_to_point.cu:403-421
//
// All four O(N) kernels are HBM/L2-bound streaming reductions: SoA float loads, one 16-byte gather
// of the matched target per point, per-thread FP64 accumulators, warp shuffles, one partial row per
// block and a fixed-order final sum by the last block to arrive (bitwise reproducible; no float
// atomics). Moments are raw sums (sum p, sum q, sum q p^T, N) in FP64 so that shards of several GPUs
// can be combined with one ncclAllReduce; centring is done afterwards: W = sum q p^T - N qbar pbar^T.
#include "common.cuh"
#include "solve_device.cuh"
#include "peer_exchange.cuh"

namespace icpb {

constexpr int RB = 256;   // threads per block of the reduction kernels

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}

// Block-wide sum of NV per-thread doubles; result row (NV values) valid in thread < NV of warp 0.
template <int NV> __device__ __forceinline__ void block_reduce(double (&v)[NV], double* smem /* [RB/32][NV] */, double* out_row)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < NV; k++) {
		double s = warp_sum(v[k]);
		if (lane == 0) smem[warp * NV + k] = s;
	}
	__syncthreads();
	if (threadIdx.x < NV) {
		double s = 0.0;
#pragma unroll
		for (int w = 0; w < RB / 32; w++) s += smem[w * NV + threadIdx.x];
		out_row[threadIdx.x] = s;
	}
}

// The last block to finish sums the per-block rows in block order (deterministic).
template <int NV> __device__ __forceinline__ bool last_block_sum(double* partials, int* ticket, double* dst)
{
	__shared__ int is_last;
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		const int t = atomicAdd(ticket, 1);
		is_last = (t == (int)gridDim.x - 1);
	}
	__syncthreads();
	if (!is_last) return false;
	__threadfence();
	if (threadIdx.x < NV) {
		double s = 0.0;
		for (unsigned b = 0; b < gridDim.x; b++) s += partials[(size_t)b * 32 + threadIdx.x];
		dst[threadIdx.x] = s;
	}
	if (threadIdx.x == 0) *ticket = 0;
	__syncthreads();
	return true;
}

__global__ void solve_kernel(IterState* st, int metric)
{
	if (st->done) return;
	if (threadIdx.x == 0) { if (metric == ICPB_POINT_TO_PLANE) solve_p2plane(st); else solve_p2p(st); }
}

// ------------------------------------------------------------------------------------------------
// K2 / K7: moment accumulation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int resolve_idx(const ReduceParams& p, int i)
{
	// A source whose every distance was >= sentinel keeps its previous correspondence
	// (src/ICP_point_to_point.cu:51-55 leaves idx[i] unwritten).
	const u64 key = p.keys[i];
	int j = p.idx[i];
	if (key != KEY_UNMATCHED) { j = (int)(uint32_t)(key & 0xffffffffull); p.idx[i] = j; p.seed[i] = j; }
	return j;
}

__global__ void __launch_bounds__(RB) moments_p2p_kernel(const ReduceParams p)
{
	if (p.st->done) return;
	__shared__ double red[(RB / 32) * 15];
	double acc[15];
#pragma unroll
	for (int k = 0; k < 15; k++) acc[k] = 0.0;
	for (int i = blockIdx.x * RB + threadIdx.x; i < p.n; i += gridDim.x * RB) {
		const int j = resolve_idx(p, i);
		const float4 q = __ldg(p.q4 + j);
		const double x = p.px[i], y = p.py[i], z = p.pz[i];
		const double qx = q.x, qy = q.y, qz = q.z;
		acc[0] += x; acc[1] += y; acc[2] += z;
		acc[3] += qx; acc[4] += qy; acc[5] += qz;
		acc[6] += qx * x; acc[7] += qy * x; acc[8] += qz * x;      // column 0 of sum q p^T (rows = q)
		acc[9] += qx * y; acc[10] += qy * y; acc[11] += qz * y;
		acc[12] += qx * z; acc[13] += qy * z; acc[14] += qz * z;
	}
	double* row = p.partials + (size_t)blockIdx.x * 32;
	block_reduce<15>(acc, red, row);
	if (last_block_sum<15>(p.partials, &p.st->ticket_a, p.st->moments)) {
		if (threadIdx.x == 0) p.st->moments[15] = (double)p.n;
		bool ok = true;
		if (p.peer.world > 1) { __syncthreads(); ok = peer_allreduce_block(p.peer, p.st->moments, 16); }
		if (threadIdx.x == 0) {
			if (!ok) { p.st->numeric_error = 100; p.st->done = 1; }
			else if (p.fuse_tail) solve_p2p(p.st);
		}
	}
}

// Cxb in registers: c = p x n, the 21 products of [c;n][c;n]^T, b = -[c;n]*((p-q).n) — float, in the
// reference's contraction order (src/ICP_point_to_plane.cu:200-233); sums in FP64.
__global__ void __launch_bounds__(RB) moments_p2plane_kernel(const ReduceParams p)
{
	if (p.st->done) return;
	__shared__ double red[(RB / 32) * 27];
	double acc[27];
#pragma unroll
	for (int k = 0; k < 27; k++) acc[k] = 0.0;
	for (int i = blockIdx.x * RB + threadIdx.x; i < p.n; i += gridDim.x * RB) {
		const int j = resolve_idx(p, i);
		const float4 q = __ldg(p.q4 + j), nn = __ldg(p.nrm4 + j);
		const float x = p.px[i], y = p.py[i], z = p.pz[i];
		float v[6];
		v[0] = __fmaf_rn(y, nn.z, -__fmul_rn(z, nn.y));
		v[1] = __fmaf_rn(z, nn.x, -__fmul_rn(x, nn.z));
		v[2] = __fmaf_rn(nn.y, x, -__fmul_rn(y, nn.x));
		v[3] = nn.x; v[4] = nn.y; v[5] = nn.z;
		int k = 0;
#pragma unroll
		for (int r = 0; r < 6; r++)
#pragma unroll
			for (int c = r; c < 6; c++) acc[k++] += (double)__fmul_rn(v[r], v[c]);
		const float d0 = __fsub_rn(x, q.x), d1 = __fsub_rn(y, q.y), d2 = __fsub_rn(z, q.z);
		const float aux = __fmaf_rn(nn.z, d2, __fmaf_rn(nn.x, d0, __fmul_rn(nn.y, d1)));
#pragma unroll
		for (int r = 0; r < 6; r++) acc[21 + r] += (double)__fmul_rn(v[r], -aux);
	}
	double* row = p.partials + (size_t)blockIdx.x * 32;
	block_reduce<27>(acc, red, row);
	if (last_block_sum<27>(p.partials, &p.st->ticket_a, p.st->moments)) {
		if (threadIdx.x == 0) p.st->moments[27] = (double)p.n;
		bool ok = true;
		if (p.peer.world > 1) { __syncthreads(); ok = peer_allreduce_block(p.peer, p.st->moments, 28); }
		if (threadIdx.x == 0) {
			if (!ok) { p.st->numeric_error = 100; p.st->done = 1; }
			else if (p.fuse_tail) solve_p2plane(p.st);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// K4: P <- R P + T with RyT's arithmetic, residual against the SAME correspondences, bookkeeping
// ------------------------------------------------------------------------------------------------
__device__ void finish_iteration(IterState* st, float* errors)
{
	// src/ICP_point_to_point.cu:415-422
	const float err = (float)(sqrt(st->err_sum) / sqrt(st->n_total));
	const int it = st->iteration;
	errors[it + 1] = err;
	st->last_err = err;
	st->iters_run += 1;
	const bool stop = st->stop_early && (((double)err < st->tol) ||
	                                     ((double)(float)fabs((double)err - (double)errors[it]) < st->tol));
	if (stop) { st->done = 1; return; }
	st->iteration = it + 1;
	if (it + 1 >= st->max_iter) st->done = 1;
}

__global__ void finish_kernel(IterState* st, float* errors)
{
	if (st->done) return;
	if (threadIdx.x == 0) finish_iteration(st, errors);
}

__global__ void __launch_bounds__(RB) transform_kernel(const ReduceParams p)
{
	if (p.st->done) return;
	__shared__ double red[(RB / 32) * 1];
	const float* R = p.st->R; const float* T = p.st->T;
	const float r0 = R[0], r1 = R[1], r2 = R[2], r3 = R[3], r4 = R[4], r5 = R[5], r6 = R[6], r7 = R[7], r8 = R[8];
	const float t0 = T[0], t1 = T[1], t2 = T[2];
	double acc[1] = { 0.0 };
	for (int i = blockIdx.x * RB + threadIdx.x; i < p.n; i += gridDim.x * RB) {
		const float x = p.px[i], y = p.py[i], z = p.pz[i];
		// RyT (src/ICP_point_to_point.cu:85-87) as nvcc schedules it: FMUL (R[.,1]*y), FFMA (R[.,0]*x), FFMA (R[.,2]*z), FADD T
		const float nx = __fadd_rn(__fmaf_rn(r6, z, __fmaf_rn(r0, x, __fmul_rn(r3, y))), t0);
		const float ny = __fadd_rn(__fmaf_rn(r7, z, __fmaf_rn(r1, x, __fmul_rn(r4, y))), t1);
		const float nz = __fadd_rn(__fmaf_rn(r8, z, __fmaf_rn(r2, x, __fmul_rn(r5, y))), t2);
		p.ox[i] = nx; p.oy[i] = ny; p.oz[i] = nz;
		const float4 q = __ldg(p.q4 + p.idx[i]);
		const float ex = __fsub_rn(nx, q.x), ey = __fsub_rn(ny, q.y), ez = __fsub_rn(nz, q.z);
		acc[0] += (double)ex * (double)ex + (double)ey * (double)ey + (double)ez * (double)ez;
		p.keys[i] = KEY_UNMATCHED;   // arm the next matching step
	}
	double* row = p.partials + (size_t)blockIdx.x * 32;
	block_reduce<1>(acc, red, row);
	if (last_block_sum<1>(p.partials, &p.st->ticket_b, &p.st->err_sum)) {
		bool ok = true;
		if (p.peer.world > 1) ok = peer_allreduce_block(p.peer, &p.st->err_sum, 1);
		if (threadIdx.x == 0) {
			if (!ok) { p.st->numeric_error = 100; p.st->done = 1; }
			else if (p.fuse_tail) finish_iteration(p.st, p.errors);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Layout kernels: AoS xyz <-> SoA, and the target re-tiling done once per registration
// ------------------------------------------------------------------------------------------------
__global__ void pack_source_kernel(const float* __restrict__ xyz, int n, int n_cap, float* px, float* py, float* pz, u64* keys, int* idx, int* seed, int reset_seed)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_cap) return;
	float x = 0.f, y = 0.f, z = 0.f;
	if (i < n) { x = xyz[3 * (size_t)i]; y = xyz[3 * (size_t)i + 1]; z = xyz[3 * (size_t)i + 2]; }
	px[i] = x; py[i] = y; pz[i] = z;
	keys[i] = KEY_UNMATCHED; idx[i] = 0;
	if (reset_seed) seed[i] = 0;
}
__global__ void unpack_source_kernel(const float* px, const float* py, const float* pz, int n, float* __restrict__ xyz)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	xyz[3 * (size_t)i] = px[i]; xyz[3 * (size_t)i + 1] = py[i]; xyz[3 * (size_t)i + 2] = pz[i];
}
__global__ void pack_target_kernel(const float* __restrict__ xyz, int m, int m_pad, float4* q4, float* qtiles)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m_pad) return;
	const float inf = __int_as_float(0x7f800000);
	float x = inf, y = inf, z = inf;
	if (j < m) { x = xyz[3 * (size_t)j]; y = xyz[3 * (size_t)j + 1]; z = xyz[3 * (size_t)j + 2]; q4[j] = make_float4(x, y, z, 0.f); }
	float* t = qtiles + (size_t)(j / K1_TT) * 3 * K1_TT + (j % K1_TT);
	t[0] = x; t[K1_TT] = y; t[2 * K1_TT] = z;
}

// after a stand-alone matching step: keys -> idx (+ the winning distance)
__global__ void resolve_kernel(const u64* keys, int* idx, int* seed, float* dmin, int n, float sentinel)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 key = keys[i];
	if (key != KEY_UNMATCHED) { idx[i] = (int)(uint32_t)(key & 0xffffffffull); seed[i] = idx[i]; dmin[i] = __uint_as_float((uint32_t)(key >> 32)); }
	else dmin[i] = sentinel;
}

// register-resident FFMA loop: the measured FP32 peak used as roofline denominator
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b)
{
	float v[8];
#pragma unroll
	for (int c = 0; c < 8; c++) v[c] = threadIdx.x * 0.001f + c;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int c = 0; c < 8; c++) v[c] = fmaf(v[c], a, b);
	}
	float s = 0;
#pragma unroll
	for (int c = 0; c < 8; c++) s += v[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static ReduceParams make_params(Ctx* c, int metric)
{
	ReduceParams p;
	p.px = c->px; p.py = c->py; p.pz = c->pz;
	p.ox = c->px; p.oy = c->py; p.oz = c->pz;
	p.q4 = c->q4; p.nrm4 = c->nrm4; p.keys = c->keys; p.idx = c->idx; p.seed = c->seed; p.n = c->n;
	p.partials = c->partials; p.st = c->st; p.errors = c->errors;
	p.peer = c->peer;
	p.fuse_tail = (c->world == 1 || c->peer.world > 1) ? 1 : 0;
	p.metric = metric;
	return p;
}
static int reduce_grid_for(Ctx* c)
{
	int g = (c->n + RB - 1) / RB;
	if (g > c->reduce_grid) g = c->reduce_grid;
	if (g < 1) g = 1;
	return g;
}

int launch_moments(Ctx* c, int metric)
{
	ReduceParams p = make_params(c, metric);
	const int g = reduce_grid_for(c);
	if (metric == ICPB_POINT_TO_PLANE) moments_p2plane_kernel<<<g, RB, 0, c->stream>>>(p);
	else moments_p2p_kernel<<<g, RB, 0, c->stream>>>(p);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}
int launch_solve(Ctx* c, int metric)
{
	solve_kernel<<<1, 32, 0, c->stream>>>(c->st, metric);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}
int launch_transform(Ctx* c)
{
	ReduceParams p = make_params(c, 0);
	transform_kernel<<<reduce_grid_for(c), RB, 0, c->stream>>>(p);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}
int launch_finish(Ctx* c)
{
	finish_kernel<<<1, 32, 0, c->stream>>>(c->st, c->errors);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}
int launch_resolve(Ctx* c, float sentinel)
{
	if (c->n <= 0) return ICPB_OK;
	resolve_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->keys, c->idx, c->seed, c->dmin, c->n, sentinel);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}
int launch_pack_source(Ctx* c, const float* d_xyz, int n, bool reset_seed)
{
	pack_source_kernel<<<(c->n_cap + 255) / 256, 256, 0, c->stream>>>(d_xyz, n, c->n_cap, c->px, c->py, c->pz, c->keys, c->idx, c->seed, reset_seed ? 1 : 0);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}
int launch_unpack_source(Ctx* c, float* d_xyz)
{
	if (c->n <= 0) return ICPB_OK;
	unpack_source_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->px, c->py, c->pz, c->n, d_xyz);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}
int launch_pack_target(Ctx* c, const float* d_xyz, int m)
{
	const int m_pad = c->nt * K1_TT;
	pack_target_kernel<<<(m_pad + 255) / 256, 256, 0, c->stream>>>(d_xyz, m, m_pad, c->q4, c->qtiles);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}
int launch_fp32_peak(Ctx* c, float* d_out, int iters, int blocks)
{
	fp32_peak_kernel<<<blocks, 256, 0, c->stream>>>(d_out, iters, 1.0001f, 0.5f);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}

} // namespace icpb
