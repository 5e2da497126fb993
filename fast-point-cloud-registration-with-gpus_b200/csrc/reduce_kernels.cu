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
constexpr int RP = 4;     // consecutive points per thread and pass: 16-byte loads of the SoA arrays, 4 gathers in flight

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

// The last block to finish sums the per-block rows in a FIXED order (deterministic: bitwise reproducible results for a
// given point count). ncu on the r1 kernels showed where their time went: a 100k-point launch took 33 us and a 1M-point
// one 55 us — a ~30 us floor that was this sum done by NV threads over up to 592 rows, one dependent L2 round trip
// after the other. Here all RB threads take part: the rows are dealt to RB / NVP groups (NVP = NV rounded up to a power
// of two, lane = value index), every group adds its rows in ascending order with 8 independent loads in flight, and the
// group sums are then added in group order.
template <int NV> __device__ __forceinline__ bool last_block_sum(double* partials, int* ticket, double* dst)
{
	constexpr int NVP = NV <= 1 ? 1 : (NV <= 16 ? 16 : 32);
	constexpr int GROUPS = RB / NVP;
	static_assert(NV <= 32, "one lane per value");
	__shared__ int is_last;
	__shared__ double gsum[GROUPS * NVP];
	if (gridDim.x == 1) {             // small clouds: the block's row is the sum — no ticket, no fences, no second trip to L2
		if (threadIdx.x < NV) dst[threadIdx.x] = partials[threadIdx.x];      // written by this same thread in block_reduce
		__syncthreads();
		return true;
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		const int t = atomicAdd(ticket, 1);
		is_last = (t == (int)gridDim.x - 1);
	}
	__syncthreads();
	if (!is_last) return false;
	__threadfence();
	const int grp = threadIdx.x / NVP, v = threadIdx.x % NVP;
	double s = 0.0;
	if (v < NV) {
		const unsigned nrows = gridDim.x;
		unsigned b = grp;
		for (; b + 7 * GROUPS < nrows; b += 8 * GROUPS) {
			double t[8];
#pragma unroll
			for (int k = 0; k < 8; k++) t[k] = __ldcg(partials + (size_t)(b + k * GROUPS) * 32 + v);
#pragma unroll
			for (int k = 0; k < 8; k++) s += t[k];
		}
		for (; b < nrows; b += GROUPS) s += __ldcg(partials + (size_t)b * 32 + v);
	}
	gsum[grp * NVP + v] = s;
	__syncthreads();
	if (threadIdx.x < NV) {
		double t = 0.0;
#pragma unroll 8
		for (int g = 0; g < GROUPS; g++) t += gsum[g * NVP + threadIdx.x];
		dst[threadIdx.x] = t;
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
// Correspondences of RP consecutive points: keys -> idx (+ seed), then the gathers. A source whose every distance was
// >= sentinel keeps its previous correspondence (src/ICP_point_to_point.cu:51-55 leaves idx[i] unwritten) — or, with
// ICPB_FLAG_REJECT_UNMATCHED, is marked -1 and dropped from the sums.
struct Corr4 { int j[RP]; bool use[RP]; };

__device__ __forceinline__ Corr4 resolve4(const ReduceParams& p, int i0, bool reject)
{
	Corr4 c;
	const ulonglong2 k01 = *reinterpret_cast<const ulonglong2*>(p.keys + i0), k23 = *reinterpret_cast<const ulonglong2*>(p.keys + i0 + 2);
	const u64 key[RP] = { k01.x, k01.y, k23.x, k23.y };
	const int4 old = *reinterpret_cast<const int4*>(p.idx + i0);
	const int prev[RP] = { old.x, old.y, old.z, old.w };
	int out[RP];
	bool any_new = false;
#pragma unroll
	for (int r = 0; r < RP; r++) {
		const bool in = i0 + r < p.n;
		const bool hit = key[r] != KEY_UNMATCHED;
		int j = hit ? (int)(uint32_t)(key[r] & 0xffffffffull) : (reject ? -1 : prev[r]);
		out[r] = in ? j : prev[r];
		c.j[r] = j; c.use[r] = in && j >= 0;
		any_new |= in && hit;
	}
	if (any_new || reject) {
		*reinterpret_cast<int4*>(p.idx + i0) = make_int4(out[0], out[1], out[2], out[3]);
		const int4 sd = *reinterpret_cast<const int4*>(p.seed + i0);
		int so[RP] = { sd.x, sd.y, sd.z, sd.w };
#pragma unroll
		for (int r = 0; r < RP; r++) if (i0 + r < p.n && key[r] != KEY_UNMATCHED) so[r] = out[r];
		*reinterpret_cast<int4*>(p.seed + i0) = make_int4(so[0], so[1], so[2], so[3]);
	}
	return c;
}

__global__ void __launch_bounds__(RB) moments_p2p_kernel(const ReduceParams p)
{
	if (p.st->done) return;
	__shared__ double red[(RB / 32) * 16];
	double acc[16];
#pragma unroll
	for (int k = 0; k < 16; k++) acc[k] = 0.0;
	const bool reject = (p.flags & ICPB_FLAG_REJECT_UNMATCHED) != 0;
	for (int i0 = RP * (blockIdx.x * RB + threadIdx.x); i0 < p.n; i0 += RP * gridDim.x * RB) {
		const Corr4 c = resolve4(p, i0, reject);
		float4 q[RP];
#pragma unroll
		for (int r = 0; r < RP; r++) q[r] = c.use[r] ? __ldg(p.q4 + c.j[r]) : make_float4(0.f, 0.f, 0.f, 0.f);
		const float4 X = *reinterpret_cast<const float4*>(p.px + i0), Y = *reinterpret_cast<const float4*>(p.py + i0), Z = *reinterpret_cast<const float4*>(p.pz + i0);
		const float xs[RP] = { X.x, X.y, X.z, X.w }, ys[RP] = { Y.x, Y.y, Y.z, Y.w }, zs[RP] = { Z.x, Z.y, Z.z, Z.w };
#pragma unroll
		for (int r = 0; r < RP; r++) {
			if (!c.use[r]) continue;
			const double x = xs[r], y = ys[r], z = zs[r];
			const double qx = q[r].x, qy = q[r].y, qz = q[r].z;
			acc[0] += x; acc[1] += y; acc[2] += z;
			acc[3] += qx; acc[4] += qy; acc[5] += qz;
			acc[6] += qx * x; acc[7] += qy * x; acc[8] += qz * x;      // column 0 of sum q p^T (rows = q)
			acc[9] += qx * y; acc[10] += qy * y; acc[11] += qz * y;
			acc[12] += qx * z; acc[13] += qy * z; acc[14] += qz * z;
			acc[15] += 1.0;                                             // N: the points that entered the sums
		}
	}
	double* row = p.partials + (size_t)blockIdx.x * 32;
	block_reduce<16>(acc, red, row);
	if (last_block_sum<16>(p.partials, &p.st->ticket_a, p.st->moments)) {
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
	__shared__ double red[(RB / 32) * 28];
	double acc[28];
#pragma unroll
	for (int k = 0; k < 28; k++) acc[k] = 0.0;
	const bool reject = (p.flags & ICPB_FLAG_REJECT_UNMATCHED) != 0;
	for (int i0 = RP * (blockIdx.x * RB + threadIdx.x); i0 < p.n; i0 += RP * gridDim.x * RB) {
		const Corr4 c = resolve4(p, i0, reject);
		float4 q[RP], nv[RP];
#pragma unroll
		for (int r = 0; r < RP; r++) {
			q[r] = c.use[r] ? __ldg(p.q4 + c.j[r]) : make_float4(0.f, 0.f, 0.f, 0.f);
			nv[r] = c.use[r] ? __ldg(p.nrm4 + c.j[r]) : make_float4(0.f, 0.f, 0.f, 0.f);
		}
		const float4 X = *reinterpret_cast<const float4*>(p.px + i0), Y = *reinterpret_cast<const float4*>(p.py + i0), Z = *reinterpret_cast<const float4*>(p.pz + i0);
		const float xs[RP] = { X.x, X.y, X.z, X.w }, ys[RP] = { Y.x, Y.y, Y.z, Y.w }, zs[RP] = { Z.x, Z.y, Z.z, Z.w };
#pragma unroll
		for (int r = 0; r < RP; r++) {
			if (!c.use[r]) continue;
			const float x = xs[r], y = ys[r], z = zs[r];
			const float4 nn = nv[r];
			float v[6];
			v[0] = __fmaf_rn(y, nn.z, -__fmul_rn(z, nn.y));
			v[1] = __fmaf_rn(z, nn.x, -__fmul_rn(x, nn.z));
			v[2] = __fmaf_rn(nn.y, x, -__fmul_rn(y, nn.x));
			v[3] = nn.x; v[4] = nn.y; v[5] = nn.z;
			int k = 0;
#pragma unroll
			for (int a = 0; a < 6; a++)
#pragma unroll
				for (int b = a; b < 6; b++) acc[k++] += (double)__fmul_rn(v[a], v[b]);
			const float d0 = __fsub_rn(x, q[r].x), d1 = __fsub_rn(y, q[r].y), d2 = __fsub_rn(z, q[r].z);
			const float aux = __fmaf_rn(nn.z, d2, __fmaf_rn(nn.x, d0, __fmul_rn(nn.y, d1)));
#pragma unroll
			for (int a = 0; a < 6; a++) acc[21 + a] += (double)__fmul_rn(v[a], -aux);
			acc[27] += 1.0;
		}
	}
	double* row = p.partials + (size_t)blockIdx.x * 32;
	block_reduce<28>(acc, red, row);
	if (last_block_sum<28>(p.partials, &p.st->ticket_a, p.st->moments)) {
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
	// src/ICP_point_to_point.cu:415-422; with ICPB_FLAG_REJECT_UNMATCHED the RMS is over the points that were matched
	const double npts = (st->flags & ICPB_FLAG_REJECT_UNMATCHED) ? fmax(st->moments[st->count_slot], 1.0) : st->n_total;
	const float err = (float)(sqrt(st->err_sum) / sqrt(npts));
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
	double acc[1] = { 0.0 };
	const int i_first = RP * (blockIdx.x * RB + threadIdx.x);
	// everything that does not depend on this iteration's R,T is loaded first (with programmatic dependent launch this
	// part overlaps the tail of the moments kernel: the last-block sum and the 3x3 / 6x6 solve)
	float4 X = make_float4(0.f, 0.f, 0.f, 0.f), Y = X, Z = X;
	int4 J = make_int4(-1, -1, -1, -1);
	if (i_first < p.n) {
		X = *reinterpret_cast<const float4*>(p.px + i_first); Y = *reinterpret_cast<const float4*>(p.py + i_first); Z = *reinterpret_cast<const float4*>(p.pz + i_first);
		J = *reinterpret_cast<const int4*>(p.idx + i_first);
	}
	const float* R = p.st->R; const float* T = p.st->T;
	const float r0 = R[0], r1 = R[1], r2 = R[2], r3 = R[3], r4 = R[4], r5 = R[5], r6 = R[6], r7 = R[7], r8 = R[8];
	const float t0 = T[0], t1 = T[1], t2 = T[2];
	for (int i0 = i_first; i0 < p.n; i0 += RP * gridDim.x * RB) {
		if (i0 != i_first) {
			X = *reinterpret_cast<const float4*>(p.px + i0); Y = *reinterpret_cast<const float4*>(p.py + i0); Z = *reinterpret_cast<const float4*>(p.pz + i0);
			J = *reinterpret_cast<const int4*>(p.idx + i0);
		}
		const int js[RP] = { J.x, J.y, J.z, J.w };
		float4 q[RP];
#pragma unroll
		for (int r = 0; r < RP; r++) q[r] = (i0 + r < p.n && js[r] >= 0) ? __ldg(p.q4 + js[r]) : make_float4(0.f, 0.f, 0.f, 0.f);
		const float xs[RP] = { X.x, X.y, X.z, X.w }, ys[RP] = { Y.x, Y.y, Y.z, Y.w }, zs[RP] = { Z.x, Z.y, Z.z, Z.w };
		float ox[RP], oy[RP], oz[RP];
#pragma unroll
		for (int r = 0; r < RP; r++) {
			const float x = xs[r], y = ys[r], z = zs[r];
			// RyT (src/ICP_point_to_point.cu:85-87) as nvcc schedules it: FMUL (R[.,1]*y), FFMA (R[.,0]*x), FFMA (R[.,2]*z), FADD T
			ox[r] = __fadd_rn(__fmaf_rn(r6, z, __fmaf_rn(r0, x, __fmul_rn(r3, y))), t0);
			oy[r] = __fadd_rn(__fmaf_rn(r7, z, __fmaf_rn(r1, x, __fmul_rn(r4, y))), t1);
			oz[r] = __fadd_rn(__fmaf_rn(r8, z, __fmaf_rn(r2, x, __fmul_rn(r5, y))), t2);
			if (i0 + r < p.n && js[r] >= 0) {
				const float ex = __fsub_rn(ox[r], q[r].x), ey = __fsub_rn(oy[r], q[r].y), ez = __fsub_rn(oz[r], q[r].z);
				acc[0] += (double)ex * (double)ex + (double)ey * (double)ey + (double)ez * (double)ez;
			}
		}
		// the arrays are padded to a block multiple: whole 16-byte stores, also across the end of the cloud (the padding
		// holds transformed zeros, which nothing reads)
		*reinterpret_cast<float4*>(p.ox + i0) = make_float4(ox[0], ox[1], ox[2], ox[3]);
		*reinterpret_cast<float4*>(p.oy + i0) = make_float4(oy[0], oy[1], oy[2], oy[3]);
		*reinterpret_cast<float4*>(p.oz + i0) = make_float4(oz[0], oz[1], oz[2], oz[3]);
		const ulonglong2 un = make_ulonglong2(KEY_UNMATCHED, KEY_UNMATCHED);     // arm the next matching step
		*reinterpret_cast<ulonglong2*>(p.keys + i0) = un; *reinterpret_cast<ulonglong2*>(p.keys + i0 + 2) = un;
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
__global__ void pack_source_kernel(const float* __restrict__ xyz, int n, int n_cap, float* px, float* py, float* pz, u64* keys, int* idx, int* seed, int reset_seed, int m)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_cap) return;
	float x = 0.f, y = 0.f, z = 0.f;
	if (i < n) { x = xyz[3 * (size_t)i]; y = xyz[3 * (size_t)i + 1]; z = xyz[3 * (size_t)i + 2]; }
	px[i] = x; py[i] = y; pz[i] = z;
	keys[i] = KEY_UNMATCHED; idx[i] = 0;
	// a fresh warm-start seed: ANY target index is a valid upper bound of the minimum, so this only affects the speed of the
	// first (cold) matching pass. The target at the same relative position in index order is a far better guess than
	// target 0 for scans of the same sensor / moved copies of one cloud, and as good as any other otherwise.
	if (reset_seed) seed[i] = (m > 0 && n > 0 && i < n) ? (int)(((long long)i * m) / n) : 0;
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
// Is the cloud just uploaded bit for bit the target already packed? (a host-driven loop re-uploads an unchanged target at
// every step; everything derived from it — tiles, grid, normals, policy state — can then stay)
__global__ void target_same_kernel(const float* __restrict__ xyz, const float4* __restrict__ q4, int m, int* differ)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m) return;
	const float4 q = q4[j];
	if (__float_as_uint(xyz[3 * (size_t)j]) != __float_as_uint(q.x) || __float_as_uint(xyz[3 * (size_t)j + 1]) != __float_as_uint(q.y) ||
	    __float_as_uint(xyz[3 * (size_t)j + 2]) != __float_as_uint(q.z)) *differ = 1;
}

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
	p.flags = c->run_flags;
	return p;
}
static int reduce_grid_for(Ctx* c)
{
	int g = (c->n + RB * RP - 1) / (RB * RP);      // RP consecutive points per thread and pass
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
	pack_source_kernel<<<(c->n_cap + 255) / 256, 256, 0, c->stream>>>(d_xyz, n, c->n_cap, c->px, c->py, c->pz, c->keys, c->idx, c->seed, reset_seed ? 1 : 0, c->m);
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
int launch_target_same(Ctx* c, const float* d_xyz, int m, int* d_differ)
{
	target_same_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(d_xyz, c->q4, m, d_differ);
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
