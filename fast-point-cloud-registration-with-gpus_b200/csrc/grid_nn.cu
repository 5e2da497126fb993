// grid_nn.cu — K1g: exact nearest neighbour through a uniform grid over the (static) target cloud.
//
// Same contract as K1 (`Matching`, src/ICP_point_to_point.cu:31-57): idx[i] = argmin_j d(P_i,Q_j) with the
// reference's distance chain, the LOWEST index on ties, nothing for a source with no target below the
// sentinel — so the indices are identical to the brute-force kernel's, bit for bit.
//
//  build (once per target):  bounding box -> cell size h (<= ~4 cells per point, <= 16M cells) ->
//      counting sort of the targets by cell (histogram, exclusive scan, scatter); sorted points are stored
//      as float4 {x, y, z, original index}.
//  query (one thread per source): visit the cells of Chebyshev ring r = 0,1,2 around the source's cell,
//      whole x-rows at a time (a row is one contiguous range of the sorted array). Candidates are compared
//      as 64-bit keys (distance bits << 32 | index), i.e. by (distance, index) — independent of the visiting
//      order. After ring r every unvisited target is at least r*h away, so the search stops once
//      best < (r*h)^2 * 0.998 (the margin covers float rounding of the chain and of the cell assignment).
//  fallback: sources still open after ring GRID_RMAX (far from the cloud, e.g. the first ICP iterations)
//      are appended to a compact list and finished by the brute-force K1 kernel through its `remap`
//      path (device-side count, no host round trip). Both paths are exact, so their union is.
//  pyramid (default; grid_tree.cuh): an occupancy pyramid over the same cells (an octree whose leaves are the grid's
//      cells) descended best-first by one thread per source closes EVERY source whatever its distance — no ring limit,
//      no brute-force fallback (1 650 points + 130 nodes per source at the initial pose of the 1M x 1M registration,
//      55 + 41 near convergence, identical keys; the whole 45-iteration registration: 215 ms on a B200 against 1.27 s
//      with rings + fallback). ICPB_GRID_PYRAMID=0 selects the ring search described above.
#include "common.cuh"
#include "k1_device.cuh"
#include "grid_tree.cuh"
#include <cmath>
#include <cstdlib>

namespace icpb {

constexpr int GRID_RMAX = 2;

__device__ __forceinline__ unsigned f2ord(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void grid_bbox_kernel(const float4* __restrict__ q4, int m, unsigned* mm /* [6] min xyz, max xyz (ordered ints) */)
{
	float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
		const float4 q = q4[j];
		lo[0] = fminf(lo[0], q.x); lo[1] = fminf(lo[1], q.y); lo[2] = fminf(lo[2], q.z);
		hi[0] = fmaxf(hi[0], q.x); hi[1] = fmaxf(hi[1], q.y); hi[2] = fmaxf(hi[2], q.z);
	}
	for (int k = 0; k < 3; k++) {
		for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
		if ((threadIdx.x & 31) == 0) { atomicMin(mm + k, f2ord(lo[k])); atomicMax(mm + 3 + k, f2ord(hi[k])); }
	}
}

__device__ __forceinline__ int cell_coord(float v, float o, float inv_h) { return (int)floorf((v - o) * inv_h); }

__global__ void grid_count_kernel(const float4* __restrict__ q4, int m, GridGeom g, int* __restrict__ counts, int* __restrict__ cell_of)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m) return;
	const float4 q = q4[j];
	int cx = min(max(cell_coord(q.x, g.ox, g.inv_h), 0), g.nx - 1);
	int cy = min(max(cell_coord(q.y, g.oy, g.inv_h), 0), g.ny - 1);
	int cz = min(max(cell_coord(q.z, g.oz, g.inv_h), 0), g.nz - 1);
	const int c = cx + g.nx * (cy + g.ny * cz);
	cell_of[j] = c;
	atomicAdd(counts + c, 1);
}

// exclusive scan of `n` ints in three passes (1024 elements per block)
__global__ void scan_block_kernel(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ block_sums, int n)
{
	__shared__ int sh[1024];
	const int i = blockIdx.x * 1024 + threadIdx.x;
	const int v = (i < n) ? in[i] : 0;
	sh[threadIdx.x] = v;
	__syncthreads();
	for (int o = 1; o < 1024; o <<= 1) {
		const int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
		__syncthreads();
		sh[threadIdx.x] += t;
		__syncthreads();
	}
	if (i < n) out[i] = sh[threadIdx.x] - v;
	if (threadIdx.x == 1023) block_sums[blockIdx.x] = sh[1023];
}
__global__ void scan_sums_kernel(int* block_sums, int nb)
{
	// single block: sequential chunks of 1024 with a running carry
	__shared__ int sh[1024];
	__shared__ int carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < nb; base += 1024) {
		const int i = base + threadIdx.x;
		const int v = (i < nb) ? block_sums[i] : 0;
		sh[threadIdx.x] = v;
		__syncthreads();
		for (int o = 1; o < 1024; o <<= 1) {
			const int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
			__syncthreads();
			sh[threadIdx.x] += t;
			__syncthreads();
		}
		if (i < nb) block_sums[i] = carry + sh[threadIdx.x] - v;
		__syncthreads();
		if (threadIdx.x == 0) carry += sh[1023];
		__syncthreads();
	}
}
__global__ void scan_add_kernel(int* out, const int* __restrict__ block_sums, int n, int total_slot)
{
	const int i = blockIdx.x * 1024 + threadIdx.x;
	if (i < n) out[i] += block_sums[blockIdx.x];
	(void)total_slot;
}
__global__ void grid_scatter_kernel(const float4* __restrict__ q4, int m, const int* __restrict__ cell_of, const int* __restrict__ cell_start,
                                    int* __restrict__ fill, float4* __restrict__ sorted4)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m) return;
	const int c = cell_of[j];
	const int pos = cell_start[c] + atomicAdd(fill + c, 1);
	float4 q = q4[j];
	q.w = __int_as_float(j);
	sorted4[pos] = q;
}

template <int MODE>
__global__ void __launch_bounds__(128) grid_query_kernel(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz, int n,
                                                         const float4* __restrict__ sorted4, const int* __restrict__ cell_start, GridGeom g, float thr0,
                                                         u64* __restrict__ keys, int* __restrict__ open_list, int* __restrict__ open_count,
                                                         unsigned long long* __restrict__ visited, const int* done)
{
	if (done != nullptr && *done) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long seen = 0;
	if (i < n) {
		const float x = px[i], y = py[i], z = pz[i];
		const int cx = cell_coord(x, g.ox, g.inv_h), cy = cell_coord(y, g.oy, g.inv_h), cz = cell_coord(z, g.oz, g.inv_h);
		u64 best = KEY_UNMATCHED;
		bool resolved = false;
		const int rcover = max(max(max(cx, g.nx - 1 - cx), max(cy, g.ny - 1 - cy)), max(cz, g.nz - 1 - cz));   // ring that covers the whole grid
		for (int r = 0; r <= GRID_RMAX && !resolved; r++) {
			for (int dz = -r; dz <= r; dz++) {
				const int zc = cz + dz;
				if (zc < 0 || zc >= g.nz) continue;
				for (int dy = -r; dy <= r; dy++) {
					const int yc = cy + dy;
					if (yc < 0 || yc >= g.ny) continue;
					const bool shell = (abs(dz) == r) || (abs(dy) == r);
					// shell rows: every x in [cx-r, cx+r]; interior rows: only the two end cells
					for (int part = 0; part < (shell ? 1 : 2); part++) {
						int x0 = shell ? cx - r : (part == 0 ? cx - r : cx + r);
						int x1 = shell ? cx + r : x0;
						if (!shell && r == 0) continue;
						x0 = max(x0, 0); x1 = min(x1, g.nx - 1);
						if (x0 > x1) continue;
						const int row = g.nx * (yc + g.ny * zc);
						const int k0 = cell_start[row + x0], k1 = cell_start[row + x1 + 1];
						for (int k = k0; k < k1; k++) {
							const float4 q = __ldg(sorted4 + k);
							float d = dist_chain(x, y, z, q.x, q.y, q.z);
							if (d < thr0) {
								if (MODE == ICPB_DIST_SQRT) d = __fsqrt_rn(d);
								const u64 key = ((u64)__float_as_uint(d) << 32) | (u64)(uint32_t)__float_as_int(q.w);
								if (key < best) best = key;
							}
						}
						seen += (unsigned long long)(k1 - k0);
					}
				}
			}
			if (best != KEY_UNMATCHED) {
				float bd = __uint_as_float((uint32_t)(best >> 32));
				if (MODE == ICPB_DIST_SQRT) bd = bd * bd;
				const float reach = (float)r * g.h;
				if (bd < reach * reach * 0.998f) resolved = true;
			}
			if (r >= rcover) resolved = true;       // every cell of the grid has been visited
		}
		if (resolved) { if (best != KEY_UNMATCHED) keys[i] = best; }
		else open_list[atomicAdd(open_count, 1)] = i;
	}
	for (int o = 16; o > 0; o >>= 1) seen += __shfl_xor_sync(0xffffffffu, seen, o);
	if ((threadIdx.x & 31) == 0 && seen) atomicAdd(visited, seen);
}

// ---- occupancy pyramid -------------------------------------------------------------------------------------------
__global__ void pyr_level0_kernel(const int* __restrict__ cell_start, long long ncell, unsigned char* __restrict__ occ)
{
	const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c < ncell) occ[c] = cell_start[c + 1] > cell_start[c];
}
__global__ void pyr_up_kernel(const unsigned char* __restrict__ child, int nxc, int nyc, int nzc, unsigned char* __restrict__ parent, int nxp, int nyp, int nzp)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (long long)nxp * nyp * nzp) return;
	const int x = (int)(i % nxp), y = (int)((i / nxp) % nyp), z = (int)(i / ((long long)nxp * nyp));
	unsigned char o = 0;               // bit k = child k (x + 2y + 4z) is occupied: one load tells the descent which children to test
#pragma unroll
	for (int k = 0; k < 8; k++) {
		const int cx = 2 * x + (k & 1), cy = 2 * y + ((k >> 1) & 1), cz = 2 * z + ((k >> 2) & 1);
		if (cx < nxc && cy < nyc && cz < nzc && child[(long long)cx + (long long)nxc * ((long long)cy + (long long)nyc * cz)]) o |= (unsigned char)(1u << k);
	}
	parent[i] = o;
}

// One thread per source: best-first descent of the pyramid (grid_tree.cuh), warm-started from the previous
// correspondence when one exists (any real candidate is a valid starting key).
template <int MODE>
__global__ void __launch_bounds__(128) grid_tree_query_kernel(const float* __restrict__ px, const float* __restrict__ py_, const float* __restrict__ pz, int n,
                                                              const float4* __restrict__ sorted4, const int* __restrict__ cell_start, const float4* __restrict__ q4, int m,
                                                              const int* __restrict__ seed, GridGeom g, const GridPyramid pyr, float thr0,
                                                              u64* __restrict__ keys, unsigned long long* __restrict__ visited, const int* done)
{
	if (done != nullptr && *done) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long seen = 0;
	if (i < n) {
		const float x = px[i], y = py_[i], z = pz[i];
		u64 best = KEY_UNMATCHED;
		if (seed != nullptr) {
			const int j0 = seed[i];
			if (j0 >= 0 && j0 < m) {
				const float4 q = __ldg(q4 + j0);
				float d = dist_chain(x, y, z, q.x, q.y, q.z);
				if (d < thr0) {
					if (MODE == ICPB_DIST_SQRT) d = __fsqrt_rn(d);
					best = ((u64)__float_as_uint(d) << 32) | (u64)(uint32_t)j0;
				}
			}
		}
		best = grid_tree_nn<MODE>(x, y, z, g, pyr, cell_start, sorted4, thr0, best, &seen);
		if (best != KEY_UNMATCHED) keys[i] = best;
	}
	for (int o = 16; o > 0; o >>= 1) seen += __shfl_xor_sync(0xffffffffu, seen, o);
	if ((threadIdx.x & 31) == 0 && seen) atomicAdd(visited, seen);
}

int launch_match_brute_remap(Ctx* c, int dist_mode, float sentinel, const int* remap, const int* count_dev);

static int build_grid(Ctx* c)
{
	const int m = c->m;
	// every buffer of the build is kept in the context and only grows: cudaMalloc/cudaFree pairs cost more than the build
	auto grow = [&](void** ptr, size_t& cap, size_t need, size_t elem) -> cudaError_t {
		if (need <= cap && *ptr) return cudaSuccess;
		cudaFree(*ptr); *ptr = nullptr; cap = 0;
		const cudaError_t e = cudaMalloc(ptr, need * elem);
		if (e == cudaSuccess) cap = need;
		return e;
	};
	if (!c->grid_mm) ICPB_CUDA(c, cudaMalloc((void**)&c->grid_mm, 8 * sizeof(unsigned)));
	unsigned* mm = c->grid_mm;
	unsigned init[6] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u };
	ICPB_CUDA(c, cudaMemcpyAsync(mm, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
	grid_bbox_kernel<<<c->sm_count, 256, 0, c->stream>>>(c->q4, m, mm);
	c->launches++;
	unsigned h_mm[6];
	ICPB_CUDA(c, cudaMemcpyAsync(h_mm, mm, sizeof h_mm, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	auto dec = [](unsigned u) { unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &v, 4); return f; };
	float lo[3], hi[3];
	for (int k = 0; k < 3; k++) { lo[k] = dec(h_mm[k]); hi[k] = dec(h_mm[3 + k]); }
	const GridGeom g = compute_grid_geom(lo, hi, m);
	const size_t ncell = (size_t)g.nx * g.ny * g.nz;
	c->grid_dim[0] = g.nx; c->grid_dim[1] = g.ny; c->grid_dim[2] = g.nz;
	c->grid_origin[0] = g.ox; c->grid_origin[1] = g.oy; c->grid_origin[2] = g.oz; c->grid_cell = g.h;

	const int nb = (int)((ncell + 1 + 1023) / 1024);
	{
		size_t cap_cells = c->grid_cells_cap, cap_cells2 = c->grid_cells_cap, cap_pts = c->grid_pts_cap, cap_pts2 = c->grid_pts_cap;
		ICPB_CUDA(c, grow((void**)&c->grid_cell_start, cap_cells, ncell + 1, sizeof(int)));
		ICPB_CUDA(c, grow((void**)&c->grid_counts, cap_cells2, ncell + 1, sizeof(int)));
		c->grid_cells_cap = cap_cells < cap_cells2 ? cap_cells : cap_cells2;
		ICPB_CUDA(c, grow((void**)&c->grid_sorted4, cap_pts, (size_t)m, sizeof(float4)));
		ICPB_CUDA(c, grow((void**)&c->grid_cell_of, cap_pts2, (size_t)m, sizeof(int)));
		c->grid_pts_cap = cap_pts < cap_pts2 ? cap_pts : cap_pts2;
		ICPB_CUDA(c, grow((void**)&c->grid_sums, c->grid_sums_cap, (size_t)nb, sizeof(int)));
	}
	int *counts = c->grid_counts, *cell_of = c->grid_cell_of, *sums = c->grid_sums;
	ICPB_CUDA(c, cudaMemsetAsync(counts, 0, sizeof(int) * (ncell + 1), c->stream));
	grid_count_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(c->q4, m, g, counts, cell_of);
	scan_block_kernel<<<nb, 1024, 0, c->stream>>>(counts, c->grid_cell_start, sums, (int)(ncell + 1));
	scan_sums_kernel<<<1, 1024, 0, c->stream>>>(sums, nb);
	scan_add_kernel<<<nb, 1024, 0, c->stream>>>(c->grid_cell_start, sums, (int)(ncell + 1), 0);
	ICPB_CUDA(c, cudaMemsetAsync(counts, 0, sizeof(int) * (ncell + 1), c->stream));
	grid_scatter_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(c->q4, m, cell_of, c->grid_cell_start, counts, c->grid_sorted4);
	c->launches += 5;
	ICPB_CUDA(c, cudaGetLastError());
	if (!c->grid_open_list || c->grid_open_cap < c->n_cap) {
		cudaFree(c->grid_open_list); c->grid_open_list = nullptr;
		ICPB_CUDA(c, cudaMalloc((void**)&c->grid_open_list, sizeof(int) * (size_t)(c->n_cap > 0 ? c->n_cap : 1)));
		c->grid_open_cap = c->n_cap;
	}
	if (!c->grid_counters) ICPB_CUDA(c, cudaMalloc((void**)&c->grid_counters, 4 * sizeof(unsigned long long)));
	ICPB_CUDA(c, cudaMemsetAsync(c->grid_counters, 0, 4 * sizeof(unsigned long long), c->stream));
	if (c->grid_pyramid) {
		GridPyramid& py = c->grid_py;
		const long long total = pyramid_layout(g, py);
		ICPB_CUDA(c, grow((void**)&c->grid_occ, c->grid_occ_cap, (size_t)total, 1));
		py.occ = c->grid_occ;
		pyr_level0_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, c->stream>>>(c->grid_cell_start, (long long)ncell, c->grid_occ);
		c->launches++;
		for (int L = 1; L < py.levels; L++) {
			const long long np = (long long)py.nx[L] * py.ny[L] * py.nz[L];
			pyr_up_kernel<<<(unsigned)((np + 255) / 256), 256, 0, c->stream>>>(c->grid_occ + py.off[L - 1], py.nx[L - 1], py.ny[L - 1], py.nz[L - 1],
			                                                                    c->grid_occ + py.off[L], py.nx[L], py.ny[L], py.nz[L]);
			c->launches++;
		}
		ICPB_CUDA(c, cudaGetLastError());
		ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	c->grid_ready = true;
	return ICPB_OK;
}

// K5 through the pyramid (ICPB_KNN_PYRAMID=1): one thread per target point, k+1 nearest targets by best-first descent.
template <int MODE>
__global__ void __launch_bounds__(128) knn_tree_kernel(const float4* __restrict__ q4, int m, int k1, const float4* __restrict__ sorted4,
                                                       const int* __restrict__ cell_start, GridGeom g, const GridPyramid pyr, int* __restrict__ nbr)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const float4 me = q4[i];
	float kd[GT_KMAX]; int ki[GT_KMAX];
	grid_tree_knn<MODE>(me.x, me.y, me.z, k1, g, pyr, cell_start, sorted4, kd, ki);
	for (int p = 0; p < k1; p++) nbr[(size_t)i * k1 + p] = ki[p] >= 0 ? ki[p] : 0;      // `minimum` returns 0 when nothing is below 10000
}

int launch_knn_tree(Ctx* c, int k1, int knn_dist_mode, int* nbr)
{
	if (k1 < 1 || k1 > GT_KMAX || !c->grid_pyramid) return ICPB_ERR_BADARG;
	int rc;
	if (!c->grid_ready) { if ((rc = build_grid(c)) != ICPB_OK) return rc; }
	GridGeom g;
	g.h = c->grid_cell; g.inv_h = (float)(1.0 / (double)c->grid_cell);
	g.ox = c->grid_origin[0]; g.oy = c->grid_origin[1]; g.oz = c->grid_origin[2];
	g.nx = c->grid_dim[0]; g.ny = c->grid_dim[1]; g.nz = c->grid_dim[2];
	auto kern = (knn_dist_mode == ICPB_DIST_SQRT) ? knn_tree_kernel<ICPB_DIST_SQRT> : knn_tree_kernel<ICPB_DIST_SQ>;
	kern<<<(c->m + 127) / 128, 128, 0, c->stream>>>(c->q4, c->m, k1, c->grid_sorted4, c->grid_cell_start, g, c->grid_py, nbr);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}

// Builds the grid of the current target OUTSIDE the timed loop (icpb_run calls this before it records its first event,
// as it does for the filter data; launch_match_grid still builds lazily for the step-wise API).
int prepare_match_grid(Ctx* c)
{
	if (c->m <= 0 || (c->grid_ready && c->grid_open_cap >= c->n_cap)) return ICPB_OK;
	return build_grid(c);
}

int launch_match_grid(Ctx* c, int dist_mode, float sentinel)
{
	if (dist_mode == ICPB_DIST_STD) { snprintf(c->err, sizeof c->err, "ICPB_NN_GRID supports ICPB_DIST_SQ and ICPB_DIST_SQRT"); return ICPB_ERR_BADARG; }
	if (c->n <= 0 || c->m <= 0) return ICPB_OK;
	int rc;
	if (!c->grid_ready || c->grid_open_cap < c->n_cap) { if ((rc = build_grid(c)) != ICPB_OK) return rc; }
	GridGeom g;
	g.h = c->grid_cell; g.inv_h = (float)(1.0 / (double)c->grid_cell);
	g.ox = c->grid_origin[0]; g.oy = c->grid_origin[1]; g.oz = c->grid_origin[2];
	g.nx = c->grid_dim[0]; g.ny = c->grid_dim[1]; g.nz = c->grid_dim[2];
	float thr0 = sentinel;
	if (dist_mode == ICPB_DIST_SQRT) {
		float y = sentinel * sentinel;
		if (!std::isinf(y)) {
			while (sqrtf(y) < sentinel) y = nextafterf(y, INFINITY);
			while (y > 0.0f && sqrtf(nextafterf(y, 0.0f)) >= sentinel) y = nextafterf(y, 0.0f);
		}
		thr0 = y;
	}
	if (c->grid_pyramid) {
		unsigned long long* visited_t = c->grid_counters + 1;
		ICPB_CUDA(c, cudaMemsetAsync(c->grid_counters, 0, sizeof(unsigned long long), c->stream));      // no open sources in this mode
		auto kt = (dist_mode == ICPB_DIST_SQRT) ? grid_tree_query_kernel<ICPB_DIST_SQRT> : grid_tree_query_kernel<ICPB_DIST_SQ>;
		kt<<<(c->n + 127) / 128, 128, 0, c->stream>>>(c->px, c->py, c->pz, c->n, c->grid_sorted4, c->grid_cell_start, c->q4, c->m,
		                                               c->kf_use_seed ? c->seed : nullptr, g, c->grid_py, thr0, c->keys, visited_t, &c->st->done);
		c->launches++;
		ICPB_CUDA(c, cudaGetLastError());
		c->pairs_acc += (double)c->n * (double)c->m;
		return ICPB_OK;
	}
	int* open_count = reinterpret_cast<int*>(c->grid_counters);                    // [0] low word: open sources of this pass
	unsigned long long* visited = c->grid_counters + 1;                           // [1] candidates visited (cumulative)
	ICPB_CUDA(c, cudaMemsetAsync(open_count, 0, sizeof(unsigned long long), c->stream));
	auto kern = (dist_mode == ICPB_DIST_SQRT) ? grid_query_kernel<ICPB_DIST_SQRT> : grid_query_kernel<ICPB_DIST_SQ>;
	kern<<<(c->n + 127) / 128, 128, 0, c->stream>>>(c->px, c->py, c->pz, c->n, c->grid_sorted4, c->grid_cell_start, g, thr0, c->keys,
	                                                 c->grid_open_list, open_count, visited, &c->st->done);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	c->pairs_acc += (double)c->n * (double)c->m;
	// sources the grid could not close within GRID_RMAX rings: exact brute force over that compact list
	return launch_match_brute_remap(c, dist_mode, sentinel, c->grid_open_list, open_count);
}

} // namespace icpb

extern "C" int icpb_get_grid_stats(icpb_ctx* ctx, double* candidates_visited, int* last_open_sources, int dims[3], float* cell)
{
	using namespace icpb;
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = reinterpret_cast<Ctx*>(ctx);
	ICPB_CUDA(c, cudaSetDevice(c->device));
	if (!c->grid_ready) { snprintf(c->err, sizeof c->err, "no grid: run a matching step with ICPB_NN_GRID first"); return ICPB_ERR_STATE; }
	unsigned long long h[2];
	ICPB_CUDA(c, cudaMemcpyAsync(h, c->grid_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (candidates_visited) *candidates_visited = (double)h[1];
	if (last_open_sources) *last_open_sources = (int)(h[0] & 0xffffffffull);
	if (dims) { dims[0] = c->grid_dim[0]; dims[1] = c->grid_dim[1]; dims[2] = c->grid_dim[2]; }
	if (cell) *cell = c->grid_cell;
	return ICPB_OK;
}
