// engine.cu — the C ABI (include/icp_b200.h): context, cloud upload, the ICP loop on one stream.
//
// Host-side mirror of the reference mains' allocation block and `while (iteration < MAX_ITER)` loop
// (src/ICP_point_to_point.cu:203-285, 295-423; src/ICP_standard.cu:369-463; src/ICP_point_to_plane.cu:517-631):
// same step order (matching -> minimisation -> transformation -> error -> convergence test) but
//   * every buffer is allocated once per cloud size, nothing inside the loop;
//   * no cudaDeviceSynchronize between kernels: one non-blocking stream, the convergence test runs on
//     the device (finish_iteration) and raises IterState::done; later kernels see the flag and return,
//     so the host may enqueue `sync_every` iterations between reads of the flag without changing results;
//   * 3 kernels per point-to-point iteration on one GPU (match, moments+SVD, transform+error+test).
#include "common.cuh"
#include <cmath>
#include <cstdlib>
#include <new>
#include <chrono>
#include <thread>

namespace icpb {

int fail_cuda(Ctx* c, cudaError_t e, const char* what, const char* file, int line)
{
	if (c) snprintf(c->err, sizeof c->err, "CUDA error '%s' in %s (%s:%d)", cudaGetErrorString(e), what, file, line);
	return ICPB_ERR_CUDA;
}
static int fail(Ctx* c, int code, const char* msg)
{
	if (c) snprintf(c->err, sizeof c->err, "%s", msg);
	return code;
}

template <typename T> static int dev_alloc(Ctx* c, T** p, size_t count)
{
	if (*p) { cudaFree(*p); *p = nullptr; }
	if (count == 0) return ICPB_OK;
	cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
	if (e != cudaSuccess) { fail_cuda(c, e, "cudaMalloc", __FILE__, __LINE__); return ICPB_ERR_NOMEM; }
	return ICPB_OK;
}
static int ensure_stage(Ctx* c, size_t bytes)
{
	if (bytes <= c->stage_cap) return ICPB_OK;
	int rc = dev_alloc(c, reinterpret_cast<unsigned char**>(&c->stage_xyz), bytes);
	c->stage_cap = rc == ICPB_OK ? bytes : 0;
	return rc;
}
static int ensure_errors(Ctx* c, int count)
{
	if (count <= c->err_cap) return ICPB_OK;
	int rc = dev_alloc(c, &c->errors, (size_t)count);
	if (rc != ICPB_OK) return rc;
	c->graph_gen++;                     // captured kernels hold the old c->errors pointer
	if (c->errors_host) cudaFreeHost(c->errors_host);
	ICPB_CUDA(c, cudaMallocHost((void**)&c->errors_host, sizeof(float) * (size_t)count));
	c->err_cap = count;
	return ICPB_OK;
}

static int reset_state(Ctx* c, const icpb_params* p)
{
	IterState* h = c->st_host;
	memset(h, 0, sizeof *h);
	h->max_iter = p->max_iter; h->stop_early = p->stop_early; h->tol = p->tol; h->flags = p->flags;
	h->count_slot = (p->metric == ICPB_POINT_TO_PLANE) ? 27 : 15;
	c->run_flags = p->flags;
	h->n_total = (double)c->n;
	for (int k = 0; k < 9; k++) h->Rtot[k] = (k % 4 == 0) ? 1.0 : 0.0;
	for (int k = 0; k < 9; k++) h->R[k] = (k % 4 == 0) ? 1.0f : 0.0f;
	// the global point count changes only when a rank uploads a shard of another size: one allreduce per such upload
	// (every rank calls icpb_set_source the same number of times), not one per run
	const bool count_known = c->world > 1 && c->n_total_valid;
	if (count_known) h->n_total = c->n_total;
	ICPB_CUDA(c, cudaMemcpyAsync(c->st, h, sizeof *h, cudaMemcpyHostToDevice, c->stream));
	if (c->world > 1 && !count_known) {
		int rc = dist_allreduce_f64(c->dist, &c->st->n_total, 1, c->stream, c->err, sizeof c->err);
		if (rc != ICPB_OK) return rc;
		double nt = 0.0;
		ICPB_CUDA(c, cudaMemcpyAsync(&nt, &c->st->n_total, sizeof nt, cudaMemcpyDeviceToHost, c->stream));
		ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
		c->n_total = nt; c->n_total_valid = true;
	}
	int rc = ensure_errors(c, p->max_iter + 2);
	if (rc != ICPB_OK) return rc;
	ICPB_CUDA(c, cudaMemsetAsync(c->errors, 0, sizeof(float) * (size_t)(p->max_iter + 2), c->stream));
	return ICPB_OK;
}

static int check_params(Ctx* c, const icpb_params* p)
{
	if (!p) return fail(c, ICPB_ERR_BADARG, "params is NULL");
	if (p->max_iter < 1 || p->max_iter > (1 << 20)) return fail(c, ICPB_ERR_BADARG, "max_iter out of range");
	if (p->metric != ICPB_POINT_TO_POINT && p->metric != ICPB_POINT_TO_PLANE) return fail(c, ICPB_ERR_BADARG, "unknown metric");
	if (p->dist_mode < ICPB_DIST_SQ || p->dist_mode > ICPB_DIST_STD) return fail(c, ICPB_ERR_BADARG, "unknown dist_mode");
	if (p->nn_method < ICPB_NN_BRUTE || p->nn_method > ICPB_NN_BRUTE_DIRECT) return fail(c, ICPB_ERR_BADARG, "unknown nn_method");
	if (c->m <= 0) return fail(c, ICPB_ERR_STATE, "no target cloud: call icpb_set_target first");
	if (c->n <= 0 && c->world == 1) return fail(c, ICPB_ERR_STATE, "no source cloud: call icpb_set_source first");
	if (p->metric == ICPB_POINT_TO_PLANE && !c->have_normals) return fail(c, ICPB_ERR_STATE, "point-to-plane needs normals: call icpb_estimate_normals or icpb_set_normals");
	return ICPB_OK;
}

int launch_match(Ctx* c, int dist_mode, int nn_method, float sentinel);   // nn dispatch (grid variant in grid_nn.cu)

// one loop body, enqueued on the stream
// `ev`: nullptr, or 2 events (around the matching step), or — `phases` — 4 events bracketing matching, minimisation and
// transformation + error, the phases the reference's instrumented programs time with dsecnd()
// (src/CUDA/ICP_point_to_point_clean.cu:320-457).
static int enqueue_iteration(Ctx* c, const icpb_params* p, cudaEvent_t* ev, bool phases)
{
	int rc;
	if (ev) ICPB_CUDA(c, cudaEventRecord(ev[0], c->stream));
	nvtxRangePushA("icpb:match");
	rc = launch_match(c, p->dist_mode, p->nn_method, p->sentinel);
	nvtxRangePop();
	if (rc != ICPB_OK) return rc;
	if (ev) ICPB_CUDA(c, cudaEventRecord(ev[1], c->stream));
	ICPB_NVTX("icpb:minimize+transform");
	if ((rc = launch_moments(c, p->metric)) != ICPB_OK) return rc;
	const bool nccl_exchange = c->world > 1 && c->peer.world < 2;   // otherwise the exchange happens inside K2/K7/K4
	if (nccl_exchange) {
		const int cnt = (p->metric == ICPB_POINT_TO_PLANE) ? 28 : 16;
		if ((rc = dist_allreduce_f64(c->dist, c->st->moments, cnt, c->stream, c->err, sizeof c->err)) != ICPB_OK) return rc;
		if ((rc = launch_solve(c, p->metric)) != ICPB_OK) return rc;
	}
	if (ev && phases) ICPB_CUDA(c, cudaEventRecord(ev[2], c->stream));
	if ((rc = launch_transform(c)) != ICPB_OK) return rc;
	if (nccl_exchange) {
		if ((rc = dist_allreduce_f64(c->dist, &c->st->err_sum, 1, c->stream, c->err, sizeof c->err)) != ICPB_OK) return rc;
		if ((rc = launch_finish(c)) != ICPB_OK) return rc;
	}
	if (ev && phases) ICPB_CUDA(c, cudaEventRecord(ev[3], c->stream));
	return ICPB_OK;
}

static int read_state(Ctx* c)
{
	ICPB_CUDA(c, cudaMemcpyAsync(c->st_host, c->st, sizeof(IterState), cudaMemcpyDeviceToHost, c->stream));
	if (c->world > 1 && dist_has_comm(c->dist) && c->peer.world < 2) {
		// NCCL path: a collective whose peer died never completes. Poll the stream and the communicator's asynchronous
		// error state (ncclCommGetAsyncError) instead of blocking in cudaStreamSynchronize; on error the communicator is
		// aborted, which lets the pending work drain.
		for (;;) {
			const cudaError_t q = cudaStreamQuery(c->stream);
			if (q == cudaSuccess) break;
			if (q != cudaErrorNotReady) return fail_cuda(c, q, "cudaStreamQuery", __FILE__, __LINE__);
			const int rc = dist_check_async(c->dist, c->err, sizeof c->err);
			if (rc != ICPB_OK) { cudaStreamSynchronize(c->stream); return rc; }
			std::this_thread::sleep_for(std::chrono::microseconds(20));
		}
	}
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return kf_policy_update(c);      // the stream is idle: a good moment to look at the filter's exact-pass rate
}

int create_context(Ctx** out, int device)
{
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ICPB_ERR_NODEVICE;
	if (device < 0 || device >= count) return ICPB_ERR_BADARG;
	Ctx* c = new (std::nothrow) Ctx();
	if (!c) return ICPB_ERR_NOMEM;
	c->device = device;
	cudaDeviceProp prop;
	if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return ICPB_ERR_CUDA; }
	if (prop.major != 10) { delete c; return ICPB_ERR_NODEVICE; }   // sm_100a only: no fallback path exists
	c->sm_count = prop.multiProcessorCount;
	cudaDeviceGetAttribute(&c->sm_clock_khz, cudaDevAttrClockRate, device);
	snprintf(c->name, sizeof c->name, "%s", prop.name);
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ICPB_ERR_CUDA; }
	for (int k = 0; k < 4; k++) cudaEventCreate(&c->ev[k]);
	c->reduce_grid = c->sm_count * 8;
	if (cudaMalloc((void**)&c->st, sizeof(IterState)) != cudaSuccess ||
	    cudaMallocHost((void**)&c->st_host, sizeof(IterState)) != cudaSuccess ||
	    cudaMalloc((void**)&c->partials, sizeof(double) * 32 * (size_t)c->reduce_grid) != cudaSuccess) {
		icpb_destroy(reinterpret_cast<icpb_ctx*>(c));
		return ICPB_ERR_NOMEM;
	}
	cudaMemset(c->st, 0, sizeof(IterState));
	if (const char* e = getenv("ICPB_K1_CFG")) { c->k1_cfg = atoi(e); c->k1_cfg_forced = true; }
	if (const char* e = getenv("ICPB_K1_GRID")) c->k1_grid_override = atoi(e);
	if (const char* e = getenv("ICPB_K1_FILTER")) c->k1_use_filter = atoi(e) != 0;
	if (const char* e = getenv("ICPB_K1_SEED")) c->kf_use_seed = atoi(e) != 0;
	if (const char* e = getenv("ICPB_K1_TC")) c->k1_use_tc = atoi(e) != 0;
	if (!c->k1_use_tc && !getenv("ICPB_K1_FILTER_MIN_PAIRS")) c->kf_min_pairs = 1e9;      // the FP32 filter's own crossover
	if (const char* e = getenv("ICPB_KT_VAR")) c->kt_variant = atoi(e);
	if (const char* e = getenv("ICPB_KT_SORT")) c->kt_sort_mode = atoi(e) != 0 ? 1 : 0;
	if (const char* e = getenv("ICPB_KT_TPC")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) { c->kt_tpc_start = v; c->kt_tpc_auto = v; c->kt_tpc_forced_start = true; } }
	if (const char* e = getenv("ICPB_K1_FILTER_MIN_PAIRS")) c->kf_min_pairs = atof(e);
	if (const char* e = getenv("ICPB_GRAPHS")) c->graphs_enabled = atoi(e) != 0;
	if (const char* e = getenv("ICPB_KF_CHUNK")) c->kf_chunk_override = atoi(e);
	if (const char* e = getenv("ICPB_GRID_PYRAMID")) c->grid_pyramid = atoi(e) != 0;
	if (const char* e = getenv("ICPB_KNN_PYRAMID")) c->knn_pyramid = atoi(e) != 0;
	if (const char* e = getenv("ICPB_KF_GSS")) { int a = 0, b = 0, d = 0; if (sscanf(e, "%d,%d,%d", &a, &b, &d) == 3 && a > 0 && b >= a && d > 0) { c->kf_gss[0] = a; c->kf_gss[1] = b; c->kf_gss[2] = d; } }
	if (const char* e = getenv("ICPB_KF_DIMS")) c->kf_dims_forced = atoi(e);
	if (const char* e = getenv("ICPB_KF_S")) c->kf_s = (atoi(e) == 16) ? 16 : 8;
	if (const char* e = getenv("ICPB_KF_DROP")) c->kf_drop_forced = atoi(e);
	*out = c;
	return ICPB_OK;
}

} // namespace icpb

using namespace icpb;
static inline Ctx* C(icpb_ctx* p) { return reinterpret_cast<Ctx*>(p); }
static inline const Ctx* C(const icpb_ctx* p) { return reinterpret_cast<const Ctx*>(p); }
#define ICPB_ENTER(ctx) do { if (!(ctx)) return ICPB_ERR_BADARG; cudaError_t e__ = cudaSetDevice(C(ctx)->device); if (e__ != cudaSuccess) return fail_cuda(C(ctx), e__, "cudaSetDevice", __FILE__, __LINE__); } while (0)

extern "C" {

int icpb_version(void) { return ICPB_VERSION; }

const char* icpb_status_string(int s)
{
	switch (s) {
	case ICPB_OK: return "ok";
	case ICPB_ERR_CUDA: return "CUDA error";
	case ICPB_ERR_BADARG: return "bad argument";
	case ICPB_ERR_NCCL: return "NCCL error";
	case ICPB_ERR_STATE: return "call out of order";
	case ICPB_ERR_NOMEM: return "out of memory";
	case ICPB_ERR_NUMERIC: return "normal equations not positive definite";
	case ICPB_ERR_NODEVICE: return "no sm_100 CUDA device";
	default: return "unknown status";
	}
}

void icpb_default_params(icpb_params* p)
{
	if (!p) return;
	p->metric = ICPB_POINT_TO_POINT; p->dist_mode = ICPB_DIST_SQ; p->nn_method = ICPB_NN_BRUTE;
	p->max_iter = 40; p->stop_early = 1; p->sync_every = 0; p->sentinel = 100000.0f; p->tol = 0.000001; p->flags = 0;
}

int icpb_device_count(int* count)
{
	if (!count) return ICPB_ERR_BADARG;
	if (cudaGetDeviceCount(count) != cudaSuccess) { *count = 0; return ICPB_ERR_NODEVICE; }
	return ICPB_OK;
}

int icpb_create(icpb_ctx** out, int device)
{
	if (!out) return ICPB_ERR_BADARG;
	Ctx* c = nullptr;
	int rc = create_context(&c, device);
	if (rc != ICPB_OK) return rc;
	*out = reinterpret_cast<icpb_ctx*>(c);
	return ICPB_OK;
}

int icpb_nccl_unique_id(void* id128) { return id128 ? dist_unique_id(id128) : ICPB_ERR_BADARG; }

int icpb_create_dist(icpb_ctx** out, int device, int rank, int world, const void* nccl_unique_id)
{
	if (!out || world < 1 || rank < 0 || rank >= world) return ICPB_ERR_BADARG;
	Ctx* c = nullptr;
	int rc = create_context(&c, device);
	if (rc != ICPB_OK) return rc;
	c->rank = rank; c->world = world;
	if (world > 1) {
		if (!nccl_unique_id) { icpb_destroy(reinterpret_cast<icpb_ctx*>(c)); return ICPB_ERR_BADARG; }
		rc = dist_init(&c->dist, rank, world, nccl_unique_id, c->err, sizeof c->err);
		if (rc == ICPB_OK) rc = dist_peer_init(c->dist, device, c->stream, &c->peer, c->err, sizeof c->err);
		if (rc != ICPB_OK) { icpb_destroy(reinterpret_cast<icpb_ctx*>(c)); return rc; }
	}
	*out = reinterpret_cast<icpb_ctx*>(c);
	return ICPB_OK;
}

int icpb_destroy(icpb_ctx* ctx)
{
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = C(ctx);
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
	dist_destroy(c->dist);
	cudaFree(c->q4); cudaFree(c->qtiles); cudaFree(c->nrm4); cudaFree(c->nbr);
	cudaFree(c->px); cudaFree(c->py); cudaFree(c->pz); cudaFree(c->keys); cudaFree(c->idx); cudaFree(c->seed); cudaFree(c->dmin);
	cudaFree(c->stage_xyz); cudaFree(c->st); cudaFree(c->partials); cudaFree(c->errors);
	cudaFree(c->kt_tiles); cudaFree(c->kt_fail); cudaFree(c->kt_work); cudaFree(c->tgt_differ); cudaFree(c->kt_skeys); cudaFree(c->kt_sperm2); cudaFree(c->kt_mkeys); cudaFree(c->kt_perm2); cudaFree(c->kt_q4s); cudaFree(c->kt_slot_index); cudaFree(c->kt_colstart); cudaFree(c->kt_scan_a); cudaFree(c->kt_scan_b); cudaFree(c->kt_cub_tmp); cudaFree(c->kt_hmax_d);
	cudaFree(c->grid_counts); cudaFree(c->grid_cell_of); cudaFree(c->grid_sums); cudaFree(c->grid_mm);
	cudaFree(c->grid_cell_start); cudaFree(c->grid_sorted4); cudaFree(c->grid_open_list); cudaFree(c->grid_counters); cudaFree(c->grid_occ); cudaFree(c->kf_tiles7); cudaFree(c->kf_scratch); cudaFree(c->kf_stats); cudaFree(c->kf_work_counter);
	cudaFree(c->k9.s); cudaFree(c->k9.t); cudaFree(c->k9.e); cudaFree(c->k9.ints); cudaFree(c->k9.dbl);
	if (c->k9.one) cudaFreeHost(c->k9.one);
	if (c->k9.ev) cudaEventDestroy(c->k9.ev);
	if (c->k9.copy_stream) cudaStreamDestroy(c->k9.copy_stream);
	if (c->st_host) cudaFreeHost(c->st_host);
	if (c->errors_host) cudaFreeHost(c->errors_host);
	for (int k = 0; k < 4; k++) if (c->ev[k]) cudaEventDestroy(c->ev[k]);
	for (int k = 0; k < c->ev_match_cap; k++) cudaEventDestroy(c->ev_match[k]);
	free(c->ev_match);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
	return ICPB_OK;
}

const char* icpb_last_error(const icpb_ctx* ctx) { return ctx ? C(ctx)->err : "null context"; }

int icpb_device_info(const icpb_ctx* ctx, int* sm_count, int* sm_clock_khz, char* name64)
{
	if (!ctx) return ICPB_ERR_BADARG;
	if (sm_count) *sm_count = C(ctx)->sm_count;
	if (sm_clock_khz) *sm_clock_khz = C(ctx)->sm_clock_khz;
	if (name64) snprintf(name64, 64, "%s", C(ctx)->name);
	return ICPB_OK;
}

long long icpb_launch_count(const icpb_ctx* ctx) { return ctx ? C(ctx)->launches : 0; }

int icpb_dist_info(const icpb_ctx* ctx, int* rank, int* world, int* peer_exchange)
{
	if (!ctx) return ICPB_ERR_BADARG;
	if (rank) *rank = C(ctx)->rank;
	if (world) *world = C(ctx)->world;
	if (peer_exchange) *peer_exchange = C(ctx)->peer.world > 1 ? 1 : 0;
	return ICPB_OK;
}

// ---- clouds ---------------------------------------------------------------------------------------
int icpb_set_target(icpb_ctx* ctx, const float* xyz, int m, int on_device)
{
	ICPB_NVTX("icpb_set_target");
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (!xyz || m <= 0) return fail(c, ICPB_ERR_BADARG, "icpb_set_target: empty target");
	int rc;
	const int nt = (m + K1_TT - 1) / K1_TT;
	if (m != c->m) {
		if ((rc = dev_alloc(c, &c->q4, (size_t)m)) != ICPB_OK) return rc;
		if ((rc = dev_alloc(c, &c->qtiles, (size_t)nt * 3 * K1_TT)) != ICPB_OK) return rc;
		if ((rc = dev_alloc(c, &c->nrm4, (size_t)0)) != ICPB_OK) return rc;
		if ((rc = dev_alloc(c, &c->nbr, (size_t)0)) != ICPB_OK) return rc;
	}
	const bool same_size = (m == c->m) && c->tgt_packed;
	const float* src = xyz;
	if (!on_device) {
		if ((rc = ensure_stage(c, sizeof(float) * 3 * (size_t)m)) != ICPB_OK) return rc;
		ICPB_CUDA(c, cudaMemcpyAsync(c->stage_xyz, xyz, sizeof(float) * 3 * (size_t)m, cudaMemcpyHostToDevice, c->stream));
		src = c->stage_xyz;
	}
	if (same_size) {
		// The same cloud again, bit for bit (a host-driven loop re-uploads its target at every step)? Then everything derived
		// from it stays: packed copies, filter tiles, grid, normals, policy state, captured graphs.
		if (!c->tgt_differ) ICPB_CUDA(c, cudaMalloc((void**)&c->tgt_differ, sizeof(int)));
		int differ = 1;
		ICPB_CUDA(c, cudaMemsetAsync(c->tgt_differ, 0, sizeof(int), c->stream));
		if ((rc = launch_target_same(c, src, m, c->tgt_differ)) != ICPB_OK) return rc;
		ICPB_CUDA(c, cudaMemcpyAsync(&differ, c->tgt_differ, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
		if (!differ) return ICPB_OK;
	}
	if (m != c->m) c->seed_n = -1;      // seeds are indices into the target: only a different size invalidates them
	c->m = m; c->nt = nt; c->have_normals = false; c->knn_k = 0; c->grid_ready = false; c->kf_ready = false; c->kt_ready = false; c->step_state_ready = false; c->graph_gen++;
	c->tgt_packed = false;
	if ((rc = launch_pack_target(c, src, m)) != ICPB_OK) return rc;
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	c->tgt_packed = true;
	return ICPB_OK;
}

int icpb_set_source(icpb_ctx* ctx, const float* xyz, int n, int on_device)
{
	ICPB_NVTX("icpb_set_source");
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (n < 0 || (n > 0 && !xyz)) return fail(c, ICPB_ERR_BADARG, "icpb_set_source: bad arguments");
	if (n == 0 && c->world == 1) return fail(c, ICPB_ERR_BADARG, "icpb_set_source: empty source");
	int rc;
	const int blk = k1_max_block_sources();
	const int cap = ((n + blk - 1) / blk) * blk + blk;
	if (cap > c->n_cap) {
		if ((rc = dev_alloc(c, &c->px, (size_t)cap)) != ICPB_OK) return rc;
		if ((rc = dev_alloc(c, &c->py, (size_t)cap)) != ICPB_OK) return rc;
		if ((rc = dev_alloc(c, &c->pz, (size_t)cap)) != ICPB_OK) return rc;
		if ((rc = dev_alloc(c, &c->keys, (size_t)cap)) != ICPB_OK) return rc;
		if ((rc = dev_alloc(c, &c->idx, (size_t)cap)) != ICPB_OK) return rc;
		if ((rc = dev_alloc(c, &c->seed, (size_t)cap)) != ICPB_OK) return rc;
		c->seed_n = -1;
		if ((rc = dev_alloc(c, &c->dmin, (size_t)cap)) != ICPB_OK) return rc;
		c->n_cap = cap;
	}
	// a captured iteration graph bakes in the buffers and the point count: it stays valid across uploads that change
	// neither (the repeated same-size registrations ICPB_FLAG_GRAPH is meant for)
	if (cap > c->n_cap_at_graph || n != c->n) c->graph_gen++;
	c->n_cap_at_graph = c->n_cap;
	c->n = n; c->step_state_ready = false;   // the control block caches the global point count
	c->kt_src_checked = false;                                // K1T looks at the new source's order before its next pass
	c->n_total_valid = false;
	const float* src = xyz;
	if (!on_device && n > 0) {
		if ((rc = ensure_stage(c, sizeof(float) * 3 * (size_t)n)) != ICPB_OK) return rc;
		ICPB_CUDA(c, cudaMemcpyAsync(c->stage_xyz, xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
		src = c->stage_xyz;
	}
	// the warm-start seeds (last resolved correspondences) survive a new source of the same size against the same
	// target: that is what a host-driven ICP loop uploads; any value in [0, m) is a valid seed, so this affects speed only
	const bool reset_seed = (c->seed_n != n);
	c->seed_n = n;
	if (reset_seed) c->kf_seeded = false;
	if ((rc = launch_pack_source(c, src, n, reset_seed)) != ICPB_OK) return rc;
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return ICPB_OK;
}

int icpb_get_source(icpb_ctx* ctx, float* xyz, int on_device)
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (!xyz) return fail(c, ICPB_ERR_BADARG, "icpb_get_source: NULL output");
	if (c->n <= 0) return ICPB_OK;
	int rc;
	float* dst = xyz;
	if (!on_device) { if ((rc = ensure_stage(c, sizeof(float) * 3 * (size_t)c->n)) != ICPB_OK) return rc; dst = c->stage_xyz; }
	if ((rc = launch_unpack_source(c, dst)) != ICPB_OK) return rc;
	if (!on_device) ICPB_CUDA(c, cudaMemcpyAsync(xyz, dst, sizeof(float) * 3 * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return ICPB_OK;
}

int icpb_get_correspondences(icpb_ctx* ctx, int* idx, int on_device)
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (!idx) return fail(c, ICPB_ERR_BADARG, "icpb_get_correspondences: NULL output");
	if (c->n <= 0) return ICPB_OK;
	ICPB_CUDA(c, cudaMemcpyAsync(idx, c->idx, sizeof(int) * (size_t)c->n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return ICPB_OK;
}

int icpb_get_min_distances(icpb_ctx* ctx, float* d, int on_device)
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (!d) return fail(c, ICPB_ERR_BADARG, "icpb_get_min_distances: NULL output");
	if (c->n <= 0) return ICPB_OK;
	ICPB_CUDA(c, cudaMemcpyAsync(d, c->dmin, sizeof(float) * (size_t)c->n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return ICPB_OK;
}

// ---- single steps ------------------------------------------------------------------------------------
static int ensure_step_state(Ctx* c)
{
	// the step-wise API runs on a fresh control block (done = 0) unless a run left one behind
	if (!c->step_state_ready) {
		icpb_params p; icpb_default_params(&p); p.max_iter = 1 << 16; p.stop_early = 0;
		int rc = reset_state(c, &p);
		if (rc != ICPB_OK) return rc;
		c->step_state_ready = true;
	}
	return ICPB_OK;
}

int icpb_match(icpb_ctx* ctx, int dist_mode, int nn_method, float sentinel)
{
	ICPB_NVTX("icpb_match");
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (c->m <= 0 || c->n <= 0) return fail(c, ICPB_ERR_STATE, "icpb_match: set the target and the source first");
	if (dist_mode < ICPB_DIST_SQ || dist_mode > ICPB_DIST_STD) return fail(c, ICPB_ERR_BADARG, "unknown dist_mode");
	int rc;
	if ((rc = ensure_step_state(c)) != ICPB_OK) return rc;
	if ((rc = launch_key_reset(c)) != ICPB_OK) return rc;
	if ((rc = launch_match(c, dist_mode, nn_method, sentinel)) != ICPB_OK) return rc;
	if ((rc = launch_resolve(c, sentinel)) != ICPB_OK) return rc;
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return filter_tc_check(c);
}

int icpb_minimize(icpb_ctx* ctx, int metric, float R[9], float T[3])
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (c->m <= 0 || c->n <= 0) return fail(c, ICPB_ERR_STATE, "icpb_minimize: set the clouds and match first");
	if (metric == ICPB_POINT_TO_PLANE && !c->have_normals) return fail(c, ICPB_ERR_STATE, "point-to-plane needs normals");
	int rc;
	if ((rc = ensure_step_state(c)) != ICPB_OK) return rc;
	if ((rc = launch_moments(c, metric)) != ICPB_OK) return rc;
	if (c->world > 1 && c->peer.world < 2) {
		if ((rc = dist_allreduce_f64(c->dist, c->st->moments, metric == ICPB_POINT_TO_PLANE ? 28 : 16, c->stream, c->err, sizeof c->err)) != ICPB_OK) return rc;
		if ((rc = launch_solve(c, metric)) != ICPB_OK) return rc;
	}
	if ((rc = read_state(c)) != ICPB_OK) return rc;
	if (c->st_host->numeric_error == 100) return fail(c, ICPB_ERR_NCCL, "peer-memory exchange timed out: a rank never published its moment sums");
	if (c->st_host->numeric_error) return fail(c, ICPB_ERR_NUMERIC, "6x6 normal equations are not positive definite");
	if (R) memcpy(R, c->st_host->R, sizeof(float) * 9);
	if (T) memcpy(T, c->st_host->T, sizeof(float) * 3);
	return ICPB_OK;
}

int icpb_transform(icpb_ctx* ctx, float* rms)
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (c->m <= 0 || c->n <= 0) return fail(c, ICPB_ERR_STATE, "icpb_transform: nothing to transform");
	int rc;
	if ((rc = ensure_step_state(c)) != ICPB_OK) return rc;
	if ((rc = launch_transform(c)) != ICPB_OK) return rc;
	if (c->world > 1 && c->peer.world < 2) {
		if ((rc = dist_allreduce_f64(c->dist, &c->st->err_sum, 1, c->stream, c->err, sizeof c->err)) != ICPB_OK) return rc;
		if ((rc = launch_finish(c)) != ICPB_OK) return rc;
	}
	if ((rc = read_state(c)) != ICPB_OK) return rc;
	if (rms) *rms = c->st_host->last_err;
	return ICPB_OK;
}

int icpb_set_transform(icpb_ctx* ctx, const float R[9], const float T[3])
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (!R || !T) return fail(c, ICPB_ERR_BADARG, "icpb_set_transform: NULL argument");
	int rc;
	if ((rc = ensure_step_state(c)) != ICPB_OK) return rc;
	ICPB_CUDA(c, cudaMemcpyAsync(c->st->R, R, sizeof(float) * 9, cudaMemcpyHostToDevice, c->stream));
	ICPB_CUDA(c, cudaMemcpyAsync(c->st->T, T, sizeof(float) * 3, cudaMemcpyHostToDevice, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	return ICPB_OK;
}

int icpb_get_moments(icpb_ctx* ctx, double* mom, int count)
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (!mom || count < 1 || count > 32) return fail(c, ICPB_ERR_BADARG, "icpb_get_moments: bad arguments");
	int rc;
	if ((rc = read_state(c)) != ICPB_OK) return rc;
	memcpy(mom, c->st_host->moments, sizeof(double) * (size_t)count);
	return ICPB_OK;
}

// ---- whole loop ------------------------------------------------------------------------------------------
int icpb_run(icpb_ctx* ctx, const icpb_params* params, float* errors, icpb_result* result)
{
	ICPB_NVTX("icpb_run");
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	int rc;
	if ((rc = check_params(c, params)) != ICPB_OK) return rc;
	icpb_params p = *params;
	// sync_every <= 0: adaptive — small clouds are launch-latency bound, so several iterations are enqueued per read
	// of the device-side stop flag (kernels launched after the flag is raised return immediately: same results)
	if (p.sync_every < 1) p.sync_every = ((double)c->n * (double)c->m < 4e9) ? 4 : 1;
	c->step_state_ready = false;
	if ((rc = reset_state(c, &p)) != ICPB_OK) return rc;
	if ((rc = launch_key_reset(c)) != ICPB_OK) return rc;

	// one (start, stop) event pair per iteration around the matching kernel; with ICPB_FLAG_PROFILE two more events
	// split the rest of the iteration into minimisation and transformation + error
	const bool phases = (p.flags & ICPB_FLAG_PROFILE) != 0;
	const int epi = phases ? 4 : 2;
	const int want = epi * p.max_iter;
	if (want > c->ev_match_cap && want <= 4096) {
		cudaEvent_t* ne = (cudaEvent_t*)realloc(c->ev_match, sizeof(cudaEvent_t) * (size_t)want);
		if (ne) {
			c->ev_match = ne;
			for (int k = c->ev_match_cap; k < want; k++) cudaEventCreate(&c->ev_match[k]);
			c->ev_match_cap = want;
		}
	}
	c->pairs_acc = 0;
	// per-target data of the matching method is built before the clock starts (as the reference's cudaMalloc block is)
	if (p.nn_method == ICPB_NN_BRUTE && c->k1_use_filter && p.dist_mode != ICPB_DIST_STD && (double)c->n * (double)c->m >= c->kf_min_pairs) {
		if ((rc = prepare_match_filter(c)) != ICPB_OK) return rc;
		if (c->k1_use_tc && (rc = ensure_filter_tc_data(c)) != ICPB_OK) return rc;
	}
	if (p.nn_method == ICPB_NN_GRID && p.dist_mode != ICPB_DIST_STD) { if ((rc = prepare_match_grid(c)) != ICPB_OK) return rc; }
	// Opt-in (ICPB_FLAG_GRAPH / ICPB_GRAPHS=1) for launch-latency-bound small problems: after a first plain iteration
	// (which also builds lazily created data and caches launch attributes) the remaining ones are replayed from a CUDA
	// graph holding `sync_every` iterations; the instantiated graph is reused by later runs on clouds of the same size.
	// Iterations past the stop flag / MAX_ITER return immediately, so results are those of the plain loop.
	// Measured on B200: capture + instantiate cost 2-90 ms, so a single 27-iteration registration of 16 384 points
	// (3.3 ms) is slower with it; it is off by default.
	const bool use_graph = (c->graphs_enabled || (p.flags & ICPB_FLAG_GRAPH)) && c->world == 1 && p.nn_method != ICPB_NN_GRID && !(p.flags & ICPB_FLAG_PROFILE) &&
	                       (double)c->n * (double)c->m < 4e9 && p.max_iter > 1;
	ICPB_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
	int enq = 0;
	bool timed_iterations = true;
	if (use_graph) {
		timed_iterations = false;
		if ((rc = enqueue_iteration(c, &p, nullptr, false)) != ICPB_OK) return rc;
		enq = 1;
		if ((rc = read_state(c)) != ICPB_OK) return rc;
		if (!c->st_host->done) {
			// the policy may just have changed K1T's group size: lay the tiles out again before the capture, not inside it
			if (p.nn_method == ICPB_NN_BRUTE && c->k1_use_filter && c->k1_use_tc && c->kf_dims_last == 4 && (rc = ensure_filter_tc_data(c)) != ICPB_OK) return rc;
			const bool valid = c->graph_exec != nullptr && c->graph_built_gen == c->graph_gen && c->graph_batch == p.sync_every &&
			                   c->graph_key[0] == p.metric && c->graph_key[1] == p.dist_mode && c->graph_key[2] == p.nn_method &&
			                   c->graph_key[3] == p.flags && c->graph_sentinel == p.sentinel;
			if (!valid) {
				if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
				const long long l0 = c->launches;
				cudaGraph_t g = nullptr;
				ICPB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
				for (int b = 0; b < p.sync_every && rc == ICPB_OK; b++) rc = enqueue_iteration(c, &p, nullptr, false);
				cudaError_t ce = cudaStreamEndCapture(c->stream, &g);
				if (rc != ICPB_OK) { if (g) cudaGraphDestroy(g); return rc; }
				if (ce != cudaSuccess) return fail_cuda(c, ce, "cudaStreamEndCapture", __FILE__, __LINE__);
				ce = cudaGraphInstantiate(&c->graph_exec, g, 0);
				cudaGraphDestroy(g);
				if (ce != cudaSuccess) return fail_cuda(c, ce, "cudaGraphInstantiate", __FILE__, __LINE__);
				c->graph_launches = (int)(c->launches - l0);
				c->launches = l0;                      // captured, not launched
				c->graph_built_gen = c->graph_gen; c->graph_batch = p.sync_every;
				c->graph_key[0] = p.metric; c->graph_key[1] = p.dist_mode; c->graph_key[2] = p.nn_method; c->graph_key[3] = p.flags;
				c->graph_sentinel = p.sentinel;
			}
			while (true) {
				ICPB_CUDA(c, cudaGraphLaunch(c->graph_exec, c->stream));
				c->launches += c->graph_launches;
				enq += p.sync_every;
				if ((rc = read_state(c)) != ICPB_OK) return rc;
				if (c->st_host->done || enq >= p.max_iter) break;
			}
		}
	} else {
		while (true) {
			for (int b = 0; b < p.sync_every && enq < p.max_iter; b++, enq++) {
				cudaEvent_t* ev = (epi * enq + epi - 1 < c->ev_match_cap) ? c->ev_match + epi * enq : nullptr;
				if ((rc = enqueue_iteration(c, &p, ev, phases)) != ICPB_OK) return rc;
			}
			if ((rc = read_state(c)) != ICPB_OK) return rc;
			if (c->st_host->done || enq >= p.max_iter) break;
		}
	}
	ICPB_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
	ICPB_CUDA(c, cudaMemcpyAsync(c->errors_host, c->errors, sizeof(float) * (size_t)(p.max_iter + 1), cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	const IterState* h = c->st_host;
	if (h->numeric_error == 100) return fail(c, ICPB_ERR_NCCL, "peer-memory exchange timed out: a rank never published its moment sums");
	if (h->numeric_error) return fail(c, ICPB_ERR_NUMERIC, "6x6 normal equations are not positive definite");
	if (errors) memcpy(errors, c->errors_host, sizeof(float) * (size_t)(p.max_iter + 1));
	if (result) {
		memset(result, 0, sizeof *result);
		result->iterations = h->iteration; result->iterations_run = h->iters_run;
		memcpy(result->R, h->Rtot, sizeof h->Rtot); memcpy(result->t, h->ttot, sizeof h->ttot);
		memcpy(result->last_R, h->R, sizeof h->R); memcpy(result->last_T, h->T, sizeof h->T);
		cudaEventElapsedTime(&result->elapsed_ms, c->ev[0], c->ev[1]);
		float mm = 0.f, mn = 0.f, mt = 0.f;
		for (int k = 0; timed_iterations && k < h->iters_run && epi * k + epi - 1 < c->ev_match_cap; k++) {
			float ms = 0.f;
			cudaEvent_t* ev = c->ev_match + epi * k;
			if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) mm += ms;
			if (phases && cudaEventElapsedTime(&ms, ev[1], ev[2]) == cudaSuccess) mn += ms;
			if (phases && cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) mt += ms;
		}
		result->match_ms = mm; result->minimize_ms = mn; result->transform_ms = mt;
		result->nn_pairs = (double)h->iters_run * (double)c->n * (double)c->m;
	}
	return ICPB_OK;
}

int icpb_iterate_host(icpb_ctx* ctx, const icpb_params* params, const float* source_xyz, int n, const float* target_xyz, int m,
                      int* idx_out, float R[9], float T[3], float* rms)
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (!params) return fail(c, ICPB_ERR_BADARG, "params is NULL");
	int rc;
	if ((rc = icpb_set_target(ctx, target_xyz, m, 0)) != ICPB_OK) return rc;
	if ((rc = icpb_set_source(ctx, source_xyz, n, 0)) != ICPB_OK) return rc;
	icpb_params p = *params;
	p.max_iter = 1; p.stop_early = 0; p.sync_every = 1;
	if (p.metric == ICPB_POINT_TO_PLANE) {
		if ((rc = icpb_estimate_normals(ctx, 4, nullptr)) != ICPB_OK) return rc;
	}
	icpb_result res;
	float err[2] = { 0.f, 0.f };
	if ((rc = icpb_run(ctx, &p, err, &res)) != ICPB_OK) return rc;
	if (idx_out) { if ((rc = icpb_get_correspondences(ctx, idx_out, 0)) != ICPB_OK) return rc; }
	if (R) memcpy(R, res.last_R, sizeof res.last_R);
	if (T) memcpy(T, res.last_T, sizeof res.last_T);
	if (rms) *rms = err[1];
	return ICPB_OK;
}

int icpb_host_alloc(void** ptr, unsigned long long bytes)
{
	if (!ptr || bytes == 0) return ICPB_ERR_BADARG;
	return cudaMallocHost(ptr, (size_t)bytes) == cudaSuccess ? ICPB_OK : ICPB_ERR_NOMEM;
}
int icpb_host_free(void* ptr) { return (ptr && cudaFreeHost(ptr) == cudaSuccess) ? ICPB_OK : ICPB_ERR_BADARG; }

// ---- measurement helpers ---------------------------------------------------------------------------------
int icpb_measure_fp32_peak(icpb_ctx* ctx, double* tflops)
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (!tflops) return fail(c, ICPB_ERR_BADARG, "NULL output");
	const int blocks = c->sm_count * 8, iters = 8192;
	int rc;
	if ((rc = ensure_stage(c, sizeof(float) * 256 * (size_t)blocks)) != ICPB_OK) return rc;
	float best = 1e30f;
	for (int r = 0; r < 7; r++) {
		ICPB_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
		if ((rc = launch_fp32_peak(c, c->stage_xyz, iters, blocks)) != ICPB_OK) return rc;
		ICPB_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
		ICPB_CUDA(c, cudaEventSynchronize(c->ev[3]));
		float ms = 0.f;
		ICPB_CUDA(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
		if (r >= 2 && ms < best) best = ms;
	}
	*tflops = (double)blocks * 256.0 * 8.0 * iters * 2.0 / (best * 1e-3) * 1e-12;
	return ICPB_OK;
}

int icpb_time_match(icpb_ctx* ctx, int dist_mode, int nn_method, float sentinel, int reps, float* mean_ms, float* min_ms)
{
	ICPB_ENTER(ctx);
	Ctx* c = C(ctx);
	if (c->m <= 0 || c->n <= 0) return fail(c, ICPB_ERR_STATE, "icpb_time_match: set the clouds first");
	if (reps < 1) return fail(c, ICPB_ERR_BADARG, "reps < 1");
	int rc;
	if ((rc = ensure_step_state(c)) != ICPB_OK) return rc;
	double sum = 0.0; float mn = 1e30f;
	for (int r = 0; r < reps; r++) {
		if ((rc = launch_key_reset(c)) != ICPB_OK) return rc;
		ICPB_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
		if ((rc = launch_match(c, dist_mode, nn_method, sentinel)) != ICPB_OK) return rc;
		ICPB_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
		ICPB_CUDA(c, cudaEventSynchronize(c->ev[3]));
		float ms = 0.f;
		ICPB_CUDA(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
		sum += ms; if (ms < mn) mn = ms;
	}
	if ((rc = launch_resolve(c, sentinel)) != ICPB_OK) return rc;
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (mean_ms) *mean_ms = (float)(sum / reps);
	if (min_ms) *min_ms = mn;
	return ICPB_OK;
}

} // extern "C"
