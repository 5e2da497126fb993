// group.cpp — several GPUs of one box driven from ONE process in plain C/C++ (icpb_group_* of include/icp_b200.h).
//
// The reference is a single no-argument C++ main() (src/ICP_point_to_point.cu:90-460); the north star keeps the host
// code C/C++ and shards the SOURCE points over the GPUs of a box with the target replicated. A group is `ndev`
// contexts, rank r on devices[r], that behave exactly like the one-process-per-GPU ranks of icpb_create_dist — same
// kernels, same fused exchange of the moment sums inside K2/K7/K4 over NVLink peer memory (here through
// cudaDeviceEnablePeerAccess instead of CUDA IPC: one address space), same ncclAllReduce fallback (ncclCommInitAll) —
// but need no launcher, no torch, no MPI and no unique id. Every group call runs one host thread per device: with the
// fused exchange the kernels of all ranks wait for each other inside the iteration, so all of them must be in flight.
#include "common.cuh"
#include <thread>
#include <vector>
#include <algorithm>

using namespace icpb;

struct icpb_group {
	int world = 0;
	std::vector<int> devices;
	std::vector<Ctx*> ctx;
	std::vector<std::vector<int>> owned;      // owned[r] = original indices of rank r's source points, ascending
	std::vector<std::vector<float>> hbuf;     // per-rank host staging (gather / scatter)
	std::vector<std::vector<int>> ibuf;
	int n = 0, m = 0;
	char err[640] = {0};
};

namespace {

template <typename F> int for_each_rank(icpb_group* g, F fn)
{
	std::vector<int> rc((size_t)g->world, ICPB_OK);
	if (g->world == 1) { rc[0] = fn(0); }
	else {
		std::vector<std::thread> th;
		th.reserve((size_t)g->world);
		for (int r = 0; r < g->world; r++) th.emplace_back([&, r] { rc[(size_t)r] = fn(r); });
		for (auto& t : th) t.join();
	}
	for (int r = 0; r < g->world; r++)
		if (rc[(size_t)r] != ICPB_OK) {
			snprintf(g->err, sizeof g->err, "rank %d (device %d): %s", r, g->devices[(size_t)r], g->ctx[(size_t)r]->err);
			return rc[(size_t)r];
		}
	return ICPB_OK;
}
inline icpb_ctx* H(Ctx* c) { return reinterpret_cast<icpb_ctx*>(c); }

} // namespace

extern "C" {

int icpb_group_create(icpb_group** out, const int* devices, int ndev)
{
	if (!out || ndev < 1 || ndev > 64) return ICPB_ERR_BADARG;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ICPB_ERR_NODEVICE;
	icpb_group* g = new icpb_group();
	g->world = ndev;
	for (int r = 0; r < ndev; r++) g->devices.push_back(devices ? devices[r] : r);
	for (int r = 0; r < ndev; r++) {
		if (g->devices[(size_t)r] < 0 || g->devices[(size_t)r] >= count) { icpb_group_destroy(g); return ICPB_ERR_BADARG; }
		for (int q = 0; q < r; q++) if (g->devices[(size_t)q] == g->devices[(size_t)r]) { icpb_group_destroy(g); return ICPB_ERR_BADARG; }   // one rank per GPU: ranks wait for each other inside kernels
	}
	for (int r = 0; r < ndev; r++) {
		Ctx* c = nullptr;
		const int rc = create_context(&c, g->devices[(size_t)r]);
		if (rc != ICPB_OK) { icpb_group_destroy(g); return rc; }
		c->rank = r; c->world = ndev;
		g->ctx.push_back(c);
	}
	if (ndev > 1) {
		std::vector<Dist*> d((size_t)ndev, nullptr);
		std::vector<PeerXchg> px((size_t)ndev);
		const int rc = dist_init_local(d.data(), px.data(), g->devices.data(), ndev, g->err, sizeof g->err);
		if (rc != ICPB_OK) { icpb_group_destroy(g); return rc; }
		for (int r = 0; r < ndev; r++) { g->ctx[(size_t)r]->dist = d[(size_t)r]; g->ctx[(size_t)r]->peer = px[(size_t)r]; }
	}
	g->owned.resize((size_t)ndev); g->hbuf.resize((size_t)ndev); g->ibuf.resize((size_t)ndev);
	*out = g;
	return ICPB_OK;
}

int icpb_group_destroy(icpb_group* g)
{
	if (!g) return ICPB_ERR_BADARG;
	for (Ctx* c : g->ctx) if (c) icpb_destroy(H(c));
	delete g;
	return ICPB_OK;
}

int icpb_group_size(const icpb_group* g) { return g ? g->world : 0; }
icpb_ctx* icpb_group_ctx(icpb_group* g, int rank) { return (g && rank >= 0 && rank < g->world) ? H(g->ctx[(size_t)rank]) : nullptr; }
const char* icpb_group_last_error(const icpb_group* g) { return g ? g->err : "null group"; }

int icpb_group_info(const icpb_group* g, int* ndev, int* peer_exchange, int* devices)
{
	if (!g) return ICPB_ERR_BADARG;
	if (ndev) *ndev = g->world;
	if (peer_exchange) *peer_exchange = (g->world > 1 && g->ctx[0]->peer.world > 1) ? 1 : 0;
	if (devices) for (int r = 0; r < g->world; r++) devices[r] = g->devices[(size_t)r];
	return ICPB_OK;
}

int icpb_group_set_target(icpb_group* g, const float* xyz, int m)
{
	if (!g || !xyz || m <= 0) return ICPB_ERR_BADARG;
	g->m = m;
	return for_each_rank(g, [&](int r) { return icpb_set_target(H(g->ctx[(size_t)r]), xyz, m, 0); });
}

int icpb_group_set_source(icpb_group* g, const float* xyz, int n, int block)
{
	if (!g || !xyz || n <= 0 || block < 0) return ICPB_ERR_BADARG;
	const int W = g->world;
	for (int r = 0; r < W; r++) g->owned[(size_t)r].clear();
	if (block == 0) {                                  // contiguous shards, sizes differing by at most one
		const int base = n / W, rem = n % W;
		int lo = 0;
		for (int r = 0; r < W; r++) { const int cnt = base + (r < rem ? 1 : 0); for (int i = lo; i < lo + cnt; i++) g->owned[(size_t)r].push_back(i); lo += cnt; }
	} else {                                           // blocks of `block` consecutive points dealt round-robin: every rank sees the same mix of regions
		const int nblocks = (n + block - 1) / block;
		for (int b = 0; b < nblocks; b++) {
			std::vector<int>& o = g->owned[(size_t)(b % W)];
			for (int i = b * block; i < std::min(n, (b + 1) * block); i++) o.push_back(i);
		}
	}
	g->n = n;
	const int rc = for_each_rank(g, [&](int r) {
		const std::vector<int>& o = g->owned[(size_t)r];
		std::vector<float>& h = g->hbuf[(size_t)r];
		h.resize(3 * o.size() + 3);
		for (size_t k = 0; k < o.size(); k++) { const float* s = xyz + 3 * (size_t)o[k]; h[3 * k] = s[0]; h[3 * k + 1] = s[1]; h[3 * k + 2] = s[2]; }
		Ctx* c = g->ctx[(size_t)r];
		const int rc2 = icpb_set_source(H(c), h.data(), (int)o.size(), 0);
		if (rc2 == ICPB_OK && W > 1) { c->n_total = (double)n; c->n_total_valid = true; }      // the global count is known here: no collective needed
		return rc2;
	});
	return rc;
}

int icpb_group_estimate_normals(icpb_group* g, int k, int knn_dist_mode, float* elapsed_ms)
{
	if (!g) return ICPB_ERR_BADARG;
	std::vector<float> ms((size_t)g->world, 0.f);
	// the target is replicated, so are its normals: every device computes them (no exchange; 1 ms-scale work)
	const int rc = for_each_rank(g, [&](int r) { return icpb_estimate_normals_ex(H(g->ctx[(size_t)r]), k, knn_dist_mode, &ms[(size_t)r]); });
	if (elapsed_ms) *elapsed_ms = *std::max_element(ms.begin(), ms.end());
	return rc;
}

int icpb_group_run(icpb_group* g, const icpb_params* params, float* errors, icpb_result* result)
{
	if (!g || !params) return ICPB_ERR_BADARG;
	std::vector<icpb_result> res((size_t)g->world);
	std::vector<std::vector<float>> errs((size_t)g->world, std::vector<float>((size_t)params->max_iter + 2, 0.f));
	const int rc = for_each_rank(g, [&](int r) { return icpb_run(H(g->ctx[(size_t)r]), params, errs[(size_t)r].data(), &res[(size_t)r]); });
	if (rc != ICPB_OK) return rc;
	// every rank solved the same 3x3 / 6x6 system on identical sums: identical trajectories, bit for bit
	for (int r = 1; r < g->world; r++)
		if (memcmp(errs[(size_t)r].data(), errs[0].data(), sizeof(float) * ((size_t)params->max_iter + 1)) != 0 || memcmp(res[(size_t)r].R, res[0].R, sizeof res[0].R) != 0) {
			snprintf(g->err, sizeof g->err, "ranks 0 and %d disagree on the trajectory: the exchanged sums were not identical", r);
			return ICPB_ERR_NCCL;
		}
	if (errors) memcpy(errors, errs[0].data(), sizeof(float) * ((size_t)params->max_iter + 1));
	if (result) {
		*result = res[0];
		for (int r = 1; r < g->world; r++) {
			result->elapsed_ms = std::max(result->elapsed_ms, res[(size_t)r].elapsed_ms);
			result->match_ms = std::max(result->match_ms, res[(size_t)r].match_ms);
			result->minimize_ms = std::max(result->minimize_ms, res[(size_t)r].minimize_ms);
			result->transform_ms = std::max(result->transform_ms, res[(size_t)r].transform_ms);
			result->nn_pairs += res[(size_t)r].nn_pairs;
		}
	}
	return ICPB_OK;
}

int icpb_group_get_source(icpb_group* g, float* xyz)
{
	if (!g || !xyz) return ICPB_ERR_BADARG;
	return for_each_rank(g, [&](int r) {
		const std::vector<int>& o = g->owned[(size_t)r];
		std::vector<float>& h = g->hbuf[(size_t)r];
		h.resize(3 * o.size() + 3);
		const int rc = icpb_get_source(H(g->ctx[(size_t)r]), h.data(), 0);
		if (rc != ICPB_OK) return rc;
		for (size_t k = 0; k < o.size(); k++) { float* d = xyz + 3 * (size_t)o[k]; d[0] = h[3 * k]; d[1] = h[3 * k + 1]; d[2] = h[3 * k + 2]; }
		return ICPB_OK;
	});
}

int icpb_group_get_correspondences(icpb_group* g, int* idx)
{
	if (!g || !idx) return ICPB_ERR_BADARG;
	return for_each_rank(g, [&](int r) {
		const std::vector<int>& o = g->owned[(size_t)r];
		std::vector<int>& h = g->ibuf[(size_t)r];
		h.resize(o.size() + 1);
		const int rc = icpb_get_correspondences(H(g->ctx[(size_t)r]), h.data(), 0);
		if (rc != ICPB_OK) return rc;
		for (size_t k = 0; k < o.size(); k++) idx[(size_t)o[k]] = h[k];
		return ICPB_OK;
	});
}

} // extern "C"
