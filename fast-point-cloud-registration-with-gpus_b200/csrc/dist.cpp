// dist.cpp — NCCL plumbing for source-sharded multi-GPU ICP.
//
// The reference has no multi-GPU code at all. Here every rank owns a contiguous
// shard of the source cloud and a replica of the target; the only exchange per iteration is an
// ncclAllReduce(sum, double) of the 16 (point-to-point) or 28 (point-to-plane) raw moment sums and
// of the squared-residual sum. Every rank then solves the same 3x3 / 6x6 problem on identical bits,
// so no broadcast of R,T is needed.
//
// libnccl is bound lazily with dlopen so that libicp_b200.so loads (and every single-GPU entry
// point works) on machines without NCCL; inside a torch process the already-loaded
// libnccl.so.2 of torch is the one that gets used.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <mutex>

namespace icpb {

struct NcclApi {
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
	ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
	char why[256] = {0};
};

static NcclApi& nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		const char* names[] = { "libnccl.so.2", "libnccl.so", nullptr };
		for (int k = 0; names[k] && !api.handle; k++) api.handle = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
		if (!api.handle) { snprintf(api.why, sizeof api.why, "dlopen(libnccl.so.2) failed: %s", dlerror()); return; }
		api.GetUniqueId    = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
		api.CommInitRank   = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
		api.AllReduce      = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
		api.AllGather      = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
		api.CommDestroy    = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
		api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
		api.CommInitAll    = (decltype(api.CommInitAll))dlsym(api.handle, "ncclCommInitAll");            // optional: in-process groups on the NCCL path
		api.CommGetAsyncError = (decltype(api.CommGetAsyncError))dlsym(api.handle, "ncclCommGetAsyncError");   // optional: failure detection
		api.CommAbort      = (decltype(api.CommAbort))dlsym(api.handle, "ncclCommAbort");
		api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.AllGather && api.CommDestroy && api.GetErrorString;
		if (!api.ok) snprintf(api.why, sizeof api.why, "libnccl is missing a required symbol");
	});
	return api;
}

struct Dist {
	ncclComm_t comm = nullptr;
	int rank = 0, world = 1;
	// peer-memory exchange (peer_exchange.cuh)
	double* mailbox = nullptr;                 // this rank's mailbox (cudaMalloc)
	u64*    seq = nullptr;
	void*   peer_ptr[PEER_MAX] = {nullptr};    // cudaIpcOpenMemHandle mappings of the other ranks' mailboxes
};

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");

int dist_unique_id(void* id128)
{
	NcclApi& a = nccl_api();
	if (!a.ok) return ICPB_ERR_NCCL;
	ncclUniqueId id;
	if (a.GetUniqueId(&id) != ncclSuccess) return ICPB_ERR_NCCL;
	memcpy(id128, &id, sizeof id);
	return ICPB_OK;
}

int dist_init(Dist** out, int rank, int world, const void* id128, char* err, size_t errlen)
{
	NcclApi& a = nccl_api();
	if (!a.ok) { snprintf(err, errlen, "%s", a.why); return ICPB_ERR_NCCL; }
	Dist* d = new Dist();
	d->rank = rank; d->world = world;
	ncclUniqueId id;
	memcpy(&id, id128, sizeof id);
	ncclResult_t r = a.CommInitRank(&d->comm, world, id, rank);
	if (r != ncclSuccess) { snprintf(err, errlen, "ncclCommInitRank: %s", a.GetErrorString(r)); delete d; return ICPB_ERR_NCCL; }
	*out = d;
	return ICPB_OK;
}

int dist_allreduce_f64(Dist* d, double* dev_buf, int count, cudaStream_t s, char* err, size_t errlen)
{
	NcclApi& a = nccl_api();
	if (!d || !d->comm) { snprintf(err, errlen, "no NCCL communicator on this rank (peer-memory group without NCCL)"); return ICPB_ERR_NCCL; }
	ncclResult_t r = a.AllReduce(dev_buf, dev_buf, (size_t)count, ncclDouble, ncclSum, d->comm, s);
	if (r != ncclSuccess) { snprintf(err, errlen, "ncclAllReduce: %s", a.GetErrorString(r)); return ICPB_ERR_NCCL; }
	return ICPB_OK;
}

// One mailbox per rank, mapped into every peer with CUDA IPC. The handles travel through ncclAllGather (the only
// channel the C ABI has between ranks), and an ncclAllReduce(min) makes the decision unanimous: the fused exchange is
// used only if EVERY rank could map EVERY mailbox (one process per GPU on one NVLink/NVSwitch node). ICPB_PEER=0
// forces the NCCL path (A/B measurements).
int dist_peer_init(Dist* d, int device, cudaStream_t s, PeerXchg* out, char* err, size_t errlen)
{
	memset(out, 0, sizeof *out);
	if (!d || d->world < 2) return ICPB_OK;
	NcclApi& a = nccl_api();
	int mine = (d->world <= PEER_MAX) ? 1 : 0;
	if (const char* e = getenv("ICPB_PEER")) if (atoi(e) == 0) mine = 0;
	const size_t mbytes = sizeof(double) * 2 * PEER_MAX * PEER_ROW;
	cudaIpcMemHandle_t* d_handles = nullptr;      // [world] on the device for the allgather
	int* d_flag = nullptr;
	cudaIpcMemHandle_t h_handles[PEER_MAX];
	memset(h_handles, 0, sizeof h_handles);
	auto cleanup = [&] { cudaFree(d_handles); cudaFree(d_flag); };
	if (cudaMalloc((void**)&d_flag, sizeof(int)) != cudaSuccess) { snprintf(err, errlen, "dist_peer_init: cudaMalloc failed"); return ICPB_ERR_NOMEM; }
	if (mine) {
		if (cudaMalloc((void**)&d->mailbox, mbytes) != cudaSuccess || cudaMalloc((void**)&d->seq, sizeof(u64)) != cudaSuccess ||
		    cudaMalloc((void**)&d_handles, sizeof(cudaIpcMemHandle_t) * PEER_MAX) != cudaSuccess) mine = 0;
	}
	if (mine) {
		cudaMemsetAsync(d->mailbox, 0, mbytes, s); cudaMemsetAsync(d->seq, 0, sizeof(u64), s);
		if (cudaIpcGetMemHandle(&h_handles[d->rank], d->mailbox) != cudaSuccess) { cudaGetLastError(); mine = 0; }
	}
	// every rank takes part in the collectives whatever its own state: the calls must match across ranks. With more than
	// PEER_MAX ranks the fused exchange is off for everybody (mine == 0 above) and the handle staging is skipped altogether:
	// h_handles / d_handles hold PEER_MAX entries, a rank >= PEER_MAX has no slot in them.
	if (d->world <= PEER_MAX) {
		if (!d_handles && cudaMalloc((void**)&d_handles, sizeof(cudaIpcMemHandle_t) * PEER_MAX) != cudaSuccess) { cleanup(); snprintf(err, errlen, "dist_peer_init: cudaMalloc failed"); return ICPB_ERR_NOMEM; }
		cudaMemcpyAsync(d_handles + d->rank, &h_handles[d->rank], sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice, s);
		ncclResult_t r = a.AllGather(d_handles + d->rank, d_handles, sizeof(cudaIpcMemHandle_t), ncclChar, d->comm, s);
		if (r != ncclSuccess) { cleanup(); snprintf(err, errlen, "ncclAllGather: %s", a.GetErrorString(r)); return ICPB_ERR_NCCL; }
		cudaMemcpyAsync(h_handles, d_handles, sizeof(cudaIpcMemHandle_t) * d->world, cudaMemcpyDeviceToHost, s);
	}
	if (cudaStreamSynchronize(s) != cudaSuccess) { cleanup(); snprintf(err, errlen, "dist_peer_init: stream error"); return ICPB_ERR_CUDA; }
	for (int r = 0; mine && r < d->world; r++) {
		if (r == d->rank) continue;
		if (cudaIpcOpenMemHandle(&d->peer_ptr[r], h_handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); d->peer_ptr[r] = nullptr; mine = 0; }
	}
	cudaMemcpyAsync(d_flag, &mine, sizeof(int), cudaMemcpyHostToDevice, s);
	ncclResult_t r = a.AllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, d->comm, s);
	if (r != ncclSuccess) { cleanup(); snprintf(err, errlen, "ncclAllReduce: %s", a.GetErrorString(r)); return ICPB_ERR_NCCL; }
	int all = 0;
	cudaMemcpyAsync(&all, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s);
	if (cudaStreamSynchronize(s) != cudaSuccess) { cleanup(); snprintf(err, errlen, "dist_peer_init: stream error"); return ICPB_ERR_CUDA; }
	cleanup();
	if (!all) return ICPB_OK;                 // somebody could not: everybody stays on ncclAllReduce
	out->rank = d->rank; out->world = d->world; out->seq = d->seq;
	for (int k = 0; k < d->world; k++) out->mailbox[k] = (k == d->rank) ? d->mailbox : reinterpret_cast<double*>(d->peer_ptr[k]);
	return ICPB_OK;
}

// Failure detection on the NCCL path: an asynchronous communicator error (a peer process died, a link
// went down) otherwise shows up as a collective that never completes. Polled whenever the engine synchronises with
// the stream anyway; on error the communicator is aborted so that pending collectives return instead of hanging.
int dist_check_async(Dist* d, char* err, size_t errlen)
{
	if (!d || !d->comm) return ICPB_OK;
	NcclApi& a = nccl_api();
	if (!a.CommGetAsyncError) return ICPB_OK;
	ncclResult_t st = ncclSuccess;
	const ncclResult_t r = a.CommGetAsyncError(d->comm, &st);
	if (r != ncclSuccess) { snprintf(err, errlen, "ncclCommGetAsyncError: %s", a.GetErrorString(r)); return ICPB_ERR_NCCL; }
	if (st != ncclSuccess && st != ncclInProgress) {
		snprintf(err, errlen, "NCCL asynchronous error on rank %d of %d: %s", d->rank, d->world, a.GetErrorString(st));
		if (a.CommAbort) { a.CommAbort(d->comm); d->comm = nullptr; }
		return ICPB_ERR_NCCL;
	}
	return ICPB_OK;
}
bool dist_has_comm(const Dist* d) { return d && d->comm; }

// In-process group (icpb_group_create): `world` ranks, one per device, all in THIS process — no launcher, no torch, no
// MPI, no unique id. Fused path (default): every device enables peer access to every other one and the mailboxes are
// plain cudaMalloc allocations that the reduction kernels of all devices address directly over NVLink (no CUDA IPC: one
// address space). Otherwise (ICPB_PEER=0, more than PEER_MAX devices, or a pair without peer access): one
// ncclCommInitAll communicator set and ncclAllReduce launches between the kernels.
int dist_init_local(Dist** outs, PeerXchg* peers, const int* devices, int world, char* err, size_t errlen)
{
	for (int r = 0; r < world; r++) { outs[r] = nullptr; memset(&peers[r], 0, sizeof(PeerXchg)); }
	if (world < 2) return ICPB_OK;
	bool peer_ok = world <= PEER_MAX;
	if (const char* e = getenv("ICPB_PEER")) if (atoi(e) == 0) peer_ok = false;
	for (int a = 0; peer_ok && a < world; a++)
		for (int b = 0; peer_ok && b < world; b++) {
			if (a == b) continue;
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can) { cudaGetLastError(); peer_ok = false; }
		}
	for (int r = 0; r < world; r++) { outs[r] = new Dist(); outs[r]->rank = r; outs[r]->world = world; }
	auto fail_all = [&](int code) { for (int r = 0; r < world; r++) { if (outs[r]) { cudaSetDevice(devices[r]); dist_destroy(outs[r]); outs[r] = nullptr; } memset(&peers[r], 0, sizeof(PeerXchg)); } return code; };
	if (peer_ok) {
		const size_t mbytes = sizeof(double) * 2 * PEER_MAX * PEER_ROW;
		for (int a = 0; a < world; a++) {
			if (cudaSetDevice(devices[a]) != cudaSuccess) { snprintf(err, errlen, "dist_init_local: cudaSetDevice(%d) failed", devices[a]); return fail_all(ICPB_ERR_CUDA); }
			for (int b = 0; b < world; b++) {
				if (a == b) continue;
				const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { snprintf(err, errlen, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[a], devices[b], cudaGetErrorString(e)); return fail_all(ICPB_ERR_CUDA); }
				cudaGetLastError();
			}
			if (cudaMalloc((void**)&outs[a]->mailbox, mbytes) != cudaSuccess || cudaMalloc((void**)&outs[a]->seq, sizeof(u64)) != cudaSuccess) { snprintf(err, errlen, "dist_init_local: cudaMalloc failed"); return fail_all(ICPB_ERR_NOMEM); }
			cudaMemset(outs[a]->mailbox, 0, mbytes); cudaMemset(outs[a]->seq, 0, sizeof(u64));
			if (cudaDeviceSynchronize() != cudaSuccess) { snprintf(err, errlen, "dist_init_local: device error"); return fail_all(ICPB_ERR_CUDA); }
		}
		for (int a = 0; a < world; a++) {
			peers[a].rank = a; peers[a].world = world; peers[a].seq = outs[a]->seq;
			for (int b = 0; b < world; b++) peers[a].mailbox[b] = outs[b]->mailbox;
		}
		return ICPB_OK;
	}
	NcclApi& api = nccl_api();
	if (!api.ok || !api.CommInitAll) { snprintf(err, errlen, "no peer access between the devices and no usable libnccl (%s)", api.ok ? "ncclCommInitAll missing" : api.why); return fail_all(ICPB_ERR_NCCL); }
	ncclComm_t* comms = new ncclComm_t[world];
	const ncclResult_t r = api.CommInitAll(comms, world, devices);
	if (r != ncclSuccess) { snprintf(err, errlen, "ncclCommInitAll: %s", api.GetErrorString(r)); delete[] comms; return fail_all(ICPB_ERR_NCCL); }
	for (int a = 0; a < world; a++) outs[a]->comm = comms[a];
	delete[] comms;
	return ICPB_OK;
}

void dist_destroy(Dist* d)
{
	if (!d) return;
	for (int r = 0; r < PEER_MAX; r++) if (d->peer_ptr[r]) cudaIpcCloseMemHandle(d->peer_ptr[r]);
	cudaFree(d->mailbox); cudaFree(d->seq);
	NcclApi& a = nccl_api();
	if (a.ok && d->comm) a.CommDestroy(d->comm);
	delete d;
}

} // namespace icpb
