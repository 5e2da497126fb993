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
#include <mutex>

namespace icpb {

struct NcclApi {
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
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
		api.CommDestroy    = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
		api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
		api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
		if (!api.ok) snprintf(api.why, sizeof api.why, "libnccl is missing a required symbol");
	});
	return api;
}

struct Dist {
	ncclComm_t comm = nullptr;
	int rank = 0, world = 1;
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
	ncclResult_t r = a.AllReduce(dev_buf, dev_buf, (size_t)count, ncclDouble, ncclSum, d->comm, s);
	if (r != ncclSuccess) { snprintf(err, errlen, "ncclAllReduce: %s", a.GetErrorString(r)); return ICPB_ERR_NCCL; }
	return ICPB_OK;
}

void dist_destroy(Dist* d)
{
	if (!d) return;
	NcclApi& a = nccl_api();
	if (a.ok && d->comm) a.CommDestroy(d->comm);
	delete d;
}

} // namespace icpb
