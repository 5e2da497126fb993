// runner.h — one registration on 1 GPU (icpb_create) or on several GPUs of the box (icpb_group_create, --gpus N):
// the same four calls either way, so the drop-in mains stay the reference's straight-line shape
// (src/ICP_point_to_point.cu:90-460). Host C++ over the C ABI only.
#pragma once
#include <cstdio>
#include "icp_b200.h"

struct Runner {
	icpb_ctx* ctx = nullptr;
	icpb_group* grp = nullptr;
	int gpus = 1;

	int create(int ngpus)
	{
		gpus = ngpus < 1 ? 1 : ngpus;
		return gpus == 1 ? icpb_create(&ctx, 0) : icpb_group_create(&grp, nullptr, gpus);
	}
	int set_clouds(const float* target, int m, const float* source, int n)
	{
		int rc;
		if (grp) {
			if ((rc = icpb_group_set_target(grp, target, m)) != ICPB_OK) return rc;
			return icpb_group_set_source(grp, source, n, 0);           // contiguous shards: each GPU's sources stay a compact piece of the cloud (what the Morton-ordered matching kernel rewards)
		}
		if ((rc = icpb_set_target(ctx, target, m, 0)) != ICPB_OK) return rc;
		return icpb_set_source(ctx, source, n, 0);
	}
	int normals(int k, float* ms) { return grp ? icpb_group_estimate_normals(grp, k, ICPB_DIST_SQRT, ms) : icpb_estimate_normals(ctx, k, ms); }
	int run(const icpb_params* p, float* err, icpb_result* res) { return grp ? icpb_group_run(grp, p, err, res) : icpb_run(ctx, p, err, res); }
	const char* last_error() const { return grp ? icpb_group_last_error(grp) : icpb_last_error(ctx); }
	const char* exchange() const
	{
		if (!grp) return "single GPU";
		int nd = 0, peer = 0;
		icpb_group_info(grp, &nd, &peer, nullptr);
		return peer ? "moment sums exchanged inside the reduction kernels over NVLink peer memory" : "moment sums combined with ncclAllReduce";
	}
	void destroy() { if (grp) icpb_group_destroy(grp); if (ctx) icpb_destroy(ctx); grp = nullptr; ctx = nullptr; }
};
