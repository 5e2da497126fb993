// pending.cu — entry points whose kernels land later in round 1 (normals, grid NN, batched).
#include "common.cuh"
namespace icpb {
int launch_match(Ctx* c, int dist_mode, int nn_method, float sentinel)
{
	if (nn_method == ICPB_NN_GRID) return launch_match_grid(c, dist_mode, sentinel);
	return launch_match_brute(c, dist_mode, sentinel);
}
int launch_match_grid(Ctx* c, int, float) { snprintf(c->err, sizeof c->err, "ICPB_NN_GRID: not built yet"); return ICPB_ERR_STATE; }
}
using namespace icpb;
extern "C" {
}
