// nn_dispatch.cu — nearest-neighbour method dispatch.
#include "common.cuh"
namespace icpb {
int launch_match(Ctx* c, int dist_mode, int nn_method, float sentinel)
{
	if (nn_method == ICPB_NN_GRID) return launch_match_grid(c, dist_mode, sentinel);
	// the filter pays off once the per-block set-up is amortised (crossover ~3e4 x 3e4 points, tools/sweep_small.py)
	if (nn_method == ICPB_NN_BRUTE && c->k1_use_filter && dist_mode != ICPB_DIST_STD && (double)c->n * (double)c->m >= c->kf_min_pairs)
		return launch_match_filter(c, dist_mode, sentinel);
	return launch_match_brute(c, dist_mode, sentinel);     // ICPB_NN_BRUTE_DIRECT, or ICP_standard's FP64 formula
}
}
using namespace icpb;
extern "C" {
}
