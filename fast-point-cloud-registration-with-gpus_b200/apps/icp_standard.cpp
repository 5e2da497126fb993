// icp_standard (also installed as icp_test, the name CMakeLists.txt:26 gives it) — drop-in for the
// reference program src/ICP_standard.cu: 32x32 saddle, hard-coded pose, double pow/sqrt matching,
// exactly 40 iterations with no convergence test, errors printed on one line with my_lib's printSarray
// (src/ICP_standard.cu:358,472-475).
#include "synth.h"
#include "icp_b200.h"

int main(int argc, char** argv)
{
	synth::Options opt;
	if (!synth::parse(argc, argv, opt)) return 2;
	const int W = opt.width > 0 ? opt.width : 32;
	const int max_iter = opt.max_iter > 0 ? opt.max_iter : 40;
	synth::Clouds c = synth::standard_clouds(W);
	const int npts = c.n;

	icpb_ctx* ctx = nullptr;
	int rc = icpb_create(&ctx, 0);
	if (rc != ICPB_OK) { printf("Error creating the ICP context: %s\n", icpb_status_string(rc)); return -1; }
	if ((rc = icpb_set_target(ctx, c.M.data(), npts, 0)) != ICPB_OK || (rc = icpb_set_source(ctx, c.D.data(), npts, 0)) != ICPB_OK) {
		printf("Error uploading the clouds: %s\n", icpb_last_error(ctx));
		return -1;
	}
	printf("Grid Size: %d, Block Size: %d\n", 8, npts / 8);

	icpb_params p;
	icpb_default_params(&p);
	p.dist_mode = ICPB_DIST_STD;
	p.max_iter = max_iter;
	p.stop_early = 0;
	p.sync_every = max_iter;       // no convergence test: enqueue the whole loop at once
	std::vector<float> err((size_t)max_iter + 1, 0.f);
	icpb_result res;
	rc = icpb_run(ctx, &p, err.data(), &res);
	if (rc != ICPB_OK) { printf("Error in the ICP loop: %s\n", icpb_last_error(ctx)); return -1; }

	printf("Error:\n");
	printSarray(err.data() + 1, max_iter);   // the reference stores iteration k's error in slot k (0-based)
	printf("Elapsed time: %f ms\n", res.elapsed_ms);
	icpb_destroy(ctx);
	return 0;
}
