// icp_point_to_point — drop-in for the reference program src/ICP_point_to_point.cu.
//
// With no arguments it does what the reference binary does — 128x128 synthetic saddle, pose
// t=(0.8,-0.3,0.2) r=(0.2,-0.2,0.05), point-to-point ICP, at most 40 iterations, tolerance 1e-6 — and
// prints the same lines in the same format (src/ICP_point_to_point.cu:290,428-433 and its private
// printSarray :500-508); only the elapsed time differs. All device work goes through the C ABI of
// libicp_b200.so. Optional flags (--width, --n, --max-iter, --tol, --nn, --report) select the larger
// configurations of BASELINE.json; they do not exist in the reference.
#include "synth.h"
#include "icp_b200.h"
#include "runner.h"

int main(int argc, char** argv)
{
	synth::Options opt;
	if (!synth::parse(argc, argv, opt)) return 2;
	const int W = opt.width > 0 ? opt.width : 128;
	const int npts = opt.n > 0 ? opt.n : W * W;
	const int max_iter = opt.max_iter > 0 ? opt.max_iter : 40;

	synth::Clouds c = synth::point_to_point_clouds(W, npts);

	Runner run;                                        // --gpus N (N > 1): source sharded over N GPUs of this box, one process
	int rc = run.create(opt.gpus);
	if (rc != ICPB_OK) { printf("Error creating the ICP context: %s\n", icpb_status_string(rc)); return -1; }
	if ((rc = run.set_clouds(c.M.data(), npts, c.D.data(), npts)) != ICPB_OK) {
		printf("Error uploading the clouds: %s\n", run.last_error());
		return -1;
	}

	// the reference prints its own launch geometry here (GridSize = 16, BlockSize = NUM_POINTS / 16)
	printf("Grid Size: %d, Block Size: %d\n", 16, npts / 16);

	icpb_params p;
	icpb_default_params(&p);
	p.max_iter = max_iter;
	p.sync_every = opt.sync_every;
	if (opt.tol >= 0) p.tol = opt.tol;
	if (opt.grid_nn) p.nn_method = ICPB_NN_GRID;
	if (opt.report) p.flags |= ICPB_FLAG_PROFILE;     // per-iteration matching times for the report
	std::vector<float> err((size_t)max_iter + 1, 0.f);
	icpb_result res;
	rc = run.run(&p, err.data(), &res);
	if (rc != ICPB_OK) { printf("Error in the ICP loop: %s\n", run.last_error()); return -1; }

	printf("Error:\n");
	for (int i = 0; i < res.iterations + 1; i++) printf("%d: %.4f\n", i + 1, err[(size_t)i]);
	printf("\n");
	printf("ICP converged successfully!\n\n");
	printf("Elapsed time: %f ms\n", res.elapsed_ms);

	if (opt.report) {
		const double pairs = res.nn_pairs;
		printf("\n[report] points %d x %d, iterations run %d, matching %.3f ms, %.4e NN pairs/s, %.2f ICP iterations/s\n",
		       npts, npts, res.iterations_run, res.match_ms, pairs / (res.match_ms * 1e-3), res.iterations_run / (res.elapsed_ms * 1e-3));
		printf("[report] R (column-major): "); for (int k = 0; k < 9; k++) printf("%.7f ", res.R[k]);
		printf("\n[report] t: %.7f %.7f %.7f\n", res.t[0], res.t[1], res.t[2]);
		printf("[report] gpus %d (%s), %.3f ms per ICP iteration\n", run.gpus, run.exchange(), res.elapsed_ms / res.iterations_run);
	}
	run.destroy();
	return 0;
}
