// icp_point_to_plane — drop-in for the reference program src/ICP_point_to_plane.cu.
//
// No arguments: 128x128 synthetic saddle, pose t=(0.8,-0.3,0.2) r=(0.2,-0.2,0.05); normals of the target
// from the 4 nearest neighbours by PCA; point-to-plane ICP (sqrt matching, 6x6 normal equations, Cholesky,
// small-angle Euler rotation), at most 50 iterations, tolerance 1e-6. Prints what the reference prints
// (src/ICP_point_to_plane.cu:381,423-427,513,623,636-641): the two launch-geometry banners, the normals
// timing line, one "Current error" line per iteration, the error array on one line, the success line and
// the elapsed time.
#include "synth.h"
#include "icp_b200.h"
#include "runner.h"

int main(int argc, char** argv)
{
	synth::Options opt;
	if (!synth::parse(argc, argv, opt)) return 2;
	const int W = opt.width > 0 ? opt.width : 128;
	const int npts = opt.n > 0 ? opt.n : W * W;
	const int max_iter = opt.max_iter > 0 ? opt.max_iter : 50;

	synth::Clouds c = synth::point_to_point_clouds(W, npts);

	Runner run;                                        // --gpus N (N > 1): source sharded over N GPUs of this box, one process
	int rc = run.create(opt.gpus);
	if (rc != ICPB_OK) { printf("Error creating the ICP context: %s\n", icpb_status_string(rc)); return -1; }
	if ((rc = run.set_clouds(c.M.data(), npts, c.D.data(), npts)) != ICPB_OK) {
		printf("Error uploading the clouds: %s\n", run.last_error());
		return -1;
	}

	printf("For normals:\nGrid Size: %d, Block Size: %d\n", 16, npts / 16);
	float normals_ms = 0.f;
	rc = run.normals(4, &normals_ms);
	if (rc != ICPB_OK) { printf("Error in knn kernel: %s\n", run.last_error()); return -1; }
	printf("\n");
	printf("Normals were calculated in %f ms\n\n", normals_ms);

	printf("For ICP loop:\nGrid Size: %d, Block Size: %d\n", 16, npts / 16);
	icpb_params p;
	icpb_default_params(&p);
	p.metric = ICPB_POINT_TO_PLANE;
	p.dist_mode = ICPB_DIST_SQRT;
	p.max_iter = max_iter;
	p.sync_every = opt.sync_every;
	if (opt.tol >= 0) p.tol = opt.tol;
	if (opt.grid_nn) p.nn_method = ICPB_NN_GRID;
	if (opt.report) p.flags |= ICPB_FLAG_PROFILE;     // per-iteration matching times for the report
	std::vector<float> err((size_t)max_iter + 1, 0.f);
	icpb_result res;
	rc = run.run(&p, err.data(), &res);
	if (rc != ICPB_OK) { printf("Error in the ICP loop: %s\n", run.last_error()); return -1; }

	// the reference prints each error as soon as it is known (src/ICP_point_to_plane.cu:623)
	for (int k = 1; k <= res.iterations_run; k++) printf("Current error (%d): %.4f\n", k, err[(size_t)k]);
	printf("Error:\n");
	for (int i = 0; i < res.iterations + 1; i++) printf("%.4f ", err[(size_t)i]);
	printf("\n");
	printf("ICP converged successfully!\n\n");
	printf("Elapsed time: %f ms\n", res.elapsed_ms);
	if (opt.report) {
		printf("\n[report] points %d x %d, iterations run %d, normals %.3f ms, matching %.3f ms, %.4e NN pairs/s\n",
		       npts, npts, res.iterations_run, normals_ms, res.match_ms, res.nn_pairs / (res.match_ms * 1e-3));
		printf("[report] R (column-major): "); for (int k = 0; k < 9; k++) printf("%.7f ", res.R[k]);
		printf("\n[report] t: %.7f %.7f %.7f\n", res.t[0], res.t[1], res.t[2]);
		printf("[report] gpus %d (%s), %.3f ms per ICP iteration\n", run.gpus, run.exchange(), res.elapsed_ms / res.iterations_run);
	}
	run.destroy();
	return 0;
}
