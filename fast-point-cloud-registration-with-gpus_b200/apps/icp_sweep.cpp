// icp_sweep — the reference's size sweeps, regenerated through the C ABI with the same CSV
// schemas so the B200 curves can be laid over the thesis' RTX 2060 ones:
//   --matching   src/CUDA/Matching_opt.cu:58-243     WIDTH 3..128, matching kernel alone, min of 10 CUDA-event timings,
//                                                    "#POINTS,TIME" rows -> Matching_loop_optimized.csv
//   --icp-point  src/CUDA/GPU_time_complexity_point.cu  one point-to-point iteration per WIDTH, "NUM_POINTS,TIME"
//                                                    -> GPU_ICP_point_to_point_TimeComp.csv
//   --icp-plane  src/CUDA/GPU_time_complexity_plane.cu  one point-to-plane iteration (normals excluded, as there)
//                                                    -> GPU_ICP_point_to_plane_TimeComp.csv
//   --phases W   the per-phase percentage report of src/CUDA/ICP_point_to_point_clean.cu:464-481 for one size
// Options: --min W --max W (default 3..128), --out DIR (default .), --direct (reference chain on every pair).
#include "synth.h"
#include "icp_b200.h"
#include <string>

static int fail(icpb_ctx* ctx, const char* what) { printf("%s: %s\n", what, ctx ? icpb_last_error(ctx) : "?"); return -1; }

int main(int argc, char** argv)
{
	int wmin = 3, wmax = 128, phases_w = 0, direct = 0;
	std::string mode, out = ".";
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		if (a == "--matching" || a == "--icp-point" || a == "--icp-plane") mode = a;
		else if (a == "--phases" && i + 1 < argc) { mode = a; phases_w = atoi(argv[++i]); }
		else if (a == "--min" && i + 1 < argc) wmin = atoi(argv[++i]);
		else if (a == "--max" && i + 1 < argc) wmax = atoi(argv[++i]);
		else if (a == "--out" && i + 1 < argc) out = argv[++i];
		else if (a == "--direct") direct = 1;
		else { fprintf(stderr, "usage: %s --matching|--icp-point|--icp-plane|--phases W [--min W] [--max W] [--out DIR] [--direct]\n", argv[0]); return 2; }
	}
	if (mode.empty()) { fprintf(stderr, "pick a sweep\n"); return 2; }
	icpb_ctx* ctx = nullptr;
	if (icpb_create(&ctx, 0) != ICPB_OK) { printf("no sm_100 device\n"); return -1; }
	const int nn = direct ? ICPB_NN_BRUTE_DIRECT : ICPB_NN_BRUTE;

	if (mode == "--phases") {
		const int W = phases_w > 0 ? phases_w : 128;
		synth::Clouds c = synth::point_to_point_clouds(W, W * W);
		if (icpb_set_target(ctx, c.M.data(), c.n, 0) != ICPB_OK || icpb_set_source(ctx, c.D.data(), c.n, 0) != ICPB_OK) return fail(ctx, "upload");
		icpb_params p; icpb_default_params(&p); p.nn_method = nn; p.flags |= ICPB_FLAG_PROFILE;
		std::vector<float> err((size_t)p.max_iter + 1);
		icpb_result res;
		if (icpb_run(ctx, &p, err.data(), &res) != ICPB_OK) return fail(ctx, "icpb_run");
		const double other = res.elapsed_ms - res.match_ms;
		printf("Points: %d, iterations: %d, elapsed %.4f ms\n", c.n, res.iterations_run, res.elapsed_ms);
		printf("Matching step: %.4f ms (%.2f %%)\n", res.match_ms, 100.0 * res.match_ms / res.elapsed_ms);
		printf("Minimization + transformation + error + convergence test: %.4f ms (%.2f %%)\n", other, 100.0 * other / res.elapsed_ms);
		icpb_destroy(ctx);
		return 0;
	}

	const char* fname = mode == "--matching" ? "Matching_loop_optimized.csv" : mode == "--icp-point" ? "GPU_ICP_point_to_point_TimeComp.csv" : "GPU_ICP_point_to_plane_TimeComp.csv";
	FILE* doc = fopen((out + "/" + fname).c_str(), "w");
	if (!doc) { printf("cannot write %s\n", fname); return -1; }
	fprintf(doc, mode == "--matching" ? "#POINTS,TIME\n" : "NUM_POINTS,TIME\n");
	printf(mode == "--matching" ? "#POINTS\tTIME\n" : "NUM_POINTS\tTIME\n");
	for (int W = wmin; W <= wmax; W++) {
		synth::Clouds c = synth::point_to_point_clouds(W, W * W);
		if (icpb_set_target(ctx, c.M.data(), c.n, 0) != ICPB_OK || icpb_set_source(ctx, c.D.data(), c.n, 0) != ICPB_OK) return fail(ctx, "upload");
		float t_ms = 0.f;
		if (mode == "--matching") {
			float mean = 0.f;
			if (icpb_time_match(ctx, ICPB_DIST_SQ, nn, 100000.0f, 10, &mean, &t_ms) != ICPB_OK) return fail(ctx, "icpb_time_match");
		} else {
			icpb_params p; icpb_default_params(&p);
			p.max_iter = 1; p.stop_early = 0; p.nn_method = nn;            // MAX_ITER 1, as GPU_time_complexity_point.cu:25
			if (mode == "--icp-plane") {
				p.metric = ICPB_POINT_TO_PLANE; p.dist_mode = ICPB_DIST_SQRT;
				if (W * W < 6) continue;                                     // fewer than k+1 points: the reference would read out of bounds
				if (icpb_estimate_normals(ctx, 4, nullptr) != ICPB_OK) return fail(ctx, "icpb_estimate_normals");
			}
			float err[2]; icpb_result res; float best = 1e30f;
			for (int rep = 0; rep < 10; rep++) {
				if (icpb_set_source(ctx, c.D.data(), c.n, 0) != ICPB_OK) return fail(ctx, "upload");
				int rc = icpb_run(ctx, &p, err, &res);
				if (rc == ICPB_ERR_NUMERIC) { best = -1.f; break; }         // degenerate tiny clouds (reference: potrf devInfo > 0, ignored)
				if (rc != ICPB_OK) return fail(ctx, "icpb_run");
				if (res.elapsed_ms < best) best = res.elapsed_ms;
			}
			t_ms = best;
		}
		fprintf(doc, "%d,%f\n", c.n, t_ms);
		printf("%d\t%f\n", c.n, t_ms);
	}
	fclose(doc);
	icpb_destroy(ctx);
	return 0;
}
