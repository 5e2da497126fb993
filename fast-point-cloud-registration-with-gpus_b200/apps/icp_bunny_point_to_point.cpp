// icp_bunny_point_to_point — drop-in for the reference program src/CUDA/GPU_point_to_point_bunny.cu.
//
// Reads "Bunny_res.csv" (8171 points of the Stanford bunny, one "x y z" per line) from the working directory, builds
// the target by moving it with t = (0.01,-0.04,0.02), r = (0.15,-0.1,0.05) rad (:137-166), runs point-to-point ICP
// (squared-distance matching, at most 40 iterations, tolerance 1e-6) and prints what the reference prints: the launch
// geometry banner, the error list and the per-phase report (:268, 414-434). Options (ours): --data DIR, --nn grid,
// --max-iter K, --report.
#include "dataset.h"

int main(int argc, char** argv)
{
	dataset::Options opt;
	if (!dataset::parse(argc, argv, opt)) return 2;
	const int npts = 8171;
	std::vector<float> D, M;
	if (dataset::read_cloud_text(dataset::path_in(opt.data_dir, "Bunny_res.csv"), D) != 3 * npts) { printf("Error reading data\n"); return -1; }
	float ti[3] = { 0.01f, -0.04f, 0.02f }, ri[3] = { 0.15f, -0.1f, 0.05f }, r[9];
	synth::euler_rotation(ri, r);
	synth::move_rigid(D, npts, r, ti, M);

	icpb_ctx* ctx = nullptr;
	int rc = icpb_create(&ctx, 0);
	if (rc != ICPB_OK) { printf("Error creating the ICP context: %s\n", icpb_status_string(rc)); return -1; }
	rc = dataset::register_clouds(ctx, opt, D, M, npts, false, 40, 100000.0f);
	icpb_destroy(ctx);
	return rc;
}
