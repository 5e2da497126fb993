// icp_bunny_point_to_plane — drop-in for the reference program src/CUDA/GPU_point_to_plane_bunny.cu.
//
// Same data and pose as icp_bunny_point_to_point; normals of the target from its 4 nearest neighbours (k-NN ranked on
// squared distances, :47-82) by PCA, then point-to-plane ICP with squared-distance matching (:187-231), at most 40
// iterations, tolerance 1e-6. Prints the reference's banners, the normals time, the error list and the phase report
// (:391, 440, 525, 656-673).
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
	rc = dataset::register_clouds(ctx, opt, D, M, npts, true, 40, 100000.0f);
	icpb_destroy(ctx);
	return rc;
}
