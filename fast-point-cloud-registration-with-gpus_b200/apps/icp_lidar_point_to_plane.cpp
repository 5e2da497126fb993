// icp_lidar_point_to_plane — drop-in for the reference program src/CUDA/GPU_point_to_plane_real.cu.
//
// Same capture, conversion and target synthesis as icp_lidar_point_to_point (only the conversion time is printed, :817);
// PCA normals from the 4 nearest neighbours, then point-to-plane ICP with squared-distance matching (sentinel 100000,
// :196), at most 100 iterations, tolerance 1e-6; banners, normals time, error list and phase report as :362-643.
#include "dataset.h"

int main(int argc, char** argv)
{
	dataset::Options opt;
	if (!dataset::parse(argc, argv, opt)) return 2;
	const int npts = 16384;
	icpb_ctx* ctx = nullptr;
	int rc = icpb_create(&ctx, 0);
	if (rc != ICPB_OK) { printf("Error creating the ICP context: %s\n", icpb_status_string(rc)); return -1; }
	std::vector<float> D, M;
	if (dataset::load_lidar_clouds(ctx, opt.data_dir, npts, false, D, M) != 0) { printf("Error when reading LiDAR data\n"); icpb_destroy(ctx); return -1; }
	rc = dataset::register_clouds(ctx, opt, D, M, npts, true, 100, 100000.0f);
	icpb_destroy(ctx);
	return rc;
}
