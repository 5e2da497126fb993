// icp_lidar_point_to_point — drop-in for the reference program src/CUDA/GPU_point_to_point_real.cu.
//
// Reads 64 Ouster OS1-16 packets ("Donut_1024x16.csv", one byte per line) and the beam angles ("beam_intrinsics.csv")
// from the working directory, converts the 16 384 ranges to Cartesian points on the GPU, synthesises the target with a
// small rigid motion, scales both clouds from millimetres to metres and registers them with point-to-point ICP
// (squared-distance matching with sentinel 1e6, at most 100 iterations, tolerance 1e-6). Prints the reference's lines:
// the two kernel times (:572, 614), the banner (:237), the error list and the phase report (:386-403).
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
	if (dataset::load_lidar_clouds(ctx, opt.data_dir, npts, true, D, M) != 0) { printf("Error when reading LiDAR data\n"); icpb_destroy(ctx); return -1; }
	rc = dataset::register_clouds(ctx, opt, D, M, npts, false, 100, 1000000.0f);
	icpb_destroy(ctx);
	return rc;
}
