// dataset_dump — host-only: runs the dataset parsers of apps/dataset.h and writes what they read as raw binary, so the
// CPU test-suite can compare the product's loaders with the oracle's restatement and the reference's own readData.
//   dataset_dump cloud FILE out.bin                 floats of a whitespace-separated cloud file
//   dataset_dump lidar DIR out.bin                  16384 ranges, 16 altitude + 16 azimuth angles (floats), then the
//                                                   encoder count as one more float-sized pair (u64)
// No device work: nothing of libicp_b200.so is called.
#include "dataset.h"

int main(int argc, char** argv)
{
	if (argc != 4) { fprintf(stderr, "usage: %s cloud FILE out.bin | lidar DIR out.bin\n", argv[0]); return 2; }
	FILE* out = fopen(argv[3], "wb");
	if (!out) return 1;
	if (!strcmp(argv[1], "cloud")) {
		std::vector<float> v;
		if (dataset::read_cloud_text(argv[2], v) < 0) return 1;
		fwrite(v.data(), sizeof(float), v.size(), out);
	} else if (!strcmp(argv[1], "lidar")) {
		dataset::LidarCapture cap;
		const int n = 16384;
		if (dataset::read_lidar_packets(dataset::path_in(argv[2], "Donut_1024x16.csv"), n, cap) != n) return 1;
		if (dataset::read_beam_intrinsics(dataset::path_in(argv[2], "beam_intrinsics.csv"), cap) != 0) return 1;
		fwrite(cap.range.data(), sizeof(float), cap.range.size(), out);
		fwrite(cap.altitude, sizeof(float), 16, out);
		fwrite(cap.azimuth, sizeof(float), 16, out);
		fwrite(&cap.encoder_count, sizeof cap.encoder_count, 1, out);
	} else return 2;
	fclose(out);
	return 0;
}
