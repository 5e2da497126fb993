// dataset.h — host-side front ends of the reference's dataset programs, shared by the
// icp_bunny_* and icp_lidar_* executables. Restates (does not copy):
//   readData       src/CUDA/GPU_point_to_point_bunny.cu:463-497   whitespace-separated floats, x0 y0 z0 x1 ...
//   Read_data      src/CUDA/GPU_point_to_point_real.cu:432-528    64 OS1-16 packets dumped one byte per line + beam angles
//   the phase report of the instrumented programs                 src/CUDA/GPU_point_to_point_bunny.cu:417-434
// Device work (conversion, transform, scaling, ICP) goes through the C ABI of libicp_b200.so.
#pragma once
#include "synth.h"
#include "icp_b200.h"

namespace dataset {

// The reference opens its files by bare name in the working directory; `--data DIR` (ours) prefixes a directory.
inline std::string path_in(const std::string& dir, const char* name) { return dir.empty() ? std::string(name) : dir + "/" + name; }

// Every token of every line (separators ' ' and '\n', lines up to 2047 characters) is one float, in file order.
// The reference trusts the file to hold exactly 3*num_points values; here the count is returned and checked.
inline int read_cloud_text(const std::string& path, std::vector<float>& out)
{
	FILE* doc = fopen(path.c_str(), "r");
	if (!doc) { perror("File opening failed\n"); return -1; }
	char line[2048];
	out.clear();
	while (fgets(line, sizeof line, doc) != NULL) {
		char* save = nullptr;
		for (char* tok = strtok_r(line, " \n", &save); tok; tok = strtok_r(nullptr, " \n", &save)) out.push_back(strtof(tok, nullptr));
	}
	fclose(doc);
	return (int)out.size();
}

struct LidarCapture {
	std::vector<float> range;             // one range word per fired beam, millimetres, block-major then beam
	unsigned long long encoder_count = 0; // encoder count of the first azimuth block
	float altitude[16] = {0}, azimuth[16] = {0};
};

// Capture layout (one byte per text line, 1-based line numbers): packet k starts at line 1 + 12608k, azimuth block b
// of it at + 788b; a block is a 16-byte header (bytes 13-14 of the first one = low 16 bits of the encoder count) and
// 64 channel records of 12 bytes whose first three bytes are the 20-bit range (little endian). The OS1-16 fills
// channels 2, 6, ..., 62. Like the reference, a range is committed when the line after its third byte is read.
inline int read_lidar_packets(const std::string& path, int n, LidarCapture& cap)
{
	FILE* doc = fopen(path.c_str(), "r");
	if (!doc) { perror("File opening failed"); return -1; }
	char line[128];
	cap.range.assign((size_t)n, 0.f);
	unsigned long enc = 0, word = 0;
	int stored = 0, channel = 2, block = 0, packet = 0, lineno = 1;
	while (fgets(line, sizeof line, doc) != NULL) {
		const int v = atoi(line);
		if (lineno == 13) enc = (unsigned long)v;
		if (lineno == 14) enc = (unsigned long)(v << 8) | enc;
		const int at = 17 + 12 * channel + 788 * block + 12608 * packet;
		if (lineno == at) word = (unsigned long)v;
		if (lineno == at + 1) word = (unsigned long)(v << 8) | word;
		if (lineno == at + 2) word = (unsigned long)((v & 0xF) << 16) | word;
		if (lineno > at + 2) { if (stored < n) cap.range[(size_t)stored] = (float)word; stored++; channel += 4; }
		if (channel >= 64) { channel = 2; block++; }
		if (block >= 16) { block = 0; packet++; }
		if (packet >= 64) break;
		lineno++;
	}
	fclose(doc);
	cap.encoder_count = enc;
	return stored;
}

// beam_intrinsics.csv: a header line then 64 altitude angles (lines 2-65), a blank line, a header line and 64 azimuth
// angles (lines 68-131); the 16 fitted beams are lines 4, 8, ..., 64 and 70, 74, ..., 130.
inline int read_beam_intrinsics(const std::string& path, LidarCapture& cap)
{
	FILE* doc = fopen(path.c_str(), "r");
	if (!doc) { perror("File opening failed"); return -1; }
	char line[128];
	int lineno = 1, k = 0;
	while (fgets(line, sizeof line, doc) != NULL) {
		if (lineno == 2) k = 0;
		if (lineno >= 2 && lineno <= 65 && lineno % 4 == 0 && k < 16) cap.altitude[k++] = (float)atof(line);
		if (lineno == 68) k = 0;
		if (lineno >= 68 && lineno <= 131 && (lineno - 66) % 4 == 0 && k < 16) cap.azimuth[k++] = (float)atof(line);
		lineno++;
	}
	fclose(doc);
	return 0;
}

// The LiDAR programs' Read_data: parse, convert on the device, synthesise the target with RyT (pose in the capture's own
// millimetre frame: t = (0.001,-0.0202,0.02), r = (0.01,-0.003,0.05) rad), then scale both clouds to metres.
// Prints the kernel-time lines the reference prints (`print_ryt`: only the point-to-point program prints the second).
inline int load_lidar_clouds(icpb_ctx* ctx, const std::string& dir, int n, bool print_ryt, std::vector<float>& D, std::vector<float>& M)
{
	LidarCapture cap;
	if (read_lidar_packets(path_in(dir, "Donut_1024x16.csv"), n, cap) < 0) return -1;
	if (read_beam_intrinsics(path_in(dir, "beam_intrinsics.csv"), cap) < 0) return -1;
	D.assign(3 * (size_t)n, 0.f); M.assign(3 * (size_t)n, 0.f);
	float ms = 0.f;
	if (icpb_lidar_convert(ctx, cap.range.data(), n, cap.encoder_count, cap.altitude, cap.azimuth, 16, 88, 90112, D.data(), 0, &ms) != ICPB_OK) {
		printf("Error in Conversion kernel: %s\n", icpb_last_error(ctx));
		return -1;
	}
	printf("Conversion kernel's elapsed time: %.3f ms\n", ms);
	float T[3] = { 0.001f, -0.0202f, 0.02f }, ri[3] = { 0.01f, -0.003f, 0.05f }, R[9];
	synth::euler_rotation(ri, R);
	if (icpb_apply_transform(ctx, R, T, D.data(), n, M.data(), 0, &ms) != ICPB_OK) {
		printf("Error in RyT kernel: %s\n", icpb_last_error(ctx));
		return -1;
	}
	if (print_ryt) printf("RyT kernel's elapsed time: %.3f ms\n", ms);
	// convert from millimetres to metres (cublasSscal, alpha = 1.0 / 1000.0 rounded to float)
	const float alpha = (float)(1.0 / 1000.0);
	if (icpb_scale_cloud(ctx, alpha, D.data(), n, 0) != ICPB_OK || icpb_scale_cloud(ctx, alpha, M.data(), n, 0) != ICPB_OK) {
		printf("Error scaling the clouds: %s\n", icpb_last_error(ctx));
		return -1;
	}
	return 0;
}

// Error list + the phase report of the instrumented programs. `count_run`: the point-to-point programs report
// iteration + 1 passes, the point-to-plane ones print `iteration` (src/CUDA/GPU_point_to_plane_bunny.cu:660-661).
// The engine fuses transformation, error and the stop test into one kernel: that time is reported on the
// transformation line and the error-estimation line shows what is left of the loop (event gaps, host reads).
inline void print_report(const std::vector<float>& err, const icpb_result& res, bool count_run)
{
	printf("Error:\n");
	for (int i = 0; i < res.iterations + 1; i++) printf("%d: %.4f\n", i + 1, err[(size_t)i]);
	printf("\n");
	const double total = res.elapsed_ms > 0 ? res.elapsed_ms : 1.0;
	double rest = total - res.match_ms - res.minimize_ms - res.transform_ms;
	if (rest < 0) rest = 0;
	printf("\nThe ICP algorithm was computed in %.4f ms with %d iterations\n\n", (double)res.elapsed_ms, count_run ? res.iterations + 1 : res.iterations);
	printf("The matching step represents the %.4f%% of the total time with %.4f ms\n\n", res.match_ms * 100.0 / total, (double)res.match_ms);
	printf("The minimization step represents the %.4f%% of the total time with %.4f ms\n\n", res.minimize_ms * 100.0 / total, (double)res.minimize_ms);
	printf("The transformation step represents the %.4f%% of the total time with %.4f ms\n\n", res.transform_ms * 100.0 / total, (double)res.transform_ms);
	printf("The error estimation step represents the %.4f%% of the total time with %.4f ms\n\n", rest * 100.0 / total, rest);
}

struct Options { std::string data_dir; int grid_nn = 0, report = 0, max_iter = 0; };
inline bool parse(int argc, char** argv, Options& o)
{
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		if (a == "--data" && i + 1 < argc) o.data_dir = argv[++i];
		else if (a == "--nn" && i + 1 < argc) o.grid_nn = (strcmp(argv[++i], "grid") == 0);
		else if (a == "--max-iter" && i + 1 < argc) o.max_iter = atoi(argv[++i]);
		else if (a == "--report") o.report = 1;
		else { fprintf(stderr, "usage: %s [--data DIR] [--nn brute|grid] [--max-iter K] [--report]\n", argv[0]); return false; }
	}
	return true;
}

inline void print_extra_report(const icpb_result& res, int n, int m)
{
	printf("[report] points %d x %d, iterations run %d, %.4e NN pairs/s\n", n, m, res.iterations_run, res.match_ms > 0 ? res.nn_pairs / (res.match_ms * 1e-3) : 0.0);
	printf("[report] R (column-major): "); for (int k = 0; k < 9; k++) printf("%.7f ", res.R[k]);
	printf("\n[report] t: %.7f %.7f %.7f\n", res.t[0], res.t[1], res.t[2]);
}

// Shared tail of the four programs: upload, (normals,) loop, report.
inline int register_clouds(icpb_ctx* ctx, const Options& opt, std::vector<float>& D, std::vector<float>& M, int npts, bool plane, int max_iter, float sentinel)
{
	int rc;
	if ((rc = icpb_set_target(ctx, M.data(), npts, 0)) != ICPB_OK || (rc = icpb_set_source(ctx, D.data(), npts, 0)) != ICPB_OK) {
		printf("Error uploading the clouds: %s\n", icpb_last_error(ctx));
		return -1;
	}
	const int grid = 60, block = npts / grid + 1;      // the reference's launch geometry, printed for fidelity
	if (plane) {
		printf("For normals:\nGrid Size: %d, Block Size: %d\n", grid, block);
		float normals_ms = 0.f;
		if (icpb_estimate_normals_ex(ctx, 4, ICPB_DIST_SQ, &normals_ms) != ICPB_OK) { printf("Error in knn kernel: %s\n", icpb_last_error(ctx)); return -1; }
		printf("Normals were calculated in %f ms\n\n", normals_ms);
		printf("For ICP loop:\nGrid Size: %d, Block Size: %d\n", grid, block);
	} else {
		printf("Grid Size: %d, Block Size: %d\n", grid, block);
	}
	icpb_params p;
	icpb_default_params(&p);
	p.metric = plane ? ICPB_POINT_TO_PLANE : ICPB_POINT_TO_POINT;
	p.dist_mode = ICPB_DIST_SQ;                        // both families match on squared distances
	p.max_iter = opt.max_iter > 0 ? opt.max_iter : max_iter;
	p.sentinel = sentinel;
	p.flags |= ICPB_FLAG_PROFILE;
	if (opt.grid_nn) p.nn_method = ICPB_NN_GRID;
	std::vector<float> err((size_t)p.max_iter + 1, 0.f);
	icpb_result res;
	if ((rc = icpb_run(ctx, &p, err.data(), &res)) != ICPB_OK) { printf("Error in the ICP loop: %s\n", icpb_last_error(ctx)); return -1; }
	print_report(err, res, !plane);
	if (opt.report) print_extra_report(res, npts, npts);
	return 0;
}

} // namespace dataset
