// icp_batched — BASELINE.json config 5 as a plain C++ program: `--batch` independent registrations of `--n`-point
// synthetic clouds (default 4096 pairs of 2048 points), dealt to `--gpus` GPUs of this box (replicas: one host thread
// and one context per GPU, no communication), every pair registered to convergence inside one persistent kernel
// (icpb_run_batched). Poses come from the counter-based generator of SURVEY.md 8(d): splitmix64, seed 20240, stream b:
// r = 0.2 u, u in U(-1,1)^3; t = (0.8,-0.3,0.2) * (0.5 + 0.5 v), v in U(0,1)^3. The reference has no such program (its
// nearest relative is src/tests/centroid.cu's 2048-point block reduction); the per-pair loop is src/ICP_point_to_point.cu:295-423.
#include <chrono>
#include <cstdint>
#include <thread>
#include "synth.h"
#include "icp_b200.h"

static uint64_t splitmix64(uint64_t& state)
{
	state += 0x9E3779B97F4A7C15ull;
	uint64_t z = state;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

int main(int argc, char** argv)
{
	int batch = 4096, n = 2048, width = 46, gpus = 1, max_iter = 40, reps = 3;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return (i + 1 < argc) ? atoi(argv[++i]) : 0; };
		if (a == "--batch") batch = next();
		else if (a == "--n") n = next();
		else if (a == "--width") width = next();
		else if (a == "--gpus") gpus = next();
		else if (a == "--max-iter") max_iter = next();
		else if (a == "--reps") reps = next();
		else { fprintf(stderr, "usage: %s [--batch B] [--n N] [--width W] [--gpus G] [--max-iter K] [--reps R]\n", argv[0]); return 2; }
	}
	if (batch < 1 || n < 1 || n > width * width || gpus < 1 || reps < 1) { fprintf(stderr, "bad arguments\n"); return 2; }

	// pinned host clouds (streamed upload); the source is the same saddle patch for every pair, each target its moved copy
	float *S = nullptr, *T = nullptr;
	const unsigned long long bytes = sizeof(float) * 3ull * (unsigned long long)batch * n;
	if (icpb_host_alloc((void**)&S, bytes) != ICPB_OK || icpb_host_alloc((void**)&T, bytes) != ICPB_OK) { printf("Error allocating pinned host memory\n"); return -1; }
	std::vector<float> D, M;
	synth::make_source(width, n, false, D);
	std::vector<float> r_true(9 * (size_t)batch), t_true(3 * (size_t)batch);
	for (int b = 0; b < batch; b++) {
		uint64_t st = (20240ull << 32) | (uint64_t)b;
		double d[6];
		for (int k = 0; k < 6; k++) d[k] = (double)(splitmix64(st) >> 11) * (1.0 / 9007199254740992.0);
		float ri[3], ti[3], r[9];
		const double base[3] = { 0.8, -0.3, 0.2 };
		for (int k = 0; k < 3; k++) { ri[k] = (float)(0.2 * (d[k] * 2.0 - 1.0)); ti[k] = (float)(base[k] * (0.5 + 0.5 * d[3 + k])); }
		synth::euler_rotation(ri, r);
		synth::move_rigid(D, n, r, ti, M);
		memcpy(S + 3 * (size_t)b * n, D.data(), sizeof(float) * 3 * (size_t)n);
		memcpy(T + 3 * (size_t)b * n, M.data(), sizeof(float) * 3 * (size_t)n);
		memcpy(&r_true[9 * (size_t)b], r, sizeof r); memcpy(&t_true[3 * (size_t)b], ti, sizeof ti);
	}

	std::vector<icpb_ctx*> ctx((size_t)gpus, nullptr);
	for (int g = 0; g < gpus; g++) {
		const int rc = icpb_create(&ctx[(size_t)g], g);
		if (rc != ICPB_OK) { printf("Error creating the ICP context on GPU %d: %s\n", g, icpb_status_string(rc)); return -1; }
	}
	icpb_params p;
	icpb_default_params(&p);
	p.max_iter = max_iter;
	std::vector<float> errors((size_t)batch * (max_iter + 1));
	std::vector<int> iters((size_t)batch);
	std::vector<double> R(9 * (size_t)batch), t(3 * (size_t)batch);
	std::vector<float> kernel_ms((size_t)gpus, 0.f);
	std::vector<int> rcs((size_t)gpus, ICPB_OK);
	double best_wall = 1e30;
	for (int rep = 0; rep < reps; rep++) {
		const auto t0 = std::chrono::steady_clock::now();
		std::vector<std::thread> th;
		for (int g = 0; g < gpus; g++)
			th.emplace_back([&, g] {
				const int lo = (int)((long long)batch * g / gpus), hi = (int)((long long)batch * (g + 1) / gpus);
				if (hi <= lo) return;
				rcs[(size_t)g] = icpb_run_batched(ctx[(size_t)g], &p, hi - lo, S + 3 * (size_t)lo * n, n, T + 3 * (size_t)lo * n, n,
				                                  errors.data() + (size_t)lo * (max_iter + 1), iters.data() + lo, R.data() + 9 * (size_t)lo, t.data() + 3 * (size_t)lo,
				                                  &kernel_ms[(size_t)g]);
			});
		for (auto& x : th) x.join();
		const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		for (int g = 0; g < gpus; g++) if (rcs[(size_t)g] != ICPB_OK) { printf("Error in the batched ICP on GPU %d: %s\n", g, icpb_last_error(ctx[(size_t)g])); return -1; }
		if (wall < best_wall) best_wall = wall;
	}
	double max_pose = 0.0, pairs = 0.0; long long its = 0; int worst = 0;
	for (int b = 0; b < batch; b++) {
		for (int k = 0; k < 9; k++) max_pose = fmax(max_pose, fabs(R[9 * (size_t)b + k] - (double)r_true[9 * (size_t)b + k]));
		for (int k = 0; k < 3; k++) max_pose = fmax(max_pose, fabs(t[3 * (size_t)b + k] - (double)t_true[3 * (size_t)b + k]));
		its += iters[(size_t)b] + 1; if (iters[(size_t)b] > worst) worst = iters[(size_t)b];
		pairs += (double)(iters[(size_t)b] + 1) * n * n;
	}
	float kms = 0.f; for (float v : kernel_ms) kms = fmaxf(kms, v);
	printf("Batched ICP: %d pairs of %d points on %d GPU(s)\n", batch, n, gpus);
	printf("ICP converged successfully!\n\n");
	printf("Elapsed time: %f ms\n", best_wall);
	printf("[report] host buffers in, results out: %.3f ms wall (best of %d), slowest GPU's kernel %.3f ms; %.0f registrations/s, %.4e NN pairs/s, %.0f ICP iterations/s\n",
	       best_wall, reps, kms, batch / (best_wall * 1e-3), pairs / (best_wall * 1e-3), its / (best_wall * 1e-3));
	printf("[report] max |pose - generating pose| over all pairs: %.3e; iterations per pair: max %d, mean %.2f\n", max_pose, worst, (double)its / batch - 1.0);
	for (auto c : ctx) icpb_destroy(c);
	icpb_host_free(S); icpb_host_free(T);
	return max_pose < 2e-5 ? 0 : 1;
}
