// CPU harness for csrc/grid_tree.cuh: builds the uniform grid and its occupancy pyramid on the host exactly as
// grid_nn.cu builds them on the device, runs the SAME traversal code (host instantiation) for every source and
// compares index and distance bits with a literal brute-force scan (reference chain via fmaf, strict `<` in ascending
// order = lowest index on ties, sentinel). No GPU needed:
//     nvcc -O2 -std=c++17 -Xcompiler -ffp-contract=off -I fast-point-cloud-registration-with-gpus_b200/csrc -I include \
//          tools/grid_tree_host_test.cu -o /tmp/grid_tree_host_test && /tmp/grid_tree_host_test
#include "grid_tree.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
using namespace icpb;

struct HostGrid {
	GridGeom g; GridPyramid py;
	std::vector<int> cell_start; std::vector<float4> sorted4; std::vector<unsigned char> occ;
};

static HostGrid build(const std::vector<float>& Q)
{
	const int m = (int)Q.size() / 3;
	float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
	for (int j = 0; j < m; j++) for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], Q[3 * j + k]); hi[k] = fmaxf(hi[k], Q[3 * j + k]); }
	HostGrid G;
	G.g = compute_grid_geom(lo, hi, m);
	const GridGeom& g = G.g;
	const size_t ncell = (size_t)g.nx * g.ny * g.nz;
	std::vector<int> cell_of(m), cnt(ncell + 1, 0);
	for (int j = 0; j < m; j++) {
		int cx = std::min(std::max(gt_cell_coord(Q[3 * j], g.ox, g.inv_h), 0), g.nx - 1);
		int cy = std::min(std::max(gt_cell_coord(Q[3 * j + 1], g.oy, g.inv_h), 0), g.ny - 1);
		int cz = std::min(std::max(gt_cell_coord(Q[3 * j + 2], g.oz, g.inv_h), 0), g.nz - 1);
		cell_of[j] = cx + g.nx * (cy + g.ny * cz);
		cnt[cell_of[j]]++;
	}
	G.cell_start.assign(ncell + 1, 0);
	for (size_t c = 0; c < ncell; c++) G.cell_start[c + 1] = G.cell_start[c] + cnt[c];
	std::vector<int> fill(ncell, 0);
	G.sorted4.resize(m);
	for (int j = m - 1; j >= 0; j--) {            // reversed on purpose: the order inside a cell must not matter
		const int c = cell_of[j];
		float4 q = make_float4(Q[3 * j], Q[3 * j + 1], Q[3 * j + 2], 0.f);
		memcpy(&q.w, &j, 4);
		G.sorted4[G.cell_start[c] + fill[c]++] = q;
	}
	const long long total = pyramid_layout(g, G.py);
	G.occ.assign((size_t)total, 0);
	for (size_t c = 0; c < ncell; c++) G.occ[c] = G.cell_start[c + 1] > G.cell_start[c];
	for (int L = 1; L < G.py.levels; L++) {
		const int nxc = G.py.nx[L - 1], nyc = G.py.ny[L - 1], nzc = G.py.nz[L - 1];
		for (int z = 0; z < G.py.nz[L]; z++) for (int y = 0; y < G.py.ny[L]; y++) for (int x = 0; x < G.py.nx[L]; x++) {
			unsigned char o = 0;
			for (int i = 0; i < 8; i++) {
				const int cx = 2 * x + (i & 1), cy = 2 * y + ((i >> 1) & 1), cz = 2 * z + ((i >> 2) & 1);
				if (cx < nxc && cy < nyc && cz < nzc && G.occ[(size_t)G.py.off[L - 1] + cx + (size_t)nxc * (cy + (size_t)nyc * cz)]) o |= (unsigned char)(1u << i);
			}
			G.occ[(size_t)G.py.off[L] + x + (size_t)G.py.nx[L] * (y + (size_t)G.py.ny[L] * z)] = o;
		}
	}
	G.py.occ = G.occ.data();
	return G;
}

template <int MODE> static u64 brute(const float* p, const std::vector<float>& Q, float sentinel)
{
	const int m = (int)Q.size() / 3;
	float best = sentinel; int bj = -1;
	for (int j = 0; j < m; j++) {
		float d = gt_chain(p[0], p[1], p[2], Q[3 * j], Q[3 * j + 1], Q[3 * j + 2]);
		if (MODE == ICPB_DIST_SQRT) d = sqrtf(d);
		if (d < best) { best = d; bj = j; }
	}
	if (bj < 0) return KEY_UNMATCHED;
	unsigned db; memcpy(&db, &best, 4);
	return ((u64)db << 32) | (u64)(unsigned)bj;
}

static float sqrt_domain_threshold(float sentinel)
{
	float y = sentinel * sentinel;
	if (std::isinf(y)) return y;
	while (sqrtf(y) < sentinel) y = nextafterf(y, INFINITY);
	while (y > 0.0f && sqrtf(nextafterf(y, 0.0f)) >= sentinel) y = nextafterf(y, 0.0f);
	return y;
}

static int failures = 0;
template <int MODE> static void check(const char* name, const std::vector<float>& P, const std::vector<float>& Q, float sentinel, int warm)
{
	HostGrid G = build(Q);
	const int n = (int)P.size() / 3;
	const float thr0 = MODE == ICPB_DIST_SQRT ? sqrt_domain_threshold(sentinel) : sentinel;
	unsigned long long seen = 0, nodes = 0; int bad = 0;
	for (int i = 0; i < n; i++) {
		const u64 want = brute<MODE>(&P[3 * i], Q, sentinel);
		u64 start = KEY_UNMATCHED;
		if (warm) {                                   // warm start from a real candidate: 1 = arbitrary (index i mod m); 2 = the target next to the
			                                          // answer (a tight bound: the near-field enumeration must still find the answer); 3 = the
			                                          // HIGHEST-indexed target as close as the answer (a tie: equality must not prune the lower index)
			int j = i % (int)(Q.size() / 3);
			if (warm >= 2 && want != KEY_UNMATCHED) {
				const int wj = (int)(unsigned)(want & 0xffffffffull), mq = (int)(Q.size() / 3);
				j = (wj + 1) % mq;
				if (warm == 3) {
					unsigned wb = (unsigned)(want >> 32);
					for (int t = mq - 1; t > wj; t--) {
						float dt = gt_chain(P[3 * i], P[3 * i + 1], P[3 * i + 2], Q[3 * t], Q[3 * t + 1], Q[3 * t + 2]);
						if (MODE == ICPB_DIST_SQRT) dt = sqrtf(dt);
						unsigned tb; memcpy(&tb, &dt, 4);
						if (tb == wb) { j = t; break; }
					}
				}
			}
			float d = gt_chain(P[3 * i], P[3 * i + 1], P[3 * i + 2], Q[3 * j], Q[3 * j + 1], Q[3 * j + 2]);
			if (d < thr0) { if (MODE == ICPB_DIST_SQRT) d = sqrtf(d); unsigned db; memcpy(&db, &d, 4); start = ((u64)db << 32) | (u64)(unsigned)j; }
		}
		const u64 got = grid_tree_nn<MODE>(P[3 * i], P[3 * i + 1], P[3 * i + 2], G.g, G.py, G.cell_start.data(), G.sorted4.data(), thr0, start, &seen, &nodes);
		if (got != want) { if (bad < 3) printf("  MISMATCH %s source %d: got %016llx want %016llx\n", name, i, (unsigned long long)got, (unsigned long long)want); bad++; }
	}
	printf("%-44s mode %d %s  grid %dx%dx%d levels %d  %6.1f points + %6.1f nodes per source (of %zu points)  %s\n", name, MODE, warm == 0 ? "cold" : warm == 1 ? "warm" : warm == 2 ? "near" : "tie ",
	       G.g.nx, G.g.ny, G.g.nz, G.py.levels, (double)seen / n, (double)nodes / n, Q.size() / 3, bad ? "FAIL" : "ok");
	failures += bad;
}

// knn + minimum of the reference (src/ICP_point_to_plane.cu:30-70), literally: all distances, then k1 argmin passes with
// strict `<` from 10000.0, each winner invalidated with 10000.0; "nothing found" returns index 0.
template <int MODE> static void check_knn(const char* name, const std::vector<float>& Q, int k1)
{
	HostGrid G = build(Q);
	const int m = (int)Q.size() / 3;
	int bad = 0;
	std::vector<float> d(m);
	for (int i = 0; i < m; i++) {
		for (int j = 0; j < m; j++) { float v = gt_chain(Q[3 * i], Q[3 * i + 1], Q[3 * i + 2], Q[3 * j], Q[3 * j + 1], Q[3 * j + 2]); d[j] = MODE == ICPB_DIST_SQRT ? sqrtf(v) : v; }
		int want[GT_KMAX];
		for (int r = 0; r < k1; r++) {
			float mn = 10000.0f; int b = 0;
			for (int j = 0; j < m; j++) if (d[j] < mn) { mn = d[j]; b = j; }
			want[r] = b; d[b] = 10000.0f;
		}
		float kd[GT_KMAX]; int ki[GT_KMAX];
		grid_tree_knn<MODE>(Q[3 * i], Q[3 * i + 1], Q[3 * i + 2], k1, G.g, G.py, G.cell_start.data(), G.sorted4.data(), kd, ki);
		for (int r = 0; r < k1; r++) {
			const int got = ki[r] >= 0 ? ki[r] : 0;
			if (got != want[r]) { if (bad < 3) printf("  KNN MISMATCH %s query %d rank %d: got %d want %d\n", name, i, r, got, want[r]); bad++; break; }
		}
	}
	printf("%-44s mode %d k+1=%d  grid %dx%dx%d  %d queries  %s\n", name, MODE, k1, G.g.nx, G.g.ny, G.g.nz, m, bad ? "FAIL" : "ok");
	failures += bad;
}

static void saddle(int W, std::vector<float>& D, std::vector<float>& M)
{
	D.resize(3 * (size_t)W * W); M.resize(D.size());
	const float r[9] = { 0.9788f, 0.0089f, 0.2045f, -0.0490f, 0.9800f, 0.1922f, -0.1987f, -0.1980f, 0.9599f };   // ~ the reference pose
	for (int i = 0; i < W * W; i++) {
		const float x = -2.0f + 4.0f * (float)(i / W) / (float)(W - 1), y = -2.0f + 4.0f * (float)(i % W) / (float)(W - 1), z = x * x - y * y;
		D[3 * i] = x; D[3 * i + 1] = y; D[3 * i + 2] = z;
		for (int k = 0; k < 3; k++) M[3 * i + k] = r[k] * x + r[k + 3] * y + r[k + 6] * z + (k == 0 ? 0.8f : k == 1 ? -0.3f : 0.2f);
	}
}

int main(int argc, char** argv)
{
	std::mt19937 rng(5);
	std::vector<float> D, M;
	if (argc > 1 && !strcmp(argv[1], "big")) {       // the bench size: 1M targets, a sample of sources at the initial pose and near convergence
		saddle(1000, D, M);
		std::vector<float> far, nearp;
		for (int i = 0; i < 1000000; i += 997) { for (int k = 0; k < 3; k++) { far.push_back(D[3 * i + k]); nearp.push_back(M[3 * i + k] + (k == 2 ? 2e-3f : 0.0f)); } }
		check<0>("1M saddle, initial pose (far field)", far, M, 100000.0f, false);
		check<0>("1M saddle, 2e-3 off the surface", nearp, M, 100000.0f, false);
		return failures ? 1 : 0;
	}
	saddle(120, D, M);
	std::vector<float> Dsub(D.begin(), D.begin() + 3 * 3000);
	for (int warm = 0; warm < 2; warm++) {
		check<0>("saddle, initial pose (far field)", Dsub, M, 100000.0f, warm);
		check<1>("saddle, initial pose (far field)", Dsub, M, 100000.0f, warm);
	}
	// near field: the source is the target itself, perturbed by a few ulps and by 1e-3
	std::vector<float> Pn(M.begin(), M.begin() + 3 * 3000);
	for (size_t k = 0; k < Pn.size(); k++) Pn[k] = (k % 3 == 0) ? nextafterf(Pn[k], INFINITY) : Pn[k] + ((k % 7 == 0) ? 1e-3f : 0.0f);
	check<0>("saddle, converged (near field)", Pn, M, 100000.0f, false);
	check<1>("saddle, converged (near field)", Pn, M, 100000.0f, true);
	// lattice: massive ties and duplicates, outliers far outside the box
	std::uniform_int_distribution<int> li(-10, 10), lp(-20, 20);
	std::vector<float> Q(3 * 6000), P(3 * 2000);
	for (auto& v : Q) v = 0.25f * (float)li(rng);
	for (auto& v : P) v = 0.125f * (float)lp(rng);
	for (int i = 0; i < 40; i++) for (int k = 0; k < 3; k++) P[3 * i + k] += 40.0f;
	for (int i = 40; i < 50; i++) for (int k = 0; k < 3; k++) P[3 * i + k] -= 1e3f;
	for (int warm = 0; warm < 2; warm++) { check<0>("lattice with ties, duplicates, outliers", P, Q, 100000.0f, warm); check<1>("lattice with ties, duplicates, outliers", P, Q, 100000.0f, warm); }
	for (int warm = 2; warm <= 3; warm++) { check<0>("lattice with ties, duplicates, outliers", P, Q, 100000.0f, warm); check<1>("lattice with ties, duplicates, outliers", P, Q, 100000.0f, warm); }
	check<0>("saddle, initial pose (far field)", Dsub, M, 100000.0f, 2);
	check<1>("saddle, converged (near field)", Pn, M, 100000.0f, 2);
	check<0>("saddle, converged (near field)", Pn, M, 100000.0f, 3);
	// sentinel rule
	std::normal_distribution<float> nd(0.f, 1.f);
	std::vector<float> Qn(3 * 2000), Pw(3 * 600);
	for (auto& v : Qn) v = nd(rng);
	for (auto& v : Pw) v = 3.0f * nd(rng);
	check<0>("random cloud, sentinel 0.05", Pw, Qn, 0.05f, false);
	check<1>("random cloud, sentinel 0.05", Pw, Qn, 0.05f, true);
	check<0>("random cloud", Pw, Qn, 100000.0f, true);
	check<0>("random cloud", Pw, Qn, 100000.0f, 2);
	check<1>("random cloud, sentinel 0.05", Pw, Qn, 0.05f, 2);
	// degenerate boxes
	std::vector<float> line(3 * 300, 0.f), pl;
	for (int j = 0; j < 300; j++) line[3 * j] = (float)j / 299.0f;
	for (int j = 0; j < 300; j += 7) { pl.push_back(line[3 * j] + 0.001f); pl.push_back(0.001f); pl.push_back(0.001f); }
	check<0>("collinear target", pl, line, 100000.0f, false);
	std::vector<float> one = { 1.f, 2.f, 3.f }, two = { 0.f, 0.f, 0.f, 5.f, 5.f, 5.f };
	check<0>("single target point", two, one, 100000.0f, false);
	// long collinear / thin / flat targets: one axis would need > 2^17 cells (368334x1x1 grid, 18 levels, cold MISMATCH);
	// a sample of sources along and beside the line, cold and warm; brute force over 100k targets for each
	{
		std::vector<float> L(3 * 100000, 0.f), PL;
		for (int j = 0; j < 100000; j++) { L[3 * j] = (float)j * 1e-3f; }
		for (int j = 0; j < 100000; j += 331) { PL.push_back(L[3 * j] + 4e-4f); PL.push_back((j % 2) ? 0.0f : 0.01f); PL.push_back((j % 3) ? 0.0f : -0.02f); }
		for (int warm = 0; warm < 3; warm++) { check<0>("100k collinear target", PL, L, 100000.0f, warm); check<1>("100k collinear target", PL, L, 100000.0f, warm); }
		std::uniform_real_distribution<float> u(0.f, 1.f);
		std::vector<float> T(3 * 100000), PT;
		for (int j = 0; j < 100000; j++) { T[3 * j] = 200.0f * u(rng); T[3 * j + 1] = 0.01f * u(rng); T[3 * j + 2] = 0.01f * u(rng); }      // thin rod
		for (int j = 0; j < 100000; j += 407) { PT.push_back(T[3 * j] + 1e-3f); PT.push_back(T[3 * j + 1] - 2e-3f); PT.push_back(0.02f * u(rng)); }
		check<0>("100k thin rod 200 x 0.01 x 0.01", PT, T, 100000.0f, false);
		check<1>("100k thin rod 200 x 0.01 x 0.01", PT, T, 100000.0f, true);
		std::vector<float> F(3 * 60000), PF;
		for (int j = 0; j < 60000; j++) { F[3 * j] = 3000.0f * u(rng); F[3 * j + 1] = 3000.0f * u(rng); F[3 * j + 2] = 7.0f; }                // flat, large extent
		for (int j = 0; j < 60000; j += 263) { PF.push_back(F[3 * j] + 0.5f); PF.push_back(F[3 * j + 1] - 0.25f); PF.push_back(7.0f + 3.0f * u(rng)); }
		check<0>("60k flat sheet 3000 x 3000 x 0", PF, F, 100000.0f, false);
		HostGrid G = build(L);
		if (G.g.nx > GP_MAX_DIM || G.py.nx[G.py.levels - 1] != 1 || G.py.ny[G.py.levels - 1] != 1 || G.py.nz[G.py.levels - 1] != 1) { printf("FAIL: pyramid of the collinear target has no single root (%d levels, top %dx%dx%d)\n", G.py.levels, G.py.nx[G.py.levels - 1], G.py.ny[G.py.levels - 1], G.py.nz[G.py.levels - 1]); failures++; }
	}
	// large offsets and scales
	std::vector<float> Qo(Qn), Po(Pw);
	for (auto& v : Qo) v = v * 30.0f - 5e4f;
	for (auto& v : Po) v = v * 30.0f - 5e4f;
	check<0>("offset -5e4, scale 30", Po, Qo, 3e38f, false);
	check<1>("offset -5e4, scale 30", Po, Qo, 3e38f, true);
	// a cloud a million units from the origin with unit spread (world / UTM coordinates): ulp(1e6) = 0.06 > the cell size
	std::vector<float> Qu(Qn), Pu(Pw);
	for (size_t k = 0; k < Qu.size(); k++) Qu[k] = Qu[k] + (k % 3 == 0 ? 1e6f : k % 3 == 1 ? -2.5e5f : 3e3f);
	for (size_t k = 0; k < Pu.size(); k++) Pu[k] = Pu[k] + (k % 3 == 0 ? 1e6f : k % 3 == 1 ? -2.5e5f : 3e3f);
	check<0>("unit cloud at (1e6,-2.5e5,3e3)", Pu, Qu, 3e38f, false);
	check<0>("unit cloud at (1e6,-2.5e5,3e3)", Pu, Qu, 3e38f, 2);
	check<1>("unit cloud at (1e6,-2.5e5,3e3)", Pu, Qu, 3e38f, true);
	// sources exactly on cell faces and corners of the grid's frame
	{
		HostGrid G = build(Qn);
		std::vector<float> Pf;
		for (int i = 0; i < 400; i++) { Pf.push_back(G.g.ox + (float)(i % 17) * G.g.h); Pf.push_back(G.g.oy + (float)((i / 3) % 15) * G.g.h); Pf.push_back(G.g.oz + (float)((i / 7) % 18) * G.g.h); }
		check<0>("sources on cell faces / corners", Pf, Qn, 100000.0f, false);
		check<0>("sources on cell faces / corners", Pf, Qn, 100000.0f, 2);
		check<1>("sources on cell faces / corners", Pf, Qn, 100000.0f, true);
	}
	// non-finite sources
	std::vector<float> Pbad(Pw.begin(), Pw.begin() + 30); Pbad[0] = NAN; Pbad[4] = INFINITY; Pbad[8] = -INFINITY;
	check<0>("NaN / inf sources", Pbad, Qn, 100000.0f, false);
	// ---- k nearest neighbours (K5) ----
	{
		std::vector<float> Dk, Mk;
		saddle(48, Dk, Mk);
		check_knn<1>("knn: saddle target", Mk, 5);
		check_knn<0>("knn: saddle target", Mk, 5);
		check_knn<1>("knn: lattice with duplicates", std::vector<float>(Q.begin(), Q.begin() + 3 * 2500), 5);
		check_knn<0>("knn: lattice with duplicates", std::vector<float>(Q.begin(), Q.begin() + 3 * 2500), 7);
		check_knn<1>("knn: random cloud", Qn, 5);
		check_knn<1>("knn: collinear", line, 5);
		std::vector<float> tiny(Qn.begin(), Qn.begin() + 9);
		check_knn<1>("knn: 3 points, k+1 = 5", tiny, 5);
		std::vector<float> wide(Qn.begin(), Qn.begin() + 3 * 400);
		for (auto& v : wide) v *= 6000.0f;                                    // most pairs farther than 10000: the cut-off matters
		check_knn<1>("knn: spread beyond the 10000 cut-off", wide, 5);
		check_knn<0>("knn: spread beyond the 10000 cut-off", wide, 5);
		check_knn<1>("knn: unit cloud at 1e6", std::vector<float>(Qu.begin(), Qu.begin() + 3 * 1200), 5);
		// sqrt merging: a ring of points whose squared distances to the centre differ by ulps
		std::vector<float> ring = { 0.f, 0.f, 0.f };
		const float e = sqrtf(nextafterf(1.0f, 2.0f) - 1.0f);
		for (int k = 0; k < 12; k++) { const float a = 6.2831853f * (float)k / 12.0f; ring.push_back(cosf(a)); ring.push_back(sinf(a)); ring.push_back((k % 2) ? e : 0.0f); }
		check_knn<1>("knn: sqrt-merged ring", ring, 5);
	}
	printf(failures ? "FAILED: %d mismatches\n" : "all exact\n", failures);
	return failures ? 1 : 0;
}
