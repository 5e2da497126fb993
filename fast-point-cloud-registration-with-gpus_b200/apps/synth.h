// synth.h — the reference's synthetic saddle z = x^2 - y^2 and its rigidly moved copy, plus the tiny
// command-line helper shared by the drop-in executables. Host-only; compile with -ffp-contract=off so
// the clouds are bit-identical to what the reference binaries generate (no FMA on the host there).
//
// Restates (not copies) src/ICP_point_to_point.cu:103-190 (= src/ICP_point_to_plane.cu:258-347) and
// src/ICP_standard.cu:160-263.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#define MY_LIB_NO_IMPL
#include "my_lib.h"

namespace synth {

struct Clouds {
	std::vector<float> D;   // data / source, AoS xyz
	std::vector<float> M;   // model / target
	int n = 0;
};

// Grid over [-2,2]^2: point i has x = lin[i / W] (slow) and y = lin[i % W] (fast); `npts` <= W*W keeps
// the first npts points in generation order. `double_linspace` selects ICP_standard's spelling of the
// lin_space expression (a double sum rounded to float, src/ICP_standard.cu:164) instead of the all-float
// one (src/ICP_point_to_point.cu:109); both round identically, it is kept for fidelity.
inline void make_source(int W, int npts, bool double_linspace, std::vector<float>& D)
{
	const float lenght = (float)(2.0 - (-2.0));
	std::vector<float> lin((size_t)W);
	for (int i = 0; i < W; i++) {
		const float step = ((float)i * lenght) / ((float)W - 1.0f);
		lin[(size_t)i] = double_linspace ? (float)(-2.0 + (double)step) : ((float)-2.0 + step);
	}
	D.resize(3 * (size_t)npts);
	for (int i = 0; i < npts; i++) {
		const float x = lin[(size_t)(i / W)], y = lin[(size_t)(i % W)];
		D[3 * (size_t)i + 0] = x;
		D[3 * (size_t)i + 1] = y;
		D[3 * (size_t)i + 2] = (float)(pow((double)x, 2) - pow((double)y, 2));   // pow(float,int) promotes to double
	}
}

// Column-major rotation built from Euler angles the way src/ICP_point_to_point.cu:167-172 does.
inline void euler_rotation(const float ri[3], float r[9])
{
	const float cx = (float)cos(ri[0]), cy = (float)cos(ri[1]), cz = (float)cos(ri[2]);
	const float sx = (float)sin(ri[0]), sy = (float)sin(ri[1]), sz = (float)sin(ri[2]);
	r[0] = cy * cz;   r[1] = (cz * sx * sy) + (cx * sz);   r[2] = -(cx * cz * sy) + (sx * sz);
	r[3] = -cy * sz;  r[4] = (cx * cz) - (sx * sy * sz);   r[5] = (cx * sy * sz) + (cz * sx);
	r[6] = sy;        r[7] = -cy * sx;                     r[8] = cx * cy;
}

// M = r * D + t with the reference's own naive product (my_lib's SmatrixMul).
inline void move_rigid(std::vector<float>& D, int npts, float r[9], const float t[3], std::vector<float>& M)
{
	M.resize(3 * (size_t)npts);
	SmatrixMul(r, D.data(), M.data(), 3, npts, 3);
	for (int i = 0; i < npts; i++)
		for (int j = 0; j < 3; j++) M[(size_t)j + 3 * (size_t)i] += t[j];
}

// Defaults of ICP_point_to_point.cu / ICP_point_to_plane.cu: t = (0.8,-0.3,0.2), r = (0.2,-0.2,0.05) rad.
inline Clouds point_to_point_clouds(int W, int npts)
{
	Clouds c; c.n = npts;
	float ti[3] = { 0.8f, -0.3f, 0.2f }, ri[3] = { 0.2f, -0.2f, 0.05f }, r[9];
	make_source(W, npts, false, c.D);
	euler_rotation(ri, r);
	move_rigid(c.D, npts, r, ti, c.M);
	return c;
}

// ICP_standard.cu: hard-coded rotation (:247-249) and t = (1,-0.3,0.2).
inline Clouds standard_clouds(int W)
{
	Clouds c; c.n = W * W;
	float r[9] = { 0.876485812f, -0.37591464f, 0.300767018f,
	               -0.04386084f, 0.559789799f, 0.827473024f,
	               -0.47942553f, -0.73846026f, 0.474159881f };
	float ti[3] = { 1.0f, -0.3f, 0.2f };
	make_source(W, c.n, true, c.D);
	move_rigid(c.D, c.n, r, ti, c.M);
	return c;
}

// ---- optional command line (the reference takes none; with no arguments behaviour is identical) ----
struct Options {
	int width = 0, n = 0, max_iter = 0, gpus = 1, grid_nn = 0, report = 0, sync_every = 0;
	double tol = -1;
};
inline bool parse(int argc, char** argv, Options& o)
{
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto val = [&](const char* name) -> const char* {
			size_t L = strlen(name);
			if (a.compare(0, L, name) == 0 && a.size() > L && a[L] == '=') return argv[i] + L + 1;
			if (a == name && i + 1 < argc) return argv[++i];
			return nullptr;
		};
		const char* v;
		if ((v = val("--width"))) o.width = atoi(v);
		else if ((v = val("--n"))) o.n = atoi(v);
		else if ((v = val("--max-iter"))) o.max_iter = atoi(v);
		else if ((v = val("--tol"))) o.tol = atof(v);
		else if ((v = val("--gpus"))) o.gpus = atoi(v);
		else if ((v = val("--sync-every"))) o.sync_every = atoi(v);
		else if ((v = val("--nn"))) o.grid_nn = (strcmp(v, "grid") == 0);
		else if (a == "--report") o.report = 1;
		else {
			fprintf(stderr, "usage: %s [--width W] [--n N] [--max-iter K] [--tol T] [--nn brute|grid] [--sync-every K] [--gpus G] [--report]\n", argv[0]);
			return false;
		}
	}
	return true;
}

} // namespace synth
