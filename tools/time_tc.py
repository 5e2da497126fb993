#!/usr/bin/env python
"""K1T (tensor-core filter) against K1F (FP32 filter) and K1 (direct): matching-kernel time over the first iterations of the
1M x 1M registration (cold pass, then warm passes), exact-pass fractions, identical correspondences."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import numpy as np
import icp_b200 as ib
import icp_synth

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
kind = os.environ.get("TC_CLOUD", "saddle")          # saddle (raster order) | volume | sphere (random order: the Morton-ordered tiles)
if kind == "saddle":
    D, M = icp_synth.p2p_clouds(W)
else:
    rng = np.random.default_rng(99)
    nf = W * W
    if kind == "volume":
        M = rng.uniform(-1.0, 1.0, size=(nf, 3)).astype(np.float32)
    else:
        u = rng.normal(size=(nf, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
        M = (u * (2.0 + 0.01 * rng.normal(size=(nf, 1)))).astype(np.float32)
    D = icp_synth.rigid_move(M, icp_synth.euler_matrix([0.02, -0.02, 0.005]), [0.02, -0.01, 0.015])
if os.environ.get("TC_SHARD"):                       # "r/w": the sources rank r of w would get from bench.py (blocks of 2048 dealt round-robin)
    parts = [int(v) for v in os.environ["TC_SHARD"].split("/")]
    r_, w_ = parts[0], parts[1]
    blk = parts[2] if len(parts) > 2 else 2048             # 0: contiguous shards
    if blk > 0:
        blocks = np.arange(D.shape[0]) // blk
        D = np.ascontiguousarray(D[blocks % w_ == r_])
    else:
        per = (D.shape[0] + w_ - 1) // w_
        D = np.ascontiguousarray(D[r_ * per:(r_ + 1) * per])
out = {}
ref_idx = None
variants = [int(v) for v in os.environ.get("TC_VARIANTS", "0,1,2,3,4").split(",")]
cases = [("k1t_var%d" % v, {"ICPB_K1_TC": "1", "ICPB_KT_VAR": str(v)}) for v in variants if v >= 0] + [("k1t_auto", {"ICPB_K1_TC": "1"})] + ([] if os.environ.get("SKIP_K1F") else [("k1f_fp32", {"ICPB_K1_TC": "0"})])
for name, env in cases:
    os.environ.pop("ICPB_K1_TC", None); os.environ.pop("ICPB_KT_VAR", None)
    os.environ.update(env)
    with ib.Context(0) as ctx:
        ctx.set_target(M); ctx.set_source(D)
        per_it = []
        for k in range(iters):
            e, r = ctx.run(ib.default_params(max_iter=1, stop_early=0))
            per_it.append(round(r.match_ms, 3))
        st = ctx.filter_stats()
        idx = ctx.correspondences()
        out[name] = {"match_ms_per_iteration": per_it, "pairs_per_sec_last": float(D.shape[0]) * M.shape[0] / (per_it[-1] * 1e-3),
                     "exact_fraction_overall": st["subtile_exact"] / max(st["subtile_tests"], 1), "final_rms": float(e[1]),
                     "tc_config": ctx.filter_tc_config()}
        if ref_idx is None:
            ref_idx = idx
        else:
            out["identical_correspondences"] = bool(np.array_equal(ref_idx, idx))
print(json.dumps(out, indent=1))
