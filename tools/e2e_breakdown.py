"""Where does a host-driven step spend its wall-clock time? Times each C-ABI call of icpb_iterate_host's sequence
separately (set_target, set_source, run(1 iteration), get_correspondences, get_source) at the bench size, and
prints the filter's exact-pass rate per step for the planar / full bound."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib      # noqa: E402
import icp_synth           # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
RETARGET = (sys.argv[3] != "0") if len(sys.argv) > 3 else True     # upload the target again at every step (as icpb_iterate_host does)
D, M = icp_synth.p2p_clouds(W)
ctx = ib.Context(0)
cur = D
p = ib.default_params(max_iter=1, stop_early=0, sync_every=1)
for step in range(STEPS):
    t = [time.perf_counter()]
    if RETARGET or step == 0:
        ctx.set_target(M)
    t.append(time.perf_counter())
    ctx.set_source(cur); t.append(time.perf_counter())
    s0 = ctx.filter_stats()
    t[-1] = time.perf_counter()
    err, res = ctx.run(p); t.append(time.perf_counter())
    idx = ctx.correspondences(); t.append(time.perf_counter())
    cur = ctx.get_source(); t.append(time.perf_counter())
    s1 = ctx.filter_stats(); cfg = ctx.filter_config()
    d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
    frac = (s1["subtile_exact"] - s0["subtile_exact"]) / max(1.0, s1["subtile_tests"] - s0["subtile_tests"])
    print("step %d: set_target %.2f set_source %.2f run %.2f (device %.2f, match %.2f) get_idx %.2f get_source %.2f ms | dims %d drop %d exact %.3f%%"
          % (step, d[0], d[1], d[2], res.elapsed_ms, res.match_ms, d[3], d[4], cfg["dims_last"], cfg["drop_axis"], 100 * frac))
ctx.close()
