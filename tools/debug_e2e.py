import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
D, M = icp_synth.p2p_clouds(1000)
ctx = ib.Context(0)
p = ib.default_params()
cur = D
for k in range(6):
    s0 = ctx.filter_stats()
    t0 = time.perf_counter(); ctx.set_target(M); t1 = time.perf_counter(); ctx.set_source(cur); t2 = time.perf_counter()
    e, r = ctx.run(ib.default_params(max_iter=1, stop_early=0)); t3 = time.perf_counter()
    cur = ctx.get_source(); t4 = time.perf_counter()
    s1 = ctx.filter_stats()
    print("step %d: set_target %.1f set_source %.1f run %.1f (match %.1f) get_source %.1f ms; exact frac %.4f rms %.4f" % (k, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, r.match_ms, (t4-t3)*1e3,
          (s1["subtile_exact"]-s0["subtile_exact"])/max(1.0, s1["subtile_tests"]-s0["subtile_tests"]), e[1]), flush=True)
print("now without re-uploading the target")
for k in range(4):
    s0 = ctx.filter_stats()
    t2 = time.perf_counter(); ctx.set_source(cur); e, r = ctx.run(ib.default_params(max_iter=1, stop_early=0)); t3 = time.perf_counter()
    cur = ctx.get_source()
    s1 = ctx.filter_stats()
    print("step %d: set_source+run %.1f (match %.1f); exact frac %.4f" % (k, (t3-t2)*1e3, r.match_ms, (s1["subtile_exact"]-s0["subtile_exact"])/max(1.0, s1["subtile_tests"]-s0["subtile_tests"])), flush=True)
