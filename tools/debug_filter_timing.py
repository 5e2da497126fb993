import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
D, M = icp_synth.p2p_clouds(317, 100000)
ctx = ib.Context(0)
for nn, name in ((ib.NN_BRUTE, "filter"), (ib.NN_BRUTE_DIRECT, "direct")):
    ctx.set_target(M); ctx.set_source(D)
    s0 = ctx.filter_stats()
    e, r = ctx.run(ib.default_params(max_iter=64, nn_method=nn))
    s1 = ctx.filter_stats()
    print(name, "single run: iters", r.iterations_run, "match_ms %.2f elapsed %.2f" % (r.match_ms, r.elapsed_ms), "exact frac %.4f" % ((s1["subtile_exact"]-s0["subtile_exact"])/max(1,s1["subtile_tests"]-s0["subtile_tests"])))
    ctx.set_source(D)
    tot = 0; per = []
    for it in range(r.iterations_run):
        e, r1 = ctx.run(ib.default_params(max_iter=1, stop_early=0, nn_method=nn)); tot += r1.match_ms; per.append(round(r1.match_ms, 2))
    print(name, "stepwise: match_ms %.2f" % tot, per[:6])
D, M = icp_synth.p2p_clouds(1000)
p = ib.default_params()
for k in range(3):
    t0 = time.perf_counter(); ctx.set_target(M); t1 = time.perf_counter(); ctx.set_source(D); t2 = time.perf_counter()
    e, r = ctx.run(ib.default_params(max_iter=1, stop_early=0)); t3 = time.perf_counter()
    print("1M host path: set_target %.1f ms set_source %.1f ms run %.1f ms (match_ms %.1f)" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, r.match_ms))
    t0 = time.perf_counter(); ctx.iterate_host(p, D, M); print("   iterate_host %.1f ms" % ((time.perf_counter()-t0)*1e3))
