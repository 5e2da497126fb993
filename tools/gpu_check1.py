# quick GPU bring-up check (scratch): matching parity + p2p run vs oracle
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import icp_b200 as ib
import oracle as orc
ctx = ib.Context(0)
print(ctx.device_info())
for W, mode in [(32, ib.DIST_SQ), (32, ib.DIST_SQRT), (32, ib.DIST_STD), (64, ib.DIST_SQ), (100, ib.DIST_SQRT)]:
    D, M = orc.synth_p2p(W)
    ctx.set_target(M); ctx.set_source(D)
    idx = ctx.match(mode)
    ref = orc.match(D, M, mode)
    print("W", W, "mode", mode, "mismatches", int((idx != ref).sum()), "of", len(ref))
# full run
for W in (32, 128):
    D, M = orc.synth_p2p(W)
    ctx.set_target(M); ctx.set_source(D)
    p = ib.default_params(max_iter=40)
    err, res = ctx.run(p)
    o = orc.icp_p2p(D, M, max_iter=40)
    print("W", W, "gpu iters", res.iterations, res.iterations_run, "oracle iters", o["iterations"], o["iterations_run"], "elapsed_ms", res.elapsed_ms, "match_ms", res.match_ms)
    print(" gpu err", np.array2string(err[:res.iterations + 2], precision=5))
    print(" orc err", np.array2string(o["errors"][:o["iterations"] + 2], precision=5))
    print(" R diff", np.abs(np.array(res.R[:]) - o["R"]).max(), "t diff", np.abs(np.array(res.t[:]) - o["t"]).max())
    print(" idx==i:", int((ctx.correspondences() == np.arange(W * W)).sum()), "/", W * W)
print("fp32 peak", ctx.fp32_peak_tflops())
for n in (16384, 100000):
    W = int(np.ceil(np.sqrt(n)))
    D, M = orc.synth_p2p(W, n)
    ctx.set_target(M); ctx.set_source(D)
    for cfg in range(6):
        os.environ["ICPB_K1_CFG"] = str(cfg)
        c2 = ib.Context(0); c2.set_target(M); c2.set_source(D)
        mean, mn = c2.time_match(ib.DIST_SQ, reps=5)
        print("n", n, "cfg", cfg, "match ms mean %.3f min %.3f  pairs/s %.3e" % (mean, mn, n * n / (mn * 1e-3)))
        c2.close()
