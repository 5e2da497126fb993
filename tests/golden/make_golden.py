"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference.

Runs on a GPU box (B200) where oracle/_ref/ was built beforehand from /root/reference by
oracle/Makefile (the reference tree itself is not needed at run time):

    python tests/golden/make_golden.py [outdir]          # default: gpurun_out/golden

  * oracle/_ref/ref_kernels run std|p2p|p2l  -> stdout of the three reference programs
  * oracle/_ref/ref_kernels match|ryt|knn|normalsA|cxb -> outputs of the reference's own device kernels
    on inputs produced by the oracle's generators (stored in the fixture too)
  * oracle/_ref/icp_cpu -> stdout of src/ICP_CPU.c (CPU; also reproducible without a GPU)
The fixtures are small (.npz, a few hundred KB in total) and are committed; tests compare both the CPU
oracle (-m "not gpu") and the CUDA engine (-m gpu) against them.
"""
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_kernels")


def run_ref(args, **kw):
    return subprocess.run([REF] + [str(a) for a in args], check=True, capture_output=True, text=True, **kw)


def partly_registered(D, M, iters):
    """Source after `iters` oracle ICP iterations: gives non-trivial, tie-rich matching inputs."""
    if iters == 0:
        return D.copy()
    return orc.icp_p2p(D, M, max_iter=iters, stop_early=False)["P"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    tmp = tempfile.mkdtemp()

    # 1. whole reference programs
    for name in ("std", "p2p", "p2l"):
        r = run_ref(["run", name])
        open(os.path.join(out, "ref_%s_stdout.txt" % name), "w").write(r.stdout)
        print(name, "stdout lines:", len(r.stdout.splitlines()))

    def f(name):
        return os.path.join(tmp, name)

    # 2. Matching kernels, three variants, several stages of a registration
    cases = {}
    for W, iters in ((32, 0), (32, 5), (32, 12), (64, 3)):
        D, M = orc.synth_p2p(W)
        P = partly_registered(D, M, iters)
        n = W * W
        P.tofile(f("P.bin")); M.tofile(f("Q.bin"))
        for which in ("p2p", "p2l", "std"):
            if which == "std" and n > 1024:
                continue
            run_ref(["match", which, n, n, f("P.bin"), f("Q.bin"), f("idx.bin")])
            cases["match_%s_W%d_it%d" % (which, W, iters)] = np.fromfile(f("idx.bin"), np.int32)
        cases["P_W%d_it%d" % (W, iters)] = P
        cases["Q_W%d" % W] = M
    # ICP_standard's own clouds (harder pose) through its own Matching kernel
    D, M = orc.synth_standard(32)
    D.tofile(f("P.bin")); M.tofile(f("Q.bin"))
    run_ref(["match", "std", 1024, 1024, f("P.bin"), f("Q.bin"), f("idx.bin")])
    cases["match_std_standard_clouds"] = np.fromfile(f("idx.bin"), np.int32)
    cases["P_standard"] = D; cases["Q_standard"] = M
    # exact duplicates / ties: target = two copies of the same cloud, and a coarse lattice
    rng = np.random.default_rng(7)
    lat = rng.integers(-3, 4, size=(1024, 3)).astype(np.float32) * 0.5
    src = rng.integers(-3, 4, size=(512, 3)).astype(np.float32) * 0.5 + np.float32(0.25)
    src.tofile(f("P.bin")); lat.tofile(f("Q.bin"))
    for which in ("p2p", "p2l", "std"):
        run_ref(["match", which, 512, 1024, f("P.bin"), f("Q.bin"), f("idx.bin")])
        cases["match_%s_lattice" % which] = np.fromfile(f("idx.bin"), np.int32)
    cases["P_lattice"] = src; cases["Q_lattice"] = lat
    np.savez_compressed(os.path.join(out, "ref_matching.npz"), **cases)
    print("matching cases:", len(cases))

    # 3. RyT
    D, M = orc.synth_p2p(32)
    R = orc.euler_matrix([0.3, -0.1, 0.7]); T = np.array([0.125, -1.5, 0.3], np.float32)
    D.tofile(f("P.bin")); R.tofile(f("R.bin")); T.tofile(f("T.bin"))
    run_ref(["ryt", 1024, f("R.bin"), f("T.bin"), f("P.bin"), f("out.bin")])
    np.savez_compressed(os.path.join(out, "ref_ryt.npz"), P=D, R=R, T=T, out=np.fromfile(f("out.bin"), np.float32).reshape(-1, 3))

    # 4. knn + Normals covariance + Cxb on the point-to-plane program's clouds
    k = {}
    for W in (32, 64):
        D, M = orc.synth_p2p(W)
        m = W * W
        M.tofile(f("Q.bin"))
        run_ref(["normalsA", m, 4, f("Q.bin"), f("nbr.bin"), f("A.bin")])
        k["nbr_W%d" % W] = np.fromfile(f("nbr.bin"), np.int32).reshape(m, 5)
        k["A_W%d" % W] = np.fromfile(f("A.bin"), np.float32).reshape(m, 9)
        k["Q_W%d" % W] = M
    np.savez_compressed(os.path.join(out, "ref_knn_normals.npz"), **k)

    D, M = orc.synth_p2p(32)
    nbr = orc.knn(M, 5); nrm = orc.normals(M, nbr, 4)
    idx = orc.match(D, M, orc.MODE_SQRT)
    D.tofile(f("P.bin")); M.tofile(f("Q.bin")); idx.tofile(f("idx.bin")); nrm.tofile(f("nrm.bin"))
    run_ref(["cxb", 1024, 1024, f("P.bin"), f("Q.bin"), f("idx.bin"), f("nrm.bin"), f("C.bin"), f("b.bin")])
    np.savez_compressed(os.path.join(out, "ref_cxb.npz"), P=D, Q=M, idx=idx, normals=nrm,
                        C=np.fromfile(f("C.bin"), np.float32), b=np.fromfile(f("b.bin"), np.float32))
    print("done ->", out)


if __name__ == "__main__":
    main()
