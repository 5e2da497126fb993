"""Generate the dataset golden fixtures from the UNMODIFIED reference dataset programs.

Runs on a GPU box (B200) where oracle/_ref/ was built beforehand from /root/reference by oracle/Makefile
(ref_datasets + the three dataset files; the reference tree itself is not needed at run time):

    python tests/golden/make_golden_datasets.py [outdir]          # default: gpurun_out/golden

  * ref_datasets run bunny_p2p|bunny_p2l|lidar_p2p|lidar_p2l   -> stdout of the four reference programs
  * ref_datasets lidar                                          -> the clouds the reference's Read_data produces
                                                                   (parser + Conversion kernel + RyT), with the parsed
                                                                   ranges / angles the oracle read from the same files
  * ref_datasets knn_sq                                         -> the squared-distance knn kernel on the moved bunny
  * ref_datasets match                                          -> the sentinel-1e6 Matching kernel on the LiDAR clouds
"""
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
REF = os.path.join(REFDIR, "ref_datasets")


def run_ref(args):
    return subprocess.run([REF] + [str(a) for a in args], check=True, capture_output=True, text=True, cwd=REFDIR)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    tmp = tempfile.mkdtemp()

    def f(name):
        return os.path.join(tmp, name)

    for name in ("bunny_p2p", "bunny_p2l", "lidar_p2p", "lidar_p2l"):
        r = run_ref(["run", name])
        open(os.path.join(out, "ref_%s_stdout.txt" % name), "w").write(r.stdout)
        print(name, "stdout lines:", len(r.stdout.splitlines()))

    # LiDAR front end: what the reference's Read_data leaves on the device (millimetres, before the scaling)
    run_ref(["lidar", f("P.bin"), f("Q.bin")])
    P = np.fromfile(f("P.bin"), np.float32).reshape(-1, 3)
    Q = np.fromfile(f("Q.bin"), np.float32).reshape(-1, 3)
    Pm, Qm, (r, enc, alt, az, _, _) = orc.lidar_clouds(REFDIR)
    # Matching with the LiDAR programs' sentinel on the scaled clouds (first iteration of lidar_p2p)
    a = np.float32(1.0 / 1000.0)
    (P * a).tofile(f("Ps.bin")); (Q * a).tofile(f("Qs.bin"))
    run_ref(["match", P.shape[0], Q.shape[0], f("Ps.bin"), f("Qs.bin"), f("idx.bin")])
    np.savez_compressed(os.path.join(out, "ref_lidar.npz"), ranges=r, encoder_count=np.uint64(enc), altitude=alt, azimuth=az,
                        P_mm=P, Q_mm=Q, idx_first=np.fromfile(f("idx.bin"), np.int32))
    print("lidar clouds:", P.shape, "zero ranges:", int((r == 0).sum()))

    # Bunny: squared-distance knn of the moved cloud (the target of both bunny programs)
    D, M = orc.bunny_clouds(REFDIR)
    M.tofile(f("M.bin"))
    run_ref(["knn_sq", M.shape[0], 5, f("M.bin"), f("nbr.bin")])
    D.tofile(f("D.bin"))
    run_ref(["match", D.shape[0], M.shape[0], f("D.bin"), f("M.bin"), f("idx.bin")])
    # the dataset itself is not committed: the fixture holds a checksum of what was read and the kernel outputs
    np.savez_compressed(os.path.join(out, "ref_bunny.npz"), nbr=np.fromfile(f("nbr.bin"), np.int32).reshape(-1, 5),
                        idx_first=np.fromfile(f("idx.bin"), np.int32),
                        D_sum=np.float64(D.astype(np.float64).sum()), D_first=D[:4], M_first=M[:4])
    print("done ->", out)


if __name__ == "__main__":
    main()
