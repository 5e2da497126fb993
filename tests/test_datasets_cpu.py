"""CPU suite: dataset front ends. The product's host-side loaders (apps/dataset.h, through
apps/dataset_dump) against the oracle restatement; the oracle against the golden outputs of the reference's own
Read_data / Conversion / knn / Matching (tests/golden/ref_lidar.npz, ref_bunny.npz, produced on a B200 by
tests/golden/make_golden_datasets.py); and the oracle's whole loops against the reference programs' stdout.
The dataset files are build outputs of oracle/Makefile under oracle/_ref (copied from the reference tree, not
committed): tests that need them skip when they are absent."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200")
REFDIR = os.path.join(ROOT, "oracle", "_ref")
DUMP = os.path.join(PKG, "apps", "dataset_dump")

need_bunny = pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "Bunny_res.csv")), reason="oracle/_ref/Bunny_res.csv not built")
need_lidar = pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "Donut_1024x16.csv")), reason="oracle/_ref/Donut_1024x16.csv not built")


def parse_errors(text):
    """The "k: e" lines after "Error:" in the dataset programs' stdout."""
    m = re.search(r"Error:\n((?:\d+: -?[\d.]+\n)+)", text)
    return np.array([float(l.split(":")[1]) for l in m.group(1).strip().splitlines()])


def test_cloud_reader_on_ragged_text(orc):
    """Tokens separated by spaces/newlines, several points per line, blank lines, exponent notation, trailing space."""
    text = "0.5 -1.25 3e-2\n\n1 2 3 4 5 6 \n-7.5e1 8 9"
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c.txt"); open(p, "w").write(text)
        want = np.array([0.5, -1.25, 3e-2, 1, 2, 3, 4, 5, 6, -75, 8, 9], np.float32)
        assert np.array_equal(orc.read_cloud_text(p), want)
        subprocess.run([DUMP, "cloud", p, os.path.join(d, "o.bin")], check=True)
        assert np.array_equal(np.fromfile(os.path.join(d, "o.bin"), np.float32), want)
        with pytest.raises(OSError):
            orc.read_cloud_text(os.path.join(d, "missing.txt"))
        assert subprocess.run([DUMP, "cloud", os.path.join(d, "missing.txt"), os.path.join(d, "o2.bin")], capture_output=True).returncode != 0


@need_bunny
def test_bunny_loader_product_vs_oracle_vs_golden(orc, golden_dir):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([DUMP, "cloud", os.path.join(REFDIR, "Bunny_res.csv"), os.path.join(d, "o.bin")], check=True)
        got = np.fromfile(os.path.join(d, "o.bin"), np.float32)
    want = orc.read_cloud_text(os.path.join(REFDIR, "Bunny_res.csv"))
    assert got.size == 3 * 8171 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    g = np.load(os.path.join(golden_dir, "ref_bunny.npz"))
    D, M = orc.bunny_clouds(REFDIR)
    assert np.array_equal(D[:4], g["D_first"]) and np.array_equal(M[:4], g["M_first"]) and float(D.astype(np.float64).sum()) == float(g["D_sum"])


@need_bunny
def test_oracle_bunny_knn_and_matching_vs_reference_kernels(orc, golden_dir):
    """Squared-distance knn (GPU_point_to_plane_bunny.cu:47-82) and the first Matching pass: bit-exact indices."""
    g = np.load(os.path.join(golden_dir, "ref_bunny.npz"))
    D, M = orc.bunny_clouds(REFDIR)
    assert np.array_equal(orc.knn(M, 5, orc.MODE_SQ), g["nbr"])
    assert np.array_equal(orc.match(D, M, orc.MODE_SQ), g["idx_first"])


@need_lidar
def test_lidar_parsers_product_vs_oracle_vs_golden(orc, golden_dir):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([DUMP, "lidar", REFDIR, os.path.join(d, "o.bin")], check=True)
        raw = np.fromfile(os.path.join(d, "o.bin"), np.uint8)
    ranges = raw[:4 * 16384].view(np.float32); alt = raw[4 * 16384:4 * 16400].view(np.float32); az = raw[4 * 16400:4 * 16416].view(np.float32)
    enc = int(raw[4 * 16416:].view(np.uint64)[0])
    r, e = orc.lidar_parse_packets(os.path.join(REFDIR, "Donut_1024x16.csv"))
    a, z = orc.lidar_read_beams(os.path.join(REFDIR, "beam_intrinsics.csv"))
    assert np.array_equal(ranges, r) and enc == e and np.array_equal(alt, a) and np.array_equal(az, z)
    g = np.load(os.path.join(golden_dir, "ref_lidar.npz"))
    assert np.array_equal(g["ranges"], r) and int(g["encoder_count"]) == e
    # 20-bit words, 16 beams x 1024 azimuth blocks; this capture has dropouts (range 0)
    assert r.size == 16384 and r.max() < 2 ** 20 and (r == 0).sum() > 0


def test_oracle_conversion_vs_reference_kernel(orc, golden_dir):
    """Conversion (GPU_point_to_point_real.cu:20-36) and the RyT that synthesises the target: the reference's device
    cosf/sinf are not glibc's, so the CPU restatement agrees to a few ulp of the range (1e-6 relative), not bitwise."""
    g = np.load(os.path.join(golden_dir, "ref_lidar.npz"))
    P = orc.lidar_convert(g["ranges"], int(g["encoder_count"]), g["altitude"], g["azimuth"])
    scale = np.maximum(g["ranges"], 1.0)[:, None]
    assert np.abs(P - g["P_mm"]).max() / scale.max() < 1e-6 and (np.abs(P - g["P_mm"]) <= 2e-6 * scale).all()
    # dropouts collapse onto the origin exactly
    assert np.array_equal(P[g["ranges"] == 0], np.zeros_like(P[g["ranges"] == 0]))
    # RyT on the reference's own converted cloud: bit-exact (same arithmetic contract as the loop's transformation)
    Q = orc.transform(g["P_mm"], orc.euler_matrix([0.01, -0.003, 0.05]), np.array([0.001, -0.0202, 0.02], np.float32))
    assert np.array_equal(Q.view(np.uint32), g["Q_mm"].view(np.uint32))
    # first matching pass of the LiDAR point-to-point program (sentinel 1e6), on the reference's clouds
    a = np.float32(1.0 / 1000.0)
    assert np.array_equal(orc.match(g["P_mm"] * a, g["Q_mm"] * a, orc.MODE_SQ, sentinel=1e6), g["idx_first"])


@need_bunny
def test_oracle_bunny_loops_vs_reference_stdout(orc, golden_dir):
    D, M = orc.bunny_clouds(REFDIR)
    ref = parse_errors(open(os.path.join(golden_dir, "ref_bunny_p2p_stdout.txt")).read())
    o = orc.icp_p2p(D, M, max_iter=40)
    k = min(len(ref), o["iterations"] + 1)
    assert abs(len(ref) - (o["iterations"] + 1)) <= 1                      # float cuBLAS/cuSOLVER noise at the 1e-6 threshold
    assert np.abs(ref[:k] - o["errors"][:k]).max() <= 1.01e-4              # values printed with 4 decimals
    ref = parse_errors(open(os.path.join(golden_dir, "ref_bunny_p2l_stdout.txt")).read())
    nrm = orc.normals(M, orc.knn(M, 5, orc.MODE_SQ), 4)
    o = orc.icp_p2plane(D, M, nrm, max_iter=40, mode=orc.MODE_SQ)
    k = min(len(ref), o["iterations"] + 1)
    assert abs(len(ref) - (o["iterations"] + 1)) <= 1
    assert np.abs(ref[:k] - o["errors"][:k]).max() <= 1.01e-4
