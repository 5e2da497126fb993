"""CPU suite: the occupancy-pyramid nearest-neighbour traversal (csrc/grid_tree.cuh) is host/device code; its host
instantiation is checked here against a literal brute-force scan on the clouds that decide exactness (far and near
field of the saddle, tie-heavy lattices with duplicates and far outliers, small sentinel, collinear / single-point
targets, large offsets, NaN / inf sources, both distance modes, cold and warm starts). The GPU kernel compiles the same
function (tests/test_gpu_grid.py::test_grid_pyramid_equals_brute checks it on the device)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200")


def test_pyramid_traversal_host_instantiation_is_exact(tmp_path):
    exe = str(tmp_path / "grid_tree_host_test")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-ffp-contract=off", "-I", os.path.join(PKG, "csrc"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tools", "grid_tree_host_test.cu"), "-o", exe], check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "all exact" in r.stdout and "MISMATCH" not in r.stdout and "FAIL" not in r.stdout
    assert r.stdout.count(" ok") >= 18
