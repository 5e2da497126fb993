import numpy as np


def test_numpy_generator_is_bit_identical_to_the_c_generator(orc):
    import icp_synth
    for W, n in ((32, None), (128, None), (317, 100000)):
        D, M = icp_synth.p2p_clouds(W, n)
        Do, Mo = orc.synth_p2p(W, n)
        assert np.array_equal(D.view(np.uint32), Do.view(np.uint32))
        assert np.array_equal(M.view(np.uint32), Mo.view(np.uint32))


def test_cpu_program_clouds_match_the_c_generator(orc):
    """icp_synth.cpu_clouds (src/ICP_CPU.c's own clouds, used by tools/bench_configs.py) against the oracle's generator."""
    import icp_synth
    Dd, Md = orc.synth_cpu_f64(100)
    D, M = icp_synth.cpu_clouds(100)
    assert np.abs(D - Dd.reshape(3, -1).T.astype(np.float32)).max() <= 1e-6
    assert np.abs(M - Md.reshape(3, -1).T.astype(np.float32)).max() <= 1e-6


def test_splitmix64_known_answers_and_pose_ranges():
    """The counter-based pose generator of BASELINE config 5: splitmix64 against its published test
    vector (state 0 -> 0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4, 0x06c45d188009454f), stream independence, and the
    ranges r = 0.2 u, u in (-1,1)^3, t = (0.8,-0.3,0.2) * (0.5 + 0.5 v)."""
    import numpy as np
    import icp_synth
    st = np.zeros(1, np.uint64)
    outs = []
    for _ in range(3):
        st, z = icp_synth._splitmix64(st)
        outs.append(int(z[0]))
    assert outs == [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4, 0x06c45d188009454f]
    r, t = icp_synth.batched_poses(4096)
    r2, t2 = icp_synth.batched_poses(17)
    assert np.array_equal(r[:17], r2) and np.array_equal(t[:17], t2), "pose of pair b depends on b only"
    assert np.abs(r).max() <= 0.2 and np.abs(r).max() > 0.19
    base = np.array([0.8, -0.3, 0.2])
    ratio = t.astype(np.float64) / base
    assert ratio.min() >= 0.5 - 1e-6 and ratio.max() <= 1.0 + 1e-6
    assert len({tuple(x) for x in np.round(r, 6)}) == 4096
