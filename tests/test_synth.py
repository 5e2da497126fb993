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
