import numpy as np


def test_numpy_generator_is_bit_identical_to_the_c_generator(orc):
    import icp_synth
    for W, n in ((32, None), (128, None), (317, 100000)):
        D, M = icp_synth.p2p_clouds(W, n)
        Do, Mo = orc.synth_p2p(W, n)
        assert np.array_equal(D.view(np.uint32), Do.view(np.uint32))
        assert np.array_equal(M.view(np.uint32), Mo.view(np.uint32))
