"""Host-side logic of the fog mirror (no GPU): ParameterSet against the oracle's, the PCG64 stream bookkeeping."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_snow_sim_b200.fog import simulation as fsim          # noqa: E402
from oracle import fog as ofog                                   # noqa: E402


def test_parameter_set_matches_oracle():
    for kw in ({}, dict(alpha=0.2, gamma=0.000001), dict(alpha=0.005), dict(beta=1e-3, tau_h=1e-8)):
        a, b = fsim.ParameterSet(**kw), ofog.ParameterSet(**kw)
        for k, v in b.__dict__.items():
            assert a.__dict__[k] == v, k
    p = fsim.ParameterSet(alpha=0.2)
    assert p.beta == 0.046 / (np.log(20) / 0.06)            # kwargs applied last: beta is NOT re-derived (:171)


@pytest.mark.parametrize('k', [0, 1, 2, 17, 1000, 123457])
def test_pcg64_advance_is_k_draws(k):
    a, b = np.random.default_rng(7), np.random.default_rng(7)
    a.integers(low=1, high=20, size=1)                      # leaves a buffered 32-bit half in the bit generator
    b.integers(low=1, high=20, size=1)
    a.random(k)
    fsim._pcg64_advance(b, k)
    assert a.bit_generator.state == b.bit_generator.state   # including has_uint32 / uinteger
    assert np.array_equal(a.random(3), b.random(3))
    assert np.array_equal(a.integers(1, 20, size=4), b.integers(1, 20, size=4))


def test_state_words():
    r = np.random.default_rng(42)
    w = fsim._pcg64_state(r)
    st = r.bit_generator.state['state']
    assert (int(w[0]) << 64) | int(w[1]) == st['state'] and (int(w[2]) << 64) | int(w[3]) == st['inc']
    with pytest.raises(TypeError):
        fsim._pcg64_state(np.random.Generator(np.random.MT19937(1)))


def test_table_loader(tmp_path):
    import pickle
    d = {round(i * 0.1, 2): (np.float64(i * 0.05), np.float64(1e-9 * i)) for i in range(2001)}
    f = tmp_path / 'integral_0m_to_200m_stepsize_0.1m_tau_h_20ns_alpha_0.06.pickle'
    f.write_bytes(pickle.dumps(d))
    (tmp_path / 'integral_0m_to_200m_stepsize_0.1m_tau_h_20ns_alpha_0.2.pickle').write_bytes(pickle.dumps(d))
    assert fsim.get_available_alphas(tmp_path) == [0.06, 0.2]
    t = fsim.load_integral_table(fsim.ParameterSet(alpha=0.07), tmp_path)          # nearest available alpha
    assert t.shape == (2001, 2) and t[300, 0] == 300 * 0.05 and t[2000, 1] == 1e-9 * 2000
    with pytest.raises(FileNotFoundError):
        os.environ.pop('LSS_FOG_LUT_DIR', None)
        fsim.load_integral_table(fsim.ParameterSet())
