"""The fog oracle (oracle/fog.py) against the vectors frozen from the reference itself (tools/make_golden_fog.py):
bit for bit, including the position of the caller's random stream after the call.  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fog as ofog                                   # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden', 'fog.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def cases(gold):
    for i in range(int(gold['n_cases'])):
        alpha, variant, noise, gain, hard, soft, nf = gold[f'case{i}_cfg']
        yield i, dict(alpha=float(alpha), variant=f'v{int(variant)}', noise=int(noise), gain=bool(gain), hard=bool(hard),
                      soft=bool(soft), pc=gold['pc4'] if int(nf) == 4 else gold['pc'])


def test_lut_key_rule(gold):
    assert np.array_equal(ofog.lut_index(gold['key_r']), gold['key_idx'])


def test_oracle_matches_reference_vectors(gold):
    n = 0
    for i, c in cases(gold):
        rng = np.random.default_rng(seed=42)
        p = ofog.ParameterSet(alpha=c['alpha'], gamma=0.000001)
        aug, fog, info = ofog.simulate_fog(p, c['pc'], c['noise'], gold[f"lut_{c['alpha']}"], rng, gain=c['gain'],
                                           noise_variant=c['variant'], hard=c['hard'], soft=c['soft'])
        want = gold[f'case{i}_aug']
        assert aug.dtype == want.dtype and aug.shape == want.shape, i
        assert np.array_equal(aug, want, equal_nan=True), (i, c['variant'])
        wf = gold[f'case{i}_fog']
        if c['soft']:
            assert (fog is None and wf.shape[0] == 0) or np.array_equal(fog, wf), i
            assert np.array_equal(np.array([info['min_fog_response'], info['max_fog_response'],
                                            info['num_fog_responses']], dtype=np.float64), gold[f'case{i}_info']), i
        else:
            assert fog is None and info is None
        assert np.array_equal(rng.random(2), gold[f'case{i}_next_u']), i       # same stream position afterwards
        n += 1
    assert n == 17


def test_parameter_set_quirk():
    """kwargs are applied last (fog_simulation.py:171): alpha=... does not re-derive beta."""
    p = ofog.ParameterSet(alpha=0.2)
    assert p.alpha == 0.2 and p.beta == 0.046 / (np.log(20) / 0.06)
