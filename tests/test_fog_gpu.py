"""Fog simulation on the GPU (csrc/fog.cu through the C ABI and the simulate_fog mirror) against the vectors frozen
from the reference (tests/golden/fog.npz) and against the oracle on larger clouds.

Exact: fog mask, ranks, counts, the position of the caller's random stream after the call.  By tolerance (the
reference is host-defined there): values that pass through the float32 np.exp / scalar float32 power /
pow -- relative 3e-7 (two float32 ulps) on intensities and coordinates, and a rounded hard-target intensity may differ
by one count where exp(-2 alpha r) * I lies within 1e-4 of a half-integer."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_snow_sim_b200.synthetic import synthetic_cloud        # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, 'tests', 'golden', 'fog.npz')


@pytest.fixture(scope='module')
def engine():
    from lidar_snow_sim_b200.engine import SnowfallEngine
    return SnowfallEngine(0)


@pytest.fixture(scope='module')
def oracle():
    from oracle import fog as ofog
    return ofog


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def close(got, want, rtol=3e-7):
    return np.allclose(got, want, rtol=rtol, atol=0, equal_nan=True)


def check_against(aug, fog, info, w_aug, w_fog, w_info, hard_alpha=None, pc=None):
    assert aug.dtype == w_aug.dtype and aug.shape == w_aug.shape
    bad = ~np.isclose(aug, w_aug, rtol=3e-7, atol=0, equal_nan=True)
    if bad.any():                                   # only hard-target roundings at a near tie may differ, by one count
        rows, cols = np.nonzero(bad)
        assert hard_alpha is not None and np.all(cols == 3) and rows.size <= max(1, aug.shape[0] // 500)
        r0 = np.linalg.norm(pc[rows, :3].astype(np.float64), axis=1)
        prod = np.exp(-2 * hard_alpha * r0) * pc[rows, 3]
        assert np.all(np.abs(prod - np.floor(prod) - 0.5) < 1e-4) and np.all(np.abs(aug[rows, 3] - w_aug[rows, 3]) <= 1)
    if w_fog is not None:
        assert (fog is None and w_fog.shape[0] == 0) or (fog.shape == w_fog.shape and close(fog, w_fog))
    if w_info is not None:
        assert info['num_fog_responses'] == int(w_info[2])
        assert close(np.array([info['min_fog_response'], info['max_fog_response']]), w_info[:2])


def test_golden_cases(engine, gold):
    """Every flag / noise-variant combination the reference was run on, through the simulate_fog mirror."""
    from lidar_snow_sim_b200.fog import ParameterSet, simulate_fog
    for i in range(int(gold['n_cases'])):
        alpha, variant, noise, gain, hard, soft, nf = gold[f'case{i}_cfg']
        pc = gold['pc4'] if int(nf) == 4 else gold['pc']
        rng = np.random.default_rng(seed=42)
        p = ParameterSet(alpha=float(alpha), gamma=0.000001)
        aug, fog, info = simulate_fog(p, pc, int(noise), gain=bool(gain), noise_variant=f'v{int(variant)}',
                                      hard=bool(hard), soft=bool(soft), engine=engine, lut=gold[f'lut_{float(alpha)}'],
                                      rng=rng)
        w_aug = gold[f'case{i}_aug']
        if soft:
            w_fog = gold[f'case{i}_fog']
            assert (0 if fog is None else len(fog)) == len(w_fog), i
            check_against(aug, fog, info, w_aug, w_fog, gold[f'case{i}_info'], float(alpha) if hard else None, pc)
        else:
            assert fog is None and info is None
            check_against(aug, None, None, w_aug, None, None, float(alpha), pc)
        assert np.array_equal(rng.random(2), gold[f'case{i}_next_u']), i      # same stream position as the reference


def test_vs_oracle_batch(engine, oracle, gold):
    """A ragged batch with per-cloud generator states == the oracle cloud by cloud; ranks are positions in point order."""
    lut = gold['lut_0.06']
    d_lut = torch.from_numpy(lut).cuda()
    clouds = [synthetic_cloud(seed=900 + b, n_azimuth=96 + 32 * b) for b in range(3)] + [np.zeros((0, 5), np.float32)]
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(clouds)).cuda()
    p = oracle.ParameterSet(alpha=0.06, gamma=0.000001)
    for variant, noise in ((1, 10), (2, 10), (3, 4), (1, 0)):
        rngs = [np.random.default_rng(100 + b) for b in range(len(clouds))]
        from lidar_snow_sim_b200.fog.simulation import _pcg64_state
        states = []
        for r in rngs:
            r.integers(low=1, high=20, size=1)
            states.append(_pcg64_state(r))
        res = engine.fog_batch(pts, off, d_lut, p.alpha, p.beta, p.beta_0, noise=noise, noise_variant=variant,
                               rng_states=np.stack(states), want_rank=True)
        engine.check()
        got = res['points'].cpu().numpy()
        mask = res['fog_mask'].cpu().numpy().astype(bool)
        rank = res['rank'].cpu().numpy()
        info = res['info'].cpu().numpy()
        for b, c in enumerate(clouds):
            sl = slice(off[b], off[b + 1])
            w_aug, w_fog, w_info = oracle.simulate_fog(p, c, noise, lut, np.random.default_rng(100 + b),
                                                       noise_variant=f'v{variant}')
            assert close(got[sl], w_aug), (variant, b)
            assert int(info[b, 2]) == w_info['num_fog_responses']
            assert np.array_equal(rank[sl][mask[sl]], np.arange(mask[sl].sum()))
            assert np.all(rank[sl][~mask[sl]] == -1)
            if w_info['num_fog_responses']:
                assert close(info[b, :2], [w_info['min_fog_response'], w_info['max_fog_response']])
                assert close(got[sl][mask[sl]], w_fog)
            else:
                assert np.isinf(info[b, 0]) and info[b, 1] == 0


def test_full_size_and_errors(engine, gold):
    """131 072-point clouds: external uniforms == generator states (same numbers, two routes); invalid arguments."""
    lut = torch.from_numpy(gold['lut_0.2']).cuda()
    clouds = [synthetic_cloud(seed=40 + b, n_azimuth=2048) for b in range(2)]
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(clouds)).cuda()
    from lidar_snow_sim_b200.fog import ParameterSet
    from lidar_snow_sim_b200.fog.simulation import _pcg64_state
    p = ParameterSet(alpha=0.2, gamma=0.000001)
    rngs = [np.random.default_rng(5 + b) for b in range(2)]
    a = engine.fog_batch(pts, off, lut, p.alpha, p.beta, p.beta_0, noise=10, noise_variant=1, gain=True,
                         rng_states=np.stack([_pcg64_state(r) for r in rngs]), want_rank=True)
    cnt = a['info'][:, 2].cpu().numpy().astype(int)
    assert cnt.min() > 1000
    ext = torch.zeros(pts.shape[0], dtype=torch.float64)
    for b in range(2):
        ext[off[b]:off[b] + cnt[b]] = torch.from_numpy(rngs[b].random(cnt[b]))
    e = engine.fog_batch(pts, off, lut, p.alpha, p.beta, p.beta_0, noise=10, noise_variant=1, gain=True,
                         ext_noise=ext.cuda())
    engine.check()
    assert torch.equal(a['points'], e['points']) and torch.equal(a['fog_mask'], e['fog_mask'])
    out = a['points'].cpu().numpy()
    for b in range(2):                                          # gain: the brightest return of each cloud is 255
        assert out[off[b]:off[b + 1], 3].max() == pytest.approx(255.0, rel=1e-12)
    with pytest.raises(ValueError):
        engine.fog_batch(pts, off, None, p.alpha, p.beta, p.beta_0)                       # soft target without table
    with pytest.raises(ValueError):
        engine.fog_batch(pts, off, lut, p.alpha, p.beta, p.beta_0, noise=5, noise_variant=7)
    from lidar_snow_sim_b200.fog import simulate_fog
    with pytest.raises(NotImplementedError):                                              # fog_simulation.py:264-266
        simulate_fog(p, clouds[0][:100], 10, noise_variant='v9', engine=engine, lut=gold['lut_0.2'])


def test_foggify_cvl_from_pickled_tables(engine, gold, tmp_path):
    """The DenseDataset.foggify 'CVL' branch through the table loader: pickles in the reference's file layout."""
    import pickle
    from lidar_snow_sim_b200.fog import ParameterSet, simulate_fog
    from lidar_snow_sim_b200.integrations.dense import foggify_cvl
    for a in (0.06, 0.2):
        lut = gold[f'lut_{a}']
        d = {round(i * 0.1, 2): (np.float64(lut[i, 0]), np.float64(lut[i, 1])) for i in range(2001)}
        (tmp_path / f'integral_0m_to_200m_stepsize_0.1m_tau_h_20ns_alpha_{a}.pickle').write_bytes(pickle.dumps(d))
    pc = gold['pc']
    got = foggify_cvl(pc, '0.200', {'FOG_NOISE_VARIANT': 'v2', 'FOG_GAIN': True}, engine=engine, lut_dir=tmp_path,
                      rng=np.random.default_rng(42))
    want, _, _ = simulate_fog(ParameterSet(alpha=0.2, gamma=0.000001), pc, 10, gain=True, noise_variant='v2',
                              engine=engine, lut=gold['lut_0.2'], rng=np.random.default_rng(42))
    assert got.dtype == np.float64 and np.array_equal(got, want)
    assert foggify_cvl(pc, '0.000', {}, engine=engine) is pc
    hard_only = foggify_cvl(pc, '0.060', {'FOG_SOFT': False}, engine=engine, lut_dir=tmp_path)
    assert hard_only.dtype == np.float32 and np.array_equal(hard_only[:, :3], pc[:, :3])
    assert np.all(hard_only[:, 3] <= pc[:, 3])
