"""CPU tests of the native dart-throwing sampler (tools/snowfall/sampling.py:90-194) -- needs no GPU."""
import os

import numpy as np
import pytest

from lidar_snow_sim_b200 import build
from lidar_snow_sim_b200.snowfall import sampling as S


@pytest.fixture(scope='module', autouse=True)
def _built():
    build.build()


def test_matches_reference_fixture(gold_dir):
    """Same Generator state -> the reference's table (count, order, values; cos/sin may differ by 1 ulp across
    hosts because NumPy dispatches them to SIMD kernels) and the same stream position afterwards."""
    g = np.load(os.path.join(gold_dir, 'dart_throwing.npz'))
    for dist in ('gunn', 'sekhon'):
        rng = np.random.default_rng(int(g[f'{dist}_seed']))
        t = S.dart_throwing(float(g['occupancy']), float(g['rainfall_rate']), float(g[f'{dist}_R0']), rng, dist)
        ref = g[f'{dist}_table']
        assert t.shape == ref.shape and t.dtype == np.float64
        assert np.array_equal(t[:, 2], ref[:, 2])                      # radii involve no libm-dependent value
        assert np.allclose(t[:, :2], ref[:, :2], rtol=4e-16, atol=1e-17)
        assert rng.bit_generator.random_raw() == int(g[f'{dist}_next_u64'][0])


def test_matches_oracle_restatement(oracle):
    occ, rr = S.compute_occupancy(1.0, 1.6), float(S.snowfall_rate_to_rainfall_rate(1.0, 1.6))
    a = S.dart_throwing(occ, rr, 9.0, np.random.default_rng(77), 'gunn')
    b = oracle.dart_throwing(occ, rr, 9.0, np.random.default_rng(77), 'gunn')
    assert a.shape == b.shape and np.array_equal(a[:, 2], b[:, 2]) and np.allclose(a, b, rtol=4e-16, atol=1e-17)


def test_full_size_plane_is_bit_identical_to_the_numpy_path(oracle):
    """One full-size plane of the bench workload (R_0 = 80 m, 2.5 mm/h: ~18 k disks, ~10^5 draws): bit-identical to the
    oracle's NumPy restatement on the same host -- the scalar squares go through libm's pow() like python's `** 2`
    (0.52 ulp, not always the rounded x*x: 15 radii of this plane differ in the last bit otherwise)."""
    occ, rr = S.compute_occupancy(2.5, 1.6), float(S.snowfall_rate_to_rainfall_rate(2.5, 1.6))
    rng_a, rng_b = np.random.default_rng(1000), np.random.default_rng(1000)
    a = S.dart_throwing(occ, rr, 80.0, rng_a, 'gunn')
    b = oracle.dart_throwing(occ, rr, 80.0, rng_b, 'gunn')
    assert a.shape == b.shape == (18439, 3)
    assert np.array_equal(a[:, 2], b[:, 2])                            # radii: no SIMD-dispatched libm call involved
    assert np.allclose(a[:, :2], b[:, :2], rtol=4e-16, atol=1e-17)     # x, y: np.cos / np.sin may be SIMD kernels
    assert rng_a.bit_generator.random_raw() == rng_b.bit_generator.random_raw()


def test_table_properties():
    """Full-size plane (R_0 = 80 m): non-overlap, origin excluded, occupancy reached, diameter cap."""
    occ, rr = S.compute_occupancy(2.5, 1.6), float(S.snowfall_rate_to_rainfall_rate(2.5, 1.6))
    t = S.dart_throwing(occ, rr, 80.0, np.random.default_rng(1003), 'gunn')
    assert t.shape[0] == 17671                                         # the reference's count for this seed
    x, y, r = t.T
    assert np.all(x * x + y * y > r * r) and np.all(np.hypot(x, y) <= 80.0) and np.all(r <= 0.01) and np.all(r > 0)
    area = np.pi * (r ** 2).sum()
    target = occ * np.pi * 80.0 ** 2
    assert area >= target and area - np.pi * r[-1] ** 2 < target       # stops right after crossing the target
    # non-overlap via a grid neighbourhood check
    order = np.lexsort((y, x))
    xs, ys, rs = x[order], y[order], r[order]
    for i in range(0, xs.shape[0], 97):
        j0, j1 = np.searchsorted(xs, [xs[i] - 0.021, xs[i] + 0.021])
        d2 = (xs[j0:j1] - xs[i]) ** 2 + (ys[j0:j1] - ys[i]) ** 2
        ok = d2 > (rs[j0:j1] + rs[i]) ** 2
        ok[i - j0] = True
        assert ok.all()


def test_table_set_and_errors():
    tabs = S.sample_table_set('gunn', 2.5, 1.6, seed=1000, n_planes=4)
    one = S.dart_throwing(S.compute_occupancy(2.5, 1.6), float(S.snowfall_rate_to_rainfall_rate(2.5, 1.6)), 80.0,
                          np.random.default_rng(1003), 'gunn')
    assert np.array_equal(tabs[3], one)
    with pytest.raises(NotImplementedError):
        S.dart_throwing(1e-6, 10.0, 5.0, np.random.default_rng(0), 'marshall_palmer')
