"""
GPU tests of the device pre-pass (ground plane, laser parameters, noise-threshold polynomial).

The reference's pre-pass is library-defined (sklearn RANSAC on the global NumPy RNG, np.argpartition's pick among the
three least populated bins, float32 LAPACK fits -- DESIGN.md "pre-pass parity"), so the bars are tolerances:
  * threshold polynomial, given the SAME plane and the portable 'first minimum' bin rule: 1e-6 relative on the
    threshold it produces over 0..120 m (float64 normal equations vs NumPy's float32 Vandermonde + LAPACK);
  * plane: normal within 2e-3 rad and offset within 5 mm of the oracle's sklearn RANSAC;
  * end to end: the keep mask of augment() with the device pre-pass agrees with the oracle's on > 99.5 % of points.
"""
import numpy as np
import pytest
import torch

from helpers import DIV
from lidar_snow_sim_b200.calib.hdl64e_s3 import sensor_arrays
from lidar_snow_sim_b200.synthetic import synthetic_cloud, synthetic_particles

pytestmark = pytest.mark.gpu


def thr_of(p, d):
    return p[0] * d ** 2 + p[1] * d + p[2]


@pytest.mark.parametrize('seed,n_az,shuffle', [(0, 2048, False), (7, 1024, True), (11, 512, False)])
def test_poly_given_plane(engine, oracle, seed, n_az, shuffle):
    pc = synthetic_cloud(seed=seed, n_azimuth=n_az, drop=0.05, shuffle_rows=shuffle)
    np.random.seed(seed)
    w, h = oracle.calculate_plane(pc)
    want = oracle.noise_threshold_poly(pc, w, h, 0.7, least_populated='first_min')
    d_pc = torch.from_numpy(pc).cuda()
    poly, plane = engine.noise_threshold_poly(d_pc, [0, pc.shape[0]], 0.7, plane=np.array([[w[0], w[1], w[2], h]]))
    engine.check()
    got = poly[0].cpu().numpy()
    d = np.linspace(1.0, 120.0, 200)
    rel = np.abs(thr_of(got, d) - thr_of(want, d)) / np.maximum(np.abs(thr_of(want, d)), 1e-3)
    assert rel.max() < 1e-6, (got, want)
    assert np.allclose(plane[0].cpu().numpy(), [w[0], w[1], w[2], h])


def test_plane_vs_sklearn(engine, oracle):
    clouds = [synthetic_cloud(seed=s, n_azimuth=2048) for s in (1, 2, 3)]
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])])
    d_pc = torch.from_numpy(np.concatenate(clouds)).cuda()
    poly, plane = engine.noise_threshold_poly(d_pc, off, 0.7)
    engine.check()
    plane = plane.cpu().numpy()
    for b, c in enumerate(clouds):
        np.random.seed(b)
        w, h = oracle.calculate_plane(c)
        ang = np.arccos(np.clip(np.dot(plane[b, :3], w) / np.linalg.norm(plane[b, :3]) / np.linalg.norm(w), -1, 1))
        assert ang < 2e-3 and abs(plane[b, 3] - h) < 5e-3, (plane[b], w, h)
        assert abs(np.linalg.norm(plane[b, :3]) - 1) < 1e-12 and plane[b, 2] < 0
    # deterministic: same answer when run again, and independent of batching
    poly2, plane2 = engine.noise_threshold_poly(d_pc, off, 0.7)
    assert torch.equal(poly, poly2) and torch.equal(torch.from_numpy(plane).cuda(), plane2)
    p1, pl1 = engine.noise_threshold_poly(torch.from_numpy(clouds[1]).cuda(), [0, clouds[1].shape[0]], 0.7)
    assert torch.equal(p1[0], poly[1]) and torch.equal(pl1[0], plane2[1])


def test_augment_end_to_end_device_prepass(engine, oracle):
    from lidar_snow_sim_b200.snowfall.simulation import augment
    tables = [synthetic_particles(4000 + k, 18000) for k in range(64)]
    pc = synthetic_cloud(seed=31, n_azimuth=1024)
    order = np.random.default_rng(1).permutation(64).tolist()
    np.random.seed(31)
    o_stats, o_aug, oi = oracle.augment(pc, tables, DIV, sensor_arrays(), order=order, stable_sort=True,
                                        return_internals=True, least_populated='first_min')
    theta = np.empty(pc.shape[0], np.float32)
    theta[oi['sort_index']] = oi['theta']
    stats, aug, gi = augment(pc, 'unused', DIV, only_camera_fov=False, engine=engine, tables=tables, order=order,
                             theta=theta, return_internals=True)
    # the per-beam solve does not depend on the pre-pass: un-filtered rows are exact
    assert np.array_equal(gi['full'], oi['full'])
    # keep mask: recompute from the rows we got back
    keep_o = oi['keep']
    key = lambda a: {tuple(r) for r in a.tolist()}
    inter = len(key(aug) & key(o_aug))
    agree = 1 - (len(aug) + len(o_aug) - 2 * inter) / pc.shape[0]
    print(f'keep-mask agreement {agree:.5f}; kept {len(aug)} vs oracle {len(o_aug)}; stats {stats} vs {o_stats}')
    assert agree > 0.995
    assert abs(stats[0] - o_stats[0]) <= 0.01 * max(o_stats[0], 1) + 5


def test_too_few_ground_points(engine):
    from lidar_snow_sim_b200.snowfall.simulation import augment
    pc = synthetic_cloud(seed=2, n_azimuth=64)
    pc = pc[(pc[:, 2] > -0.5) & (pc[:, 2] < 0.9)]               # no ground returns, nothing near the flat-earth fallback plane
    tables = [synthetic_particles(k, 2000) for k in range(64)]
    with pytest.raises(TypeError):                              # estimate_laser_parameters -> None, simulation.py:457-462
        augment(pc, 'unused', DIV, only_camera_fov=False, engine=engine, tables=tables)
