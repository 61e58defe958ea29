"""
GPU parity tests of the wet-ground kernel (tools/wet_ground/augmentation.py:25-161 + phy_equations.py:35-108).

Given the same ground plane and the portable 'first minimum' rule for the least populated histogram bin (DESIGN.md
"pre-pass parity"), the device output must match the oracle: same rows in the same order (non-ground first, kept
ground after), same drop mask, labels exact, new intensities within 1e-9 relative (float64 on both sides; libm vs
CUDA transcendentals).  The reference returns float64; so does the mirror.
"""
import numpy as np
import pytest
import torch

from helpers import DIV
from lidar_snow_sim_b200.calib.hdl64e_s3 import sensor_arrays
from lidar_snow_sim_b200.synthetic import synthetic_cloud, synthetic_particles
from lidar_snow_sim_b200.wet_ground.augmentation import ground_water_augmentation

pytestmark = pytest.mark.gpu


def compare(got, want):
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.array_equal(got[:, [0, 1, 2, 4]], want[:, [0, 1, 2, 4]])
    assert np.allclose(got[:, 3], want[:, 3], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('kw', [dict(), dict(water_height=0.0005, pavement_depth=0.002, noise_floor=0.5, power_factor=10),
                                dict(flat_earth=True), dict(replace=False, delta=0.3), dict(water_height=0.01)])
def test_vs_oracle_given_plane(engine, oracle, kw):
    pc = synthetic_cloud(seed=3, n_azimuth=512, shuffle_rows=True)
    np.random.seed(3)
    w, h = oracle.calculate_plane(pc)
    want = oracle.ground_water_augmentation(pc, plane=(w, h), least_populated='first_min', **kw)
    got = ground_water_augmentation(pc, debug=False, engine=engine, plane=(w, h), **kw)
    compare(got, want)
    n_ground_kept = int((got[:, 4] == 1).sum())
    assert 0 < n_ground_kept < pc.shape[0]


def test_device_plane_and_passthrough(engine, oracle):
    pc = synthetic_cloud(seed=4, n_azimuth=1024)
    got, info = ground_water_augmentation(pc, debug=False, engine=engine, return_internals=True)
    pl = info['plane']
    want = oracle.ground_water_augmentation(pc, plane=(pl[:3], pl[3]), least_populated='first_min')
    compare(got, want)
    # fewer than 1000 ground points: the INPUT object comes back unchanged (augmentation.py:51-52)
    small = synthetic_cloud(seed=5, n_azimuth=16)
    out = ground_water_augmentation(small, debug=False, engine=engine)
    assert out is small
    with pytest.raises(NotImplementedError):
        ground_water_augmentation(pc, estimation_method='poly', debug=False, engine=engine)


def test_fused_snow_then_wet(engine, oracle):
    """BASELINE.json configs[2]: snowfall + wet ground back to back on the device (no host round trip)."""
    B = 3
    clouds = [synthetic_cloud(seed=60 + b, n_azimuth=512) for b in range(B)]
    tables = [synthetic_particles(8000 + k, 18000) for k in range(64)]
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
    orders = np.stack([np.random.default_rng(b).permutation(64) for b in range(B)]).astype(np.int32)
    poly = np.tile(np.array([1e-3, -0.2, 9.0]), (B, 1))
    tid = engine.upload_tables(tables)
    d_pts = torch.from_numpy(np.concatenate(clouds)).cuda()
    snow = engine.snowfall_batch(tid, d_pts, off, orders, DIV, thresh_poly=poly)
    wet = engine.wet_ground_batch(snow['points'], off, counts=snow['counts'], water_height=0.001, replace=False,
                                  want_intensity64=True)
    engine.check()
    counts = wet['counts'].cpu().numpy()
    planes = wet['plane'].cpu().numpy()
    pts = wet['points'].cpu().numpy()
    i64 = wet['intensity64'].cpu().numpy()
    for b in range(B):
        idx = clouds[b][:, 4].argsort(kind='stable')
        o_stats, o_aug, oi = oracle.augment(clouds[b], tables, DIV, sensor_arrays(), order=orders[b].tolist(),
                                            thresh_poly=poly[b], stable_sort=True, return_internals=True)
        # same theta caveat as everywhere: replay the device's snow output instead of the oracle's if they differ
        sn = snow['points'].cpu().numpy()[off[b]:off[b] + int(snow['counts'][b])]
        want = oracle.ground_water_augmentation(sn, water_height=0.001, replace=False,
                                                plane=(planes[b, :3], planes[b, 3]), least_populated='first_min')
        got = pts[off[b]:off[b] + counts[b]].astype(np.float64)
        got[:, 3] = i64[off[b]:off[b] + counts[b]]
        compare(got, want)
        assert set(np.unique(got[:, 4])) <= {0.0, 1.0, 2.0}          # replace=False keeps the snow labels (viewer path)
    engine.free_tables(tid)
