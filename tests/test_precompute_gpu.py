"""GPU test of the offline batch driver (reference: tools/snowfall/precompute.py:47-106)."""
import numpy as np
import pytest

from lidar_snow_sim_b200.synthetic import synthetic_cloud
from lidar_snow_sim_b200.snowfall import precompute as pre
from lidar_snow_sim_b200.snowfall.simulation import augment
from lidar_snow_sim_b200.snowfall.sampling import sample_table_set

pytestmark = pytest.mark.gpu


def test_precompute_layout_skip_and_content(engine, oracle, tmp_path):
    lidar = tmp_path / 'lidar_hdl64_strongest'
    lidar.mkdir()
    ids = ['2018-02-03_20-48-35_00400', '2018-02-03_20-48-35_00500', '2018-02-04_10-00-00_00100']
    for k, s in enumerate(ids):
        synthetic_cloud(seed=70 + k, n_azimuth=256).tofile(str(lidar / f'{s}.bin'))
    n = pre.precompute(ids, lidar, modes=('gunn',), sample_tables=True, engine=engine, batch_frames=2, shuffle=False)
    assert n == 3 * 5
    rain = [int(c[2]) for c in pre.combos()]
    for r in rain:
        d = tmp_path / 'snowfall_simulation' / 'gunn' / f'lidar_hdl64_strongest_rainrate_{r}'
        assert sorted(p.name for p in d.iterdir()) == sorted(f'{s}.bin' for s in ids)
    # idempotent: nothing is recomputed
    assert pre.precompute(ids, lidar, modes=('gunn',), sample_tables=True, engine=engine, shuffle=False) == 0
    # content == a direct augment() of the FOV-filtered frame with the same tables / order
    rs, tv, rr, occ = pre.combos()[3]
    pts = np.fromfile(str(lidar / f'{ids[1]}.bin'), dtype=np.float32).reshape(-1, 5)
    pts = pts[pre.get_fov_flag(pts[:, :3])]
    tabs = sample_table_set('gunn', rs, tv, seed=42)
    stats, want = augment(pts, 'unused', float(np.degrees(3e-3)), shuffle=False, engine=engine, tables=tabs)
    got = np.fromfile(str(tmp_path / 'snowfall_simulation' / 'gunn' / f'lidar_hdl64_strongest_rainrate_{int(rr)}' /
                          f'{ids[1]}.bin'), dtype=np.float32).reshape(-1, 5)
    assert np.array_equal(got, want) and got.shape[0] > 0
    # ... and == the ORACLE's augment() of that frame (same tables / order / plane; the device's own azimuths and
    # first-minimum bin rule): the un-filtered solve exactly on the beams whose float32 azimuth agrees, the kept rows
    # to > 99 %
    from helpers import DIV
    from lidar_snow_sim_b200.calib.hdl64e_s3 import sensor_arrays
    from lidar_snow_sim_b200.calib.dense_camera import STF_HDL64_CAMERA
    _, _, gi = augment(pts, 'unused', DIV, shuffle=False, engine=engine, tables=tabs, return_internals=True)
    poly, plane = engine.noise_threshold_poly(__import__('torch').from_numpy(pts).cuda(), [0, pts.shape[0]], 0.7)
    pl = plane[0].cpu().numpy()
    o_stats, o_aug, oi = oracle.augment(pts, tabs, DIV, sensor_arrays(), order=list(range(64)), plane=(pl[:3], pl[3]),
                                        only_camera_fov=True, calib=STF_HDL64_CAMERA, stable_sort=True,
                                        return_internals=True, least_populated='first_min')
    th64 = np.arctan2(pts[:, 1].astype(np.float64), pts[:, 0].astype(np.float64)).astype(np.float32)
    same_theta = (th64[oi['sort_index']] == oi['theta'])
    assert np.array_equal(gi['full'][same_theta], oi['full'][same_theta])
    key = lambda a: {tuple(r) for r in a.tolist()}
    assert len(key(got) & key(o_aug)) > 0.99 * max(len(got), len(o_aug))
    with pytest.raises(FileNotFoundError):
        pre.precompute(['x'], lidar, modes=('sekhon',), npy_root=tmp_path / 'no_tables', engine=engine)
