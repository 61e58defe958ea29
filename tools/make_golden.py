"""
Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (/root/reference, imported via
oracle/ref_harness.py with its three in-memory shims) in the build container, and check the CPU oracle against it
on the way.  The GPU box has no /root/reference: tests only read the committed fixtures.

    python tools/make_golden.py            # writes tests/golden/*.npz, *.json ; exits non-zero on oracle mismatch

Host note: the reference's float32 np.arctan2 is host/SIMD dependent; the fixtures store the
theta bits this host produced so that every consumer can replay them exactly.
"""
import hashlib
import json
import os
import platform
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh            # noqa: E402
from oracle import oracle as orc                # noqa: E402
from lidar_snow_sim_b200.synthetic import synthetic_cloud, synthetic_particles      # noqa: E402
from lidar_snow_sim_b200.calib.hdl64e_s3 import sensor_arrays, HDL64E_S3            # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
DIV = float(np.degrees(3e-3))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def canon(a):
    return a[np.lexsort(a.T[::-1])]


def channel_infos():
    infos = []
    for (lid, fd, fs, mi, vc) in HDL64E_S3:
        d = {'focal_distance': fd, 'focal_slope': fs}
        if mi is not None:
            d['min_intensity'] = mi
        infos.append(d)
    return infos


def write_tables(root, prefix, tables):
    d = os.path.join(root, 'training', 'snowflakes', 'npy')
    os.makedirs(d, exist_ok=True)
    for k, t in enumerate(tables):
        np.save(os.path.join(d, f'{prefix}_{k + 1}.npy'), t)


class CapturePrepass:
    """Records, while the unmodified reference runs, the library-defined choices of its pre-pass so that they can be
    replayed on the device: the RANSAC plane (planes.py:35), the np.argpartition picks (augmentation.py:236), the two
    linregress fits (:216, :249) and the np.polyfit result (simulation.py:467)."""

    def __init__(self, ns):
        self.ns = ns
        self.planes, self.ymins, self.fits, self.polys = [], [], [], []

    def __enter__(self):
        ns = self.ns
        self._cp_sim, self._cp_wet = ns.sim.calculate_plane, ns.wet_aug.calculate_plane
        self._argpart, self._polyfit, self._linreg = np.argpartition, np.polyfit, ns.wet_aug.linregress
        cap = self

        def cp(p, *a, **k):
            w, h = cap._cp_wet(p, *a, **k)
            cap.planes.append((np.asarray(w, dtype=np.float64), float(h)))
            return w, h

        def ap(a, kth, axis=-1, *r, **k):
            out = cap._argpart(a, kth, axis, *r, **k)
            if getattr(a, 'shape', None) == (50, 2555) and kth == 2 and axis == 1:
                cap.ymins.append(np.asarray(out[:, 0], dtype=np.int32).copy())
            return out

        def pf(x, y, deg, *a, **k):
            r = cap._polyfit(x, y, deg, *a, **k)
            cap.polys.append(np.asarray(r, dtype=np.float64))
            return r

        def lr(x, y=None, *a, **k):
            r = cap._linreg(x, y, *a, **k)
            cap.fits.append((float(r[0]), float(r[1])))
            return r

        ns.sim.calculate_plane = cp
        ns.wet_aug.calculate_plane = cp
        np.argpartition = ap
        np.polyfit = pf
        ns.wet_aug.linregress = lr
        return self

    def __exit__(self, *exc):
        ns = self.ns
        ns.sim.calculate_plane, ns.wet_aug.calculate_plane = self._cp_sim, self._cp_wet
        np.argpartition, np.polyfit, ns.wet_aug.linregress = self._argpart, self._polyfit, self._linreg
        return False


def main():
    ns = rh.load()
    os.makedirs(GOLD, exist_ok=True)
    ok = True
    meta = {'numpy': np.__version__, 'machine': platform.machine(), 'processor': platform.processor(),
            'NPY_DISABLE_CPU_FEATURES': os.environ.get('NPY_DISABLE_CPU_FEATURES', '')}

    # ---------------------------------------------------------------- B-1 scalars (sampling.py:23-87)
    scal = {}
    for rs, v in [(0.5, 2.0), (1.0, 1.6), (2.0, 2.0), (2.5, 1.6), (1.5, 0.6), (10.0, 0.2)]:
        rr = float(ns.sampling.snowfall_rate_to_rainfall_rate(rs, v))
        scal[f'{rs}_{v}'] = {'occupancy': float(ns.sampling.compute_occupancy(rs, v)), 'rainfall_rate': rr,
                             'gunn': float(ns.sampling.gunn_marshall(rr)),
                             'sekhon': float(ns.sampling.sekhon_srivastava(rr)),
                             'back': float(ns.sampling.rainfall_rate_to_snowfall_rate(rr, v))}
    json.dump({'meta': meta, 'scalars': scal}, open(os.path.join(GOLD, 'kat_scalars.json'), 'w'), indent=1)

    # ---------------------------------------------------------------- B-2 compute_occlusion_dict (simulation.py:231-295)
    PI = np.pi
    cases = []
    base = [[2 * PI - 0.0010, 2 * PI - 0.0006, 5.0], [0.0004, 0.0008, 7.0], [2 * PI - 0.0002, 0.0001, 9.0]]
    for rot in (0.0, 1.0):
        beam = ((2 * PI - 0.0015 + rot) % (2 * PI) if rot else 2 * PI - 0.0015, 0.0015 + rot)
        iv = np.array([[(a + rot) % (2 * PI) if rot else a, (b + rot) % (2 * PI) if rot else b, c] for a, b, c in base])
        d = ns.sim.compute_occlusion_dict(beam, iv.copy(), 30.0, DIV)
        cases.append({'beam': list(beam), 'intervals': iv.tolist(),
                      'dict': {str(k): [float(v[0]), float(v[1])] for k, v in d.items()}})
    json.dump({'meta': meta, 'beam_divergence_deg': DIV, 'cases': cases},
              open(os.path.join(GOLD, 'kat_occlusion_dict.json'), 'w'), indent=1)

    # ---------------------------------------------------------------- B-3 hand-checkable channel (simulation.py:50-194)
    root = tempfile.mkdtemp()
    kat_particles = np.array([[5, 0.004, 0.003], [12, 0.030, 0.004], [20, -0.010, 0.002], [8, 8, 0.005],
                              [30, 30, 0.008], [0, 15, 0.006]], dtype=np.float64)
    write_tables(root, 'kat', [kat_particles] * 64)
    c45 = np.float32(25 * np.cos(np.pi / 4))
    kat_pts = np.array([[40, 0, 0, 100, 2], [c45, c45, 0, 80, 2], [0, 15.1, 0, 60, 2], [-30, 0, 1, 50, 2]],
                       dtype=np.float32)
    infos = channel_infos()
    s, idx, out = ns.sim.process_single_channel(root, 'kat', kat_pts, DIV, list(range(64)), infos, 2)
    occl = ns.sim.get_occlusions(
        beam_angles=None if False else _beam_angles(kat_pts), ranges_orig=np.linalg.norm(kat_pts[:, :3].T, axis=0),
        root_path=root, particle_file='kat_3.npy', beam_divergence=DIV)
    o_out, o_s, o_n, o_th = orc.snow_channel(kat_pts, kat_particles, DIV, infos[2]['focal_distance'],
                                             infos[2]['focal_slope'], infos[2].get('min_intensity', 0), 255,
                                             theta=np.arctan2(kat_pts[:, 1], kat_pts[:, 0]))
    good = np.array_equal(out, o_out) and float(s) == o_s
    print('KAT channel  oracle==reference:', good)
    ok &= good
    np.savez(os.path.join(GOLD, 'kat_channel.npz'), particles=kat_particles, points=kat_pts, out=out,
             intensity_diff_sum=float(s), theta=np.arctan2(kat_pts[:, 1], kat_pts[:, 0]),
             n_occluders=np.array([len(d) - 1 for d in occl], dtype=np.int32),
             occl_json=json.dumps([{str(k): [float(v[0]), float(v[1])] for k, v in d.items()} for d in occl]))

    # ---------------------------------------------------------------- seeded channel cases incl. seam beams, near flakes
    sensor = sensor_arrays()
    rec = {}
    ci = 0
    for (seed, ch, n_part, M) in [(11, 0, 27000, 192), (12, 33, 49000, 192), (13, 53, 18000, 192), (14, 39, 27000, 160)]:
        rng = np.random.default_rng(seed)
        table = synthetic_particles(seed, n_part)
        # add a few flakes very close to the sensor (large angular width) and around the azimuth-0 seam
        extra = np.array([[0.45, 0.02, 0.004], [0.9, -0.3, 0.006], [0.95, 0.0005, 0.003], [3.0, -0.0004, 0.002],
                          [6.0, 0.0009, 0.0031], [10.0, -0.011, 0.009], [-2.0, 0.003, 0.004], [1.2, 0.0, 0.0045]])
        table = np.vstack((table, extra))
        az = rng.uniform(-np.pi, np.pi, M)
        az[:48] = rng.uniform(-0.004, 0.004, 48)                  # seam beams (straddling azimuth 0)
        az[48:56] = np.pi + rng.uniform(-0.002, 0.002, 8)         # around +-pi
        az[56:60] = np.pi / 2                                     # vertical limit-line special cases
        d = rng.uniform(1.5, 110.0, M)
        d[60:70] = rng.uniform(0.85, 1.05, 10)                    # xsi ramp
        el = rng.uniform(-0.4, 0.03, M)
        pts = np.stack([d * np.cos(el) * np.cos(az), d * np.cos(el) * np.sin(az), d * np.sin(el),
                        np.round(rng.uniform(1, 255, M)), np.full(M, ch)], axis=1).astype(np.float32)
        root = tempfile.mkdtemp()
        write_tables(root, 'g', [table] * 64)
        s, idx, out = ns.sim.process_single_channel(root, 'g', pts, DIV, list(range(64)), infos, ch)
        theta = np.arctan2(pts[:, 1], pts[:, 0])
        o_out, o_s, o_n, o_th = orc.snow_channel(pts, table, DIV, sensor[0][ch], sensor[1][ch], sensor[2][ch],
                                                 sensor[3][ch], theta=theta)
        good = np.array_equal(out, o_out) and float(s) == o_s
        print(f'channel case {ci} (ch {ch}, Np {table.shape[0]}): oracle==reference: {good}; labels',
              [(out[:, 4] == l).sum() for l in (0, 1, 2)], 'max occluders', o_n.max())
        ok &= good
        rec[f'c{ci}_seed'] = seed
        rec[f'c{ci}_channel'] = ch
        rec[f'c{ci}_npart'] = n_part
        rec[f'c{ci}_extra'] = extra
        rec[f'c{ci}_table_sha'] = sha(table)
        rec[f'c{ci}_points'] = pts
        rec[f'c{ci}_theta'] = theta
        rec[f'c{ci}_out'] = out
        rec[f'c{ci}_sum'] = float(s)
        rec[f'c{ci}_nocc'] = o_n
        ci += 1
    rec['n_cases'] = ci
    np.savez_compressed(os.path.join(GOLD, 'channel_cases.npz'), **rec)

    # ---------------------------------------------------------------- end-to-end augment (simulation.py:427-544)
    for name, seed, n_az, n_part, fov in [('augment_a', 0, 192, 27000, False), ('augment_b', 1, 128, 18000, True)]:
        pc = synthetic_cloud(seed=seed, n_azimuth=n_az, drop=0.08 if seed else 0.0, shuffle_rows=bool(seed))
        tables = [synthetic_particles(5000 + 64 * seed + k, n_part) for k in range(64)]
        root = tempfile.mkdtemp()
        write_tables(root, 'g', tables)
        with CapturePrepass(ns) as cp_:
            random.seed(seed)
            np.random.seed(seed)
            stats, aug = ns.sim.augment(pc, 'g', DIV, shuffle=True, show_progressbar=True, only_camera_fov=fov,
                                        root_path=root)
        cap = {'plane': cp_.planes[0], 'poly': cp_.polys[0], 'ymins': cp_.ymins[0], 'fits': np.array(cp_.fits[:2])}
        random.seed(seed)
        order = list(range(64))
        random.shuffle(order)
        pcs = pc[pc[:, 4].argsort()]
        theta_sorted = np.arctan2(pcs[:, 1], pcs[:, 0])
        calib = fov_calib(ns)
        o_stats, o_aug, internals = orc.augment(pc, tables, DIV, sensor, order=order, plane=cap['plane'],
                                                theta_sorted=theta_sorted, only_camera_fov=fov, calib=calib,
                                                return_internals=True)
        stats = tuple(int(v) for v in stats)
        good = (stats == o_stats) and aug.shape == o_aug.shape and np.array_equal(canon(aug), canon(o_aug))
        good_poly = np.allclose(internals['thresh_poly'], cap['poly'], rtol=1e-12, atol=0)
        print(f'{name}: oracle==reference: {good} (poly match {good_poly}) stats {stats} out {aug.shape}')
        ok &= good and good_poly
        # theta in ORIGINAL row order so that consumers with a different (stable) channel sort can use it
        theta_orig = np.arctan2(pc[:, 1], pc[:, 0])
        np.savez_compressed(os.path.join(GOLD, f'{name}.npz'), seed=seed, n_azimuth=n_az, n_part=n_part,
                            drop=0.08 if seed else 0.0, shuffle_rows=bool(seed), fov=fov,
                            cloud_sha=sha(pc), table_sha=np.array([sha(t) for t in tables]),
                            order=np.array(order, dtype=np.int32), plane_w=cap['plane'][0], plane_h=cap['plane'][1],
                            thresh_poly=cap['poly'], ymins=cap['ymins'], fits=cap['fits'], theta=theta_orig,
                            stats=np.array(stats, dtype=np.int64), out=canon(aug))

    # ---------------------------------------------------------------- FULL-SIZE augment: 64 x 2048 cloud, dart-throwing tables
    # (BASELINE.json configs[0]).  Only hashes / small arrays are stored: the cloud and the tables are regenerated from
    # seeds, the host's float32 arctan2 bits are stored as the ulp offset from the correctly rounded value.
    from lidar_snow_sim_b200.snowfall.sampling import sample_table_set
    pc = synthetic_cloud(seed=0, n_azimuth=2048, drop=0.08)
    tables = sample_table_set('gunn', 1.0, 1.6, seed=1000)
    root = tempfile.mkdtemp()
    write_tables(root, 'g', tables)
    with CapturePrepass(ns) as cp_:
        random.seed(7)
        np.random.seed(7)
        stats, aug = ns.sim.augment(pc, 'g', DIV, shuffle=True, show_progressbar=True, only_camera_fov=False, root_path=root)
    cap = {'plane': cp_.planes[0], 'poly': cp_.polys[0], 'ymins': cp_.ymins[0], 'fits': np.array(cp_.fits[:2])}
    random.seed(7)
    order = list(range(64))
    random.shuffle(order)
    theta = np.arctan2(pc[:, 1], pc[:, 0])
    theta_cr = np.arctan2(pc[:, 1].astype(np.float64), pc[:, 0].astype(np.float64)).astype(np.float32)
    ulp = (theta.view(np.int32).astype(np.int64) - theta_cr.view(np.int32).astype(np.int64))
    assert np.abs(ulp).max() < 100
    idx = pc[:, 4].argsort(kind='stable')
    o_stats, o_aug, o_int = orc.augment(pc, tables, DIV, sensor, order=order, thresh_poly=cap['poly'],
                                        theta_sorted=theta[idx], stable_sort=True, return_internals=True)
    stats = tuple(int(v) for v in stats)
    good = (stats == o_stats) and aug.shape == o_aug.shape and np.array_equal(canon(aug), canon(o_aug))
    print(f'augment_full: oracle==reference: {good} stats {stats} out {aug.shape} labels',
          [(aug[:, 4] == l).sum() for l in (0, 1, 2)])
    ok &= good
    np.savez_compressed(os.path.join(GOLD, 'augment_full.npz'), seed=0, n_azimuth=2048, drop=0.08, cloud_sha=sha(pc),
                        table_sha=np.array([sha(t) for t in tables]), table_counts=np.array([t.shape[0] for t in tables]),
                        order=np.array(order, dtype=np.int32), plane_w=cap['plane'][0], plane_h=cap['plane'][1],
                        thresh_poly=cap['poly'], ymins=cap['ymins'], fits=cap['fits'], theta_ulp=ulp.astype(np.int8),
                        stats=np.array(stats, dtype=np.int64),
                        out_sha=sha(canon(aug)), out_shape=np.array(aug.shape),
                        label_counts=np.array([(aug[:, 4] == l).sum() for l in (0, 1, 2)]),
                        label_counts_unfiltered=np.array([(o_int['full'][:, 4] == l).sum() for l in (0, 1, 2)]))

    # ---------------------------------------------------------------- BASELINE.json configs[1] / configs[2], one cloud each:
    # 64 x 2048 cloud (no drop), 2.5 mm/h Gunn-Marshall dart-throwing tables; then the wet-ground model on the snow
    # output the way the viewer chains them (pointcloud_viewer.py:2804-2821: replace=False).  Stored: hashes, the
    # pre-pass choices of both stages, and the wet stage's float64 intensities in canonical row order.
    pc = synthetic_cloud(seed=1, n_azimuth=2048)
    tables = sample_table_set('gunn', 2.5, 1.6, seed=1000)
    root = tempfile.mkdtemp()
    write_tables(root, 'g', tables)
    with CapturePrepass(ns) as cp_:
        random.seed(11)
        np.random.seed(11)
        stats, aug = ns.sim.augment(pc, 'g', DIV, shuffle=True, show_progressbar=True, only_camera_fov=False, root_path=root)
    cap = {'plane': cp_.planes[0], 'poly': cp_.polys[0], 'ymins': cp_.ymins[0], 'fits': np.array(cp_.fits[:2])}
    random.seed(11)
    order = list(range(64))
    random.shuffle(order)
    theta = np.arctan2(pc[:, 1], pc[:, 0])
    theta_cr = np.arctan2(pc[:, 1].astype(np.float64), pc[:, 0].astype(np.float64)).astype(np.float32)
    ulp = (theta.view(np.int32).astype(np.int64) - theta_cr.view(np.int32).astype(np.int64))
    assert np.abs(ulp).max() < 100
    idx = pc[:, 4].argsort(kind='stable')
    o_stats, o_aug, o_int = orc.augment(pc, tables, DIV, sensor, order=order, plane=cap['plane'],
                                        theta_sorted=theta[idx], stable_sort=True, return_internals=True)
    stats = tuple(int(v) for v in stats)
    good = (stats == o_stats) and aug.shape == o_aug.shape and np.array_equal(canon(aug), canon(o_aug))
    good_poly = np.allclose(o_int['thresh_poly'], cap['poly'], rtol=1e-12, atol=0)
    print(f'augment_cfg1: oracle==reference: {good} (poly match {good_poly}) stats {stats} out {aug.shape} labels',
          [(aug[:, 4] == l).sum() for l in (0, 1, 2)])
    ok &= good and good_poly
    with CapturePrepass(ns) as cw_:
        np.random.seed(12)
        wet = ns.wet_aug.ground_water_augmentation(aug, water_height=0.001, debug=False, replace=False)
    o_wet = orc.ground_water_augmentation(aug, water_height=0.001, replace=False, plane=cw_.planes[0])
    good = wet.shape == o_wet.shape and np.array_equal(wet, o_wet)
    print('snow -> wet (cfg2): oracle==reference:', good, wet.shape, 'kept ground', int((wet[:, 4] == 1).sum()))
    ok &= good
    wkey = np.lexsort((wet[:, 4], wet[:, 2], wet[:, 1], wet[:, 0]))          # canonical order WITHOUT the intensity column
    wet_c = wet[wkey]
    np.savez_compressed(os.path.join(GOLD, 'augment_cfg1.npz'), seed=1, n_azimuth=2048, drop=0.0, cloud_sha=sha(pc),
                        snowfall_rate=2.5, terminal_velocity=1.6, table_seed=1000,
                        table_sha=np.array([sha(t) for t in tables]), table_counts=np.array([t.shape[0] for t in tables]),
                        order=np.array(order, dtype=np.int32), plane_w=cap['plane'][0], plane_h=cap['plane'][1],
                        thresh_poly=cap['poly'], ymins=cap['ymins'], fits=cap['fits'], theta_ulp=ulp.astype(np.int8),
                        stats=np.array(stats, dtype=np.int64), out_sha=sha(canon(aug)), out_shape=np.array(aug.shape),
                        label_counts=np.array([(aug[:, 4] == l).sum() for l in (0, 1, 2)]),
                        label_counts_unfiltered=np.array([(o_int['full'][:, 4] == l).sum() for l in (0, 1, 2)]),
                        wet_plane_w=cw_.planes[0][0], wet_plane_h=cw_.planes[0][1], wet_ymins=cw_.ymins[0],
                        wet_fits=np.array(cw_.fits[:2]), wet_shape=np.array(wet.shape),
                        wet_xyzl_sha=sha(wet_c[:, [0, 1, 2, 4]]), wet_intensity=wet_c[:, 3],
                        wet_label_counts=np.array([(wet[:, 4] == l).sum() for l in (0, 1, 2)]))

    # ---------------------------------------------------------------- wet ground (wet_ground/augmentation.py:25-161)
    pc = synthetic_cloud(seed=3, n_azimuth=256)
    with CapturePrepass(ns) as cp_:
        np.random.seed(3)
        wet = ns.wet_aug.ground_water_augmentation(pc, water_height=0.001, debug=False)
    cap = {'plane': cp_.planes[0], 'ymins': cp_.ymins[0], 'fits': np.array(cp_.fits[:2])}
    o_wet = orc.ground_water_augmentation(pc, water_height=0.001, plane=cap['plane'])
    good = wet.shape == o_wet.shape and np.array_equal(wet, o_wet)
    print('wet ground: oracle==reference:', good, wet.shape, wet.dtype)
    ok &= good
    np.savez_compressed(os.path.join(GOLD, 'wet_ground.npz'), seed=3, n_azimuth=256, cloud_sha=sha(pc),
                        plane_w=cap['plane'][0], plane_h=cap['plane'][1], ymins=cap['ymins'], fits=cap['fits'], out=wet)

    # ---------------------------------------------------------------- dart throwing (sampling.py:90-194)
    occ = float(ns.sampling.compute_occupancy(2.5, 1.6))
    rr = float(ns.sampling.snowfall_rate_to_rainfall_rate(2.5, 1.6))
    darts = {}
    for dist, seed, R0 in [('gunn', 1000, 12.0), ('sekhon', 1001, 10.0)]:
        ref_t = ns.sampling.dart_throwing(occ, rr, R0, np.random.default_rng(seed), dist)
        rng = np.random.default_rng(seed)
        o_t = orc.dart_throwing(occ, rr, R0, rng, dist)
        good = np.array_equal(ref_t, o_t)
        print(f'dart_throwing {dist}: oracle==reference: {good}, N={ref_t.shape[0]}')
        ok &= good
        darts[f'{dist}_table'] = ref_t
        darts[f'{dist}_seed'] = seed
        darts[f'{dist}_R0'] = R0
        darts[f'{dist}_next_u64'] = np.array([rng.bit_generator.random_raw()], dtype=np.uint64)
    np.savez_compressed(os.path.join(GOLD, 'dart_throwing.npz'), occupancy=occ, rainfall_rate=rr, **darts)

    print('ALL OK' if ok else 'MISMATCH')
    return 0 if ok else 1


def _beam_angles(pts):
    th = np.arctan2(pts[:, 1], pts[:, 0])
    th[th < 0] = th[th < 0] + 2 * np.pi
    ba = -np.ones((pts.shape[0], 2))
    ba[:, 0] = th - np.radians(DIV / 2)
    ba[:, 1] = th + np.radians(DIV / 2)
    ba[ba < 0] = ba[ba < 0] + 2 * np.pi
    ba[ba > 2 * np.pi] = ba[ba > 2 * np.pi] - 2 * np.pi
    return ba


def fov_calib(ns):
    c = ns.sim.get_calib()
    return {'P2': c.P2, 'R0': c.R0, 'V2C': c.V2C}


if __name__ == '__main__':
    sys.exit(main())
