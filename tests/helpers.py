"""Shared helpers of the test-suite: fixture re-generation from seeds, canonical row order."""
import hashlib

import numpy as np

from lidar_snow_sim_b200.synthetic import synthetic_cloud, synthetic_particles

DIV = float(np.degrees(3e-3))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def canon(a):
    """Rows in lexicographic order (the reference's within-channel order is implementation-defined)."""
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


def channel_case(rec, ci):
    """Rebuild the particle table of channel case `ci` of tests/golden/channel_cases.npz and check its hash."""
    table = synthetic_particles(int(rec[f'c{ci}_seed']), int(rec[f'c{ci}_npart']))
    table = np.vstack((table, rec[f'c{ci}_extra']))
    assert sha(table) == str(rec[f'c{ci}_table_sha']), 'numpy RNG stream changed: regenerate tests/golden'
    return table


def augment_case(g):
    pc = synthetic_cloud(seed=int(g['seed']), n_azimuth=int(g['n_azimuth']), drop=float(g['drop']),
                         shuffle_rows=bool(g['shuffle_rows']))
    assert sha(pc) == str(g['cloud_sha']), 'synthetic cloud changed: regenerate tests/golden'
    tables = [synthetic_particles(5000 + 64 * int(g['seed']) + k, int(g['n_part'])) for k in range(64)]
    for k in (0, 63):
        assert sha(tables[k]) == str(g['table_sha'][k])
    return pc, tables


def augment_full_case(g):
    """BASELINE.json configs[0]: full-size STF-shaped cloud + dart-throwing tables (1.0 mm/h, 1.6 m/s), from seeds."""
    from lidar_snow_sim_b200.snowfall.sampling import sample_table_set
    pc = synthetic_cloud(seed=int(g['seed']), n_azimuth=int(g['n_azimuth']), drop=float(g['drop']))
    assert sha(pc) == str(g['cloud_sha']), 'synthetic cloud changed: regenerate tests/golden'
    tables = sample_table_set('gunn', 1.0, 1.6, seed=1000)
    assert [t.shape[0] for t in tables] == g['table_counts'].tolist()
    assert [sha(t) for t in tables] == [str(v) for v in g['table_sha']], 'sampler output changed (libm?): regenerate'
    theta_cr = np.arctan2(pc[:, 1].astype(np.float64), pc[:, 0].astype(np.float64)).astype(np.float32)
    theta = (theta_cr.view(np.int32) + g['theta_ulp'].astype(np.int32)).view(np.float32)     # the reference host's bits
    return pc, tables, theta


def augment_cfg1_case(g):
    """BASELINE.json configs[1]/[2], one cloud: full 64 x 2048 cloud + 2.5 mm/h Gunn-Marshall dart-throwing tables."""
    from lidar_snow_sim_b200.snowfall.sampling import sample_table_set
    pc = synthetic_cloud(seed=int(g['seed']), n_azimuth=int(g['n_azimuth']), drop=float(g['drop']))
    assert sha(pc) == str(g['cloud_sha']), 'synthetic cloud changed: regenerate tests/golden'
    tables = sample_table_set('gunn', float(g['snowfall_rate']), float(g['terminal_velocity']), seed=int(g['table_seed']))
    assert [t.shape[0] for t in tables] == g['table_counts'].tolist()
    assert [sha(t) for t in tables] == [str(v) for v in g['table_sha']], 'sampler output changed (libm?): regenerate'
    theta_cr = np.arctan2(pc[:, 1].astype(np.float64), pc[:, 0].astype(np.float64)).astype(np.float32)
    theta = (theta_cr.view(np.int32) + g['theta_ulp'].astype(np.int32)).view(np.float32)     # the reference host's bits
    return pc, tables, theta


def canon_no_intensity(a):
    """Rows ordered by (x, y, z, label) -- the key tools/make_golden.py uses for the wet-ground fixture of configs[2]."""
    a = np.asarray(a)
    return a[np.lexsort((a[:, 4], a[:, 2], a[:, 1], a[:, 0]))]
