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
