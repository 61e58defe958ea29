"""Generate tests/golden/fog.npz by running the REFERENCE's fog simulation (imported from /root/reference, build
container only) on seeded synthetic clouds.  Stores the inputs, the integral look-up tables the reference read (as plain
(2001, 2) float64 arrays per alpha) and the reference's outputs for every noise variant / flag combination.
    python tools/make_golden_fog.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/lib/LiDAR_fog_sim')
import fog_simulation as ref                                    # noqa: E402  (the reference itself)
from lidar_snow_sim_b200.synthetic import synthetic_cloud        # noqa: E402


def lut_array(p):
    d = ref.get_integral_dict(p)
    keys = sorted(d.keys())
    assert len(keys) == 2001 and keys[0] == 0 and keys[-1] == 200.0
    return np.array([[float(d[k][0]), float(d[k][1])] for k in keys], dtype=np.float64)


def main():
    out = {}
    alphas = [0.005, 0.06, 0.2]
    for a in alphas:
        out[f'lut_{a}'] = lut_array(ref.ParameterSet(alpha=a, gamma=0.000001))
    pc = synthetic_cloud(seed=77, n_azimuth=16)
    pc[5, :3] *= 250.0 / np.linalg.norm(pc[5, :3])              # beyond the 200 m cap of the table
    pc[6, :3] *= 0.04 / np.linalg.norm(pc[6, :3])               # LUT entry 0
    out['pc'] = pc
    out['pc4'] = np.ascontiguousarray(pc[:, :4])
    cases = []
    for a in alphas:
        for variant in ('v1', 'v2', 'v3', 'v4'):
            cases.append(dict(alpha=a, variant=variant, noise=10, gain=False, hard=True, soft=True, key='pc'))
    cases += [dict(alpha=0.06, variant='v1', noise=0, gain=False, hard=True, soft=True, key='pc'),
              dict(alpha=0.06, variant='v1', noise=10, gain=True, hard=True, soft=True, key='pc'),
              dict(alpha=0.06, variant='v1', noise=10, gain=False, hard=False, soft=True, key='pc'),
              dict(alpha=0.06, variant='v1', noise=10, gain=False, hard=True, soft=False, key='pc'),
              dict(alpha=0.2, variant='v2', noise=3, gain=True, hard=True, soft=True, key='pc4')]
    out['n_cases'] = np.array(len(cases))
    for i, c in enumerate(cases):
        ref.RNG = np.random.default_rng(seed=42)                # the module-level generator, fresh (fog_simulation.py:15)
        p = ref.ParameterSet(alpha=float(c['alpha']), gamma=0.000001)           # dense_dataset.py:990
        aug, fog, info = ref.simulate_fog(p, pc=out[c['key']], noise=c['noise'], gain=c['gain'],
                                          noise_variant=c['variant'], hard=c['hard'], soft=c['soft'])
        out[f'case{i}_cfg'] = np.array([c['alpha'], int(c['variant'][1]), c['noise'], c['gain'], c['hard'], c['soft'],
                                        4 if c['key'] == 'pc4' else 5], dtype=np.float64)
        out[f'case{i}_aug'] = aug
        out[f'case{i}_fog'] = np.zeros((0, aug.shape[1])) if fog is None else fog
        if info is not None:
            out[f'case{i}_info'] = np.array([info['min_fog_response'], info['max_fog_response'],
                                             info['num_fog_responses']], dtype=np.float64)
        # two more calls on the same generator: the stream position after a call is part of the contract
        out[f'case{i}_next_u'] = ref.RNG.random(2)
    # the reference's LUT key rule on a dense set of ranges
    r = np.linspace(0.0, 205.0, 8201).astype(np.float32)
    d = ref.get_integral_dict(ref.ParameterSet(alpha=0.06))
    keys = sorted(d.keys())
    idx = np.array([keys.index(min(float(str(round(x, 1))), 200)) for x in r], dtype=np.int32)
    out['key_r'] = r
    out['key_idx'] = idx
    path = os.path.join(ROOT, 'tests', 'golden', 'fog.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes,', len(cases), 'cases')


if __name__ == '__main__':
    main()
