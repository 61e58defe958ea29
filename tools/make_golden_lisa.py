"""
Golden vectors of the LISA Monte-Carlo augmenter: runs the UNMODIFIED reference (lib/LISA/python/lisa.py, imported with two
in-memory shims: a PyMieScatt stub -- only called when the Mie table file is missing, it is not -- and
scipy.integrate.trapz -> numpy.trapezoid for SciPy >= 1.14) with fixed_seed=True, return by return in one thread (the
reference's own ThreadPool shares NumPy's global generator between threads), checks the oracle restatement
(oracle/lisa.py) bit for bit and freezes inputs / outputs in tests/golden/lisa.npz.

    python tools/make_golden_lisa.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference/lib/LISA/python'


def load_reference():
    sys.dont_write_bytecode = True
    sys.modules.setdefault('PyMieScatt', types.ModuleType('PyMieScatt'))
    import scipy.integrate as si
    if not hasattr(si, 'trapz'):
        si.trapz = np.trapezoid
    sys.path.insert(0, REF)
    import lisa
    return lisa


def main():
    from oracle import lisa as ol
    from lidar_snow_sim_b200.synthetic import synthetic_cloud
    lisa = load_reference()
    pc32 = synthetic_cloud(seed=21, n_azimuth=24, shuffle_rows=True)
    pc = np.zeros((pc32.shape[0], 4))
    pc[:, :3] = pc32[:, :3]
    pc[:, 3] = pc32[:, 3] / 255                                        # dense_dataset.py:732-734
    pc[:4, :3] *= 0.01                                                 # returns inside r_min
    pc[4, :3] = 0.0                                                    # r == 0
    ok = True
    rec = {'points': pc}
    cases = [('gunn', 34.97475775452152, 'strongest'), ('rain', 20.0, 'strongest'), ('sekhon', 70.78393287483148, 'last'),
             ('gunn', 200.20719573938692, 'last'), ('rain', 2.0, 'strongest')]
    for ci, (mode, Rr, signal) in enumerate(cases):
        L = lisa.LISA(mode=mode, signal=signal)
        a = L.alpha(L.Nd(L.D, Rr))
        ref = np.array([lisa.multi_lisa(Rr, True, L.r_min, L.r_max, L.beam_divergence, L.min_diameter, L.refractive_index,
                                        L.range_accuracy, a, L.signal, L.density, L.diameters, tuple(p)) for p in pc])
        got = ol.monte_carlo_augment(pc, Rr, mode, ol.alpha(mode, Rr, L.D, L.qext), signal=signal)
        good = np.array_equal(ref, got)
        print(f'{mode} Rr={Rr:.3f} {signal}: oracle==reference {good}; labels', [(ref[:, 4] == l).sum() for l in (0, 1, 2)])
        ok &= good
        rec[f'c{ci}_mode'] = mode
        rec[f'c{ci}_Rr'] = Rr
        rec[f'c{ci}_signal'] = signal
        rec[f'c{ci}_alpha'] = float(a)
        rec[f'c{ci}_out'] = ref
        if ci == 0:
            rec['D'] = L.D
            rec['qext_ice'] = L.qext
        if mode == 'rain' and 'qext_water' not in rec:
            rec['qext_water'] = L.qext
    rec['n_cases'] = len(cases)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'lisa.npz'), **rec)
    print('ALL OK' if ok else 'MISMATCH')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
