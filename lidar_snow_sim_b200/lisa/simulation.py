"""
Drop-in mirror of the reference's LISA augmenter for its Monte-Carlo modes (lib/LISA/python/lisa.py:191-341; caller:
DenseDataset.__getitem__, lib/OpenPCDet/pcdet/datasets/dense/dense_dataset.py:713-746), backed by the CUDA engine:

    lisa = LISA(mode='gunn')                                     # 'rain' | 'gunn' | 'sekhon'; signal 'strongest' | 'last'
    after = lisa.augment(pc=before, Rr=rainfall_rate)            # (N, 4) float64 x, y, z, intensity in [0, 1]
                                                                 # -> (N, 6): x, y, z, intensity, label, intensity_diff
                                                                 # label 0 lost, 1 not scattered, 2 randomly scattered

Same constructor arguments and defaults, same `augment(pc, Rr, fixed_seed=False)` call and return layout.  The
extinction coefficient alpha(Rr) is integrated on the host from the tabulated Mie efficiencies exactly like LISA.alpha
(:468-482); the table is the reference's data file `mie_<refractive index>_λ_<wavelength>.npz` (arrays D, qext; pass its
path / directory as `mie_table`, or the arrays themselves) -- generating it needs PyMieScatt and is offline tooling.

Randomness: with `fixed_seed=True` (every return re-seeds NumPy's generator with 666, lisa.py:54-55) the device replays
NumPy's own draw sequence and reproduces the reference up to libm rounding.  Without it the reference is not
reproducible itself (a thread pool shares the global generator, :333-339); the device then uses a counter-based
generator seeded from NumPy's global state (so np.random.seed() still controls it): same distribution, different draws.

The fog / haze / spray modes of the reference (`average_augment`, `goodin_augment`) are not part of this path.
"""
import os
from pathlib import Path

import numpy as np
import torch

from .. import _lib
from ..engine import default_engine, _ptr

_MODES = {'rain': (0, 1.328), 'gunn': (1, 1.3031), 'sekhon': (2, 1.3031)}
_SEED = 666                                         # lisa.py:55


def _size_law(mode, Rr):
    """N0, Lambda of the exponential size distribution N(D) = N0 exp(-Lambda D) (lisa.py:497-664)."""
    if mode == 'rain':
        return 8000.0, 4.1 * Rr ** (-0.21)
    if mode == 'gunn':
        return 7.6e3 * Rr ** (-0.87), 2.55 * Rr ** (-0.48)
    return 5.0e3 * Rr ** (-0.94), 2.29 * Rr ** (-0.45)


class LISA:
    def __init__(self, wavelength: float = 905, r_min: float = 0.9, r_max: float = 120, beam_divergence: float = 3e-3,
                 min_diameter: float = 0.05, range_accuracy: float = 0.09, signal: str = 'strongest', mode: str = 'rain',
                 show_progressbar: bool = False, *, mie_table=None, engine=None) -> None:
        if mode not in _MODES:
            raise NotImplementedError(f"mode '{mode}': only the Monte-Carlo modes 'rain', 'gunn', 'sekhon' run on the engine")
        if signal not in ('strongest', 'last'):
            raise ValueError('Invalid lidar return mode')
        self.r_min, self.r_max, self.signal, self.atm_model = r_min, r_max, signal, mode
        self.wavelength, self.min_diameter = wavelength, min_diameter
        self.range_accuracy, self.beam_divergence = range_accuracy, beam_divergence
        self.show_progressbar = show_progressbar
        self.refractive_index = _MODES[mode][1]
        self.engine = engine
        self.D, self.qext = self._load_mie(mie_table)
        self._tables = {}

    def _load_mie(self, mie_table):
        if isinstance(mie_table, (tuple, list)):
            return np.asarray(mie_table[0], dtype=np.float64), np.asarray(mie_table[1], dtype=np.float64)
        name = f'mie_{self.refractive_index}_λ_{self.wavelength}.npz'
        cands = []
        if mie_table is not None:
            p = Path(mie_table)
            cands += [p, p / name]
        if os.environ.get('LSS_LISA_MIE_DIR'):
            cands.append(Path(os.environ['LSS_LISA_MIE_DIR']) / name)
        for c in cands:
            if c.is_file():
                dat = np.load(str(c))
                return np.asarray(dat['D'], dtype=np.float64), np.asarray(dat['qext'], dtype=np.float64)
        raise FileNotFoundError(f"Mie coefficient table '{name}' not found (pass mie_table=<path | directory | (D, qext)> or "
                                f"set LSS_LISA_MIE_DIR; the reference ships it in lib/LISA/python/)")

    # ---- the reference's helpers the callers use (pointcloud_viewer.py:2794-2796) ------------------------------------
    def Nd(self, D, Rr):
        n0, lam = _size_law(self.atm_model, Rr)
        return n0 * np.exp(-lam * D)

    def alpha(self, curve):
        """lisa.py:468-482"""
        curve = np.asarray(curve)
        if curve.size == 1:
            return 0.01 * curve ** 0.6
        return 1e-6 * np.trapezoid(self.D ** 2 * self.qext * curve, self.D) * np.pi / 4

    def density(self, Rr, dstart):
        n0, lam = _size_law(self.atm_model, Rr)
        return n0 * np.exp(-lam * dstart) / lam

    # ---- augment --------------------------------------------------------------------------------------------------------
    def _draw_table(self, engine, n_draws):
        key = (id(engine), 'fixed')
        t = self._tables.get(key)
        if t is None or t.numel() < n_draws:
            n = max(n_draws, 1 << 14)
            host = np.random.RandomState(_SEED).random_sample(n)       # == the first n doubles after np.random.seed(666)
            t = torch.from_numpy(host).to(engine.device)
            self._tables[key] = t
        return t

    def augment(self, pc: np.ndarray, Rr, fixed_seed: bool = False) -> np.ndarray:
        """LISA.monte_carlo_augment (lisa.py:293-341)."""
        engine = self.engine or default_engine()
        lib = engine.lib
        pc = np.ascontiguousarray(pc, dtype=np.float64)
        n, F = pc.shape
        if F < 4:
            raise ValueError('pc must be (N, >= 4): x, y, z, intensity')
        Rr = float(Rr)
        a = float(self.alpha(self.Nd(self.D, Rr)))
        d_pc = torch.from_numpy(pc).to(engine.device)
        out = torch.empty((n, F + 2), dtype=torch.float64, device=engine.device)
        seed = 0
        if not fixed_seed:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        # upper bound on the draws of one return: 1 + ranges + diameters + the Gaussian's rejection pairs
        half = 1e-3 * (1e3 * np.tan(self.beam_divergence) * self.r_max) / 2
        r_far = float(np.sqrt((pc[:, :3] ** 2).sum(axis=1).max())) if n else 0.0
        n_max = self.density(Rr, self.min_diameter) * (np.pi / 3) * max(r_far, 1.0) * (half * max(r_far, 1.0) / self.r_max) ** 2
        need = int(2 * (n_max + 2) + 128)
        while True:
            table = self._draw_table(engine, need) if fixed_seed else None
            with torch.cuda.device(engine.device):
                st = lib.lss_lisa_batch(engine.h, _ptr(d_pc), F, n, Rr, _MODES[self.atm_model][0], a, float(self.r_min),
                                        float(self.r_max), float(self.beam_divergence), float(self.min_diameter),
                                        float(self.range_accuracy), 1 if self.signal == 'last' else 0, _ptr(table),
                                        0 if table is None else int(table.numel()), seed, _ptr(out), engine._stream())
            _lib.check(st, engine.h)
            try:
                engine.check()
                break
            except RuntimeError:
                if not fixed_seed or need > (1 << 26):
                    raise
                need *= 4                                              # the draw table was too short for some return
        return out.cpu().numpy()
