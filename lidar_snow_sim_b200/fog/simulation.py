"""
Drop-in mirror of the reference's fog simulation API (lib/LiDAR_fog_sim/fog_simulation.py): `ParameterSet` (:52-171) and
`simulate_fog(p, pc, noise, gain, noise_variant, hard, soft)` (:299-316), computed by the B200 engine (csrc/fog.cu).
Caller in the reference: DenseDataset.foggify, lib/OpenPCDet/pcdet/datasets/dense/dense_dataset.py:990-1009.

Like the reference, the noise draws come from a module-level `RNG = np.random.default_rng(seed=42)` (:15) unless the
caller passes `rng=`; the generator's stream is consumed exactly as the reference consumes it (one `integers` draw per
call, then one draw per fog point in point order), so a run is reproducible against the reference draw for draw.

The integral look-up tables are the reference's data files (`integral_lookup_tables/original/*.pickle`, 1.7 MB, :19);
point `LSS_FOG_LUT_DIR` (or `lut_dir=`) at that directory, or pass `lut=` a (2001, 2) float64 array directly.
"""
import math
import os
import pickle
from pathlib import Path

import numpy as np
import torch

from ..engine import default_engine

speed_of_light = 299792458.0            # scipy.constants.speed_of_light

RNG = np.random.default_rng(seed=42)    # fog_simulation.py:15

AVAILABLE_TAU_Hs = [20]                 # :17

_PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645
_MASK128 = (1 << 128) - 1


class ParameterSet:
    """fog_simulation.py:52-171: same fields, defaults and derivation order (keyword overrides are applied LAST, so
    `ParameterSet(alpha=...)` keeps the default-derived `mor` / `beta`, exactly like the reference)."""

    def __init__(self, **kwargs) -> None:
        self.n = 500
        self.n_min = 100
        self.n_max = 1000
        self.r_range = 100
        self.r_range_min = 50
        self.r_range_max = 250
        # soft target a.k.a. fog
        self.alpha = 0.06                               # attenuation coefficient
        self.alpha_min = 0.003
        self.alpha_max = 0.5
        self.alpha_scale = 1000
        self.mor = np.log(20) / self.alpha              # meteorological optical range (m)
        self.beta = 0.046 / self.mor                    # backscattering coefficient (1/sr)
        self.beta_min = 0.023 / self.mor
        self.beta_max = 0.092 / self.mor
        self.beta_scale = 1000 * self.mor
        # sensor
        self.p_0 = 80                                   # pulse peak power (W)
        self.p_0_min = 60
        self.p_0_max = 100
        self.tau_h = 2e-8                               # half-power pulse width (s)
        self.tau_h_min = 5e-9
        self.tau_h_max = 8e-8
        self.tau_h_scale = 1e9
        self.e_p = self.p_0 * self.tau_h                # total pulse energy (J)
        self.a_r = 0.25                                 # receiver aperture area (m^2)
        self.a_r_min = 0.01
        self.a_r_max = 0.1
        self.a_r_scale = 1000
        self.l_r = 0.05                                 # loss of the receiver's optics
        self.l_r_min = 0.01
        self.l_r_max = 0.10
        self.l_r_scale = 100
        self.c_a = speed_of_light * self.l_r * self.a_r / 2
        self.linear_xsi = True
        self.D = 0.1
        self.ROH_T = 0.01
        self.ROH_R = 0.01
        self.GAMMA_T_DEG = 2
        self.GAMMA_R_DEG = 3.5
        self.GAMMA_T = math.radians(self.GAMMA_T_DEG)
        self.GAMMA_R = math.radians(self.GAMMA_R_DEG)
        self.r_1 = 0.9
        self.r_1_min = 0
        self.r_1_max = 10
        self.r_1_scale = 10
        self.r_2 = 1.0
        self.r_2_min = 0
        self.r_2_max = 10
        self.r_2_scale = 10
        # hard target
        self.r_0 = 30
        self.r_0_min = 1
        self.r_0_max = 200
        self.gamma = 0.000001                           # reflectivity of the hard target
        self.gamma_min = 0.0000001
        self.gamma_max = 0.00001
        self.gamma_scale = 10000000
        self.beta_0 = self.gamma / np.pi                # differential reflectivity of the target
        self.__dict__.update(kwargs)


def _lut_dir(lut_dir=None):
    d = lut_dir or os.environ.get('LSS_FOG_LUT_DIR')
    if d is None:
        raise FileNotFoundError('integral look-up tables: set LSS_FOG_LUT_DIR (or pass lut_dir= / lut=) to the '
                                "reference's lib/LiDAR_fog_sim/integral_lookup_tables/original directory")
    return Path(d)


def get_available_alphas(lut_dir=None):
    """fog_simulation.py:38-50"""
    alphas = []
    for file in os.listdir(_lut_dir(lut_dir)):
        if file.endswith('.pickle'):
            alphas.append(float(file.split('_')[-1].replace('.pickle', '')))
    return sorted(alphas)


def load_integral_table(p, lut_dir=None):
    """get_integral_dict (fog_simulation.py:174-180) as a (2001, 2) float64 array (fog_distance, fog_response): the
    table of the available alpha nearest to p.alpha; row k is the reference's dictionary entry for key k / 10."""
    alphas = get_available_alphas(lut_dir)
    alpha = min(alphas, key=lambda x: abs(x - p.alpha))
    tau_h = min(AVAILABLE_TAU_Hs, key=lambda x: abs(x - int(p.tau_h * 1e9)))
    filename = _lut_dir(lut_dir) / f'integral_0m_to_200m_stepsize_0.1m_tau_h_{tau_h}ns_alpha_{alpha}.pickle'
    with open(filename, 'rb') as handle:
        integral_dict = pickle.load(handle)
    keys = sorted(integral_dict.keys())
    if len(keys) != 2001:
        raise ValueError(f'{filename}: expected 2001 entries (0 .. 200 m in 0.1 m steps), found {len(keys)}')
    return np.array([[float(integral_dict[k][0]), float(integral_dict[k][1])] for k in keys], dtype=np.float64)


def _pcg64_state(rng):
    st = rng.bit_generator.state
    if st['bit_generator'] != 'PCG64':
        raise TypeError('simulate_fog needs a numpy Generator on PCG64 (np.random.default_rng)')
    s, inc = st['state']['state'], st['state']['inc']
    return np.array([s >> 64, s & (2 ** 64 - 1), inc >> 64, inc & (2 ** 64 - 1)], dtype=np.uint64)


def _pcg64_advance(rng, delta):
    """Advance the generator by `delta` 64-bit outputs, keeping its buffered 32-bit half (Generator.integers) intact --
    PCG64.advance() would drop it and a later `integers` call would leave the reference's stream."""
    st = rng.bit_generator.state
    s, inc = st['state']['state'], st['state']['inc']
    acc_mult, acc_plus, cur_mult, cur_plus = 1, 0, _PCG_MULT, inc
    while delta > 0:
        if delta & 1:
            acc_mult = (acc_mult * cur_mult) & _MASK128
            acc_plus = (acc_plus * cur_mult + cur_plus) & _MASK128
        cur_plus = ((cur_mult + 1) * cur_plus) & _MASK128
        cur_mult = (cur_mult * cur_mult) & _MASK128
        delta >>= 1
    st['state']['state'] = (acc_mult * s + acc_plus) & _MASK128
    rng.bit_generator.state = st


def simulate_fog(p, pc, noise, gain=False, noise_variant='v1', hard=True, soft=True, *, engine=None, lut=None,
                 lut_dir=None, rng=None):
    """
    fog_simulation.py:299-316.  pc: (N, F >= 4) array (x, y, z, intensity, ...).  Returns
    (augmented_pc, simulated_fog_pc or None, info_dict or None) with the reference's dtypes: float64 (N, F) when `soft`,
    the input's float32 when only `hard`.
    """
    variants = {'v1': 1, 'v2': 2, 'v3': 3, 'v4': 4}
    if soft and noise > 0 and noise_variant not in variants:
        raise NotImplementedError(f"noise variant '{noise_variant}' is not implemented (yet)")      # :264-266
    eng = engine or default_engine()
    rng = RNG if rng is None else rng
    pc32 = np.ascontiguousarray(pc, dtype=np.float32)
    N, F = pc32.shape
    off = np.array([0, N], dtype=np.int64)
    d_pts = torch.from_numpy(pc32).to(eng.device)
    d_lut = None
    if soft:
        table = load_integral_table(p, lut_dir) if lut is None else np.ascontiguousarray(lut, dtype=np.float64)
        d_lut = torch.from_numpy(table).to(eng.device)
        rng.integers(low=1, high=20, size=1)                # :207 (the value is overwritten by 10 at :208)
    variant = variants.get(noise_variant, 1)
    kw = dict(hard=hard, soft=soft, gain=gain, noise=int(noise), noise_variant=variant)
    draws = soft and noise > 0
    if draws and variant == 4:
        # Generator.beta is rejection sampling (no jump-ahead): first pass for ranks and the count, draw, second pass
        first = eng.fog_batch(d_pts, off, d_lut, p.alpha, p.beta, p.beta_0, **dict(kw, noise=0))
        cnt = int(first['info'][0, 2].item())
        ext = torch.zeros((max(N, 1),), dtype=torch.float64)
        if cnt:
            ext[:cnt] = torch.from_numpy(rng.beta(a=2, b=20, size=cnt))
        res = eng.fog_batch(d_pts, off, d_lut, p.alpha, p.beta, p.beta_0, ext_noise=ext.to(eng.device), **kw)
    else:
        res = eng.fog_batch(d_pts, off, d_lut, p.alpha, p.beta, p.beta_0,
                            rng_states=_pcg64_state(rng)[None] if draws else None, **kw)
    eng.check()
    aug = res['points'].cpu().numpy()
    if not soft:
        return aug.astype(np.float32), None, None           # P_R_fog_hard keeps the input's dtype (:183-189)
    info = res['info'].cpu().numpy()[0]
    cnt = int(info[2])
    if draws and variant != 4 and cnt:
        _pcg64_advance(rng, cnt)
    mask = res['fog_mask'].cpu().numpy().astype(bool)
    simulated_fog_pc = aug[mask] if cnt > 0 else None       # :287-291
    info_dict = {'min_fog_response': float(info[0]) if cnt else np.inf,
                 'max_fog_response': float(info[1]) if cnt else 0,
                 'num_fog_responses': cnt}
    return aug, simulated_fog_pc, info_dict
