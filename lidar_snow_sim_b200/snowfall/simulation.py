"""
Drop-in mirror of the reference's snowfall augmentation API (tools/snowfall/simulation.py:427-544), backed by the
CUDA engine.  Same name, same positional arguments, same return value and exception types:

    stats, aug_pc = augment(pc, particle_file_prefix, beam_divergence, shuffle=True, show_progressbar=False,
                            only_camera_fov=True, noise_floor=0.7, root_path=None)

    stats  = (num_attenuated, num_removed, avg_intensity_diff)
    aug_pc = float32 (N', 5): x, y, z, intensity, label (0 untouched, 1 attenuated, 2 scattered), sorted by channel

Differences a caller can observe, all documented in DESIGN.md:
  * rows of one channel keep their input order (stable sort); the reference's argsort order within a channel is
    implementation-defined (simulation.py:447);
  * `show_progressbar` is accepted and ignored (it only selects the reference's process pool);
  * keyword-only extras (`engine`, `tables`, `order`, `plane`, `ymins`, `thresh_poly`, `theta`) let tests inject what
    the reference draws from global state (random.shuffle, RANSAC) or computes host-dependently (float32 arctan2,
    np.argpartition's pick among the least populated histogram bins).
"""
import os
import random
from pathlib import Path
from typing import Tuple

import numpy as np
import torch

from ..engine import default_engine, DEFAULT_MAX_DIVERGENCE_RAD

_table_cache = {}


def _load_tables(engine, particle_file_prefix: str, root_path, max_div_rad: float):
    """Particle-table lookup of simulation.py:324-329, cached per (engine, directory, prefix)."""
    if root_path:
        base = Path(root_path) / 'training' / 'snowflakes' / 'npy'
    else:
        base = Path(os.environ.get('LSS_NPY_DIR', Path(__file__).parent.parent.parent.absolute() / 'npy'))
    key = (id(engine), str(base), particle_file_prefix)
    hit = _table_cache.get(key)
    if hit is not None and hit[1] >= max_div_rad:
        return hit[0]
    tables = []
    for k in range(1, 65):
        path = base / f'{particle_file_prefix}_{k}.npy'
        if not path.is_file():
            raise FileNotFoundError(f"[Errno 2] No such file or directory: '{path}'")
        tables.append(np.load(str(path)))
    if hit is not None:
        engine.free_tables(hit[0])
    build_div = max(max_div_rad, DEFAULT_MAX_DIVERGENCE_RAD)
    tid = engine.upload_tables(tables, max_beam_divergence_rad=build_div)
    _table_cache[key] = (tid, build_div)
    return tid


def augment(pc: np.ndarray, particle_file_prefix: str, beam_divergence: float, shuffle: bool = True,
            show_progressbar: bool = False, only_camera_fov: bool = True, noise_floor: float = 0.7,
            root_path: str = None, *, engine=None, tables=None, order=None, plane=None, ymins=None, thresh_poly=None,
            theta=None, return_internals: bool = False) -> Tuple:
    """
    :param pc:                      N-by-5 array containing original pointcloud (x, y, z, intensity, channel).
    :param particle_file_prefix:    Prefix of the files where sampled particles are stored (x, y, r).
    :param beam_divergence:         Beam divergence in degrees.
    :param shuffle:                 Flag if order of sampled snowflakes should be shuffled.
    :param show_progressbar:        Ignored (kept for signature compatibility).
    :param only_camera_fov:         Flag if the camera field of view (FOV) filter should be applied.
    :param noise_floor:             Noise floor threshold.
    :param root_path:               Optional root path of '<root>/training/snowflakes/npy'.
    :return:                        ((num_attenuated, num_removed, avg_intensity_diff), augmented pointcloud)
    """
    engine = engine or default_engine()
    max_div = float(np.radians(beam_divergence))
    own_tables = tables is not None and not isinstance(tables, int)
    if tables is not None:
        table_id = tables if isinstance(tables, int) else engine.upload_tables(
            tables, max_beam_divergence_rad=max(max_div, DEFAULT_MAX_DIVERGENCE_RAD))
    else:
        table_id = _load_tables(engine, particle_file_prefix, root_path, max_div)
    try:
        return _augment_uploaded(engine, table_id, pc, beam_divergence, shuffle, only_camera_fov, noise_floor, order,
                                 plane, ymins, thresh_poly, theta, return_internals)
    finally:
        if own_tables:
            engine.free_tables(table_id)


def _augment_uploaded(engine, table_id, pc, beam_divergence, shuffle, only_camera_fov, noise_floor, order, plane, ymins,
                      thresh_poly, theta, return_internals):

    if order is None:
        order = list(range(64))
        if shuffle:
            random.shuffle(order)                       # simulation.py:485-486 (python's global RNG, like the reference)

    pc = np.ascontiguousarray(pc, dtype=np.float32)
    n = pc.shape[0]
    dev = engine.device
    d_pc = torch.from_numpy(pc).to(dev, non_blocking=False)
    d_theta = None
    if theta is not None:
        d_theta = torch.from_numpy(np.ascontiguousarray(theta, dtype=np.float32)).to(dev)
    off = np.array([0, n], dtype=np.int64)
    pl = None
    if plane is not None and thresh_poly is None:       # (w, h) as calculate_plane returns it (planes.py:50)
        pl = np.array([[plane[0][0], plane[0][1], plane[0][2], plane[1]]], dtype=np.float64)
    ym = None if (ymins is None or thresh_poly is not None) else np.asarray(ymins, dtype=np.int32).reshape(1, 50)
    res = engine.snowfall_batch(table_id, d_pc, off, np.asarray(order, dtype=np.int32)[None, :], float(beam_divergence),
                                theta=d_theta, thresh_poly=thresh_poly, plane=pl, ymins=ym, noise_floor=noise_floor,
                                threshold_filter=True, camera_fov=bool(only_camera_fov),
                                device_prepass=thresh_poly is None, want_full=return_internals,
                                want_perm=return_internals, want_nocc=return_internals)
    engine.check()                                      # synchronises; raises IndexError / AssertionError like the reference
    count = int(res['counts'][0].item())
    st = res['stats'][0].cpu().numpy()
    aug_pc = res['points'][:count].cpu().numpy()
    stats = (int(st[0]), int(st[1]), int(st[2]))
    if return_internals:
        return stats, aug_pc, dict(order=list(order), full=res['full'].cpu().numpy(), perm=res['perm'].cpu().numpy(),
                                   n_occluders=res['nocc'].cpu().numpy(), intensity_diff_sum=float(st[3]))
    return stats, aug_pc
