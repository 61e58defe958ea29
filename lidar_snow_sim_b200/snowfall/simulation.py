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
from .sampling import table_dir, load_table_set, load_or_sample_table_set, particle_file_prefix

_table_cache = {}


def _load_tables(engine, particle_file_prefix: str, root_path, max_div_rad: float, tables=None):
    """Particle-table lookup of simulation.py:324-329, uploaded once and cached per (engine, directory, prefix)."""
    base = table_dir(root_path)
    key = (id(engine), str(base), particle_file_prefix)
    hit = _table_cache.get(key)
    if hit is not None and hit[1] >= max_div_rad:
        return hit[0]
    if tables is None:
        tables = load_table_set(particle_file_prefix, base)
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


def augment_snowfall(pc: np.ndarray, snowfall_rate: float, terminal_velocity: float = 1.6, mode: str = 'gunn',
                     beam_divergence: float = float(np.degrees(3e-3)), shuffle: bool = True,
                     show_progressbar: bool = False, only_camera_fov: bool = True, noise_floor: float = 0.7,
                     root_path: str = None, *, engine=None, write_tables: bool = False, table_seed: int = 1000,
                     **extras) -> Tuple:
    """
    augment() addressed by the physical parameters instead of a particle-file prefix:

        stats, aug_pc = augment_snowfall(pc, snowfall_rate=2.5, terminal_velocity=1.6, mode='gunn')

    Derives the prefix '<mode>_<rain_rate>_<occupancy>' exactly like the reference's callers do
    (tools/snowfall/precompute.py:57-58,101, pointcloud_viewer.py:2798-2802), reads '<prefix>_<1..64>.npy' from the
    directory augment() would read them from if they are all there, and otherwise draws the 64 planes with the engine's
    native dart-throwing sampler (no 2.3 GB download; `write_tables=True` stores them under the reference's file names,
    sampling.py:344, so that the next process -- or the reference itself -- finds them).  The table set is uploaded once
    per process and engine.  Everything else is augment(): same defaults (beam_divergence = degrees(3e-3),
    precompute.py:104), same return value, same exceptions; `extras` are augment()'s keyword-only test hooks.
    """
    engine = engine or default_engine()
    prefix = particle_file_prefix(mode, snowfall_rate, terminal_velocity)
    key = (id(engine), str(table_dir(root_path)), prefix)
    max_div = float(np.radians(beam_divergence))
    hit = _table_cache.get(key)
    if hit is None or hit[1] < max_div:
        tables, _, _ = load_or_sample_table_set(mode, snowfall_rate, terminal_velocity, root_path=root_path,
                                                write=write_tables, seed=table_seed)
        _load_tables(engine, prefix, root_path, max_div, tables=tables)
    return augment(pc, prefix, beam_divergence, shuffle=shuffle, show_progressbar=show_progressbar,
                   only_camera_fov=only_camera_fov, noise_floor=noise_floor, root_path=root_path, engine=engine, **extras)
