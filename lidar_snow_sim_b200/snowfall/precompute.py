"""
Offline batch driver with the reference's layout and skip-if-exists semantics (tools/snowfall/precompute.py:47-106):

    for mode in ('gunn', 'sekhon'):
        for every sample id of the split:
            read   <lidar_folder>/<id>.bin                                  (N x 5 float32, :78)
            keep   the points inside the camera field of view                (:96-99)
            for the five (snowfall_rate, terminal_velocity) combos (:20-21):
                write <lidar_folder>/../snowfall_simulation/<mode>/<lidar_folder.name>_rainrate_<int(rain)>/<id>.bin
                      = augment(pc, '<mode>_<rain>_<occupancy>', beam_divergence=degrees(3e-3))      (:101-106)

Differences: frames of one (mode, combo) go through the engine in batches (`batch_frames` clouds per call, pinned
host -> chunked H2D / kernels / D2H pipeline); the snowflake tables are read from `npy_root` like the reference does,
or -- `sample_tables=True` -- drawn on the fly by the engine's dart-throwing sampler (no 2.3 GB download).
"""
import random
from pathlib import Path

import numpy as np
import torch

from ..calib.dense_camera import STF_HDL64_CAMERA
from ..engine import default_engine, DEFAULT_MAX_DIVERGENCE_RAD
from .sampling import compute_occupancy, snowfall_rate_to_rainfall_rate, sample_table_set

SNOWFALL_RATES = [0.5, 1.0, 2.0, 2.5, 1.5]       # mm/h   (precompute.py:20)
TERMINAL_VELOCITIES = [2.0, 1.6, 2.0, 1.6, 0.6]  # m/s    (precompute.py:21)


def get_fov_flag(points_xyz, camera=STF_HDL64_CAMERA):
    """Host-side camera FOV mask of the input frames, as precompute.py:96-98 (calibration_kitti.py:65-84)."""
    P2, R0, V2C = camera['P2'], camera['R0'], camera['V2C']
    h, w = camera.get('img_shape', (1024, 1920))
    pts = np.asarray(points_xyz, dtype=np.float32)
    hom = np.hstack((pts, np.ones((pts.shape[0], 1), dtype=np.float32)))
    rect = np.dot(hom, np.dot(V2C.T, R0.T))
    rect_hom = np.hstack((rect, np.ones((rect.shape[0], 1), dtype=np.float32)))
    img = np.dot(rect_hom, P2.T)
    uv = (img[:, 0:2].T / rect_hom[:, 2]).T
    depth = img[:, 2] - P2.T[3, 2]
    return (uv[:, 0] >= 0) & (uv[:, 0] < w) & (uv[:, 1] >= 0) & (uv[:, 1] < h) & (depth >= 0)


def combos():
    out = []
    for rs, tv in zip(SNOWFALL_RATES, TERMINAL_VELOCITIES):
        out.append((rs, tv, float(snowfall_rate_to_rainfall_rate(rs, tv)), compute_occupancy(rs, tv)))
    return out


def _load_tables(mode, rain, occ, npy_root):
    tabs = []
    for k in range(1, 65):
        path = Path(npy_root) / f'{mode}_{rain}_{occ}_{k}.npy'
        if not path.is_file():
            raise FileNotFoundError(f"[Errno 2] No such file or directory: '{path}'")
        tabs.append(np.load(str(path)))
    return tabs


def precompute(sample_ids, lidar_folder, modes=('gunn', 'sekhon'), npy_root=None, sample_tables=False, table_seed=42,
               engine=None, batch_frames=32, only_camera_fov=True, shuffle=True, camera=STF_HDL64_CAMERA,
               progress=None):
    """Returns the number of files written.  Existing outputs are skipped (precompute.py:91-92)."""
    engine = engine or default_engine()
    lidar_folder = Path(lidar_folder)
    div_deg = float(np.degrees(3e-3))                                # precompute.py:104
    written = 0
    for mode in modes:
        for (rs, tv, rain, occ) in combos():
            save_dir = lidar_folder.parent / 'snowfall_simulation' / mode / f'{lidar_folder.name}_rainrate_{int(rain)}'
            save_dir.mkdir(parents=True, exist_ok=True)
            todo = [s for s in sample_ids if not (save_dir / f'{s}.bin').is_file()]
            if not todo:
                continue
            tabs = sample_table_set(mode, rs, tv, seed=table_seed) if sample_tables else _load_tables(mode, rain, occ, npy_root)
            tid = engine.upload_tables(tabs, max_beam_divergence_rad=DEFAULT_MAX_DIVERGENCE_RAD)
            def finish(job):
                # wait for a submitted batch and write its files (precompute.py:105-106)
                ticket, ids_j, off_j = job
                res = engine.snowfall_batch_host_wait(ticket)
                out = res['points'].numpy()
                for b, s in enumerate(ids_j):
                    out[off_j[b]:off_j[b] + int(res['counts'][b])].astype(np.float32).tofile(str(save_dir / f'{s}.bin'))
                if progress:
                    progress(mode, rain, len(ids_j))
                return len(ids_j)

            pending = in_flight = None
            try:
                # double buffered: while the GPU works on batch k, the host reads the files of batch k+1 and writes
                # those of batch k-1
                for i0 in range(0, len(todo), batch_frames):
                    ids = todo[i0:i0 + batch_frames]
                    clouds = []
                    for s in ids:
                        pts = np.fromfile(str(lidar_folder / f'{s}.bin'), dtype=np.float32).reshape(-1, 5)
                        clouds.append(np.ascontiguousarray(pts[get_fov_flag(pts[:, 0:3], camera)]))
                    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
                    orders = []
                    for _ in ids:
                        o = list(range(64))
                        if shuffle:
                            random.shuffle(o)                        # simulation.py:485-486
                        orders.append(o)
                    host = torch.from_numpy(np.concatenate(clouds)).pin_memory()
                    ticket = engine.snowfall_batch_host_submit(tid, host, off, np.asarray(orders, dtype=np.int32), div_deg,
                                                               device_prepass=True, camera_fov=only_camera_fov,
                                                               n_chunks=min(4, len(ids)))
                    # the new ticket is tracked BEFORE the previous batch is finished: if writing that one raises, the
                    # `finally` below still waits for both before the tables are freed
                    previous, pending = pending, (ticket, ids, off)
                    if previous is not None:
                        in_flight = previous
                        written += finish(previous)
                        in_flight = None
                if pending is not None:
                    job, pending = pending, None
                    in_flight = job
                    written += finish(job)
                    in_flight = None
            finally:
                for job in (in_flight, pending):                     # an exception above: leave no batch in flight
                    if job is not None:
                        try:
                            engine.snowfall_batch_host_wait(job[0])
                        except Exception:
                            pass
                engine.free_tables(tid)
    return written
