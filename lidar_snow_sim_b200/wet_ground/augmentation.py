"""
Drop-in mirror of the reference's wet-ground augmentation (tools/wet_ground/augmentation.py:25-161), backed by the
CUDA engine.  Same name, arguments and return value:

    out = ground_water_augmentation(pointcloud, water_height=0.001, pavement_depth=0.0012, noise_floor=0.7,
                                    power_factor=15, estimation_method='linear', flat_earth=False, debug=True,
                                    delta=0.5, replace=True)

    out: float64 (N'', 5): all non-ground rows first (unchanged), then the kept ground rows with their new intensity;
         column 4 is 0 (if `replace`) / the input value for non-ground rows and 1 for the kept ground rows.
         With fewer than 1000 ground points the INPUT array is returned unchanged (augmentation.py:51-52).

Notes: `debug` is accepted and ignored (it only draws matplotlib plots); estimation_method='poly' (a RANSAC polyfit
on np.random, augmentation.py:171-192,223-228,243-246) is not implemented; coordinates are processed as float32
(STF clouds are float32 on disk, precompute.py:78).  Keyword-only extras: `engine`, `plane`, `ymins` (the reference
host's RANSAC plane and np.argpartition picks, replayed for parity tests).
"""
import numpy as np
import torch

from ..engine import default_engine


def ground_water_augmentation(pointcloud, water_height=0.001, pavement_depth=0.0012, noise_floor=0.7, power_factor=15,
                              estimation_method='linear', flat_earth=False, debug=True,
                              delta=0.5, replace=True, *, engine=None, plane=None, ymins=None, return_internals=False):
    if estimation_method != 'linear':
        raise NotImplementedError("only estimation_method='linear' is implemented")
    if not isinstance(flat_earth, (bool, np.bool_)):
        assert False, 'flat earth tag has be bool'                      # augmentation.py:64-65
    engine = engine or default_engine()
    pc32 = np.ascontiguousarray(pointcloud[:, :5], dtype=np.float32)
    n = pc32.shape[0]
    d_pc = torch.from_numpy(pc32).to(engine.device)
    pl = None if plane is None else np.asarray([[plane[0][0], plane[0][1], plane[0][2], plane[1]]], dtype=np.float64)
    res = engine.wet_ground_batch(d_pc, np.array([0, n], dtype=np.int64), None, water_height, pavement_depth,
                                  noise_floor, power_factor, bool(flat_earth), delta, bool(replace), plane=pl,
                                  want_intensity64=True,
                                  ymins=None if ymins is None else np.asarray(ymins, dtype=np.int32).reshape(1, 50))
    engine.check()
    if int(res['passthrough'][0].item()):
        return (pointcloud, dict(passthrough=True)) if return_internals else pointcloud
    cnt = int(res['counts'][0].item())
    out = res['points'][:cnt].cpu().numpy().astype(np.float64)
    out[:, 3] = res['intensity64'][:cnt].cpu().numpy()
    if return_internals:
        return out, dict(passthrough=False, plane=res['plane'][0].cpu().numpy())
    return out
