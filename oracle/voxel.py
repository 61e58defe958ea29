"""
CPU oracle of the detector-input stage that follows the augmentation in the reference's data path -- TEST INFRASTRUCTURE,
NOT PRODUCT CODE (only tests/ may import it).

Restated:
  * mask_points_by_range                lib/OpenPCDet/pcdet/utils/common_utils.py:60-63 (x / y only, both ends inclusive),
    called by DataProcessor.mask_points_and_boxes_outside_range, lib/OpenPCDet/pcdet/datasets/processor/data_processor.py:78-91
  * DataProcessor.transform_points_to_voxels          data_processor.py:115-143, which delegates to
    VoxelGeneratorWrapper (data_processor.py:15-58) -> spconv's point-to-voxel generator.

PARITY UNPINNED for the voxel rule: spconv is a third-party dependency that is NOT vendored in the reference tree
(`pip install spconv-cu113`, README.md:139-140: spconv 2.x, `Point2VoxelCPU3d`; the wrapper also accepts spconv 1.x
`VoxelGeneratorV2`) and is not installed in this image, so it cannot be run here.  What is restated is its published
algorithm (spconv 1.2.1 src/spconv/.../points_to_voxel_3d_np, the same rule Point2VoxelCPU3d implements), in float32
like the C++ template instantiated for float32 points:

    grid_size[j] = round((range[3 + j] - range[j]) / voxel_size[j])
    for every point i, in order:
        c[j] = floor((p[i, j] - range[j]) / voxel_size[j])  for j = x, y, z        (float32 arithmetic)
        skip the point if any c[j] < 0 or c[j] >= grid_size[j]
        voxel = the voxel with coordinate (c_z, c_y, c_x); a NEW voxel gets the next index, unless max_voxels voxels
                exist already, in which case the point is skipped
        if the voxel holds fewer than max_points points, the point is appended to it

Outputs as the reference consumes them (data_processor.py:133-142): voxels (M, max_points, F) float32 zero padded,
coordinates (M, 3) int32 in (z, y, x) order, num_points_per_voxel (M,) int32; voxels are numbered by first appearance.
"""
import numpy as np


def mask_points_by_range(points, limit_range):
    """common_utils.py:60-63"""
    limit_range = np.asarray(limit_range, dtype=np.float32)
    return (points[:, 0] >= limit_range[0]) & (points[:, 0] <= limit_range[3]) \
        & (points[:, 1] >= limit_range[1]) & (points[:, 1] <= limit_range[4])


def grid_size(point_cloud_range, voxel_size):
    """data_processor.py:117-118: np.round((range[3:6] - range[0:3]) / voxel_size).astype(int64) (float32 inputs)."""
    r = np.asarray(point_cloud_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    return np.round((r[3:6] - r[0:3]) / v).astype(np.int64)


def points_to_voxels(points, point_cloud_range, voxel_size, max_points, max_voxels):
    """spconv's point-to-voxel rule (see the module docstring), vectorised where the rule allows it."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    r = np.asarray(point_cloud_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    gs = grid_size(r, v)
    c = np.floor((pts[:, :3] - r[:3]) / v)                        # float32 throughout, like the C++ template
    ok = np.all((c >= 0) & (c < gs.astype(np.float32)), axis=1)
    ci = c.astype(np.int64)
    key = (ci[:, 2] * gs[1] + ci[:, 1]) * gs[0] + ci[:, 0]
    F = pts.shape[1]
    voxels = np.zeros((max_voxels, max_points, F), dtype=np.float32)
    coords = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros(max_voxels, dtype=np.int32)
    index = {}
    n_vox = 0
    for i in np.nonzero(ok)[0]:                                    # the sequential rule itself
        k = int(key[i])
        vid = index.get(k, -1)
        if vid == -1:
            if n_vox >= max_voxels:
                continue
            vid = n_vox
            n_vox += 1
            index[k] = vid
            coords[vid] = (ci[i, 2], ci[i, 1], ci[i, 0])
        if num[vid] < max_points:
            voxels[vid, num[vid]] = pts[i]
            num[vid] += 1
    return voxels[:n_vox], coords[:n_vox], num[:n_vox]


def mask_and_voxelize(points, point_cloud_range, voxel_size, max_points, max_voxels):
    """data_processor.py:78-91 followed by :115-143 (use_lead_xyz = True)."""
    pts = points[mask_points_by_range(points, point_cloud_range)]
    return (pts,) + points_to_voxels(pts, point_cloud_range, voxel_size, max_points, max_voxels)
