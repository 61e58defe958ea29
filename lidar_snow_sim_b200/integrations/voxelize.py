"""
Point-range mask + voxelisation on the device: the two DataProcessor steps that sit between the augmentation and the
detector in the reference's training path (lib/OpenPCDet/pcdet/datasets/processor/data_processor.py:78-91, 115-143;
config: lib/OpenPCDet/tools/cfgs/dataset_configs/dense_dataset.yaml:4,66-78), so that an augmented batch can stay on
the GPU all the way to the detector input.

    proc = DeviceVoxelizer(point_cloud_range=[0, -40, -3, 70.4, 40, 1], voxel_size=[0.05, 0.05, 0.1],
                           max_points_per_voxel=5, max_number_of_voxels=16000)
    data_dict = proc(data_dict)          # same keys as DataProcessor: 'voxels', 'voxel_coords', 'voxel_num_points'

    out = proc.batch(points_cuda, cloud_offsets, counts)     # batched, device in / device out (collated layout)

The voxel rule is spconv's (third party, not part of the reference tree): see oracle/voxel.py and csrc/voxelize.cu.
"""
import numpy as np
import torch

from ..engine import default_engine


class DeviceVoxelizer:
    def __init__(self, point_cloud_range, voxel_size, max_points_per_voxel, max_number_of_voxels, engine=None,
                 mask_outside_range=True):
        self.point_cloud_range = np.asarray(point_cloud_range, dtype=np.float32)
        self.voxel_size = np.asarray(voxel_size, dtype=np.float32)
        self.max_points = int(max_points_per_voxel)
        self.max_voxels = int(max_number_of_voxels)
        self.mask_outside_range = bool(mask_outside_range)
        self.engine = engine
        # data_processor.py:117-118
        self.grid_size = np.round((self.point_cloud_range[3:6] - self.point_cloud_range[0:3]) / self.voxel_size).astype(np.int64)

    def _engine(self):
        if self.engine is None:
            self.engine = default_engine()
        return self.engine

    def batch(self, points, cloud_offsets, counts=None):
        """points: CUDA float32 (N, F); cloud_offsets: host int64 (B + 1); counts: optional CUDA int32 (B,) valid rows per
        cloud slot (the slot-compacted output of the augmentation kernels).  Returns the engine's dict: voxels
        (B, max_voxels, max_points, F), coords (B, max_voxels, 4) = (batch index, z, y, x), num_points (B, max_voxels),
        n_voxels (B,) -- all CUDA tensors; rows >= n_voxels[b] of cloud b are zero."""
        return self._engine().voxelize_batch(points, cloud_offsets, self.point_cloud_range, self.voxel_size,
                                             self.max_points, self.max_voxels, counts=counts,
                                             mask_xy_range=self.mask_outside_range)

    def collate(self, out):
        """The collated arrays DatasetTemplate.collate_batch builds (lib/OpenPCDet/pcdet/datasets/dataset.py:190-204):
        voxels (M, max_points, F), voxel_coords (M, 4), voxel_num_points (M,), clouds concatenated in batch order."""
        n = out['n_voxels'].cpu().numpy()
        sel = torch.cat([torch.arange(int(n[b]), device=out['voxels'].device) + b * self.max_voxels for b in range(len(n))])
        F = out['voxels'].shape[-1]
        return dict(voxels=out['voxels'].reshape(-1, self.max_points, F)[sel],
                    voxel_coords=out['coords'].reshape(-1, 4)[sel],
                    voxel_num_points=out['num_points'].reshape(-1)[sel])

    def __call__(self, data_dict):
        """One cloud, NumPy in / NumPy out: DataProcessor.mask_points_and_boxes_outside_range (points part) followed by
        transform_points_to_voxels, with the reference's keys (use_lead_xyz honoured, data_processor.py:137-138)."""
        pts = np.ascontiguousarray(data_dict['points'], dtype=np.float32)
        eng = self._engine()
        d = torch.from_numpy(pts).to(eng.device)
        out = self.batch(d, np.array([0, pts.shape[0]], dtype=np.int64))
        eng.check()
        n = int(out['n_voxels'][0].item())
        voxels = out['voxels'][0, :n].cpu().numpy()
        if self.mask_outside_range:
            r = self.point_cloud_range
            m = (pts[:, 0] >= r[0]) & (pts[:, 0] <= r[3]) & (pts[:, 1] >= r[1]) & (pts[:, 1] <= r[4])
            data_dict['points'] = data_dict['points'][m]
        if not data_dict.get('use_lead_xyz', True):
            voxels = voxels[..., 3:]
        data_dict['voxels'] = voxels
        data_dict['voxel_coords'] = out['coords'][0, :n, 1:].cpu().numpy()
        data_dict['voxel_num_points'] = out['num_points'][0, :n].cpu().numpy()
        return data_dict
