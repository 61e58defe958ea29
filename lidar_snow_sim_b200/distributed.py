"""
Multi-GPU use of the engine: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box, gloo
in the CPU tests).  Clouds are independent, so a batch is sharded over ranks by cloud with NO
data-path collective; the only exchange is the final all-gather that reassembles the augmented batch on every rank
(BASELINE.json configs[3]) -- and a consumer that lives on the same rank (a per-rank DataLoader) can skip even that.

Layout of the gathered batch: fixed stride per rank.  Every rank contributes a (slot_rows, 5) float32 buffer in which
cloud b occupies rows local_offsets[b] .. local_offsets[b] + counts[b] (the slot-compacted output of
SnowfallEngine.snowfall_batch), padded to the largest rank's row count, plus its per-cloud counts and stats padded to
the largest rank's cloud count.
"""
import numpy as np
import torch
import torch.distributed as dist


def bind_host_to_gpu(device: int = 0):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (first-touch then places pinned host buffers
    there), so that one rank's PCIe traffic does not cross the socket interconnect.  Call before allocating host
    buffers.  Returns the cpu list used, or None when the topology cannot be read (no-op)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device)
        pci = f'{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{pci}/local_cpulist') as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(','):
            if '-' in part:
                lo, hi = part.split('-')
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def shard_range(n_clouds: int, rank: int, world: int):
    """Contiguous block partition of cloud indices: rank r owns [start, stop)."""
    base, rem = divmod(n_clouds, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_augmented(points, local_offsets, counts, stats, group=None):
    """
    points: (N_local, 5) float32 slot-compacted rows; local_offsets: host int64 (B_local + 1); counts: (B_local,) int32;
    stats: (B_local, 4) float64.  Works for ragged shards (different N_local / B_local per rank).
    Returns dict(points (world, N_max, 5), counts (world, B_max), stats (world, B_max, 4), offsets list per rank,
                 n_clouds list per rank); `cloud(r, b)` slices are `points[r, offsets[r][b] : offsets[r][b] + counts[r, b]]`.
    One all-gather for the rows (the only large message), two small ones for counts / stats, one for the shapes.
    """
    world = dist.get_world_size(group)
    dev = points.device
    local_offsets = np.ascontiguousarray(local_offsets, dtype=np.int64)
    b_local = local_offsets.shape[0] - 1
    n_local = int(local_offsets[-1])
    meta = torch.tensor([n_local, b_local], dtype=torch.int64, device=dev)
    metas = torch.empty((world * 2,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas_h = metas.cpu().numpy().reshape(world, 2)
    n_max, b_max = int(metas_h[:, 0].max()), int(metas_h[:, 1].max())

    def padded(t, rows, tail_shape, dtype):
        if t.shape[0] == rows:
            return t.contiguous()
        buf = torch.zeros((rows,) + tail_shape, dtype=dtype, device=dev)
        buf[:t.shape[0]] = t
        return buf

    p = padded(points, n_max, (5,), torch.float32)
    c = padded(counts, b_max, (), torch.int32)
    s = padded(stats, b_max, (4,), torch.float64)
    o = padded(torch.from_numpy(local_offsets[:-1].copy()).to(dev), b_max, (), torch.int64)
    # outputs are the concatenation along dim 0 (the layout every backend accepts), viewed per rank afterwards
    g_p = torch.empty((world * n_max, 5), dtype=torch.float32, device=dev)
    g_c = torch.empty((world * b_max,), dtype=torch.int32, device=dev)
    g_s = torch.empty((world * b_max, 4), dtype=torch.float64, device=dev)
    g_o = torch.empty((world * b_max,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(g_p, p, group=group)
    dist.all_gather_into_tensor(g_c, c, group=group)
    dist.all_gather_into_tensor(g_s, s, group=group)
    dist.all_gather_into_tensor(g_o, o, group=group)
    g_p = g_p.view(world, n_max, 5)
    g_c = g_c.view(world, b_max)
    g_s = g_s.view(world, b_max, 4)
    g_o_h = g_o.cpu().numpy().reshape(world, b_max)
    return dict(points=g_p, counts=g_c, stats=g_s, offsets=[g_o_h[r, :metas_h[r, 1]] for r in range(world)],
                n_clouds=[int(v) for v in metas_h[:, 1]])


def unpack_clouds(gathered):
    """Host-side view of a gathered batch as a flat list of per-cloud (count, 5) tensors in global cloud order."""
    out = []
    counts = gathered['counts'].cpu().numpy()
    for r, nb in enumerate(gathered['n_clouds']):
        for b in range(nb):
            o = int(gathered['offsets'][r][b])
            out.append(gathered['points'][r, o:o + int(counts[r, b])])
    return out


class ShardedAugmenter:
    """
    Batch-level driver: every rank augments its contiguous block of the global batch on its own GPU and (optionally)
    all ranks exchange the results.  `augment_fn(points, offsets, orders) -> dict(points, counts, stats)` is
    SnowfallEngine.snowfall_batch bound to a table set (tests pass a CPU stand-in to exercise the plumbing on gloo).
    """

    def __init__(self, augment_fn, group=None):
        self.augment_fn = augment_fn
        self.group = group

    def run(self, clouds, orders, device, gather=True):
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = shard_range(len(clouds), rank, world)
        mine = clouds[lo:hi]
        off = np.concatenate([[0], np.cumsum([c.shape[0] for c in mine])]).astype(np.int64)
        if len(mine):
            pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(mine), dtype=np.float32)).to(device)
        else:
            pts = torch.zeros((0, 5), dtype=torch.float32, device=device)
        res = self.augment_fn(pts, off, np.asarray(orders[lo:hi], dtype=np.int32).reshape(-1, 64))
        if not gather:
            return res, (lo, hi)
        return all_gather_augmented(res['points'], off, res['counts'], res['stats'], group=self.group), (lo, hi)
