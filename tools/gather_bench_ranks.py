#!/usr/bin/env python
"""
Stand-alone timing of the exchange step on N ranks with NO kernels next to it: what do the fabric and each
gather kind deliver for the bench's batch (32 x 131072 rows per rank, ~75 % of the rows kept)?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/gather_bench_ranks.py

Per variant: ms per exchange (CUDA events on the launching rank, max over ranks, 20 exchanges back to back, double
buffered) and the bytes that landed on one rank from its peers.  Rank 0 prints one JSON object.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    lr = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(lr)
    dev = torch.device('cuda', lr)
    dist.init_process_group('nccl', device_id=dev)
    from lidar_snow_sim_b200.distributed import BatchGather
    from lidar_snow_sim_b200.engine import SnowfallEngine
    eng = SnowfallEngine(lr)
    B, n_per = 32, 131072
    off = (np.arange(B + 1, dtype=np.int64) * n_per)
    n_rows = int(off[-1])
    g0 = np.random.default_rng(rank)
    cnt = (n_per * g0.uniform(0.65, 0.85, size=B)).astype(np.int32)
    d_pts = torch.randn((n_rows, 5), dtype=torch.float32, device=dev)
    d_cnt = torch.from_numpy(cnt).to(dev)
    res = {'world': world, 'rows_per_rank': n_rows, 'kept_fraction': float(cnt.sum() / n_rows)}
    variants = [('push_mc_b37', 'push', {'LSS_GATHER_BLOCKS': '37'}), ('push_mc_b74', 'push', {'LSS_GATHER_BLOCKS': '74'}),
                ('push_mc_b148', 'push', {'LSS_GATHER_BLOCKS': '148'}), ('push_mc_b296', 'push', {'LSS_GATHER_BLOCKS': '296'}),
                ('push_uni_b32', 'push', {'LSS_GATHER_MULTICAST': '0', 'LSS_GATHER_BLOCKS': '32'}),
                ('push_uni_b64', 'push', {'LSS_GATHER_MULTICAST': '0', 'LSS_GATHER_BLOCKS': '64'}),
                ('push_uni_b148', 'push', {'LSS_GATHER_MULTICAST': '0', 'LSS_GATHER_BLOCKS': '148'}),
                ('ce', 'ce', {}), ('nccl', 'nccl', {})]
    for name, kind, env in variants:
        for k, v in env.items():
            os.environ[k] = v
        g = BatchGather(n_rows, B, dev, depth=2, kind=kind, engine=eng, cloud_offsets=off)
        for k in env:
            del os.environ[k]

        def run(n):
            for s in range(n):
                g.wait(s & 1)
                g.start(s & 1, d_pts, d_cnt)
            g.wait_all()
        run(4)
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(20)
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / 20], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        moved = (cnt.sum() if g.kind == 'push' else n_rows) * 20.0 * (world - 1)
        res[name] = {'kind_used': g.kind, 'multicast': bool(getattr(g, 'multicast', False)), 'ms': round(float(t.item()), 4),
                     'inbound_GBs_per_rank': round(moved / (float(t.item()) * 1e-3) / 1e9, 1)}
        del g
        torch.cuda.synchronize(dev)
        dist.barrier()
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
