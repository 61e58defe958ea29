"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: sharding by cloud + the all-gather of the augmented batch.
The augmentation itself has no CPU path; a stand-in `augment_fn` that mimics the engine's slot-compacted output
exercises exactly the host logic that runs around it on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidar_snow_sim_b200.distributed import shard_range, all_gather_augmented, unpack_clouds, ShardedAugmenter


def test_shard_range_partitions():
    for n in (0, 1, 5, 32, 33, 256):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                lo, hi = shard_range(n, r, w)
                seen += list(range(lo, hi))
                assert 0 <= hi - lo <= -(-n // w) if n else hi == lo
            assert seen == list(range(n))


def fake_augment(points, off, orders):
    """Keeps every second row of each cloud and tags column 4 with the plane of channel 0 (stand-in for the engine)."""
    B = off.shape[0] - 1
    out = torch.zeros_like(points)
    counts = torch.zeros(B, dtype=torch.int32)
    stats = torch.zeros((B, 4), dtype=torch.float64)
    for b in range(B):
        rows = points[off[b]:off[b + 1]][::2].clone()
        rows[:, 4] = float(orders[b, 0])
        out[off[b]:off[b] + rows.shape[0]] = rows
        counts[b] = rows.shape[0]
        stats[b, 0] = rows.shape[0]
        stats[b, 3] = float(rows[:, 3].sum())
    return dict(points=out, counts=counts, stats=stats)


def _worker(rank, world, port, n_clouds, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        clouds = [rng.normal(size=(int(rng.integers(3, 40)), 5)).astype(np.float32) for _ in range(n_clouds)]
        orders = np.stack([np.random.default_rng(b).permutation(64) for b in range(n_clouds)]) if n_clouds else \
            np.zeros((0, 64), np.int32)
        aug = ShardedAugmenter(fake_augment)
        gathered, (lo, hi) = aug.run(clouds, orders, torch.device('cpu'))
        got = unpack_clouds(gathered)
        ok = len(got) == n_clouds
        for b in range(n_clouds):
            want = clouds[b][::2].copy()
            want[:, 4] = orders[b, 0]
            ok &= np.array_equal(got[b].numpy(), want)
        tot = sum(int(gathered['counts'][r, :nb].sum()) for r, nb in enumerate(gathered['n_clouds']))
        ok &= tot == sum(c[::2].shape[0] for c in clouds)
        q.put((rank, bool(ok), lo, hi))
    except Exception as exc:                      # surface worker failures instead of a queue timeout
        q.put((rank, False, repr(exc), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_clouds', [7, 2, 1])
def test_gloo_world2_sharded_gather(n_clouds):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clouds, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    spans = sorted((r[2], r[3]) for r in res)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == n_clouds


def _gather_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from lidar_snow_sim_b200.distributed import BatchGather
        n_rows, n_clouds = 50, 3
        dev = torch.device('cpu')
        # asking for the push kernel without an engine / off a GPU must fall back to the library all-gather, not fail
        g = BatchGather(n_rows, n_clouds, dev, depth=2, kind='push')
        ok = g.kind == 'nccl'
        for step in range(5):                     # double-buffered: buffer j is reused every second step
            j = step & 1
            g.wait(j)
            pts = torch.full((n_rows, 5), float(100 * step + rank))
            cnt = torch.full((n_clouds,), 10 * step + rank, dtype=torch.int32)
            g.start(j, pts, cnt)
            g.wait(j)
            for r in range(world):
                ok &= bool((g.points[j][r * n_rows:(r + 1) * n_rows] == float(100 * step + r)).all())
                ok &= bool((g.counts[j][r * n_clouds:(r + 1) * n_clouds] == 10 * step + r).all())
        g.wait_all()
        q.put((rank, bool(ok)))
    except Exception as exc:
        q.put((rank, False, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_batch_gather_falls_back_to_the_library_all_gather():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
