#!/usr/bin/env python
"""The bench workload (32 clouds of 64 x 2048, 2.5 mm/h tables) for K steps, nothing else -- the command ncu wraps.
    python tools/profile_step.py --steps 6 [--config 2]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                     # noqa: E402
from lidar_snow_sim_b200.engine import SnowfallEngine            # noqa: E402
from lidar_snow_sim_b200.snowfall.sampling import sample_table_set   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=6)
ap.add_argument('--config', type=int, default=1)
ap.add_argument('--batch', type=int, default=bench.BATCH_PER_GPU)
ap.add_argument('--rate', type=float, default=bench.SNOWFALL_RATE)
ap.add_argument('--velocity', type=float, default=bench.TERMINAL_VELOCITY)
args = ap.parse_args()
eng = SnowfallEngine(0)
tid = eng.upload_tables(sample_table_set(bench.MODE, args.rate, args.velocity, seed=bench.TABLE_SEED))
clouds, orders = bench.make_workload(0, args.batch)
clouds2, orders2 = bench.make_workload(0, args.batch, seed0=500000)
off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
d = [torch.from_numpy(np.concatenate(clouds)).cuda(), torch.from_numpy(np.concatenate(clouds2)).cuda()]
o = [orders, orders2]
outs = [{}, {}]
for k in range(args.steps):
    r = eng.snowfall_batch(tid, d[k & 1], off, o[k & 1], bench.DIV_DEG, device_prepass=True, out=outs[k & 1])
    if args.config == 2:
        eng.wet_ground_batch(r['points'], off, counts=r['counts'], water_height=bench.WATER_HEIGHT, replace=False)
torch.cuda.synchronize()
eng.check()
print('info', eng.table_info(tid), 'launches', eng.launch_count())
