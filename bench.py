#!/usr/bin/env python
"""
bench.py -- snowfall augmentation throughput on B200 (BASELINE.json metric: augmented LiDAR points/s).

    python bench.py --gpus 1 --steps 10 --warmup 3                     # our arm (CUDA engine)
    python bench.py --impl reference --steps 2 --warmup 1              # CPU arm: the oracle port on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                          # N ranks, one per GPU, weak scaling

Workload (N=1): BASELINE.json configs[1] -- batch = 32 synthetic 64 x 2048 clouds per GPU, snowfall_rate 2.5 mm/h,
terminal velocity 1.6 m/s, Gunn-Marshall size distribution; tables drawn by the engine's dart-throwing sampler.
One step = one pass of the whole augment() pipeline over the batch: channel sort, [pre-pass], per-beam solve,
threshold filter, compaction, stats.  With N > 1 every rank augments its own 32 clouds (clouds are independent, no
data-path collective) and one NCCL all-gather reassembles the augmented batch on every rank (configs[3]).

`value`     device-resident inputs; K steps bracketed by one CUDA-event pair on the launching stream (barrier +
            synchronize on both sides), max over ranks; two input batches alternate so that no step finds its rows in L2;
            with N > 1 the all-gather of step k overlaps the kernels of step k+1 (the last gathers are inside the bracket)
`e2e`       the public API with pinned HOST buffers: H2D of the batch + augment + D2H of the augmented batch, per step
`roofline`  the dominant kernel (k_snowfall) against the measured HBM copy peak; algorithmic bytes = 40 B/point
             + the candidate index it may touch once per launch
`cpu_baseline`  the CPU oracle port (oracle/) on a bounded sample of the same workload, rank 0, N=1 only
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SNOWFALL_RATE = 2.5
TERMINAL_VELOCITY = 1.6
MODE = 'gunn'
BATCH_PER_GPU = 32
N_AZIMUTH = 2048
DIV_DEG = float(np.degrees(3e-3))
ALGO_BYTES_PER_POINT = 40           # read 5 x f32, write 5 x f32
FIXED_POLY = (2e-3, -0.3, 12.0)     # only used with --host-threshold


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler runs from
    before the warm-up (nvidia-smi needs ~0.2 s to start); samples are stamped on arrival and the ones that fall inside
    the timed window are reported (the window is tens of ms: if none falls inside, the nearest ones are used and
    `window` says so)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '20', '-i', str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [t.strip() for t in line.split(',')]))

    def window_begin(self):
        self.t0 = time.perf_counter()

    def window_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if self.t0 is not None and self.t0 - 0.02 <= t <= self.t1 + 0.03]
        window = 'inside the timed region'
        if not rows and self.rows:
            mid = 0.5 * ((self.t0 or 0) + (self.t1 or 0))
            rows = [r for (t, r) in sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            window = 'nearest samples (timed region shorter than the sampling period)'
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nme in enumerate(names):
                    if r[5 + k].lower().startswith('active'):
                        reasons.add(nme)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


def make_workload(rank, batch, seed0=0):
    from lidar_snow_sim_b200.synthetic import synthetic_cloud
    clouds = [synthetic_cloud(seed=seed0 + rank * 10000 + b, n_azimuth=N_AZIMUTH) for b in range(batch)]
    orders = []
    for b in range(batch):
        r = random.Random(seed0 + rank * 10000 + b)
        o = list(range(64))
        r.shuffle(o)                                # the reference's random.shuffle(order), simulation.py:486
        orders.append(o)
    return clouds, np.array(orders, dtype=np.int32)


_REF_STATE = None


def _ref_init(tables, sensor, clouds, orders, poly, threads):
    global _REF_STATE
    from oracle import oracle as orc
    orc.lib()
    _REF_STATE = (orc, tables, sensor, clouds, orders, poly, threads)


def _ref_one_cloud(k):
    orc, tables, sensor, clouds, orders, poly, threads = _REF_STATE
    stats, aug = orc.augment(clouds[k], tables, DIV_DEG, sensor, order=orders[k].tolist(), thresh_poly=poly,
                             threads=threads, stable_sort=True)
    return stats


def run_reference(args):
    """CPU arm: the oracle port (oracle/, restating tools/snowfall/simulation.py) on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import oracle as orc
    from lidar_snow_sim_b200.snowfall.sampling import sample_table_set
    from lidar_snow_sim_b200.calib.hdl64e_s3 import sensor_arrays
    orc.build()
    cores = os.cpu_count() or 1
    tables = sample_table_set(MODE, SNOWFALL_RATE, TERMINAL_VELOCITY, seed=1000)
    # bounded sample per step so that `--steps K --warmup W` ends within a few minutes whatever K the driver passes
    # (one cloud costs ~3-6 core-seconds of the pre-pass + ~20 core-seconds of channel work)
    total_steps = args.steps + args.warmup
    clouds_per_step = args.cpu_clouds if args.cpu_clouds > 0 else (8 if total_steps <= 6 else 4 if total_steps <= 14
                                                                    else 2 if total_steps <= 30 else 1)
    clouds, orders = make_workload(0, clouds_per_step)
    sensor = sensor_arrays()
    poly = np.array(FIXED_POLY)

    # all host cores: clouds are independent -> one worker process per cloud of the step (like the reference's
    # process_map over channels, simulation.py:490-494), each with its share of threads for the 64 channel tasks
    import multiprocessing as mp
    n_workers = max(1, min(cores, clouds_per_step))
    threads_each = max(1, cores // n_workers)
    pool = mp.get_context('spawn').Pool(n_workers, initializer=_ref_init,
                                        initargs=(tables, sensor, clouds, orders, poly if args.host_threshold else None,
                                                  threads_each))
    pool.map(_ref_one_cloud, [], chunksize=1)      # workers up before timing

    def step():
        pool.map(_ref_one_cloud, range(clouds_per_step), chunksize=1)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    pts = sum(c.shape[0] for c in clouds)
    value = pts / dt
    pool.close()
    sample = (f'{clouds_per_step} clouds of 64x{N_AZIMUTH} per step ({pts} points), full augment() incl. pre-pass; '
              f'{n_workers} worker processes x {threads_each} threads')
    line = {'impl': 'reference', 'metric': 'augmented LiDAR points/sec', 'value': value, 'unit': 'points/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': value, 'unit': 'points/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'clouds_per_s': value / (64 * N_AZIMUTH)}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {'workload': f'BASELINE.json configs[1]: batch={BATCH_PER_GPU} synthetic 64x{N_AZIMUTH} clouds per GPU, '
                        f'snowfall_rate={SNOWFALL_RATE} mm/h, v={TERMINAL_VELOCITY} m/s, {MODE} DSD, '
                        f'beam_divergence=3 mrad' + ('' if n_gpus == 1 else f'; x{n_gpus} GPUs + NCCL all-gather of the '
                                                     f'augmented batch (configs[3])'),
            'batch_per_gpu': BATCH_PER_GPU, 'points_per_cloud': 64 * N_AZIMUTH, 'parallelism': f'clouds sharded x{n_gpus}',
            'l2': 'no explicit flush: two different input batches alternate (2 x 84 MB of rows + 92 MB index > 126 MB L2); '
                  'K steps timed with one CUDA-event pair on the launching stream'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH_PER_GPU)
    ap.add_argument('--cpu-clouds', type=int, default=0, help='clouds per step of the CPU arm / cpu_baseline sample')
    ap.add_argument('--host-threshold', action='store_true',
                    help='skip the device pre-pass and use a fixed threshold polynomial (debug only; reported in config)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-inflight', type=int, default=3, help='batches in flight in the e2e leg (1..3)')
    ap.add_argument('--e2e-chunks', type=int, default=2, help='chunks of the host-to-host pipeline (e2e leg)')
    args = ap.parse_args()
    args.steps = max(1, args.steps)
    args.warmup = max(3, args.warmup) if args.impl == 'b200' else max(0, args.warmup)

    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from lidar_snow_sim_b200.engine import SnowfallEngine
    from lidar_snow_sim_b200.snowfall.sampling import sample_table_set

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback for the b200 arm)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'

    numa_cpus = None
    if world > 1 and os.environ.get('LSS_NUMA_BIND'):     # opt-in: measured no consistent gain on the 2-GPU box
        from lidar_snow_sim_b200.distributed import bind_host_to_gpu
        numa_cpus = bind_host_to_gpu(local_rank)          # pinned host buffers land on the GPU's own NUMA node
    eng = SnowfallEngine(local_rank)
    tables = sample_table_set(MODE, SNOWFALL_RATE, TERMINAL_VELOCITY, seed=1000)
    tid = eng.upload_tables(tables)
    tinfo = eng.table_info(tid)
    B = args.batch
    # two different batches per rank, used alternately: 2 x 84 MB of rows (+ the 92 MB index) per pair of steps is more
    # than the 126 MB L2, so no step finds its inputs cached by the previous one (no explicit flush needed)
    clouds, orders = make_workload(rank, B)
    clouds2, orders2 = make_workload(rank, B, seed0=500000)
    n_per = [c.shape[0] for c in clouds]
    off = np.concatenate([[0], np.cumsum(n_per)]).astype(np.int64)
    N = int(off[-1])
    assert [c.shape[0] for c in clouds2] == n_per
    host_pts = torch.from_numpy(np.concatenate(clouds)).pin_memory()
    d_pts = [host_pts.to(dev), torch.from_numpy(np.concatenate(clouds2)).to(dev)]
    d_orders = [orders, orders2]
    poly = np.tile(np.array(FIXED_POLY), (B, 1)) if args.host_threshold else None
    device_prepass = not args.host_threshold
    outs = [{}, {}]
    gathered = gathered_counts = None
    pending = [None, None]
    if world > 1:
        gathered = [torch.empty((world * N, 5), dtype=torch.float32, device=dev) for _ in range(2)]
        gathered_counts = [torch.empty((world * B,), dtype=torch.int32, device=dev) for _ in range(2)]

    def step(k):
        """One pass of the augment() pipeline over this rank's batch; with N > 1 followed by the all-gather of the
        augmented batch, issued asynchronously on NCCL's stream so that it overlaps the next step's
        kernels (double-buffered; a buffer is reused only after its gather has completed)."""
        j = k & 1
        if pending[j] is not None:
            for wk in pending[j]:
                wk.wait()
            pending[j] = None
        r = eng.snowfall_batch(tid, d_pts[j], off, d_orders[j], DIV_DEG, thresh_poly=poly, device_prepass=device_prepass,
                               out=outs[j])
        if world > 1:
            pending[j] = [dist.all_gather_into_tensor(gathered[j], r['points'], async_op=True),
                          dist.all_gather_into_tensor(gathered_counts[j], r['counts'], async_op=True)]
        return r

    def drain():
        for j in range(2):
            if pending[j] is not None:
                for wk in pending[j]:
                    wk.wait()
                pending[j] = None

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) ------------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    for k in range(args.warmup):
        step(k)
    drain()
    eng.check()
    eng.set_profiling(True)
    eng.kernel_times(reset=True)
    launches0 = eng.launch_count()
    sync_all()
    clocks.window_begin()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step(k)
    drain()                                                 # the last gathers are inside the timed region
    e1.record()
    sync_all()
    clocks.window_end()
    clk = clocks.stop()
    eng.check()
    total_ms = float(e0.elapsed_time(e1))
    launches = eng.launch_count() - launches0
    ktimes = eng.kernel_times(reset=True)
    eng.set_profiling(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    points_all = N * world
    value = points_all / (ms_per_step * 1e-3)
    orders = d_orders[0]

    # ---- end to end through the public API with host buffers (`e2e`) --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # The public host-to-host API, called the way a prefetching data loader calls it: up to `depth` batches are in
        # flight (one pinned result buffer each), so that batch k+1's copy-in overlaps batch k's kernels and batch k-1's
        # copy-out.  Every step moves its own input H2D and its own result D2H inside the timed region.
        depth = max(1, min(3, args.e2e_inflight))
        host_outs = [{} for _ in range(depth)]
        e2e_kw = dict(thresh_poly=poly, device_prepass=device_prepass, n_chunks=args.e2e_chunks)

        def e2e_run(steps):
            tickets = []
            for k in range(steps):
                if len(tickets) == depth:
                    eng.snowfall_batch_host_wait(tickets.pop(0))        # the caller consumes the oldest batch here
                tickets.append(eng.snowfall_batch_host_submit(tid, host_pts, off, orders, DIV_DEG,
                                                              host_out=host_outs[k % depth], **e2e_kw))
            for t in tickets:
                eng.snowfall_batch_host_wait(t)

        e2e_run(3)
        sync_all()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        sync_all()
        dt = (time.perf_counter() - t0) / args.steps
        # the same API called synchronously (one batch at a time, nothing in flight across calls): latency per batch
        t0 = time.perf_counter()
        for _ in range(max(3, args.steps // 2)):
            eng.snowfall_batch_host(tid, host_pts, off, orders, DIV_DEG, host_out=host_outs[0], **e2e_kw)
        dt_sync = (time.perf_counter() - t0) / max(3, args.steps // 2)
        tt = torch.tensor([dt, dt_sync], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_sync = float(tt[0].item()), float(tt[1].item())
        e2e = {'value': points_all / dt, 'unit': 'points/s', 'h2d_bytes_per_step': int(N * 20),
               'd2h_bytes_per_step': int(N * 20 + B * 4 + B * 32), 'ms_per_step': dt * 1e3,
               'sync_call': {'value': points_all / dt_sync, 'ms_per_step': dt_sync * 1e3},
               'chunks': args.e2e_chunks, 'batches_in_flight': depth,
               'host_numa_bound_cpus': None if numa_cpus is None else len(numa_cpus),
               'timing': 'host wall clock around K steps of the C-ABI host-buffer calls lss_snowfall_batch_host_submit / '
                         '_wait (pinned host in -> copy-in / pre-pass / beam / copy-out streams -> pinned host out), '
                         'with up to batches_in_flight steps submitted before the oldest is awaited; sync_call = the same batches '
                         'through the synchronous lss_snowfall_batch_host, one at a time; with N > 1 every rank feeds '
                         'its own host-side consumer, no gather'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------------
    peak, peak_src = load_peaks()
    k_ms, k_calls = ktimes.get('snowfall', (0.0, 0))
    k_avg_ms = k_ms / max(k_calls, 1)
    algo_bytes = ALGO_BYTES_PER_POINT * N + tinfo['bytes']
    achieved = algo_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    roofline = {'bound': 'hbm', 'kernel': 'k_snowfall beam stage: scan <24,0> + list sort + solve <24,1> (dominant) + overflow <128,1> launches',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': int(algo_bytes), 'kernel_ms': k_avg_ms,
                'kernel_share_of_step': k_avg_ms / ms_per_step,
                'kernel_ms_all': {k: (v[0] / max(v[1], 1)) * (v[1] / args.steps) for k, v in ktimes.items() if v[1]},
                'note': 'latency / FP64-issue bound, not HBM bound (DESIGN.md 4, profiles/): duration = CUDA events around the '
                        'beam-stage launches while the pre-pass runs concurrently on a side stream (kernel_ms_all therefore '
                        'sums to more than the step); traffic = summed dram bytes of the scan, solve and overflow launches '
                        'from profiles/traffic.json'}
    prof = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(prof):
        try:
            roofline['traffic'] = json.load(open(prof)).get('k_snowfall_dram_bytes_per_launch')
        except Exception:
            pass

    # ---- CPU baseline (oracle port, bounded sample) ---------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        from lidar_snow_sim_b200.calib.hdl64e_s3 import sensor_arrays
        import multiprocessing as mp
        orc.build()
        cores = os.cpu_count() or 1
        sample = clouds[:(args.cpu_clouds if args.cpu_clouds > 0 else 8)]
        n_workers = max(1, min(cores, len(sample)))
        threads_each = max(1, cores // n_workers)
        with mp.get_context('spawn').Pool(n_workers, initializer=_ref_init,
                                          initargs=(tables, sensor_arrays(), sample, orders,
                                                    None if device_prepass else np.array(FIXED_POLY), threads_each)) as pool:
            pool.map(_ref_one_cloud, list(range(min(n_workers, len(sample)))), chunksize=1)     # warm up the workers
            t0 = time.perf_counter()
            pool.map(_ref_one_cloud, range(len(sample)), chunksize=1)
            dt = time.perf_counter() - t0
        cpu = {'value': sum(c.shape[0] for c in sample) / dt, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
               'sample': f'{len(sample)} of the {B} clouds of one step ({sum(c.shape[0] for c in sample)} points), '
                         f'oracle port (C core + numpy/scipy/sklearn pre-pass), {n_workers} processes x '
                         f'{threads_each} threads, {dt:.1f} s'}

    cfg = workload_config(args.gpus)
    cfg['prepass'] = 'device' if device_prepass else 'DEBUG: fixed host-supplied threshold polynomial'
    cfg['table_particles'] = tinfo['n_particles']
    cfg['table_index_bytes'] = tinfo['bytes']
    line = {'metric': 'augmented LiDAR points/sec', 'value': value, 'unit': 'points/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
            'clouds_per_s': value / (64 * N_AZIMUTH), 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clk,
            'roofline': roofline, 'cpu_baseline': cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
