"""
GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against
  (1) the golden vectors frozen from the unmodified reference (tests/golden/),
  (2) the CPU oracle on the same seeded inputs at sizes it finishes in seconds,
  (3) size-independent properties at BASELINE.json's full batch size.
Bar: labels (the occluded-point mask), integer intensities and the keep mask exact; xyz within 1e-4 relative
(in practice bit-identical).  The beam azimuth theta is injected where exact-mask parity is asserted, because the
reference's float32 arctan2 is host dependent; the device-computed theta is checked separately.
"""
import os

import numpy as np
import pytest
import torch

from helpers import DIV, canon, channel_case, augment_case, augment_full_case, sha
from lidar_snow_sim_b200.calib.hdl64e_s3 import sensor_arrays
from lidar_snow_sim_b200.synthetic import synthetic_cloud, synthetic_particles

pytestmark = pytest.mark.gpu


def run_full(engine, tid, pc, order, theta=None, thresh_poly=None, **kw):
    """Un-filtered, channel-sorted rows + perm + occluder counts for one cloud."""
    d_pc = torch.from_numpy(np.ascontiguousarray(pc, dtype=np.float32)).cuda()
    d_th = None if theta is None else torch.from_numpy(np.ascontiguousarray(theta, dtype=np.float32)).cuda()
    off = np.array([0, pc.shape[0]], dtype=np.int64)
    res = engine.snowfall_batch(tid, d_pc, off, np.asarray(order, dtype=np.int32)[None], DIV, theta=d_th,
                                thresh_poly=thresh_poly, threshold_filter=thresh_poly is not None, want_full=True,
                                want_perm=True, want_nocc=True, **kw)
    engine.check()
    return {k: v.cpu().numpy() for k, v in res.items()}


def assert_rows_match(got, want, what=''):
    assert got.shape == want.shape, what
    assert np.array_equal(got[:, 4], want[:, 4]), f'{what}: label mask differs'
    assert np.array_equal(got[:, 3], want[:, 3]), f'{what}: intensities differ'
    rel = np.abs(got[:, :3] - want[:, :3]) / np.maximum(np.abs(want[:, :3]), 1e-6)
    assert rel.max() <= 1e-4, f'{what}: xyz off by {rel.max()}'


# ----------------------------------------------------------------------------------------------------------------------
# (1) golden vectors from the reference
# ----------------------------------------------------------------------------------------------------------------------
def test_golden_kat_channel(engine, gold_dir):
    g = np.load(os.path.join(gold_dir, 'kat_channel.npz'))
    tid = engine.upload_tables([g['particles']] * 64)
    r = run_full(engine, tid, g['points'], list(range(64)), theta=g['theta'])
    engine.free_tables(tid)
    assert np.array_equal(r['full'], g['out'])              # all points are channel 2: sorted order == input order
    assert r['stats'][0, 3] == float(g['intensity_diff_sum'])
    assert np.array_equal(r['nocc'], g['n_occluders'])


def test_golden_channel_cases(engine, gold_dir):
    rec = np.load(os.path.join(gold_dir, 'channel_cases.npz'))
    for ci in range(int(rec['n_cases'])):
        table = channel_case(rec, ci)
        tid = engine.upload_tables([table] * 64)
        r = run_full(engine, tid, rec[f'c{ci}_points'], list(range(64)), theta=rec[f'c{ci}_theta'])
        engine.free_tables(tid)
        assert np.array_equal(r['full'], rec[f'c{ci}_out']), f'case {ci}'
        # the reference adds (0.9 * max_intensity - new_i) beam by beam; the device adds exact integer partial sums
        assert np.isclose(r['stats'][0, 3], float(rec[f'c{ci}_sum']), rtol=1e-12, atol=0)
        assert np.array_equal(r['nocc'], rec[f'c{ci}_nocc'])


@pytest.mark.parametrize('name', ['augment_a', 'augment_b'])
def test_golden_augment_api(engine, gold_dir, name):
    """The reference-signature wrapper end to end (threshold polynomial injected from the reference run)."""
    from lidar_snow_sim_b200.snowfall.simulation import augment
    g = np.load(os.path.join(gold_dir, f'{name}.npz'))
    pc, tables = augment_case(g)
    stats, aug = augment(pc, 'unused', DIV, only_camera_fov=bool(g['fov']), engine=engine, tables=tables,
                         order=g['order'].tolist(), thresh_poly=g['thresh_poly'], theta=g['theta'])
    assert stats == tuple(int(v) for v in g['stats'])
    assert aug.dtype == np.float32
    assert np.array_equal(canon(aug), g['out'])


def test_golden_augment_full_size(engine, gold_dir):
    """BASELINE.json configs[0] end to end against the reference's own output (stats + SHA-256 of the rows)."""
    from lidar_snow_sim_b200.snowfall.simulation import augment
    g = np.load(os.path.join(gold_dir, 'augment_full.npz'))
    pc, tables, theta = augment_full_case(g)
    stats, aug = augment(pc, 'unused', DIV, only_camera_fov=False, engine=engine, tables=tables,
                         order=g['order'].tolist(), thresh_poly=g['thresh_poly'], theta=theta)
    assert stats == tuple(int(v) for v in g['stats'])
    assert aug.shape == tuple(g['out_shape']) and sha(canon(aug)) == str(g['out_sha'])
    # With the device pre-pass the kept set differs from THIS reference run: the reference's threshold polynomial hinges
    # on np.argpartition's implementation-defined pick (on this cloud the AVX-512 NumPy pick and the
    # portable first-minimum pick give polynomials of opposite curvature).  tests/test_prepass_gpu.py pins the device
    # pre-pass to the oracle run with the portable rule; here only the un-filtered solve must be unaffected.
    stats2, aug2, gi = augment(pc, 'unused', DIV, only_camera_fov=False, engine=engine, tables=tables,
                               order=g['order'].tolist(), theta=theta, return_internals=True)
    print('reference stats', stats, 'device pre-pass stats', stats2)
    assert [(gi['full'][:, 4] == l).sum() for l in (0, 1, 2)] == g['label_counts_unfiltered'].tolist()


# ----------------------------------------------------------------------------------------------------------------------
# (2) against the oracle on seeded inputs
# ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def tables18k():
    return [synthetic_particles(7000 + k, 18000) for k in range(64)]


def test_vs_oracle_cloud(engine, oracle, tables18k):
    rng = np.random.default_rng(5)
    pc = synthetic_cloud(seed=5, n_azimuth=384, drop=0.08, shuffle_rows=True)
    order = rng.permutation(64).tolist()
    poly = np.array([1e-3, -0.2, 14.0])
    idx = pc[:, 4].argsort(kind='stable')
    pcs = pc[idx]
    o_stats, o_aug, oi = oracle.augment(pc, tables18k, DIV, sensor_arrays(), order=order, thresh_poly=poly,
                                        stable_sort=True, return_internals=True)
    tid = engine.upload_tables(tables18k)
    theta_orig = np.empty(pc.shape[0], dtype=np.float32)
    theta_orig[idx] = oi['theta']                          # oracle host's atan2f bits, back in original row order
    r = run_full(engine, tid, pc, order, theta=theta_orig, thresh_poly=poly)
    assert np.array_equal(r['perm'], idx)                  # stable channel sort
    full_o = oi['full']
    assert_rows_match(r['full'], full_o, 'vs oracle')
    assert np.array_equal(r['full'], full_o)               # in practice bit-identical
    assert np.array_equal(r['nocc'], oi['n_occluders'])
    n = int(r['counts'][0])
    assert np.array_equal(r['points'][:n], o_aug)
    assert (int(r['stats'][0, 0]), int(r['stats'][0, 1]), int(r['stats'][0, 2])) == o_stats
    assert np.isclose(r['stats'][0, 3], oi['intensity_diff_sum'], rtol=1e-12, atol=0)

    # device-computed theta (correctly rounded float32 of the float64 atan2): the only differences allowed are beams
    # whose azimuth differs by an ulp from the host libm's atan2f -- report and bound the mismatch rate
    r2 = run_full(engine, tid, pc, order, thresh_poly=poly)
    mism = (r2['full'][:, 4] != full_o[:, 4]).mean()
    print(f'label mismatch rate with device theta: {mism:.2e}')
    assert mism < 2e-3
    th64 = np.arctan2(pcs[:, 1].astype(np.float64), pcs[:, 0].astype(np.float64)).astype(np.float32)
    same_theta = th64 == oi['theta']
    assert np.array_equal(r2['full'][same_theta], full_o[same_theta])
    engine.free_tables(tid)


def test_batch_ragged_and_empty(engine, oracle, tables18k):
    """Ragged batch with an empty cloud, a tiny cloud and rows with invalid channel ids."""
    clouds = [synthetic_cloud(seed=20, n_azimuth=64), np.zeros((0, 5), np.float32),
              synthetic_cloud(seed=21, n_azimuth=96, shuffle_rows=True)[:777], synthetic_cloud(seed=22, n_azimuth=32)]
    clouds[3] = clouds[3].copy()
    clouds[3][5, 4] = 64.0          # not a channel: passes through untouched
    clouds[3][9, 4] = 7.5
    clouds[3][11, 4] = -1.0
    rng = np.random.default_rng(9)
    orders = np.stack([rng.permutation(64) for _ in clouds]).astype(np.int32)
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(clouds)).cuda()
    tid = engine.upload_tables(tables18k)
    th = []
    want = []
    for c, o in zip(clouds, orders):
        if c.shape[0] == 0:
            want.append(np.zeros((0, 5), np.float32))
            th.append(np.zeros(0, np.float32))
            continue
        idx = c[:, 4].argsort(kind='stable')
        cs = c[idx]
        aug, s, nocc, theta = oracle.snow_cloud(cs, tables18k, o.tolist(), sensor_arrays(), DIV)
        aug[:, 3] = np.round(aug[:, 3])
        want.append(aug)
        t = np.empty(c.shape[0], np.float32)
        t[idx] = theta
        th.append(t)
    theta = torch.from_numpy(np.concatenate(th)).cuda()
    res = engine.snowfall_batch(tid, pts, off, orders, DIV, theta=theta, threshold_filter=False, want_full=True)
    engine.check()
    full = res['full'].cpu().numpy()
    counts = res['counts'].cpu().numpy()
    for b, c in enumerate(clouds):
        got = full[off[b]:off[b + 1]]
        valid = (want[b][:, 4] >= 0) if got.shape[0] else np.zeros(0, bool)
        assert counts[b] == c.shape[0]                      # no filter: everything kept
        if got.shape[0] == 0:
            continue
        # rows with an invalid channel id sort to the end (stable) and keep the channel value in column 4
        cs = c[c[:, 4].argsort(kind='stable')]
        ok_ch = (cs[:, 4] >= 0) & (cs[:, 4] < 64) & (cs[:, 4] == np.floor(cs[:, 4]))
        n_ok = int(ok_ch.sum())
        good = canon(want[b][ok_ch])
        assert np.array_equal(canon(got[:n_ok]), good)
        bad_rows = got[n_ok:]
        assert bad_rows.shape[0] == (~ok_ch).sum()
        assert np.array_equal(canon(bad_rows), canon(cs[~ok_ch]))
        assert np.array_equal(res['points'].cpu().numpy()[off[b]:off[b] + counts[b]], got)
    engine.free_tables(tid)


def test_errors(engine, tables18k):
    from lidar_snow_sim_b200.snowfall.simulation import augment
    tid = engine.upload_tables(tables18k[:8] * 8)
    pc = synthetic_cloud(seed=1, n_azimuth=64)
    far = pc.copy()
    far[:, :3] *= (125.0 / np.linalg.norm(far[:, :3], axis=1))[:, None]      # every return beyond the 1230-sample grid
    with pytest.raises(IndexError):                                          # simulation.py:149
        run_full(engine, tid, far, list(range(64)))
    r = run_full(engine, tid, pc, list(range(64)))                           # engine still usable afterwards
    assert set(np.unique(r['full'][:, 4])) <= {0.0, 1.0, 2.0}
    with pytest.raises(FileNotFoundError):                                   # simulation.py:329
        augment(pc, 'no_such_prefix', DIV, root_path='/nonexistent', engine=engine)
    with pytest.raises(FileNotFoundError):
        engine.snowfall_batch(tid, torch.from_numpy(pc).cuda(), np.array([0, pc.shape[0]]), np.full((1, 64), 99), DIV,
                              threshold_filter=False)
    with pytest.raises(ValueError):                                          # divergence beyond what the index was built for
        engine.snowfall_batch(tid, torch.from_numpy(pc).cuda(), np.array([0, pc.shape[0]]),
                              np.arange(64)[None], float(np.degrees(1e-2)), threshold_filter=False)
    engine.free_tables(tid)
    with pytest.raises(FileNotFoundError):
        engine.free_tables(tid)


def test_wider_beam_and_bucket_counts(engine, oracle):
    """Other beam divergences / index resolutions give the same answers (index is a pure accelerator)."""
    tables = [synthetic_particles(300 + k, 9000) for k in range(64)]
    pc = synthetic_cloud(seed=8, n_azimuth=96)
    div = float(np.degrees(6e-3))
    aug, s, nocc, theta = oracle.snow_cloud(pc, tables, list(range(64)), sensor_arrays(), div)
    aug[:, 3] = np.round(aug[:, 3])
    for nb in (512, 2048, 8192):
        tid = engine.upload_tables(tables, max_beam_divergence_rad=6e-3, n_buckets=nb)
        d_pc = torch.from_numpy(pc).cuda()
        res = engine.snowfall_batch(tid, d_pc, np.array([0, pc.shape[0]]), np.arange(64)[None], div,
                                    theta=torch.from_numpy(theta).cuda(), threshold_filter=False, want_full=True,
                                    want_nocc=True, assume_sorted=True)
        engine.check()
        assert np.array_equal(res['full'].cpu().numpy(), aug), f'n_buckets={nb}'
        assert np.array_equal(res['nocc'].cpu().numpy(), nocc)
        engine.free_tables(tid)


# ----------------------------------------------------------------------------------------------------------------------
# (3) full-size properties (BASELINE.json config 1/2: batch of 64 x 2048 clouds)
# ----------------------------------------------------------------------------------------------------------------------
def test_full_size_properties(engine, tables18k):
    B = 8
    clouds = [synthetic_cloud(seed=100 + b) for b in range(B)]
    N = clouds[0].shape[0]
    assert N == 64 * 2048
    pts = torch.from_numpy(np.concatenate(clouds)).cuda()
    off = (np.arange(B + 1) * N).astype(np.int64)
    rng = np.random.default_rng(3)
    orders = np.stack([rng.permutation(64) for _ in range(B)]).astype(np.int32)
    poly = np.tile(np.array([2e-3, -0.3, 12.0]), (B, 1))
    tid = engine.upload_tables(tables18k)
    r1 = engine.snowfall_batch(tid, pts, off, orders, DIV, thresh_poly=poly, want_full=True, want_perm=True)
    engine.check()
    r1 = {k: v.clone() for k, v in r1.items()}
    r2 = engine.snowfall_batch(tid, pts, off, orders, DIV, thresh_poly=poly, want_full=True, want_perm=True)
    engine.check()
    # determinism / idempotence of the whole pipeline
    for k in ('full', 'counts', 'stats', 'perm'):
        assert torch.equal(r1[k], r2[k]), k
    full = r1['full'].cpu().numpy().reshape(B, N, 5)
    counts = r1['counts'].cpu().numpy()
    stats = r1['stats'].cpu().numpy()
    src = np.stack(clouds)
    perm = r1['perm'].cpu().numpy().reshape(B, N)
    for b in range(B):
        lab = full[b, :, 4]
        assert set(np.unique(lab)) <= {0.0, 1.0, 2.0}
        s = src[b][perm[b]]
        assert np.array_equal(np.sort(perm[b]), np.arange(N))
        assert np.all(np.diff(s[:, 4]) >= 0)                                 # sorted by channel
        un = lab == 0
        assert np.array_equal(full[b][un][:, :3], s[un][:, :3])              # untouched beams keep xyz
        assert np.array_equal(full[b][un][:, 3], np.round(s[un][:, 3]))
        att = lab == 1
        assert np.array_equal(full[b][att][:, :3], s[att][:, :3])            # attenuated: only intensity changes
        sc = lab == 2
        d0 = np.linalg.norm(s[sc][:, :3].astype(np.float64), axis=1)
        d1 = np.linalg.norm(full[b][sc][:, :3].astype(np.float64), axis=1)
        assert np.all(d1 < d0)                                               # scattered points move towards the sensor
        cosang = np.sum(s[sc][:, :3].astype(np.float64) * full[b][sc][:, :3], axis=1) / (d0 * d1)
        assert np.all(np.abs(cosang) > 1 - 1e-6)                             # ... along the beam
        # reference quirk kept on purpose: a beam fully blocked inside the receiver's blind zone (xsi = 0 below 0.9 m)
        # has an all-zero waveform, np.argmax gives index 0 and the point lands at d_max = -c*tau/2 BEHIND the sensor
        back = cosang < 0
        assert np.allclose(d1[back], 299792458.0 * 1e-8 / 2, rtol=1e-5) and back.mean() < 0.2
        assert counts[b] + stats[b, 1] == N                                  # kept + removed == input
        kept = r1['points'].cpu().numpy()[off[b]:off[b] + counts[b]]
        assert (kept[:, 4] == 1).sum() == stats[b, 0]
        # a single-cloud call gives the same rows as the batched call (clouds are independent)
    rs = engine.snowfall_batch(tid, pts[off[3]:off[4]].contiguous(), np.array([0, N]), orders[3:4], DIV,
                               thresh_poly=poly[3:4], want_full=True)
    engine.check()
    assert torch.equal(rs['full'], r1['full'][off[3]:off[4]])
    assert int(rs['counts'][0]) == counts[3]
    frac = [(full[..., 4] == l).mean() for l in (0, 1, 2)]
    print('label fractions', frac)
    assert frac[1] > 0.05 and frac[2] > 0.005
    engine.free_tables(tid)


def test_host_pipeline_matches_device(engine, tables18k):
    """lss_snowfall_batch_host (host in/out, chunks over the engine's copy / pre-pass / beam streams) == snowfall_batch
    on device-resident input, bit for bit, for any chunking; ragged batch with an empty cloud in it."""
    B = 7
    clouds = [synthetic_cloud(seed=300 + b, n_azimuth=256 + 64 * b) for b in range(B)]
    clouds[3] = clouds[3][:0]
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
    orders = np.stack([np.random.default_rng(b).permutation(64) for b in range(B)]).astype(np.int32)
    host = torch.from_numpy(np.concatenate(clouds)).pin_memory()
    tid = engine.upload_tables(tables18k)
    poly = np.tile(np.array([1e-4, -2e-3, 0.02]), (B, 1))
    for kw in (dict(thresh_poly=poly), dict(threshold_filter=False)):
        dev = engine.snowfall_batch(tid, host.cuda(), off, orders, DIV, **kw)
        engine.check()
        dev = {k: v.cpu() for k, v in dev.items()}
        for chunks, src in ((1, host), (3, host.numpy().copy()), (7, host), (0, host), (50, host)):
            got = engine.snowfall_batch_host(tid, src, off, orders, DIV, n_chunks=chunks, **kw)
            assert torch.equal(got['counts'], dev['counts']) and torch.equal(got['stats'], dev['stats'])
            for b in range(B):
                n = int(dev['counts'][b])
                assert torch.equal(got['points'][off[b]:off[b] + n], dev['points'][off[b]:off[b] + n])
    engine.free_tables(tid)


def test_host_pipeline_device_prepass(engine, tables18k):
    """Same with the device pre-pass running on the pipeline's own streams (needs ground: no empty cloud)."""
    B = 6
    clouds = [synthetic_cloud(seed=300 + b, n_azimuth=256 + 64 * b) for b in range(B)]
    off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
    orders = np.stack([np.random.default_rng(b).permutation(64) for b in range(B)]).astype(np.int32)
    host = torch.from_numpy(np.concatenate(clouds)).pin_memory()
    tid = engine.upload_tables(tables18k)
    dev = engine.snowfall_batch(tid, host.cuda(), off, orders, DIV, device_prepass=True)
    engine.check()
    dev = {k: v.cpu() for k, v in dev.items()}
    out = {}
    for chunks in (1, 4, 6, 4):
        got = engine.snowfall_batch_host(tid, host, off, orders, DIV, device_prepass=True, n_chunks=chunks, host_out=out)
        assert torch.equal(got['counts'], dev['counts']) and torch.equal(got['stats'], dev['stats'])
        for b in range(B):
            n = int(dev['counts'][b])
            assert torch.equal(got['points'][off[b]:off[b] + n], dev['points'][off[b]:off[b] + n])
    # a device-side error inside one chunk surfaces as the reference's exception type from the synchronous call
    far = clouds[0].copy()
    far[:, :3] *= (125.0 / np.linalg.norm(far[:, :3], axis=1))[:, None]     # returns beyond the 1230-sample grid
    with pytest.raises(IndexError):
        engine.snowfall_batch_host(tid, np.concatenate([clouds[1], far]), np.array([0, len(clouds[1]), len(clouds[1]) + len(far)]),
                                   orders[:2], DIV, threshold_filter=False, n_chunks=2)
    got = engine.snowfall_batch_host(tid, host, off, orders, DIV, device_prepass=True, n_chunks=3)   # engine still usable
    assert torch.equal(got['counts'], dev['counts'])
    engine.free_tables(tid)


def test_host_pipeline_batches_in_flight(engine, tables18k):
    """submit / wait: three batches in flight at once (one of them failing on the device) give the same results as the
    synchronous call; the error belongs to the batch that caused it; a fourth submit or a second wait is refused."""
    tid = engine.upload_tables(tables18k)
    batches = []
    for k in range(3):
        clouds = [synthetic_cloud(seed=700 + 10 * k + b, n_azimuth=192 + 64 * k) for b in range(3 + k)]
        if k == 1:
            far = clouds[1]
            far[:, :3] *= (125.0 / np.linalg.norm(far[:, :3], axis=1))[:, None]         # IndexError batch
        off = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
        orders = np.stack([np.random.default_rng(50 + b).permutation(64) for b in range(len(clouds))]).astype(np.int32)
        batches.append((torch.from_numpy(np.concatenate(clouds)).pin_memory(), off, orders))
    poly = np.array([1e-4, -2e-3, 0.02])
    want = []
    for k, (host, off, orders) in enumerate(batches):
        if k == 1:
            want.append(None)
            continue
        r = engine.snowfall_batch(tid, host.cuda(), off, orders, DIV, thresh_poly=np.tile(poly, (len(off) - 1, 1)))
        engine.check()
        want.append({n: v.cpu() for n, v in r.items()})
    for rep in range(2):                                        # second round reuses the slots
        tickets = [engine.snowfall_batch_host_submit(tid, host, off, orders, DIV, n_chunks=2,
                                                     thresh_poly=np.tile(poly, (len(off) - 1, 1)))
                   for host, off, orders in batches]
        with pytest.raises(ValueError):                         # a fourth batch needs a wait first
            engine.snowfall_batch_host_submit(tid, *batches[0], DIV, n_chunks=1, threshold_filter=False)
        for k in (0, 2, 1):                                     # any order
            if k == 1:
                with pytest.raises(IndexError):
                    engine.snowfall_batch_host_wait(tickets[k])
                continue
            got = engine.snowfall_batch_host_wait(tickets[k])
            off = batches[k][1]
            assert torch.equal(got['counts'], want[k]['counts']) and torch.equal(got['stats'], want[k]['stats'])
            for b in range(len(off) - 1):
                n = int(got['counts'][b])
                assert torch.equal(got['points'][off[b]:off[b] + n], want[k]['points'][off[b]:off[b] + n])
        with pytest.raises(ValueError):
            engine.snowfall_batch_host_wait(tickets[0])
    engine.free_tables(tid)


def test_degenerate_rows(engine, oracle, tables18k):
    """Rows at the origin, NaN coordinates, zero / negative intensity: same answer as the oracle, no crash."""
    pc = synthetic_cloud(seed=41, n_azimuth=64)
    pc[3, :3] = 0.0                      # at the sensor: range 0, no particle is nearer
    pc[7, 0] = np.nan
    pc[11, 1] = np.nan
    pc[13, 3] = 0.0
    pc[17, 3] = -5.0
    pc[19, :3] = [1e-3, -1e-3, 2e-3]
    order = list(range(64))
    aug, s, nocc, theta = oracle.snow_cloud(pc, tables18k, order, sensor_arrays(), DIV)
    aug[:, 3] = np.round(aug[:, 3])
    tid = engine.upload_tables(tables18k)
    r = run_full(engine, tid, pc, order, theta=theta)
    got, want = r['full'], aug
    same = np.isnan(want) & np.isnan(got) | (want == got)
    assert same.all()
    engine.free_tables(tid)


# ----------------------------------------------------------------------------------------------------------------------
# (4) many occluders on one beam: the solve kernel's deferral to the overflow kernel, and the hard cap
# ----------------------------------------------------------------------------------------------------------------------
def _column_of_flakes(n, seed):
    """n small disks strung along azimuth ~0 between 2 and 45 m, plus background flakes elsewhere."""
    rng = np.random.default_rng(seed)
    r = np.sort(rng.uniform(10.0, 28.0, n))
    col = np.column_stack((r, rng.uniform(-1.2e-3, 1.2e-3, n) * r, rng.uniform(1e-4, 3e-4, n)))    # ~1e-5 rad wide each
    return np.vstack((col, synthetic_particles(seed, 3000)))


def test_beams_with_dozens_of_occluders_match_the_oracle(engine, oracle):
    """40 and 100 occluders on one beam: more than the solve kernel's shared-memory arena takes per beam (63), so the
    second case is deferred to the overflow kernel -- both must equal the oracle (labels, intensities, occluder counts)."""
    fd, fs, mi, mx = sensor_arrays()
    for n_col, seed in ((40, 21), (100, 22)):
        table = _column_of_flakes(n_col, seed)
        az = np.concatenate(([0.0, 1e-4, -2e-4, 3e-4], np.linspace(-np.pi, np.pi, 60, endpoint=False)))
        d = np.concatenate(([50.0, 48.0, 60.0, 30.0], np.full(60, 35.0)))
        pts = np.stack([d * np.cos(az), d * np.sin(az), np.zeros_like(d), np.full_like(d, 90.0), np.full_like(d, 5.0)],
                       axis=1).astype(np.float32)
        theta = np.arctan2(pts[:, 1], pts[:, 0])
        want, s, nocc, _ = oracle.snow_channel(pts, table, DIV, fd[5], fs[5], mi[5], mx[5], theta=theta)
        assert nocc.max() >= n_col * 0.6, 'test set-up: most flakes of the column must claim a piece of the first beams'
        tid = engine.upload_tables([table] * 64)
        r = run_full(engine, tid, pts, list(range(64)), theta=theta)
        engine.free_tables(tid)
        assert np.array_equal(r['full'], want), f'{n_col} flakes'
        assert np.array_equal(r['nocc'], nocc)
        assert np.isclose(r['stats'][0, 3], s, rtol=1e-12, atol=0)


def test_more_than_128_occluders_is_an_error(engine):
    """The engine's only hard cap (LSS_ERR_OCCLUDER_OVERFLOW, no reference analogue: the reference's lists are unbounded;
    the surveyed densities give at most 14-28 occluders per beam)."""
    table = _column_of_flakes(400, 23)
    pts = np.array([[50.0, 0.0, 0.0, 90.0, 5.0], [0.0, 30.0, 0.0, 80.0, 5.0]], dtype=np.float32)
    tid = engine.upload_tables([table] * 64)
    d_pc = torch.from_numpy(pts).cuda()
    engine.snowfall_batch(tid, d_pc, np.array([0, 2], dtype=np.int64), np.arange(64, dtype=np.int32)[None], DIV,
                          threshold_filter=False)
    with pytest.raises(RuntimeError, match='occluders'):
        engine.check()
    engine.free_tables(tid)
    engine.check()                                         # the latched status is cleared by the failing check


def test_augment_snowfall_rate_signature(engine, tmp_path, monkeypatch):
    """The north-star call shape augment_snowfall(pc, snowfall_rate, terminal_velocity, mode): prefix derived like the
    reference's callers do (precompute.py:57-58,101), tables sampled once, written under the reference's file names
    and found there by plain augment() afterwards."""
    from lidar_snow_sim_b200.snowfall import simulation as sim
    from lidar_snow_sim_b200.snowfall.sampling import particle_file_prefix, sample_table_set
    monkeypatch.setenv('LSS_NPY_DIR', str(tmp_path))
    pc = synthetic_cloud(seed=9, n_azimuth=256)
    order = np.random.default_rng(3).permutation(64).tolist()
    poly = np.array([1e-3, -0.2, 14.0])
    s1, a1 = sim.augment_snowfall(pc, 2.5, 1.6, 'gunn', only_camera_fov=False, engine=engine, write_tables=True,
                                  order=order, thresh_poly=poly)
    prefix = particle_file_prefix('gunn', 2.5, 1.6)
    assert sorted(p.name for p in tmp_path.iterdir()) == sorted(f'{prefix}_{k}.npy' for k in range(1, 65))
    s2, a2 = sim.augment(pc, prefix, DIV, only_camera_fov=False, engine=engine, order=order, thresh_poly=poly)
    s3, a3 = sim.augment(pc, 'unused', DIV, only_camera_fov=False, engine=engine, order=order, thresh_poly=poly,
                         tables=sample_table_set('gunn', 2.5, 1.6, seed=1000))
    assert s1 == s2 == s3 and np.array_equal(a1, a2) and np.array_equal(a1, a3) and a1.shape[0] > 0
    with pytest.raises(FileNotFoundError):
        sim.augment(pc, 'gunn_1.0_2.0', DIV, engine=engine)


def test_device_azimuth_is_the_rounded_float64_atan2(engine):
    """The kernels' beam azimuth (fast table + series path with a library fall-back near float32 rounding boundaries) equals
    float32(atan2(float64 y, float64 x)) bit for bit -- on the bench cloud, on random arguments and on special values."""
    import ctypes
    from lidar_snow_sim_b200.engine import _ptr
    rng = np.random.default_rng(3)
    pc = synthetic_cloud(seed=8, n_azimuth=2048)
    xs = [pc[:, 0], rng.uniform(-120, 120, 2_000_000).astype(np.float32), (10.0 ** rng.uniform(-20, 20, 200_000)).astype(np.float32),
          np.array([0, 0, 1, -1, 0.0, -0.0, 1e-30, 3e38, np.inf, -np.inf, np.nan, 1, 1, -1, 5, -5], dtype=np.float32)]
    ys = [pc[:, 1], rng.uniform(-120, 120, 2_000_000).astype(np.float32), (-(10.0 ** rng.uniform(-20, 20, 200_000))).astype(np.float32),
          np.array([0, 1, 0, 0, -0.0, -0.0, 1e-30, 3e38, np.inf, 1, 1, np.nan, 1, -1, 5e-8, -5e-8], dtype=np.float32)]
    x = np.concatenate(xs)
    y = np.concatenate(ys)
    d_x, d_y = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    out = torch.empty_like(d_x)
    st = engine.lib.lss_debug_azimuth(engine.h, _ptr(d_y), _ptr(d_x), x.shape[0], _ptr(out), engine._stream())
    assert st == 0
    engine.check()
    got = out.cpu().numpy()
    with np.errstate(invalid='ignore'):
        want = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32)
    same = (got.view(np.int32) == want.view(np.int32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (int((~same).sum()), x[~same][:5], y[~same][:5], got[~same][:5], want[~same][:5])


@pytest.mark.gpu
def test_gather_push_writes_kept_rows_into_every_peer_buffer(engine):
    """lss_gather_push on one GPU: the 'peers' are three separate buffers of the same device.  Ragged clouds
    whose offsets are not multiples of four rows (16-byte misalignment), counts below the slot sizes, an empty cloud."""
    eng = engine
    dev = eng.device
    rng = np.random.default_rng(5)
    sizes = [1000, 0, 37, 4099, 2, 513]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n_rows, B, world = int(off[-1]), len(sizes), 3
    counts = np.array([rng.integers(0, s + 1) for s in sizes], dtype=np.int32)
    counts[3] = sizes[3]
    pts = torch.from_numpy(rng.normal(size=(n_rows, 5)).astype(np.float32)).to(dev)
    d_counts = torch.from_numpy(counts).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    for rank in range(world):
        for use_counts in (True, False):
            peers = [torch.full((world * n_rows, 5), -7.0, dtype=torch.float32, device=dev) for _ in range(world)]
            pcnt = [torch.full((world * B,), -7, dtype=torch.int32, device=dev) for _ in range(world)]
            eng.gather_push(pts, d_counts if use_counts else None, d_off, n_rows, world, rank, peers, pcnt, blocks=3 + rank)
            torch.cuda.synchronize(dev)
            want = np.full((world * n_rows, 5), -7.0, dtype=np.float32)
            wcnt = np.full((world * B,), -7, dtype=np.int32)
            src = pts.cpu().numpy()
            for b in range(B):
                c = int(counts[b]) if use_counts else sizes[b]
                want[rank * n_rows + off[b]: rank * n_rows + off[b] + c] = src[off[b]: off[b] + c]
                wcnt[rank * B + b] = c
            for p in range(world):
                assert np.array_equal(peers[p].cpu().numpy(), want)
                assert np.array_equal(pcnt[p].cpu().numpy(), wcnt)
