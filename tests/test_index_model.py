"""
Why the 8-byte candidate index is a pure accelerator: a CPU model of its encoding (k_particle_records / k_fill_entries,
lidar_snow_sim_b200/csrc/tables.cu; lss_decode, csrc/common.cuh) and of the scan kernel's float32 broad phase (k_scan,
csrc/solve.cu) against the reference's exact occlusion test (tools/snowfall/simulation.py:345-385, restated like exact_hit in
solve.cu).  For random planes of disks and random beams -- seam beams, disks next to the sensor, wide disks, targets a hair
beyond a disk -- every particle the exact test accepts must (1) be registered in the ONE azimuth bucket the beam reads and
(2) pass the quantised float32 range / azimuth / half-width comparison of that bucket's entry.  The kernel itself is
compared with the oracle bit for bit in the `-m gpu` tests (also across bucket counts); this pins the arithmetic the
quantisation rests on, without a GPU.
"""
import numpy as np

TWO_PI = 6.283185307179586
PI = 3.141592653589793
ANG_MARGIN = 1e-5                                   # LSS_ANG_MARGIN
RHO_UNIT = np.float32(0.0025)                       # LSS_RHO_UNIT
PHI_UNIT = 9.587672516830327e-05                    # LSS_PHI_UNIT = pi / 32767


def build_entries(x, y, r, n_buckets, max_div):
    """tables.cu: per particle (phi, rho, alpha), the span of buckets it is registered in and its encoded entry fields."""
    rho = np.sqrt(x * x + y * y)
    phi = np.arctan2(y, x)
    phi = np.where(phi < 0, phi + TWO_PI, phi)
    alpha = np.arcsin(r / rho)
    half_div_margin = max_div / 2 + ANG_MARGIN
    zbase = np.float32(max_div / 2 + ANG_MARGIN)
    w = TWO_PI / n_buckets
    hw = alpha + half_div_margin
    everywhere = 2 * hw + 2 * w >= TWO_PI
    blo = np.floor((phi - hw) / w).astype(np.int64)
    bhi = np.floor((phi + hw) / w).astype(np.int64)
    n_span = np.where(everywhere, n_buckets, np.minimum(bhi - blo + 1, n_buckets))
    lo = np.where(everywhere, 0, np.mod(blo, n_buckets))
    rq = np.clip(np.floor(rho * 400.0).astype(np.int64) - 1, 0, 65535)
    want = (alpha + half_div_margin + 0.5 * PHI_UNIT)
    want32 = want.astype(np.float32)
    want32 = np.where(want32.astype(np.float64) < want, np.nextafter(want32, np.float32(np.inf)), want32)   # __double2float_ru
    zc = np.maximum(np.ceil(32.0 * np.log2(want32.astype(np.float64) / float(zbase))).astype(np.int64), 0)
    for _ in range(4):                              # the kernel's fix-up loop against the float32 decode
        dec = zbase * np.exp2(zc.astype(np.float32) * np.float32(1.0 / 32.0)).astype(np.float32)
        zc = np.where((dec < want32) & (zc < 1023), zc + 1, zc)
    return dict(rho=rho, phi=phi, alpha=alpha, lo=lo, n_span=n_span, rq=rq, zc=zc, zbase=zbase, w=w)


def exact_hit(tab, d, right, left):
    """solve.cu exact_hit == simulation.py:345-385 (float64)."""
    phi, alpha = tab['phi'], tab['alpha']
    straddle = right > left
    inside = (right <= phi) & (phi <= left)
    if straddle:
        inside |= ((right - TWO_PI <= phi) & (phi <= left)) | ((right <= phi) & (phi <= left + TWO_PI))

    def within(diff, tol):
        return (np.abs(diff) < tol) | (np.abs(diff - TWO_PI) < tol) | (np.abs(diff + TWO_PI) < tol)
    return (tab['rho'] < d) & (inside | within(right - phi, alpha) | within(left - phi, alpha))


def broad_phase(tab, n_buckets, th32, d32):
    """k_scan phase A for one beam: which particles are in the beam's bucket, and which of those pass the float32 test."""
    thd = float(th32)
    thm = thd - TWO_PI if thd >= TWO_PI else thd
    bk = min(max(int(thm * (n_buckets / TWO_PI)), 0), n_buckets - 1)
    th_rel = np.float32(thm - (bk + 0.5) * tab['w'])
    in_bucket = np.mod(bk - tab['lo'], n_buckets) < tab['n_span']
    centre = (bk + 0.5) * tab['w']
    rel = tab['phi'] - centre
    rel = np.where(rel > PI, rel - TWO_PI, rel)
    rel = np.where(rel <= -PI, rel + TWO_PI, rel)
    pq = np.rint(rel / PHI_UNIT)
    en_x = tab['rq'].astype(np.float32) * RHO_UNIT
    en_y = pq.astype(np.float32) * np.float32(PHI_UNIT)
    en_z = (tab['zbase'] * np.exp2(tab['zc'].astype(np.float32) * np.float32(1.0 / 32.0))).astype(np.float32)
    passes = (en_x < d32) & (np.abs(en_y - th_rel) <= en_z)
    return in_bucket, passes, en_x


def test_broad_phase_of_the_compact_index_is_a_superset_of_the_exact_test():
    rng = np.random.default_rng(77)
    hits = survivors = 0
    for n_buckets, max_div, div in ((2048, 3e-3, 3e-3), (512, 3e-3, 3e-3), (8192, 3e-3, 1.5e-3), (2048, 6e-3, 6e-3)):
        n_p = 4000
        rho = np.concatenate([rng.uniform(0.2, 80.0, n_p - 400), rng.uniform(0.05, 0.3, 200), rng.uniform(80.0, 200.0, 200)])
        ang = rng.uniform(0, TWO_PI, n_p)
        ang[:60] = rng.choice([1e-7, TWO_PI - 1e-7, 1e-4, TWO_PI - 1e-4], 60)          # disks on the seam
        r = np.concatenate([rng.uniform(1e-4, 3e-3, n_p - 200), rng.uniform(5e-3, 4e-2, 200)])
        r = np.minimum(r, 0.5 * rho)
        tab = build_entries(rho * np.cos(ang), rho * np.sin(ang), r, n_buckets, max_div)
        for trial in range(1500):
            mode = rng.random()
            if mode < 0.15:                                                              # seam beams
                th = np.float32(rng.choice([0.0, 1e-4, TWO_PI - 1e-4, 1.4e-3, TWO_PI - 1.4e-3]))
            elif mode < 0.55:                                                            # aimed at a disk's rim / centre
                q = int(rng.integers(0, n_p))
                th = np.float32(np.mod(tab['phi'][q] + rng.choice([-1, 0, 1]) * (tab['alpha'][q] + div / 2) *
                                       rng.uniform(0.98, 1.02), TWO_PI))
            else:
                th = np.float32(rng.uniform(0, TWO_PI))
            if th < 0:
                th = np.float32(th + np.float32(6.2831855))
            d32 = np.float32(rng.uniform(0.3, 120.0))
            if mode < 0.55 and rng.random() < 0.5:                                       # target a hair beyond a disk
                q = int(rng.integers(0, n_p))
                d32 = np.nextafter(np.float32(tab['rho'][q]), np.float32(np.inf))
            thd = float(th)
            right, left = thd - div / 2, thd + div / 2
            right = right + TWO_PI if right < 0 else right
            left = left + TWO_PI if left < 0 else left
            right = right - TWO_PI if right > TWO_PI else right
            left = left - TWO_PI if left > TWO_PI else left
            ex = exact_hit(tab, float(d32), right, left)
            in_bucket, passes, en_x = broad_phase(tab, n_buckets, th, d32)
            assert not (ex & ~in_bucket).any(), 'an occluder is not registered in the bucket the beam reads'
            assert not (ex & ~passes).any(), 'the float32 broad phase rejects an occluder'
            # early exit of the walk: entries are sorted by the quantised range, which stays strictly below the true one
            assert (en_x.astype(np.float64) < tab['rho']).all()
            hits += int(ex.sum())
            survivors += int((in_bucket & passes).sum())
    assert hits > 2000                                   # the beams do meet disks
    assert survivors < 1.25 * hits                       # and the broad phase is tight, not merely safe (measured: 1.11)
