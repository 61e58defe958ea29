"""
Why the solve kernel's argmax is exact: a CPU model of what k_solve (lidar_snow_sim_b200/csrc/solve.cu) evaluates --
pieces of constant pulse set, three samples around the analytic peak of each piece, pieces pruned against the largest
pulse -- against the reference's way (sum every window sample by sample, np.argmax over the 1230-sample grid,
tools/snowfall/simulation.py:118-153).  The kernel itself is compared with the oracle bit for bit in the `-m gpu` tests; this
test pins the MATHEMATICS those shortcuts rest on, on thousands of random pulse sets (amplitude ratios down to 1e-12, zero
amplitudes, pulses at both ends of the grid, cancelling pairs half a period apart, the hard target's float32 window
arithmetic), with no GPU.
"""
import math

import numpy as np

from oracle import oracle as orc

CTAU = 299792458.0 * 1e-8
M = 1230
R = orc.range_grid()
INV_STEP = (M - 1) / (120 + CTAU)
_A = np.fmod(R / CTAU, 2.0)
SIN_T, COS_T = np.sin(np.pi * _A), np.cos(np.pi * _A)          # the kernel's phase table (host_phase_table, api.cu)


def reference_argmax(pulses):
    """simulation.py:137-153: i[ks:ke] += A sin^2(pi (R - r) / (c tau)) per pulse in dict order, first index of the maximum."""
    wave = np.zeros(M)
    for amp, r, ks, ke in pulses:
        b = r / CTAU
        sb, cb = math.sin(math.pi * b), math.cos(math.pi * b)
        k = np.arange(ks, ke)
        sn = SIN_T[k] * cb - COS_T[k] * sb
        wave[k] += amp * (sn * sn)
    k = int(np.argmax(wave))
    return k, wave[k]


def kernel_model_argmax(pulses):
    """The sweep of k_solve: pieces [k, pend) with active pulses qa..qb; a piece is evaluated only if the sum of its active
    amplitudes can reach 0.99 x the largest amplitude; per piece the samples kc-1..kc+1 (clipped) around the analytic peak."""
    ext = [(amp, r, ks, ke, math.sin(math.pi * (r / CTAU)), math.cos(math.pi * (r / CTAU))) for amp, r, ks, ke in pulses]
    n = len(ext)
    lb = 0.99 * max(p[0] for p in ext)
    best, kbest = 0.0, 0
    kept = pieces = 0
    qa, qb, nxt, asum = 0, -1, 0, 0.0
    k = ext[0][2]
    while True:
        while nxt < n and ext[nxt][2] <= k:
            qb = nxt
            asum += ext[qb][0]
            nxt += 1
        while qa <= qb and ext[qa][3] <= k:
            asum -= ext[qa][0]
            qa += 1
        if qa > qb:
            asum = 0.0
            if nxt >= n:
                break
            k = ext[nxt][2]
            continue
        pend = min(ext[qa][3], ext[nxt][2] if nxt < n else 1 << 30)
        lo, hi = k, pend
        pieces += 1
        if asum * 1.0001 >= lb:
            kept += 1
            act = ext[qa:qb + 1]
            whole = False
            if len(act) == 1:
                kc = int(np.rint((act[0][1] + CTAU / 2) * INV_STEP))
            else:                                                    # float32, like the kernel: it only SELECTS samples
                zx = zy = a32 = np.float32(0)
                for amp, _, _, _, sb, cb in act:
                    a_, s_, c_ = np.float32(amp), np.float32(sb), np.float32(cb)
                    zx += a_ * (c_ * c_ - s_ * s_)
                    zy += a_ * (np.float32(2) * s_ * c_)
                    a32 += a_
                whole = float(zx * zx + zy * zy) < 1e-6 * float(a32 * a32)
                r0 = (float(np.arctan2(zy, zx)) + math.pi) * CTAU / (2 * math.pi)
                rc = 0.5 * (lo + hi - 1) / INV_STEP
                kc = int(np.rint((r0 + round((rc - r0) / CTAU) * CTAU) * INV_STEP))
            c_lo = lo if whole else max(lo, min(hi - 1, kc - 1))
            c_hi = hi - 1 if whole else min(hi - 1, max(lo, kc + 1))
            for c in range(c_lo, c_hi + 1):
                v = 0.0
                for amp, _, _, _, sb, cb in act:                     # dict order, like the reference's i[k] +=
                    sn = SIN_T[c] * cb - COS_T[c] * sb
                    v += amp * (sn * sn)
                if v > best:                                         # pieces and samples ascend: the first maximum stays
                    best, kbest = v, c
        k = pend
    return kbest, best, kept, pieces


def random_pulses(rng):
    n = int(rng.integers(1, 9))
    mode = rng.random()
    base = rng.uniform(0.95, 119.9) if mode < 0.8 else (rng.uniform(0.9, 1.5) if mode < 0.9 else rng.uniform(110, 119.9))
    gaps = rng.uniform(0.0, 3.2, n - 1) if rng.random() < 0.7 else rng.uniform(0, 0.3, n - 1)
    rs = np.concatenate([[base], base + np.cumsum(gaps)])
    rs = rs[rs <= 119.95]
    pulses = []
    for q, r in enumerate(rs):
        u = rng.random()
        amp = rng.uniform(0.001, 50.0) if u < 0.6 else (10.0 ** rng.uniform(-12, 2) if u < 0.9 else (0.0 if u < 0.93 else 1.0))
        if q == len(rs) - 1 and rng.random() < 0.5:                  # hard target: float32 range and window arithmetic
            d32 = np.float32(r)
            ks = int(np.ceil(np.float32(d32 * np.float32(10.0))))
            ke = int(np.floor(np.float32((d32 + np.float32(CTAU)) * np.float32(10.0))) + np.float32(1.0))
            r = float(d32)
        else:
            ks, ke = int(math.ceil(r * 10)), int(math.floor((r + CTAU) * 10) + 1)
        if ke <= M:
            pulses.append((amp, float(r), ks, ke))
    if len(pulses) >= 2 and rng.random() < 0.1:                      # equal amplitudes half a period apart: the pulses cancel
        amp, r = pulses[0][0], pulses[0][1] + CTAU / 2
        if math.floor((r + CTAU) * 10) + 1 <= M:
            pulses[1] = (amp, r, int(math.ceil(r * 10)), int(math.floor((r + CTAU) * 10) + 1))
    pulses.sort(key=lambda p: (p[2], p[3]))
    if any(pulses[i][3] > pulses[i + 1][3] for i in range(len(pulses) - 1)):
        return []                                                    # the kernel's pulses ascend in range: so do both window ends
    return pulses


def test_pieces_three_samples_and_amplitude_pruning_find_the_reference_argmax():
    rng = np.random.default_rng(2024)
    trials = kept = pieces = 0
    worst = np.inf
    while trials < 20000:
        pulses = random_pulses(rng)
        if not pulses:
            continue
        trials += 1
        k_ref, v_ref = reference_argmax(pulses)
        k_mod, v_mod, nk, npc = kernel_model_argmax(pulses)
        assert (k_mod, v_mod) == (k_ref, v_ref), pulses
        kept += nk
        pieces += npc
        amax = max(p[0] for p in pulses)
        if amax > 0:
            worst = min(worst, v_ref / amax)
    # the bound the pruning rests on: the waveform's maximum is at least 0.99 x the largest pulse amplitude
    # (cos^2 of half a grid step + the grid's 0.005 m rounding = 0.9967)
    assert worst >= 0.99
    assert kept < 0.7 * pieces                                       # and the rule does prune
