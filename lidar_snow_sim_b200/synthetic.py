"""
Seeded synthetic HDL-64E-shaped clouds for tests and bench.py -- there is no dataset on the GPU box.

64 channels x `n_azimuth` azimuth steps.  Elevation of channel c = HDL-64E S3 vert_correction; azimuth_k = -pi + 2 pi k / n.
Range = min(ground hit at sensor height 1.7 m, wall range ~U(20, 60) m, 100 m cap), clipped to [1.5, 110] m so that
every return stays inside the reference's 0..120 m waveform grid (tools/snowfall/simulation.py:106-116).
Intensity: ground round(clip(60 - 0.6 d + N(0, 4), 1, 255)), walls round(U(5, 200)).
Layout: float32 (N, 5) = x, y, z, intensity, channel -- the STF / reference layout (precompute.py:78).
Rows are emitted channel-major, azimuth ascending; `drop` removes a seeded random fraction (config "STF-shaped").
"""
import numpy as np

from .calib.hdl64e_s3 import vert_corrections


def synthetic_cloud(seed=0, n_azimuth=2048, drop=0.0, sensor_height=1.7, shuffle_rows=False):
    rng = np.random.default_rng(seed)
    elev = vert_corrections()                                   # rad, (64,)
    n_ch = elev.shape[0]
    az = -np.pi + 2 * np.pi * np.arange(n_azimuth) / n_azimuth
    # walls: piecewise-constant range over 32 azimuth sectors (so neighbouring beams see similar targets)
    n_sectors = 32
    wall_sector = rng.uniform(20.0, 60.0, n_sectors)
    wall = wall_sector[(np.arange(n_azimuth) * n_sectors) // n_azimuth]
    wall = wall[None, :] + rng.normal(0.0, 0.05, (n_ch, n_azimuth))
    with np.errstate(divide='ignore'):
        ground = np.where(elev < 0, sensor_height / np.sin(-elev), np.inf)[:, None] * np.ones((1, n_azimuth))
    ground = ground * (1.0 + rng.normal(0.0, 0.002, (n_ch, n_azimuth)))
    rng_wall = wall / np.cos(elev)[:, None]
    d = np.minimum(np.minimum(ground, rng_wall), 100.0)
    d = np.clip(d, 1.5, 110.0)
    is_ground = ground <= np.minimum(rng_wall, 100.0)
    i_ground = np.round(np.clip(60.0 - 0.6 * d + rng.normal(0.0, 4.0, d.shape), 1, 255))
    i_wall = np.round(rng.uniform(5.0, 200.0, d.shape))
    inten = np.where(is_ground, i_ground, i_wall)
    ce = np.cos(elev)[:, None]
    se = np.sin(elev)[:, None]
    x = d * ce * np.cos(az)[None, :]
    y = d * ce * np.sin(az)[None, :]
    z = d * se * np.ones((1, n_azimuth))
    ch = np.arange(n_ch, dtype=np.float64)[:, None] * np.ones((1, n_azimuth))
    pc = np.stack([x, y, z, inten, ch], axis=-1).reshape(-1, 5).astype(np.float32)
    if drop > 0:
        keep = rng.uniform(size=pc.shape[0]) >= drop
        pc = pc[keep]
    if shuffle_rows:
        pc = pc[rng.permutation(pc.shape[0])]
    return np.ascontiguousarray(pc)


def synthetic_particles(seed, n_particles, R_0=80.0, mean_diameter_mm=2.0):
    """
    Cheap seeded stand-in for one plane of a snowflake table (x, y, r) float64 -- uniform in the disk of radius R_0,
    exponential diameters truncated at 20 mm, random slice height (same marginals as sampling.py:145-163) but WITHOUT
    the non-overlap rejection.  Used for parity fixtures, where any table works as long as both sides see the same one.
    """
    rng = np.random.default_rng(seed)
    length = np.sqrt(rng.uniform(0, R_0 ** 2, n_particles))
    angle = rng.uniform(0, 2, n_particles) * np.pi
    dia = np.minimum(rng.exponential(mean_diameter_mm, n_particles), 20.0) / 1000.0
    height = rng.uniform(-0.5, 0.5, n_particles) * dia
    r = np.sqrt((dia / 2) ** 2 - height ** 2)
    x = length * np.cos(angle)
    y = length * np.sin(angle)
    ok = x ** 2 + y ** 2 > r ** 2
    ok &= r > 0
    return np.ascontiguousarray(np.column_stack((x, y, r))[ok])
