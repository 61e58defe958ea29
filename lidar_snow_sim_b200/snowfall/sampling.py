"""
Snow physics helpers with the reference's names and argument meaning (tools/snowfall/sampling.py:23-87).
Scalar host code: these are evaluated once per (snowfall_rate, terminal_velocity) to name / size a particle table.
"""
import numpy as np


def compute_occupancy(snowfall_rate: float, terminal_velocity: float, snow_density: float = 0.1) -> float:
    """Occupancy ratio of the medium (sampling.py:23-32).  snowfall_rate [mm/h], terminal_velocity [m/s],
    snow_density [g/cm^3]."""
    water_density = 1.0
    return (water_density * snowfall_rate) / ((3.6 * 10 ** 6) * (snow_density * terminal_velocity))


def rainfall_rate_to_snowfall_rate(rainfall_rate: float, terminal_velocity: float,
                                   snowflake_density: float = 0.1, snowflake_diameter: float = 0.003) -> float:
    """sampling.py:35-52"""
    return 487 * snowflake_density * snowflake_diameter * terminal_velocity * (rainfall_rate ** (2 / 3))


def snowfall_rate_to_rainfall_rate(snowfall_rate: float, terminal_velocity: float,
                                   snowflake_density: float = 0.1, snowflake_diameter: float = 0.003) -> float:
    """sampling.py:55-69"""
    return np.sqrt((snowfall_rate / (487 * snowflake_density * snowflake_diameter * terminal_velocity)) ** 3)


def sekhon_srivastava(precipitation_rate: float) -> float:
    """Rate parameter [1/cm] of the snowflake diameter distribution, Sekhon & Srivastava 1970 (sampling.py:72-78)."""
    return 22.9 * precipitation_rate ** -0.45


def gunn_marshall(precipitation_rate: float) -> float:
    """Rate parameter [1/cm], Gunn & Marshall 1958 (sampling.py:81-87)."""
    return 25.5 * precipitation_rate ** -0.48


def particle_file_prefix(mode: str, snowfall_rate: float, terminal_velocity: float) -> str:
    """The '<mode>_<rain_rate>_<occupancy>' prefix callers build (precompute.py:101, pointcloud_viewer.py:2798-2802)."""
    rain_rate = snowfall_rate_to_rainfall_rate(float(snowfall_rate), float(terminal_velocity))
    occupancy = compute_occupancy(float(snowfall_rate), float(terminal_velocity))
    return f'{mode}_{rain_rate}_{occupancy}'


# ----------------------------------------------------------------------------------------------------------------------
# particle tables
# ----------------------------------------------------------------------------------------------------------------------
_DIST = {'gunn': 0, 'sekhon': 1}


def _pcg_state(rng):
    st = rng.bit_generator.state
    if st['bit_generator'] != 'PCG64':
        raise NotImplementedError('dart_throwing needs a PCG64-backed numpy Generator (np.random.default_rng)')
    s, inc = st['state']['state'], st['state']['inc']
    mask = (1 << 64) - 1
    return st, np.array([s >> 64, s & mask, inc >> 64, inc & mask], dtype=np.uint64)


def _expected_capacity(occupancy_ratio, precipitation_rate, R_0, distribution):
    """Upper bound on the number of accepted disks: total area / mean disk area, with head-room."""
    rate = gunn_marshall(precipitation_rate) if distribution == 'gunn' else sekhon_srivastava(precipitation_rate)
    mean_d = 10.0 / rate / 1000.0                    # mean diameter [m]
    mean_area = np.pi * (2.0 / 3.0) * mean_d ** 2 / 2.0     # E[pi (d^2/4 - h^2)] = pi d^2/6, E[d^2] = 2 mean^2
    n = occupancy_ratio * np.pi * R_0 ** 2 / max(mean_area, 1e-12)
    return int(n * 1.5) + 4096


def dart_throwing(occupancy_ratio: float, precipitation_rate: float, R_0: float, rng: np.random.Generator,
                  distribution: str = 'sekhon_srivastava', show_progessbar: bool = False) -> np.ndarray:
    """
    Same call as the reference's dart_throwing (sampling.py:90-194): N-by-3 float64 array of disks (x, y, r) [m].
    Runs in the engine's native sampler and advances `rng` exactly as the reference would have.
    """
    import ctypes
    from .. import _lib
    if distribution not in _DIST:
        raise NotImplementedError('Distribution model unknown.')            # sampling.py:113
    lib = _lib.load()
    st, words = _pcg_state(rng)
    cap = _expected_capacity(occupancy_ratio, precipitation_rate, R_0, distribution)
    while True:
        w = words.copy()
        out = np.empty((cap, 3), dtype=np.float64)
        n = ctypes.c_int64(0)
        rc = lib.lss_dart_throwing(float(occupancy_ratio), float(precipitation_rate), float(R_0), _DIST[distribution],
                                   ctypes.c_void_p(w.ctypes.data), ctypes.c_void_p(out.ctypes.data), cap,
                                   ctypes.byref(n))
        if rc == _lib.LSS_ERR_WORKSPACE:
            cap *= 2
            continue
        _lib.check(rc)
        break
    st['state']['state'] = (int(w[0]) << 64) | int(w[1])
    rng.bit_generator.state = st
    return out[:n.value].copy()


def sample_table_set(mode: str, snowfall_rate: float, terminal_velocity: float, seed: int = 1000, R_0: float = 80.0,
                     n_planes: int = 64, n_threads: int = 0):
    """
    The 64 planes of one (snowfall_rate, terminal_velocity) configuration (sampling.py:360-413), plane k drawn from
    np.random.default_rng(seed + k).  Returns a list of (Np_k, 3) float64 arrays.
    """
    import ctypes
    from .. import _lib
    lib = _lib.load()
    occ = compute_occupancy(float(snowfall_rate), float(terminal_velocity))
    rr = float(snowfall_rate_to_rainfall_rate(float(snowfall_rate), float(terminal_velocity)))
    states = np.stack([_pcg_state(np.random.default_rng(seed + k))[1] for k in range(n_planes)])
    cap = _expected_capacity(occ, rr, R_0, mode)
    while True:
        w = np.ascontiguousarray(states.copy())
        out = np.empty((n_planes, cap, 3), dtype=np.float64)
        counts = np.zeros(n_planes, dtype=np.int64)
        rc = lib.lss_dart_throwing_planes(n_planes, occ, rr, float(R_0), _DIST[mode], ctypes.c_void_p(w.ctypes.data),
                                          ctypes.c_void_p(out.ctypes.data), cap, ctypes.c_void_p(counts.ctypes.data),
                                          int(n_threads))
        if rc == _lib.LSS_ERR_WORKSPACE:
            cap *= 2
            continue
        _lib.check(rc)
        break
    return [out[k, :counts[k]].copy() for k in range(n_planes)]


# ----------------------------------------------------------------------------------------------------------------------
# persistent table cache with the reference's file naming
# ----------------------------------------------------------------------------------------------------------------------
def table_dir(root_path=None):
    """Directory of the particle files, as augment() looks them up (tools/snowfall/simulation.py:324-327):
    '<root_path>/training/snowflakes/npy' or '<repo>/npy' (override: environment variable LSS_NPY_DIR)."""
    import os
    from pathlib import Path
    if root_path:
        return Path(root_path) / 'training' / 'snowflakes' / 'npy'
    return Path(os.environ.get('LSS_NPY_DIR', Path(__file__).parent.parent.parent.absolute() / 'npy'))


def table_files(prefix: str, directory, n_planes: int = 64):
    """'<dist>_<rate>_<ratio>_<line>.npy', line = 1 .. 64 (sampling.py:344, simulation.py:78)."""
    from pathlib import Path
    return [Path(directory) / f'{prefix}_{k}.npy' for k in range(1, n_planes + 1)]


def load_table_set(prefix: str, directory):
    """The 64 (x, y, r) float64 tables of one prefix; FileNotFoundError like np.load in the reference (simulation.py:329)."""
    tabs = []
    for path in table_files(prefix, directory):
        if not path.is_file():
            raise FileNotFoundError(f"[Errno 2] No such file or directory: '{path}'")
        tabs.append(np.load(str(path)))
    return tabs


def load_or_sample_table_set(mode: str, snowfall_rate: float, terminal_velocity: float, root_path=None, directory=None,
                             write: bool = False, seed: int = 1000, R_0: float = 80.0):
    """
    The table set of (mode, snowfall_rate, terminal_velocity): read from '<prefix>_<k>.npy' if all 64 files exist (the
    reference's published tables, or files written earlier), else drawn by the native sampler (plane k from
    np.random.default_rng(seed + k)) and -- with `write` -- saved under the reference's names, skipping files that
    exist (sampling.py:346-347).  Returns (tables, prefix, 'files' | 'sampled').
    """
    prefix = particle_file_prefix(mode, snowfall_rate, terminal_velocity)
    directory = table_dir(root_path) if directory is None else directory
    files = table_files(prefix, directory)
    if all(f.is_file() for f in files):
        return [np.load(str(f)) for f in files], prefix, 'files'
    if mode not in _DIST:
        raise NotImplementedError('Distribution model unknown.')
    tables = sample_table_set(mode, snowfall_rate, terminal_velocity, seed=seed, R_0=R_0)
    if write:
        files[0].parent.mkdir(parents=True, exist_ok=True)
        for f, t in zip(files, tables):
            if not f.is_file():
                np.save(str(f), t)
    return tables, prefix, 'sampled'
