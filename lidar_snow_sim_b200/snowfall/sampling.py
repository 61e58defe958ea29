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
