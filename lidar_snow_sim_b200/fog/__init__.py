from .simulation import ParameterSet, simulate_fog, load_integral_table, get_available_alphas  # noqa: F401
