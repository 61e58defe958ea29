from .simulation import LISA  # noqa: F401
