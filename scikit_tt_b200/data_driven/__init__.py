"""Data-driven methods that re-use the ALS sweep skeleton."""
from . import regression  # noqa: F401
