from . import sle, evp, ode, multi  # noqa: F401
