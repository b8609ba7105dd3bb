from . import sle, evp, ode  # noqa: F401
