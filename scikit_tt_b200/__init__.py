"""B200-native ALS/MALS sweep hot path of scikit_tt: same Python surface (TT, solvers.sle, solvers.evp,
solvers.ode.implicit_euler), hand-written sm_100a CUDA underneath (libsktt_b200.so, include/sktt_b200.h).
There is no CPU fallback: every solver call needs the built library and a Blackwell GPU."""
from .tensor_train import TT  # noqa: F401
from . import tensor_train  # noqa: F401
from . import solvers  # noqa: F401
