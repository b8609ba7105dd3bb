"""Phase breakdown of one bench step (GPU box): synchronised wall time per sweep phase + Krylov iteration counts."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import workload_cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle, _local
from scikit_tt_b200._device import get_device
dev = get_device()
r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
opc, rhsc, x0c = workload_cores(32, 64, r)
op, rhs = TT(opc), TT(rhsc)
x0 = TT(x0c).ortho_right()
st = sle._State(op, x0, rhs)
x0_dev = list(st.x)
for _ in range(2):
    st.reset(x0_dev); sle._run_als(st, 1, 'solve')
torch.cuda.synchronize()
t = time.perf_counter(); st.reset(x0_dev); l0 = dev.launches(); sle._run_als(st, 1, 'solve'); torch.cuda.synchronize()
print("unprofiled step", time.perf_counter() - t, "launches", dev.launches() - l0)
_local.PROFILE = {}
t = time.perf_counter(); st.reset(x0_dev); sle._run_als(st, 1, 'solve'); torch.cuda.synchronize()
print("profiled step", time.perf_counter() - t)
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in _local.PROFILE.items()}))
