"""Diagnostic (GPU box): which Krylov method the sweep picks at the bench shape and why."""
import os, sys, time, json
os.environ["SKTT_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import workload_cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle, _local
from scikit_tt_b200._device import get_device
dev = get_device()
opc, rhsc, x0c = workload_cores(32, 64, 64)
op, rhs = TT(opc), TT(rhsc)
x0 = TT(x0c).ortho_right()
st = sle._State(op, x0, rhs)
for i in range(31, -1, -1):
    st.right(i)
for i in (0, 1, 2):
    st.left(i)
    L, R, A = st.Lop[i], st.Rop[i], st.A[i]
    opl = dev.local_op(L, A, R)
    shape = (L.shape[0], A.shape[2], R.shape[0])
    g = torch.Generator(device=dev.device).manual_seed(1234)
    u = torch.randn(shape, dtype=torch.float64, device=dev.device, generator=g)
    v = torch.randn(shape, dtype=torch.float64, device=dev.device, generator=g)
    Mu, Mv = dev.local_matvec(opl, u), dev.local_matvec(opl, v)
    a, b = dev.dotc(u, Mv), dev.dotc(Mu, v)
    print(i, shape, "uMv", a, "Muv", b, "rel", abs(a - b) / (dev.nrm2(u) * dev.nrm2(Mv)), flush=True)
    u_, _ = sle._micro_als(st, i, 'solve')
    q = dev.qr(u_.reshape(shape[0] * shape[1], shape[2]))
    st.x[i] = q.reshape(shape[0], shape[1], q.shape[1])
print(st.cache)
