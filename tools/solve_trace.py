"""Per-micro-step trace of one bench step (GPU box): Krylov iterations, cycles and synchronised time of every call."""
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
rec = []
orig = {}
def wrap(name):
    f = getattr(dev, name)
    orig[name] = f
    def g(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter(); l0 = dev.launches()
        out = f(*a, **k)
        torch.cuda.synchronize()
        info = dict(fn=name, ms=round((time.perf_counter() - t) * 1e3, 3), launches=dev.launches() - l0)
        if name in ('krylov_solve', 'krylov_solve_refined'):
            info.update(N=a[1].numel(), iters=out[1], relres=out[2], status=out[0], cycles=(out[3] if len(out) > 3 else None))
            if name == 'krylov_solve_refined':
                import ctypes
                info['persistent_out'] = [float(v) for v in dev.scratch_peek(65536 + 4 * 256 * 8, 4, ctype=ctypes.c_double)]
        rec.append(info)
        return out
    setattr(dev, name, g)
for n in ('krylov_solve', 'krylov_solve_refined', 'local_matvec', 'prepare_local_op', 'qr', 'rq', 'nrm2', 'axpby', 'dotc'):
    wrap(n)
t = time.perf_counter(); st.reset(x0_dev); sle._run_als(st, 1, 'solve'); torch.cuda.synchronize()
print("traced step", time.perf_counter() - t)
agg = {}
for x in rec:
    a = agg.setdefault(x['fn'], dict(calls=0, ms=0.0, launches=0, iters=0))
    a['calls'] += 1; a['ms'] += x['ms']; a['launches'] += x['launches']; a['iters'] += x.get('iters', 0)
print(json.dumps(agg))
for x in rec:
    if x['fn'].startswith('krylov_solve'):
        print(json.dumps(x))
