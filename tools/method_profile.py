"""GPU probe: device time of one resident C3 bench step by Device method (CUDA events around every call of the C-ABI wrappers,
no synchronisation inside the step): where the step goes besides the persistent CG."""
import os, sys, collections, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import workload_cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle
from scikit_tt_b200._device import get_device, Device
dev = get_device()
r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
opc, rhsc, x0c = workload_cores(32, 64, r)
op, rhs = TT(opc), TT(rhsc)
x0 = TT(x0c).ortho_right()
st = sle._State(op, x0, rhs)
x0_dev = list(st.x)
for _ in range(3):
    st.reset(x0_dev); sle._run_als(st, 1, 'solve')
torch.cuda.synchronize()
records = []
names = [n for n in dir(Device) if not n.startswith('_') and callable(getattr(Device, n))
         and n not in ('launches', 'sync', 'empty', 'work', 'to_device', 'close', 'set_debug', 'set_gemm_mode', 'scratch_peek', 'tiled_len')]
for n in names:
    orig = getattr(dev, n)
    def make(n, orig):
        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = orig(*a, **k); e1.record()
            records.append((n, e0, e1))
            return out
        return wrapped
    setattr(dev, n, make(n, orig))
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.reset(x0_dev)
s0.record(); sle._run_als(st, 1, 'solve'); s1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for n, e0, e1 in records:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
tot = s0.elapsed_time(s1)
print(f"step {tot:.2f} ms (with the event records in the stream)")
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:34s} {c:5d} calls {ms:8.3f} ms {100 * ms / tot:5.1f} %  {1e3 * ms / c:8.1f} us per call")
print(f"  {'(outside the wrapped calls)':34s} {'':5s}       {tot - sum(v[1] for v in agg.values()):8.3f} ms")
