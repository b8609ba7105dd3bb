"""Diagnostic (GPU box): first micro-step at which evp.als on C2 departs from the oracle, with the local spectra."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, scipy.linalg as sla
from util import load, cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import evp
from oracle import evp as oevp
z = load("c2_cooxidation20")
op, x0 = cores(z, "op"), cores(z, "x0")
rec_g, rec_o = [], []
orig = evp._local_eig
def patched(dev, M, B, k, solver, sigma):
    Mh = M.detach().cpu().numpy().copy()
    lam, vec = orig(dev, M, B, k, solver, sigma)
    rec_g.append((Mh, lam.detach().cpu().numpy().copy(), vec.detach().cpu().numpy().copy()))
    return lam, vec
evp._local_eig = patched
oorig = oevp._local_eig
def opatched(M, B, k, solver, sigma, real):
    lam, vec = oorig(M, B, k, solver, sigma, real)
    rec_o.append((np.array(M).copy(), np.array(lam).copy(), np.array(vec).copy()))
    return lam, vec
oevp._local_eig = opatched
lam, x, it = evp.als(TT(op), TT(x0), repeats=1, conv_eps=0, solver='eig')
lam_o, x_o, it_o = oevp.als(op, x0, repeats=1, conv_eps=0, solver='eig')
print("final", lam, lam_o, len(rec_g), len(rec_o))
for i, ((Mg, lg, vg), (Mo, lo, vo)) in enumerate(zip(rec_g, rec_o)):
    dM = np.linalg.norm(Mg - Mo) / np.linalg.norm(Mo) if Mg.shape == Mo.shape else -1
    ok = abs(lg[0] - lo[0]) <= 1e-8 * max(1, abs(lo[0]))
    print(i, Mg.shape, "dM", f"{dM:.2e}", "lam gpu", lg[0], "lam oracle", lo[0], "OK" if ok else "DIVERGED")
    if not ok or dM > 1e-8:
        w = sla.eigvals(Mo)
        idx = np.argsort(np.abs(w - 1))
        print("  oracle M: closest eigenvalues to 1:", w[idx[:6]])
        wg = sla.eigvals(Mg)
        idg = np.argsort(np.abs(wg - 1))
        print("  gpu    M: closest eigenvalues to 1:", wg[idg[:6]])
        print("  |lam - 1|: gpu", abs(lg[0] - 1), "oracle", abs(lo[0] - 1))
        # eigenvector agreement up to phase for the oracle's choice
        break
