"""Shared helpers of the test-suite: golden fixture loading and gauge-invariant TT metrics."""
import os

import numpy as np

from oracle import tt as ott

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def cores(z, prefix):
    return [z[f"{prefix}/{i}"] for i in range(int(z[prefix + "/n"]))]


def rel_diff(a, b):
    """|| a - b || / || b || for two tensor trains given as core lists (gauge invariant)."""
    a = [np.asarray(c) for c in a]
    b = [np.asarray(c) for c in b]
    return ott.norm(ott.sub(a, b)) / ott.norm(b)


def rel_diff_up_to_phase(a, b):
    """min over unit phases of || a e^{i phi} - b || / || b || (eigenvectors are defined up to phase)."""
    a = [np.asarray(c) for c in a]
    b = [np.asarray(c) for c in b]
    # <b, a> through the transfer matrices, then align the phase of a with b and subtract in TT form
    t = np.ones((1, 1))
    for ca, cb in zip(a, b):
        t = np.einsum('pq,pmnr,qmns->rs', t, np.conj(cb), ca)
    ip = complex(t.reshape(()))
    phase = ip / abs(ip) if abs(ip) > 0 else 1.0
    if abs(phase.imag) < 1e-300:
        phase = phase.real
    return rel_diff(ott.scale(a, 1.0 / phase), b)


def cascade_operator(z):
    d = int(z["d"])
    return [z["op/first"]] + [z["op/mid"].copy() for _ in range(d - 2)] + [z["op/last"]]
