"""CPU oracle for the ALS/MALS sweep hot path of scikit_tt  --  TEST INFRASTRUCTURE ONLY.

This package restates, in plain NumPy/SciPy on lists of 4-D cores, the arithmetic of the reference
(PGelss/scikit_tt: scikit_tt/solvers/sle.py, evp.py, ode.py:249-330, tensor_train.py:1092-1430).
Every function cites the reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
it, and only as the checker / timed CPU baseline -- never as the shipped product path
(scikit_tt_b200 never imports oracle).

Parity pin: the oracle is checked against golden vectors produced by the live reference
(tests/golden/make_golden.py, run in the build container with PYTHONPATH=/root/reference;
numpy 2.3.5, scipy 1.18.1, OPENBLAS_NUM_THREADS=1) -- see tests/test_oracle_golden.py.
"""
