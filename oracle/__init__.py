"""CPU oracle for the DMRG/TDVP sweep hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain NumPy/SciPy restatement of the reference's algorithm for the path named in
BASELINE.json's north_star (H_eff*C, environment update, SVD/QR bond truncation and the
Davidson / Krylov / sweep drivers that call them).  Every function cites the reference
file:line it follows.  Parity is PINNED: tests/test_oracle_golden.py checks each function here
against golden vectors produced by running the unmodified reference itself
(tests/golden/make_golden.py, fixtures in tests/golden/*.npz).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker / timed CPU baseline.  renormalizer_b200/ never
imports it.
"""
