"""CPU oracle for sopht_b200 — TEST INFRASTRUCTURE ONLY.

A numpy / scipy.fft (and, for the CPU baseline, plain C + OpenMP) restatement of the reference's
algorithms for the Eulerian flow step. Nothing under sopht_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs do, and only as the
checker or the reported CPU baseline.

Parity status: stencils, Poisson and simulator steps are pinned against the numpy references and
seeded inputs inside the reference's own test-suite, the IB communicator against the reference's
numba implementation executed in the build container (tests/golden/make_golden.py ->
tests/golden/*.npz, checked by tests/test_oracle_golden.py). Periodic Poisson has no reference
counterpart: parity unpinned.
"""
