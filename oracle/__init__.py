"""CPU oracle for the sigma-vector path -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's algorithm for the Davidson sigma build sigma = A.X of the
spin-adapted TDA family (X-TDA, SF-TDA, XSF-TDA) in Quantum-Chemistry-Group-BNU/XTDDFT, plus the
explicit-matrix definitions of A the reference uses for its own dense-vs-iterative checks.  Every
function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this package, and only as the checker or the timed CPU baseline.  The product
(`xtddft_b200`) never imports it; the product path fails loudly when the CUDA library is missing.

Pinning status: the reference ships no golden sigma vectors and its arithmetic lives
in PySCF (not installed here, not vendored).  The oracle is pinned three ways:
  1. `tests/golden/*.npz` were produced by executing the reference's OWN `vind` closures
     (xtddft/XTDA.py, xtddft/SF_TDA.py, xtddft/XSF_TDA.py, xtddft/XSF_TDA_GPU.py) in this container on seeded synthetic
     inputs, with only the absent third-party PySCF calls (get_jk, nr_uks_fxc, block_loop ...)
     supplied by a stub that restates their published semantics (tests/golden/make_golden.py).
  2. the reference's pure-NumPy helpers (`order_pyscf2my`, `so2st`, `st2so`, `get_vect`,
     `deal_v_davidson`) were executed verbatim for fixtures.
  3. explicit-A == sigma-build identities (the reference's own validation strategy).
End-to-end excitation energies of real molecules need PySCF integrals and are
"parity unpinned" here.
"""
