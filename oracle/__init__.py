"""CPU oracle for the replay-step hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a CPU restatement (numpy / torch-CPU fp32 and fp64) of the
algorithms on the reference's replay-step path.  It exists
only so that the CUDA path can be checked against it.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``online-continual-learning_b200/`` imports it; the product path raises when the
CUDA library is missing instead of falling back to this code.

How it is pinned: the reference ships no tests and no golden vectors (SURVEY.md
section 4), so the pin is the reference itself, executed in the build container:
``tests/golden/make_golden.py`` imports the reference modules from
``/root/reference`` (with the three optional-import stubs of SURVEY.md Appendix
B), runs them on seeded inputs and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every oracle function against those
files (`gss.npz` for oracle/gss.py).  Two things stay "parity unpinned" (SURVEY.md
section 8c): the DEFINITION of the SCR augmentation pipeline (kornia 0.4.1 is not in the
image and not vendored; oracle/augment.py restates its published formulas and is checked
against torch's grid_sample and the standard library's colorsys instead,
tests/test_oracle_augment.py) and anything that depends on the exact conv/BN kernels of
the pinned torch 1.7.1.

Every function cites the reference file:line it restates.
"""
