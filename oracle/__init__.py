"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the gate-application path of the reference (`qc_shor.c`):

* ``Restatement``  -- ctypes view of ``_build/libqcsoracle.so`` (qcs_oracle.c), the
  matrix-free restatement with the reference's floating-point order;
  ``RestatementAllCores`` is the same file built with -fopenmp (row loops on all host cores).
* ``Reference``    -- ctypes view of ``_ref/libqcref.so`` (ref_bridge.c), the
  unmodified reference compiled in place against the GSL stand-in.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``quantumcomputer_b200`` / ``libqcs.so``) never does.
"""
from .bindings import (Reference, Restatement, RestatementAllCores, build, have_reference,  # noqa: F401
                       have_restatement, have_restatement_omp)
