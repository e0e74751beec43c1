"""CPU oracle for the Eikonal hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this package.  See oracle/eikonal_oracle.c for what it restates.
The reference itself (C++ needing Eigen + TensorFlow, driven from Julia) cannot be
built in this image, so there is no oracle/_ref.
"""
from .oracle import *  # noqa: F401,F403
