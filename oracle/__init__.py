"""CPU oracle for the Eikonal hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this package.  See oracle/eikonal_oracle.c for what it restates.
oracle/ref.py binds oracle/_ref/libref_eikonal.so: the reference's OWN solver sources
compiled unmodified (from /root/reference, where mounted) against a stub of the absent
third-party Eigen (oracle/eigen_stub); it pins the restatement and generates goldens.
"""
from .oracle import *  # noqa: F401,F403
