"""TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy) of the reference's per-timestep update loop
(erthward/geonomics v1.4.9), used as the parity checker for the CUDA path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this package.  Nothing under geonomics_b200/ imports it.
"""
