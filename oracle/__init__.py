"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithms for the hot path (and, under oracle/_ref/,
the unmodified reference CUDA kernels compiled for sm_100a).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; nothing under decnet_b200/ does.
"""
