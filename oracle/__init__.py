"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithms for the fusion hot path, plus a
loader for the reference's own compiled code (oracle/_ref, built from
/root/reference where that exists).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs may import this package; the
product (ffb6d_b200/) never does.
"""
