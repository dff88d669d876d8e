"""CPU oracle for the MA -> FFT -> Pk/XPk path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (pylians3_b200) never does.
"""
