"""CPU oracle for the ray walk -- TEST INFRASTRUCTURE ONLY (see wgrt_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product package never does.
"""
