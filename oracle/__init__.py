"""CPU oracle for libbagpu -- TEST INFRASTRUCTURE ONLY (see oracle/ba_ref.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this package. PARITY UNPINNED: the reference ships no golden vectors for its BA path.
"""
