"""TEST INFRASTRUCTURE — the CPU oracle.  Not importable from the product package.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import anything from here.
"""
