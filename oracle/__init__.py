"""CPU oracle for the pillarization hot path -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this package.  hgsfusion_b200/ never does.
"""
