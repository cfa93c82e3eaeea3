"""CPU oracle for the RepCONC constrained-PQ hot path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under repconc_b200/
may import this package.
"""
