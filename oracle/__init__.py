"""ORACLE package — test infrastructure only.

CPU restatements (plain PyTorch fp32/fp64) of the reference algorithms on the hot path,
used as the parity checker by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package `medmoe_b200` never imports it.
"""
