"""CPU oracles — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
anything from this package.  The product (gs_dynamics_b200) never does.
"""
